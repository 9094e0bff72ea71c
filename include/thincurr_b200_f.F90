!---------------------------------------------------------------------------------
! thincurr_b200_f.F90 -- ISO_C_BINDING interface to libthincurr_b200.so
!
! What a Fortran host (src/physics/thin_wall.F90 of the Open FUSION Toolkit) USEs to hand
! the dense operator builds to the B200 backend.  Every procedure is a plain C function
! declared in include/thincurr_b200.h; scalars are passed by VALUE, arrays by reference
! (assumed-size, contiguous, column-major exactly as tw_type stores them).
! All functions return 0 on success; thincurr_b200_last_error() returns the message.
!
! Shipped as source: this image has no Fortran compiler, the C ABI is exercised from
! C++/ctypes by the test-suite (tests/test_host_cpu.py, tests/test_gpu_lmat.py).
!---------------------------------------------------------------------------------
MODULE thincurr_b200
USE, INTRINSIC :: iso_c_binding, ONLY: c_int, c_int64_t, c_double, c_ptr, c_char, c_null_ptr
IMPLICIT NONE
INTERFACE
  !> Number of CUDA devices the backend will use
  FUNCTION thincurr_b200_device_count() BIND(C,NAME="thincurr_b200_device_count") RESULT(n)
    IMPORT :: c_int
    INTEGER(c_int) :: n
  END FUNCTION thincurr_b200_device_count
  !> Message of the last failed call (NUL-terminated C string)
  FUNCTION thincurr_b200_last_error() BIND(C,NAME="thincurr_b200_last_error") RESULT(msg)
    IMPORT :: c_ptr
    TYPE(c_ptr) :: msg
  END FUNCTION thincurr_b200_last_error
  !> Model handle from the arrays of an initialized tw_type (after tw_setup)
  FUNCTION thincurr_b200_model_from_tw(np,r,nc,lc,reg,pmap,np_active,nholes,kfh,lfh,ca,qbasis,tw_ptr) &
    BIND(C,NAME="thincurr_b200_model_from_tw") RESULT(ierr)
    IMPORT :: c_int, c_double, c_ptr
    INTEGER(c_int), VALUE, INTENT(in) :: np,nc,np_active,nholes
    REAL(c_double), INTENT(in) :: r(3,*)        !< mesh%r
    INTEGER(c_int), INTENT(in) :: lc(3,*)       !< mesh%lc (1-based, after orientation sync)
    INTEGER(c_int), INTENT(in) :: reg(*)        !< mesh%reg
    INTEGER(c_int), INTENT(in) :: pmap(*)       !< self%pmap
    INTEGER(c_int), INTENT(in) :: kfh(*)        !< self%kfh(nc+1)
    INTEGER(c_int), INTENT(in) :: lfh(2,*)      !< self%lfh(2,nfh)
    REAL(c_double), INTENT(in) :: ca(*)         !< mesh%ca
    REAL(c_double), INTENT(in) :: qbasis(3,3,*) !< self%qbasis
    TYPE(c_ptr), INTENT(out) :: tw_ptr
    INTEGER(c_int) :: ierr
  END FUNCTION thincurr_b200_model_from_tw
  SUBROUTINE thincurr_b200_destroy(tw_ptr) BIND(C,NAME="thincurr_b200_destroy")
    IMPORT :: c_ptr
    TYPE(c_ptr), VALUE :: tw_ptr
  END SUBROUTINE thincurr_b200_destroy
  !> Replaces the loop nest of tw_compute_LmatDirect: full Lmat(nelems,nelems) into host memory
  FUNCTION thincurr_b200_Lmat_host(tw_ptr,Lmat) BIND(C,NAME="thincurr_b200_Lmat_host") RESULT(ierr)
    IMPORT :: c_int, c_double, c_ptr
    TYPE(c_ptr), VALUE :: tw_ptr
    REAL(c_double), INTENT(out) :: Lmat(*)
    INTEGER(c_int) :: ierr
  END FUNCTION thincurr_b200_Lmat_host
  !> Number of rows of shard `shard` (0-based) out of `nshards`
  FUNCTION thincurr_b200_plan(tw_ptr,nshards,shard,nrows) BIND(C,NAME="thincurr_b200_plan") RESULT(ierr)
    IMPORT :: c_int, c_ptr
    TYPE(c_ptr), VALUE :: tw_ptr
    INTEGER(c_int), VALUE, INTENT(in) :: nshards,shard
    INTEGER(c_int), INTENT(out) :: nrows
    INTEGER(c_int) :: ierr
  END FUNCTION thincurr_b200_plan
  !> 0-based reference DOF ids of the rows of a shard
  FUNCTION thincurr_b200_shard_rows(tw_ptr,nshards,shard,row_ids) BIND(C,NAME="thincurr_b200_shard_rows") RESULT(ierr)
    IMPORT :: c_int, c_ptr
    TYPE(c_ptr), VALUE :: tw_ptr
    INTEGER(c_int), VALUE, INTENT(in) :: nshards,shard
    INTEGER(c_int), INTENT(out) :: row_ids(*)
    INTEGER(c_int) :: ierr
  END FUNCTION thincurr_b200_shard_rows
  !> Symmetric multi-device partition: row count and/or 0-based reference DOF ids of a shard (either may be c_null_ptr)
  FUNCTION thincurr_b200_shard_rows_sym(tw_ptr,nshards,shard,nrows,row_ids) BIND(C,NAME="thincurr_b200_shard_rows_sym") RESULT(ierr)
    IMPORT :: c_int, c_ptr
    TYPE(c_ptr), VALUE :: tw_ptr
    INTEGER(c_int), VALUE, INTENT(in) :: nshards,shard
    TYPE(c_ptr), VALUE :: nrows   !< c_null_ptr or C_LOC of an INTEGER(c_int)
    TYPE(c_ptr), VALUE :: row_ids !< c_null_ptr or C_LOC of INTEGER(c_int) :: row_ids(nrows)
    INTEGER(c_int) :: ierr
  END FUNCTION thincurr_b200_shard_rows_sym
  !> Upper-trapezoid rows of a shard into DEVICE memory d_out(ld,nrows): columns of the DOFs of earlier shards stay zero
  !> and are the transposes of blocks those shards computed (exchange once after the assembly)
  FUNCTION thincurr_b200_Lmat_shard_sym(tw_ptr,nshards,shard,d_out,ld,stream,stats) &
    BIND(C,NAME="thincurr_b200_Lmat_shard_sym") RESULT(ierr)
    IMPORT :: c_int, c_int64_t, c_ptr
    TYPE(c_ptr), VALUE :: tw_ptr
    INTEGER(c_int), VALUE, INTENT(in) :: nshards,shard
    TYPE(c_ptr), VALUE :: d_out   !< device pointer
    INTEGER(c_int64_t), VALUE, INTENT(in) :: ld
    TYPE(c_ptr), VALUE :: stream  !< cudaStream_t or c_null_ptr
    TYPE(c_ptr), VALUE :: stats   !< c_null_ptr or INTEGER(c_int64_t) :: stats(8)
    INTEGER(c_int) :: ierr
  END FUNCTION thincurr_b200_Lmat_shard_sym
  !> Dense block d_out(j,i) = Lmat(col_ids(j)+1,row_ids(i)+1) into DEVICE memory d_out(ld,nrows), ld >= ncols
  FUNCTION thincurr_b200_Lmat_block(tw_ptr,nrows,row_ids,ncols,col_ids,d_out,ld,stream) &
    BIND(C,NAME="thincurr_b200_Lmat_block") RESULT(ierr)
    IMPORT :: c_int, c_int64_t, c_ptr
    TYPE(c_ptr), VALUE :: tw_ptr
    INTEGER(c_int), VALUE, INTENT(in) :: nrows,ncols
    INTEGER(c_int), INTENT(in) :: row_ids(*),col_ids(*)
    TYPE(c_ptr), VALUE :: d_out   !< device pointer
    INTEGER(c_int64_t), VALUE, INTENT(in) :: ld
    TYPE(c_ptr), VALUE :: stream  !< cudaStream_t or c_null_ptr
    INTEGER(c_int) :: ierr
  END FUNCTION thincurr_b200_Lmat_block
  !> Rows of one shard into host memory h_out(ld,nrows) (row r = Lmat(:,row_ids(r)+1))
  FUNCTION thincurr_b200_Lmat_shard_host(tw_ptr,nshards,shard,h_out,ld,stats) &
    BIND(C,NAME="thincurr_b200_Lmat_shard_host") RESULT(ierr)
    IMPORT :: c_int, c_int64_t, c_double, c_ptr
    TYPE(c_ptr), VALUE :: tw_ptr
    INTEGER(c_int), VALUE, INTENT(in) :: nshards,shard
    REAL(c_double), INTENT(out) :: h_out(*)
    INTEGER(c_int64_t), VALUE, INTENT(in) :: ld
    TYPE(c_ptr), VALUE :: stats !< c_null_ptr or INTEGER(c_int64_t) :: stats(8)
    INTEGER(c_int) :: ierr
  END FUNCTION thincurr_b200_Lmat_shard_host
  !> Coil sets from memory (kind: 0 = Vcoils, 1 = Icoils), see include/thincurr_b200.h
  FUNCTION thincurr_b200_set_coils(tw_ptr,kind,nsets,set_ptr,fil_ptr,pts,scales,radius,res_per_len,sens_mask,sizes) &
    BIND(C,NAME="thincurr_b200_set_coils") RESULT(ierr)
    IMPORT :: c_int, c_double, c_ptr
    TYPE(c_ptr), VALUE :: tw_ptr
    INTEGER(c_int), VALUE, INTENT(in) :: kind,nsets
    INTEGER(c_int), INTENT(in) :: set_ptr(*),fil_ptr(*),sens_mask(*)
    REAL(c_double), INTENT(in) :: pts(3,*),scales(*),radius(*),res_per_len(*)
    INTEGER(c_int), INTENT(out) :: sizes(9)
    INTEGER(c_int) :: ierr
  END FUNCTION thincurr_b200_set_coils
END INTERFACE
CONTAINS
!---------------------------------------------------------------------------------
!> Example host-side replacement of the hot loop of tw_compute_LmatDirect
!! (thin_wall.F90:887-1186) for the self-inductance case without V-coils: the caller keeps
!! allocation, caching, timing print and the V-coil block fill exactly as in the reference.
!---------------------------------------------------------------------------------
SUBROUTINE tw_lmat_b200(np,r,nc,lc,reg,pmap,np_active,nholes,kfh,lfh,ca,qbasis,nelems,Lmat,ierr)
INTEGER(c_int), INTENT(in) :: np,nc,np_active,nholes,nelems
REAL(c_double), INTENT(in) :: r(3,np),ca(nc),qbasis(3,3,nc)
INTEGER(c_int), INTENT(in) :: lc(3,nc),reg(nc),pmap(np),kfh(nc+1),lfh(2,*)
REAL(c_double), INTENT(out) :: Lmat(nelems,nelems)
INTEGER(c_int), INTENT(out) :: ierr
TYPE(c_ptr) :: tw_ptr
ierr=thincurr_b200_model_from_tw(np,r,nc,lc,reg,pmap,np_active,nholes,kfh,lfh,ca,qbasis,tw_ptr)
IF(ierr/=0)RETURN
ierr=thincurr_b200_Lmat_host(tw_ptr,Lmat)
CALL thincurr_b200_destroy(tw_ptr)
END SUBROUTINE tw_lmat_b200
END MODULE thincurr_b200
