// tw_unity.cu -- single CUDA translation unit (the __constant__/__device__ tables in
// tw_device.cuh must have exactly one instance per device image).
#include "tw_lmat.cu"
#include "tw_ops.cu"
#include "tw_blocks.cu"
#ifdef TW_TEST_HOOKS
#include "tw_probe.cu"
#endif
#include "tw_capi.cu"
#include "tw_shard.cu"
#include "tw_solve.cu"
