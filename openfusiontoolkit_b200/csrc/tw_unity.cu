// tw_unity.cu -- single CUDA translation unit (the __constant__/__device__ tables in
// tw_device.cuh must have exactly one instance per device image).
#include "tw_lmat.cu"
#include "tw_ops.cu"
#include "tw_probe.cu"
#include "tw_capi.cu"
