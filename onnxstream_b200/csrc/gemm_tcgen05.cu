// gemm_tcgen05.cu -- placeholder until the tcgen05 path lands (next commit): reports "not eligible".
#include "common.cuh"
int osb_tc_gemm_launch(const void*, const void*, void*, const void*, const void*, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int, cudaStream_t) { return (int)cudaErrorNotSupported; }
int osb_tc_conv_launch(const void*, const void*, const void*, const void*, void*, int64_t, int64_t, int64_t, int64_t, int, int, int, int, int, int64_t, int64_t, cudaStream_t) { return (int)cudaErrorNotSupported; }
bool osb_tc_gemm_ok(int64_t, int64_t, int64_t, int, const void*, const void*, const void*, int64_t, int64_t, int64_t) { return false; }
bool osb_tc_conv_ok(int64_t, int64_t, int64_t, int64_t, int, int, int, const void*, const void*, const void*) { return false; }
