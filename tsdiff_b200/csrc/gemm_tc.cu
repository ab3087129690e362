// placeholder until the tcgen05 kernel lands: reports "unsupported" so tsd_gemm falls
// through to the FFMA kernel.
#include "gemm.cuh"
int tsd_gemm_tf32(const GemmArgs& g, cudaStream_t stream) { (void)g; (void)stream; return TSD_ERR_UNSUPPORTED; }
