// K2 tensor-core edge kernel (tcgen05 / TMEM / TMA, 3xTF32) -- placeholder until the kernel lands:
// every shape is reported as "not served", so K2 runs its FP32 SIMT GEMM.
#include "bc_internal.h"

int bc_k2_umma_edge(bc_model*, const uint8_t*, size_t, int, size_t, int, int, float*, int, float*, int, int, cudaStream_t) {
    return BC_ELIMIT;
}
void bc_k2_umma_free(bc_model*) {}
