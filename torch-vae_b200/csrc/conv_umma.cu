// tcgen05 / TMEM implicit-GEMM convolution kernels (bf16).  Placeholder until
// the UMMA path lands: reports "not handled" so the CUDA-core kernels run.
#include "common.cuh"

namespace vae {
int umma_conv_fprop(const void*, const void*, const float*, void*, double*, int, int, int, int, int,
                    int, cudaStream_t, bool* handled) { *handled = false; return VAE_OK; }
int umma_conv_dgrad(const void*, const void*, const float*, void*, double*, int, int, int, int, int,
                    int, cudaStream_t, bool* handled) { *handled = false; return VAE_OK; }
int umma_conv_wgrad(const void*, const void*, float*, int, int, int, int, int, int, cudaStream_t,
                    bool* handled) { *handled = false; return VAE_OK; }
}  // namespace vae

extern "C" int vae_has_umma(void) { return 0; }
