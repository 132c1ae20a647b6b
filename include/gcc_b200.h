/* gcc_b200 C-ABI: B200 (sm_100a) kernels for the GCC cooperative-compression training step.
 *
 * The reference (SJLeo/GCC) is pure Python on torch; it has no FFI of its own.  The boundary that
 * these entry points replace is the set of ATen calls reached from Pix2PixModel.optimize_parameters
 * / optimizer_netD_arch (models/Pix2Pix.py:565-593).  Each declaration cites the reference call
 * site whose arithmetic it implements.  See INTEGRATION.md for the ctypes binding.
 *
 * Conventions: plain pointers are DEVICE pointers unless stated; activations are NHWC bf16 with a
 * physical channel count that is a multiple of 8 (extra channels are zero); `stream` is a
 * cudaStream_t passed as void*; every function returns 0 on success, non-zero on failure
 * (gcc_last_error() describes it).  No function allocates or keeps caller memory.
 */
#ifndef GCC_B200_H
#define GCC_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- runtime ---- */
int gcc_abi_version(void);
const char* gcc_last_error(void);
int gcc_check_device(void); /* non-zero unless the current device is sm_100 */

/* ---- dense contractions on tcgen05 tensor cores (conv_gemm.cu) ----
 * gcc_conv_gemm_bf16: y[n,oy,ox,y_coff+r] = act(bias[r] + sum_{kh,kw,c} x[n,iy,ix,c] * w[r][kh*KW+kw][c])
 *   transposed=0: iy = stride*oy + kh - pad             nn.Conv2d          (models/Pix2Pix.py:31,216,227,280-300)
 *   transposed=1: oy = stride*iy + kh - pad             nn.ConvTranspose2d (models/Pix2Pix.py:40-56,253-256)
 *   (also used for the data gradients of both, with the transposed weight pack)
 *   x: [N,H,W,Cx] bf16, w: [R][T=KH*KW][Cw] bf16, y: [N,OH,OW,Cy] bf16, bias: [R] fp32 or NULL.
 *   act: 0 none, 1 leaky-relu(slope), 2 tanh.  stride in {1,2}.
 */
int gcc_conv_gemm_bf16(const void* x, int N, int H, int W, int Cx, const void* w, int R, int T, int Cw,
                       const float* bias, void* y, int OH, int OW, int Cy, int y_coff, int transposed, int KH,
                       int KW, int stride, int pad, int act, float slope, void* stream);
/* gcc_wgrad_gemm_bf16: dw[b][r][kh*KW+kw][c] (+)= scale * sum_{n,oy,ox} p[n,oy,ox,r] * q[n,stride*oy+kh-pad,stride*ox+kw-pad,c]
 *   weight gradient of Conv2d (p = dy, q = x) and ConvTranspose2d (p = x, q = dy); with batched=1,
 *   KH=KW=1, p == q it is the per-sample Gram matrix f f^T (models/Pix2Pix.py:733-740).
 *   p: [N,OH,OW,Cp] bf16, q: [N,H,W,Cq] bf16, dw: fp32 [N if batched][R][KH*KW][C]. */
int gcc_wgrad_gemm_bf16(const void* p, int N, int OH, int OW, int Cp, const void* q, int H, int W, int Cq,
                        float* dw, int R, int C, int KH, int KW, int stride, int pad, int batched, int accumulate,
                        float scale, void* stream);
/* CUDA-core cross-checks with identical signatures (tests only). */
int gcc_conv_direct_bf16(const void* x, int N, int H, int W, int Cx, const void* w, int R, int T, int Cw,
                         const float* bias, void* y, int OH, int OW, int Cy, int y_coff, int transposed, int KH,
                         int KW, int stride, int pad, int act, float slope, void* stream);
int gcc_wgrad_direct_bf16(const void* p, int N, int OH, int OW, int Cp, const void* q, int H, int W, int Cq,
                          float* dw, int R, int C, int KH, int KW, int stride, int pad, int batched, int accumulate,
                          float scale, void* stream);
void gcc_debug_force_block_n(int bn);

#ifdef __cplusplus
}
#endif
#endif /* GCC_B200_H */
