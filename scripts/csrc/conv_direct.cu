// CUDA-core direct convolution over the same NHWC bf16 / packed-weight layouts as conv_gemm.cu.
// Same argument meaning as gcc_conv_gemm_bf16 / gcc_wgrad_gemm_bf16.  These are the on-device
// cross-check for the tcgen05 kernels (tests/) and are not used on the training path.
#include "common.cuh"

namespace gcc {

__global__ void conv_direct_kernel(const bf16* __restrict__ x, int N, int H, int W, int Cx,
                                   const bf16* __restrict__ w, int R, int T, int Cw,
                                   const float* __restrict__ bias, bf16* __restrict__ y, int OH, int OW, int Cy,
                                   int y_coff, int transposed, int KH, int KW, int stride, int pad, int act,
                                   float slope, int Rp) {
  const long long total = (long long)N * OH * OW * Rp;
  const int Ck = Cx < Cw ? Cx : Cw;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx % Rp);
    long long t = idx / Rp;
    const int ox = (int)(t % OW);
    t /= OW;
    const int oy = (int)(t % OH);
    const int n = (int)(t / OH);
    float acc = 0.f;
    if (r < R) {
      for (int kh = 0; kh < KH; ++kh) {
        int iy;
        if (!transposed) iy = oy * stride + kh - pad;
        else {
          const int ty = oy + pad - kh;
          if (ty < 0 || (ty % stride)) continue;
          iy = ty / stride;
        }
        if (iy < 0 || iy >= H) continue;
        for (int kw = 0; kw < KW; ++kw) {
          int ix;
          if (!transposed) ix = ox * stride + kw - pad;
          else {
            const int tx = ox + pad - kw;
            if (tx < 0 || (tx % stride)) continue;
            ix = tx / stride;
          }
          if (ix < 0 || ix >= W) continue;
          const bf16* xp = x + (((long long)n * H + iy) * W + ix) * Cx;
          const bf16* wp = w + ((long long)r * T + kh * KW + kw) * Cw;
          for (int c = 0; c < Ck; ++c) acc += __bfloat162float(xp[c]) * __bfloat162float(wp[c]);
        }
      }
      if (bias) acc += bias[r];
      if (act == 1) acc = acc > 0.f ? acc : acc * slope;
      else if (act == 2) acc = tanhf(acc);
    }
    y[(((long long)n * OH + oy) * OW + ox) * Cy + y_coff + r] = __float2bfloat16(acc);
  }
}

__global__ void wgrad_direct_kernel(const bf16* __restrict__ pm, int N, int OH, int OW, int Cp,
                                    const bf16* __restrict__ qm, int H, int W, int Cq, float* __restrict__ dw, int R,
                                    int C, int KH, int KW, int stride, int pad, int batched, int accumulate,
                                    float scale) {
  const int T = KH * KW;
  const long long per = (long long)R * T * C;
  const long long total = per * (batched ? N : 1);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    long long t = idx / C;
    const int tap = (int)(t % T);
    t /= T;
    const int r = (int)(t % R);
    const int img = (int)(t / R);
    const int kh = tap / KW, kw = tap % KW;
    float acc = 0.f;
    const int nb = batched ? img : 0, ne = batched ? img + 1 : N;
    for (int n = nb; n < ne; ++n)
      for (int oy = 0; oy < OH; ++oy) {
        const int iy = oy * stride + kh - pad;
        if (iy < 0 || iy >= H) continue;
        for (int ox = 0; ox < OW; ++ox) {
          const int ix = ox * stride + kw - pad;
          if (ix < 0 || ix >= W) continue;
          acc += __bfloat162float(pm[(((long long)n * OH + oy) * OW + ox) * Cp + r]) *
                 __bfloat162float(qm[(((long long)n * H + iy) * W + ix) * Cq + c]);
        }
      }
    acc *= scale;
    if (accumulate) dw[idx] += acc;
    else dw[idx] = acc;
  }
}

}  // namespace gcc

extern "C" int gcc_conv_direct_bf16(const void* x, int N, int H, int W, int Cx, const void* w, int R, int T, int Cw,
                                    const float* bias, void* y, int OH, int OW, int Cy, int y_coff, int transposed,
                                    int KH, int KW, int stride, int pad, int act, float slope, int w_per_image,
                                    float* splitk_ws, long long ws_elems, float* stats, int stats_ld, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (w_per_image) { gcc_set_error(__FILE__, __LINE__, "direct conv: per-image weights unsupported"); return GCC_ERR_ARG; }
  const int Rp = (R + 7) / 8 * 8;
  const long long total = (long long)N * OH * OW * Rp;
  const int blocks = (int)((total + 255) / 256 > 148 * 32 ? 148 * 32 : (total + 255) / 256);
  gcc::conv_direct_kernel<<<blocks, 256, 0, st>>>((const bf16*)x, N, H, W, Cx, (const bf16*)w, R, T, Cw, bias,
                                                  (bf16*)y, OH, OW, Cy, y_coff, transposed, KH, KW, stride, pad, act,
                                                  slope, Rp);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}

extern "C" int gcc_wgrad_direct_bf16(const void* pmat, int N, int OH, int OW, int Cp, const void* qmat, int H, int W,
                                     int Cq, float* dw, int R, int C, int KH, int KW, int stride, int pad,
                                     int batched, int accumulate, float scale, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long total = (long long)R * KH * KW * C * (batched ? N : 1);
  const int blocks = (int)((total + 127) / 128 > 148 * 64 ? 148 * 64 : (total + 127) / 128);
  gcc::wgrad_direct_kernel<<<blocks, 128, 0, st>>>((const bf16*)pmat, N, OH, OW, Cp, (const bf16*)qmat, H, W, Cq, dw,
                                                   R, C, KH, KW, stride, pad, batched, accumulate, scale);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
