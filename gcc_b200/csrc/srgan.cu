// Kernels the SRGAN step needs on top of the pix2pix set (reference: models/SRGAN.py): nn.PReLU (single
// learnable slope, :49-50,:89), nn.PixelShuffle(2) (:88), nn.MaxPool2d(2, 2) of the truncated VGG19
// (models/GANLoss.py:95-144), AdaptiveAvgPool2d((1,1)) + Linear(C, 1) discriminator head (:231-245) and
// convert_image('[-1, 1]' -> 'imagenet-norm') (data/sr_dataset.py:15-64).  All HBM-bound, NHWC bf16, 16-byte vectors.
#include "common.cuh"

namespace gcc {

__device__ __forceinline__ void unpack8f(const uint4 u, float* f) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8f(const float* f) {
  return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

// y = x > 0 ? x : a * x, a = *slope (device memory: it is a learnable parameter)
__global__ void prelu_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, long long nvec,
                                 const float* __restrict__ slope) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const float a = *slope;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    float f[8];
    unpack8f(x[i], f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = f[k] > 0.f ? f[k] : a * f[k];
    y[i] = pack8f(f);
  }
}
// dx = dy * (x > 0 ? 1 : a);  dslope += sum_{x <= 0} dy * x   (torch's prelu_backward: weight grad where x <= 0)
__global__ void prelu_bwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy, uint4* __restrict__ dx,
                                 long long nvec, const float* __restrict__ slope, float* __restrict__ dslope) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const float a = *slope;
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    float f[8], d[8];
    unpack8f(x[i], f);
    unpack8f(dy[i], d);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (f[k] > 0.f) {
      } else {
        acc += d[k] * f[k];
        d[k] *= a;
      }
    }
    if (dx != nullptr) dx[i] = pack8f(d);
  }
  if (dslope != nullptr) {
    acc = warp_sum(acc);
    __shared__ float part[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) part[w] = acc;
    __syncthreads();
    if (w == 0) {
      float v = lane < (blockDim.x >> 5) ? part[lane] : 0.f;
      v = warp_sum(v);
      if (lane == 0) atomicAdd(dslope, v);
    }
  }
}

// PixelShuffle(2) on NHWC: out[n, 2h+i, 2w+j, c] = in[n, h, w, c*4 + i*2 + j]   (C = logical output channels,
// in has Cin_p >= 4C physical channels, out Cout_p >= C; pad channels of out are written as zero).
// inverse = 1: in[n, h, w, c*4 + i*2 + j] = out[...] (the backward pass), pad channels of `in` zeroed.
__global__ void pixel_shuffle2_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, int N, int H, int W, int C,
                                      int Cin_p, int Cout_p, int inverse) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  if (!inverse) {
    const long long total = (long long)N * 2 * H * 2 * W * Cout_p;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
      const int c = (int)(idx % Cout_p);
      long long t = idx / Cout_p;
      const int ox = (int)(t % (2 * W));
      t /= 2 * W;
      const int oy = (int)(t % (2 * H));
      const long long n = t / (2 * H);
      bf16 v = __float2bfloat16(0.f);
      if (c < C) v = src[((n * H + (oy >> 1)) * W + (ox >> 1)) * Cin_p + c * 4 + (oy & 1) * 2 + (ox & 1)];
      dst[idx] = v;
    }
  } else {
    const long long total = (long long)N * H * W * Cin_p;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
      const int ci = (int)(idx % Cin_p);
      long long t = idx / Cin_p;
      const int x = (int)(t % W);
      t /= W;
      const int y = (int)(t % H);
      const long long n = t / H;
      bf16 v = __float2bfloat16(0.f);
      if (ci < 4 * C) {
        const int c = ci >> 2, i = (ci >> 1) & 1, j = ci & 1;
        v = src[((n * 2 * H + 2 * y + i) * (2 * W) + 2 * x + j) * Cout_p + c];
      }
      dst[idx] = v;
    }
  }
}

// MaxPool2d(kernel 2, stride 2) on NHWC, H and W even.
__global__ void maxpool2_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int N, int H, int W, int G) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const int OH = H / 2, OW = W / 2;
  const long long total = (long long)N * OH * OW * G;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % G);
    long long t = idx / G;
    const int ox = (int)(t % OW);
    t /= OW;
    const int oy = (int)(t % OH);
    const long long n = t / OH;
    const long long base = ((n * H + 2 * oy) * W + 2 * ox) * G + g;
    float a[8], b[8];
    unpack8f(x[base], a);
    unpack8f(x[base + G], b);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fmaxf(a[k], b[k]);
    unpack8f(x[base + (long long)W * G], b);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fmaxf(a[k], b[k]);
    unpack8f(x[base + (long long)W * G + G], b);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fmaxf(a[k], b[k]);
    y[idx] = pack8f(a);
  }
}
// dx[window] = dy at the FIRST position (row-major scan, as ATen's max_pool2d) that holds the window maximum.
__global__ void maxpool2_bwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy, uint4* __restrict__ dx,
                                    int N, int H, int W, int G) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const int OH = H / 2, OW = W / 2;
  const long long total = (long long)N * OH * OW * G;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % G);
    long long t = idx / G;
    const int ox = (int)(t % OW);
    t /= OW;
    const int oy = (int)(t % OH);
    const long long n = t / OH;
    const long long base = ((n * H + 2 * oy) * W + 2 * ox) * G + g;
    const long long off[4] = {0, G, (long long)W * G, (long long)W * G + G};
    float v[4][8], d[8], o[4][8];
#pragma unroll
    for (int p = 0; p < 4; ++p) unpack8f(x[base + off[p]], v[p]);
    unpack8f(dy[idx], d);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int best = 0;
      float m = v[0][k];
#pragma unroll
      for (int p = 1; p < 4; ++p)
        if (v[p][k] > m) { m = v[p][k]; best = p; }
#pragma unroll
      for (int p = 0; p < 4; ++p) o[p][k] = (p == best) ? d[k] : 0.f;
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) dx[base + off[p]] = pack8f(o[p]);
  }
}

// y[., c] = x[., c] * scale[c] + shift[c] for c < C, 0 for pad channels (C <= 8 = one vector per pixel)
__global__ void channel_affine8_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, long long npix, int C,
                                       const float* __restrict__ scale, const float* __restrict__ shift) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sc[k] = k < C ? scale[k] : 0.f;
    sh[k] = (k < C && shift != nullptr) ? shift[k] : 0.f;
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix;
       i += (long long)gridDim.x * blockDim.x) {
    float f[8];
    unpack8f(x[i], f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = f[k] * sc[k] + sh[k];
    y[i] = pack8f(f);
  }
}

// logits[n] = b + sum_c w[c] * (sums[n][c] / HW)   (sums: per-sample channel sums from gcc_norm_stats_bf16)
__global__ void pool_linear_fwd_kernel(const float* __restrict__ sums, int Cp, int C, float inv_hw,
                                       const float* __restrict__ w, const float* __restrict__ b,
                                       bf16* __restrict__ logits) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const int n = blockIdx.x;
  float acc = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) acc += w[c] * sums[(long long)n * 2 * Cp + c] * inv_hw;
  acc = warp_sum(acc);
  __shared__ float part[32];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  if (lane == 0) part[wp] = acc;
  __syncthreads();
  if (wp == 0) {
    float v = lane < (blockDim.x >> 5) ? part[lane] : 0.f;
    v = warp_sum(v);
    if (lane < 8) logits[n * 8 + lane] = __float2bfloat16(lane == 0 ? v + (b != nullptr ? b[0] : 0.f) : 0.f);
  }
}
// dx[n, pix, c] = dlogit[n] * w[c] / HW ;  dw[c] += sum_n dlogit[n] * sums[n][c] / HW ;  db += sum_n dlogit[n]
__global__ void pool_linear_bwd_dx_kernel(const bf16* __restrict__ dlogit, const float* __restrict__ w, int N,
                                          long long HW, int Cp, int C, float inv_hw, bf16* __restrict__ dx) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const long long total = (long long)N * HW * Cp;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % Cp);
    const long long n = idx / ((long long)HW * Cp);
    dx[idx] = __float2bfloat16(c < C ? __bfloat162float(dlogit[n * 8]) * w[c] * inv_hw : 0.f);
  }
}
__global__ void pool_linear_bwd_param_kernel(const bf16* __restrict__ dlogit, const float* __restrict__ sums, int N,
                                             int Cp, int C, float inv_hw, float* __restrict__ dw,
                                             float* __restrict__ db) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int n = 0; n < N; ++n) acc += __bfloat162float(dlogit[n * 8]) * sums[(long long)n * 2 * Cp + c];
    if (dw != nullptr) dw[c] += acc * inv_hw;
  }
  if (db != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    float acc = 0.f;
    for (int n = 0; n < N; ++n) acc += __bfloat162float(dlogit[n * 8]);
    db[0] += acc;
  }
}

static inline int sr_blocks(long long n) {
  long long b = (n + 255) / 256;
  if (b < 1) b = 1;
  if (b > 148 * 16) b = 148 * 16;
  return (int)b;
}

}  // namespace gcc

using namespace gcc;

extern "C" int gcc_prelu_fwd_bf16(const void* x, void* y, long long n, const float* slope_dev, void* stream) {
  if (n % 8) { gcc_set_error(__FILE__, __LINE__, "prelu: element count must be a multiple of 8"); return GCC_ERR_ARG; }
  gcc_launch(prelu_fwd_kernel, sr_blocks(n / 8), 256, 0, (cudaStream_t)stream, (const uint4*)x, (uint4*)y, n / 8, slope_dev);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_prelu_bwd_bf16(const void* x, const void* dy, void* dx, long long n, const float* slope_dev,
                                  float* dslope, void* stream) {
  if (n % 8) { gcc_set_error(__FILE__, __LINE__, "prelu: element count must be a multiple of 8"); return GCC_ERR_ARG; }
  int b = sr_blocks(n / 8);
  if (b > 148 * 4) b = 148 * 4;
  gcc_launch(prelu_bwd_kernel, b, 256, 0, (cudaStream_t)stream, (const uint4*)x, (const uint4*)dy, (uint4*)dx, n / 8,
                                                        slope_dev, dslope);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_pixel_shuffle2_bf16(const void* src, void* dst, int N, int H, int W, int C, int Cin_p, int Cout_p,
                                       int inverse, void* stream) {
  if (4 * C > Cin_p || C > Cout_p) { gcc_set_error(__FILE__, __LINE__, "pixel_shuffle: bad channel counts"); return GCC_ERR_ARG; }
  const long long total = inverse ? (long long)N * H * W * Cin_p : (long long)N * 4 * H * W * Cout_p;
  gcc_launch(pixel_shuffle2_kernel, sr_blocks(total), 256, 0, (cudaStream_t)stream, (const bf16*)src, (bf16*)dst, N, H, W, C,
                                                                           Cin_p, Cout_p, inverse);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_maxpool2_fwd_bf16(const void* x, void* y, int N, int H, int W, int Cp, void* stream) {
  if ((H % 2) || (W % 2) || (Cp % 8)) { gcc_set_error(__FILE__, __LINE__, "maxpool2: H, W even and Cp % 8 == 0"); return GCC_ERR_ARG; }
  gcc_launch(maxpool2_fwd_kernel, sr_blocks((long long)N * (H / 2) * (W / 2) * (Cp / 8)), 256, 0, (cudaStream_t)stream, 
      (const uint4*)x, (uint4*)y, N, H, W, Cp / 8);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_maxpool2_bwd_bf16(const void* x, const void* dy, void* dx, int N, int H, int W, int Cp,
                                     void* stream) {
  if ((H % 2) || (W % 2) || (Cp % 8)) { gcc_set_error(__FILE__, __LINE__, "maxpool2: H, W even and Cp % 8 == 0"); return GCC_ERR_ARG; }
  gcc_launch(maxpool2_bwd_kernel, sr_blocks((long long)N * (H / 2) * (W / 2) * (Cp / 8)), 256, 0, (cudaStream_t)stream, 
      (const uint4*)x, (const uint4*)dy, (uint4*)dx, N, H, W, Cp / 8);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_channel_affine8_bf16(const void* x, void* y, long long npix, int C, const float* scale_dev,
                                        const float* shift_dev, void* stream) {
  if (C > 8) { gcc_set_error(__FILE__, __LINE__, "channel_affine8: C must be <= 8"); return GCC_ERR_ARG; }
  gcc_launch(channel_affine8_kernel, sr_blocks(npix), 256, 0, (cudaStream_t)stream, (const uint4*)x, (uint4*)y, npix, C,
                                                                           scale_dev, shift_dev);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_pool_linear_fwd(const float* sums, int N, long long HW, int Cp, int C, const float* w,
                                   const float* b, void* logits, void* stream) {
  gcc_launch(pool_linear_fwd_kernel, N, 128, 0, (cudaStream_t)stream, sums, Cp, C, 1.f / (float)HW, w, b, (bf16*)logits);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_pool_linear_bwd(const void* dlogit, const float* sums, const float* w, int N, long long HW, int Cp,
                                   int C, void* dx, float* dw, float* db, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (dx != nullptr) {
    gcc_launch(pool_linear_bwd_dx_kernel, sr_blocks((long long)N * HW * Cp), 256, 0, st, (const bf16*)dlogit, w, N, HW, Cp, C,
                                                                                1.f / (float)HW, (bf16*)dx);
    GCC_CHECK_LAUNCH();
  }
  if (dw != nullptr || db != nullptr) {
    gcc_launch(pool_linear_bwd_param_kernel, (C + 127) / 128, 128, 0, st, (const bf16*)dlogit, sums, N, Cp, C, 1.f / (float)HW, dw, db);
    GCC_CHECK_LAUNCH();
  }
  return GCC_OK;
}
