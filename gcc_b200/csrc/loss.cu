// Loss reductions of the GCC step (all produce fp32 scalars on the device, no host sync):
//   GAN losses ............. models/GANLoss.py:38-59 (hinge / lsgan / vanilla / wgangp means)
//   L1 ..................... models/Pix2Pix.py:520   (criterionL1(fake_B, real_B))
//   content / gram RMSE .... models/Pix2Pix.py:542-543 (sqrt(MSE(.)))
// Inputs are NHWC bf16 activations [npix][Cp] of which the first C channels are logical, or fp32
// matrices (Gram).  Warp-shuffle + one atomicAdd per block into a pre-zeroed fp32 accumulator.
#include "common.cuh"

namespace gcc {

__device__ __forceinline__ void block_atomic_sum(float v, float* out) {
  __shared__ float wsum[32];
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) wsum[wid] = v;
  __syncthreads();
  if (wid == 0) {
    float t = lane < (blockDim.x + 31) / 32 ? wsum[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) atomicAdd(out, t);
  }
}

// per-element GAN loss term and its derivative.  mode: 0 hinge, 1 lsgan, 2 vanilla, 3 wgangp.
// kind: 0 = D on real, 1 = D on fake, 2 = G (target real, for_discriminator = False)
__device__ __forceinline__ float gan_term(float x, int mode, int kind, float* grad) {
  if (mode == 0) {
    // torch.min(v, 0) splits the gradient evenly at an exact tie (v == 0); bf16 logits do hit it
    if (kind == 0) { const float v = x - 1.f; *grad = v < 0.f ? -1.f : (v == 0.f ? -0.5f : 0.f); return -fminf(v, 0.f); }
    if (kind == 1) { const float v = -x - 1.f; *grad = v < 0.f ? 1.f : (v == 0.f ? 0.5f : 0.f); return -fminf(v, 0.f); }
    *grad = -1.f;
    return -x;
  }
  const float target = (kind == 1) ? 0.f : 1.f;
  if (mode == 1) { const float d = x - target; *grad = 2.f * d; return d * d; }
  if (mode == 2) {
    // BCE with logits: max(x,0) - x*t + log(1 + exp(-|x|))
    const float l = fmaxf(x, 0.f) - x * target + log1pf(expf(-fabsf(x)));
    *grad = 1.f / (1.f + expf(-x)) - target;
    return l;
  }
  *grad = (kind == 1) ? 1.f : -1.f;
  return (kind == 1) ? x : -x;
}

__global__ void gan_loss_fwd_kernel(const bf16* __restrict__ pred, long long npix, int Cp, int C, int mode, int kind,
                                    float inv_count, float* __restrict__ out) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  float acc = 0.f;
  const long long total = npix * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / C;
    const int c = (int)(i % C);
    float g;
    acc += gan_term(__bfloat162float(pred[p * Cp + c]), mode, kind, &g);
  }
  block_atomic_sum(acc * inv_count, out);
}
// dpred = gout * d(term)/dx / count   (pad channels get zero)
__global__ void gan_loss_bwd_kernel(const bf16* __restrict__ pred, long long npix, int Cp, int C, int mode, int kind,
                                    float inv_count, const float* __restrict__ gout, bf16* __restrict__ dpred) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const float go = *gout * inv_count;
  const long long total = npix * Cp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cp);
    float g = 0.f;
    if (c < C) gan_term(__bfloat162float(pred[i]), mode, kind, &g);
    dpred[i] = __float2bfloat16(c < C ? g * go : 0.f);
  }
}

// out[0] += sum |a-b| * inv_count   (mode 0)   or   sum (a-b)^2 * inv_count   (mode 1)
__global__ void diff_reduce_bf16_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, long long npix,
                                        int Cp, int C, int mode, float inv_count, float* __restrict__ out) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  float acc = 0.f;
  if (C == Cp && (Cp % 8) == 0) {
    const long long nvec = npix * Cp / 8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
         i += (long long)gridDim.x * blockDim.x) {
      const uint4 u = reinterpret_cast<const uint4*>(a)[i];
      const uint4 v = reinterpret_cast<const uint4*>(b)[i];
      const uint32_t uw[4] = {u.x, u.y, u.z, u.w}, vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float d0 = bf16_lo(uw[k]) - bf16_lo(vw[k]), d1 = bf16_hi(uw[k]) - bf16_hi(vw[k]);
        acc += mode ? d0 * d0 + d1 * d1 : fabsf(d0) + fabsf(d1);  // modes 1, 2: squares
      }
    }
  } else {
    const long long total = npix * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
      const long long p = i / C;
      const int c = (int)(i % C);
      const float d = __bfloat162float(a[p * Cp + c]) - __bfloat162float(b[p * Cp + c]);
      acc += mode ? d * d : fabsf(d);
    }
  }
  block_atomic_sum(acc * inv_count, out);
}
// mode 0 (L1 mean):      da = gout * sign(a-b) * inv_count
// mode 1 (RMSE = sqrt(mean sq)): da = gout * (a-b) * inv_count / rmse,  rmse = sqrt(*msq)
// mode 2 (MSE, CycleGAN.py:513-514): da = gout * 2 (a-b) * inv_count
// one thread per 8-channel vector (Cp % 8 == 0); the pad-channel test needs no division when C == Cp
__global__ void diff_bwd_bf16_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, long long npix, int Cp,
                                     int C, int mode, float inv_count, const float* __restrict__ gout,
                                     const float* __restrict__ msq, bf16* __restrict__ da) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  float go = *gout * inv_count;
  if (mode == 1) go /= fmaxf(sqrtf(*msq), 1e-20f);
  if (mode == 2) go *= 2.f;  // plain MSE
  const int G = Cp / 8;
  const long long nvec = npix * G;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (C == Cp) ? 0 : (int)(i % G) * 8;
    const uint4 u = reinterpret_cast<const uint4*>(a)[i];
    const uint4 v = reinterpret_cast<const uint4*>(b)[i];
    const uint32_t uw[4] = {u.x, u.y, u.z, u.w}, vw[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float d0 = bf16_lo(uw[k]) - bf16_lo(vw[k]), d1 = bf16_hi(uw[k]) - bf16_hi(vw[k]);
      float r0 = mode ? d0 * go : (d0 > 0.f ? go : (d0 < 0.f ? -go : 0.f));
      float r1 = mode ? d1 * go : (d1 > 0.f ? go : (d1 < 0.f ? -go : 0.f));
      if (C != Cp) {
        if (c0 + 2 * k >= C) r0 = 0.f;
        if (c0 + 2 * k + 1 >= C) r1 = 0.f;
      }
      o[k] = pack_bf16(r0, r1);
    }
    reinterpret_cast<uint4*>(da)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// fp32 matrices (Gram): out += sum (a-b)^2 * inv_count
__global__ void sqdiff_reduce_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                                         float inv_count, float* __restrict__ out) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    acc += d * d;
  }
  block_atomic_sum(acc * inv_count, out);
}
// Gram RMSE backward: m[b][i][j] = bf16( coef * ((Gs-Gt)[i][j] + (Gs-Gt)[j][i]) ),
// coef = gout * inv_count / rmse * gram_scale, so that dF = F m (one 1x1 conv_gemm per sample).
// grid = (j tiles, i, b): no integer division; the transposed read (j, i) goes through a 32x32 smem tile.
__global__ void gram_bwd_matrix_kernel(const float* __restrict__ gs, const float* __restrict__ gt, int B, int C,
                                       int Cp, float inv_count, float gram_scale, const float* __restrict__ gout,
                                       const float* __restrict__ msq, int mse, bf16* __restrict__ m) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  __shared__ float tile[32][33];
  const float coef = mse ? *gout * inv_count * 2.f * gram_scale
                         : *gout * inv_count / fmaxf(sqrtf(*msq), 1e-20f) * gram_scale;
  const int b = blockIdx.z;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float* ds = gs + (long long)b * C * C;
  const float* dt = gt + (long long)b * C * C;
  // tile[jj][ii] = D[j0+jj][i0+ii]  (rows j, coalesced along i)
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int jj = ty + k * 8, j = j0 + jj, i = i0 + tx;
    tile[jj][tx] = (j < C && i < C) ? ds[(long long)j * C + i] - dt[(long long)j * C + i] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ii = ty + k * 8, i = i0 + ii, j = j0 + tx;
    if (i < C && j < Cp) {
      float v = 0.f;
      if (j < C) v = coef * ((ds[(long long)i * C + j] - dt[(long long)i * C + j]) + tile[tx][ii]);
      m[((long long)b * C + i) * Cp + j] = __float2bfloat16(v);
    }
  }
}

// tiny scalar program: out = sqrt(in)   (RMSE from mean-square, kept on device)
__global__ void scalar_sqrt_kernel(const float* in, float* out) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents(); *out = sqrtf(fmaxf(*in, 0.f)); }

static inline int rblocks(long long n) {
  long long b = (n + 1023) / 1024;
  return (int)(b < 1 ? 1 : (b > 148 * 4 ? 148 * 4 : b));
}

}  // namespace gcc

using namespace gcc;

// out: fp32 scalar, must be zeroed by the caller (several terms may accumulate into one scalar).
extern "C" int gcc_gan_loss_fwd_bf16(const void* pred, long long npix, int Cp, int C, int mode, int kind, float* out,
                                     void* stream) {
  const float inv = 1.f / (float)(npix * C);
  gcc_launch(gan_loss_fwd_kernel, rblocks(npix * C), 256, 0, (cudaStream_t)stream, (const bf16*)pred, npix, Cp, C, mode, kind,
                                                                          inv, out);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_gan_loss_bwd_bf16(const void* pred, long long npix, int Cp, int C, int mode, int kind,
                                     const float* gout, void* dpred, void* stream) {
  const float inv = 1.f / (float)(npix * C);
  gcc_launch(gan_loss_bwd_kernel, rblocks(npix * Cp), 256, 0, (cudaStream_t)stream, (const bf16*)pred, npix, Cp, C, mode, kind,
                                                                           inv, gout, (bf16*)dpred);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
// mode 0: out += mean|a-b| ; mode 1: out += mean (a-b)^2   over the logical [npix][C] elements
extern "C" int gcc_diff_reduce_bf16(const void* a, const void* b, long long npix, int Cp, int C, int mode, float* out,
                                    void* stream) {
  const float inv = 1.f / (float)(npix * C);
  gcc_launch(diff_reduce_bf16_kernel, rblocks(npix * Cp / 4), 256, 0, (cudaStream_t)stream, (const bf16*)a, (const bf16*)b,
                                                                                   npix, Cp, C, mode, inv, out);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_diff_bwd_bf16(const void* a, const void* b, long long npix, int Cp, int C, int mode,
                                 const float* gout, const float* msq, void* da, void* stream) {
  const float inv = 1.f / (float)(npix * C);
  if (Cp % 8) { gcc_set_error(__FILE__, __LINE__, "diff_bwd: Cp must be a multiple of 8"); return GCC_ERR_ARG; }
  long long nb = (npix * (Cp / 8) + 255) / 256;
  if (nb > 148 * 16) nb = 148 * 16;
  if (nb < 1) nb = 1;
  gcc_launch(diff_bwd_bf16_kernel, (unsigned)nb, 256, 0, (cudaStream_t)stream, (const bf16*)a, (const bf16*)b, npix, Cp, C, mode,
                                                                      inv, gout, msq, (bf16*)da);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_sqdiff_reduce_f32(const float* a, const float* b, long long n, float* out, void* stream) {
  gcc_launch(sqdiff_reduce_f32_kernel, rblocks(n), 256, 0, (cudaStream_t)stream, a, b, n, 1.f / (float)n, out);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_gram_bwd_matrix(const float* gs, const float* gt, int B, int C, int Cp, float gram_scale,
                                   const float* gout, const float* msq, int mse, void* m, void* stream) {
  const float inv = 1.f / ((float)B * C * C);
  gcc_launch(gram_bwd_matrix_kernel, dim3((Cp + 31) / 32, (C + 31) / 32, B), 256, 0, (cudaStream_t)stream, 
      gs, gt, B, C, Cp, inv, gram_scale, gout, msq, mse, (bf16*)m);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_scalar_sqrt(const float* in, float* out, void* stream) {
  gcc_launch(scalar_sqrt_kernel, 1, 1, 0, (cudaStream_t)stream, in, out);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
