// "Slab" kernels for the InstanceNorm networks (MobileResNet generator blocks, models/Pix2Pix.py:132-197; CycleGAN's
// InstanceNorm discriminator, models/CycleGAN.py:140-178) at the resolutions where ONE CTA can own all H*W pixels of
// 8 channels of one sample (H*W <= 4096: 64 KB of bf16): InstanceNorm's statistics are then a block reduction and
// the whole  dw3x3 -> IN  or  IN -> activation (+ residual)  chain is one launch that reads its input once and writes
// its output once, instead of 3 (2) launches that move 10 (6) bytes per element plus a memset.  At batch 8 the
// separate kernels of these blocks run 5-25 us each (6 500 launches per CycleGAN iteration,
// profiles/r02_kernels_cyclegan.txt): the per-launch ramp, not HBM, bounds them.
//
//   dw_in_slab_fwd:  z = IN(dw3x3(reflect_pad(x)) + b)             SeparableConv2d conv.0 + conv.1 (Pix2Pix.py:137-141)
//   dw_in_slab_bwd:  dx, dw, db from dz, recomputing dw3x3(x) from the saved input (nothing but x, mean, rstd is saved)
//   in_act_slab_fwd: z = act(IN(y)) (+ residual)                   norm after the pointwise conv, ReLU, block skip
//   in_act_slab_bwd: dy from dz
// grid = (channel groups, samples), 256 threads, thread t owns pixels t, t + 256, ...
#include "common.cuh"

namespace gcc {

static constexpr int kSlabThreads = 256;
static constexpr int kSlabMaxPix = 4096;  // 16 pixels per thread (dw kernels), 8 per thread (IN kernels, 512 threads)
static constexpr int kInThreads = 512;
static constexpr int kInPix = kSlabMaxPix / kInThreads;

__device__ __forceinline__ void unpack8f(const uint4 u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8f(const float (&f)[8]) {
  return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}
__device__ __forceinline__ int refl(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}
// block-wide sums of 16 per-thread values (NWARPS warps); every thread gets the totals
template <int NWARPS>
__device__ __forceinline__ void block_sum16(float (&v)[16], float* red /* [NWARPS][16] */) {
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = warp_sum(v[i]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < 16; ++i) red[warp * 16 + i] = v[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < NWARPS; ++w) a += red[w * 16 + i];
    v[i] = a;
  }
}
// Slab index of pixel i: one 16-byte pad after every 16 pixels, so that threads walking consecutive 16-pixel row
// segments (start addresses 272 bytes apart) hit distinct shared-memory banks.
__device__ __forceinline__ int sidx(int i) { return i + (i >> 4); }
__host__ __device__ constexpr int slab_elems(int hw) { return hw + (hw >> 4) + 1; }

// Depthwise 3x3 with reflection padding 1 over the pixels [c0, c1) of row r, from the shared-memory slab: a 3 x 3
// window of unpacked fp32 vectors slides along the row (3 shared-memory loads and 24 conversions per pixel instead of
// 9 and 72; the loop is unrolled by 3 so that the window columns rotate by index, not by register moves).
// f(c, y, win, left, mid, right): y[8] = conv result at column c, win[a][col][8] the input window (col indices given).
template <typename F>
__device__ __forceinline__ void dw_row_run(const uint4* sx, int r, int c0, int c1, int H, int W, const float (&we)[3][3][8],
                                           const float (&bs)[8], F&& f) {
  int rowoff[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) rowoff[a] = refl(r + a - 1, H) * W;
  float win[3][3][8];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    unpack8f(sx[sidx(rowoff[a] + refl(c0 - 1, W))], win[a][0]);
    unpack8f(sx[sidx(rowoff[a] + c0)], win[a][1]);
  }
  for (int c = c0; c < c1; c += 3) {
#pragma unroll
    for (int ph = 0; ph < 3; ++ph) {
      const int cc = c + ph;
      if (cc < c1) {
        constexpr int kL[3] = {0, 1, 2}, kM[3] = {1, 2, 0}, kR[3] = {2, 0, 1};
        const int L = kL[ph], M = kM[ph], R = kR[ph];
        const int cn = refl(cc + 1, W);
#pragma unroll
        for (int a = 0; a < 3; ++a) unpack8f(sx[sidx(rowoff[a] + cn)], win[a][R]);
        float y[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) y[k] = bs[k];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int k = 0; k < 8; ++k)
            y[k] += win[a][L][k] * we[a][0][k] + win[a][M][k] * we[a][1][k] + win[a][R][k] * we[a][2][k];
        f(cc, y, win, L, M, R);
      }
    }
  }
}
// segment length so that the CTA's threads share the H * W pixels as row segments
__device__ __forceinline__ int slab_seg(int H, int W) {
  int sg = (H * W + kSlabThreads - 1) / kSlabThreads;
  if (sg < 1) sg = 1;
  if (sg > W) sg = W;
  return sg;
}

// stats: fp32 [N][Cp][2] = (mean, rstd) of the depthwise output, saved for the backward pass
__global__ void __launch_bounds__(kSlabThreads)
dw_in_slab_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                      bf16* __restrict__ z, float* __restrict__ stats, int H, int W, int G, int C, float eps) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ uint4 slab[];  // [H*W] input pixels of this (sample, channel group)
  __shared__ float red[8 * 16];
  const int g = blockIdx.x, n = blockIdx.y, HW = H * W;
  const uint4* xg = reinterpret_cast<const uint4*>(x) + (long long)n * HW * G + g;
  for (int p = threadIdx.x; p < HW; p += kSlabThreads) slab[sidx(p)] = xg[(long long)p * G];
  float we[3][3][8], bs[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = g * 8 + k;
    bs[k] = (bias != nullptr && c < C) ? bias[c] : 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) we[a][b][k] = c < C ? w[c * 9 + a * 3 + b] : 0.f;
  }
  __syncthreads();
  const int seg = slab_seg(H, W), segs = (W + seg - 1) / seg, nseg = H * segs;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int sgi = threadIdx.x; sgi < nseg; sgi += kSlabThreads) {
    const int r = sgi / segs, c0 = (sgi % segs) * seg;
    dw_row_run(slab, r, c0, min(W, c0 + seg), H, W, we, bs,
               [&](int, const float (&y)[8], const float (&)[3][3][8], int, int, int) {
#pragma unroll
                 for (int k = 0; k < 8; ++k) {
                   acc[k] += y[k];
                   acc[8 + k] += y[k] * y[k];
                 }
               });
  }
  block_sum16<kSlabThreads / 32>(acc, red);
  float mean[8], rstd[8];
  const float inv = 1.f / (float)HW;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    mean[k] = acc[k] * inv;
    rstd[k] = rsqrtf(fmaxf(acc[8 + k] * inv - mean[k] * mean[k], 0.f) + eps);
    if (g * 8 + k >= C) { mean[k] = 0.f; rstd[k] = 0.f; }
  }
  if (threadIdx.x < 8 && stats != nullptr) {
    float* so = stats + ((long long)n * G * 8 + g * 8 + threadIdx.x) * 2;
    so[0] = mean[threadIdx.x];
    so[1] = rstd[threadIdx.x];
  }
  uint4* zg = reinterpret_cast<uint4*>(z) + (long long)n * HW * G + g;
  for (int sgi = threadIdx.x; sgi < nseg; sgi += kSlabThreads) {
    const int r = sgi / segs, c0 = (sgi % segs) * seg;
    dw_row_run(slab, r, c0, min(W, c0 + seg), H, W, we, bs,
               [&](int c, const float (&y)[8], const float (&)[3][3][8], int, int, int) {
                 float o[8];
#pragma unroll
                 for (int k = 0; k < 8; ++k) o[k] = (y[k] - mean[k]) * rstd[k];
                 zg[(long long)(r * W + c) * G] = pack8f(o);
               });
  }
}

// dz: gradient of z = IN(dw(x) + b).  dx (bf16), dw [C][9] and dbias [C] (fp32, ACCUMULATED with atomics).
__global__ void __launch_bounds__(kSlabThreads)
dw_in_slab_bwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dz, const float* __restrict__ w,
                      const float* __restrict__ bias, const float* __restrict__ stats, bf16* __restrict__ dx,
                      float* __restrict__ dw, float* __restrict__ dbias, int H, int W, int G, int C) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ uint4 slab[];  // [H*W] x, then [H*W] dy1 (gradient of the depthwise output, bf16)
  __shared__ float red[8 * 16];
  __shared__ float fr[8 * 80];     // per-warp partial sums of the 72 + 8 parameter gradients
  const int g = blockIdx.x, n = blockIdx.y, HW = H * W;
  uint4* sx = slab;
  uint4* sd = slab + slab_elems(HW);
  const uint4* xg = reinterpret_cast<const uint4*>(x) + (long long)n * HW * G + g;
  const uint4* dzg = reinterpret_cast<const uint4*>(dz) + (long long)n * HW * G + g;
  for (int p = threadIdx.x; p < HW; p += kSlabThreads) sx[sidx(p)] = xg[(long long)p * G];
  float we[3][3][8], bs[8], mean[8], rstd[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = g * 8 + k;
    bs[k] = (bias != nullptr && c < C) ? bias[c] : 0.f;
    const float* so = stats + ((long long)n * G * 8 + c) * 2;
    mean[k] = so[0];
    rstd[k] = so[1];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) we[a][b][k] = c < C ? w[c * 9 + a * 3 + b] : 0.f;
  }
  __syncthreads();
  const int seg = slab_seg(H, W), segs = (W + seg - 1) / seg, nseg = H * segs;
  // pass 1: S1 = sum dz, S2 = sum dz * xhat  (xhat from the recomputed depthwise output)
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int sgi = threadIdx.x; sgi < nseg; sgi += kSlabThreads) {
    const int r = sgi / segs, c0 = (sgi % segs) * seg;
    dw_row_run(sx, r, c0, min(W, c0 + seg), H, W, we, bs,
               [&](int c, const float (&y)[8], const float (&)[3][3][8], int, int, int) {
                 float d[8];
                 unpack8f(dzg[(long long)(r * W + c) * G], d);
#pragma unroll
                 for (int k = 0; k < 8; ++k) {
                   acc[k] += d[k];
                   acc[8 + k] += d[k] * (y[k] - mean[k]) * rstd[k];
                 }
               });
  }
  block_sum16<kSlabThreads / 32>(acc, red);
  const float inv = 1.f / (float)HW;
  // pass 2: dy1 = rstd (dz - S1/M - xhat S2/M) into shared memory; weight / bias gradient accumulators
  {
    float gw[3][3][8], gb[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      gb[k] = 0.f;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) gw[a][b][k] = 0.f;
    }
    for (int sgi = threadIdx.x; sgi < nseg; sgi += kSlabThreads) {
      const int r = sgi / segs, c0 = (sgi % segs) * seg;
      dw_row_run(sx, r, c0, min(W, c0 + seg), H, W, we, bs,
                 [&](int c, const float (&y)[8], const float (&win)[3][3][8], int L, int M, int R) {
                   float d[8];
                   unpack8f(dzg[(long long)(r * W + c) * G], d);
#pragma unroll
                   for (int k = 0; k < 8; ++k) {
                     const float xh = (y[k] - mean[k]) * rstd[k];
                     d[k] = rstd[k] * (d[k] - acc[k] * inv - xh * acc[8 + k] * inv);
                   }
                   const uint4 packed = pack8f(d);
                   sd[sidx(r * W + c)] = packed;
                   unpack8f(packed, d);  // the bf16 values the data-gradient pass reads
#pragma unroll
                   for (int k = 0; k < 8; ++k) gb[k] += d[k];
#pragma unroll
                   for (int a = 0; a < 3; ++a)
#pragma unroll
                     for (int k = 0; k < 8; ++k) {
                       gw[a][0][k] += d[k] * win[a][L][k];
                       gw[a][1][k] += d[k] * win[a][M][k];
                       gw[a][2][k] += d[k] * win[a][R][k];
                     }
                 });
    }
    // block reduction of the 72 + 8 parameter gradients: warp shuffles, per-warp partials, one atomic per value
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      gb[k] = warp_sum(gb[k]);
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) gw[a][b][k] = warp_sum(gw[a][b][k]);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) fr[warp * 80 + k * 10 + a * 3 + b] = gw[a][b][k];
        fr[warp * 80 + k * 10 + 9] = gb[k];
      }
    }
  }
  __syncthreads();   // dy1 slab and the per-warp partials are complete
  if (threadIdx.x < 80) {
    float a = 0.f;
#pragma unroll
    for (int wp = 0; wp < kSlabThreads / 32; ++wp) a += fr[wp * 80 + threadIdx.x];
    const int k = threadIdx.x / 10, j = threadIdx.x % 10;
    const int c = g * 8 + k;
    if (c < C) {
      if (j < 9) {
        if (dw != nullptr) atomicAdd(dw + c * 9 + j, a);
      } else if (dbias != nullptr) {
        atomicAdd(dbias + c, a);
      }
    }
  }
  // pass 3: data gradient = zero-padded correlation of dy1 with the flipped taps + the mirrored border taps of the
  // fused ReflectionPad2d(1) (dx[1] += w[kh=0] dy1[0], dx[H-2] += w[kh=2] dy1[H-1]; same for columns)
  if (dx != nullptr) {
    uint4* dxg = reinterpret_cast<uint4*>(dx) + (long long)n * HW * G + g;
    for (int p = threadIdx.x; p < HW; p += kSlabThreads) {
      const int r = p / W, c = p % W;
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = 0.f;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const int rr = r + a - 1;
        if (rr < 0 || rr >= H) continue;
        // vertical tap(s) of this source row: kh = 2 - a, plus the mirrored one next to the border
        const bool top = (a == 0 && r == 1), bot = (a == 2 && r == H - 2);
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          const int cc = c + b - 1;
          if (cc < 0 || cc >= W) continue;
          float dv[8];
          unpack8f(sd[sidx(rr * W + cc)], dv);
          const bool lft = (b == 0 && c == 1), rgt = (b == 2 && c == W - 2);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            // horizontal taps: kw = 2 - b, plus the mirrored kw (0 at c == 1 from column 0, 2 at c == W-2 from column W-1)
            float wv = we[2 - a][2 - b][k];
            if (top) wv += we[0][2 - b][k];
            if (bot) wv += we[2][2 - b][k];
            if (lft) {
              wv += we[2 - a][0][k];
              if (top) wv += we[0][0][k];
              if (bot) wv += we[2][0][k];
            }
            if (rgt) {
              wv += we[2 - a][2][k];
              if (top) wv += we[0][2][k];
              if (bot) wv += we[2][2][k];
            }
            o[k] += dv[k] * wv;
          }
        }
      }
      dxg[(long long)p * G] = pack8f(o);
    }
  }
}

__device__ __forceinline__ float slab_act(float v, int act, float slope) {
  if (act == 1) return v > 0.f ? v : v * slope;
  if (act == 2) return v > 0.f ? v : 0.f;
  return v;
}
__device__ __forceinline__ float slab_act_grad(float v, int act, float slope) {
  if (act == 1) return v > 0.f ? 1.f : slope;
  if (act == 2) return v > 0.f ? 1.f : 0.f;
  return 1.f;
}

// z = act(IN(y)) (+ res).  act: 0 none, 1 leaky-relu(slope), 2 relu.  512 threads; the thread's <= 8 pixels stay in registers.
__global__ void __launch_bounds__(kInThreads)
in_act_slab_fwd_kernel(const bf16* __restrict__ y, const bf16* __restrict__ res, bf16* __restrict__ z,
                       float* __restrict__ stats, int HW, int G, int C, float eps, int act, float slope) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float red[(kInThreads / 32) * 16];
  const int g = blockIdx.x, n = blockIdx.y;
  const uint4* yg = reinterpret_cast<const uint4*>(y) + (long long)n * HW * G + g;
  uint4 v[kInPix];
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
#pragma unroll
  for (int i = 0; i < kInPix; ++i) {
    const int p = threadIdx.x + i * kInThreads;
    v[i] = p < HW ? yg[(long long)p * G] : make_uint4(0, 0, 0, 0);
  }
#pragma unroll
  for (int i = 0; i < kInPix; ++i) {
    float f[8];
    unpack8f(v[i], f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      acc[k] += f[k];
      acc[8 + k] += f[k] * f[k];
    }
  }
  block_sum16<kInThreads / 32>(acc, red);
  float mean[8], rstd[8];
  const float inv = 1.f / (float)HW;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    mean[k] = acc[k] * inv;
    rstd[k] = rsqrtf(fmaxf(acc[8 + k] * inv - mean[k] * mean[k], 0.f) + eps);
    if (g * 8 + k >= C) { mean[k] = 0.f; rstd[k] = 0.f; }
  }
  if (threadIdx.x < 8 && stats != nullptr) {
    float* so = stats + ((long long)n * G * 8 + g * 8 + threadIdx.x) * 2;
    so[0] = mean[threadIdx.x];
    so[1] = rstd[threadIdx.x];
  }
  uint4* zg = reinterpret_cast<uint4*>(z) + (long long)n * HW * G + g;
  const uint4* rg = res != nullptr ? reinterpret_cast<const uint4*>(res) + (long long)n * HW * G + g : nullptr;
#pragma unroll
  for (int i = 0; i < kInPix; ++i) {
    const int p = threadIdx.x + i * kInThreads;
    if (p < HW) {
      float f[8], r8[8];
      unpack8f(v[i], f);
      if (rg != nullptr) unpack8f(rg[(long long)p * G], r8);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        f[k] = slab_act((f[k] - mean[k]) * rstd[k], act, slope);
        if (rg != nullptr) f[k] += r8[k];
      }
      zg[(long long)p * G] = pack8f(f);
    }
  }
}

// dy from dz (gradient of z = act(IN(y)) [+ res]: the residual's own gradient is dz itself)
__global__ void __launch_bounds__(kInThreads)
in_act_slab_bwd_kernel(const bf16* __restrict__ y, const bf16* __restrict__ dz, const float* __restrict__ stats,
                       bf16* __restrict__ dy, int HW, int G, int C, int act, float slope) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float red[(kInThreads / 32) * 16];
  const int g = blockIdx.x, n = blockIdx.y;
  const uint4* yg = reinterpret_cast<const uint4*>(y) + (long long)n * HW * G + g;
  const uint4* dg = reinterpret_cast<const uint4*>(dz) + (long long)n * HW * G + g;
  float mean[8], rstd[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float* so = stats + ((long long)n * G * 8 + g * 8 + k) * 2;
    mean[k] = so[0];
    rstd[k] = so[1];
  }
  uint4 v[kInPix], d[kInPix];
#pragma unroll
  for (int i = 0; i < kInPix; ++i) {
    const int p = threadIdx.x + i * kInThreads;
    v[i] = p < HW ? yg[(long long)p * G] : make_uint4(0, 0, 0, 0);
    d[i] = p < HW ? dg[(long long)p * G] : make_uint4(0, 0, 0, 0);
  }
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
#pragma unroll
  for (int i = 0; i < kInPix; ++i) {
    float f[8], e[8];
    unpack8f(v[i], f);
    unpack8f(d[i], e);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float xh = (f[k] - mean[k]) * rstd[k];
      const float dgv = e[k] * slab_act_grad(xh, act, slope);
      acc[k] += dgv;
      acc[8 + k] += dgv * xh;
    }
  }
  block_sum16<kInThreads / 32>(acc, red);
  const float inv = 1.f / (float)HW;
  uint4* og = reinterpret_cast<uint4*>(dy) + (long long)n * HW * G + g;
#pragma unroll
  for (int i = 0; i < kInPix; ++i) {
    const int p = threadIdx.x + i * kInThreads;
    if (p < HW) {
      float f[8], e[8];
      unpack8f(v[i], f);
      unpack8f(d[i], e);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float xh = (f[k] - mean[k]) * rstd[k];
        const float dgv = e[k] * slab_act_grad(xh, act, slope);
        f[k] = rstd[k] * (dgv - acc[k] * inv - xh * acc[8 + k] * inv);
      }
      og[(long long)p * G] = pack8f(f);
    }
  }
}

}  // namespace gcc

using namespace gcc;

static int slab_check(int H, int W, int Cp, const char* what) {
  if ((Cp % 8) || H < 2 || W < 2 || (long long)H * W > kSlabMaxPix) {
    gcc_set_error(__FILE__, __LINE__, what);
    return GCC_ERR_ARG;
  }
  return GCC_OK;
}
static int slab_smem_attr(const void* fn) {
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
    gcc_set_error(__FILE__, __LINE__, "cudaFuncSetAttribute failed");
    return GCC_ERR_CUDA;
  }
  return GCC_OK;
}

// 1 when the slab kernels can run this shape (all H*W pixels of 8 channels of one sample in one CTA)
extern "C" int gcc_slab_supported(int H, int W) { return (H >= 2 && W >= 2 && (long long)H * W <= kSlabMaxPix) ? 1 : 0; }

extern "C" int gcc_dw_in_slab_fwd_bf16(const void* x, const float* w, const float* bias, void* z, float* stats, int N, int H,
                                       int W, int Cp, int C, float eps, void* stream) {
  if (int rc = slab_check(H, W, Cp, "dw_in_slab_fwd: needs Cp % 8 == 0 and 4 <= H*W <= 4096")) return rc;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  dev = (dev >= 0 && dev < 64) ? dev : 0;
  if (!configured[dev]) {
    if (int rc = slab_smem_attr((const void*)dw_in_slab_fwd_kernel)) return rc;
    configured[dev] = true;
  }
  gcc_launch(dw_in_slab_fwd_kernel, dim3(Cp / 8, N), kSlabThreads, (size_t)slab_elems(H * W) * 16, (cudaStream_t)stream, (const bf16*)x, w,
             bias, (bf16*)z, stats, H, W, Cp / 8, C, eps);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_dw_in_slab_bwd_bf16(const void* x, const void* dz, const float* w, const float* bias, const float* stats,
                                       void* dx, float* dw, float* dbias, int N, int H, int W, int Cp, int C, void* stream) {
  if (int rc = slab_check(H, W, Cp, "dw_in_slab_bwd: needs Cp % 8 == 0 and 4 <= H*W <= 4096")) return rc;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  dev = (dev >= 0 && dev < 64) ? dev : 0;
  if (!configured[dev]) {
    if (int rc = slab_smem_attr((const void*)dw_in_slab_bwd_kernel)) return rc;
    configured[dev] = true;
  }
  const size_t smem = (size_t)slab_elems(H * W) * 32;
  gcc_launch(dw_in_slab_bwd_kernel, dim3(Cp / 8, N), kSlabThreads, smem, (cudaStream_t)stream, (const bf16*)x, (const bf16*)dz,
             w, bias, stats, (bf16*)dx, dw, dbias, H, W, Cp / 8, C);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_in_act_slab_fwd_bf16(const void* y, const void* res, void* z, float* stats, int N, long long HW, int Cp, int C,
                                        float eps, int act, float slope, void* stream) {
  if ((Cp % 8) || HW < 1 || HW > kSlabMaxPix) {
    gcc_set_error(__FILE__, __LINE__, "in_act_slab_fwd: needs Cp % 8 == 0 and H*W <= 4096");
    return GCC_ERR_ARG;
  }
  gcc_launch(in_act_slab_fwd_kernel, dim3(Cp / 8, N), kInThreads, 0, (cudaStream_t)stream, (const bf16*)y, (const bf16*)res,
             (bf16*)z, stats, (int)HW, Cp / 8, C, eps, act, slope);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_in_act_slab_bwd_bf16(const void* y, const void* dz, const float* stats, void* dy, int N, long long HW, int Cp,
                                        int C, int act, float slope, void* stream) {
  if ((Cp % 8) || HW < 1 || HW > kSlabMaxPix) {
    gcc_set_error(__FILE__, __LINE__, "in_act_slab_bwd: needs Cp % 8 == 0 and H*W <= 4096");
    return GCC_ERR_ARG;
  }
  gcc_launch(in_act_slab_bwd_kernel, dim3(Cp / 8, N), kInThreads, 0, (cudaStream_t)stream, (const bf16*)y, (const bf16*)dz, stats,
             (bf16*)dy, (int)HW, Cp / 8, C, act, slope);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
