// Batch-norm / instance-norm / channel-gate / activation kernels over NHWC bf16 activations.
//
// Forward of one "norm block" (reference: nn.BatchNorm2d / nn.InstanceNorm2d + DifferentiableOP +
// (Leaky)ReLU, models/Pix2Pix.py:26-35,201,272-341 and models/DifferentiableOp.py:44-49):
//     z = gamma * (x - mean) * rstd + beta          (identity when there is no norm)
//     g = mask * z,  mask = (sign(alpha - t) + 1)/2 (1 when there is no gate)
//     y = act(g)                                     act in {none, leaky-relu, relu}
// and optionally a second output y2 = act2(g) written into a channel window of a wider buffer
// (the U-Net skip concat: the reference's in-place LeakyReLU/ReLU aliasing, Pix2Pix.py:33,35,77).
//
// All kernels are HBM-bound: 16-byte vector accesses, channel index = vector index % (Cp/8), fp32
// statistics accumulated per block in shared memory then one atomicAdd per (block, channel).
#include <stdlib.h>

#include "common.cuh"

namespace gcc {

struct Vec8 {
  float v[8];
};

__device__ __forceinline__ Vec8 load8(const bf16* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  Vec8 r;
  r.v[0] = bf16_lo(u.x); r.v[1] = bf16_hi(u.x);
  r.v[2] = bf16_lo(u.y); r.v[3] = bf16_hi(u.y);
  r.v[4] = bf16_lo(u.z); r.v[5] = bf16_hi(u.z);
  r.v[6] = bf16_lo(u.w); r.v[7] = bf16_hi(u.w);
  return r;
}
__device__ __forceinline__ Vec8 unpack8(const uint4 u) {
  Vec8 r;
  r.v[0] = bf16_lo(u.x); r.v[1] = bf16_hi(u.x);
  r.v[2] = bf16_lo(u.y); r.v[3] = bf16_hi(u.y);
  r.v[4] = bf16_lo(u.z); r.v[5] = bf16_hi(u.z);
  r.v[6] = bf16_lo(u.w); r.v[7] = bf16_hi(u.w);
  return r;
}
__device__ __forceinline__ void store8(bf16* p, const Vec8& r) {
  uint4 u;
  u.x = pack_bf16(r.v[0], r.v[1]);
  u.y = pack_bf16(r.v[2], r.v[3]);
  u.z = pack_bf16(r.v[4], r.v[5]);
  u.w = pack_bf16(r.v[6], r.v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

__device__ __forceinline__ float act_fwd(float g, int act, float slope) {
  if (act == 1) return g > 0.f ? g : g * slope;
  if (act == 2) return g > 0.f ? g : 0.f;
  return g;
}
__device__ __forceinline__ float act_grad(float g, int act, float slope) {
  if (act == 1) return g > 0.f ? 1.f : slope;
  if (act == 2) return g > 0.f ? 1.f : 0.f;
  return 1.f;
}
__device__ __forceinline__ float gate_value(const float* alpha, float thr, int c) {
  if (alpha == nullptr) return 1.f;
  const float d = alpha[c] - thr;
  const float s = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
  return (s + 1.f) * 0.5f;
}

// ---------------------------------------------------------------------------------- statistics
// sums[(n)][0:Cp) += sum x ; sums[(n)][Cp:2Cp) += sum x^2   (grid.y = n when per_sample)
__global__ void __launch_bounds__(256, 4) norm_stats_kernel(const bf16* __restrict__ x, long long npix, int Cp, int G,
                                                            int lanes, float* __restrict__ sums) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  extern __shared__ float red[];  // [lanes][G][16]
  const int tid = threadIdx.x;
  const int g = tid % G, lane = tid / G;
  const bf16* xb = x + (long long)blockIdx.y * npix * Cp + g * 8;
  float* sb = sums + (long long)blockIdx.y * 2 * Cp;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  if (lane < lanes) {
    const long long step = (long long)gridDim.x * lanes;
    const uint4 zero = make_uint4(0, 0, 0, 0);
    auto ld = [&](long long q) { return q < npix ? *reinterpret_cast<const uint4*>(xb + q * Cp) : zero; };
    // register double buffer: the next four 16-byte loads are issued before the current four are consumed
    long long p = (long long)blockIdx.x * lanes + lane;
    uint4 cur[4], nxt[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) cur[j] = ld(p + j * step);
    while (p < npix) {
      p += 4 * step;
#pragma unroll
      for (int j = 0; j < 4; ++j) nxt[j] = ld(p + j * step);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const Vec8 a = unpack8(cur[j]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s[i] += a.v[i];
          q[i] += a.v[i] * a.v[i];
        }
        cur[j] = nxt[j];
      }
    }
    float* r = red + ((long long)lane * G + g) * 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      r[i] = s[i];
      r[8 + i] = q[i];
    }
  }
  __syncthreads();
  for (int e = tid; e < G * 16; e += blockDim.x) {
    const int gg = e / 16, j = e % 16;
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += red[((long long)l * G + gg) * 16 + j];
    const int c = gg * 8 + (j & 7);
    atomicAdd(sb + (j < 8 ? 0 : Cp) + c, acc);
  }
}

struct NormArgs {
  const bf16* x;
  long long npix;  // pixels per statistics group (N*H*W for batch norm, H*W for instance norm) on THIS device
  long long stat_npix;  // pixels behind `sums` / `red` (= npix, or the global count under sync batch norm)
  int Cp, C, G;
  int per_sample;
  const float* sums;   // nullptr: identity (no normalisation)
  const float* gamma;  // nullptr: 1
  const float* beta;   // nullptr: 0
  const float* alpha;  // nullptr: no gate
  float thr, eps;
  int act;
  float slope;
  int gate_after;  // 1: y = mask * act(z) (identity norm only: PatchGAN layer 0, Pix2Pix.py:320-322)
};

__device__ __forceinline__ void channel_affine(const NormArgs& a, int n, int c, float& mean, float& rstd,
                                               float& gam, float& bet, float& mask) {
  mean = 0.f;
  rstd = 1.f;
  if (a.sums != nullptr) {
    const float* sb = a.sums + (long long)(a.per_sample ? n : 0) * 2 * a.Cp;
    const float inv = 1.f / (float)a.stat_npix;
    mean = sb[c] * inv;
    const float var = fmaxf(sb[a.Cp + c] * inv - mean * mean, 0.f);
    rstd = rsqrtf(var + a.eps);
  }
  const bool live = c < a.C;
  gam = (a.gamma != nullptr && live) ? a.gamma[c] : (live ? 1.f : 0.f);
  bet = (a.beta != nullptr && live) ? a.beta[c] : 0.f;
  mask = (a.alpha != nullptr && live) ? gate_value(a.alpha, a.thr, c) : 1.f;
  if (!live) { mean = 0.f; rstd = 0.f; }
}

// -------------------------------------------------------------------------------------- forward
// Thread layout as in the statistics kernel: thread = (channel group g of 8, pixel lane); the 8 channels'
// coefficients live in registers for the whole kernel and the pixel loop has no integer division:
//   y = act(x*p + q) * m   with p = rstd*gamma*mask, q = (beta - mean*rstd*gamma)*mask, m = 1
//   (gate_after: p, q without the mask and m = mask).  grid.y = sample index for instance norm.
template <bool DUAL>
__global__ void __launch_bounds__(256, DUAL ? 2 : 4)
norm_apply_kernel(NormArgs a, int lanes, bf16* __restrict__ y, bf16* __restrict__ y2, int y2_Cp, int y2_coff, int act2,
                  float* running_mean, float* running_var, float momentum) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const int tid = threadIdx.x;
  const int g = tid % a.G, lane = tid / a.G;
  const int n = blockIdx.y;
  if (lane < lanes) {
    float cp[8], cq[8], cm[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float mean, rstd, gam, bet, mask;
      channel_affine(a, n, g * 8 + k, mean, rstd, gam, bet, mask);
      const float mk = a.gate_after ? 1.f : mask;
      cp[k] = rstd * gam * mk;
      cq[k] = (bet - mean * rstd * gam) * mk;
      cm[k] = a.gate_after ? mask : 1.f;
    }
    const long long pix0 = a.per_sample ? (long long)n * a.npix : 0;
    const long long step = (long long)gridDim.x * lanes;
    const bf16* xb = a.x + pix0 * a.Cp + g * 8;
    bf16* yb = y ? y + pix0 * a.Cp + g * 8 : nullptr;
    bf16* y2b = DUAL ? y2 + pix0 * y2_Cp + y2_coff + g * 8 : nullptr;
    auto emit = [&](const uint4 u, long long p) {
      const Vec8 xv = unpack8(u);
      Vec8 o, o2;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float z = xv.v[k] * cp[k] + cq[k];
        o.v[k] = act_fwd(z, a.act, a.slope) * cm[k];
        if (DUAL) o2.v[k] = act_fwd(z, act2, a.slope) * cm[k];
      }
      if (yb) store8(yb + p * a.Cp, o);
      if (DUAL) store8(y2b + p * y2_Cp, o2);
    };
    long long p = (long long)blockIdx.x * lanes + lane;
    for (; p + 3 * step < a.npix; p += 4 * step) {  // four independent 16-byte loads in flight per thread
      uint4 u[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) u[j] = *reinterpret_cast<const uint4*>(xb + (p + j * step) * a.Cp);
#pragma unroll
      for (int j = 0; j < 4; ++j) emit(u[j], p + j * step);
    }
    for (; p < a.npix; p += step) emit(*reinterpret_cast<const uint4*>(xb + p * a.Cp), p);
  }
  // running statistics (train-mode BatchNorm2d side effect; momentum 0.1, unbiased variance)
  if (running_mean != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && a.sums != nullptr) {
    for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
      const float inv = 1.f / (float)a.stat_npix;
      const float mean = a.sums[c] * inv;
      const float var = fmaxf(a.sums[a.Cp + c] * inv - mean * mean, 0.f);
      const float unb = a.stat_npix > 1 ? var * ((float)a.stat_npix / (float)(a.stat_npix - 1)) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * unb;
    }
  }
}

// eval-mode batch norm: statistics come from the running buffers
__global__ void norm_apply_eval_kernel(NormArgs a, const float* __restrict__ rmean, const float* __restrict__ rvar,
                                       long long total_pix, bf16* __restrict__ y, bf16* __restrict__ y2, int y2_Cp,
                                       int y2_coff, int act2) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const long long nvec = total_pix * a.G;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % a.G);
    const long long pix = i / a.G;
    const Vec8 xv = load8(a.x + pix * a.Cp + g * 8);
    Vec8 o, o2;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = g * 8 + k;
      float z = 0.f, mask = 1.f;
      if (c < a.C) {
        const float rstd = rsqrtf(rvar[c] + a.eps);
        z = (xv.v[k] - rmean[c]) * rstd * (a.gamma ? a.gamma[c] : 1.f) + (a.beta ? a.beta[c] : 0.f);
        mask = gate_value(a.alpha, a.thr, c);
      }
      if (a.gate_after) {
        o.v[k] = act_fwd(z, a.act, a.slope) * mask;
        o2.v[k] = act_fwd(z, act2, a.slope) * mask;
      } else {
        const float gg = z * mask;
        o.v[k] = act_fwd(gg, a.act, a.slope);
        o2.v[k] = act_fwd(gg, act2, a.slope);
      }
    }
    if (y != nullptr) store8(y + pix * a.Cp + g * 8, o);
    if (y2 != nullptr) store8(y2 + pix * y2_Cp + y2_coff + g * 8, o2);
  }
}

// ------------------------------------------------------------------------------------- backward
// red[(n)][0:Cp) += sum dg ; red[(n)][Cp:2Cp) += sum dg * xhat      dg = dy*act'(g) + dy2*act2'(g)
template <int U, bool HAS_D2, bool PIPE = true>
__global__ void __launch_bounds__(256, 2)
norm_bwd_reduce_kernel(NormArgs a, int lanes, const bf16* __restrict__ dy, int dy_Cp, int dy_coff,
                       const bf16* __restrict__ dy2, int dy2_Cp, int dy2_coff, int act2, float* __restrict__ red) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  extern __shared__ float sred[];  // [lanes][G][16]
  const int tid = threadIdx.x;
  const int g = tid % a.G, lane = tid / a.G;
  const int n = blockIdx.y;
  const long long pix0 = (long long)n * a.npix;  // blockIdx.y == 0 for batch norm
  float s1[8], s2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
  if (lane < lanes) {
    // xhat = x*rs + ms ;  z = xhat*gam + bet ;  gg = gate_after ? z : z*mask  ==  x*cg + cb.
    // The loop accumulates S1 = sum dg and sum dg*x on the RAW x (two coefficient arrays live instead of four);
    // sum dg*xhat = rs * sum dg*x + ms * S1 is formed once after the loop.
    float cg[8], cb[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float mean, rstd, gam, bet, mask;
      channel_affine(a, n, g * 8 + k, mean, rstd, gam, bet, mask);
      const float mk = a.gate_after ? 1.f : mask;
      cg[k] = rstd * gam * mk;
      cb[k] = (bet - mean * rstd * gam) * mk;
    }
    const bf16* xb = a.x + pix0 * a.Cp + g * 8;
    const bf16* d1b = dy ? dy + pix0 * dy_Cp + dy_coff + g * 8 : nullptr;
    const bf16* d2b = (HAS_D2 && dy2) ? dy2 + pix0 * dy2_Cp + dy2_coff + g * 8 : nullptr;
    auto accum = [&](const uint4 ux, const uint4 u1, const uint4 u2) {
      const Vec8 xv = unpack8(ux), d1 = unpack8(u1), d2 = unpack8(u2);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float gg = xv.v[k] * cg[k] + cb[k];
        float dg = 0.f;
        if (d1b) dg += d1.v[k] * act_grad(gg, a.act, a.slope);
        if (HAS_D2 && d2b) dg += d2.v[k] * act_grad(gg, act2, a.slope);
        s1[k] += dg;
        if (a.gate_after) {  // S2 = sum dy * act(z): the gate gradient of y = mask * act(z)
          if (d1b) s2[k] += d1.v[k] * act_fwd(gg, a.act, a.slope);
          if (HAS_D2 && d2b) s2[k] += d2.v[k] * act_fwd(gg, act2, a.slope);
        } else {
          s2[k] += dg * xv.v[k];
        }
      }
    };
    const uint4 zero = make_uint4(0, 0, 0, 0);
    const long long step = (long long)gridDim.x * lanes;
    auto ldx = [&](long long q) { return q < a.npix ? *reinterpret_cast<const uint4*>(xb + q * a.Cp) : zero; };
    auto ld1 = [&](long long q) { return (d1b && q < a.npix) ? *reinterpret_cast<const uint4*>(d1b + q * dy_Cp) : zero; };
    auto ld2 = [&](long long q) { return (HAS_D2 && d2b && q < a.npix) ? *reinterpret_cast<const uint4*>(d2b + q * dy2_Cp) : zero; };
    // register double buffer over groups of U pixels (out-of-range pixels load zeros and contribute nothing)
    long long p = (long long)blockIdx.x * lanes + lane;
    uint4 cx[U], c1[U], c2[U], nx[U], n1[U], n2[U];
    if (!PIPE) {  // U loads in flight, no register double buffer (the other resident warps cover the compute phase)
      for (; p < a.npix; p += U * step) {
#pragma unroll
        for (int j = 0; j < U; ++j) { cx[j] = ldx(p + j * step); c1[j] = ld1(p + j * step); c2[j] = ld2(p + j * step); }
#pragma unroll
        for (int j = 0; j < U; ++j) accum(cx[j], c1[j], c2[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < U; ++j) { cx[j] = ldx(p + j * step); c1[j] = ld1(p + j * step); c2[j] = ld2(p + j * step); }
    while (p < a.npix) {
      p += U * step;
#pragma unroll
      for (int j = 0; j < U; ++j) { nx[j] = ldx(p + j * step); n1[j] = ld1(p + j * step); n2[j] = ld2(p + j * step); }
#pragma unroll
      for (int j = 0; j < U; ++j) {
        accum(cx[j], c1[j], c2[j]);
        cx[j] = nx[j]; c1[j] = n1[j]; c2[j] = n2[j];
      }
    }
    float* r = sred + ((long long)lane * a.G + g) * 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (!a.gate_after) {
        float mean, rstd, gam, bet, mask;
        channel_affine(a, n, g * 8 + i, mean, rstd, gam, bet, mask);
        s2[i] = rstd * (s2[i] - mean * s1[i]);
      }
      r[i] = s1[i];
      r[8 + i] = s2[i];
    }
  }
  __syncthreads();
  float* rb = red + (long long)n * 2 * a.Cp;
  for (int e = tid; e < a.G * 16; e += blockDim.x) {
    const int gg = e / 16, j = e % 16;
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += sred[((long long)l * a.G + gg) * 16 + j];
    atomicAdd(rb + (j < 8 ? 0 : a.Cp) + gg * 8 + (j & 7), acc);
  }
}

// dx = gamma*rstd*mask * (dg - (S1 + xhat*S2)/M)   [norm]   or   dx = mask*dg [identity]
// with per-thread register coefficients: gg = x*p + q, dx = c1*dg - c2 - c3*x  (thread = channel group x lane).
// block (0,0) also accumulates dgamma += mask*S2, dbeta += mask*S1, dalpha += gamma*S2 + beta*S1 (summed over n).
template <bool HAS_D2, int U>
__global__ void __launch_bounds__(256, 2) norm_bwd_apply_kernel(NormArgs a, int nimg, int lanes, const bf16* __restrict__ dy, int dy_Cp,
                                      int dy_coff, const bf16* __restrict__ dy2, int dy2_Cp, int dy2_coff, int act2,
                                      const float* __restrict__ red, const float* __restrict__ red_param,
                                      bf16* __restrict__ dx, float* dgamma, float* dbeta, float* dalpha) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const int tid = threadIdx.x;
  const int g = tid % a.G, lane = tid / a.G;
  const int n = blockIdx.y;
  const float invM = 1.f / (float)a.stat_npix;
  if (dx != nullptr && lane < lanes) {
    const float* rb = red + (long long)n * 2 * a.Cp;
    float cp[8], cq[8], c1[8], c2[8], c3[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = g * 8 + k;
      float mean, rstd, gam, bet, mask;
      channel_affine(a, n, c, mean, rstd, gam, bet, mask);
      const float mk = a.gate_after ? 1.f : mask;
      cp[k] = rstd * gam * mk;
      cq[k] = (bet - mean * rstd * gam) * mk;
      const float k1 = gam * rstd * mask;
      c1[k] = (c < a.C) ? k1 : 0.f;
      if (a.sums != nullptr && c < a.C) {
        c2[k] = k1 * invM * (rb[c] - rb[a.Cp + c] * rstd * mean);
        c3[k] = k1 * invM * rb[a.Cp + c] * rstd;
      } else {
        c2[k] = 0.f;
        c3[k] = 0.f;
      }
    }
    const long long pix0 = a.per_sample ? (long long)n * a.npix : 0;
    const long long step = (long long)gridDim.x * lanes;
    const bf16* xb = a.x + pix0 * a.Cp + g * 8;
    const bf16* d1b = dy ? dy + pix0 * dy_Cp + dy_coff + g * 8 : nullptr;
    const bf16* d2b = (HAS_D2 && dy2) ? dy2 + pix0 * dy2_Cp + dy2_coff + g * 8 : nullptr;
    bf16* ob = dx + pix0 * a.Cp + g * 8;
    auto emit = [&](const uint4 ux, const uint4 u1, const uint4 u2, long long p) {
      const Vec8 xv = unpack8(ux), d1 = unpack8(u1), d2 = unpack8(u2);
      Vec8 o;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float gg = xv.v[k] * cp[k] + cq[k];
        float dg = 0.f;
        if (d1b) dg += d1.v[k] * act_grad(gg, a.act, a.slope);
        if (HAS_D2 && d2b) dg += d2.v[k] * act_grad(gg, act2, a.slope);
        o.v[k] = c1[k] * dg - c2[k] - c3[k] * xv.v[k];
      }
      store8(ob + p * a.Cp, o);
    };
    const uint4 zero = make_uint4(0, 0, 0, 0);
    auto ldx = [&](long long q) { return q < a.npix ? *reinterpret_cast<const uint4*>(xb + q * a.Cp) : zero; };
    auto ld1 = [&](long long q) { return (d1b && q < a.npix) ? *reinterpret_cast<const uint4*>(d1b + q * dy_Cp) : zero; };
    auto ld2 = [&](long long q) { return (HAS_D2 && d2b && q < a.npix) ? *reinterpret_cast<const uint4*>(d2b + q * dy2_Cp) : zero; };
    // register double buffer over groups of U pixels
    long long p = (long long)blockIdx.x * lanes + lane;
    uint4 cx[U], e1[U], e2[U], nx[U], n1[U], n2[U];
#pragma unroll
    for (int j = 0; j < U; ++j) { cx[j] = ldx(p + j * step); e1[j] = ld1(p + j * step); e2[j] = ld2(p + j * step); }
    while (p < a.npix) {
      const long long pn = p + U * step;
#pragma unroll
      for (int j = 0; j < U; ++j) { nx[j] = ldx(pn + j * step); n1[j] = ld1(pn + j * step); n2[j] = ld2(pn + j * step); }
#pragma unroll
      for (int j = 0; j < U; ++j) {
        if (p + j * step < a.npix) emit(cx[j], e1[j], e2[j], p + j * step);
        cx[j] = nx[j]; e1[j] = n1[j]; e2[j] = n2[j];
      }
      p = pn;
    }
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && (dgamma != nullptr || dbeta != nullptr || dalpha != nullptr)) {
    const int groups = a.per_sample ? nimg : 1;
    for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
      float S1 = 0.f, S2 = 0.f;
      for (int m = 0; m < groups; ++m) {  // parameter gradients come from THIS device's sums (red_param)
        S1 += red_param[(long long)m * 2 * a.Cp + c];
        S2 += red_param[(long long)m * 2 * a.Cp + a.Cp + c];
      }
      const float gam = a.gamma ? a.gamma[c] : 1.f;
      const float bet = a.beta ? a.beta[c] : 0.f;
      const float mask = a.alpha ? gate_value(a.alpha, a.thr, c) : 1.f;
      // parameter gradients ACCUMULATE (the arena is zeroed by zero_grad)
      if (dgamma) dgamma[c] += mask * S2;
      if (dbeta) dbeta[c] += mask * S1;
      if (dalpha) dalpha[c] += a.gate_after ? S2 : gam * S2 + bet * S1;
    }
  }
}

static inline int stats_threads(int G, int* lanes) {
  int l = 256 / G;
  if (l < 1) l = 1;
  *lanes = l;
  int t = l * G;
  return (t + 31) / 32 * 32;
}

}  // namespace gcc

using namespace gcc;

static int fill_args(NormArgs& a, const void* x, int N, long long HW, int Cp, int C, int per_sample,
                     const float* sums, const float* gamma, const float* beta, const float* alpha, float thr,
                     float eps, int act, float slope, int gate_after = 0, long long stat_count = 0) {
  if (Cp % 8 || C > Cp || Cp / 8 > 512) {
    gcc_set_error(__FILE__, __LINE__, "norm: channel count must be a multiple of 8 and <= 4096");
    return GCC_ERR_ARG;
  }
  a.x = (const bf16*)x;
  a.npix = per_sample ? HW : (long long)N * HW;
  a.stat_npix = (stat_count > 0 && !per_sample) ? stat_count : a.npix;
  a.Cp = Cp; a.C = C; a.G = Cp / 8;
  a.per_sample = per_sample;
  a.sums = sums; a.gamma = gamma; a.beta = beta; a.alpha = alpha;
  a.thr = thr; a.eps = eps; a.act = act; a.slope = slope;
  a.gate_after = gate_after;
  if (gate_after && sums != nullptr) {
    gcc_set_error(__FILE__, __LINE__, "norm: gate_after_act is only defined for the identity norm");
    return GCC_ERR_ARG;
  }
  return GCC_OK;
}

// blocks for a (group, lane) kernel: each thread visits >= `per` pixels, whole grid <= ~16 CTAs per SM
// `waves` = CTAs per SM the grid is capped at.  The forward apply kernel (4 resident CTAs per SM, 4 B / element) does
// not care (16); the backward apply kernel (2 resident CTAs, 6 B / element, a dependent coefficient prologue per CTA)
// wants exactly ONE resident wave: measured on the PatchGAN activations at batch 32 (scripts/exp_norm_unroll.py,
// L2 flushed): 96 -> 76 us (134 MB tensors), 66 -> 45, 39 -> 29, 62 -> 43 us with the cap at 2 and 4 pixels in flight.
static int lane_blocks(long long npix, int lanes, int groups, int per, int waves = 16) {
  long long b = (npix + (long long)lanes * per * 4 - 1) / ((long long)lanes * per * 4);
  long long cap = (148LL * waves + groups - 1) / groups;
  if (cap < 1) cap = 1;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

static int ew_blocks(long long nvec) {
  long long b = (nvec + 255) / 256;
  const long long cap = 148LL * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// sums: fp32 [N if per_sample else 1][2][Cp]; zeroed here (zero_sums) or by the caller (gcc_norm_stats_acc_bf16: the
// host side hands out pre-zeroed scratch that one fill per training phase clears, instead of one memset node per call).
static int norm_stats_launch(const void* x, int N, long long HW, int Cp, int per_sample, float* sums, int zero_sums,
                             void* stream);
extern "C" int gcc_norm_stats_bf16(const void* x, int N, long long HW, int Cp, int per_sample, float* sums,
                                   void* stream) {
  return norm_stats_launch(x, N, HW, Cp, per_sample, sums, 1, stream);
}
extern "C" int gcc_norm_stats_acc_bf16(const void* x, int N, long long HW, int Cp, int per_sample, float* sums,
                                       void* stream) {
  return norm_stats_launch(x, N, HW, Cp, per_sample, sums, 0, stream);
}
static int norm_stats_launch(const void* x, int N, long long HW, int Cp, int per_sample, float* sums, int zero_sums,
                             void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (Cp % 8 || Cp / 8 > 512) {
    gcc_set_error(__FILE__, __LINE__, "norm_stats: bad channel count");
    return GCC_ERR_ARG;
  }
  const int G = Cp / 8;
  const int groups = per_sample ? N : 1;
  const long long npix = per_sample ? HW : (long long)N * HW;
  if (zero_sums && cudaMemsetAsync(sums, 0, sizeof(float) * 2 * Cp * groups, st) != cudaSuccess) return GCC_ERR_CUDA;
  int lanes;
  const int threads = stats_threads(G, &lanes);
  long long bx = (npix + lanes * 8 - 1) / (lanes * 8);
  const long long cap = (148LL * 4 + groups - 1) / groups;  // one resident wave: few same-address reductions
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  gcc_launch(norm_stats_kernel, dim3((unsigned)bx, groups), threads, sizeof(float) * lanes * G * 16, st, 
      (const bf16*)x, npix, Cp, G, lanes, sums);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}

extern "C" int gcc_norm_apply_bf16(const void* x, void* y, int N, long long HW, int Cp, int C, int per_sample,
                                   const float* sums, const float* gamma, const float* beta, const float* alpha,
                                   float thr, float eps, float* running_mean, float* running_var, float momentum,
                                   int act, float slope, int gate_after, void* y2, int y2_Cp, int y2_coff, int act2,
                                   long long stat_count, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  NormArgs a;
  int rc = fill_args(a, x, N, HW, Cp, C, per_sample, sums, gamma, beta, alpha, thr, eps, act, slope, gate_after,
                     stat_count);
  if (rc) return rc;
  if (y2 != nullptr && ((y2_Cp % 8) || (y2_coff % 8))) {
    gcc_set_error(__FILE__, __LINE__, "norm_apply: second output window must be 8-channel aligned");
    return GCC_ERR_ARG;
  }
  const int groups = per_sample ? N : 1;
  int lanes;
  const int threads = stats_threads(a.G, &lanes);
  const int bx = lane_blocks(a.npix, lanes, groups, 4);
  if (y2 != nullptr)
    gcc_launch(norm_apply_kernel<true>, dim3(bx, groups), threads, 0, st, a, lanes, (bf16*)y, (bf16*)y2, y2_Cp, y2_coff, act2,
                                                                 running_mean, running_var, momentum);
  else
    gcc_launch(norm_apply_kernel<false>, dim3(bx, groups), threads, 0, st, a, lanes, (bf16*)y, nullptr, 0, 0, 0, running_mean,
                                                                  running_var, momentum);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}

extern "C" int gcc_norm_apply_eval_bf16(const void* x, void* y, int N, long long HW, int Cp, int C,
                                        const float* running_mean, const float* running_var, const float* gamma,
                                        const float* beta, const float* alpha, float thr, float eps, int act,
                                        float slope, void* y2, int y2_Cp, int y2_coff, int act2, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  NormArgs a;
  int rc = fill_args(a, x, N, HW, Cp, C, 0, nullptr, gamma, beta, alpha, thr, eps, act, slope, 0);
  if (rc) return rc;
  const long long nvec = (long long)N * HW * a.G;
  gcc_launch(norm_apply_eval_kernel, ew_blocks(nvec), 256, 0, st, a, running_mean, running_var, (long long)N * HW, (bf16*)y,
                                                          (bf16*)y2, y2_Cp, y2_coff, act2);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}

// red: fp32 workspace [N if per_sample else 1][2][Cp] (zeroed here unless phase has bit 2 set: `red` is then
// pre-zeroed scratch of the caller).  dx may be NULL (only parameter gradients wanted); dgamma/dbeta/dalpha may be NULL.
extern "C" int gcc_norm_bwd_bf16(const void* x, int N, long long HW, int Cp, int C, int per_sample, const float* sums,
                                 const float* gamma, const float* beta, const float* alpha, float thr, float eps,
                                 int act, float slope, int gate_after, const void* dy, int dy_Cp, int dy_coff,
                                 const void* dy2, int dy2_Cp, int dy2_coff, int act2, float* red, void* dx,
                                 float* dgamma, float* dbeta, float* dalpha, long long stat_count, int phase,
                                 const float* red_param, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  NormArgs a;
  int rc = fill_args(a, x, N, HW, Cp, C, per_sample, sums, gamma, beta, alpha, thr, eps, act, slope, gate_after,
                     stat_count);
  if (rc) return rc;
  if ((dy && ((dy_Cp % 8) || (dy_coff % 8))) || (dy2 && ((dy2_Cp % 8) || (dy2_coff % 8)))) {
    gcc_set_error(__FILE__, __LINE__, "norm_bwd: gradient windows must be 8-channel aligned");
    return GCC_ERR_ARG;
  }
  const int groups = per_sample ? N : 1;
  const bool prezeroed = (phase & 4) != 0;
  phase &= ~4;
  if (phase < 0 || phase > 2) {
    gcc_set_error(__FILE__, __LINE__, "norm_bwd: phase must be 0 (both), 1 (reduce) or 2 (apply)");
    return GCC_ERR_ARG;
  }
  if (red_param == nullptr) red_param = red;
  const bool need_red = ((sums != nullptr) || dgamma || dbeta || dalpha || phase == 1) && phase != 2;
  if (need_red) {
    if (!prezeroed && cudaMemsetAsync(red, 0, sizeof(float) * 2 * Cp * groups, st) != cudaSuccess) return GCC_ERR_CUDA;
    int lanes;
    const int threads = stats_threads(a.G, &lanes);
    long long bx = (a.npix + lanes * 8 - 1) / (lanes * 8);
    const long long cap = (148LL * 2 + groups - 1) / groups;  // one resident wave (launch bounds: 2 CTAs / SM)
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    static const int flat = getenv("GCC_B200_NORM_REDUCE_FLAT") ? atoi(getenv("GCC_B200_NORM_REDUCE_FLAT")) : 0;
    if (dy2 == nullptr && flat == 4)
      gcc_launch(norm_bwd_reduce_kernel<4, false, false>, dim3((unsigned)bx, groups), threads, sizeof(float) * lanes * a.G * 16, st, a,
                 lanes, (const bf16*)dy, dy_Cp, dy_coff, (const bf16*)dy2, dy2_Cp, dy2_coff, act2, red);
    else if (dy2 == nullptr && flat == 8)
      gcc_launch(norm_bwd_reduce_kernel<8, false, false>, dim3((unsigned)bx, groups), threads, sizeof(float) * lanes * a.G * 16, st, a,
                 lanes, (const bf16*)dy, dy_Cp, dy_coff, (const bf16*)dy2, dy2_Cp, dy2_coff, act2, red);
    else if (dy2 == nullptr)   // (4 pixels in flight spill in this kernel and measured 25 % slower)
      gcc_launch(norm_bwd_reduce_kernel<2, false>, dim3((unsigned)bx, groups), threads, sizeof(float) * lanes * a.G * 16, st, a,
                 lanes, (const bf16*)dy, dy_Cp, dy_coff, (const bf16*)dy2, dy2_Cp, dy2_coff, act2, red);
    else
      gcc_launch(norm_bwd_reduce_kernel<2, true>, dim3((unsigned)bx, groups), threads, sizeof(float) * lanes * a.G * 16, st, a,
                 lanes, (const bf16*)dy, dy_Cp, dy_coff, (const bf16*)dy2, dy2_Cp, dy2_coff, act2, red);
    GCC_CHECK_LAUNCH();
  }
  if (phase == 1) return GCC_OK;
  int alanes;
  const int athreads = stats_threads(a.G, &alanes);
  const int bx = dx ? lane_blocks(a.npix, alanes, groups, 2, 2) : 1;
  if (dy2 != nullptr)
    gcc_launch(norm_bwd_apply_kernel<true, 2>, dim3(bx, dx ? groups : 1), athreads, 0, st, 
        a, N, alanes, (const bf16*)dy, dy_Cp, dy_coff, (const bf16*)dy2, dy2_Cp, dy2_coff, act2, red, red_param,
        (bf16*)dx, dgamma, dbeta, dalpha);
  else
    gcc_launch(norm_bwd_apply_kernel<false, 4>, dim3(bx, dx ? groups : 1), athreads, 0, st, 
        a, N, alanes, (const bf16*)dy, dy_Cp, dy_coff, nullptr, 0, 0, 0, red, red_param, (bf16*)dx, dgamma, dbeta,
        dalpha);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
