// Elementwise / layout kernels: NCHW fp32 <-> NHWC bf16 boundary conversion, channel-window copies
// (torch.cat of the reference, models/Pix2Pix.py:77,467,471,516), activations, dropout, residual add,
// weight packing, bias gradients, reflection padding and the depthwise 3x3 convolution of the
// MobileResNet blocks (models/Pix2Pix.py:132-145).  All HBM-bound, vectorised where alignment allows.
#include "common.cuh"

namespace gcc {

__device__ __forceinline__ uint32_t hash_u32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU;
  x ^= x >> 15; x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}

// ---- boundary layout conversion -------------------------------------------------------------
// dst[n,h,w,c_off + c] = bf16(src[n,c,h,w]); channels [c_off + C, zero_to) are zeroed.
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int N, int C, long long HW,
                                    int Cp, int c_off, int zero_to) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const long long total = (long long)N * HW;
  // the <= 8-channel images of the step (one 16-byte pixel): the whole pixel is assembled in registers and written with
  // ONE vector store (the general path below writes eight 2-byte values per pixel)
  const bool pixel8 = (Cp == 8 && c_off == 0 && zero_to == 8 && C <= 8);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HW, p = i % HW;
    if (pixel8) {
      float v[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] = c < C ? src[(n * C + c) * HW + p] : 0.f;
      *reinterpret_cast<uint4*>(dst + i * 8) =
          make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
      continue;
    }
    bf16* d = dst + i * Cp;
    for (int c = 0; c < C; ++c) d[c_off + c] = __float2bfloat16(src[(n * C + c) * HW + p]);
    for (int c = c_off + C; c < zero_to; ++c) d[c] = __float2bfloat16(0.f);
  }
}
// dst[n,c,h,w] (+)= float(src[n,h,w,c_off + c])
__global__ void nhwc_to_nchw_kernel(const bf16* __restrict__ src, float* __restrict__ dst, int N, int C, long long HW,
                                    int Cp, int c_off, int accumulate) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const long long total = (long long)N * C * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long p = i % HW;
    const long long nc = i / HW;
    const int c = (int)(nc % C);
    const long long n = nc / C;
    const float v = __bfloat162float(src[(n * HW + p) * Cp + c_off + c]);
    if (accumulate) dst[i] += v;
    else dst[i] = v;
  }
}

// ---- channel window copy (concat / split) ---------------------------------------------------
// dst[p, d_off + c] (=|+=) src[p, s_off + c], c < C.  Vector path when everything is 8-aligned.
__global__ void copy_channels_vec_kernel(const bf16* __restrict__ src, int Cs, int s_off, bf16* __restrict__ dst,
                                         int Cd, int d_off, int G, long long npix, int accumulate) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const long long nvec = npix * G;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    const long long p = i / G;
    uint4 v = *reinterpret_cast<const uint4*>(src + p * Cs + s_off + g * 8);
    uint4* d = reinterpret_cast<uint4*>(dst + p * Cd + d_off + g * 8);
    if (accumulate) {
      const uint4 o = *d;
      v.x = pack_bf16(bf16_lo(v.x) + bf16_lo(o.x), bf16_hi(v.x) + bf16_hi(o.x));
      v.y = pack_bf16(bf16_lo(v.y) + bf16_lo(o.y), bf16_hi(v.y) + bf16_hi(o.y));
      v.z = pack_bf16(bf16_lo(v.z) + bf16_lo(o.z), bf16_hi(v.z) + bf16_hi(o.z));
      v.w = pack_bf16(bf16_lo(v.w) + bf16_lo(o.w), bf16_hi(v.w) + bf16_hi(o.w));
    }
    *d = v;
  }
}
__global__ void copy_channels_scalar_kernel(const bf16* __restrict__ src, int Cs, int s_off, bf16* __restrict__ dst,
                                            int Cd, int d_off, int C, long long npix, int accumulate) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const long long total = npix * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long p = i / C;
    const float v = __bfloat162float(src[p * Cs + s_off + c]);
    bf16* d = dst + p * Cd + d_off + c;
    *d = __float2bfloat16(accumulate ? v + __bfloat162float(*d) : v);
  }
}

// torch.cat of two <= 8-channel images whose channels fit one 16-byte pixel (pix2pix's cat(real_A, fake_B): 3 + 3):
// one vector load per input pixel, one vector store.  split = the backward (either output may be NULL).
__global__ void cat_small_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ y,
                                 int ca, int cb, long long npix) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix;
       i += (long long)gridDim.x * blockDim.x) {
    const uint4 ua = a[i], ub = b[i];
    const unsigned short* pa = reinterpret_cast<const unsigned short*>(&ua);
    const unsigned short* pb = reinterpret_cast<const unsigned short*>(&ub);
    uint4 o = make_uint4(0, 0, 0, 0);
    unsigned short* po = reinterpret_cast<unsigned short*>(&o);
#pragma unroll
    for (int k = 0; k < 8; ++k) po[k] = k < ca ? pa[k] : (k < ca + cb ? pb[k - ca] : (unsigned short)0);
    y[i] = o;
  }
}
__global__ void split_small_kernel(const uint4* __restrict__ dy, uint4* __restrict__ da, uint4* __restrict__ db, int ca,
                                   int cb, long long npix) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix;
       i += (long long)gridDim.x * blockDim.x) {
    const uint4 u = dy[i];
    const unsigned short* p = reinterpret_cast<const unsigned short*>(&u);
    if (da != nullptr) {
      uint4 o = make_uint4(0, 0, 0, 0);
      unsigned short* po = reinterpret_cast<unsigned short*>(&o);
#pragma unroll
      for (int k = 0; k < 8; ++k) po[k] = k < ca ? p[k] : (unsigned short)0;
      da[i] = o;
    }
    if (db != nullptr) {
      uint4 o = make_uint4(0, 0, 0, 0);
      unsigned short* po = reinterpret_cast<unsigned short*>(&o);
#pragma unroll
      for (int k = 0; k < 8; ++k) po[k] = (k < cb && k + ca < 8) ? p[k + ca] : (unsigned short)0;
      db[i] = o;
    }
  }
}

// ---- activations / dropout / add ------------------------------------------------------------
// mode 1 leaky-relu, 2 relu, 3 tanh
__global__ void act_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long nvec, int mode,
                               float slope) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const uint4 u = reinterpret_cast<const uint4*>(x)[i];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float a = bf16_lo(w[k]), b = bf16_hi(w[k]);
      if (mode == 1) { a = a > 0.f ? a : a * slope; b = b > 0.f ? b : b * slope; }
      else if (mode == 2) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
      else if (mode == 3) { a = tanhf(a); b = tanhf(b); }
      o[k] = pack_bf16(a, b);
    }
    reinterpret_cast<uint4*>(y)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
// dx = dy * f'(.) where ref is the forward INPUT for (leaky-)relu and the forward OUTPUT for tanh
__global__ void act_bwd_kernel(const bf16* __restrict__ ref, const bf16* __restrict__ dy, bf16* __restrict__ dx,
                               long long nvec, int mode, float slope) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const uint4 r = reinterpret_cast<const uint4*>(ref)[i];
    const uint4 d = reinterpret_cast<const uint4*>(dy)[i];
    const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
    const uint32_t dw[4] = {d.x, d.y, d.z, d.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float ra = bf16_lo(rw[k]), rb = bf16_hi(rw[k]);
      float da = bf16_lo(dw[k]), db = bf16_hi(dw[k]);
      if (mode == 1) { da *= ra > 0.f ? 1.f : slope; db *= rb > 0.f ? 1.f : slope; }
      else if (mode == 2) { da = ra > 0.f ? da : 0.f; db = rb > 0.f ? db : 0.f; }
      else if (mode == 3) { da *= 1.f - ra * ra; db *= 1.f - rb * rb; }
      o[k] = pack_bf16(da, db);
    }
    reinterpret_cast<uint4*>(dx)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
// Counter-based Bernoulli(keep = 1-p) dropout, scale 1/(1-p); the same (seed, index) regenerates the
// mask in backward (nn.Dropout(0.5), Pix2Pix.py:64).  seed is read from device memory so that a
// captured CUDA graph draws a fresh mask on every replay.
__global__ void dropout_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long n, float p,
                               const unsigned long long* __restrict__ seed_ptr, unsigned int salt) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const unsigned long long seed = *seed_ptr;
  const uint32_t s0 = hash_u32((uint32_t)seed ^ salt), s1 = hash_u32((uint32_t)(seed >> 32) + 0x9e3779b9U);
  const float scale = 1.f / (1.f - p);
  const uint32_t cut = (uint32_t)(p * 4294967296.0);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const uint32_t r = hash_u32((uint32_t)i * 0x9e3779b1U + s0) ^ hash_u32((uint32_t)(i >> 32) + s1);
    const float v = __bfloat162float(x[i]);
    y[i] = __float2bfloat16(r >= cut ? v * scale : 0.f);
  }
}
__global__ void add_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ y,
                           long long nvec) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const uint4 u = reinterpret_cast<const uint4*>(a)[i];
    const uint4 v = reinterpret_cast<const uint4*>(b)[i];
    uint4 o;
    o.x = pack_bf16(bf16_lo(u.x) + bf16_lo(v.x), bf16_hi(u.x) + bf16_hi(v.x));
    o.y = pack_bf16(bf16_lo(u.y) + bf16_lo(v.y), bf16_hi(u.y) + bf16_hi(v.y));
    o.z = pack_bf16(bf16_lo(u.z) + bf16_lo(v.z), bf16_hi(u.z) + bf16_hi(v.z));
    o.w = pack_bf16(bf16_lo(u.w) + bf16_lo(v.w), bf16_hi(u.w) + bf16_hi(v.w));
    reinterpret_cast<uint4*>(y)[i] = o;
  }
}

// ---- weight packing --------------------------------------------------------------------------
// src fp32 [D0][T][D1] (channels_last storage of an OIHW / IOHW parameter)
//   direct     bf16 [D0][T][D1p]      transposed bf16 [D1][T][D0p]    (pads stay zero)
__global__ void pack_weight_kernel(const float* __restrict__ src, bf16* __restrict__ direct,
                                   bf16* __restrict__ transposed, int D0, int T, int D1, int D1p, int D0p) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const long long total = (long long)D0 * T * D1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int d1 = (int)(i % D1);
    const long long r = i / D1;
    const int t = (int)(r % T);
    const int d0 = (int)(r / T);
    const bf16 v = __float2bfloat16(src[i]);
    if (direct) direct[((long long)d0 * T + t) * D1p + d1] = v;
    if (transposed) transposed[((long long)d1 * T + t) * D0p + d0] = v;
  }
}

// ---- bias gradient: db[c] = sum over pixels of dy[p, c_off + c] --------------------------------
// thread = (channel group of 8, pixel lane); 16-byte loads; shared-memory reduce over lanes; one atomic per
// (block, channel).  Requires Cp % 8 == 0 and c_off % 8 == 0 (activations always satisfy this).
__global__ void colsum_vec_kernel(const bf16* __restrict__ dy, long long npix, int Cp, int c_off, int C, int G,
                                  int lanes, float* __restrict__ out) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  extern __shared__ float red[];  // [lanes][G][8]
  const int tid = threadIdx.x;
  const int g = tid % G, lane = tid / G;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (lane < lanes) {
    for (long long p = (long long)blockIdx.x * lanes + lane; p < npix; p += (long long)gridDim.x * lanes) {
      const uint4 u = *reinterpret_cast<const uint4*>(dy + p * Cp + c_off + g * 8);
      s[0] += bf16_lo(u.x); s[1] += bf16_hi(u.x);
      s[2] += bf16_lo(u.y); s[3] += bf16_hi(u.y);
      s[4] += bf16_lo(u.z); s[5] += bf16_hi(u.z);
      s[6] += bf16_lo(u.w); s[7] += bf16_hi(u.w);
    }
    float* r = red + ((long long)lane * G + g) * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = s[i];
  }
  __syncthreads();
  for (int e = tid; e < G * 8; e += blockDim.x) {
    if (e >= C) continue;
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += red[(long long)l * G * 8 + e];
    atomicAdd(out + e, acc);
  }
}

// ---- reflection padding (nn.ReflectionPad2d) fwd / bwd --------------------------------------
__device__ __forceinline__ int reflect(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}
__global__ void reflect_pad_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int N, int H, int W, int G,
                                   int pad) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const int OH = H + 2 * pad, OW = W + 2 * pad;
  const long long nvec = (long long)N * OH * OW * G;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    long long t = i / G;
    const int ow = (int)(t % OW); t /= OW;
    const int oh = (int)(t % OH);
    const long long n = t / OH;
    const int ih = reflect(oh - pad, H), iw = reflect(ow - pad, W);
    reinterpret_cast<uint4*>(y)[i] = reinterpret_cast<const uint4*>(x)[((n * H + ih) * W + iw) * G + g];
  }
}
// dx[n,ih,iw] = sum of dy over all padded positions that mirror onto (ih, iw)
__global__ void reflect_pad_bwd_kernel(const bf16* __restrict__ dy, bf16* __restrict__ dx, int N, int H, int W, int G,
                                       int pad) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const int OH = H + 2 * pad, OW = W + 2 * pad;
  const long long nvec = (long long)N * H * W * G;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    long long t = i / G;
    const int iw = (int)(t % W); t /= W;
    const int ih = (int)(t % H);
    const long long n = t / H;
    int hs[3], ws[3], nh = 0, nw = 0;
    hs[nh++] = ih + pad;
    if (ih >= 1 && ih <= pad) hs[nh++] = pad - ih;
    if (ih <= H - 2 && ih >= H - 1 - pad) hs[nh++] = pad + 2 * (H - 1) - ih;
    ws[nw++] = iw + pad;
    if (iw >= 1 && iw <= pad) ws[nw++] = pad - iw;
    if (iw <= W - 2 && iw >= W - 1 - pad) ws[nw++] = pad + 2 * (W - 1) - iw;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int a = 0; a < nh; ++a)
      for (int b = 0; b < nw; ++b) {
        const uint4 u = reinterpret_cast<const uint4*>(dy)[((n * OH + hs[a]) * OW + ws[b]) * G + g];
        acc[0] += bf16_lo(u.x); acc[1] += bf16_hi(u.x);
        acc[2] += bf16_lo(u.y); acc[3] += bf16_hi(u.y);
        acc[4] += bf16_lo(u.z); acc[5] += bf16_hi(u.z);
        acc[6] += bf16_lo(u.w); acc[7] += bf16_hi(u.w);
      }
    reinterpret_cast<uint4*>(dx)[i] = make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]),
                                                 pack_bf16(acc[4], acc[5]), pack_bf16(acc[6], acc[7]));
  }
}

// ---- depthwise 3x3 (groups = C) with fused reflection padding 1 -------------------------------
// w: fp32 [C][9], bias fp32 [C].  y[n,h,w,c] = b[c] + sum_k w[c][k] * x[n, refl(h+kh-1), refl(w+kw-1), c]
__global__ void dw3x3_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ w,
                                 const float* __restrict__ bias, bf16* __restrict__ y, int N, int H, int W, int G,
                                 int C) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const long long nvec = (long long)N * H * W * G;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    long long t = i / G;
    const int ow = (int)(t % W); t /= W;
    const int oh = (int)(t % H);
    const long long n = t / H;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = (bias && g * 8 + k < C) ? bias[g * 8 + k] : 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = reflect(oh + kh - 1, H);
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int iw = reflect(ow + kw - 1, W);
        const uint4 u = reinterpret_cast<const uint4*>(x)[((n * H + ih) * W + iw) * G + g];
        const float xv[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                             bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int c = g * 8 + k;
          if (c < C) acc[k] += xv[k] * w[c * 9 + kh * 3 + kw];
        }
      }
    }
    reinterpret_cast<uint4*>(y)[i] = make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]),
                                                pack_bf16(acc[4], acc[5]), pack_bf16(acc[6], acc[7]));
  }
}
// Row-sliding depthwise 3x3: thread = (channel group of 8, pixel lane); an item is a segment of one output row.  The
// 72 tap weights of the thread's 8 channels live in registers and a 3 x 3 window of 16-byte vectors slides along the
// row, so every output costs 3 new loads (instead of 9 loads + 72 weight loads in the per-pixel kernels above).
//   MODE 0: forward with the fused ReflectionPad2d(1):  y = b + sum_k w[k] x[refl(h+kh-1), refl(w+kw-1)]
//   MODE 1: data gradient of the same: the zero-padded correlation of dy with the flipped taps, where the rows /
//           columns next to the border also collect the mirrored taps (dx[1] += w[kh=0] dy[0], dx[H-2] += w[kh=2]
//           dy[H-1], same for columns) -- the fold of reflect_pad_bwd applied to the tap weights instead of the data.
template <int MODE>
__global__ void __launch_bounds__(256)
dw3x3_rows_kernel(const bf16* __restrict__ src, const float* __restrict__ w, const float* __restrict__ bias,
                  bf16* __restrict__ dst, int H, int W, int G, int C, int lanes, int seg, long long items) {
  pdl_wait();
  pdl_launch_dependents();
  const int g = threadIdx.x % G, lane = threadIdx.x / G;
  if (lane >= lanes) return;
  float we[3][3][8];  // [row offset][column offset][channel]
  float bs[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) bs[k] = (MODE == 0 && bias != nullptr && g * 8 + k < C) ? bias[g * 8 + k] : 0.f;
  if (MODE == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int k = 0; k < 8; ++k) we[a][b][k] = (g * 8 + k < C) ? w[(g * 8 + k) * 9 + a * 3 + b] : 0.f;
  }
  const int segs = (W + seg - 1) / seg;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  const uint4* sv = reinterpret_cast<const uint4*>(src);
  uint4* dv = reinterpret_cast<uint4*>(dst);
  for (long long it = (long long)blockIdx.x * lanes + lane; it < items; it += (long long)gridDim.x * lanes) {
    const int sgi = (int)(it % segs);
    const long long t = it / segs;
    const int r = (int)(t % H);
    const long long n = t / H;
    const int w0 = sgi * seg, w1 = min(W, w0 + seg);
    long long rowoff[3];
    bool rv[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      int rr = r + a - 1;
      if (MODE == 0) {
        rr = reflect(rr, H);
        rv[a] = true;
      } else {
        rv[a] = rr >= 0 && rr < H;
      }
      rowoff[a] = ((n * H + (rv[a] ? rr : 0)) * (long long)W) * G + g;
    }
    if (MODE == 1) {
      // we[a][b] = w[kh = 2 - a][kw = 2 - b] (+ the mirrored tap on the rows next to the border)
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int c = g * 8 + k;
            float v = 0.f;
            if (c < C) {
              v = __ldg(w + c * 9 + (2 - a) * 3 + (2 - b));
              if (a == 0 && r == 1) v += __ldg(w + c * 9 + 0 * 3 + (2 - b));
              if (a == 2 && r == H - 2) v += __ldg(w + c * 9 + 2 * 3 + (2 - b));
            }
            we[a][b][k] = v;
          }
    }
    auto ld = [&](int a, int col) -> uint4 {
      if (MODE == 0) return sv[rowoff[a] + (long long)reflect(col, W) * G];
      return (rv[a] && col >= 0 && col < W) ? sv[rowoff[a] + (long long)col * G] : zero;
    };
    uint4 win[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      win[a][0] = ld(a, w0 - 1);
      win[a][1] = ld(a, w0);
    }
    for (int ow = w0; ow < w1; ++ow) {
#pragma unroll
      for (int a = 0; a < 3; ++a) win[a][2] = ld(a, ow + 1);
      float acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = bs[k];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          const uint4 u = win[a][b];
          const float xv[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                               bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] += xv[k] * we[a][b][k];
        }
      if (MODE == 1 && (ow == 1 || ow == W - 2)) {
        // mirrored column taps: dx[.., 1] += w[.., kw = 0] dy[.., 0];  dx[.., W - 2] += w[.., kw = 2] dy[.., W - 1]
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          if (ow == 1) {
            const uint4 u = win[a][0];
            const float xv[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                                 bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += xv[k] * we[a][2][k];
          }
          if (ow == W - 2) {
            const uint4 u = win[a][2];
            const float xv[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                                 bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += xv[k] * we[a][0][k];
          }
        }
      }
      dv[((n * H + r) * (long long)W + ow) * G + g] =
          make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]), pack_bf16(acc[4], acc[5]),
                     pack_bf16(acc[6], acc[7]));
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        win[a][0] = win[a][1];
        win[a][1] = win[a][2];
      }
    }
  }
}

// data gradient: scatter form of the reflected gather = gather over the (<= 3x3 x mirror) sources.
// Implemented as: dxp = zero-padded correlation on the reflect-PADDED grid, then folded by
// reflect_pad_bwd.  Here: dyp-style direct accumulation over taps with explicit mirror bookkeeping
// is avoided by computing on the padded grid: dxp[n, ph, pw, c] = sum_k w[c][k] * dy[n, ph-kh, pw-kw, c].
__global__ void dw3x3_bwd_data_padded_kernel(const bf16* __restrict__ dy, const float* __restrict__ w,
                                             bf16* __restrict__ dxp, int N, int H, int W, int G, int C) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const int PH = H + 2, PW = W + 2;
  const long long nvec = (long long)N * PH * PW * G;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    long long t = i / G;
    const int pw = (int)(t % PW); t /= PW;
    const int ph = (int)(t % PH);
    const long long n = t / PH;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int oh = ph - kh;
      if (oh < 0 || oh >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int ow = pw - kw;
        if (ow < 0 || ow >= W) continue;
        const uint4 u = reinterpret_cast<const uint4*>(dy)[((n * H + oh) * W + ow) * G + g];
        const float dv[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                             bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int c = g * 8 + k;
          if (c < C) acc[k] += dv[k] * w[c * 9 + kh * 3 + kw];
        }
      }
    }
    reinterpret_cast<uint4*>(dxp)[i] = make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]),
                                                  pack_bf16(acc[4], acc[5]), pack_bf16(acc[6], acc[7]));
  }
}
// Row-sliding weight / bias gradient of the depthwise conv (same thread layout and items as dw3x3_rows_kernel): the
// 3 x 3 window of x slides along the row (3 new loads + one dy load per pixel), 72 + 8 accumulators per thread, one
// shared-memory reduction over the CTA's pixel lanes and one atomic per (CTA, channel, tap).
__global__ void __launch_bounds__(256)
dw3x3_wgrad_rows_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, float* __restrict__ dw,
                        float* __restrict__ dbias, int H, int W, int G, int C, int lanes, int seg, long long items) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ float sred[];  // [lanes][G][80]
  const int g = threadIdx.x % G, lane = threadIdx.x / G;
  float acc[3][3][8], accb[8];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[a][b][k] = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) accb[k] = 0.f;
  if (lane < lanes) {
    const int segs = (W + seg - 1) / seg;
    const uint4* xv = reinterpret_cast<const uint4*>(x);
    const uint4* dv = reinterpret_cast<const uint4*>(dy);
    for (long long it = (long long)blockIdx.x * lanes + lane; it < items; it += (long long)gridDim.x * lanes) {
      const int sgi = (int)(it % segs);
      const long long t = it / segs;
      const int r = (int)(t % H);
      const long long n = t / H;
      const int w0 = sgi * seg, w1 = min(W, w0 + seg);
      long long rowoff[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) rowoff[a] = ((n * H + reflect(r + a - 1, H)) * (long long)W) * G + g;
      const long long dyoff = ((n * H + r) * (long long)W) * G + g;
      uint4 win[3][3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        win[a][0] = xv[rowoff[a] + (long long)reflect(w0 - 1, W) * G];
        win[a][1] = xv[rowoff[a] + (long long)w0 * G];
      }
      for (int ow = w0; ow < w1; ++ow) {
#pragma unroll
        for (int a = 0; a < 3; ++a) win[a][2] = xv[rowoff[a] + (long long)reflect(ow + 1, W) * G];
        const uint4 du = dv[dyoff + (long long)ow * G];
        const float d8[8] = {bf16_lo(du.x), bf16_hi(du.x), bf16_lo(du.y), bf16_hi(du.y),
                             bf16_lo(du.z), bf16_hi(du.z), bf16_lo(du.w), bf16_hi(du.w)};
#pragma unroll
        for (int k = 0; k < 8; ++k) accb[k] += d8[k];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) {
            const uint4 u = win[a][b];
            const float x8[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                                 bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[a][b][k] += d8[k] * x8[k];
          }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          win[a][0] = win[a][1];
          win[a][1] = win[a][2];
        }
      }
    }
    float* rr = sred + ((long long)lane * G + g) * 80;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) rr[k * 10 + a * 3 + b] = acc[a][b][k];
      rr[k * 10 + 9] = accb[k];
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < G * 80; e += blockDim.x) {
    const int gg = e / 80, kj = e % 80;
    const int k = kj / 10, j = kj % 10;
    const int c = gg * 8 + k;
    if (c >= C) continue;
    float a = 0.f;
    for (int l = 0; l < lanes; ++l) a += sred[((long long)l * G + gg) * 80 + kj];
    if (j < 9) atomicAdd(dw + c * 9 + j, a);
    else if (dbias) atomicAdd(dbias + c, a);
  }
}

// weight / bias gradient: dw[c][k] = sum_{n,h,w} dy[n,h,w,c] * x[n, refl(h+kh-1), refl(w+kw-1), c]
// block = (G groups) x lanes; per-thread 8 channels x 10 accumulators; smem reduce; atomics.
__global__ void dw3x3_bwd_weight_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                        float* __restrict__ dw, float* __restrict__ dbias, int N, int H, int W, int G,
                                        int C, int lanes) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  extern __shared__ float sred[];  // [lanes][G][80]
  const int tid = threadIdx.x;
  const int g = tid % G, lane = tid / G;
  float acc[8][10];
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int j = 0; j < 10; ++j) acc[k][j] = 0.f;
  const long long npix = (long long)N * H * W;
  if (lane < lanes) {
    for (long long p = (long long)blockIdx.x * lanes + lane; p < npix; p += (long long)gridDim.x * lanes) {
      const int ow = (int)(p % W);
      const int oh = (int)((p / W) % H);
      const long long n = p / ((long long)W * H);
      const uint4 du = reinterpret_cast<const uint4*>(dy)[p * G + g];
      const float dv[8] = {bf16_lo(du.x), bf16_hi(du.x), bf16_lo(du.y), bf16_hi(du.y),
                           bf16_lo(du.z), bf16_hi(du.z), bf16_lo(du.w), bf16_hi(du.w)};
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k][9] += dv[k];
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int ih = reflect(oh + kh - 1, H);
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int iw = reflect(ow + kw - 1, W);
          const uint4 u = reinterpret_cast<const uint4*>(x)[((n * H + ih) * W + iw) * G + g];
          const float xv[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                               bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k][kh * 3 + kw] += dv[k] * xv[k];
        }
      }
    }
    float* r = sred + ((long long)lane * G + g) * 80;
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int j = 0; j < 10; ++j) r[k * 10 + j] = acc[k][j];
  }
  __syncthreads();
  for (int e = tid; e < G * 80; e += blockDim.x) {
    const int gg = e / 80, kj = e % 80;
    const int k = kj / 10, j = kj % 10;
    const int c = gg * 8 + k;
    if (c >= C) continue;
    float a = 0.f;
    for (int l = 0; l < lanes; ++l) a += sred[((long long)l * G + gg) * 80 + kj];
    if (j < 9) atomicAdd(dw + c * 9 + j, a);
    else if (dbias) atomicAdd(dbias + c, a);
  }
}

// ---- im2col / col2im for k4 s2 p1 layers whose image side has <= 8 channels (Cp == 8) -----------
// These layers (PatchGAN / U-Net first conv, U-Net last ConvTranspose) have K or N = 3..6 channels: as an
// implicit GEMM every tap would be a 64-wide k-block that is 7/8 zeros.  Instead the 16 taps x 8 channels
// are laid out as ONE 128-wide K (or N) so the same tcgen05 kernels run them as 1x1 convs.
// col[n,oh,ow,(kh*4+kw)*8 + c] = img[n, 2*oh+kh-1, 2*ow+kw-1, c]   (zero outside the image)
__global__ void im2col_k4s2_c8_kernel(const uint4* __restrict__ img, uint4* __restrict__ col, int N, int H, int W) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const int OH = H / 2, OW = W / 2;
  const long long total = (long long)N * OH * OW * 16;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i & 15);
    long long t = i >> 4;
    const int ow = (int)(t % OW); t /= OW;
    const int oh = (int)(t % OH);
    const long long n = t / OH;
    const int ih = 2 * oh + (tap >> 2) - 1, iw = 2 * ow + (tap & 3) - 1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = img[(n * H + ih) * W + iw];
    col[i] = v;
  }
}
// img[n,iy,ix,c] = act(bias[c] + sum_{kh,kw} col[n,(iy+1-kh)/2,(ix+1-kw)/2, idx(kh,kw,c)]) over the taps whose
// source index is integral and inside the col grid.  order 0: idx = (kh*4+kw)*8+c, order 1: idx = c*16+kh*4+kw.
// idx_t = unsigned when N*H*W < 2^31 (always, in practice): the three divisions per pixel are then 32-bit.
template <typename idx_t>
__global__ void col2im_k4s2_c8_kernel(const bf16* __restrict__ col, int Ccol, int order, int C,
                                      const float* __restrict__ bias, int act, uint4* __restrict__ img, int N, int H,
                                      int W) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const int OH = H / 2, OW = W / 2;
  const idx_t total = (idx_t)N * (idx_t)H * (idx_t)W;
  for (idx_t i = (idx_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (idx_t)gridDim.x * blockDim.x) {
    const int ix = (int)(i % (idx_t)W);
    idx_t t = i / (idx_t)W;
    const int iy = (int)(t % (idx_t)H);
    const long long n = (long long)(t / (idx_t)H);
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = (bias != nullptr && c < C) ? bias[c] : 0.f;
    const int kh0 = (iy + 1) & 1, kw0 = (ix + 1) & 1;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int kh = kh0 + 2 * a;
      const int oh = (iy + 1 - kh) / 2;
      if (iy + 1 - kh < 0 || oh >= OH) continue;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int kw = kw0 + 2 * b;
        const int ow = (ix + 1 - kw) / 2;
        if (ix + 1 - kw < 0 || ow >= OW) continue;
        const bf16* row = col + ((n * OH + oh) * OW + ow) * (long long)Ccol;
        const int tap = kh * 4 + kw;
        if (order == 0) {
          const uint4 u = *reinterpret_cast<const uint4*>(row + tap * 8);
          acc[0] += bf16_lo(u.x); acc[1] += bf16_hi(u.x);
          acc[2] += bf16_lo(u.y); acc[3] += bf16_hi(u.y);
          acc[4] += bf16_lo(u.z); acc[5] += bf16_hi(u.z);
          acc[6] += bf16_lo(u.w); acc[7] += bf16_hi(u.w);
        } else if (order == 2) {  // tap-major, 4 channels per tap: one 8-byte load
          const uint2 u = *reinterpret_cast<const uint2*>(row + tap * 4);
          acc[0] += bf16_lo(u.x); acc[1] += bf16_hi(u.x);
          acc[2] += bf16_lo(u.y); acc[3] += bf16_hi(u.y);
        } else {
          for (int c = 0; c < C; ++c) acc[c] += __bfloat162float(row[c * 16 + tap]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c >= C) acc[c] = 0.f;
      else if (act == 2) acc[c] = tanhf(acc[c]);
    }
    img[i] = make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]), pack_bf16(acc[4], acc[5]),
                        pack_bf16(acc[6], acc[7]));
  }
}
// g[r][tap][c] += tmp[r][tap*8 + c]   (c < C <= 8): weight gradient of a col-path layer back to [R][16][C]
__global__ void unpad_wgrad_c8_kernel(const float* __restrict__ tmp, float* __restrict__ g, int R, int C) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const int total = R * 16 * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % C;
    const int rt = i / C;
    g[i] += tmp[(long long)rt * 8 + c];
  }
}

// ---- k4 s1 p1 Conv2d with <= 8 OUTPUT channels (PatchGAN's 1-channel logits head, Pix2Pix.py:300,343) --------
// fprop: ycol[pix_in, co*16+tap] = x[pix_in,:] . w[co,tap,:] is one 1x1 GEMM that reads x ONCE (instead of once
// per tap); the 16 partial dot products are then folded:  y[n,oy,ox,co] = b[co] + sum_taps ycol[n,oy+kh-1,ox+kw-1,.]
__global__ void fold_k4s1_kernel(const bf16* __restrict__ ycol, int Ccol, int C, const float* __restrict__ bias,
                                 uint4* __restrict__ y, int N, int H, int W) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const int OH = H - 1, OW = W - 1;  // k4 s1 p1
  const long long total = (long long)N * OH * OW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % OW);
    long long t = i / OW;
    const int oy = (int)(t % OH);
    const long long n = t / OH;
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = (bias != nullptr && c < C) ? bias[c] : 0.f;
    for (int kh = 0; kh < 4; ++kh) {
      const int iy = oy + kh - 1;
      if (iy < 0 || iy >= H) continue;
      for (int kw = 0; kw < 4; ++kw) {
        const int ix = ox + kw - 1;
        if (ix < 0 || ix >= W) continue;
        const bf16* row = ycol + ((n * H + iy) * W + ix) * (long long)Ccol;
        for (int c = 0; c < C; ++c) acc[c] += __bfloat162float(row[c * 16 + kh * 4 + kw]);
      }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c >= C) acc[c] = 0.f;
    y[i] = make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]), pack_bf16(acc[4], acc[5]),
                      pack_bf16(acc[6], acc[7]));
  }
}
// backward expansion: dcol[n,iy,ix,tap*8+c] = dy[n, iy-kh+1, ix-kw+1, c]  (zero outside), dy [N,H-1,W-1,8]
__global__ void unfold_k4s1_kernel(const uint4* __restrict__ dy, uint4* __restrict__ dcol, int N, int H, int W) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const int OH = H - 1, OW = W - 1;
  const long long total = (long long)N * H * W * 16;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i & 15);
    long long t = i >> 4;
    const int ix = (int)(t % W); t /= W;
    const int iy = (int)(t % H);
    const long long n = t / H;
    const int oy = iy - (tap >> 2) + 1, ox = ix - (tap & 3) + 1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (oy >= 0 && oy < OH && ox >= 0 && ox < OW) v = dy[(n * OH + oy) * OW + ox];
    dcol[i] = v;
  }
}
// g[c][tap][k] += tmp[tap*8 + c][k]   (c < C): weight gradient of the head back to [C][16][K]
__global__ void unpad_wgrad_rows_kernel(const float* __restrict__ tmp, float* __restrict__ g, int C, int K) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const long long total = (long long)C * 16 * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const int ct = (int)(i / K);
    const int tap = ct % 16, c = ct / 16;
    g[i] += tmp[(long long)(tap * 8 + c) * K + k];
  }
}

static inline int blocks_for(long long n, int per = 256) {
  long long b = (n + per - 1) / per;
  const long long cap = 148LL * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ---------------------------------------------------------------------------------------- image pool
// ImagePool.query (utils/image_pool.py:22-54) with the pool, its fill count and the random stream resident on the
// device, so that a captured CUDA graph replays the query with fresh decisions.  state[0] = images stored so far,
// state[1] = random counter.  One thread decides sequentially for the b images of the batch (the reference loops
// over them in order: a later image may swap out an earlier one of the same batch):
//   dec[2i]   = where out[i] comes from: -1 own image, s >= 0 pool slot s (content before this call),
//               -(2 + j) image j of this batch (it was stored into the chosen slot earlier in the loop)
//   dec[2i+1] = pool slot that receives image i at the end of the call, or -1
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
__global__ void image_pool_decide_kernel(long long* state, int pool_size, int b, int* dec) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  long long count = state[0];
  unsigned long long ctr = (unsigned long long)state[1];
  for (int i = 0; i < b; ++i) {
    int src = -1, dst = -1;
    if (count < pool_size) {
      dst = (int)count++;
    } else {
      const unsigned long long r = splitmix64(ctr++);
      if ((r >> 11) * (1.0 / 9007199254740992.0) > 0.5) {
        const int slot = (int)(splitmix64(ctr++) % (unsigned long long)pool_size);
        src = slot;
        for (int j = i - 1; j >= 0; --j)       // the slot may hold an image of this very batch
          if (dec[2 * j + 1] == slot) { src = -(2 + j); dec[2 * j + 1] = -1; break; }
        dst = slot;
      }
    }
    dec[2 * i] = src;
    dec[2 * i + 1] = dst;
  }
  state[0] = count;
  state[1] = (long long)ctr;
}
__global__ void image_pool_gather_kernel(const uint4* __restrict__ images, const uint4* __restrict__ pool,
                                         const int* __restrict__ dec, uint4* __restrict__ out, long long vec_per_image) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const int i = blockIdx.y;
  const int src = dec[2 * i];
  const uint4* s = src == -1 ? images + (long long)i * vec_per_image
                             : (src >= 0 ? pool + (long long)src * vec_per_image
                                         : images + (long long)(-(src + 2)) * vec_per_image);
  uint4* o = out + (long long)i * vec_per_image;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < vec_per_image;
       v += (long long)gridDim.x * blockDim.x)
    o[v] = s[v];
}
__global__ void image_pool_scatter_kernel(const uint4* __restrict__ images, uint4* __restrict__ pool,
                                          const int* __restrict__ dec, long long vec_per_image) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const int i = blockIdx.y;
  const int dst = dec[2 * i + 1];
  if (dst < 0) return;
  const uint4* s = images + (long long)i * vec_per_image;
  uint4* o = pool + (long long)dst * vec_per_image;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < vec_per_image;
       v += (long long)gridDim.x * blockDim.x)
    o[v] = s[v];
}

// ---- stem / head convolutions with <= 4 (8) channels on the image side and a large kernel (k7 of the MobileResNet
// generator, Pix2Pix.py:216,259; k9 of the SRResNet, SRGAN.py:150,190; k3 of VGG's first conv) ---------------------
// As implicit GEMMs these layers waste the tensor cores: the stem has K = 3 channels per tap (a 64-wide k-block per tap
// is 95 % zeros), the head has N = 3 output columns.  Instead:
//   stem ("row window"):  the taps of one kernel ROW are contiguous in an 8-channel NHWC image (kw pixels x 8
//        channels = 16 kw bytes), so a TMA tensor map with a 16-byte stride between window positions delivers, per
//        kernel row, a K-major [pixels][8 kw x 8 c] operand tile; K = KH x ceil(KW / 8) x 64 (conv_gemm.cu, rowwin).
//   head ("fold"):        one 1x1 GEMM over the (padded) input computes every tap's partial dot product,
//        ycol[pix][(tap, c)] = x[pix, :] . w[c][tap][:], and fold sums the KH*KW shifted partials per output pixel.
// Column index of a (tap, channel) pair: tap * CG + c with CG = 4 or 8 channels per group.

// y[n,oy,ox,c] = act(bias[c] + sum_{kh,kw} ycol[n, oy + dir*kh + off, ox + dir*kw + off, (kh*KW+kw)*CG + c])
// (terms outside the ycol grid are skipped).  dir = +1, off = 0: forward of a stride-1 conv over a pre-padded input;
// dir = -1, off = 0 with a ycol grid of the OUTPUT size: its data gradient on the padded input grid.
template <int CG>
__global__ void fold_taps_kernel(const bf16* __restrict__ ycol, int Ccol, int KH, int KW, int C, const float* __restrict__ bias,
                                 int act, uint4* __restrict__ y, int N, int GH, int GW, int OH, int OW, int dir, int off) {
  pdl_wait();
  pdl_launch_dependents();
  const long long total = (long long)N * OH * OW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % OW);
    long long t = i / OW;
    const int oy = (int)(t % OH);
    const long long n = t / OH;
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = (bias != nullptr && c < C) ? bias[c] : 0.f;
    for (int kh = 0; kh < KH; ++kh) {
      const int gy = oy + dir * kh + off;
      if (gy < 0 || gy >= GH) continue;
      const bf16* rowp = ycol + ((n * GH + gy) * (long long)GW) * Ccol + kh * KW * CG;
      for (int kw = 0; kw < KW; ++kw) {
        const int gx = ox + dir * kw + off;
        if (gx < 0 || gx >= GW) continue;
        const bf16* p = rowp + (long long)gx * Ccol + kw * CG;
        if (CG == 4) {
          const uint2 u = *reinterpret_cast<const uint2*>(p);
          acc[0] += bf16_lo(u.x); acc[1] += bf16_hi(u.x); acc[2] += bf16_lo(u.y); acc[3] += bf16_hi(u.y);
        } else {
          const uint4 u = *reinterpret_cast<const uint4*>(p);
          acc[0] += bf16_lo(u.x); acc[1] += bf16_hi(u.x); acc[2] += bf16_lo(u.y); acc[3] += bf16_hi(u.y);
          acc[4] += bf16_lo(u.z); acc[5] += bf16_hi(u.z); acc[6] += bf16_lo(u.w); acc[7] += bf16_hi(u.w);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = (c < C) ? (act == 2 ? tanhf(acc[c]) : acc[c]) : 0.f;
    y[i] = make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]), pack_bf16(acc[4], acc[5]),
                      pack_bf16(acc[6], acc[7]));
  }
}
// dcol[n,gy,gx,(kh*KW+kw)*CG + c] = dy[n, gy - kh - off, gx - kw - off, c] (zero outside the dy grid and in the column padding):
// the operand of the head conv's data- and weight-gradient GEMMs.  One thread per (pixel, 16-byte column chunk).
template <int CG>
__global__ void unfold_taps_kernel(const bf16* __restrict__ dy, bf16* __restrict__ dcol, int Ccol, int KH, int KW, int N,
                                   int GH, int GW, int OH, int OW, int off) {
  pdl_wait();
  pdl_launch_dependents();
  const int chunks = Ccol / 8;
  const long long total = (long long)N * GH * GW * chunks;
  constexpr int TPC = 8 / CG;  // taps per 16-byte chunk
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % chunks);
    long long t = i / chunks;
    const int gx = (int)(t % GW); t /= GW;
    const int gy = (int)(t % GH);
    const long long n = t / GH;
    uint32_t o[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < TPC; ++j) {
      const int tap = ch * TPC + j;
      if (tap >= KH * KW) break;
      const int oy = gy - tap / KW - off, ox = gx - tap % KW - off;
      if (oy < 0 || oy >= OH || ox < 0 || ox >= OW) continue;
      const bf16* p = dy + ((n * OH + oy) * (long long)OW + ox) * 8;
      if (CG == 4) {
        const uint2 u = *reinterpret_cast<const uint2*>(p);
        o[2 * j] = u.x;
        o[2 * j + 1] = u.y;
      } else {
        const uint4 u = *reinterpret_cast<const uint4*>(p);
        o[0] = u.x; o[1] = u.y; o[2] = u.z; o[3] = u.w;
      }
    }
    reinterpret_cast<uint4*>(dcol)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
// Weight operands of the fold path from the arena's bf16 packs:
//   mode 0: out[(t*CG + c)][k]   = direct[c][t][k]       (c < C, k < Kp)     rows = round8(T*CG), pitch Kp
//   mode 1: out[k][t*CG + c]     = transposed[k][t][c]   (c < C)             rows = K, pitch round8(T*CG); src pitch Cp
__global__ void fold_weight_pack_kernel(const bf16* __restrict__ src, bf16* __restrict__ out, int mode, int C, int T, int CG,
                                        int K, int Kp, int Cp, int rows_p) {
  pdl_wait();
  pdl_launch_dependents();
  if (mode == 0) {
    const long long total = (long long)rows_p * Kp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
      const int k = (int)(i % Kp);
      const int row = (int)(i / Kp);
      const int t = row / CG, c = row % CG;
      out[i] = (t < T && c < C) ? src[((long long)c * T + t) * Kp + k] : __float2bfloat16(0.f);
    }
  } else {
    const long long total = (long long)K * rows_p;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
      const int col = (int)(i % rows_p);
      const int k = (int)(i / rows_p);
      const int t = col / CG, c = col % CG;
      out[i] = (t < T && c < C) ? src[((long long)k * T + t) * Cp + c] : __float2bfloat16(0.f);
    }
  }
}
// g[c][t][k] += tmp[(t*CG + c)][k]  (c < C): fp32 weight gradient of the head GEMM back into the arena layout [C][T][K]
__global__ void fold_wgrad_unpack_kernel(const float* __restrict__ tmp, float* __restrict__ g, int C, int T, int CG, int K) {
  pdl_wait();
  pdl_launch_dependents();
  const long long total = (long long)C * T * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const long long ct = i / K;
    const int t = (int)(ct % T), c = (int)(ct / T);
    g[i] += tmp[((long long)t * CG + c) * K + k];
  }
}
// Row-window stem: weights [R][KH*KW][8] (the arena's direct pack, Cin padded to 8) -> [R][KH*KB][64] with the kernel
// row split into KB = ceil(KW / 8) blocks of 8 taps x 8 channels (zeros beyond KW); and the fp32 gradient back:
// g[r][kh*KW + kw][c] += tmp[r][kh*KB + kw/8][(kw%8)*8 + c]  (c < Cin), arena layout [R][T][Cin].
__global__ void rowwin_weight_pack_kernel(const bf16* __restrict__ src, bf16* __restrict__ out, int R, int KH, int KW, int KB) {
  pdl_wait();
  pdl_launch_dependents();
  const long long total = (long long)R * KH * KB * 64;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % 64);
    long long t = i / 64;
    const int kb = (int)(t % KB); t /= KB;
    const int kh = (int)(t % KH);
    const long long r = t / KH;
    const int kw = kb * 8 + e / 8, c = e % 8;
    out[i] = (kw < KW) ? src[((r * KH + kh) * KW + kw) * 8 + c] : __float2bfloat16(0.f);
  }
}
__global__ void rowwin_wgrad_unpack_kernel(const float* __restrict__ tmp, float* __restrict__ g, int R, int KH, int KW, int KB,
                                           int Cin) {
  pdl_wait();
  pdl_launch_dependents();
  const long long total = (long long)R * KH * KW * Cin;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cin);
    long long t = i / Cin;
    const int kw = (int)(t % KW); t /= KW;
    const int kh = (int)(t % KH);
    const long long r = t / KH;
    g[i] += tmp[((r * KH + kh) * KB + kw / 8) * 64 + (kw % 8) * 8 + c];
  }
}
// zero padding of an NHWC tensor (SRGAN's k9 p4 / VGG's k3 p1 convs feed the row-window stem, which wants a pre-padded
// image) and its backward (crop); `slack` extra zero rows at the end of every image keep the window reads in bounds.
__global__ void zero_pad_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int N, int H, int W, int G, int pad,
                                int slack, int backward) {
  pdl_wait();
  pdl_launch_dependents();
  const int PH = H + 2 * pad + slack, PW = W + 2 * pad;
  if (!backward) {
    const long long total = (long long)N * PH * PW * G;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
      const int g = (int)(i % G);
      long long t = i / G;
      const int pw = (int)(t % PW); t /= PW;
      const int ph = (int)(t % PH);
      const long long n = t / PH;
      const int h = ph - pad, w = pw - pad;
      y[i] = (h >= 0 && h < H && w >= 0 && w < W) ? x[((n * H + h) * W + w) * G + g] : make_uint4(0, 0, 0, 0);
    }
  } else {  // x = gradient on the padded grid, y = gradient on the original grid
    const long long total = (long long)N * H * W * G;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
      const int g = (int)(i % G);
      long long t = i / G;
      const int w = (int)(t % W); t /= W;
      const int h = (int)(t % H);
      const long long n = t / H;
      y[i] = x[((n * PH + h + pad) * PW + w + pad) * G + g];
    }
  }
}

}  // namespace gcc

using namespace gcc;

extern "C" int gcc_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int N, int C, long long HW, int Cp, int c_off,
                                         int zero_to, void* stream) {
  gcc_launch(nchw_to_nhwc_kernel, blocks_for((long long)N * HW), 256, 0, (cudaStream_t)stream, src, (bf16*)dst, N, C, HW, Cp,
                                                                                       c_off, zero_to);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_nhwc_bf16_to_nchw_f32(const void* src, float* dst, int N, int C, long long HW, int Cp, int c_off,
                                         int accumulate, void* stream) {
  gcc_launch(nhwc_to_nchw_kernel, blocks_for((long long)N * C * HW), 256, 0, (cudaStream_t)stream, (const bf16*)src, dst, N, C,
                                                                                           HW, Cp, c_off, accumulate);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_copy_channels_bf16(const void* src, int Cs, int s_off, void* dst, int Cd, int d_off, int C,
                                      long long npix, int accumulate, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!(Cs % 8) && !(s_off % 8) && !(Cd % 8) && !(d_off % 8) && !(C % 8)) {
    const int G = C / 8;
    gcc_launch(copy_channels_vec_kernel, blocks_for(npix * G), 256, 0, st, (const bf16*)src, Cs, s_off, (bf16*)dst, Cd, d_off,
                                                                   G, npix, accumulate);
  } else {
    gcc_launch(copy_channels_scalar_kernel, blocks_for(npix * C), 256, 0, st, (const bf16*)src, Cs, s_off, (bf16*)dst, Cd,
                                                                      d_off, C, npix, accumulate);
  }
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_cat_small_bf16(const void* a, const void* b, void* y, int ca, int cb, long long npix, void* stream) {
  if (ca < 0 || cb < 0 || ca + cb > 8) { gcc_set_error(__FILE__, __LINE__, "cat_small: ca + cb must be <= 8"); return GCC_ERR_ARG; }
  gcc_launch(cat_small_kernel, blocks_for(npix), 256, 0, (cudaStream_t)stream, (const uint4*)a, (const uint4*)b, (uint4*)y, ca, cb,
                                                                      npix);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_split_small_bf16(const void* dy, void* da, void* db, int ca, int cb, long long npix, void* stream) {
  if (ca < 0 || cb < 0 || ca + cb > 8) { gcc_set_error(__FILE__, __LINE__, "split_small: ca + cb must be <= 8"); return GCC_ERR_ARG; }
  gcc_launch(split_small_kernel, blocks_for(npix), 256, 0, (cudaStream_t)stream, (const uint4*)dy, (uint4*)da, (uint4*)db, ca, cb,
                                                                        npix);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_act_fwd_bf16(const void* x, void* y, long long n, int mode, float slope, void* stream) {
  if (n % 8) { gcc_set_error(__FILE__, __LINE__, "act: element count must be a multiple of 8"); return GCC_ERR_ARG; }
  gcc_launch(act_fwd_kernel, blocks_for(n / 8), 256, 0, (cudaStream_t)stream, (const bf16*)x, (bf16*)y, n / 8, mode, slope);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_act_bwd_bf16(const void* ref, const void* dy, void* dx, long long n, int mode, float slope,
                                void* stream) {
  if (n % 8) { gcc_set_error(__FILE__, __LINE__, "act: element count must be a multiple of 8"); return GCC_ERR_ARG; }
  gcc_launch(act_bwd_kernel, blocks_for(n / 8), 256, 0, (cudaStream_t)stream, (const bf16*)ref, (const bf16*)dy, (bf16*)dx,
                                                                      n / 8, mode, slope);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_dropout_bf16(const void* x, void* y, long long n, float p, const void* seed_dev, int salt,
                                void* stream) {
  gcc_launch(dropout_kernel, blocks_for(n), 256, 0, (cudaStream_t)stream, (const bf16*)x, (bf16*)y, n, p,
                                                                  (const unsigned long long*)seed_dev,
                                                                  (unsigned int)salt);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_add_bf16(const void* a, const void* b, void* y, long long n, void* stream) {
  if (n % 8) { gcc_set_error(__FILE__, __LINE__, "add: element count must be a multiple of 8"); return GCC_ERR_ARG; }
  gcc_launch(add_kernel, blocks_for(n / 8), 256, 0, (cudaStream_t)stream, (const bf16*)a, (const bf16*)b, (bf16*)y, n / 8);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_pack_weight_bf16(const float* src, void* direct, void* transposed, int D0, int T, int D1, int D1p,
                                    int D0p, void* stream) {
  gcc_launch(pack_weight_kernel, blocks_for((long long)D0 * T * D1), 256, 0, (cudaStream_t)stream, 
      src, (bf16*)direct, (bf16*)transposed, D0, T, D1, D1p, D0p);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_bias_grad_bf16(const void* dy, long long npix, int Cp, int c_off, int C, float* out, int accumulate,
                                  void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate && cudaMemsetAsync(out, 0, sizeof(float) * C, st) != cudaSuccess) return GCC_ERR_CUDA;
  if ((Cp % 8) || (c_off % 8)) {
    gcc_set_error(__FILE__, __LINE__, "bias_grad: channel window must be 8-aligned");
    return GCC_ERR_ARG;
  }
  const int G = (C + 7) / 8;
  int lanes = 256 / G;
  if (lanes < 1) lanes = 1;
  const int threads = (lanes * G + 31) / 32 * 32;
  long long bx = (npix + lanes * 16 - 1) / (lanes * 16);
  if (bx > 148 * 8) bx = 148 * 8;
  if (bx < 1) bx = 1;
  gcc_launch(colsum_vec_kernel, (unsigned)bx, threads, sizeof(float) * lanes * G * 8, st, (const bf16*)dy, npix, Cp, c_off, C, G,
                                                                                lanes, out);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_reflect_pad_bf16(const void* x, void* y, int N, int H, int W, int Cp, int pad, int backward,
                                    void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (Cp % 8 || pad >= H || pad >= W) { gcc_set_error(__FILE__, __LINE__, "reflect_pad: bad arguments"); return GCC_ERR_ARG; }
  const int G = Cp / 8;
  if (!backward)
    gcc_launch(reflect_pad_kernel, blocks_for((long long)N * (H + 2 * pad) * (W + 2 * pad) * G), 256, 0, st, 
        (const bf16*)x, (bf16*)y, N, H, W, G, pad);
  else  // x = dy on the padded grid [N, H+2p, W+2p, Cp], y = dx [N, H, W, Cp]
    gcc_launch(reflect_pad_bwd_kernel, blocks_for((long long)N * H * W * G), 256, 0, st, (const bf16*)x, (bf16*)y, N, H, W, G,
                                                                                 pad);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
// launch geometry of dw3x3_rows_kernel: lanes pixel lanes per CTA, row segments short enough to fill the machine
static void dw_rows_geometry(int N, int H, int W, int G, int* lanes, int* threads, int* seg, long long* items, int* blocks) {
  int l = 256 / G;
  if (l < 1) l = 1;
  int sg = 64;
  while (sg > 8 && (long long)N * H * ((W + sg - 1) / sg) < 2LL * 148 * l) sg >>= 1;
  if (sg > W) sg = W;
  const long long it = (long long)N * H * ((W + sg - 1) / sg);
  long long b = (it + l - 1) / l;
  if (b > 148 * 8) b = 148 * 8;
  *lanes = l;
  *threads = (l * G + 31) / 32 * 32;
  *seg = sg;
  *items = it;
  *blocks = (int)(b < 1 ? 1 : b);
}
extern "C" int gcc_dw3x3_fwd_bf16(const void* x, const float* w, const float* bias, void* y, int N, int H, int W,
                                  int Cp, int C, void* stream) {
  if (Cp % 8 || Cp / 8 > 256 || H < 2 || W < 2) {
    gcc_set_error(__FILE__, __LINE__, "dw3x3: bad channel count or extent");
    return GCC_ERR_ARG;
  }
  int lanes, threads, seg, blocks;
  long long items;
  dw_rows_geometry(N, H, W, Cp / 8, &lanes, &threads, &seg, &items, &blocks);
  gcc_launch(dw3x3_rows_kernel<0>, blocks, threads, 0, (cudaStream_t)stream, (const bf16*)x, w, bias, (bf16*)y, H, W,
             Cp / 8, C, lanes, seg, items);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_dw3x3_bwd_bf16(const void* x, const void* dy, const float* w, void* dxp, void* dx, float* dw,
                                  float* dbias, int N, int H, int W, int Cp, int C, int accumulate, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (Cp % 8 || Cp / 8 > 256) { gcc_set_error(__FILE__, __LINE__, "dw3x3: bad channel count"); return GCC_ERR_ARG; }
  const int G = Cp / 8;
  if (dx != nullptr) {
    if (H < 2 || W < 2) { gcc_set_error(__FILE__, __LINE__, "dw3x3: extent"); return GCC_ERR_ARG; }
    int lanes, threads, seg, blocks;
    long long items;
    dw_rows_geometry(N, H, W, G, &lanes, &threads, &seg, &items, &blocks);
    gcc_launch(dw3x3_rows_kernel<1>, blocks, threads, 0, st, (const bf16*)dy, w, (const float*)nullptr, (bf16*)dx, H, W, G,
               C, lanes, seg, items);
    GCC_CHECK_LAUNCH();
  }
  if (dw != nullptr) {
    if (!accumulate) {
      if (cudaMemsetAsync(dw, 0, sizeof(float) * C * 9, st) != cudaSuccess) return GCC_ERR_CUDA;
      if (dbias && cudaMemsetAsync(dbias, 0, sizeof(float) * C, st) != cudaSuccess) return GCC_ERR_CUDA;
    }
    int lanes, threads, seg, blocks;
    long long items;
    dw_rows_geometry(N, H, W, G, &lanes, &threads, &seg, &items, &blocks);
    if (blocks > 148 * 2) blocks = 148 * 2;      // every CTA ends with G * 80 atomics
    const size_t smem = sizeof(float) * lanes * G * 80;
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    dev = (dev >= 0 && dev < 64) ? dev : 0;
    if (!configured[dev]) {
      cudaFuncSetAttribute(dw3x3_wgrad_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      configured[dev] = true;
    }
    gcc_launch(dw3x3_wgrad_rows_kernel, blocks, threads, smem, st, (const bf16*)x, (const bf16*)dy, dw, dbias, H, W, G, C, lanes,
               seg, items);
    GCC_CHECK_LAUNCH();
  }
  return GCC_OK;
}

extern "C" int gcc_im2col_k4s2_c8(const void* img, void* col, int N, int H, int W, void* stream) {
  if ((H % 2) || (W % 2)) { gcc_set_error(__FILE__, __LINE__, "im2col: H, W must be even"); return GCC_ERR_ARG; }
  gcc_launch(im2col_k4s2_c8_kernel, blocks_for((long long)N * (H / 2) * (W / 2) * 16), 256, 0, (cudaStream_t)stream, 
      (const uint4*)img, (uint4*)col, N, H, W);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_col2im_k4s2_c8(const void* col, int Ccol, int order, int C, const float* bias, int act, void* img,
                                  int N, int H, int W, void* stream) {
  if ((H % 2) || (W % 2) || C > 8 || (Ccol % 8) || order < 0 || order > 2 || (order == 2 && C > 4)) {
    gcc_set_error(__FILE__, __LINE__, "col2im: bad arguments");
    return GCC_ERR_ARG;
  }
  if ((long long)N * H * W < (1LL << 31) - (1LL << 24))   // (headroom for the grid-stride increment)
    gcc_launch(col2im_k4s2_c8_kernel<unsigned>, blocks_for((long long)N * H * W), 256, 0, (cudaStream_t)stream,
               (const bf16*)col, Ccol, order, C, bias, act, (uint4*)img, N, H, W);
  else
    gcc_launch(col2im_k4s2_c8_kernel<long long>, blocks_for((long long)N * H * W), 256, 0, (cudaStream_t)stream,
               (const bf16*)col, Ccol, order, C, bias, act, (uint4*)img, N, H, W);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_unpad_wgrad_c8(const float* tmp, float* g, int R, int C, void* stream) {
  if (C > 8) { gcc_set_error(__FILE__, __LINE__, "unpad_wgrad: C must be <= 8"); return GCC_ERR_ARG; }
  gcc_launch(unpad_wgrad_c8_kernel, (R * 16 * C + 255) / 256, 256, 0, (cudaStream_t)stream, tmp, g, R, C);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}

extern "C" int gcc_fold_k4s1_c8(const void* ycol, int Ccol, int C, const float* bias, void* y, int N, int H, int W,
                                void* stream) {
  if (C > 8 || H < 2 || W < 2) { gcc_set_error(__FILE__, __LINE__, "fold_k4s1: bad arguments"); return GCC_ERR_ARG; }
  gcc_launch(fold_k4s1_kernel, blocks_for((long long)N * (H - 1) * (W - 1)), 256, 0, (cudaStream_t)stream, 
      (const bf16*)ycol, Ccol, C, bias, (uint4*)y, N, H, W);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_unfold_k4s1_c8(const void* dy, void* dcol, int N, int H, int W, void* stream) {
  gcc_launch(unfold_k4s1_kernel, blocks_for((long long)N * H * W * 16), 256, 0, (cudaStream_t)stream, (const uint4*)dy,
                                                                                            (uint4*)dcol, N, H, W);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_unpad_wgrad_rows(const float* tmp, float* g, int C, int K, void* stream) {
  if (C > 8) { gcc_set_error(__FILE__, __LINE__, "unpad_wgrad_rows: C must be <= 8"); return GCC_ERR_ARG; }
  gcc_launch(unpad_wgrad_rows_kernel, blocks_for((long long)C * 16 * K), 256, 0, (cudaStream_t)stream, tmp, g, C, K);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}

extern "C" int gcc_image_pool_query_bf16(const void* images, void* pool, long long* state_dev, int* dec_ws, void* out,
                                         int b, long long elems_per_image, int pool_size, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (elems_per_image % 8 || b <= 0 || pool_size <= 0) {
    gcc_set_error(__FILE__, __LINE__, "image_pool_query: bad arguments");
    return GCC_ERR_ARG;
  }
  const long long vec = elems_per_image / 8;
  gcc_launch(image_pool_decide_kernel, 1, 32, 0, st, state_dev, pool_size, b, dec_ws);
  GCC_CHECK_LAUNCH();
  long long bx = (vec + 255) / 256;
  if (bx > 148 * 4) bx = 148 * 4;
  gcc_launch(image_pool_gather_kernel, dim3((unsigned)bx, b), 256, 0, st, (const uint4*)images, (const uint4*)pool, dec_ws,
                                                                  (uint4*)out, vec);
  GCC_CHECK_LAUNCH();
  gcc_launch(image_pool_scatter_kernel, dim3((unsigned)bx, b), 256, 0, st, (const uint4*)images, (uint4*)pool, dec_ws, vec);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}

extern "C" int gcc_fold_taps_bf16(const void* ycol, int Ccol, int CG, int KH, int KW, int C, const float* bias, int act,
                                  void* y, int N, int GH, int GW, int OH, int OW, int dir, int off, void* stream) {
  if ((CG != 4 && CG != 8) || C > CG || (Ccol % 8) || Ccol < KH * KW * CG || (dir != 1 && dir != -1)) {
    gcc_set_error(__FILE__, __LINE__, "fold_taps: bad arguments");
    return GCC_ERR_ARG;
  }
  const long long total = (long long)N * OH * OW;
  if (CG == 4)
    gcc_launch(fold_taps_kernel<4>, blocks_for(total), 256, 0, (cudaStream_t)stream, (const bf16*)ycol, Ccol, KH, KW, C, bias,
               act, (uint4*)y, N, GH, GW, OH, OW, dir, off);
  else
    gcc_launch(fold_taps_kernel<8>, blocks_for(total), 256, 0, (cudaStream_t)stream, (const bf16*)ycol, Ccol, KH, KW, C, bias,
               act, (uint4*)y, N, GH, GW, OH, OW, dir, off);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_unfold_taps_bf16(const void* dy, void* dcol, int Ccol, int CG, int KH, int KW, int N, int GH, int GW,
                                    int OH, int OW, int off, void* stream) {
  if ((CG != 4 && CG != 8) || (Ccol % 8) || Ccol < KH * KW * CG) {
    gcc_set_error(__FILE__, __LINE__, "unfold_taps: bad arguments");
    return GCC_ERR_ARG;
  }
  const long long total = (long long)N * GH * GW * (Ccol / 8);
  if (CG == 4)
    gcc_launch(unfold_taps_kernel<4>, blocks_for(total), 256, 0, (cudaStream_t)stream, (const bf16*)dy, (bf16*)dcol, Ccol, KH,
               KW, N, GH, GW, OH, OW, off);
  else
    gcc_launch(unfold_taps_kernel<8>, blocks_for(total), 256, 0, (cudaStream_t)stream, (const bf16*)dy, (bf16*)dcol, Ccol, KH,
               KW, N, GH, GW, OH, OW, off);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_fold_weight_pack_bf16(const void* src, void* out, int mode, int C, int T, int CG, int K, int Kp, int Cp,
                                         int rows_p, void* stream) {
  if (rows_p < T * CG || (rows_p % 8)) { gcc_set_error(__FILE__, __LINE__, "fold_weight_pack: bad rows_p"); return GCC_ERR_ARG; }
  const long long total = mode == 0 ? (long long)rows_p * Kp : (long long)K * rows_p;
  gcc_launch(fold_weight_pack_kernel, blocks_for(total), 256, 0, (cudaStream_t)stream, (const bf16*)src, (bf16*)out, mode, C, T,
             CG, K, Kp, Cp, rows_p);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_fold_wgrad_unpack_f32(const float* tmp, float* g, int C, int T, int CG, int K, void* stream) {
  gcc_launch(fold_wgrad_unpack_kernel, blocks_for((long long)C * T * K), 256, 0, (cudaStream_t)stream, tmp, g, C, T, CG, K);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_rowwin_weight_pack_bf16(const void* src, void* out, int R, int KH, int KW, void* stream) {
  const int KB = (KW + 7) / 8;
  gcc_launch(rowwin_weight_pack_kernel, blocks_for((long long)R * KH * KB * 64), 256, 0, (cudaStream_t)stream, (const bf16*)src,
             (bf16*)out, R, KH, KW, KB);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_rowwin_wgrad_unpack_f32(const float* tmp, float* g, int R, int KH, int KW, int Cin, void* stream) {
  const int KB = (KW + 7) / 8;
  gcc_launch(rowwin_wgrad_unpack_kernel, blocks_for((long long)R * KH * KW * Cin), 256, 0, (cudaStream_t)stream, tmp, g, R, KH,
             KW, KB, Cin);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_zero_pad_bf16(const void* x, void* y, int N, int H, int W, int Cp, int pad, int slack, int backward,
                                 void* stream) {
  if (Cp % 8) { gcc_set_error(__FILE__, __LINE__, "zero_pad: bad channel count"); return GCC_ERR_ARG; }
  const int G = Cp / 8;
  const long long total = backward ? (long long)N * H * W * G : (long long)N * (H + 2 * pad + slack) * (W + 2 * pad) * G;
  gcc_launch(zero_pad_kernel, blocks_for(total), 256, 0, (cudaStream_t)stream, (const uint4*)x, (uint4*)y, N, H, W, G, pad, slack,
             backward);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
