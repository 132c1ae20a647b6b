// Kernels the SAGAN step needs on top of the pix2pix set (reference: models/SAGAN.py): SpectralNorm's power
// iteration / weight scaling and its backward (:14-71) and the Self_Attn block (:73-107).  Small, launch-bound
// layers (the whole model runs at 64 x 64): straightforward CUDA-core kernels, fp32 math, bf16 activations.
#include "common.cuh"

namespace gcc {

__device__ __forceinline__ float block_sum(float v, float* part) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) part[w] = v;
  __syncthreads();
  float r = 0.f;
  if (w == 0) {
    r = lane < (blockDim.x >> 5) ? part[lane] : 0.f;
    r = warp_sum(r);
    if (lane == 0) part[0] = r;
  }
  __syncthreads();
  r = part[0];
  __syncthreads();
  return r;
}

// ---------------------------------------------------------------------------------------- spectral norm
// W: fp32 [height][width] row-major (the arena's channels-last conv weight; height = weight.shape[0]).
// v_raw[j] = sum_r W[r][j] u[r]      (written into v), scratch[0] += |v_raw|^2
__global__ void sn_power_v_kernel(const float* __restrict__ W, const float* __restrict__ u, float* __restrict__ v,
                                  int height, int width, float* __restrict__ scratch) {
  __shared__ float part[32];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f;
  if (j < width)
    for (int r = 0; r < height; ++r) acc += W[(long long)r * width + j] * u[r];
  if (j < width) v[j] = acc;
  const float s = block_sum(j < width ? acc * acc : 0.f, part);
  if (threadIdx.x == 0) atomicAdd(scratch, s);
}
// t[r] = sum_j W[r][j] v_raw[j] / (|v_raw| + eps), scratch[1] += t[r]^2
__global__ void sn_power_u_kernel(const float* __restrict__ W, const float* __restrict__ v, float* __restrict__ t,
                                  int height, int width, float* __restrict__ scratch) {
  __shared__ float part[32];
  const int r = blockIdx.x;
  const float inv = 1.f / (sqrtf(scratch[0]) + 1e-12f);
  float acc = 0.f;
  for (int j = threadIdx.x; j < width; j += blockDim.x) acc += W[(long long)r * width + j] * v[j];
  const float s = block_sum(acc, part) * inv;
  if (threadIdx.x == 0) {
    t[r] = s;
    atomicAdd(scratch + 1, s * s);
  }
}
// v = v_raw / (|v_raw| + eps);  u = t / (|t| + eps);  sigma = u . t
__global__ void sn_finish_kernel(float* __restrict__ u, float* __restrict__ v, const float* __restrict__ t, int height,
                                 int width, const float* __restrict__ scratch, float* __restrict__ sigma) {
  const float invv = 1.f / (sqrtf(scratch[0]) + 1e-12f);
  const float nu = sqrtf(scratch[1]);
  const float invu = 1.f / (nu + 1e-12f);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < width; i += gridDim.x * blockDim.x) v[i] *= invv;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < height; i += gridDim.x * blockDim.x) u[i] = t[i] * invu;
  if (blockIdx.x == 0 && threadIdx.x == 0) *sigma = scratch[1] * invu;
}
// bf16 operand packs of W / sigma
__global__ void pack_weight_scaled_kernel(const float* __restrict__ src, const float* __restrict__ sigma,
                                          bf16* __restrict__ direct, bf16* __restrict__ transposed, int D0, int T, int D1,
                                          int D1p, int D0p) {
  const float inv = 1.f / *sigma;
  const long long total = (long long)D0 * T * D1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int d1 = (int)(i % D1);
    const long long r = i / D1;
    const int t = (int)(r % T);
    const int d0 = (int)(r / T);
    const bf16 v = __float2bfloat16(src[i] * inv);
    if (direct) direct[((long long)d0 * T + t) * D1p + d1] = v;
    if (transposed) transposed[((long long)d1 * T + t) * D0p + d0] = v;
  }
}
// scratch[0] += sum dWeff * Wbar
__global__ void sn_dot_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                              float* __restrict__ scratch) {
  __shared__ float part[32];
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc += a[i] * b[i];
  const float s = block_sum(acc, part);
  if (threadIdx.x == 0) atomicAdd(scratch, s);
}
// dWbar[r][j] += (dWeff[r][j] - (s_raw / sigma) u[r] v[j]) / sigma
__global__ void sn_bwd_w_kernel(const float* __restrict__ dweff, const float* __restrict__ u, const float* __restrict__ v,
                                const float* __restrict__ sigma, const float* __restrict__ scratch, int height, int width,
                                float* __restrict__ dwbar) {
  const float inv = 1.f / *sigma;
  const float s = scratch[0] * inv;
  const long long n = (long long)height * width;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / width), j = (int)(i % width);
    dwbar[i] += (dweff[i] - s * u[r] * v[j]) * inv;
  }
}
// du[r] += dsigma * t[r];  dv[j] += dsigma * sum_r Wbar[r][j] u[r];   dsigma = -s_raw / sigma^2
__global__ void sn_bwd_uv_kernel(const float* __restrict__ W, const float* __restrict__ u, const float* __restrict__ t,
                                 const float* __restrict__ sigma, const float* __restrict__ scratch, int height, int width,
                                 float* __restrict__ du, float* __restrict__ dv) {
  const float inv = 1.f / *sigma;
  const float ds = -scratch[0] * inv * inv;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < width) {
    float acc = 0.f;
    for (int r = 0; r < height; ++r) acc += W[(long long)r * width + j] * u[r];
    dv[j] += ds * acc;
  }
  if (j < height) du[j] += ds * t[j];
}

// ------------------------------------------------------------------------------------------- attention
// q, k: bf16 [N][L][dp] (d logical), v: bf16 [N][L][Cp].  One CTA per (n, query i):
//   e_j = q_i . k_j ; p = softmax_j(e) ; out[i][c] = sum_j p_j v[j][c]          (Self_Attn.forward, SAGAN.py:96-104)
__global__ void attn_fwd_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                                int L, int d, int dp, int C, int Cp, bf16* __restrict__ probs, bf16* __restrict__ out) {
  extern __shared__ float sm[];  // [L] energies/probs, [d] query
  float* e = sm;
  float* qs = sm + L;
  __shared__ float part[32];
  const int n = blockIdx.y, i = blockIdx.x;
  const bf16* qi = q + ((long long)n * L + i) * dp;
  for (int x = threadIdx.x; x < d; x += blockDim.x) qs[x] = __bfloat162float(qi[x]);
  __syncthreads();
  float mx = -3.4e38f;
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    const bf16* kj = k + ((long long)n * L + j) * dp;
    float acc = 0.f;
    for (int x = 0; x < d; ++x) acc += qs[x] * __bfloat162float(kj[x]);
    e[j] = acc;
    mx = fmaxf(mx, acc);
  }
  // block max
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = part[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) mx = fmaxf(mx, part[w]);
  __syncthreads();
  float sum = 0.f;
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    const float p = __expf(e[j] - mx);
    e[j] = p;
    sum += p;
  }
  sum = block_sum(sum, part);
  const float inv = 1.f / sum;
  bf16* prow = probs + ((long long)n * L + i) * L;
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    const float p = e[j] * inv;
    e[j] = p;
    prow[j] = __float2bfloat16(p);
  }
  __syncthreads();
  bf16* orow = out + ((long long)n * L + i) * Cp;
  for (int c = threadIdx.x; c < Cp; c += blockDim.x) {
    float acc = 0.f;
    if (c < C)
      for (int j = 0; j < L; ++j) acc += e[j] * __bfloat162float(v[((long long)n * L + j) * Cp + c]);
    orow[c] = __float2bfloat16(acc);
  }
}
// One CTA per (n, query i): dp_j = do_i . v_j ; de_j = p_j (dp_j - sum_j' p_j' dp_j') ; dq_i = sum_j de_j k_j.
// de is written (bf16) over a scratch [N][L][L] for the key/value pass.
__global__ void attn_bwd_q_kernel(const bf16* __restrict__ k, const bf16* __restrict__ v, const bf16* __restrict__ probs,
                                  const bf16* __restrict__ dout, int L, int d, int dp, int C, int Cp,
                                  bf16* __restrict__ de, bf16* __restrict__ dq) {
  extern __shared__ float sm[];  // [L] de, [C] do
  float* es = sm;
  float* ds = sm + L;
  __shared__ float part[32];
  const int n = blockIdx.y, i = blockIdx.x;
  const bf16* dor = dout + ((long long)n * L + i) * Cp;
  for (int c = threadIdx.x; c < C; c += blockDim.x) ds[c] = __bfloat162float(dor[c]);
  __syncthreads();
  const bf16* prow = probs + ((long long)n * L + i) * L;
  float dot = 0.f;
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    const bf16* vj = v + ((long long)n * L + j) * Cp;
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc += ds[c] * __bfloat162float(vj[c]);
    es[j] = acc;
    dot += acc * __bfloat162float(prow[j]);
  }
  dot = block_sum(dot, part);
  bf16* derow = de + ((long long)n * L + i) * L;
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    const float g = __bfloat162float(prow[j]) * (es[j] - dot);
    es[j] = g;
    derow[j] = __float2bfloat16(g);
  }
  __syncthreads();
  bf16* dqr = dq + ((long long)n * L + i) * dp;
  for (int x = threadIdx.x; x < dp; x += blockDim.x) {
    float acc = 0.f;
    if (x < d)
      for (int j = 0; j < L; ++j) acc += es[j] * __bfloat162float(k[((long long)n * L + j) * dp + x]);
    dqr[x] = __float2bfloat16(acc);
  }
}
// One CTA per (n, key j): dv_j = sum_i p_ij do_i ; dk_j = sum_i de_ij q_i
__global__ void attn_bwd_kv_kernel(const bf16* __restrict__ q, const bf16* __restrict__ probs, const bf16* __restrict__ de,
                                   const bf16* __restrict__ dout, int L, int d, int dp, int C, int Cp,
                                   bf16* __restrict__ dk, bf16* __restrict__ dv) {
  extern __shared__ float sm[];  // [L] p_.j, [L] de_.j
  float* ps = sm;
  float* es = sm + L;
  const int n = blockIdx.y, j = blockIdx.x;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    ps[i] = __bfloat162float(probs[((long long)n * L + i) * L + j]);
    es[i] = __bfloat162float(de[((long long)n * L + i) * L + j]);
  }
  __syncthreads();
  bf16* dvr = dv + ((long long)n * L + j) * Cp;
  for (int c = threadIdx.x; c < Cp; c += blockDim.x) {
    float acc = 0.f;
    if (c < C)
      for (int i = 0; i < L; ++i) acc += ps[i] * __bfloat162float(dout[((long long)n * L + i) * Cp + c]);
    dvr[c] = __float2bfloat16(acc);
  }
  bf16* dkr = dk + ((long long)n * L + j) * dp;
  for (int x = threadIdx.x; x < dp; x += blockDim.x) {
    float acc = 0.f;
    if (x < d)
      for (int i = 0; i < L; ++i) acc += es[i] * __bfloat162float(q[((long long)n * L + i) * dp + x]);
    dkr[x] = __float2bfloat16(acc);
  }
}
// y = gamma * a + x   (gamma: learnable device scalar)
__global__ void scale_add_kernel(const bf16* __restrict__ a, const bf16* __restrict__ x, const float* __restrict__ gamma,
                                 bf16* __restrict__ y, long long n) {
  const float g = *gamma;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16(g * __bfloat162float(a[i]) + __bfloat162float(x[i]));
}
// da = gamma * dy ; dgamma += sum dy * a
__global__ void scale_add_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ a,
                                     const float* __restrict__ gamma, bf16* __restrict__ da, float* __restrict__ dgamma,
                                     long long n) {
  __shared__ float part[32];
  const float g = *gamma;
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = __bfloat162float(dy[i]);
    acc += d * __bfloat162float(a[i]);
    da[i] = __float2bfloat16(g * d);
  }
  if (dgamma != nullptr) {
    const float s = block_sum(acc, part);
    if (threadIdx.x == 0) atomicAdd(dgamma, s);
  }
}

static inline int sa_blocks(long long n) {
  long long b = (n + 255) / 256;
  if (b < 1) b = 1;
  if (b > 148 * 8) b = 148 * 8;
  return (int)b;
}

}  // namespace gcc

using namespace gcc;

extern "C" int gcc_spectral_norm_fwd(const float* w_bar, float* u, float* v, int height, int width, float* t_out,
                                     float* sigma_out, float* scratch2, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(scratch2, 0, 2 * sizeof(float), st) != cudaSuccess) return GCC_ERR_CUDA;
  sn_power_v_kernel<<<(width + 255) / 256, 256, 0, st>>>(w_bar, u, v, height, width, scratch2);
  GCC_CHECK_LAUNCH();
  sn_power_u_kernel<<<height, 256, 0, st>>>(w_bar, v, t_out, height, width, scratch2);
  GCC_CHECK_LAUNCH();
  const int m = height > width ? height : width;
  sn_finish_kernel<<<(m + 255) / 256, 256, 0, st>>>(u, v, t_out, height, width, scratch2, sigma_out);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_pack_weight_scaled_bf16(const float* src, const float* sigma_dev, void* direct, void* transposed,
                                           int D0, int T, int D1, int D1p, int D0p, void* stream) {
  pack_weight_scaled_kernel<<<sa_blocks((long long)D0 * T * D1), 256, 0, (cudaStream_t)stream>>>(
      src, sigma_dev, (bf16*)direct, (bf16*)transposed, D0, T, D1, D1p, D0p);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_spectral_norm_bwd(const float* dw_eff, const float* w_bar, const float* u, const float* v,
                                     const float* sigma, const float* t_saved, int height, int width, float* dw_bar,
                                     float* du, float* dv, float* scratch1, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)height * width;
  if (cudaMemsetAsync(scratch1, 0, sizeof(float), st) != cudaSuccess) return GCC_ERR_CUDA;
  int b = sa_blocks(n);
  if (b > 148 * 2) b = 148 * 2;
  sn_dot_kernel<<<b, 256, 0, st>>>(dw_eff, w_bar, n, scratch1);
  GCC_CHECK_LAUNCH();
  sn_bwd_w_kernel<<<sa_blocks(n), 256, 0, st>>>(dw_eff, u, v, sigma, scratch1, height, width, dw_bar);
  GCC_CHECK_LAUNCH();
  if (du != nullptr && dv != nullptr) {
    const int m = height > width ? height : width;
    sn_bwd_uv_kernel<<<(m + 255) / 256, 256, 0, st>>>(w_bar, u, t_saved, sigma, scratch1, height, width, du, dv);
    GCC_CHECK_LAUNCH();
  }
  return GCC_OK;
}
extern "C" int gcc_attn_fwd_bf16(const void* q, const void* k, const void* v, int N, int L, int d, int dp, int C, int Cp,
                                 void* probs, void* out, void* stream) {
  if (L > 4096 || d > 512) { gcc_set_error(__FILE__, __LINE__, "attention: L <= 4096 and d <= 512"); return GCC_ERR_ARG; }
  attn_fwd_kernel<<<dim3(L, N), 128, sizeof(float) * (L + d), (cudaStream_t)stream>>>(
      (const bf16*)q, (const bf16*)k, (const bf16*)v, L, d, dp, C, Cp, (bf16*)probs, (bf16*)out);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_attn_bwd_bf16(const void* q, const void* k, const void* v, const void* probs, const void* dout, int N,
                                 int L, int d, int dp, int C, int Cp, void* de_scratch, void* dq, void* dk, void* dv,
                                 void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (L > 4096 || C > 4096) { gcc_set_error(__FILE__, __LINE__, "attention: L, C <= 4096"); return GCC_ERR_ARG; }
  attn_bwd_q_kernel<<<dim3(L, N), 128, sizeof(float) * (L + C), st>>>((const bf16*)k, (const bf16*)v, (const bf16*)probs,
                                                                     (const bf16*)dout, L, d, dp, C, Cp, (bf16*)de_scratch,
                                                                     (bf16*)dq);
  GCC_CHECK_LAUNCH();
  attn_bwd_kv_kernel<<<dim3(L, N), 128, sizeof(float) * 2 * L, st>>>((const bf16*)q, (const bf16*)probs,
                                                                    (const bf16*)de_scratch, (const bf16*)dout, L, d, dp, C,
                                                                    Cp, (bf16*)dk, (bf16*)dv);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_scale_add_bf16(const void* a, const void* x, const float* gamma_dev, void* y, long long n,
                                  void* stream) {
  scale_add_kernel<<<sa_blocks(n), 256, 0, (cudaStream_t)stream>>>((const bf16*)a, (const bf16*)x, gamma_dev, (bf16*)y, n);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_scale_add_bwd_bf16(const void* dy, const void* a, const float* gamma_dev, void* da, float* dgamma,
                                      long long n, void* stream) {
  int b = sa_blocks(n);
  if (b > 148 * 2) b = 148 * 2;
  scale_add_bwd_kernel<<<b, 256, 0, (cudaStream_t)stream>>>((const bf16*)dy, (const bf16*)a, gamma_dev, (bf16*)da, dgamma, n);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
