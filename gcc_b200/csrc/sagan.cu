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
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
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
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
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
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
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
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
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
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
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
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
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
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
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
// Self_Attn.forward (SAGAN.py:96-104) as batched tensor-core GEMMs (conv_gemm.cu, one weight matrix per image) plus
// three row-wise kernels.  q, k: bf16 [N][L][dp] (d logical), v: bf16 [N][L][Cp]:
//   energy = q k^T  (tcgen05, fp32 out)  ->  probs = softmax_j(energy) (bf16)  ->  out = probs v  (tcgen05)
// and in backward  dprobs = dout v^T (tcgen05, fp32),  de = probs * (dprobs - sum_j probs dprobs)  (bf16),
//   dq = de k (tcgen05),  dk = de^T q,  dv = probs^T dout  (batched weight-gradient GEMMs over the query positions).

// probs[n][i][:] = softmax(energy[n][i][:]); one warp per row, fp32 in, bf16 out
__global__ void attn_softmax_kernel(const float* __restrict__ energy, bf16* __restrict__ probs, long long rows, int L) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* e = energy + row * L;
  float mx = -3.4e38f;
  for (int j = lane; j < L; j += 32) mx = fmaxf(mx, e[j]);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int j = lane; j < L; j += 32) sum += __expf(e[j] - mx);
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  bf16* p = probs + row * L;
  for (int j = lane; j < L; j += 32) p[j] = __float2bfloat16(__expf(e[j] - mx) * inv);
}
// de[n][i][j] = p_ij * (dp_ij - sum_j' p_ij' dp_ij')   (softmax backward), one warp per row
__global__ void attn_softmax_bwd_kernel(const bf16* __restrict__ probs, const float* __restrict__ dprobs,
                                        bf16* __restrict__ de, long long rows, int L) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const bf16* p = probs + row * L;
  const float* dp = dprobs + row * L;
  float dot = 0.f;
  for (int j = lane; j < L; j += 32) dot += __bfloat162float(p[j]) * dp[j];
  dot = warp_sum(dot);
  bf16* o = de + row * L;
  for (int j = lane; j < L; j += 32) o[j] = __float2bfloat16(__bfloat162float(p[j]) * (dp[j] - dot));
}
// dst[n][c][l] = src[n][l][c] for c < C (rows C..Cd-1 of dst are not touched), 32 x 32 tiles through shared memory
__global__ void attn_transpose_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, int L, int Cs, int C) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ bf16 tile[32][33];
  const int n = blockIdx.z, l0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int l = l0 + r, c = c0 + tx;
    tile[r][tx] = (l < L && c < C) ? src[((long long)n * L + l) * Cs + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, l = l0 + tx;
    if (c < C && l < L) dst[((long long)n * C + c) * L + l] = tile[tx][r];
  }
}
// dst (bf16 [rows][Cd]) = src (fp32 [rows][Cs]) for the first C columns, zero in the padding
__global__ void attn_f32_to_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long rows, int Cs, int Cd,
                                        int C) {
  pdl_wait();
  pdl_launch_dependents();
  const long long total = rows * Cd;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cd);
    const long long r = i / Cd;
    dst[i] = __float2bfloat16(c < C ? src[r * Cs + c] : 0.f);
  }
}
// y = gamma * a + x   (gamma: learnable device scalar)
__global__ void scale_add_kernel(const bf16* __restrict__ a, const bf16* __restrict__ x, const float* __restrict__ gamma,
                                 bf16* __restrict__ y, long long n) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const float g = *gamma;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16(g * __bfloat162float(a[i]) + __bfloat162float(x[i]));
}
// da = gamma * dy ; dgamma += sum dy * a
__global__ void scale_add_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ a,
                                     const float* __restrict__ gamma, bf16* __restrict__ da, float* __restrict__ dgamma,
                                     long long n) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
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
  gcc_launch(sn_power_v_kernel, (width + 255) / 256, 256, 0, st, w_bar, u, v, height, width, scratch2);
  GCC_CHECK_LAUNCH();
  gcc_launch(sn_power_u_kernel, height, 256, 0, st, w_bar, v, t_out, height, width, scratch2);
  GCC_CHECK_LAUNCH();
  const int m = height > width ? height : width;
  gcc_launch(sn_finish_kernel, (m + 255) / 256, 256, 0, st, u, v, t_out, height, width, scratch2, sigma_out);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_pack_weight_scaled_bf16(const float* src, const float* sigma_dev, void* direct, void* transposed,
                                           int D0, int T, int D1, int D1p, int D0p, void* stream) {
  gcc_launch(pack_weight_scaled_kernel, sa_blocks((long long)D0 * T * D1), 256, 0, (cudaStream_t)stream, 
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
  gcc_launch(sn_dot_kernel, b, 256, 0, st, dw_eff, w_bar, n, scratch1);
  GCC_CHECK_LAUNCH();
  gcc_launch(sn_bwd_w_kernel, sa_blocks(n), 256, 0, st, dw_eff, u, v, sigma, scratch1, height, width, dw_bar);
  GCC_CHECK_LAUNCH();
  if (du != nullptr && dv != nullptr) {
    const int m = height > width ? height : width;
    gcc_launch(sn_bwd_uv_kernel, (m + 255) / 256, 256, 0, st, w_bar, u, t_saved, sigma, scratch1, height, width, du, dv);
    GCC_CHECK_LAUNCH();
  }
  return GCC_OK;
}
// workspace layout helpers (bytes, 256-byte aligned pieces)
static inline long long al256(long long b) { return (b + 255) / 256 * 256; }
extern "C" long long gcc_attn_workspace_bytes(int N, int L, int dp, int Cp, int backward) {
  const long long nll = (long long)N * L * L;
  if (!backward) return al256(nll * 4) + al256((long long)N * Cp * L * 2);
  return al256(nll * 4) + al256((long long)N * dp * L * 2) + al256((long long)N * L * Cp * 4) + al256((long long)N * L * dp * 4);
}

extern "C" int gcc_attn_fwd_bf16(const void* q, const void* k, const void* v, int N, int L, int d, int dp, int C, int Cp,
                                 void* probs, void* out, void* ws, long long ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if ((L % 8) || (dp % 8) || (Cp % 8) || d > dp || C > Cp || ws == nullptr ||
      ws_bytes < gcc_attn_workspace_bytes(N, L, dp, Cp, 0)) {
    gcc_set_error(__FILE__, __LINE__, "attention: L, dp, Cp must be multiples of 8 and the workspace large enough");
    return GCC_ERR_ARG;
  }
  const long long nll = (long long)N * L * L;
  float* energy = (float*)ws;
  bf16* vT = (bf16*)((char*)ws + al256(nll * 4));
  // energy[n][i][j] = sum_c q[n][i][c] k[n][j][c]: pixel-major GEMM over the L positions, weights = this image's keys
  int rc = gcc_conv_gemm_launch(q, N, 1, L, dp, k, L, 1, dp, nullptr, energy, 1, L, L, 0, 0, 1, 1, 1, 0, 0, 0.f, 1, energy,
                                nll, nullptr, 0, 1, 0, stream);
  if (rc) return rc;
  const long long rows = (long long)N * L;
  gcc_launch(attn_softmax_kernel, (unsigned)((rows + 7) / 8), 256, 0, st, (const float*)energy, (bf16*)probs, rows, L);
  GCC_CHECK_LAUNCH();
  gcc_launch(attn_transpose_kernel, dim3((L + 31) / 32, (C + 31) / 32, N), 256, 0, st, (const bf16*)v, vT, L, Cp, C);
  GCC_CHECK_LAUNCH();
  // out[n][i][c] = sum_j probs[n][i][j] v[n][j][c]: weights = this image's values, transposed to [C][L]
  return gcc_conv_gemm_launch(probs, N, 1, L, L, vT, C, 1, L, nullptr, out, 1, L, Cp, 0, 0, 1, 1, 1, 0, 0, 0.f, 1, nullptr, 0,
                              nullptr, 0, 0, 0, stream);
}
extern "C" int gcc_attn_bwd_bf16(const void* q, const void* k, const void* v, const void* probs, const void* dout, int N,
                                 int L, int d, int dp, int C, int Cp, void* de_scratch, void* dq, void* dk, void* dv,
                                 void* ws, long long ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if ((L % 8) || (dp % 8) || (Cp % 8) || d > dp || C > Cp || ws == nullptr ||
      ws_bytes < gcc_attn_workspace_bytes(N, L, dp, Cp, 1)) {
    gcc_set_error(__FILE__, __LINE__, "attention: L, dp, Cp must be multiples of 8 and the workspace large enough");
    return GCC_ERR_ARG;
  }
  const long long nll = (long long)N * L * L, rows = (long long)N * L;
  char* w8 = (char*)ws;
  float* dprobs = (float*)w8;
  w8 += al256(nll * 4);
  bf16* kT = (bf16*)w8;
  w8 += al256((long long)N * dp * L * 2);
  float* dv32 = (float*)w8;
  w8 += al256(rows * Cp * 4);
  float* dk32 = (float*)w8;
  // dprobs[n][i][j] = sum_c dout[n][i][c] v[n][j][c]
  int rc = gcc_conv_gemm_launch(dout, N, 1, L, Cp, v, L, 1, Cp, nullptr, dprobs, 1, L, L, 0, 0, 1, 1, 1, 0, 0, 0.f, 1, dprobs,
                                nll, nullptr, 0, 1, 0, stream);
  if (rc) return rc;
  gcc_launch(attn_softmax_bwd_kernel, (unsigned)((rows + 7) / 8), 256, 0, st, (const bf16*)probs, (const float*)dprobs,
             (bf16*)de_scratch, rows, L);
  GCC_CHECK_LAUNCH();
  // dq[n][i][c] = sum_j de[n][i][j] k[n][j][c]: weights = this image's keys transposed to [d][L]
  gcc_launch(attn_transpose_kernel, dim3((L + 31) / 32, (d + 31) / 32, N), 256, 0, st, (const bf16*)k, kT, L, dp, d);
  GCC_CHECK_LAUNCH();
  rc = gcc_conv_gemm_launch(de_scratch, N, 1, L, L, kT, d, 1, L, nullptr, dq, 1, L, dp, 0, 0, 1, 1, 1, 0, 0, 0.f, 1, nullptr, 0,
                            nullptr, 0, 0, 0, stream);
  if (rc) return rc;
  // dk[n][j][c] = sum_i de[n][i][j] q[n][i][c];  dv[n][j][c] = sum_i probs[n][i][j] dout[n][i][c]: contractions over
  // the query positions = the batched weight-gradient GEMM (MN-major operands straight from the row-major tensors)
  rc = gcc_wgrad_gemm_bf16(de_scratch, N, 1, L, L, q, 1, L, dp, dk32, L, dp, 1, 1, 1, 0, 1, 0, 1.f, stream);
  if (rc) return rc;
  rc = gcc_wgrad_gemm_bf16(probs, N, 1, L, L, dout, 1, L, Cp, dv32, L, Cp, 1, 1, 1, 0, 1, 0, 1.f, stream);
  if (rc) return rc;
  gcc_launch(attn_f32_to_bf16_kernel, sa_blocks(rows * dp), 256, 0, st, (const float*)dk32, (bf16*)dk, rows, dp, dp, d);
  GCC_CHECK_LAUNCH();
  gcc_launch(attn_f32_to_bf16_kernel, sa_blocks(rows * Cp), 256, 0, st, (const float*)dv32, (bf16*)dv, rows, Cp, Cp, C);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_scale_add_bf16(const void* a, const void* x, const float* gamma_dev, void* y, long long n,
                                  void* stream) {
  gcc_launch(scale_add_kernel, sa_blocks(n), 256, 0, (cudaStream_t)stream, (const bf16*)a, (const bf16*)x, gamma_dev, (bf16*)y, n);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_scale_add_bwd_bf16(const void* dy, const void* a, const float* gamma_dev, void* da, float* dgamma,
                                      long long n, void* stream) {
  int b = sa_blocks(n);
  if (b > 148 * 2) b = 148 * 2;
  gcc_launch(scale_add_bwd_kernel, b, 256, 0, (cudaStream_t)stream, (const bf16*)dy, (const bf16*)a, gamma_dev, (bf16*)da, dgamma, n);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
