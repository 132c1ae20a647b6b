// Optimizer-side kernels: fused Adam over a flat parameter arena (torch.optim.Adam semantics as
// used at models/Pix2Pix.py:382,415,430,431,440: no weight decay, no amsgrad, eps 1e-8), the
// L1-sparsity gradient term (models/Pix2Pix.py:554-563) and the table-driven bf16 weight re-pack
// that follows every optimizer step.
#include "common.cuh"

namespace gcc {

// hyper (device memory): [0] lr, [1] beta1, [2] beta2, [3] eps, [4] step (as float bits of int)
__global__ void adam_tick_kernel(float* hyper) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  int* step = reinterpret_cast<int*>(hyper + 4);
  *step += 1;
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, const float* __restrict__ hyper) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3];
  const int step = *reinterpret_cast<const int*>(hyper + 4);
  // bias corrections in double: 1 - beta^t = -expm1(t * log(beta))
  const float bc1 = (float)(-expm1((double)step * log((double)b1)));
  const float bc2_sqrt = (float)sqrt(-expm1((double)step * log((double)b2)));
  const float step_size = lr / bc1;
  const long long nvec = n / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
#define GCC_ADAM1(F)                                        \
  mm.F = b1 * mm.F + (1.f - b1) * gg.F;                     \
  vv.F = b2 * vv.F + (1.f - b2) * gg.F * gg.F;              \
  pp.F -= step_size * mm.F / (sqrtf(vv.F) / bc2_sqrt + eps);
    GCC_ADAM1(x) GCC_ADAM1(y) GCC_ADAM1(z) GCC_ADAM1(w)
#undef GCC_ADAM1
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // tail
  for (long long i = nvec * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step_size * mi / (sqrtf(vi) / bc2_sqrt + eps);
  }
}

__global__ void l1_sparsity_kernel(const float* __restrict__ w, float* __restrict__ g, long long n, float lambda) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = w[i];
    g[i] += lambda * (x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f));
  }
}

__global__ void clamp_kernel(float* __restrict__ x, long long n, float lo, float hi) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = fminf(fmaxf(x[i], lo), hi);
}

// table entry (8 x int64): src, direct, transposed, D0, T, D1, D1p, D0p
// The 64 x 64 (d0 x d1) tiles of ALL entries form one flat index space walked grid-stride (layers differ in size
// by four orders of magnitude: a per-layer grid leaves most CTAs idle).  Per tile: 16-byte fp32 reads along d1 (256 B
// per row), 8-byte bf16 writes of the direct pack along d1 and, through a shared-memory transpose, of the transposed
// pack along d0 (128 B per row).  (Round 1 used 32 x 32 tiles with scalar accesses: 4 KB per CTA iteration behind a
// binary search, three 64-bit divisions and two barriers ran at 1.9 TB/s = 227 us for the 54 M-parameter arena.)
static constexpr int kPackTile = 64;
__device__ __forceinline__ void store_bf16x4(bf16* dst, float a, float b, float c, float d, int valid) {
  if (valid == 4) {
    *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16(a, b), pack_bf16(c, d));
  } else {
    const float v[4] = {a, b, c, d};
    for (int e = 0; e < valid; ++e) dst[e] = __float2bfloat16(v[e]);
  }
}
__global__ void __launch_bounds__(256) pack_table_kernel(const long long* __restrict__ table, int count) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  __shared__ float tile[kPackTile][kPackTile + 1];
  extern __shared__ long long cum[];  // [count + 1] running tile counts
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    const long long* e = table + (long long)i * 8;
    cum[i + 1] = ((e[3] + kPackTile - 1) / kPackTile) * ((e[5] + kPackTile - 1) / kPackTile) * e[4];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    cum[0] = 0;
    for (int i = 0; i < count; ++i) cum[i + 1] += cum[i];
  }
  __syncthreads();
  const long long total = cum[count];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4 columns x 4 rows each
  for (long long flat = blockIdx.x; flat < total; flat += gridDim.x) {
    int lo = 0, hi = count - 1;  // last entry with cum[entry] <= flat
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (cum[mid] <= flat) lo = mid; else hi = mid - 1;
    }
    const long long* e = table + (long long)lo * 8;
    const float* src = reinterpret_cast<const float*>(e[0]);
    bf16* direct = reinterpret_cast<bf16*>(e[1]);
    bf16* transposed = reinterpret_cast<bf16*>(e[2]);
    const int D0 = (int)e[3], T = (int)e[4], D1 = (int)e[5], D1p = (int)e[6], D0p = (int)e[7];
    const int t1 = (D1 + kPackTile - 1) / kPackTile;
    const int tile_id = (int)(flat - cum[lo]);  // (a layer has far fewer than 2^31 tiles)
    const int b1 = tile_id % t1;
    const int r = tile_id / t1;
    const int t = r % T;
    const int b0 = r / T;
    const bool vec = (D1 % 4) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d0 = b0 * kPackTile + ty + j * 16, d1 = b1 * kPackTile + tx * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (d0 < D0 && d1 < D1) {
        const long long row = (long long)d0 * T + t;
        const int valid = min(4, D1 - d1);
        if (vec) {  // D1 % 4 == 0: whole float4
          v = *reinterpret_cast<const float4*>(src + row * D1 + d1);
        } else {
          const float* sp = src + row * D1 + d1;
          v.x = sp[0];
          if (valid > 1) v.y = sp[1];
          if (valid > 2) v.z = sp[2];
          if (valid > 3) v.w = sp[3];
        }
        if (direct) store_bf16x4(direct + row * D1p + d1, v.x, v.y, v.z, v.w, valid);
      }
      float* tr = &tile[ty + j * 16][tx * 4];
      tr[0] = v.x; tr[1] = v.y; tr[2] = v.z; tr[3] = v.w;
    }
    __syncthreads();
    if (transposed) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int d1 = b1 * kPackTile + ty + j * 16, d0 = b0 * kPackTile + tx * 4;
        if (d0 < D0 && d1 < D1) {
          const int l1 = ty + j * 16, l0 = tx * 4;
          store_bf16x4(transposed + ((long long)d1 * T + t) * D0p + d0, tile[l0][l1], tile[l0 + 1][l1], tile[l0 + 2][l1],
                       tile[l0 + 3][l1], min(4, D0 - d0));
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace gcc

using namespace gcc;

extern "C" int gcc_adam_step_f32(float* p, const float* g, float* m, float* v, long long n, float* hyper_dev,
                                 void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  gcc_launch(adam_tick_kernel, 1, 1, 0, st, hyper_dev);
  GCC_CHECK_LAUNCH();
  long long b = (n / 4 + 255) / 256;
  if (b < 1) b = 1;
  if (b > 148 * 8) b = 148 * 8;
  gcc_launch(adam_kernel, (unsigned)b, 256, 0, st, p, g, m, v, n, hyper_dev);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_l1_sparsity_f32(const float* w, float* g, long long n, float lambda, void* stream) {
  long long b = (n + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  gcc_launch(l1_sparsity_kernel, (unsigned)b, 256, 0, (cudaStream_t)stream, w, g, n, lambda);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
extern "C" int gcc_clamp_f32(float* x, long long n, float lo, float hi, void* stream) {
  long long b = (n + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  gcc_launch(clamp_kernel, (unsigned)b, 256, 0, (cudaStream_t)stream, x, n, lo, hi);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
// table: device int64 [count][8] = {src fp32 ptr, direct bf16 ptr, transposed bf16 ptr, D0, T, D1, D1p, D0p}
extern "C" int gcc_pack_weights_table(const void* table_dev, int count, void* stream) {
  if (count <= 0) return GCC_OK;
  gcc_launch(pack_table_kernel, 148 * 8, 256, sizeof(long long) * (count + 1), (cudaStream_t)stream, 
      (const long long*)table_dev, count);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}
