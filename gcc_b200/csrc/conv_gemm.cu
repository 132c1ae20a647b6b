// tcgen05 implicit-GEMM convolution kernels for sm_100a (B200).
//
// Two kernels cover every dense contraction of the GCC training step
// (reference call sites: nn.Conv2d / nn.ConvTranspose2d in models/Pix2Pix.py:31-56,
// 216-260, 280-300 and the Gram bmm in models/Pix2Pix.py:733-740):
//
//  * conv_gemm_persistent_kernel ("pixel-major" GEMM):  Y[pix, r] = sum_{tap, c} X[pix (+) tap, c] * Wp[r][tap][c]
//      A = activation patches, fetched by TMA straight from the NHWC bf16 tensor as a 4-D box
//          (64 channels x Wt x Ht x Nt pixels, zero fill outside the image = conv zero padding),
//      B = packed weights [rows][taps][channels] (K-major), D = 128 pixels x BLOCK_N rows in TMEM (double buffered).
//      Persistent (one CTA per SM walks the tiles), 320 threads: warp 0 = TMA producer, warp 1 = TMEM alloc +
//      single-thread MMA issuer, warps 2..9 = epilogue.  Used for Conv2d fprop, Conv2d dgrad, ConvTranspose2d
//      fprop/dgrad, 1x1 convs, Gram backward.  Stride-2 gathers use four parity views of the input (one tensor
//      map each); stride-2 scatters (transposed conv) run as four sub-pixel classes of the same launch.
//
//  * wgrad_gemm_kernel ("channel-major" GEMM): dW[r][tap][c] = sum_pix P[pix, r] * Q[pix (+) tap, c]
//      both operands are MN-major tiles (64 pixels x 64 channels boxes), D = 128|256 r x BLOCK_N c.
//      192 threads: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = epilogue (fp32 stores / reductions).
//      Used for all weight gradients and (batched, tap = 0, P = Q) for the Gram matrices.
//
// smem rings of STAGES slots guarded by full/empty mbarriers; accumulator hand-off through tmem_full (/ tmem_empty)
// mbarriers; every mbarrier wait is bounded and traps instead of hanging.
#include "common.cuh"
#include <stdlib.h>

namespace gcc {

static constexpr int kMaxTaps = 81;  // up to 9 x 9 (SRGAN's first / last conv)
static constexpr int kBlockM = 128;
static constexpr int kBlockK = 64;  // bf16 elements = 128 bytes = one SWIZZLE_128B row

struct GemmGeom {
  CUtensorMap a_maps[4];  // A-side activation views (parity variants for stride-2 gathers)
  CUtensorMap b_map;      // conv_gemm: packed weights (3-D).  wgrad: unused
  CUtensorMap p_map;      // wgrad: M-side activation (4-D).  conv_gemm: unused
  int num_taps;
  int k_chunks;  // conv_gemm: ceil(C / 64) channel chunks per tap
  int w_per_image;  // conv_gemm: 1 = weights [N][R][C] (1x1 only), tap index = image index
  short tap_map[kMaxTaps];
  short tap_dh[kMaxTaps];
  short tap_dw[kMaxTaps];
  short tap_widx[kMaxTaps];
  // pixel-tile geometry (powers of two): tile = Nt x Ht x Wt pixels
  int log_wt, log_ht, log_nt;
  int tiles_w, tiles_h, tiles_n;
  int GN, GH, GW;  // valid extents of the pixel grid
  // conv_gemm output (bf16 NHWC, possibly strided / channel-offset)
  bf16* out;
  long long out_sn, out_sh, out_sw;  // element strides of grid coords (n, a, b)
  int out_cols;                      // number of physical channels to write (multiple of 8)
  int bias_cols;                     // logical rows that receive bias
  const float* bias;
  int act;  // 0 none, 1 leaky-relu(slope), 2 tanh
  float slope;
  // wgrad output (fp32 [batch][R][T][C])
  float* dw;
  int R, C, T_total;     // logical sizes of the fp32 output
  int splits;            // split-K factor over pixel blocks
  int batched;           // 1: one output matrix per image n (Gram)
  int atomic_out;        // 1: red.add into dw, 0: plain store
  float scale;
  int debug;             // timing experiments: bit0 skip epilogue stores
  int c8;                // wgrad: 1 = Q is an 8-channel image, the N dimension is (tap, channel) = T * 8 columns
};

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  if (act == 1) return v > 0.f ? v : v * slope;
  if (act == 2) return tanhf(v);
  return v;
}

// ---------------------------------------------------------------------------------------------
// dW[r][tap][c] (+)= scale * sum_pix P[pix, r] * Q[pix (+) tap, c]
// grid = (r tiles of 128, c tiles of BLOCK_N, taps * splits [* batch])
// MT = number of 128-row accumulators per CTA (MT = 2: a 256 x BLOCK_N tile, one third less L2->SM traffic per
// flop, which is what bounds this kernel: both operands are activations streamed from L2).
// kC8 = image mode (GemmGeom::c8) as a compile-time switch (see conv_gemm_persistent_kernel: the single-thread loops
// are instruction-bound on the small tiles).
template <int BLOCK_N, int STAGES, int MT, bool kC8>
__global__ void __launch_bounds__(192) wgrad_gemm_kernel(const __grid_constant__ GemmGeom p) {
  constexpr uint32_t kBoxBytes = 64 * 128;  // 64 pixels x 64 channels bf16
  constexpr uint32_t kABytes = 2 * MT * kBoxBytes;
  constexpr uint32_t kBBytes = (BLOCK_N / 64) * kBoxBytes;
  constexpr uint32_t kStageBytes = kABytes + kBBytes;
  constexpr uint32_t kTmemCols = MT * BLOCK_N;
  constexpr uint32_t kIdesc = make_idesc_bf16(kBlockM, BLOCK_N, 1, 1);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (base & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int r_tile = blockIdx.x, c_tile = blockIdx.y;
  int bz = blockIdx.z;
  const int split = bz % p.splits;
  bz /= p.splits;
  const int tap = bz % p.num_taps;
  const int img = bz / p.num_taps;  // only meaningful when batched

  // pixel blocks handled by this CTA
  const int tiles_n = p.batched ? 1 : p.tiles_n;
  const int total_pb = p.tiles_w * p.tiles_h * tiles_n;
  const int per = (total_pb + p.splits - 1) / p.splits;
  const int pb_begin = split * per;
  const int pb_end = min(total_pb, pb_begin + per);
  const int num_kb = max(0, pb_end - pb_begin);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a_maps[0]);
    tma_prefetch_desc(&p.p_map);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // programmatic dependent launch: barrier init, tensor-map prefetch and the TMEM allocation above overlapped the
  // previous kernel's tail; no global memory has been touched yet
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // The whole producer warp stays converged: lane 0 waits for the slot and arms the barrier, then lanes
    // 0 .. (2 + BLOCK_N/64 - 1) each issue ONE of the 64-channel boxes, so the per-thread TMA issue latency of
    // up to six boxes per k-block overlaps instead of serialising.
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t smem0 = smem_u32(smem), full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
    int dh = p.tap_dh[tap], dw = p.tap_dw[tap];
    const CUtensorMap* qmap = &p.a_maps[p.tap_map[tap]];
    constexpr int kBoxes = 2 * MT + BLOCK_N / 64;
    if (kC8 && lane >= 2 * MT && lane < 2 * MT + BLOCK_N / 8) {  // image mode: this lane owns tap (lane - 2 MT)
      const int tq = lane - 2 * MT;
      dh = p.tap_dh[tq];
      dw = p.tap_dw[tq];
      qmap = &p.a_maps[p.tap_map[tq]];
    }
    // this lane's box: destination offset inside a stage, tensor map, channel coordinate, spatial offsets
    const bool is_p = lane < 2 * MT;
    const bool active = kC8 ? (lane < 2 * MT + BLOCK_N / 8) : (lane < kBoxes);
    const uint32_t dst_off = is_p ? lane * kBoxBytes
                                  : (kC8 ? kABytes + (lane - 2 * MT) * 1024 : kABytes + (lane - 2 * MT) * kBoxBytes);
    const CUtensorMap* map = is_p ? &p.p_map : qmap;
    const int c0 = is_p ? r_tile * (128 * MT) + lane * 64 : (kC8 ? 0 : c_tile * BLOCK_N + (lane - 2 * MT) * 64);
    const int ob = is_p ? 0 : dw, oa = is_p ? 0 : dh;
    // pixel-block coordinates advance incrementally (no integer division in the steady state)
    int tw = pb_begin % p.tiles_w;
    int th = (pb_begin / p.tiles_w) % p.tiles_h;
    int tn = pb_begin / (p.tiles_w * p.tiles_h);
    for (int pb = pb_begin; pb < pb_end; ++pb) {
      const int b0 = tw << p.log_wt, a0 = th << p.log_ht;
      const int n0 = p.batched ? img : (tn << p.log_nt);
      if (++tw == p.tiles_w) {
        tw = 0;
        if (++th == p.tiles_h) { th = 0; ++tn; }
      }
      if (lane == 0) {
        mbar_wait_s(empty0 + stage * 8, phase ^ 1);
        mbar_expect_tx_s(full0 + stage * 8, kStageBytes);
      }
      __syncwarp();
      if (active) tma_load_4d_s(smem0 + stage * kStageBytes + dst_off, map, full0 + stage * 8, c0, b0 + ob, a0 + oa, n0);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    if (elect_one()) {  // (elect.sync: straight-line MMA issue, see conv_gemm_persistent_kernel)
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t smem0 = smem_u32(smem), full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
      // MN-major SWIZZLE_128B operands: 8-pixel groups are 1024 B apart (SBO), 64-channel atoms 8192 B apart (LBO), a K
      // step (16 pixels) is 2048 B.  Image mode B: a k-step = 16 pixels = two 8-pixel core-matrix rows 128 B apart (LBO),
      // the taps (8-column groups of N) are 1 KB apart (SBO), K step 256 B.
      constexpr uint32_t kSwHi = 0x40004040u, kSwLo = (kBoxBytes >> 4) << 16, kSwStep = 2048 >> 4;
      constexpr uint32_t kBHi = kC8 ? (0x4000u | (1024u >> 4)) : kSwHi;
      constexpr uint32_t kBLo = kC8 ? (128u >> 4) << 16 : kSwLo;
      constexpr uint32_t kBStep = kC8 ? 256 >> 4 : kSwStep;
      uint32_t accumulate = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait_s(full0 + stage * 8, phase);
        tc_fence_after();
        const uint32_t alo = ((smem0 + stage * kStageBytes) >> 4) | kSwLo;
        const uint32_t blo = ((smem0 + stage * kStageBytes + kABytes) >> 4) | kBLo;
#pragma unroll
        for (int m = 0; m < MT; ++m) {
#pragma unroll
          for (int k = 0; k < 64 / 16; ++k)
            umma_bf16(tmem_base + m * BLOCK_N, ((uint64_t)kSwHi << 32) | (alo + m * ((2 * kBoxBytes) >> 4) + k * kSwStep),
                      ((uint64_t)kBHi << 32) | (blo + k * kBStep), kIdesc, k == 0 ? accumulate : 1u);
        }
        accumulate = 1;
        umma_commit_s(empty0 + stage * 8);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tmem_full_bar);
    }
  } else if (num_kb > 0) {
    const int q = warp & 3;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int m = 0; m < MT; ++m) {
      const int r = r_tile * (128 * MT) + m * 128 + q * 32 + lane;
      const bool valid = r < p.R;
      float* orow = p.dw + ((long long)(p.batched ? img : 0) * p.R + r) * ((long long)p.T_total * p.C) +
                    (long long)p.tap_widx[tap] * p.C;
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        const int col0 = c_tile * BLOCK_N + c0;
        if (col0 >= p.C) break;
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + m * BLOCK_N + c0, v);
        tmem_ld_wait();
        if (valid && !(p.debug & 1)) {
          // 16-byte vector reductions / stores where the row segment is aligned (C % 4 == 0), scalar otherwise
          const bool vec = ((reinterpret_cast<uintptr_t>(orow + col0) & 15) == 0);
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const int col = col0 + i;
            const float x0 = __uint_as_float(v[i]) * p.scale, x1 = __uint_as_float(v[i + 1]) * p.scale;
            const float x2 = __uint_as_float(v[i + 2]) * p.scale, x3 = __uint_as_float(v[i + 3]) * p.scale;
            if (vec && col + 3 < p.C) {
              if (p.atomic_out)
                red_add_v4(orow + col, x0, x1, x2, x3);
              else
                *reinterpret_cast<float4*>(orow + col) = make_float4(x0, x1, x2, x3);
            } else {
              const float xs[4] = {x0, x1, x2, x3};
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (col + j < p.C) {
                  if (p.atomic_out)
                    atomicAdd(orow + col + j, xs[j]);
                  else
                    orow[col + j] = xs[j];
                }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------------------------
// The implicit-GEMM convolution kernel (persistent).
//  * grid = min(#tiles, #SMs); each CTA walks tiles  tile = blockIdx.x + i * gridDim.x
//  * the smem ring keeps running across tile boundaries (the TMA producer never drains)
//  * TWO accumulator stages in TMEM (2 x BLOCK_N columns): the 8 epilogue warps drain tile i while the MMA
//    warp already accumulates tile i+1  (tmem_full / tmem_empty mbarriers)
//  * all sub-pixel classes of a stride-2 transposed conv are tiles of ONE launch
//  * optional split-K (tiny pixel counts, long K: the U-Net's inner levels): fp32 red.add into a workspace,
//    bias/activation/bf16 conversion by splitk_finalize_kernel
struct ConvGeom2 {
  CUtensorMap a_maps[4];
  CUtensorMap b_map;
  CUtensorMap out_maps[4];  // per sub-pixel class: bf16 output view (BLOCK_N <= 128 kernels store tiles by TMA)
  int num_classes;
  int cls_tap_begin[5];
  int cls_GH[4], cls_GW[4];
  int cls_tiles_w[4], cls_tiles_h[4];
  int cls_mtile_begin[5];
  long long cls_out_off[4];
  long long cls_part_off[4];
  short tap_map[kMaxTaps];
  short tap_dh[kMaxTaps];
  short tap_dw[kMaxTaps];
  short tap_widx[kMaxTaps];
  int k_chunks, w_per_image;
  int log_wt, log_ht, log_nt, GN;
  int n_tiles, k_splits, total_tiles;
  bf16* out;
  long long out_sn, out_sh, out_sw;
  int out_cols, bias_cols;
  const float* bias;
  int act;
  float slope;
  float* partial;
  long long part_sn, part_sh, part_sw;
  int debug;  // bit0: skip global stores, bit1: skip TMEM loads (timing experiments only)
  int f32_out;  // 1: the result stays fp32 in `partial` (red.add into the zeroed workspace), no bf16 output
  float* stats;  // optional fp32 [2][stats_ld]: per-output-channel sum / sum of squares of the bf16 outputs
  int stats_ld;
  // 1: "image" mode for 8-channel inputs (first conv of the U-Net / PatchGAN).  One k-block = 8 taps x 8 channels:
  // eight 2 KB boxes (128 pixels x 16 B, un-swizzled) land as canonical core matrices, K index = tap * 8 + c,
  // which is exactly the order of the packed weights [R][T][8] read as [R][T * 8].
  int c8;
  // Tail-wave split: a layer whose tile count leaves the last wave of the persistent grid less than half full (512
  // tiles on 148 SMs = 3.46 waves: the PatchGAN 1024 -> 512 data gradient) cuts the K loop of those last tiles into
  // `tail_splits` parts, so the last wave takes 1 / tail_splits of a tile time.  Work items >= tail_begin are
  // (tile, split) pairs; their fp32 partial tiles go to tail_ws [item][128][BLOCK_N] with plain stores (no atomics, no
  // memset) and conv_tail_finalize_kernel sums the parts in a fixed order, adds bias / activation / statistics and
  // writes the bf16 output.  tail_begin == total_tiles: no tail.
  int tail_begin, tail_splits;
  float* tail_ws;
};

struct TileInfo {
  int cls, n_tile, a0, b0, n0, tap0, kb0, kb1;
  int item;  // >= 0: tail work item (index into tail_ws), -1: an ordinary tile
};

__device__ __forceinline__ TileInfo decode_tile(const ConvGeom2& p, int tile_id) {
  TileInfo t;
  int r = tile_id, split, nsplit;
  if (tile_id >= p.tail_begin) {  // tail work item (k_splits == 1 whenever a tail exists)
    t.item = tile_id - p.tail_begin;
    nsplit = p.tail_splits;
    split = t.item % nsplit;
    r = p.tail_begin + t.item / nsplit;
  } else {
    t.item = -1;
    nsplit = p.k_splits;
    split = r % nsplit;
    r /= nsplit;
  }
  const int m_total = p.cls_mtile_begin[p.num_classes];
  int ml = r % m_total;
  t.n_tile = r / m_total;
  int cls = 0;
  while (cls + 1 < p.num_classes && ml >= p.cls_mtile_begin[cls + 1]) ++cls;
  t.cls = cls;
  ml -= p.cls_mtile_begin[cls];
  const int tw = ml % p.cls_tiles_w[cls];
  ml /= p.cls_tiles_w[cls];
  const int th = ml % p.cls_tiles_h[cls];
  const int tn = ml / p.cls_tiles_h[cls];
  t.b0 = tw << p.log_wt;
  t.a0 = th << p.log_ht;
  t.n0 = tn << p.log_nt;
  t.tap0 = p.cls_tap_begin[cls];
  const int ntap = p.cls_tap_begin[cls + 1] - t.tap0;
  const int num_kb = p.c8 ? ntap / 8 : ntap * p.k_chunks;
  const int per = (num_kb + nsplit - 1) / nsplit;
  t.kb0 = split * per;
  t.kb1 = min(num_kb, t.kb0 + per);
  return t;
}

// kC8 = image mode (ConvGeom2::c8) as a compile-time switch: the producer / MMA loops are single-thread instruction
// streams (~500 cycles per k-block before the clean-up below, measured with the pipeline stages switched off:
// profiles/r02_conv_pipeline_bound.txt), which is what bounded every BLOCK_N <= 128 layer, so nothing that can be
// decided at compile time or per tile is left inside them.
template <int BLOCK_N, int STAGES, bool kTmaStore, bool kC8>
__global__ void __launch_bounds__(320, 1) conv_gemm_persistent_kernel(const __grid_constant__ ConvGeom2 p) {
  constexpr uint32_t kABytes = kBlockM * 128;
  constexpr uint32_t kBBytes = BLOCK_N * 128;
  constexpr uint32_t kStageBytes = kABytes + kBBytes;
  constexpr uint32_t kTmemCols = 2 * BLOCK_N;
  constexpr uint32_t kIdesc = make_idesc_bf16(kBlockM, BLOCK_N, 0, 0);
  constexpr int kHalf = BLOCK_N / 2;  // columns per epilogue warp group

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (base & 1023u)) & 1023u);
  // kTmaStore (BLOCK_N <= 128 and a short K loop: store-bound layers): the bf16 tile is staged in 128B-swizzled
  // 64-column units (128 rows x 128 B = 16 KB each, double buffered) and written by TMA; otherwise (long K loops,
  // which want the shared memory for a deeper operand ring) warp-private staging below.
  static_assert(!kTmaStore || BLOCK_N <= 128, "the TMA-store epilogue stages at most two 64-column units");
  constexpr int kUnits = BLOCK_N / 64;
  constexpr uint32_t kOutBytes = kTmaStore ? 2 * kUnits * 16384 : 0;
  uint8_t* out_base = smem + STAGES * kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(out_base + kOutBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;    // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;    // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  constexpr int kStagePitch = 80;  // 64 B of data + 16 B pad per staged row (spreads banks)
  uint8_t* stage_base = out_base + kOutBytes + 256;                    // 8 warps x 32 rows x 80 B
  float* bias_base = reinterpret_cast<float*>(stage_base + 8 * 32 * kStagePitch);  // 8 x 32 floats (round 1's per-warp bias slots, now unused: tbias below)
  float* sstat = bias_base + 8 * 32;                                                // [2][BLOCK_N] BN statistics
  // biases of the current block of output columns: loaded ONCE per column-block change by the epilogue threads (the
  // per-chunk __ldg they replace put an L2 round trip on the critical path of every 32-column chunk of every tile of
  // the short-K, store-bound layers); zeros when the conv has no bias
  float* tbias = sstat + 2 * BLOCK_N;                                               // [BLOCK_N]
  for (int i = threadIdx.x; i < BLOCK_N; i += blockDim.x) tbias[i] = 0.f;
  if (p.stats != nullptr)
    for (int i = threadIdx.x; i < 2 * BLOCK_N; i += blockDim.x) sstat[i] = 0.f;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a_maps[0]);
    tma_prefetch_desc(&p.b_map);
    if (kTmaStore) tma_prefetch_desc(&p.out_maps[0]);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 8);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // programmatic dependent launch: barrier init, tensor-map prefetch and the TMEM allocation above overlapped the
  // previous kernel's tail; no global memory has been touched yet
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // One ELECTED thread (elect.sync, not `lane == 0`: the compiler then knows the block is executed by exactly one
    // thread and emits the TMA / MMA / commit instructions straight instead of inside per-instruction waterfall loops).
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t smem0 = smem_u32(smem), full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const TileInfo t = decode_tile(p, tile);
        if (kC8) {
          for (int kb = t.kb0; kb < t.kb1; ++kb) {
            mbar_wait_s(empty0 + stage * 8, phase ^ 1);
            mbar_expect_tx_s(full0 + stage * 8, kStageBytes);
            const uint32_t sa = smem0 + stage * kStageBytes;
#pragma unroll 1
            for (int j = 0; j < 8; ++j) {
              const int tap = t.tap0 + kb * 8 + j;
              tma_load_4d_s(sa + j * (kBlockM * 16), &p.a_maps[p.tap_map[tap]], full0 + stage * 8, 0,
                            t.b0 + p.tap_dw[tap], t.a0 + p.tap_dh[tap], t.n0);
            }
            tma_load_3d_s(sa + kABytes, &p.b_map, full0 + stage * 8, kb * kBlockK, 0, t.n_tile * BLOCK_N);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        } else {
          // k-block kb = (tap tl, channel chunk kc): the tap's table entries are read once per tap, not per k-block
          int tl = t.kb0 / p.k_chunks;
          int kc = t.kb0 - tl * p.k_chunks;
          const int brow = t.n_tile * BLOCK_N;
          for (int kb = t.kb0; kb < t.kb1; ++tl, kc = 0) {
            const int tap = t.tap0 + tl;
            const CUtensorMap* amap = &p.a_maps[p.tap_map[tap]];
            const int cb = t.b0 + p.tap_dw[tap], ca = t.a0 + p.tap_dh[tap];
            const int wrow = p.w_per_image ? t.n0 : (int)p.tap_widx[tap];
            const int kc_end = min(p.k_chunks, kc + (t.kb1 - kb));
            for (; kc < kc_end; ++kc, ++kb) {
              mbar_wait_s(empty0 + stage * 8, phase ^ 1);
              mbar_expect_tx_s(full0 + stage * 8, kStageBytes);
              const uint32_t sa = smem0 + stage * kStageBytes;
              tma_load_4d_s(sa, amap, full0 + stage * 8, kc * kBlockK, cb, ca, t.n0);
              tma_load_3d_s(sa + kABytes, &p.b_map, full0 + stage * 8, kc * kBlockK, wrow, brow);
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t smem0 = smem_u32(smem), full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
      // Shared-memory matrix descriptors (common.cuh): the upper word is a constant, the lower word = (address >> 4) |
      // (leading byte offset >> 4) << 16, and a K step of 16 elements just adds to the address field.
      //  K-major SWIZZLE_128B (A of the channel mode, B always): LBO 16, SBO 1024, K step 32 bytes.
      //  image mode A: one MMA (K = 16) spans two taps = two core-matrix columns 2 KB apart (LBO); the 8-pixel row
      //  groups of a tap are 128 B apart (SBO); K step = two taps = 4 KB.
      constexpr uint32_t kSwHi = 0x40004040u, kSwLo = 1u << 16, kSwStep = 32 >> 4;
      constexpr uint32_t kAHi = kC8 ? (0x4000u | (128u >> 4)) : kSwHi;
      constexpr uint32_t kALo = kC8 ? ((uint32_t)(kBlockM * 16) >> 4) << 16 : kSwLo;
      constexpr uint32_t kAStep = kC8 ? (2 * kBlockM * 16) >> 4 : kSwStep;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const TileInfo t = decode_tile(p, tile);
        if (t.kb1 <= t.kb0) continue;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        uint32_t accumulate = 0;  // the first MMA of a tile overwrites the accumulator
        for (int kb = t.kb0; kb < t.kb1; ++kb) {
          mbar_wait_s(full0 + stage * 8, phase);
          tc_fence_after();
          const uint32_t alo = ((smem0 + stage * kStageBytes) >> 4) | kALo;
          const uint32_t blo = ((smem0 + stage * kStageBytes + kABytes) >> 4) | kSwLo;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            umma_bf16(tmem_d, ((uint64_t)kAHi << 32) | (alo + k * kAStep), ((uint64_t)kSwHi << 32) | (blo + k * kSwStep),
                      kIdesc, k == 0 ? accumulate : 1u);
          accumulate = 1;
          umma_commit_s(empty0 + stage * 8);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // epilogue warps 2..9: lane quarter = warp % 4, column half = (warp - 2) / 4
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int wt_mask = (1 << p.log_wt) - 1, ht_mask = (1 << p.log_ht) - 1;
    int acc = 0;
    uint32_t acc_phase = 0;
    int cur_nt = -1, bias_nt = -1;
    int tile_iter = 0;
    const int et = threadIdx.x - 64;  // 0..255 among the epilogue threads
    // per-CTA statistics live in smem and are flushed (one global atomic per channel) when the CTA moves on to
    // another block of output channels; tiles are visited with non-decreasing n_tile
    auto flush_stats = [&](int nt) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int i = et; i < 2 * BLOCK_N; i += 256) {
        const int col = nt * BLOCK_N + (i % BLOCK_N);
        if (col < p.out_cols) atomicAdd(p.stats + (i / BLOCK_N) * p.stats_ld + col, sstat[i]);
        sstat[i] = 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    };
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileInfo t = decode_tile(p, tile);
      if (t.kb1 <= t.kb0) continue;
      if (p.stats != nullptr && t.n_tile != cur_nt) {
        if (cur_nt >= 0) flush_stats(cur_nt);
        cur_nt = t.n_tile;
      }
      if (p.bias != nullptr && t.n_tile != bias_nt) {
        asm volatile("bar.sync 1, 256;" ::: "memory");  // every epilogue thread is done with the previous block's biases
        for (int i = et; i < BLOCK_N; i += 256) {
          const int c = t.n_tile * BLOCK_N + i;
          tbias[i] = c < p.bias_cols ? __ldg(p.bias + c) : 0.f;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        bias_nt = t.n_tile;
      }
      const int b = t.b0 + (r & wt_mask);
      const int a = t.a0 + ((r >> p.log_wt) & ht_mask);
      const int n = t.n0 + (r >> (p.log_wt + p.log_ht));
      const bool valid = (n < p.GN) && (a < p.cls_GH[t.cls]) && (b < p.cls_GW[t.cls]);
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BLOCK_N + ((uint32_t)(q * 32) << 16);
      if (t.item >= 0) {
        // tail work item: this thread's row of the fp32 partial tile goes to the workspace as it is (128 contiguous
        // bytes per chunk); conv_tail_finalize_kernel sums the parts and applies bias / activation / statistics
        float* wrow = p.tail_ws + ((long long)t.item * kBlockM + r) * BLOCK_N;
#pragma unroll 1
        for (int c0 = half * kHalf; c0 < (half + 1) * kHalf; c0 += 32) {
          if (t.n_tile * BLOCK_N + c0 >= p.out_cols) break;  // warp-uniform
          uint32_t v[32];
          tmem_ld_32x32(tmem_d + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(wrow + c0 + i) =
                make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                            __uint_as_float(v[i + 3]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        continue;
      }
      if (kTmaStore && p.k_splits == 1 && !p.f32_out) {
        uint8_t* obuf = out_base + (uint32_t)(tile_iter & 1) * (kUnits * 16384);
        // the store issued two tiles ago read this buffer: wait for it, then tell everybody
        if (et == 0) tma_store_wait_read<1>();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
#pragma unroll 1
        for (int c0 = half * kHalf; c0 < (half + 1) * kHalf; c0 += 32) {
          const int col0 = t.n_tile * BLOCK_N + c0;
          if (col0 >= p.out_cols) break;  // warp-uniform
          uint32_t v[32];
          tmem_ld_32x32(tmem_d + c0, v);
          tmem_ld_wait();
          __syncwarp();
          // row r of unit (c0 / 64): 128 B, 16-byte chunks XOR-swizzled with (r & 7) (= CU_TENSOR_MAP_SWIZZLE_128B)
          uint8_t* urow = obuf + (c0 >> 6) * 16384 + r * 128;
          const int ch0 = (c0 & 63) >> 3;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
              f[i] = apply_act(__uint_as_float(v[g * 8 + i]) + tbias[c0 + g * 8 + i], p.act, p.slope);
            *reinterpret_cast<uint4*>(urow + (((ch0 + g) ^ (r & 7)) << 4)) =
                make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
          }
          __syncwarp();
          if (p.stats != nullptr && col0 + lane < p.out_cols) {
            // lane j sums column j of the warp's 32 staged rows (the bf16 values that are stored)
            float s1 = 0.f, s2 = 0.f;
            const uint8_t* ubase = obuf + (c0 >> 6) * 16384 + (q * 32) * 128 + (lane & 7) * 2;
            const int chl = ch0 + (lane >> 3);
#pragma unroll 8
            for (int rr = 0; rr < 32; ++rr) {
              if ((vmask >> rr) & 1u) {
                const float xv = __bfloat162float(
                    *reinterpret_cast<const bf16*>(ubase + rr * 128 + ((chl ^ (rr & 7)) << 4)));
                s1 += xv;
                s2 += xv * xv;
              }
            }
            atomicAdd(&sstat[c0 + lane], s1);
            atomicAdd(&sstat[BLOCK_N + c0 + lane], s2);
          }
        }
        // TMEM reads done: hand the accumulator stage back before the (slower) store hand-off
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et == 0 && !(p.debug & 1)) {
#pragma unroll
          for (int u = 0; u < kUnits; ++u)
            if (t.n_tile * BLOCK_N + u * 64 < p.out_cols)
              tma_store_4d(&p.out_maps[t.cls], obuf + u * 16384, t.n_tile * BLOCK_N + u * 64, t.b0, t.a0, t.n0);
          tma_store_commit();
        }
        ++tile_iter;
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        continue;
      }
      if (p.k_splits == 1 && !p.f32_out) {
        // Each thread owns one pixel row; its 32-column chunk (64 B) is staged in warp-private shared memory and
        // written out with 4 lanes per row, so every store instruction covers whole 32-byte sectors (8 rows x
        // 64 contiguous bytes) instead of 32 half-written sectors 1 row apart.
        bf16* orow = p.out + p.cls_out_off[t.cls] + (long long)n * p.out_sn + (long long)a * p.out_sh +
                     (long long)b * p.out_sw;
        const unsigned long long orow_u = reinterpret_cast<unsigned long long>(orow);
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        uint8_t* stage = stage_base + (warp - 2) * (32 * kStagePitch);
#pragma unroll 1
        for (int c0 = half * kHalf; c0 < (half + 1) * kHalf; c0 += 32) {
          const int col0 = t.n_tile * BLOCK_N + c0;
          if (col0 >= p.out_cols) break;  // warp-uniform
          uint32_t v[32];
          if (!(p.debug & 2)) {
            tmem_ld_32x32(tmem_d + c0, v);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0;
          }
          if (!(p.debug & 2)) tmem_ld_wait();
          __syncwarp();
          uint4* srow = reinterpret_cast<uint4*>(stage + lane * kStagePitch);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
              f[i] = apply_act(__uint_as_float(v[g * 8 + i]) + tbias[c0 + g * 8 + i], p.act, p.slope);
            srow[g] = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                                 pack_bf16(f[6], f[7]));
          }
          __syncwarp();
          if (p.stats != nullptr && col0 + lane < p.out_cols) {
            // lane j sums column j of the warp's 32 staged rows (the bf16 values that are stored)
            float s1 = 0.f, s2 = 0.f;
            const bf16* colp = reinterpret_cast<const bf16*>(stage) + lane;
#pragma unroll 8
            for (int rr = 0; rr < 32; ++rr) {
              if ((vmask >> rr) & 1u) {
                const float xv = __bfloat162float(colp[rr * (kStagePitch / 2)]);
                s1 += xv;
                s2 += xv * xv;
              }
            }
            atomicAdd(&sstat[c0 + lane], s1);
            atomicAdd(&sstat[BLOCK_N + c0 + lane], s2);
          }
          if (!(p.debug & 1)) {
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int piece = it * 32 + lane;
              const int rr = piece >> 2, ch = piece & 3;
              const unsigned long long pr = __shfl_sync(0xffffffffu, orow_u, rr);
              const int col = col0 + ch * 8;
              if (((vmask >> rr) & 1u) && col < p.out_cols) {
                const uint4 val = *reinterpret_cast<const uint4*>(stage + rr * kStagePitch + ch * 16);
                *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(pr) + col) = val;
              }
            }
          }
          __syncwarp();
        }
      } else {
        float* prow = p.partial + p.cls_part_off[t.cls] + (long long)n * p.part_sn + (long long)a * p.part_sh +
                      (long long)b * p.part_sw;
#pragma unroll 1
        for (int c0 = half * kHalf; c0 < (half + 1) * kHalf; c0 += 32) {
          const int col0 = t.n_tile * BLOCK_N + c0;
          if (col0 >= p.out_cols) break;
          uint32_t v[32];
          tmem_ld_32x32(tmem_d + c0, v);
          tmem_ld_wait();
          if (valid) {  // out_cols and the workspace row pitch are multiples of 8: whole 16-byte reductions
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              if (col0 + i < p.out_cols)
                red_add_v4(prow + col0 + i, __uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                           __uint_as_float(v[i + 3]));
          }
        }
      }
      // all TMEM reads of this warp are complete (wait::ld above): hand the accumulator stage back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (p.stats != nullptr && cur_nt >= 0) flush_stats(cur_nt);
    if (kTmaStore && et == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// Completes the tail tiles of conv_gemm_persistent_kernel (ConvGeom2::tail_*): grid = (tail tiles, 128 / kTailRows),
// thread = output column (coalesced in the workspace and in the NHWC output), kTailRows rows per CTA with all their
// loads in flight.  out = bf16(act(sum_parts ws + bias)); the per-channel statistics use the stored bf16 values of the
// valid rows, exactly like the fused epilogue.
static constexpr int kTailRows = 16;
static constexpr int kTailMaxParts = 4;
template <int BLOCK_N>
__global__ void __launch_bounds__(BLOCK_N) conv_tail_finalize_kernel(const __grid_constant__ ConvGeom2 p) {
  pdl_wait();
  pdl_launch_dependents();
  const int item0 = blockIdx.x * p.tail_splits;
  const TileInfo t = decode_tile(p, p.tail_begin + item0);
  const int col = t.n_tile * BLOCK_N + threadIdx.x;
  if (col >= p.out_cols) return;
  // parts with an empty K range were never written
  int nparts = 0;
  for (int s = 0; s < p.tail_splits; ++s) {
    const TileInfo ts = decode_tile(p, p.tail_begin + item0 + s);
    if (ts.kb1 > ts.kb0) nparts = s + 1;
  }
  const float bias = (p.bias != nullptr && col < p.bias_cols) ? __ldg(p.bias + col) : 0.f;
  const int r0 = blockIdx.y * kTailRows;
  const float* __restrict__ ws = p.tail_ws + ((long long)item0 * kBlockM + r0) * BLOCK_N + threadIdx.x;
  float v[kTailRows];
#pragma unroll
  for (int i = 0; i < kTailRows; ++i) {
    v[i] = 0.f;
#pragma unroll
    for (int s = 0; s < kTailMaxParts; ++s)
      if (s < nparts) v[i] += __ldg(ws + ((long long)s * kBlockM + i) * BLOCK_N);
  }
  const int wt_mask = (1 << p.log_wt) - 1, ht_mask = (1 << p.log_ht) - 1;
  bf16* __restrict__ out = p.out + p.cls_out_off[t.cls] + col;
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < kTailRows; ++i) {
    const int r = r0 + i;
    const int b = t.b0 + (r & wt_mask);
    const int a = t.a0 + ((r >> p.log_wt) & ht_mask);
    const int n = t.n0 + (r >> (p.log_wt + p.log_ht));
    if (!((n < p.GN) && (a < p.cls_GH[t.cls]) && (b < p.cls_GW[t.cls]))) continue;  // block-uniform
    const bf16 o = __float2bfloat16(apply_act(v[i] + bias, p.act, p.slope));
    out[(long long)n * p.out_sn + (long long)a * p.out_sh + (long long)b * p.out_sw] = o;
    const float xv = __bfloat162float(o);
    s1 += xv;
    s2 += xv * xv;
  }
  if (p.stats != nullptr) {
    atomicAdd(p.stats + col, s1);
    atomicAdd(p.stats + p.stats_ld + col, s2);
  }
}

// out[pix, y_coff + c] = bf16(act(partial[pix, c] + bias[c]))
__global__ void splitk_finalize_kernel(const float* __restrict__ partial, bf16* __restrict__ out, long long npix, int Rp,
                                       int Cy, int y_coff, int bias_cols, const float* __restrict__ bias, int act,
                                       float slope) {
  pdl_wait();  // programmatic dependent launch: everything above overlapped the previous kernel's tail
  pdl_launch_dependents();
  const long long total = npix * Rp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Rp);
    const long long pix = i / Rp;
    float v = partial[i];
    if (bias != nullptr && c < bias_cols) v += bias[c];
    out[pix * Cy + y_coff + c] = __float2bfloat16(apply_act(v, act, slope));
  }
}

}  // namespace gcc

// =============================================================================== host side
using namespace gcc;

static int ilog2_ceil(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

// choose a power-of-two pixel tile (Wt, Ht, Nt) with Wt*Ht*Nt == pixels
static void choose_tile(int pixels, int GN, int GH, int GW, int* lw, int* lh, int* ln) {
  int lp = ilog2_ceil(pixels);
  int w = ilog2_ceil(GW);
  if (w > lp) w = lp;
  int h = ilog2_ceil(GH);
  if (h > lp - w) h = lp - w;
  int n = lp - w - h;
  *lw = w;
  *lh = h;
  *ln = n;
}

// 4-D activation view: dims (C, W', H', N) over an NHWC bf16 tensor, optionally the stride-2
// parity sub-grid (ph, pw).  box = (64, 2^lw, 2^lh, 2^ln).
static int make_act_map(CUtensorMap* m, const void* x, int N, int H, int W, int C, int sub, int ph, int pw,
                        int lw, int lh, int ln, int c8 = 0) {
  const bf16* base = reinterpret_cast<const bf16*>(x);
  uint64_t dims[4], strides[3];
  if (sub == 1) {
    dims[0] = C; dims[1] = W; dims[2] = H; dims[3] = N;
    strides[0] = (uint64_t)C * 2;
    strides[1] = (uint64_t)W * C * 2;
    strides[2] = (uint64_t)H * W * C * 2;
  } else {
    base += ((long long)ph * W + pw) * C;
    dims[0] = C;
    dims[1] = (W - pw + 1) / 2;
    dims[2] = (H - ph + 1) / 2;
    dims[3] = N;
    strides[0] = (uint64_t)C * 4;
    strides[1] = (uint64_t)W * C * 4;
    strides[2] = (uint64_t)H * W * C * 2;
  }
  uint32_t box[4] = {c8 ? 8u : 64u, 1u << lw, 1u << lh, 1u << ln};
  return gcc_make_tmap_bf16_sw(m, base, 4, dims, strides, box, c8 ? 0 : 1);
}

// Row-window view of an 8-channel NHWC image [N][Hrows][Wp][8] (see elementwise.cu, "stem / head convolutions"):
// dim 0 = 64 contiguous elements = 8 pixels x 8 channels starting at a window position, dim 1 = window position
// (stride ONE pixel = 16 bytes: the windows overlap), dim 2 = image row, dim 3 = image.  A window that starts in the
// last 7 pixels of a row runs into the next row / image (finite data, multiplied by zero weights); the caller keeps
// >= 128 readable bytes behind the last image.
static int make_rowwin_map(CUtensorMap* m, const void* x, int N, int Hrows, int Wp, int lw, int lh, int ln) {
  uint64_t dims[4] = {64, (uint64_t)Wp, (uint64_t)Hrows, (uint64_t)N};
  uint64_t strides[3] = {16, (uint64_t)Wp * 16, (uint64_t)Hrows * Wp * 16};
  uint32_t box[4] = {64u, 1u << lw, 1u << lh, 1u << ln};
  return gcc_make_tmap_bf16_sw(m, x, 4, dims, strides, box, 1);
}

static int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// per-device caches (one process may drive several devices): function attributes and the SM count
static constexpr int kMaxDevices = 64;
static int cur_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}

template <int BLOCK_N, int STAGES, int MT, bool C8 = false>
static int launch_wgrad_gemm(const GemmGeom& g, dim3 grid, cudaStream_t st) {
  const int smem = STAGES * (2 * MT * 8192 + (BLOCK_N / 64) * 8192) + 1024 + 256;
  static bool configured[kMaxDevices] = {};
  const int dev = cur_device();
  if (!configured[dev]) {
    if (cudaFuncSetAttribute(wgrad_gemm_kernel<BLOCK_N, STAGES, MT, C8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             smem) != cudaSuccess) {
      gcc_set_error(__FILE__, __LINE__, "cudaFuncSetAttribute failed");
      return GCC_ERR_CUDA;
    }
    configured[dev] = true;
  }
  gcc_launch(wgrad_gemm_kernel<BLOCK_N, STAGES, MT, C8>, grid, 192, smem, st, g);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}

static int g_force_block_n = 0;  // test hook
extern "C" void gcc_debug_force_block_n(int bn) { g_force_block_n = bn; }

static int pick_block_n(int rows) {
  if (g_force_block_n) return g_force_block_n;
  if (rows <= 64) return 64;
  if (rows <= 128) return 128;
  if (rows % 256 == 0 || rows > 512) return 256;
  return 128;
}

// Fill taps for a gather (conv) relation: in = s*o + k - p.
static int fill_gather_taps(GemmGeom& g, int KH, int KW, int stride, int pad) {
  if (KH * KW > kMaxTaps) return GCC_ERR_ARG;
  g.num_taps = KH * KW;
  for (int kh = 0; kh < KH; ++kh)
    for (int kw = 0; kw < KW; ++kw) {
      const int t = kh * KW + kw;
      const int th = kh - pad, tw = kw - pad;
      if (stride == 1) {
        g.tap_map[t] = 0;
        g.tap_dh[t] = (short)th;
        g.tap_dw[t] = (short)tw;
      } else {
        const int ph = ((th % 2) + 2) % 2, pw = ((tw % 2) + 2) % 2;
        g.tap_map[t] = (short)(ph * 2 + pw);
        g.tap_dh[t] = (short)floordiv(th, 2);
        g.tap_dw[t] = (short)floordiv(tw, 2);
      }
      g.tap_widx[t] = (short)t;
    }
  return GCC_OK;
}

// Tail-wave split of the persistent conv kernel (ConvGeom2::tail_begin): number of K parts (0 = no split) of the last
// `*rem` tiles of a launch of `base_tiles` tiles with >= `min_kb` k-blocks each on `sms` SMs: few waves, a last wave that
// is at most half full, a K loop of >= tail_min_kb (and >= 32) k-blocks so that every part keeps >= 16.  Pure host
// arithmetic (exported for the CPU tests as gcc_plan_conv_tail).
static int plan_conv_tail(int base_tiles, int min_kb, int sms, int tail_min_kb, int* rem_out) {
  const int S = sms;
  const int waves = base_tiles / S, rem = base_tiles % S;
  if (rem_out) *rem_out = rem;
  if (!(waves >= 1 && waves <= 8 && rem > 0 && rem * 2 <= S && min_kb >= 32 && min_kb >= tail_min_kb)) return 0;
  int ks = S / rem;
  if (ks > min_kb / 16) ks = min_kb / 16;
  if (ks > kTailMaxParts) ks = kTailMaxParts;
  return ks >= 2 ? ks : 0;
}
extern "C" int gcc_plan_conv_tail(int base_tiles, int min_kb, int sms, int tail_min_kb) {
  return plan_conv_tail(base_tiles, min_kb, sms, tail_min_kb, nullptr);
}

static int g_debug_flags = 0;
extern "C" void gcc_debug_set_flags(int f) { g_debug_flags = f; }
static int g_tail_min_kb = 0;  // test hook: k-blocks per tile from which the tail-wave split is taken (0 = default)
extern "C" void gcc_debug_set_tail_min_kb(int kb) { g_tail_min_kb = kb; }
// bit 5 of the debug flags: time every GEMM launch with events (serialises the stream) and log shape + time
static cudaEvent_t g_trace_ev[2];
static void trace_begin(cudaStream_t st) {
  if (!(g_debug_flags & 32)) return;
  if (g_trace_ev[0] == nullptr) {
    cudaEventCreate(&g_trace_ev[0]);
    cudaEventCreate(&g_trace_ev[1]);
  }
  cudaEventRecord(g_trace_ev[0], st);
}
static void trace_end(cudaStream_t st, const char* kind, int N, int H, int W, int C, int R, int OH, int OW, int k,
                      int s, int mode, int BN, int tiles, int ksplit, int kb, int stats, double flop) {
  if (!(g_debug_flags & 32)) return;
  cudaEventRecord(g_trace_ev[1], st);
  cudaEventSynchronize(g_trace_ev[1]);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, g_trace_ev[0], g_trace_ev[1]);
  fprintf(stderr, "GCCTRACE %s N=%d H=%d W=%d C=%d R=%d OH=%d OW=%d k=%d s=%d mode=%d BN=%d tiles=%d ksplit=%d kb=%d stats=%d us=%.1f tflops=%.0f\n",
          kind, N, H, W, C, R, OH, OW, k, s, mode, BN, tiles, ksplit, kb, stats, ms * 1e3, flop / (ms * 1e9));
}
static int g_num_sms[kMaxDevices] = {};
static int num_sms() {
  const int dev = cur_device();
  if (g_num_sms[dev] == 0) {
    cudaDeviceGetAttribute(&g_num_sms[dev], cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms[dev] <= 0) g_num_sms[dev] = kNumSMs;
  }
  return g_num_sms[dev];
}

template <int BLOCK_N, int STAGES, bool TS, bool C8>
static int launch_conv_persistent(const ConvGeom2& g, cudaStream_t st) {
  const int smem = STAGES * (kBlockM * 128 + BLOCK_N * 128) + (TS ? 2 * (BLOCK_N / 64) * 16384 : 0) + 1024 + 256 +
                   8 * 32 * 80 + 8 * 32 * 4 + 3 * BLOCK_N * 4;
  static bool configured[kMaxDevices] = {};
  const int dev = cur_device();
  if (!configured[dev]) {
    if (cudaFuncSetAttribute(conv_gemm_persistent_kernel<BLOCK_N, STAGES, TS, C8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             smem) != cudaSuccess) {
      gcc_set_error(__FILE__, __LINE__, "cudaFuncSetAttribute failed");
      return GCC_ERR_CUDA;
    }
    configured[dev] = true;
  }
  const int grid = g.total_tiles < num_sms() ? g.total_tiles : num_sms();
  gcc_launch(conv_gemm_persistent_kernel<BLOCK_N, STAGES, TS, C8>, grid, 320, smem, st, g);
  GCC_CHECK_LAUNCH();
  return GCC_OK;
}

// splitk_ws: optional fp32 workspace of >= N*OH*OW*round8(R) elements; when given, layers with very few pixel
// tiles and a long contraction split K across CTAs.  NULL disables split-K.
// f32_out = 1: no bf16 output; `splitk_ws` (>= N*OH*OW*round8(R) floats, zeroed here) receives the fp32 result with
// row pitch round8(R) (used by the attention scores, whose softmax wants un-rounded logits).
int gcc_conv_gemm_launch(const void* x, int N, int H, int W, int Cx, const void* w, int R, int T, int Cw,
                         const float* bias, void* y, int OH, int OW, int Cy, int y_coff, int transposed, int KH, int KW,
                         int stride, int pad, int act, float slope, int w_per_image, float* splitk_ws,
                         long long ws_elems, float* stats, int stats_ld, int f32_out, int rowwin, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // rowwin = 1: x is a pre-padded 8-channel image [N][H][W][8] followed by >= 128 readable bytes, w = [R][KH * KB][64]
  // with KB = ceil(KW / 8) (gcc_rowwin_weight_pack_bf16), T = KH * KB, Cw = 64, stride 1, pad 0.
  if (rowwin && (transposed || w_per_image || Cx != 8 || Cw != 64 || stride != 1 || pad != 0 || T != KH * ((KW + 7) / 8) ||
                 OH + KH - 1 > H || OW + KW - 1 > W)) {
    gcc_set_error(__FILE__, __LINE__, "gcc_conv_gemm: bad row-window arguments");
    return GCC_ERR_ARG;
  }
  if (f32_out && (splitk_ws == nullptr || stats != nullptr || bias != nullptr || act != 0)) {
    gcc_set_error(__FILE__, __LINE__, "conv gemm: fp32 output needs a workspace and excludes bias / activation / statistics");
    return GCC_ERR_ARG;
  }
  if (stats != nullptr && stats_ld < ((R + 7) / 8 * 8)) {
    gcc_set_error(__FILE__, __LINE__, "gcc_conv_gemm_bf16: fused statistics need stats_ld >= round8(R)");
    return GCC_ERR_ARG;
  }
  if (w_per_image && (KH != 1 || KW != 1 || stride != 1 || pad != 0 || transposed)) {
    gcc_set_error(__FILE__, __LINE__, "gcc_conv_gemm_bf16: per-image weights need a 1x1 stride-1 conv");
    return GCC_ERR_ARG;
  }
  if ((Cx % 8) || (Cw % 8) || (Cy % 8) || (y_coff % 8) || (!rowwin && T != KH * KW) || (stride != 1 && stride != 2) ||
      KH * KW > kMaxTaps) {
    gcc_set_error(__FILE__, __LINE__, "gcc_conv_gemm_bf16: bad arguments");
    return GCC_ERR_ARG;
  }
  // physical output columns: the whole row of y when the conv owns it (y_coff == 0: pad channels are written as
  // zeros, whatever multiple of 8 the caller pads to), round8(R) inside a channel window of a wider buffer
  const int Rp = (y_coff == 0 && Cy >= R) ? Cy : (R + 7) / 8 * 8;
  if (y_coff + Rp > Cy) {
    gcc_set_error(__FILE__, __LINE__, "gcc_conv_gemm_bf16: output channel window out of range");
    return GCC_ERR_ARG;
  }
  if (!transposed && stride == 2 && ((H % 2) || (W % 2))) {
    gcc_set_error(__FILE__, __LINE__, "stride-2 conv needs even H, W");
    return GCC_ERR_ARG;
  }
  const int BN = pick_block_n(Rp);
  const int Ck = rowwin ? 64 : (Cx < Cw ? Cx : Cw);  // contraction extent (both are zero padded to their physical size)

  // 8-channel image input with a multiple of 8 taps (k4 convs): gather taps x channels into K = T * 8
  const int c8 = (!rowwin && !transposed && !w_per_image && Cx == 8 && Cw == 8 && (T % 8) == 0 && !(g_debug_flags & 128)) ? 1 : 0;
  ConvGeom2 g;
  memset(&g, 0, sizeof(g));
  g.c8 = c8;
  g.n_tiles = (Rp + BN - 1) / BN;
  g.k_chunks = (Ck + kBlockK - 1) / kBlockK;
  g.GN = N;
  g.w_per_image = w_per_image;
  const int classes = (transposed && stride == 2) ? 4 : 1;
  const int os = (transposed && stride == 2) ? 2 : 1;
  // tile geometry from the largest class grid
  const int GH0 = (transposed && stride == 2) ? (OH + 1) / 2 : OH;
  const int GW0 = (transposed && stride == 2) ? (OW + 1) / 2 : OW;
  choose_tile(kBlockM, N, GH0, GW0, &g.log_wt, &g.log_ht, &g.log_nt);
  if (w_per_image) {  // a tile must not straddle images
    g.log_ht += g.log_nt;
    g.log_nt = 0;
  }
  const int tiles_n = (N + (1 << g.log_nt) - 1) >> g.log_nt;
  int ntap = 0, ncls = 0, max_kb = 0, min_kb = 1 << 30;
  for (int cls = 0; cls < classes; ++cls) {
    int GH, GW, qh = 0, qw = 0;
    const int tap_begin = ntap;
    if (rowwin) {
      GH = OH; GW = OW;
      const int KB = (KW + 7) / 8;
      for (int kh = 0; kh < KH; ++kh)
        for (int kb = 0; kb < KB; ++kb) {
          g.tap_map[ntap] = 0; g.tap_dh[ntap] = (short)kh; g.tap_dw[ntap] = (short)(kb * 8);
          g.tap_widx[ntap] = (short)(kh * KB + kb);
          ++ntap;
        }
    } else if (!transposed) {
      GH = OH; GW = OW;
      GemmGeom tmp;
      int rc = fill_gather_taps(tmp, KH, KW, stride, pad);
      if (rc) return rc;
      for (int t = 0; t < tmp.num_taps; ++t) {
        g.tap_map[ntap] = tmp.tap_map[t]; g.tap_dh[ntap] = tmp.tap_dh[t];
        g.tap_dw[ntap] = tmp.tap_dw[t]; g.tap_widx[ntap] = tmp.tap_widx[t];
        ++ntap;
      }
    } else {
      qh = cls / 2; qw = cls % 2;
      if (stride == 1) { GH = OH; GW = OW; }
      else { GH = (OH - qh + 1) / 2; GW = (OW - qw + 1) / 2; }
      if (GH <= 0 || GW <= 0) continue;
      // scatter relation: out = s*in + k - p  <=>  in = (out + p - k) / s when divisible
      for (int kh = 0; kh < KH; ++kh)
        for (int kw = 0; kw < KW; ++kw) {
          int dh, dw;
          if (stride == 1) { dh = pad - kh; dw = pad - kw; }
          else {
            if (((qh + pad - kh) % 2) || ((qw + pad - kw) % 2)) continue;
            dh = floordiv(qh + pad - kh, 2);
            dw = floordiv(qw + pad - kw, 2);
          }
          g.tap_map[ntap] = 0; g.tap_dh[ntap] = (short)dh; g.tap_dw[ntap] = (short)dw;
          g.tap_widx[ntap] = (short)(kh * KW + kw);
          ++ntap;
        }
      if (ntap == tap_begin) {
        gcc_set_error(__FILE__, __LINE__, "transposed conv class with no taps is not supported");
        return GCC_ERR_ARG;
      }
    }
    g.cls_tap_begin[ncls] = tap_begin;
    g.cls_GH[ncls] = GH;
    g.cls_GW[ncls] = GW;
    g.cls_tiles_w[ncls] = (GW + (1 << g.log_wt) - 1) >> g.log_wt;
    g.cls_tiles_h[ncls] = (GH + (1 << g.log_ht) - 1) >> g.log_ht;
    g.cls_mtile_begin[ncls + 1] = g.cls_mtile_begin[ncls] + g.cls_tiles_w[ncls] * g.cls_tiles_h[ncls] * tiles_n;
    g.cls_out_off[ncls] = ((long long)qh * OW + qw) * Cy;
    g.cls_part_off[ncls] = ((long long)qh * OW + qw) * Rp;
    const int kb = c8 ? (ntap - tap_begin) / 8 : (ntap - tap_begin) * g.k_chunks;
    if (kb > max_kb) max_kb = kb;
    if (kb < min_kb) min_kb = kb;
    ++ncls;
  }
  g.cls_tap_begin[ncls] = ntap;
  g.num_classes = ncls;
  const int m_total = g.cls_mtile_begin[ncls];

  int rc = 0;
  if (rowwin) {
    rc |= make_rowwin_map(&g.a_maps[0], x, N, H, W, g.log_wt, g.log_ht, g.log_nt);
  } else if (!transposed && stride == 2) {
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw)
        rc |= make_act_map(&g.a_maps[ph * 2 + pw], x, N, H, W, Cx, 2, ph, pw, g.log_wt, g.log_ht, g.log_nt, c8);
  } else {
    rc |= make_act_map(&g.a_maps[0], x, N, H, W, Cx, 1, 0, 0, g.log_wt, g.log_ht, g.log_nt, c8);
  }
  if (c8) {  // weights [R][T][8] read as a K-major [R][T * 8] matrix
    uint64_t dims[3] = {(uint64_t)T * 8, 1, (uint64_t)R};
    uint64_t strides[2] = {(uint64_t)T * 16, (uint64_t)T * 16};
    uint32_t box[3] = {64u, 1u, (uint32_t)BN};
    rc |= gcc_make_tmap_bf16(&g.b_map, w, 3, dims, strides, box);
  } else {
    uint64_t dims[3] = {(uint64_t)Cw, (uint64_t)(w_per_image ? N : T), (uint64_t)R};
    uint64_t strides[2] = {(uint64_t)Cw * 2 * (w_per_image ? R : 1), (uint64_t)Cw * 2 * (w_per_image ? 1 : T)};
    uint32_t box[3] = {64u, 1u, (uint32_t)BN};
    rc |= gcc_make_tmap_bf16(&g.b_map, w, 3, dims, strides, box);
  }
  if (rc) return GCC_ERR_DRIVER;

  if (BN <= 128) {  // TMA-store views of the output, one per sub-pixel class
    for (int cls = 0; cls < ncls; ++cls) {
      const bf16* ob = reinterpret_cast<const bf16*>(y) + y_coff + g.cls_out_off[cls];
      uint64_t dims[4] = {(uint64_t)Rp, (uint64_t)g.cls_GW[cls], (uint64_t)g.cls_GH[cls], (uint64_t)N};
      uint64_t strides[3] = {(uint64_t)Cy * 2 * os, (uint64_t)OW * Cy * 2 * os, (uint64_t)OH * OW * Cy * 2};
      uint32_t box[4] = {64u, 1u << g.log_wt, 1u << g.log_ht, 1u << g.log_nt};
      rc |= gcc_make_tmap_bf16(&g.out_maps[cls], ob, 4, dims, strides, box);
    }
    if (rc) return GCC_ERR_DRIVER;
  }
  g.out = reinterpret_cast<bf16*>(y) + y_coff;
  g.out_sn = (long long)OH * OW * Cy;
  g.out_sh = (long long)OW * Cy * os;
  g.out_sw = (long long)Cy * os;
  g.out_cols = Rp;
  g.bias_cols = R;
  g.bias = bias;
  g.act = act;
  g.slope = slope;

  // split-K only when the tile count leaves most SMs idle and K is long
  g.k_splits = 1;
  const long long out_elems = (long long)N * OH * OW * Rp;
  const int base_tiles = m_total * g.n_tiles;
  if (f32_out && ws_elems < out_elems) {
    gcc_set_error(__FILE__, __LINE__, "conv gemm: fp32 output workspace too small");
    return GCC_ERR_ARG;
  }
  g.f32_out = f32_out;
  if (f32_out) {
    g.partial = splitk_ws;
    g.part_sn = (long long)OH * OW * Rp;
    g.part_sh = (long long)OW * Rp * os;
    g.part_sw = (long long)Rp * os;
    if (cudaMemsetAsync(splitk_ws, 0, sizeof(float) * out_elems, st) != cudaSuccess) return GCC_ERR_CUDA;
  }
  // (the fused statistics are summed from the stored bf16 tile: not available on the workspace path)
  if (splitk_ws != nullptr && (stats == nullptr || f32_out) && ws_elems >= out_elems && base_tiles * 2 <= num_sms() &&
      max_kb >= 8) {
    int ks = num_sms() / base_tiles;
    if (ks > max_kb / 2) ks = max_kb / 2;
    if (ks > 64) ks = 64;
    if (ks > 1 && f32_out) {
      g.k_splits = ks;
    } else if (ks > 1) {
      g.k_splits = ks;
      g.partial = splitk_ws;
      g.part_sn = (long long)OH * OW * Rp;
      g.part_sh = (long long)OW * Rp * os;
      g.part_sw = (long long)Rp * os;
      if (cudaMemsetAsync(splitk_ws, 0, sizeof(float) * out_elems, st) != cudaSuccess) return GCC_ERR_CUDA;
    }
  }
  g.total_tiles = base_tiles * g.k_splits;
  g.tail_begin = g.total_tiles;
  g.tail_splits = 1;
  // Tail-wave split (ConvGeom2::tail_begin): few waves, a last wave that is at most half full and a LONG K loop.
  // Measured after the issue-loop clean-up (profiles/r02_conv_pipeline_bound.txt, last table): 1024 -> 512 k4 s1 data
  // gradient (256 k-blocks) 367 -> 351 us including the finalize kernel; with 64 / 32 k-blocks per tile (256 -> 512 and
  // 128 -> 512 k4 s2 forward) the split loses 5-8 us, so it is taken from 128 k-blocks per tile on.
  // GCC_B200_TAIL_SPLIT=0 / debug bit 11 switch it off, GCC_B200_TAIL_MIN_KB moves the threshold (A/B measurements).
  static const int tail_on = getenv("GCC_B200_TAIL_SPLIT") ? atoi(getenv("GCC_B200_TAIL_SPLIT")) : 1;
  static const int tail_min_kb_env = getenv("GCC_B200_TAIL_MIN_KB") ? atoi(getenv("GCC_B200_TAIL_MIN_KB")) : 128;
  const int tail_min_kb = g_tail_min_kb > 0 ? g_tail_min_kb : tail_min_kb_env;
  if (tail_on && !(g_debug_flags & 2048) && splitk_ws != nullptr && g.k_splits == 1 && !f32_out && !c8 && BN >= 128) {
    int rem = 0;
    const int ks = plan_conv_tail(base_tiles, min_kb, num_sms(), tail_min_kb, &rem);
    if (ks >= 2 && ws_elems >= (long long)rem * ks * kBlockM * BN) {
      g.tail_begin = base_tiles - rem;
      g.tail_splits = ks;
      g.tail_ws = splitk_ws;
      g.total_tiles = g.tail_begin + rem * ks;
    }
  }
  g.debug = g_debug_flags;
  g.stats = stats;
  g.stats_ld = stats_ld;

  trace_begin(st);
  if (c8) {  // image mode: K = 16 taps x 8 channels = 2 k-blocks, always the short-K kernels
    if (BN == 64) rc = launch_conv_persistent<64, 6, true, true>(g, st);
    else if (BN == 128) rc = launch_conv_persistent<128, 4, true, true>(g, st);
    else rc = launch_conv_persistent<256, 4, false, true>(g, st);
  } else if (BN == 64) rc = launch_conv_persistent<64, 6, true, false>(g, st);
  else if (BN == 128 && max_kb <= 8) rc = launch_conv_persistent<128, 4, true, false>(g, st);
  else if (BN == 128) rc = launch_conv_persistent<128, 6, false, false>(g, st);
  else rc = launch_conv_persistent<256, 4, false, false>(g, st);
  if (rc) return rc;
  if (g.tail_splits > 1) {
    const unsigned tail_tiles = (unsigned)((g.total_tiles - g.tail_begin) / g.tail_splits);
    const dim3 fgrid(tail_tiles, kBlockM / kTailRows);
    if (BN == 256) gcc_launch(conv_tail_finalize_kernel<256>, fgrid, 256, 0, st, g);
    else gcc_launch(conv_tail_finalize_kernel<128>, fgrid, 128, 0, st, g);
    GCC_CHECK_LAUNCH();
  }
  if (g.k_splits > 1 && !f32_out) {
    long long b = (out_elems + 255) / 256;
    if (b > 148 * 8) b = 148 * 8;
    gcc_launch(splitk_finalize_kernel, (unsigned)b, 256, 0, st, splitk_ws, reinterpret_cast<bf16*>(y), (long long)N * OH * OW, Rp,
                                                       Cy, y_coff, R, bias, act, slope);
    GCC_CHECK_LAUNCH();
  }
  trace_end(st, "conv", N, H, W, Cx, R, OH, OW, KH, stride, transposed, BN, base_tiles, g.k_splits, max_kb,
            stats != nullptr, 2.0 * N * OH * OW * (double)R * Ck * KH * KW / (transposed && stride == 2 ? 4 : 1));
  return GCC_OK;
}

extern "C" int gcc_conv_gemm_bf16(const void* x, int N, int H, int W, int Cx, const void* w, int R, int T, int Cw,
                                  const float* bias, void* y, int OH, int OW, int Cy, int y_coff, int transposed,
                                  int KH, int KW, int stride, int pad, int act, float slope, int w_per_image,
                                  float* splitk_ws, long long ws_elems, float* stats, int stats_ld, void* stream) {
  return gcc_conv_gemm_launch(x, N, H, W, Cx, w, R, T, Cw, bias, y, OH, OW, Cy, y_coff, transposed, KH, KW, stride, pad,
                              act, slope, w_per_image, splitk_ws, ws_elems, stats, stats_ld, 0, 0, stream);
}

extern "C" int gcc_conv_rowwin_bf16(const void* x, int N, int Hrows, int Wp, const void* w_rowpack, int R, int KH, int KW,
                                    const float* bias, void* y, int OH, int OW, int Cy, int act, float slope, float* stats,
                                    int stats_ld, void* stream) {
  return gcc_conv_gemm_launch(x, N, Hrows, Wp, 8, w_rowpack, R, KH * ((KW + 7) / 8), 64, bias, y, OH, OW, Cy, 0, 0, KH, KW, 1,
                              0, act, slope, 0, nullptr, 0, stats, stats_ld, 0, 1, stream);
}

// Split-K factor of a weight-gradient launch from a small cost model (cycles): CTAs run one per SM for the big tiles,
// every split adds a full tile of fp32 atomics, and a partially filled last wave costs a whole wave.  Every integer
// factor is a candidate (16 base CTAs x 9 = 144 fills the 148 SMs where 16 x 8 = 128 leaves 20 idle); among the factors
// within 3 % of the cheapest the smallest wins (fewer atomic epilogues).  Pure host arithmetic (exported for the CPU
// tests as gcc_plan_wgrad_splits).
static int plan_wgrad_splits(int base_ctas, int total_pb, int BN, int MT, int c8) {
  const double t_kb = 2.0 * BN * MT * 1.4;             // MMA cycles per 64-pixel block (x1.4: L2-bound operands)
  const double t_epi = 40.0 * BN * MT + 6000.0;        // atomics epilogue + prologue
  const int slots = kNumSMs * ((BN <= 128 && MT == 1) ? 2 : 1);
  const int sp_max = c8 ? 256 : 64;
  double best = 1e30;
  auto cost_of = [&](int sp) {
    const int kb = (total_pb + sp - 1) / sp;
    const long long ctas = (long long)base_ctas * sp;
    const double waves = (double)((ctas + slots - 1) / slots);
    return waves * (kb * t_kb + t_epi);
  };
  for (int sp = 1; sp <= sp_max; ++sp) {
    if (sp > 1 && (total_pb + sp - 1) / sp < 8) break;
    best = cost_of(sp) < best ? cost_of(sp) : best;
  }
  for (int sp = 1; sp <= sp_max; ++sp) {
    if (sp > 1 && (total_pb + sp - 1) / sp < 8) break;
    if (cost_of(sp) <= best * 1.03) return sp;
  }
  return 1;
}
extern "C" int gcc_plan_wgrad_splits(int base_ctas, int total_pb, int BN, int MT, int c8) {
  return plan_wgrad_splits(base_ctas, total_pb, BN, MT, c8);
}

// dW[b][r][t][c] (+)= scale * sum_{pix} P[n, oh, ow, r] * Q[n, s*oh + kh - p, s*ow + kw - p, c]
// P: [N, OH, OW, Cp] bf16, Q: [N, H, W, Cq] bf16, dW fp32 [batch?][R][KH*KW][C].
static int gcc_wgrad_gemm_launch(const void* pmat, int N, int OH, int OW, int Cp, const void* qmat, int H, int W, int Cq,
                                 float* dw, int R, int C, int KH, int KW, int stride, int pad, int batched,
                                 int accumulate, float scale, int rowwin, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // rowwin = 1: q is the pre-padded 8-channel image of the row-window stem, C = 64 (8 taps x 8 channels per kernel-row
  // block), dw = fp32 [R][KH * KB][64]
  if (rowwin && (batched || Cq != 8 || C != 64 || stride != 1 || pad != 0 || OH + KH - 1 > H || OW + KW - 1 > W)) {
    gcc_set_error(__FILE__, __LINE__, "gcc_wgrad_gemm: bad row-window arguments");
    return GCC_ERR_ARG;
  }
  if ((Cp % 8) || (Cq % 8) || (stride != 1 && stride != 2) || KH * KW > kMaxTaps || R > Cp || (!rowwin && C > Cq)) {
    gcc_set_error(__FILE__, __LINE__, "gcc_wgrad_gemm_bf16: bad arguments");
    return GCC_ERR_ARG;
  }
  if (stride == 2 && ((H % 2) || (W % 2))) {
    gcc_set_error(__FILE__, __LINE__, "stride-2 wgrad needs even H, W");
    return GCC_ERR_ARG;
  }
  GemmGeom g;
  memset(&g, 0, sizeof(g));
  int rc = 0;
  const int KB = (KW + 7) / 8;
  if (rowwin) {
    g.num_taps = KH * KB;
    for (int kh = 0; kh < KH; ++kh)
      for (int kb = 0; kb < KB; ++kb) {
        const int t = kh * KB + kb;
        g.tap_map[t] = 0; g.tap_dh[t] = (short)kh; g.tap_dw[t] = (short)(kb * 8); g.tap_widx[t] = (short)t;
      }
  } else {
    rc = fill_gather_taps(g, KH, KW, stride, pad);
  }
  if (rc) return rc;
  g.GN = N; g.GH = OH; g.GW = OW;
  if (batched) {
    choose_tile(64, 1, OH, OW, &g.log_wt, &g.log_ht, &g.log_nt);
    // a batched tile must not straddle images
    if (g.log_nt != 0) { g.log_nt = 0; }
  } else {
    choose_tile(64, N, OH, OW, &g.log_wt, &g.log_ht, &g.log_nt);
  }
  // when the grid is smaller than 64 pixels per image in batched mode the box is still 64 pixels:
  // widen h/w logs so that the product stays 64 (extra rows are OOB -> zero)
  while (g.log_wt + g.log_ht + g.log_nt < 6) ++g.log_ht;
  g.tiles_w = (OW + (1 << g.log_wt) - 1) >> g.log_wt;
  g.tiles_h = (OH + (1 << g.log_ht) - 1) >> g.log_ht;
  g.tiles_n = (N + (1 << g.log_nt) - 1) >> g.log_nt;

  // 8-channel image on the Q side of a 16-tap conv: all taps become columns of ONE 128-wide tile
  // (dw is then the dense [R][16 * 8] matrix; the caller drops the padded channels)
  const int c8 = (!rowwin && !batched && Cq == 8 && C == 8 && KH * KW == 16 && !(g_debug_flags & 128)) ? 1 : 0;
  g.c8 = c8;
  rc = make_act_map(&g.p_map, pmat, N, OH, OW, Cp, 1, 0, 0, g.log_wt, g.log_ht, g.log_nt);
  if (rowwin) {
    rc |= make_rowwin_map(&g.a_maps[0], qmat, N, H, W, g.log_wt, g.log_ht, g.log_nt);
  } else if (stride == 2) {
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw)
        rc |= make_act_map(&g.a_maps[ph * 2 + pw], qmat, N, H, W, Cq, 2, ph, pw, g.log_wt, g.log_ht, g.log_nt, c8);
  } else {
    rc |= make_act_map(&g.a_maps[0], qmat, N, H, W, Cq, 1, 0, 0, g.log_wt, g.log_ht, g.log_nt, c8);
  }
  if (rc) return GCC_ERR_DRIVER;
  if (c8) {  // one "tap" (the whole receptive field), 128 output columns
    g.num_taps = 1;
    C = 128;
    KH = KW = 1;
  }

  const int BN = c8 ? 128 : (g_force_block_n ? g_force_block_n : (C <= 64 ? 64 : (C <= 128 ? 128 : 256)));
  // 256-row tiles (two accumulators) when there are enough rows: less operand traffic per flop
  // 256-row tiles pay when the K loop is long or there are plenty of tiles anyway; short-K layers (U-Net inner
  // levels) prefer more, smaller CTAs
  const int c_tiles_pre = (C + BN - 1) / BN;
  const int pb_pre = g.tiles_w * g.tiles_h * (batched ? 1 : g.tiles_n);
  const int base2 = ((R + 255) / 256) * c_tiles_pre * g.num_taps * (batched ? N : 1);
  const int MT = (BN >= 128 && R > 128 && !(g_debug_flags & 16) && (pb_pre >= 256 || base2 >= 100)) ? 2 : 1;
  const int r_tiles = (R + 128 * MT - 1) / (128 * MT);
  const int c_tiles = (C + BN - 1) / BN;
  const int total_pb = g.tiles_w * g.tiles_h * (batched ? 1 : g.tiles_n);
  const int base_ctas = r_tiles * c_tiles * g.num_taps * (batched ? N : 1);
  // split-K factor from a small cost model (cycles): CTAs run one per SM for the big tiles, every split adds a
  // full tile of fp32 atomics, and a partially filled last wave costs a whole wave
  const int splits = plan_wgrad_splits(base_ctas, total_pb, BN, MT, c8);
  g.splits = splits;
  g.batched = batched;
  g.atomic_out = (splits > 1 || accumulate) ? 1 : 0;
  g.scale = scale;
  g.debug = g_debug_flags;
  g.dw = dw;
  g.R = R;
  g.C = C;
  g.T_total = rowwin ? KH * KB : KH * KW;
  if (splits > 1 && !accumulate) {
    const size_t bytes = (size_t)(batched ? N : 1) * R * g.T_total * C * sizeof(float);
    if (cudaMemsetAsync(dw, 0, bytes, st) != cudaSuccess) return GCC_ERR_CUDA;
  }
  dim3 grid(r_tiles, c_tiles, g.num_taps * splits * (batched ? N : 1));
  trace_begin(st);
  if (c8 && MT == 2) rc = launch_wgrad_gemm<128, 4, 2, true>(g, grid, st);
  else if (c8) rc = launch_wgrad_gemm<128, 3, 1, true>(g, grid, st);
  else if (BN == 64) rc = launch_wgrad_gemm<64, 4, 1>(g, grid, st);
  else if (BN == 128 && MT == 2) rc = launch_wgrad_gemm<128, 4, 2>(g, grid, st);
  else if (BN == 128) rc = launch_wgrad_gemm<128, 3, 1>(g, grid, st);
  else if (MT == 2) rc = launch_wgrad_gemm<256, 3, 2>(g, grid, st);
  else rc = launch_wgrad_gemm<256, 4, 1>(g, grid, st);
  trace_end(st, "wgrad", N, H, W, C, R, OH, OW, KH, stride, batched, BN * 10 + MT, base_ctas, splits, total_pb, 0,
            2.0 * N * OH * OW * (double)R * C * KH * KW);
  return rc;
}

extern "C" int gcc_wgrad_gemm_bf16(const void* pmat, int N, int OH, int OW, int Cp, const void* qmat, int H, int W,
                                   int Cq, float* dw, int R, int C, int KH, int KW, int stride, int pad,
                                   int batched, int accumulate, float scale, void* stream) {
  return gcc_wgrad_gemm_launch(pmat, N, OH, OW, Cp, qmat, H, W, Cq, dw, R, C, KH, KW, stride, pad, batched, accumulate, scale,
                               0, stream);
}
extern "C" int gcc_wgrad_rowwin_bf16(const void* dy, int N, int OH, int OW, int Cp, const void* x, int Hrows, int Wp, float* dw,
                                     int R, int KH, int KW, void* stream) {
  return gcc_wgrad_gemm_launch(dy, N, OH, OW, Cp, x, Hrows, Wp, 8, dw, R, 64, KH, KW, 1, 0, 0, 0, 1.f, 1, stream);
}
