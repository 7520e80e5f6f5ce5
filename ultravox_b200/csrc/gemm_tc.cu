// uvx_gemm_bf16: bf16 x bf16 -> fp32 GEMM on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
//   C[row(b,m), n] = act(alpha * sum_k A[b,m,k] W[n,k] + bias[n]) + R[b,m,n]
//
// Persistent, warp-specialised kernel: grid = min(#tiles, #SMs), each CTA walks tiles t = blockIdx.x + i*gridDim.x.
// A CTA tile is (MT*128) x BN: MT in {1,2} row sub-tiles share one W tile (so at M <= 256 - the LLM prefill - every
// weight byte is fetched from L2/HBM exactly once), BN in {64,128,208,256} (208: ragged last column tile allowed).
//   warp 0      TMA producer: cp.async.bulk.tensor of the A box {64 k, MT*128 rows, 1 batch} and the W box {64 k, BN
//               rows} into a kStages-deep 128B-swizzled shared-memory ring (mbarrier expect_tx); runs ahead across tiles.
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16, fp32 accumulators in TMEM, MT
//               accumulators per tile); tcgen05.commit releases ring slots and publishes finished accumulators.
//   warps 2..   epilogue (4 warps at MT=2, 8 at MT=1): tcgen05.ld 32 lanes x 32 columns per warp, fused bias / GELU /
//               residual / row remap, transposed through shared memory into full-sector 16-byte stores.  When MT*BN <= 256 the accumulators are double-buffered in TMEM, so the epilogue of tile i
//               overlaps the main loop of tile i+1.
// A is described by a 3-D tensor map (k, row, batch) with caller-chosen strides, so overlapping rows (implicit-GEMM
// conv over a time-major activation) and batch-strided inputs need no im2col copy; rows or k beyond the tensor bounds
// are zero-filled by TMA, which is how M / K tails are handled.
#include <stdlib.h>

#include "uvx_common.cuh"
#include "tc_ptx.cuh"

namespace uvx {

bool gemm_ws_eligible(const uvx_gemm_args* a);                   // gemm_ws.cu
int launch_gemm_ws(const uvx_gemm_args* a, cudaStream_t stream);

static int g_gemm_dbg = 0;  // uvx_debug_gemm_mode

struct GemmParams {
  int64_t a_rows, a_batch, K, N;
  void* C;
  int64_t c_row_stride, c_batch_rows, c_row_offset;
  const int32_t* c_row_map;
  const bf16* bias;
  const bf16* R;
  int64_t r_row_stride, r_batch_stride;
  float alpha;
  int act, out_f32;
  int m_tiles;  // per batch (of MT*128 rows)
  int n_tiles, num_tiles;
  int splits, kb_per_split;  // split-K: units = num_tiles * splits
  float* ws_partial;         // [splits][a_batch * a_rows][N] fp32 partial sums (reduced by splitk_reduce_kernel)
  const bf16* norm_w;        // optional fused RMSNorm of the finished output rows: norm_out = w * bf16(C * rstd)
  bf16* norm_out;
  float norm_eps;
  int a_box_bytes;  // bytes of the A box in one ring stage (box rows * 128): fewer rows than MT*128 when M is small
  int stage_bytes;  // a_box_bytes + BN * 128
  int stages;       // ring depth that fits the shared-memory budget (2..8)
  // thread-block cluster of cm x cn CTAs computing cm row tiles x cn column tiles: the A box of a row tile is loaded once
  // (1/cn slice per CTA, TMA multicast) for its cn CTAs and the W box once for its cm CTAs - the GEMMs are bound by bytes
  // INTO the SMs (L2 -> SM fabric, ~9 TB/s aggregate), so sharing operand tiles is what raises the ceiling
  int cm, cn;
  int a_slice_rows, w_slice_rows;
  int m_groups, n_groups;  // ceil(tiles / cm), ceil(n_tiles / cn)
  int dbg_mode;            // tuning only: 1 = skip the MMAs (load pipeline alone), 2 = skip the TMA loads (MMA pipeline alone)
  // ---- round 2: weight-stream pipeline + fused epilogues (production kernel only)
  int w_tiled;             // W is the pre-tiled image [n_tiles][K/64][BN][64]: every TMA box is ONE contiguous BN*128-byte run
  int pf;                  // L2 prefetch distance in k-blocks: UTMAPF of the W box `pf` k-blocks ahead of the smem ring, so the
                           // HBM stream runs ahead of (and is decoupled from) the ring slots - 0 = off
  int swiglu;              // epilogue pairs tile columns (8 gate | 8 up interleaved): out[:, 16 per 32-column chunk]
  const float* rope_cos;   // fused RoPE (hf:modeling_llama.py:124-168) on tile columns < rope_cols, head_dim 128 == BN
  const float* rope_sin;
  const int32_t* rope_pos;
  int64_t rope_rows_per_seq, rope_pos_offset;
  int rope_cols;
  int tma_store;           // direct epilogue writes bf16 through TMA stores (32 x 32 boxes staged in the per-warp pads)
  int epi_ring;            // every CTA has at most one unit: once its accumulators are complete the operand ring is idle, and each
                           // epilogue warp stages its boxes in 16 KB of it (8 stores in flight instead of 2)
  int pdl;                 // launched with programmatic stream serialization: request the first W boxes before griddepcontrol.wait
  long long* dbg_times;    // DIAG twin only: per CTA {t0 kernel entry, setup done, first stage landed, last MMA issued, accumulators
                           // complete, epilogue done, globaltimer at entry, smid} (clock64 ticks)
};


// Fast epilogue of one 32-row x 32-column accumulator chunk (bf16 output, plain row order): alpha, bias, activation, residual in
// registers (+ the residual row of the lane), 64 bytes per row into a 2 KB half of the warp's pad, ONE TMA store of the 32 x 32 box (rows past M and columns past N
// are clipped by the tensor map).  The transposing epilogue_chunk costs ~1400 cycles per chunk on the exposed tail of a one-tile
// CTA (profiles/r2_ws_times_v2.txt: 24 % of the gate|up kernel); this is a tcgen05.ld, 16 packs, 4 stores and one UTMASTG.
__device__ __forceinline__ void epilogue_chunk_tma(const GemmParams& p, const CUtensorMap* tmC, const uint32_t* raw, uint8_t* half, int lane,
                                                   int64_t m_warp0, int64_t n) {
  // residual row of this lane (64 contiguous bytes), issued first so its latency hides behind the math
  uint4 rres[4];
  const bool has_r = p.R != nullptr && m_warp0 + lane < p.a_rows;
  if (has_r) {
    const uint4* rr = reinterpret_cast<const uint4*>(p.R + (m_warp0 + lane) * p.r_row_stride + n);
#pragma unroll
    for (int g = 0; g < 4; ++g) rres[g] = rr[g];
  }
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]) * p.alpha;
  if (p.bias) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float t[8];
      unpack8(*reinterpret_cast<const bf16x8*>(p.bias + n + g * 8), t);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[g * 8 + i] += t[i];
    }
  }
  if (p.act == UVX_ACT_GELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_fast(v[i]);
  }
  if (has_r) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float t[8];
      unpack8(*reinterpret_cast<const bf16x8*>(&rres[g]), t);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[g * 8 + i] += t[i];
    }
  }
  uint4* dst = reinterpret_cast<uint4*>(half + lane * 64);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const bf16x8 pk = pack8(v + 8 * g);
    dst[g] = *reinterpret_cast<const uint4*>(&pk);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) tma_store_2d(half, tmC, (int)n, (int)m_warp0);
}

// Fused SwiGLU, fast form: 16 finished activations per row and chunk -> 32 bytes per row into a 1 KB slice of the pad -> one
// TMA store of the 32-row x 16-column box of the [M, N/2] output.
__device__ __forceinline__ void epilogue_chunk_swiglu_tma(const GemmParams& p, const CUtensorMap* tmC, const uint32_t* raw, uint8_t* slice,
                                                          int lane, int64_t m_warp0, int64_t n_out) {
  float o[16];
#pragma unroll
  for (int g2 = 0; g2 < 2; ++g2) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float gate = __bfloat162float(__float2bfloat16_rn(__uint_as_float(raw[g2 * 16 + i]) * p.alpha));
      const float up = __bfloat162float(__float2bfloat16_rn(__uint_as_float(raw[g2 * 16 + 8 + i]) * p.alpha));
      o[g2 * 8 + i] = __bfloat162float(__float2bfloat16_rn(silu_fast(gate))) * up;
    }
  }
  uint4* dst = reinterpret_cast<uint4*>(slice + lane * 32);
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const bf16x8 pk = pack8(o + 8 * g);
    dst[g] = *reinterpret_cast<const uint4*>(&pk);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) tma_store_2d(slice, tmC, (int)n_out, (int)m_warp0);
}


// ---------------------------------------------------------------------------------- coalesced epilogue
// A warp owns 32 accumulator rows (TMEM lanes); tcgen05.ld hands each LANE one row x 32 columns.  Storing that
// directly makes every 16-byte store hit a different row (measured 8x write amplification on the SM->L2 path), so
// the 32x32 fp32 chunk is transposed through a per-warp shared-memory pad and written back with 4 lanes per row:
// every store instruction covers 8 rows x 64 contiguous bytes (full 32-byte sectors), residual reads likewise.
static constexpr int kStageBytesPerWarp = 32 * 32 * 4;  // 4096 B: 32 rows x 32 fp32, 16-byte chunks XOR-swizzled by row

__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, const uint32_t* raw, float* stage, int lane, int b,
                                               int64_t m_warp0, int64_t n, int64_t n_lim) {
  // phase 0: this lane's share of the coalesced write-back (rows r_j, 8 (bf16) / 4 (f32) columns) - issue the residual
  // loads first so that their latency hides behind the activation math and the transpose
  constexpr int kJ = 8;
  const bool f32 = p.out_f32 != 0;
  const int nj = f32 ? 8 : 4;
  int64_t orow[kJ];
  uint4 rres[kJ];
#pragma unroll
  for (int j = 0; j < kJ; ++j) {
    orow[j] = -1;
    rres[j] = make_uint4(0u, 0u, 0u, 0u);
    if (j >= nj) continue;
    const int r = f32 ? (lane >> 3) + 4 * j : (lane >> 2) + 8 * j;
    const int pc = f32 ? (lane & 7) * 4 : (lane & 3) * 8;
    const int64_t m = m_warp0 + r;
    if (m >= p.a_rows || n + pc >= n_lim) continue;
    orow[j] = p.c_row_map ? (int64_t)p.c_row_map[(int64_t)b * p.a_rows + m] : (int64_t)b * p.c_batch_rows + m + p.c_row_offset;
    if (orow[j] < 0 || !p.R) continue;
    const bf16* rr = p.R + (int64_t)b * p.r_batch_stride + m * p.r_row_stride + n + pc;
    if (f32) {
      const uint2 t = *reinterpret_cast<const uint2*>(rr);
      rres[j].x = t.x; rres[j].y = t.y;
    } else {
      rres[j] = *reinterpret_cast<const uint4*>(rr);
    }
  }
  // phase 1: this lane's row, 32 columns: alpha, bias, activation (fp32)
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]) * p.alpha;
  if (p.bias) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float t[8];
      if (n + g * 8 >= n_lim) break;  // columns past the tile / matrix edge are never stored
      unpack8(*reinterpret_cast<const bf16x8*>(p.bias + n + g * 8), t);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[g * 8 + i] += t[i];
    }
  }
  if (p.act == UVX_ACT_GELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_fast(v[i]);
  }
  float4* srow = reinterpret_cast<float4*>(stage + lane * 32);
#pragma unroll
  for (int g = 0; g < 8; ++g) srow[g ^ (lane & 7)] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
  __syncwarp();
  // phase 2: transposed read-back, residual add, coalesced stores (full 32-byte sectors)
  if (f32) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (orow[j] < 0) continue;
      const int r = (lane >> 3) + 4 * j, pc = (lane & 7) * 4;
      float4 t = reinterpret_cast<const float4*>(stage + r * 32)[(lane & 7) ^ (r & 7)];
      if (p.R) {
        const float2 r01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rres[j].x));
        const float2 r23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rres[j].y));
        t.x += r01.x; t.y += r01.y; t.z += r23.x; t.w += r23.y;
      }
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + orow[j] * p.c_row_stride + n + pc) = t;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (orow[j] < 0) continue;
      const int r = (lane >> 2) + 8 * j, pc = (lane & 3) * 8;
      const float4* row4 = reinterpret_cast<const float4*>(stage + r * 32);
      const float4 t0 = row4[(2 * (lane & 3)) ^ (r & 7)];
      const float4 t1 = row4[(2 * (lane & 3) + 1) ^ (r & 7)];
      float o[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
      if (p.R) {
        float rv[8];
        unpack8(*reinterpret_cast<const bf16x8*>(&rres[j]), rv);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += rv[i];
      }
      *reinterpret_cast<bf16x8*>(reinterpret_cast<bf16*>(p.C) + orow[j] * p.c_row_stride + n + pc) = pack8(o);
    }
  }
  __syncwarp();  // the pad is reused by the next chunk
}

// Fused SwiGLU epilogue (Llama MLP act(gate) * up, hf:modeling_llama.py:183): the pre-tiled gate|up image interleaves 8 gate
// rows with the 8 up rows of the same features, so one 32-column accumulator chunk holds 2 x (8 gate | 8 up) and yields 16
// finished activations per row - the [M, 2*ffn] intermediate is never written or re-read.  Rounding order is the unfused
// path's: gate and up rounded to bf16 (what the GEMM would have stored), silu rounded to bf16, product rounded to bf16.
__device__ __forceinline__ void epilogue_chunk_swiglu(const GemmParams& p, const uint32_t* raw, float* stage, int lane, int b,
                                                      int64_t m_warp0, int64_t n_out, int64_t n_out_lim) {
  float o[16];
#pragma unroll
  for (int g2 = 0; g2 < 2; ++g2) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float gate = __bfloat162float(__float2bfloat16_rn(__uint_as_float(raw[g2 * 16 + i]) * p.alpha));
      const float up = __bfloat162float(__float2bfloat16_rn(__uint_as_float(raw[g2 * 16 + 8 + i]) * p.alpha));
      o[g2 * 8 + i] = __bfloat162float(__float2bfloat16_rn(silu(gate))) * up;
    }
  }
  // 32 rows x 16 floats, 16-byte chunks XOR-swizzled by (row >> 1) & 3 (conflict-free both ways)
  float4* srow = reinterpret_cast<float4*>(stage + lane * 16);
#pragma unroll
  for (int g = 0; g < 4; ++g) srow[g ^ ((lane >> 1) & 3)] = make_float4(o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int r = (lane >> 1) + 16 * j, half = lane & 1;
    const int64_t m = m_warp0 + r;
    if (m >= p.a_rows || n_out + half * 8 >= n_out_lim) continue;
    const int64_t orow = p.c_row_map ? (int64_t)p.c_row_map[(int64_t)b * p.a_rows + m] : (int64_t)b * p.c_batch_rows + m + p.c_row_offset;
    if (orow < 0) continue;
    const float4* row4 = reinterpret_cast<const float4*>(stage + r * 16);
    const float4 t0 = row4[(2 * half) ^ ((r >> 1) & 3)];
    const float4 t1 = row4[(2 * half + 1) ^ ((r >> 1) & 3)];
    const float v8[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
    *reinterpret_cast<bf16x8*>(reinterpret_cast<bf16*>(p.C) + orow * p.c_row_stride + n_out + half * 8) = pack8(v8);
  }
  __syncwarp();
}

// Shared-memory plan of the 1-SM kernel (dynamic, 1024-byte aligned base, always the full 227 KB - one CTA per SM):
//   [0, stages * stage_bytes)   ring of {A box, W box} stages; the geometry is a launch parameter because the A box
//                               only holds round8(M) rows when one row tile covers the whole problem (M = 201 prefill:
//                               208 rows instead of 256 -> 4 ring stages instead of 3, i.e. more weight bytes in flight)
//   [kPadOff, kBarOff)          epilogue transpose pads, 4 KB per epilogue warp
//   [kBarOff, kSmemTotal)       mbarriers + the TMEM base slot
static constexpr int kSmemTotal = 227 * 1024;
static constexpr int kBarOff = kSmemTotal - 256;
static constexpr int kMaxStages = 8;

template <int MT, int BN, int EW = 0>
struct SmemLayout {
  static constexpr int kEpiWarps = EW ? EW : (MT == 1 ? 8 : 4);   // tensor-bound shapes (MT = 1) get two epilogue warps per TMEM lane quarter
  static constexpr int kThreads = 64 + 32 * kEpiWarps;
  static constexpr int kWBytes = BN * kBK * 2;
  static constexpr int kBNT = BN <= 64 ? 64 : (BN <= 128 ? 128 : 256);   // TMEM column stride of one 128-row accumulator
  static constexpr int kAcc = (MT * kBNT * 2 <= 512) ? 2 : 1;            // TMEM accumulator stages
  static constexpr int kTmemCols = kAcc * MT * kBNT < 32 ? 32 : kAcc * MT * kBNT;
  static constexpr int kChunks = (BN + 31) / 32;                        // 32-column epilogue chunks (last may be partial)
  static constexpr int kPadOff = kBarOff - kEpiWarps * kStageBytesPerWarp;
  static_assert((kTmemCols & (kTmemCols - 1)) == 0 && kTmemCols <= 512, "TMEM columns must be a power of two <= 512");
  static_assert(2 * (MT * kBM * kBK * 2 + kWBytes) <= kPadOff, "ring too shallow");
};

// EW: epilogue warps (0 = default for MT).  DIAG: tuning twin with the pipeline-isolation modes compiled in (uvx_debug_gemm_mode:
// 1 = loads only, 2 = MMAs only, 3 = no epilogue stores) - same convergent producer / MMA loops as production, unlike
// gemm_tc_kernel_x; never on the product path.
// FEAT: bit 0 = weight-stream producer (pre-tiled images, PDL early W loads, L2 prefetch hook), bit 1 = round-2 epilogues (TMA stores,
// fused RoPE / SwiGLU).  FEAT = 0 compiles the round-1 loops unchanged: the single-thread producer / MMA loops are latency-critical and
// every extra select or branch in them showed up as 2-3 us per launch in situ (profiles/r2_insitu_timeline_v2.txt).
template <int MT, int BN, int EW = 0, bool DIAG = false, int FEAT = 3>
__global__ void __launch_bounds__(SmemLayout<MT, BN, EW>::kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmC,
               const GemmParams p) {
  pdl_trigger();
  long long t_entry = 0;
  if constexpr (DIAG) t_entry = clock64();
  using L = SmemLayout<MT, BN, EW>;
  constexpr bool kWS = (FEAT & 1) != 0, kEpi = (FEAT & 2) != 0;
  constexpr int kAcc = L::kAcc;
  constexpr int kEpiWarps = L::kEpiWarps;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full_bar = (uint64_t*)(smem + kBarOff);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full = empty_bar + kMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m_total = p.m_tiles * (int)p.a_batch;
  const int num_kb = (int)((p.K + kBK - 1) / kBK);
  const int num_units = p.num_tiles * p.splits;
  const int stages = p.stages;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();  // the 128B-swizzle atoms need a 1024-byte aligned ring
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < kAcc; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], kEpiWarps);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)L::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // set-up above overlapped the previous kernel's tail; its outputs are visible after griddepcontrol.wait.  The producer warp
  // delays its own wait until the first weight prefetches are issued (weights do not depend on the previous kernel).
  if (warp != 0) pdl_wait();
  if constexpr (DIAG) {
    if (p.dbg_times && threadIdx.x == 64) {
      long long* t = p.dbg_times + (size_t)blockIdx.x * 8;
      unsigned long long gt;
      unsigned smid;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      t[0] = t_entry;
      t[1] = clock64();
      t[6] = (long long)gt;
      t[7] = (long long)smid;
    }
  }

  if (warp == 0) {
    // ---- TMA producer: all 32 lanes run the loop (convergent), elect.sync issues
    {
      if (lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
      }
      __syncwarp();
      int s = 0;           // ring slot and its phase: the ring runs ahead across tile boundaries
      uint32_t ph = 0;
      bool waited = false;
      const int tiled = kWS ? p.w_tiled : 0;
      // Programmatic dependent launch: this CTA may be resident while the previous kernel is still running.  Weights do not
      // depend on it, so the W boxes of the first ring round are requested BEFORE griddepcontrol.wait (their latency - and the
      // launch / set-up above - hide under the previous kernel's tail); the activation boxes follow after the wait.
      int pre = 0;
      if (kWS && p.pdl && (int)blockIdx.x < num_units) {
        const int unit = blockIdx.x;
        const int tile = unit / p.splits, split = unit % p.splits;
        const int tn_idx = tile / tiles_m_total;
        const int kb_begin = split * p.kb_per_split;
        const int nk = min(num_kb, kb_begin + p.kb_per_split) - kb_begin;
        const int rot = (int)(((unsigned)tile * 7u + (unsigned)split * 3u) % (unsigned)nk);
        pre = stages < nk ? stages : nk;
        for (int i = 0; i < pre; ++i) {
          const int kb = kb_begin + (i + rot < nk ? i + rot : i + rot - nk);
          mbar_expect_tx_e(&full_bar[i], (uint32_t)p.stage_bytes);
          tma_load_2d_e(smem + i * p.stage_bytes + p.a_box_bytes, &tmW, tiled ? 0 : kb * kBK, tiled ? (tn_idx * num_kb + kb) * BN : tn_idx * BN,
                        &full_bar[i]);
        }
      }
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
        const int tile = unit / p.splits, split = unit % p.splits;
        const int tm_idx = tile % tiles_m_total;  // consecutive tiles share the W tile
        const int tn_idx = tile / tiles_m_total;
        const int b = tm_idx / p.m_tiles;
        const int m0 = (tm_idx % p.m_tiles) * (MT * kBM);
        const int n0 = tn_idx * BN;
        const int kb_begin = split * p.kb_per_split;
        const int kb_end = min(num_kb, kb_begin + p.kb_per_split);
        // every CTA walks K starting at a different k-block (and wraps): at M <= 256 all CTAs read the SAME activation
        // tile, and in lockstep they would all hit the same L2 lines at the same time
        const int nk = kb_end - kb_begin;
        const int rot = (int)(((unsigned)tile * 7u + (unsigned)split * 3u) % (unsigned)nk);
        // W box of k-block kb: canonical [N, K] weights -> (kb * 64, n0); pre-tiled image [tile][kb][BN][64] viewed as
        // [rows, 64] -> (0, (tn_idx * num_kb + kb) * BN): one contiguous BN * 128-byte run of DRAM
        const int wt_base = tn_idx * num_kb * BN;
        // the HBM -> L2 stream of the weights runs `pf` k-blocks ahead of the shared-memory ring: the ring (4-5 stages, half of
        // each stage is the re-read activation tile) holds too few weight bytes in flight to cover DRAM latency at full rate
        const int pf = kWS ? (p.pf < nk ? p.pf : nk) : 0;
        for (int i = 0; i < pf; ++i) {
          const int kb = kb_begin + (i + rot < nk ? i + rot : i + rot - nk);
          tma_prefetch_2d_e(&tmW, tiled ? 0 : kb * kBK, tiled ? wt_base + kb * BN : n0);
        }
        if (!waited) {
          pdl_wait();
          waited = true;
        }
        for (int i = 0; i < nk; ++i) {
          const int kb = kb_begin + (i + rot < nk ? i + rot : i + rot - nk);
          if (kWS && pf > 0 && i + pf < nk) {
            const int j = i + pf;
            const int kbp = kb_begin + (j + rot < nk ? j + rot : j + rot - nk);
            tma_prefetch_2d_e(&tmW, tiled ? 0 : kbp * kBK, tiled ? wt_base + kbp * BN : n0);
          }
          uint8_t* sa = smem + s * p.stage_bytes;
          if (kWS && pre > 0) {   // first ring round of the first unit: the W box is already in flight, only the activation box is missing
            --pre;
            if constexpr (DIAG) {
              if (p.dbg_mode == 2) {   // (isolation mode: the early W loads still land; nothing else to do)
                tma_load_3d_e(sa, &tmA, kb * kBK, m0, b, &full_bar[s]);
                if (++s == stages) { s = 0; ph ^= 1u; }
                continue;
              }
            }
            tma_load_3d_e(sa, &tmA, kb * kBK, m0, b, &full_bar[s]);
            if (++s == stages) { s = 0; ph ^= 1u; }
            continue;
          }
          mbar_wait(&empty_bar[s], ph ^ 1u);
          if constexpr (DIAG) {
            if (p.dbg_mode == 2) {  // MMAs only: hand the (never loaded) slot over
              if (lane == 0) mbar_arrive(&full_bar[s]);
              __syncwarp();
              if (++s == stages) { s = 0; ph ^= 1u; }
              continue;
            }
          }
          mbar_expect_tx_e(&full_bar[s], (uint32_t)p.stage_bytes);
          tma_load_3d_e(sa, &tmA, kb * kBK, m0, b, &full_bar[s]);
          tma_load_2d_e(sa + p.a_box_bytes, &tmW, tiled ? 0 : kb * kBK, tiled ? wt_base + kb * BN : n0, &full_bar[s]);
          if (++s == stages) { s = 0; ph ^= 1u; }
        }
      }
      if (!waited) pdl_wait();
    }
  } else if (warp == 1) {
    // ---- MMA issuer: convergent warp, elect.sync issues
    {
      constexpr uint32_t idesc = make_idesc(BN);
      int s = 0;
      uint32_t ph = 0, tcount = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++tcount) {
        const int split = unit % p.splits;
        const int kb_begin = split * p.kb_per_split;
        const int kb_end = min(num_kb, kb_begin + p.kb_per_split);
        const uint32_t acc = tcount % kAcc;
        const uint32_t aph = (tcount / kAcc) & 1u;
        mbar_wait(&tmem_empty[acc], aph ^ 1u);  // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (MT * L::kBNT);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if constexpr (DIAG) {
            if (p.dbg_times && lane == 0 && kb == kb_begin && tcount == 0) p.dbg_times[(size_t)blockIdx.x * 8 + 2] = clock64();
            if (p.dbg_mode == 1) {  // loads only: free the slot at once
              if (lane == 0) mbar_arrive(&empty_bar[s]);
              __syncwarp();
              if (++s == stages) { s = 0; ph ^= 1u; }
              continue;
            }
          }
          const uint32_t sa = smem_u32(smem + s * p.stage_bytes);
          const uint64_t dw = make_smem_desc(sa + p.a_box_bytes);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            // rows past the A box (M small) read whatever follows in shared memory: they only feed accumulator rows
            // >= a_rows, which the epilogue never stores
            const uint64_t da = make_smem_desc(sa + mt * (kBM * kBK * 2));
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in the (addr >> 4) field
              umma_f16_e(d_tmem + mt * L::kBNT, da + (uint64_t)(2 * k), dw + (uint64_t)(2 * k), idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
            }
          }
          umma_commit_e(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
          if (++s == stages) { s = 0; ph ^= 1u; }
        }
        if constexpr (DIAG) {
          if (p.dbg_times && lane == 0) p.dbg_times[(size_t)blockIdx.x * 8 + 3] = clock64();
          if (p.dbg_mode == 1) {
            if (lane == 0) mbar_arrive(&tmem_full[acc]);
            __syncwarp();
            continue;
          }
        }
        umma_commit_e(&tmem_full[acc]);  // accumulators of this tile complete
      }
    }
  } else {
    // ---- epilogue: warp w reads TMEM lane quarter w % 4; with 8 warps the two warps of a quarter take alternate chunks
    const int q = warp & 3;
    const int cpar = (warp - 2) >> 2;              // 0 for the first four epilogue warps, 1 for the second four
    constexpr int cstep = kEpiWarps / 4;
    uint32_t tcount = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++tcount) {
      const int tile = unit / p.splits, split = unit % p.splits;
      const int tm_idx = tile % tiles_m_total;
      const int tn_idx = tile / tiles_m_total;
      const int b = tm_idx / p.m_tiles;
      const int m0 = (tm_idx % p.m_tiles) * (MT * kBM);
      const int n0 = tn_idx * BN;
      const uint32_t acc = tcount % kAcc;
      const uint32_t aph = (tcount / kAcc) & 1u;
      mbar_wait(&tmem_full[acc], aph);
      tc_fence_after();
      if constexpr (DIAG) {
        if (p.dbg_times && threadIdx.x == 64) p.dbg_times[(size_t)blockIdx.x * 8 + 4] = clock64();
        if (p.dbg_mode == 1 || p.dbg_mode == 3) {  // no epilogue work: hand the accumulator stage straight back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
          continue;
        }
      }
      // staging buffers of the TMA-store epilogue: 2 KB each; two in the warp's pad, or eight in the idle operand ring
      uint8_t* const epi_pad = smem + L::kPadOff + (warp - 2) * kStageBytesPerWarp;
      uint8_t* const epi_ring = smem + (warp - 2) * 16384;
      const bool ring = kEpi && p.epi_ring != 0;
      const bool tma = kEpi && p.tma_store != 0;
      auto epi_acquire = [&](uint32_t n) -> uint8_t* {
        if (ring) {
          if (n >= 8) {   // the buffer written now was last read by the store issued eight boxes ago
            if (lane == 0) tma_store_wait_read<7>();
            __syncwarp();
          }
          return epi_ring + (n & 7u) * 2048;
        }
        if (n >= 2) {
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
        }
        return epi_pad + (n & 1u) * 2048;
      };
      const bool direct = p.splits == 1;
      if (!direct) {
        // split-K: park the raw fp32 partial tile in the workspace; splitk_reduce_kernel sums the splits in a fixed
        // order and applies the epilogue (deterministic, and the reduction is spread over every SM)
        float* part = p.ws_partial + (size_t)split * (size_t)(p.a_batch * p.a_rows) * (size_t)p.N;
        uint32_t pstore = 0;
#pragma unroll 1
        for (int mt = 0; mt < MT; ++mt) {
          if ((int64_t)m0 + mt * kBM + q * 32 >= p.a_rows) continue;
          const int64_t m = (int64_t)m0 + mt * kBM + q * 32 + lane;
          const bool row_ok = m < p.a_rows;
#pragma unroll 1
          for (int c = cpar; c < L::kChunks; c += cstep) {
            uint32_t raw[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * (MT * L::kBNT) + (uint32_t)(mt * L::kBNT + c * 32), raw);
            tmem_ld_wait();
            if (tma) {
              // fp32 partial tile through the pad as two 32-row x 16-column boxes (64-byte rows), TMA-stored into
              // [split][row][col]; rows past M are clipped by the map
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                uint8_t* buf = epi_acquire(pstore);
                uint4* dsts = reinterpret_cast<uint4*>(buf + lane * 64);
#pragma unroll
                for (int g = 0; g < 4; ++g)
                  dsts[g] = make_uint4(raw[hh * 16 + 4 * g], raw[hh * 16 + 4 * g + 1], raw[hh * 16 + 4 * g + 2], raw[hh * 16 + 4 * g + 3]);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) tma_store_3d(buf, &tmC, n0 + c * 32 + hh * 16, m0 + mt * kBM + q * 32, split);
                ++pstore;
              }
            } else if (row_ok) {
              uint4* dst = reinterpret_cast<uint4*>(part + ((size_t)b * p.a_rows + m) * p.N + n0 + c * 32);
#pragma unroll
              for (int g = 0; g < 8; ++g) dst[g] = make_uint4(raw[4 * g], raw[4 * g + 1], raw[4 * g + 2], raw[4 * g + 3]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);  // accumulator stage is free again
        if (pstore > 0) {
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
          pstore = 0;
        }
        continue;
      }
      float* pad = reinterpret_cast<float*>(smem + L::kPadOff) + (warp - 2) * (kStageBytesPerWarp / 4);
      const int64_t n_lim = (int64_t)n0 + BN < p.N ? (int64_t)n0 + BN : p.N;  // ragged last column tile (N % BN != 0)
      bool handled = false;
      uint32_t nstore = 0;   // TMA stores issued by this warp for this tile (pad halves alternate)
      if constexpr (BN == 208 && kEpi) {
        if (p.swiglu) {  // warp-uniform: tile columns are (8 gate | 8 up) groups, BN / 2 finished activations per row
          handled = true;
          const int64_t n_out0 = (int64_t)tn_idx * (BN / 2);
          const int64_t n_out_lim = n_out0 + BN / 2 < p.N / 2 ? n_out0 + BN / 2 : p.N / 2;
#pragma unroll 1
          for (int mt = 0; mt < MT; ++mt) {
            const int64_t m_warp0 = (int64_t)m0 + mt * kBM + q * 32;
            if (m_warp0 >= p.a_rows) continue;
#pragma unroll 1
            for (int c = cpar; c < L::kChunks; c += cstep) {
              if (n_out0 + c * 16 >= n_out_lim) break;
              uint32_t raw[32];
              tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * (MT * L::kBNT) + (uint32_t)(mt * L::kBNT + c * 32), raw);
              tmem_ld_wait();
              if (tma && c * 32 + 32 <= BN && n_out0 + c * 16 + 16 <= n_out_lim) {
                uint8_t* buf = epi_acquire(nstore);
                epilogue_chunk_swiglu_tma(p, &tmC, raw, buf, lane, m_warp0, n_out0 + c * 16);
                ++nstore;
              } else {
                if (!ring && nstore > 0) {   // the transposing path uses the whole pad: drain the stores that still read it
                  if (lane == 0) tma_store_wait_read<0>();
                  __syncwarp();
                  nstore = 0;
                }
                epilogue_chunk_swiglu(p, raw, pad, lane, b, m_warp0, n_out0 + c * 16, n_out_lim);
              }
            }
          }
        }
      }
      if constexpr (MT == 1 && BN == 128 && kEpi) {
        if (p.rope_cos != nullptr && n0 < p.rope_cols) {
          // fused RoPE: the tile is one 128-wide head, columns j and j + 64 rotate together; this warp owns chunks cpar and
          // cpar + 2, i.e. columns [32 cpar, +32) and their partners.  Same rounding as GEMM-then-uvx_rope: the projection is
          // rounded to bf16 first, the rotation runs in fp32 on those values and is rounded once more by the store.
          handled = true;
          const int64_t m_warp0 = (int64_t)m0 + q * 32;
          if (m_warp0 < p.a_rows) {
            uint32_t r0[32], r1[32];
            const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (MT * L::kBNT);
            tmem_ld32(tbase + (uint32_t)(cpar * 32), r0);
            tmem_ld32(tbase + (uint32_t)(64 + cpar * 32), r1);
            tmem_ld_wait();
            const int64_t m = m_warp0 + lane;
            if (m < p.a_rows) {
              const int64_t pos = p.rope_pos ? (int64_t)p.rope_pos[m] : p.rope_pos_offset + (m % p.rope_rows_per_seq);
              const float4* cp = reinterpret_cast<const float4*>(p.rope_cos + pos * 64 + cpar * 32);
              const float4* sp = reinterpret_cast<const float4*>(p.rope_sin + pos * 64 + cpar * 32);
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float4 c4 = cp[g], s4 = sp[g];
                const float cc[4] = {c4.x, c4.y, c4.z, c4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float x1 = __bfloat162float(__float2bfloat16_rn(__uint_as_float(r0[4 * g + e]) * p.alpha));
                  const float x2 = __bfloat162float(__float2bfloat16_rn(__uint_as_float(r1[4 * g + e]) * p.alpha));
                  float o1, o2;
                  rope_pair(x1, x2, cc[e], ss[e], o1, o2);
                  r0[4 * g + e] = __float_as_uint(o1);
                  r1[4 * g + e] = __float_as_uint(o2);
                }
              }
            }
            if (tma) {
              epilogue_chunk_tma(p, &tmC, r0, epi_acquire(0), lane, m_warp0, (int64_t)n0 + cpar * 32);
              epilogue_chunk_tma(p, &tmC, r1, epi_acquire(1), lane, m_warp0, (int64_t)n0 + 64 + cpar * 32);
              nstore = 2;
            } else {
              epilogue_chunk(p, r0, pad, lane, b, m_warp0, (int64_t)n0 + cpar * 32, n_lim);
              epilogue_chunk(p, r1, pad, lane, b, m_warp0, (int64_t)n0 + 64 + cpar * 32, n_lim);
            }
          }
        }
      }
      if (!handled) {
#pragma unroll 1
        for (int mt = 0; mt < MT; ++mt) {
          const int64_t m_warp0 = (int64_t)m0 + mt * kBM + q * 32;
          if (m_warp0 >= p.a_rows) continue;  // sub-tile entirely out of range (warp-uniform)
#pragma unroll 1
          for (int c = cpar; c < L::kChunks; c += cstep) {
            if ((int64_t)n0 + c * 32 >= n_lim) break;
            uint32_t raw[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * (MT * L::kBNT) + (uint32_t)(mt * L::kBNT + c * 32), raw);
            tmem_ld_wait();
            if (tma && (int64_t)n0 + c * 32 + 32 <= n_lim) {
              uint8_t* buf = epi_acquire(nstore);
              epilogue_chunk_tma(p, &tmC, raw, buf, lane, m_warp0, (int64_t)n0 + c * 32);
              ++nstore;
            } else {
              if (!ring && nstore > 0) {   // the transposing path uses the whole pad: drain the stores that still read it
                if (lane == 0) tma_store_wait_read<0>();
                __syncwarp();
                nstore = 0;
              }
              epilogue_chunk(p, raw, pad, lane, b, m_warp0, (int64_t)n0 + c * 32, n_lim);
            }
          }
        }
      }
      // all TMEM reads of this warp are complete (wait::ld above): hand the accumulator stage back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (nstore > 0) {   // shared memory of the pads must outlive the bulk reads (next tile / CTA exit)
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (DIAG) {
    if (p.dbg_times && threadIdx.x == 64) {
      unsigned long long gt;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      p.dbg_times[(size_t)blockIdx.x * 8 + 5] = clock64();
      p.dbg_times[(size_t)blockIdx.x * 8 + 7] = (long long)gt;   // with [6]: wall-clock length of the CTA -> SM clock during the kernel
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)L::kTmemCols)
                 : "memory");
  }
}


// Experimental twin of gemm_tc_kernel (tuning hooks only: uvx_debug_gemm_cluster / uvx_debug_gemm_mode): adds thread-block
// clusters that share operand tiles by TMA multicast and the pipeline-isolation modes.  It is a separate kernel because the
// single-thread producer / MMA loops are latency-critical - compiling the extra code into the production kernel cost 3-8 %
// of the whole prefill step.  Measured outcome (profiles/r1_gemm_pipeline_isolation.md): multicast at cluster sizes <= 4
// does not reduce L2 traffic (the L2 already merges identical requests inside a short window), so production does not use it.
template <int MT, int BN>
__global__ void __launch_bounds__(SmemLayout<MT, BN>::kThreads, 1)
gemm_tc_kernel_x(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const GemmParams p) {
  pdl_trigger();
  using L = SmemLayout<MT, BN>;
  constexpr int kAcc = L::kAcc;
  constexpr int kEpiWarps = L::kEpiWarps;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full_bar = (uint64_t*)(smem + kBarOff);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full = empty_bar + kMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m_total = p.m_tiles * (int)p.a_batch;
  const int num_kb = (int)((p.K + kBK - 1) / kBK);
  const int stages = p.stages;
  // cluster geometry (cm = cn = 1: plain launch, no cluster instructions on the data path)
  const int cm = p.cm, cn = p.cn;
  const int csize = cm * cn;
  const bool clustered = csize > 1;
  const uint32_t crank = clustered ? cluster_rank() : 0u;
  const int rm = (int)crank % cm, rn = (int)crank / cm;
  uint16_t mask_a = 0, mask_w = 0;
  for (int j = 0; j < cn; ++j) mask_a |= (uint16_t)(1u << (rm + cm * j));   // CTAs working on the same row tile
  for (int i = 0; i < cm; ++i) mask_w |= (uint16_t)(1u << (i + cm * rn));   // CTAs working on the same column tile
  const uint16_t mask_e = mask_a | mask_w;
  const int m_groups = p.m_groups;
  const int num_units = p.m_groups * p.n_groups * p.splits;   // (cluster-level) work units
  const int unit0 = (int)blockIdx.x / csize, unit_step = (int)gridDim.x / csize;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();  // the 128B-swizzle atoms need a 1024-byte aligned ring
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], cm + cn - 1);  // one (multicast) commit from every CTA that reads what lands here
    }
    for (int a = 0; a < kAcc; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], kEpiWarps);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)L::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (clustered) cluster_barrier();  // peers' barriers are initialised before anything is multicast at them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // set-up above overlapped the previous kernel's tail; its outputs are visible from here on

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
      int s = 0;           // ring slot and its phase: the ring runs ahead across tile boundaries
      uint32_t ph = 0;
      for (int unit = unit0; unit < num_units; unit += unit_step) {
        const int tile = unit / p.splits, split = unit % p.splits;
        const int tm_idx = (tile % m_groups) * cm + rm;  // consecutive units share the W tiles
        const int tn_idx = (tile / m_groups) * cn + rn;
        const int b = tm_idx / p.m_tiles;  // padding tiles (index past the end) load zero-filled boxes
        const int m0 = (tm_idx % p.m_tiles) * (MT * kBM);
        const int n0 = tn_idx * BN;
        const int kb_begin = split * p.kb_per_split;
        const int kb_end = min(num_kb, kb_begin + p.kb_per_split);
        // every CTA walks K starting at a different k-block (and wraps): at M <= 256 all CTAs read the SAME activation
        // tile, and in lockstep they would all hit the same L2 lines at the same time
        const int nk = kb_end - kb_begin;
        const int rot = (int)(((unsigned)tile * 7u + (unsigned)split * 3u) % (unsigned)nk);
        for (int i = 0; i < nk; ++i) {
          const int kb = kb_begin + (i + rot < nk ? i + rot : i + rot - nk);
          mbar_wait(&empty_bar[s], ph ^ 1u);
          uint8_t* sa = smem + s * p.stage_bytes;
          if (p.dbg_mode == 2) {
            mbar_arrive(&full_bar[s]);
            if (++s == stages) { s = 0; ph ^= 1u; }
            continue;
          }
          mbar_expect_tx(&full_bar[s], (uint32_t)p.stage_bytes);
          if (cn == 1) tma_load_3d(sa, &tmA, kb * kBK, m0, b, &full_bar[s]);
          else tma_load_3d_mc(sa + rn * p.a_slice_rows * (kBK * 2), &tmA, kb * kBK, m0 + rn * p.a_slice_rows, b, &full_bar[s], mask_a);
          if (cm == 1) tma_load_2d(sa + p.a_box_bytes, &tmW, kb * kBK, n0, &full_bar[s]);
          else tma_load_2d_mc(sa + p.a_box_bytes + rm * p.w_slice_rows * (kBK * 2), &tmW, kb * kBK, n0 + rm * p.w_slice_rows, &full_bar[s], mask_w);
          if (++s == stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN);
      int s = 0;
      uint32_t ph = 0, tcount = 0;
      for (int unit = unit0; unit < num_units; unit += unit_step, ++tcount) {
        const int split = unit % p.splits;
        const int kb_begin = split * p.kb_per_split;
        const int kb_end = min(num_kb, kb_begin + p.kb_per_split);
        const uint32_t acc = tcount % kAcc;
        const uint32_t aph = (tcount / kAcc) & 1u;
        mbar_wait(&tmem_empty[acc], aph ^ 1u);  // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (MT * L::kBNT);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * p.stage_bytes);
          const uint64_t dw = make_smem_desc(sa + p.a_box_bytes);
          if (p.dbg_mode == 1) {
            mbar_arrive(&empty_bar[s]);
            if (++s == stages) { s = 0; ph ^= 1u; }
            continue;
          }
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            // rows past the A box (M small) read whatever follows in shared memory: they only feed accumulator rows
            // >= a_rows, which the epilogue never stores
            const uint64_t da = make_smem_desc(sa + mt * (kBM * kBK * 2));
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in the (addr >> 4) field
              umma_f16(d_tmem + mt * L::kBNT, da + (uint64_t)(2 * k), dw + (uint64_t)(2 * k), idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
            }
          }
          if (clustered) umma_commit_mc(&empty_bar[s], mask_e);  // tells every CTA that loads into this CTA's slot
          else umma_commit(&empty_bar[s]);                       // frees the smem slot once these MMAs have read it
          if (++s == stages) { s = 0; ph ^= 1u; }
        }
        umma_commit(&tmem_full[acc]);  // accumulators of this tile complete
      }
    }
  } else {
    // ---- epilogue: warp w reads TMEM lane quarter w % 4; with 8 warps the two warps of a quarter take alternate chunks
    const int q = warp & 3;
    const int cpar = (warp - 2) >> 2;              // 0 for the first four epilogue warps, 1 for the second four
    constexpr int cstep = kEpiWarps / 4;
    uint32_t tcount = 0;
    for (int unit = unit0; unit < num_units; unit += unit_step, ++tcount) {
      const int tile = unit / p.splits, split = unit % p.splits;
      const int tm_idx = (tile % m_groups) * cm + rm;
      const int tn_idx = (tile / m_groups) * cn + rn;
      const bool valid_tile = tm_idx < tiles_m_total && tn_idx < p.n_tiles;  // cluster padding computes zeros, stores nothing
      const int b = tm_idx / p.m_tiles;
      const int m0 = (tm_idx % p.m_tiles) * (MT * kBM);
      const int n0 = tn_idx * BN;
      const uint32_t acc = tcount % kAcc;
      const uint32_t aph = (tcount / kAcc) & 1u;
      mbar_wait(&tmem_full[acc], aph);
      tc_fence_after();
      const bool direct = p.splits == 1;
      if (!direct) {
        // split-K: park the raw fp32 partial tile in the workspace; splitk_reduce_kernel sums the splits in a fixed
        // order and applies the epilogue (deterministic, and the reduction is spread over every SM)
        float* part = p.ws_partial + (size_t)split * (size_t)(p.a_batch * p.a_rows) * (size_t)p.N;
#pragma unroll 1
        for (int mt = 0; mt < MT; ++mt) {
          if (!valid_tile || (int64_t)m0 + mt * kBM + q * 32 >= p.a_rows) continue;
          const int64_t m = (int64_t)m0 + mt * kBM + q * 32 + lane;
          const bool row_ok = m < p.a_rows;
#pragma unroll 1
          for (int c = cpar; c < L::kChunks; c += cstep) {
            uint32_t raw[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * (MT * L::kBNT) + (uint32_t)(mt * L::kBNT + c * 32), raw);
            tmem_ld_wait();
            if (row_ok) {
              uint4* dst = reinterpret_cast<uint4*>(part + ((size_t)b * p.a_rows + m) * p.N + n0 + c * 32);
#pragma unroll
              for (int g = 0; g < 8; ++g) dst[g] = make_uint4(raw[4 * g], raw[4 * g + 1], raw[4 * g + 2], raw[4 * g + 3]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);  // accumulator stage is free again
        continue;
      }
      float* pad = reinterpret_cast<float*>(smem + L::kPadOff) + (warp - 2) * (kStageBytesPerWarp / 4);
      const int64_t n_lim = (int64_t)n0 + BN < p.N ? (int64_t)n0 + BN : p.N;  // ragged last column tile (N % BN != 0)
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const int64_t m_warp0 = (int64_t)m0 + mt * kBM + q * 32;
        if (!valid_tile || m_warp0 >= p.a_rows) continue;  // sub-tile entirely out of range (warp-uniform)
#pragma unroll 1
        for (int c = cpar; c < L::kChunks; c += cstep) {
          if ((int64_t)n0 + c * 32 >= n_lim) break;
          uint32_t raw[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * (MT * L::kBNT) + (uint32_t)(mt * L::kBNT + c * 32), raw);
          tmem_ld_wait();
          epilogue_chunk(p, raw, pad, lane, b, m_warp0, (int64_t)n0 + c * 32, n_lim);
        }
      }
      // all TMEM reads of this warp are complete (wait::ld above): hand the accumulator stage back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (clustered) cluster_barrier();  // no CTA leaves while a peer can still multicast into it or signal its barriers
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)L::kTmemCols)
                 : "memory");
  }
}

// =====================================================================================================
// 2-SM variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x BN tile with ONE tcgen05.mma per k-step.
// CTA r stages rows [r*128, r*128+128) of A and rows [r*BN/2, (r+1)*BN/2) of W; the tensor core of the leader reads both
// halves, so each SM ingests only 16 KB + BN*64 B per k-block for 2*128*BN*64 flops of its own accumulator - twice the
// flops per byte of the 1-SM 128x128 tile (the 1-SM kernel is bound by bytes into the SM, see profiles/).
//   - both CTAs run a TMA producer; all bytes are signalled on the LEADER's full barrier (peer bit masked off);
//   - only the leader issues MMAs; tcgen05.commit.multicast frees the ring slot / publishes the accumulator in both CTAs;
//   - each CTA's epilogue drains its own 128 TMEM lanes; all 8 epilogue warps arrive on the leader's tmem_empty barrier.
static constexpr uint32_t kPeerMask = 0xFEFFFFFFu;  // clears the CTA-pair bit of a shared::cluster address -> leader CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerMask) : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(tm), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(tm), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

// Pair tile = 256 rows x (NH * BNH) columns: NH tcgen05.mma (N = BNH) per k-step into NH accumulators.  NH = 2 doubles the
// columns a row tile of A is used for (A is re-read once per column tile, and the GEMMs are bound by bytes delivered from
// L2): 256 x 416 for the Llama gate/up projection at M = 201, 256 x 512 for the encoder fc1 / qkv.
__device__ __forceinline__ void tma2_load_2d_e(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
      "}\n" ::"r"(smem_u32(dst)),
      "l"(tm), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d_e(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "@q cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
      "}\n" ::"r"(smem_u32(dst)),
      "l"(tm), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader_e(uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "@q mbarrier.arrive.shared::cluster.b64 _, [%0];\n"
      "}\n" ::"r"(smem_u32(bar) & kPeerMask)
      : "memory");
}
__device__ __forceinline__ void umma2_f16_e(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc_e(uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
      "}\n" ::"r"(smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

template <int BNH, int NH, int EPW>
struct Smem2 {
  static constexpr int kThreads = 64 + 32 * EPW;
  static constexpr int kABytes = kBM * kBK * 2;               // this CTA's 128 rows
  static constexpr int kWHalfBytes = (BNH / 2) * kBK * 2;     // this CTA's half of one N = BNH operand
  static constexpr int kWBytes = NH * kWHalfBytes;
  static constexpr int kStageBytes = kABytes + kWBytes;
  static constexpr int kBNT = BNH <= 128 ? 128 : 256;         // TMEM column stride of one accumulator
  static constexpr int kAcc = (NH * kBNT * 2 <= 512) ? 2 : 1;
  static constexpr int kTmemCols = kAcc * NH * kBNT;
  static constexpr int kChunks = (BNH + 31) / 32;
  static constexpr int kPadOff = kBarOff - EPW * kStageBytesPerWarp;
  static constexpr int kStages = kPadOff / kStageBytes > kMaxStages ? kMaxStages : kPadOff / kStageBytes;
  static_assert(kStages >= 2, "ring too shallow");
  static_assert((kTmemCols & (kTmemCols - 1)) == 0 && kTmemCols <= 512, "TMEM columns must be a power of two <= 512");
  static_assert((BNH / 2) % 8 == 0 && BNH % 16 == 0 && BNH <= 256, "UMMA N / swizzle-atom constraints");
};

template <int BNH, int NH, int EPW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Smem2<BNH, NH, EPW>::kThreads, 1)
gemm_tc2sm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const GemmParams p) {
  pdl_trigger();
  using L = Smem2<BNH, NH, EPW>;
  constexpr int kStages = L::kStages;
  constexpr int kAcc = L::kAcc;
  constexpr int BN = NH * BNH;  // columns of the pair tile
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full_bar = (uint64_t*)(smem + kBarOff);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full = empty_bar + kMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int tiles_m_total = p.m_tiles * (int)p.a_batch;  // m_tiles counts 256-row pair tiles
  const int num_kb = (int)((p.K + kBK - 1) / kBK);

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 2);   // leader's expect_tx arrive + the peer producer's arrive
      mbar_init(&empty_bar[s], 1);  // one multicast commit from the leader's MMA thread
    }
    for (int a = 0; a < kAcc; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 2 * EPW);  // epilogue warps of both CTAs (used on the leader only)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)L::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // set-up above overlapped the previous kernel's tail; its outputs are visible from here on

  if (warp == 0) {
    {  // convergent producer warp, elect.sync issues (see umma_f16_e)
      if (lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
      }
      __syncwarp();
      uint32_t it = 0;
      for (int tile = pair; tile < p.num_tiles; tile += num_pairs) {
        const int tm_idx = tile % tiles_m_total;
        const int tn_idx = tile / tiles_m_total;
        const int b = tm_idx / p.m_tiles;
        const int m0 = (tm_idx % p.m_tiles) * (2 * kBM) + (int)rank * kBM;
        const int n0 = tn_idx * BN + (int)rank * (BNH / 2);  // operand h of this CTA: W rows [n0 + h*BNH, + BNH/2)
        const int rot = (int)(((unsigned)tile * 7u) % (unsigned)num_kb);
        for (int i = 0; i < num_kb; ++i, ++it) {
          const int kb = i + rot < num_kb ? i + rot : i + rot - num_kb;
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          uint8_t* sa = smem + s * L::kStageBytes;
          tma2_load_3d_e(sa, &tmA, kb * kBK, m0, b, &full_bar[s]);
#pragma unroll
          for (int h = 0; h < NH; ++h)
            tma2_load_2d_e(sa + L::kABytes + h * L::kWHalfBytes, &tmW, kb * kBK, n0 + h * BNH, &full_bar[s]);
          if (leader) mbar_expect_tx_e(&full_bar[s], (uint32_t)(2 * L::kStageBytes));
          else mbar_arrive_leader_e(&full_bar[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {  // convergent MMA warp of the leader CTA
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BNH >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      uint32_t it = 0, tcount = 0;
      for (int tile = pair; tile < p.num_tiles; tile += num_pairs, ++tcount) {
        const uint32_t acc = tcount % kAcc;
        const uint32_t aph = (tcount / kAcc) & 1u;
        mbar_wait(&tmem_empty[acc], aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (NH * L::kBNT);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1u;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * L::kStageBytes);
          const uint64_t da = make_smem_desc(sa);
#pragma unroll
          for (int h = 0; h < NH; ++h) {
            const uint64_t dw = make_smem_desc(sa + L::kABytes + h * L::kWHalfBytes);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              umma2_f16_e(d_tmem + h * L::kBNT, da + (uint64_t)(2 * k), dw + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma2_commit_mc_e(&empty_bar[s]);
        }
        umma2_commit_mc_e(&tmem_full[acc]);
      }
      // keep the leader's barriers alive until the last epilogue arrivals from the peer have landed
      if (tcount > 0) {
        const uint32_t last = tcount - 1;
        mbar_wait(&tmem_empty[last % kAcc], (last / kAcc) & 1u);
      }
    }
  } else {
    const int q = warp & 3;
    const int cpar = (warp - 2) >> 2;
    constexpr int cstep = EPW / 4;
    uint32_t tcount = 0;
    for (int tile = pair; tile < p.num_tiles; tile += num_pairs, ++tcount) {
      const int tm_idx = tile % tiles_m_total;
      const int tn_idx = tile / tiles_m_total;
      const int b = tm_idx / p.m_tiles;
      const int m0 = (tm_idx % p.m_tiles) * (2 * kBM) + (int)rank * kBM;
      const uint32_t acc = tcount % kAcc;
      const uint32_t aph = (tcount / kAcc) & 1u;
      mbar_wait(&tmem_full[acc], aph);
      tc_fence_after();
      float* pad = reinterpret_cast<float*>(smem + L::kPadOff) + (warp - 2) * (kStageBytesPerWarp / 4);
      const int64_t m_warp0 = (int64_t)m0 + q * 32;
      if (m_warp0 < p.a_rows) {
#pragma unroll 1
        for (int h = 0; h < NH; ++h) {
          const int64_t n0 = (int64_t)tn_idx * BN + h * BNH;
          const int64_t n_lim = n0 + BNH < p.N ? n0 + BNH : p.N;  // ragged last column tile
#pragma unroll 1
          for (int c = cpar; c < L::kChunks; c += cstep) {
            if (n0 + c * 32 >= n_lim) break;
            uint32_t raw[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * (NH * L::kBNT) + (uint32_t)(h * L::kBNT + c * 32), raw);
            tmem_ld_wait();
            epilogue_chunk(p, raw, pad, lane, b, m_warp0, n0 + c * 32, n_lim);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)L::kTmemCols)
                 : "memory");
  }
}

// Split-K second pass: sums the fp32 partials in split order and applies the same epilogue as the direct path.
// One thread per 8 consecutive output columns; spread over the whole grid so no single SM has to pull all partials.
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const GemmParams p) {
  pdl_trigger();
  pdl_wait();
  const int64_t rows = p.a_batch * p.a_rows;
  const int64_t vec_per_row = p.N / 8;
  const int64_t total = rows * vec_per_row;
  const size_t split_stride = (size_t)rows * (size_t)p.N;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / vec_per_row, n = (idx % vec_per_row) * 8;
    const int64_t b = r / p.a_rows, m = r % p.a_rows;
    const float* src = p.ws_partial + (size_t)r * p.N + n;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int sidx = 0; sidx < p.splits; ++sidx) {
      const float4 t0 = __ldcg(reinterpret_cast<const float4*>(src + sidx * split_stride));
      const float4 t1 = __ldcg(reinterpret_cast<const float4*>(src + sidx * split_stride) + 1);
      v[0] += t0.x; v[1] += t0.y; v[2] += t0.z; v[3] += t0.w;
      v[4] += t1.x; v[5] += t1.y; v[6] += t1.z; v[7] += t1.w;
    }
    const int64_t orow = p.c_row_map ? (int64_t)p.c_row_map[r] : b * p.c_batch_rows + m + p.c_row_offset;
    if (orow < 0) continue;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= p.alpha;
    if (p.bias) {
      float t[8];
      unpack8(*reinterpret_cast<const bf16x8*>(p.bias + n), t);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += t[i];
    }
    if (p.act == UVX_ACT_GELU) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = gelu_erf(v[i]);
    }
    if (p.R) {
      float t[8];
      unpack8(*reinterpret_cast<const bf16x8*>(p.R + b * p.r_batch_stride + m * p.r_row_stride + n), t);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += t[i];
    }
    if (p.out_f32) {
      float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + orow * p.c_row_stride + n);
      dst[0] = make_float4(v[0], v[1], v[2], v[3]);
      dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      *reinterpret_cast<bf16x8*>(reinterpret_cast<bf16*>(p.C) + orow * p.c_row_stride + n) = pack8(v);
    }
  }
}

// Split-K second pass fused with the RMSNorm that follows o_proj / down_proj in every Llama layer: one CTA per output row
// sums the partials, applies the epilogue (bias / residual), writes the residual stream C and, from the same registers,
// norm_out = w * bf16(C * rsqrt(mean(C^2) + eps)) (LlamaRMSNorm rounding order on the bf16-rounded C).  bf16 output only.
__global__ void __launch_bounds__(256) splitk_reduce_rmsnorm_kernel(const GemmParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[32];
  const int64_t r = blockIdx.x;
  const int64_t b = r / p.a_rows, m = r % p.a_rows;
  const int64_t rows = p.a_batch * p.a_rows;
  const size_t split_stride = (size_t)rows * (size_t)p.N;
  const int64_t orow = p.c_row_map ? (int64_t)p.c_row_map[r] : b * p.c_batch_rows + m + p.c_row_offset;
  constexpr int kMaxV = 4;  // up to 256 * 4 * 8 = 8192 columns
  float v[kMaxV][8];
  float sq = 0.f;
  const int nvec = (int)(p.N / 8);
#pragma unroll
  for (int i = 0; i < kMaxV; ++i) {
    const int j = threadIdx.x + i * 256;
    if (j < nvec) {
      const int64_t n = (int64_t)j * 8;
      const float* src = p.ws_partial + (size_t)r * p.N + n;
#pragma unroll
      for (int e = 0; e < 8; ++e) v[i][e] = 0.f;
      // loads of four splits are issued together (the kernel is a latency chain: partials -> row sum -> write), sums stay in
      // split order so the result does not depend on the batching
      for (int s0 = 0; s0 < p.splits; s0 += 4) {
        float4 t0[4], t1[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (s0 + u < p.splits) {
            t0[u] = __ldcg(reinterpret_cast<const float4*>(src + (s0 + u) * split_stride));
            t1[u] = __ldcg(reinterpret_cast<const float4*>(src + (s0 + u) * split_stride) + 1);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (s0 + u < p.splits) {
            v[i][0] += t0[u].x; v[i][1] += t0[u].y; v[i][2] += t0[u].z; v[i][3] += t0[u].w;
            v[i][4] += t1[u].x; v[i][5] += t1[u].y; v[i][6] += t1[u].z; v[i][7] += t1[u].w;
          }
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) v[i][e] *= p.alpha;
      if (p.bias) {
        float t[8];
        unpack8(*reinterpret_cast<const bf16x8*>(p.bias + n), t);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[i][e] += t[e];
      }
      if (p.R) {
        float t[8];
        unpack8(*reinterpret_cast<const bf16x8*>(p.R + b * p.r_batch_stride + m * p.r_row_stride + n), t);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[i][e] += t[e];
      }
      const bf16x8 packed = pack8(v[i]);
      if (orow >= 0) *reinterpret_cast<bf16x8*>(reinterpret_cast<bf16*>(p.C) + orow * p.c_row_stride + n) = packed;
      unpack8(packed, v[i]);  // the norm sees the bf16-rounded residual stream, exactly like a separate kernel would
#pragma unroll
      for (int e = 0; e < 8; ++e) sq += v[i][e] * v[i][e];
    }
  }
  const float rstd = rsqrtf(block_sum(sq, red) / (float)p.N + p.norm_eps);
  if (orow < 0) return;
#pragma unroll
  for (int i = 0; i < kMaxV; ++i) {
    const int j = threadIdx.x + i * 256;
    if (j < nvec) {
      float wv[8], o[8];
      unpack8(*reinterpret_cast<const bf16x8*>(p.norm_w + (int64_t)j * 8), wv);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = wv[e] * __bfloat162float(__float2bfloat16_rn(v[i][e] * rstd));
      *reinterpret_cast<bf16x8*>(p.norm_out + orow * p.N + (int64_t)j * 8) = pack8(o);
    }
  }
}

// ---------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

static int encode_map(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, CUtensorMapL2promotion promo) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return UVX_ERR_CUDA;
  }
  cuuint64_t gd[3];
  cuuint64_t gs[2];
  cuuint32_t bx[3], es[3];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i < rank - 1; ++i) gs[i] = strides_bytes[i];
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu %llu stride %llu)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)strides_bytes[0]);
    return UVX_ERR_CUDA;
  }
  return UVX_OK;
}

// tensor map without swizzle (output boxes of the TMA-store epilogue): rank 2 bf16 (C) or rank 3 fp32 (split-K partials)
static int encode_map_plain(CUtensorMap* tm, const void* base, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                            int rank = 2, bool f32 = false) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return UVX_ERR_CUDA;
  }
  cuuint64_t gd[3] = {dims[0], dims[1], rank > 2 ? dims[2] : 1};
  cuuint64_t gs[2] = {strides_bytes[0], rank > 2 ? strides_bytes[1] : 0};
  cuuint32_t bx[3] = {box[0], box[1], 1}, es[3] = {1, 1, 1};
  CUresult r = enc(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd,
                   gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (output map) failed with CUresult %d (dims %llu %llu stride %llu)", (int)r, (unsigned long long)dims[0],
              (unsigned long long)dims[1], (unsigned long long)strides_bytes[0]);
    return UVX_ERR_CUDA;
  }
  return UVX_OK;
}

// tuning switch (uvx_debug_gemm_tma_store / UVX_TMA_STORE), bit mask: 1 = bf16 outputs without residual through TMA stores, 2 = also
// with a residual (within noise of 1 once those GEMMs run the unchanged round-1 loops, profiles/r2_ab_bench_v5.txt; default 1),
// 4 = fp32 split-K partials
// (measured slower than the direct 16-byte stores); 0 = the transposing epilogue everywhere
static int g_gemm_tma_store = -1;  // -1: UVX_TMA_STORE env or 1
static int g_gemm_epi_ring = -1;   // -1: UVX_EPI_RING env or 1 (tuning: 0 = stage TMA-store boxes in the per-warp pads only)

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static int g_gemm_pf = -1;  // L2 prefetch distance in k-blocks (UVX_GEMM_PF / uvx_debug_gemm_pf; default 0)
static int gemm_pf() {
  if (g_gemm_pf < 0) {
    const char* e = getenv("UVX_GEMM_PF");
    g_gemm_pf = e ? atoi(e) : 0;  // measured: prefetching ahead of the ring only slows the stream down (profiles/r2_ws_sweep_v1.txt)
    if (g_gemm_pf < 0) g_gemm_pf = 0;
  }
  return g_gemm_pf;
}

static int g_gemm_stage_cap = 0;  // tuning only (uvx_debug_gemm_stages): upper bound on the ring depth
static long long* g_gemm_times = nullptr;  // tuning only (uvx_debug_gemm_times): device buffer [grid][8] of the DIAG twin

template <int MT, int BN, int EW = 0, bool DIAG = false>
static int launch_gemm(const uvx_gemm_args* a, int splits, int cm, int cn, cudaStream_t stream) {
  using L = SmemLayout<MT, BN, EW>;
  CUtensorMap tmA, tmW;
  // one row tile covers the whole problem: stage only round8(M) rows of A per k-block (more ring stages fit)
  if (cm < 1 || cn < 1 || cm * cn > 8 || (BN / cm) % 8 != 0 || BN % cm != 0) cm = cn = 1;
  int a_box_rows = MT * kBM;
  if (a->a_batch == 1 && a->a_rows <= MT * kBM) a_box_rows = (int)((a->a_rows + 8 * cn - 1) / (8 * cn) * (8 * cn));
  if (a_box_rows % (8 * cn) != 0) cn = 1;
  const int a_slice_rows = a_box_rows / cn, w_slice_rows = BN / cm;
  {
    uint64_t dims[3] = {(uint64_t)a->K, (uint64_t)a->a_rows, (uint64_t)a->a_batch};
    uint64_t st[2] = {(uint64_t)a->a_row_stride * 2, (uint64_t)(a->a_batch > 1 ? a->a_batch_stride : a->a_row_stride) * 2};
    uint32_t box[3] = {kBK, (uint32_t)a_slice_rows, 1};
    int rc = encode_map(&tmA, a->A, 3, dims, st, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (rc) return rc;
  }
  const bool tiled = a->w_tiled != 0;
  if (tiled) {
    // pre-tiled image [n_tiles][K/64][BN][64] (built once by the host, ops.tile_weight): a 2-D map over [rows, 64] whose
    // BN-row boxes are contiguous BN*128-byte runs - the DRAM access pattern of a plain copy (measured 6.33 TB/s vs 5.66 TB/s
    // for 128-byte pieces of 208 rows 8 KB apart, profiles/r2_probe_ws.txt)
    UVX_REQUIRE(a->w_tiled == BN && cm == 1 && cn == 1 && a->K % kBK == 0, "uvx_gemm_bf16: tiled weights need BN == tile rows and K %% 64 == 0");
    const uint64_t n_tiles_w = (uint64_t)((a->N + BN - 1) / BN);
    uint64_t dims[2] = {(uint64_t)kBK, n_tiles_w * (uint64_t)(a->K / kBK) * (uint64_t)BN};
    uint64_t st[1] = {(uint64_t)kBK * 2};
    uint32_t box[2] = {kBK, (uint32_t)BN};
    int rc = encode_map(&tmW, a->W, 2, dims, st, box, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (rc) return rc;
  } else {
    uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->N};
    uint64_t st[1] = {(uint64_t)a->w_row_stride * 2};
    uint32_t box[2] = {kBK, (uint32_t)w_slice_rows};
    int rc = encode_map(&tmW, a->W, 2, dims, st, box, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (rc) return rc;
  }
  GemmParams p;
  CUtensorMap tmC = tmA;
  p.tma_store = 0;
  p.w_tiled = tiled ? 1 : 0;
  p.dbg_times = DIAG ? g_gemm_times : nullptr;
  p.pf = gemm_pf();
  p.swiglu = a->act == UVX_ACT_SWIGLU ? 1 : 0;
  p.rope_cos = a->rope_cos;
  p.rope_sin = a->rope_sin;
  p.rope_pos = a->rope_positions;
  p.rope_rows_per_seq = a->rope_rows_per_seq > 0 ? a->rope_rows_per_seq : 1;
  p.rope_pos_offset = a->rope_pos_offset;
  p.rope_cols = a->rope_cos ? a->rope_cols : 0;
  if (p.swiglu || p.rope_cos) splits = 1;  // fused epilogues are tile-local: no split-K
  p.a_rows = a->a_rows;
  p.a_batch = a->a_batch;
  p.K = a->K;
  p.N = a->N;
  p.C = a->C;
  p.c_row_stride = a->c_row_stride;
  p.c_batch_rows = a->c_batch_rows;
  p.c_row_offset = a->c_row_offset;
  p.c_row_map = a->c_row_map;
  p.bias = (const bf16*)a->bias;
  p.R = (const bf16*)a->R;
  p.r_row_stride = a->r_row_stride;
  p.r_batch_stride = a->r_batch_stride;
  p.alpha = a->alpha;
  p.act = a->act;
  p.out_f32 = a->out_dtype == UVX_DT_F32;
  p.norm_w = (const bf16*)a->norm_w;
  p.norm_out = (bf16*)a->norm_out;
  p.norm_eps = a->norm_eps;
  p.m_tiles = (int)((a->a_rows + MT * kBM - 1) / (MT * kBM));
  p.n_tiles = (int)((a->N + BN - 1) / BN);
  p.num_tiles = p.m_tiles * (int)a->a_batch * p.n_tiles;
  const int num_kb = (int)((a->K + kBK - 1) / kBK);
  if (BN % 32 != 0 || a->N % BN != 0) splits = 1;  // ragged column tiles exist only in the direct epilogue
  // split-K needs the caller's workspace: [splits][rows][N] fp32 partial sums
  const size_t per_split = (size_t)a->a_batch * (size_t)a->a_rows * (size_t)a->N * 4;
  // (the last 1 KB of the workspace belongs to gemm_ws.cu's slot flags)
  while (splits > 1 && (!a->workspace || per_split * (size_t)splits + 1024 > (size_t)a->workspace_bytes)) --splits;
  if (splits > num_kb) splits = num_kb;
  if (splits < 1) splits = 1;
  p.kb_per_split = (num_kb + splits - 1) / splits;
  p.splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;  // no empty split
  p.ws_partial = (float*)a->workspace;
  if (g_gemm_tma_store < 0) {
    const char* e = getenv("UVX_TMA_STORE");
    g_gemm_tma_store = e ? atoi(e) : 1;
  }
  if (g_gemm_epi_ring < 0) {
    const char* e = getenv("UVX_EPI_RING");
    g_gemm_epi_ring = e ? atoi(e) : 1;
  }
  if (cm == 1 && cn == 1 && g_gemm_tma_store && a->a_batch == 1 && !(a->flags & 1)) {
    if (p.splits > 1) {
      if (g_gemm_tma_store & 4) {
      // split-K: fp32 partial tiles leave through TMA stores into [split][row][col] (rows past M clipped by the map)
      uint64_t dims[3] = {(uint64_t)a->N, (uint64_t)a->a_rows, (uint64_t)p.splits};
      uint64_t st[2] = {(uint64_t)a->N * 4, (uint64_t)a->a_rows * (uint64_t)a->N * 4};
      uint32_t box[2] = {16, 32};
      int rc = encode_map_plain(&tmC, a->workspace, dims, st, box, 3, true);
      if (rc) return rc;
      p.tma_store = 1;
      }
    } else if (a->out_dtype == UVX_DT_BF16 && !a->c_row_map && a->c_row_offset == 0 && (!a->R || (g_gemm_tma_store & 2))) {
      // bf16 output in plain row order: the epilogue stores 32 x 32 boxes by TMA (tensor map of C, no swizzle: the 64-byte box
      // rows are written by 32 lanes with at most a 4-way bank conflict on four stores per chunk); fused SwiGLU: 32 x 16 boxes
      const bool sw = a->act == UVX_ACT_SWIGLU;   // [M, N/2] output, 16 columns per accumulator chunk
      uint64_t dims[2] = {(uint64_t)(sw ? a->N / 2 : a->N), (uint64_t)a->a_rows};
      uint64_t st[1] = {(uint64_t)a->c_row_stride * 2};
      uint32_t box[2] = {sw ? 16u : 32u, 32};
      int rc = encode_map_plain(&tmC, a->C, dims, st, box);
      if (rc) return rc;
      p.tma_store = 1;
    }
  }
  p.a_box_bytes = a_box_rows * kBK * 2;
  p.stage_bytes = p.a_box_bytes + L::kWBytes;
  p.stages = L::kPadOff / p.stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  if (g_gemm_stage_cap >= 2 && p.stages > g_gemm_stage_cap) p.stages = g_gemm_stage_cap;
  p.cm = cm;
  p.dbg_mode = g_gemm_dbg;
  p.cn = cn;
  p.a_slice_rows = a_slice_rows;
  p.w_slice_rows = w_slice_rows;
  p.m_groups = (p.m_tiles * (int)a->a_batch + cm - 1) / cm;
  p.n_groups = (p.n_tiles + cn - 1) / cn;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<MT, BN, EW, DIAG, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc_kernel<MT, BN, EW, DIAG, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc_kernel_x<MT, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(gemm_tc_kernel<%d,%d>, smem %d): %s", MT, BN, kSmemTotal, cudaGetErrorString(e));
      return UVX_ERR_CUDA;
    }
    attr_set = true;
  }
  const int units = p.m_groups * p.n_groups * p.splits;  // cluster-level work units
  const int csize = cm * cn;
  p.epi_ring = (g_gemm_epi_ring && p.tma_store && csize == 1 && units <= num_sms() && p.stages * p.stage_bytes >= L::kEpiWarps * 16384) ? 1 : 0;
  p.pdl = pdl_enabled() ? 1 : 0;
  UVX_REQUIRE(!(p.dbg_mode && csize > 1), "uvx_gemm_bf16: pipeline-isolation modes are for unclustered launches");
  UVX_REQUIRE(EW == 0 || csize == 1, "uvx_gemm_bf16: cluster launches use the default epilogue width");
  if (csize == 1) {
    const int grid = units < num_sms() ? units : num_sms();
    if (p.dbg_mode && !DIAG && EW == 0) launch_k(gemm_tc_kernel_x<MT, BN>, dim3((unsigned)grid), dim3(SmemLayout<MT, BN>::kThreads), kSmemTotal, stream, tmA, tmW, p);
    else if (tiled || p.pf || p.tma_store || p.swiglu || p.rope_cos || DIAG || EW)
      launch_k(gemm_tc_kernel<MT, BN, EW, DIAG, 3>, dim3((unsigned)grid), dim3(L::kThreads), kSmemTotal, stream, tmA, tmW, tmC, p);
    else   // none of the round-2 features applies (residual / split-K / fp32 / remapped outputs): the round-1 loops, unchanged
      launch_k(gemm_tc_kernel<MT, BN, EW, DIAG, 0>, dim3((unsigned)grid), dim3(L::kThreads), kSmemTotal, stream, tmA, tmW, tmC, p);
  } else {
    static int max_clusters[9] = {0};
    if (max_clusters[csize] == 0) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(num_sms() / csize * csize));
      cfg.blockDim = dim3(SmemLayout<MT, BN>::kThreads);
      cfg.dynamicSmemBytes = kSmemTotal;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = (unsigned)csize;
      at[0].val.clusterDim.y = at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int n = 0;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, gemm_tc_kernel_x<MT, BN>, &cfg);
      if (e != cudaSuccess || n < 1) {
        set_error("cudaOccupancyMaxActiveClusters(gemm_tc_kernel<%d,%d>, cluster %d): %s", MT, BN, csize, cudaGetErrorString(e));
        return UVX_ERR_CUDA;
      }
      max_clusters[csize] = n;
    }
    const int clusters = units < max_clusters[csize] ? units : max_clusters[csize];
    launch_k_cluster(gemm_tc_kernel_x<MT, BN>, dim3((unsigned)(clusters * csize)), dim3(SmemLayout<MT, BN>::kThreads), kSmemTotal, stream,
                     (unsigned)csize, tmA, tmW, p);
  }
  int rc = check_launch("gemm_tc_kernel");
  if (rc) return rc;
  const bool want_norm = a->norm_w && a->norm_out;
  if (p.splits == 1) {
    if (!want_norm) return UVX_OK;
    // direct epilogue (tile-local): the row norm runs as its own kernel on the finished rows
    return uvx_rmsnorm(a->C, a->norm_w, a->norm_out, a->a_batch * a->a_rows, a->N, a->c_row_stride, 0, 0, 0, a->norm_eps, stream);
  }
  if (want_norm && !p.out_f32 && a->N <= 8192 && !a->c_row_map && a->act == UVX_ACT_NONE) {
    launch_k(splitk_reduce_rmsnorm_kernel, dim3((unsigned)(a->a_batch * a->a_rows)), dim3(256), 0, stream, p);
    return check_launch("splitk_reduce_rmsnorm_kernel");
  }
  const int64_t total = a->a_batch * a->a_rows * (a->N / 8);
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)num_sms() * 8) blocks = (int64_t)num_sms() * 8;
  launch_k(splitk_reduce_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, p);
  rc = check_launch("splitk_reduce_kernel");
  if (rc || !want_norm) return rc;
  return uvx_rmsnorm(a->C, a->norm_w, a->norm_out, a->a_batch * a->a_rows, a->N, a->c_row_stride, 0, 0, 0, a->norm_eps, stream);
}

template <int BNH, int NH, int EPW>
static int launch_gemm_2sm(const uvx_gemm_args* a, cudaStream_t stream) {
  using L = Smem2<BNH, NH, EPW>;
  constexpr int BN = NH * BNH;
  CUtensorMap tmA, tmW;
  {
    uint64_t dims[3] = {(uint64_t)a->K, (uint64_t)a->a_rows, (uint64_t)a->a_batch};
    uint64_t st[2] = {(uint64_t)a->a_row_stride * 2, (uint64_t)(a->a_batch > 1 ? a->a_batch_stride : a->a_row_stride) * 2};
    uint32_t box[3] = {kBK, (uint32_t)kBM, 1};
    int rc = encode_map(&tmA, a->A, 3, dims, st, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->N};
    uint64_t st[1] = {(uint64_t)a->w_row_stride * 2};
    uint32_t box[2] = {kBK, (uint32_t)(BNH / 2)};
    int rc = encode_map(&tmW, a->W, 2, dims, st, box, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (rc) return rc;
  }
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.a_rows = a->a_rows;
  p.a_batch = a->a_batch;
  p.K = a->K;
  p.N = a->N;
  p.C = a->C;
  p.c_row_stride = a->c_row_stride;
  p.c_batch_rows = a->c_batch_rows;
  p.c_row_offset = a->c_row_offset;
  p.c_row_map = a->c_row_map;
  p.bias = (const bf16*)a->bias;
  p.R = (const bf16*)a->R;
  p.r_row_stride = a->r_row_stride;
  p.r_batch_stride = a->r_batch_stride;
  p.alpha = a->alpha;
  p.act = a->act;
  p.out_f32 = a->out_dtype == UVX_DT_F32;
  p.m_tiles = (int)((a->a_rows + 2 * kBM - 1) / (2 * kBM));
  p.n_tiles = (int)((a->N + BN - 1) / BN);
  p.num_tiles = p.m_tiles * (int)a->a_batch * p.n_tiles;
  p.splits = 1;
  p.kb_per_split = (int)((a->K + kBK - 1) / kBK);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc2sm_kernel<BNH, NH, EPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(gemm_tc2sm_kernel<%d,%d>, smem %d): %s", BNH, NH, kSmemTotal, cudaGetErrorString(e));
      return UVX_ERR_CUDA;
    }
    attr_set = true;
  }
  const int max_pairs = num_sms() / 2;
  const int pairs = p.num_tiles < max_pairs ? p.num_tiles : max_pairs;
  launch_k(gemm_tc2sm_kernel<BNH, NH, EPW>, dim3((unsigned)(2 * pairs)), dim3(L::kThreads), kSmemTotal, stream, tmA, tmW, p);
  int rc = check_launch("gemm_tc2sm_kernel");
  if (rc || !(a->norm_w && a->norm_out)) return rc;
  return uvx_rmsnorm(a->C, a->norm_w, a->norm_out, a->a_batch * a->a_rows, a->N, a->c_row_stride, 0, 0, 0, a->norm_eps, stream);
}

}  // namespace uvx

// Tile / split selection.  cfg = MT*1000 + BN; UVX_GEMM_CFG / UVX_GEMM_SPLITS (env, tuning only) override the heuristic.
static int forced = -1, forced_splits = -1, forced_cm = 0, forced_cn = 0;

// tuning hook (scripts/gemm_sweep.py): force a tile config (MT*1000+BN, 0 = heuristic) and a split count (0 = heuristic)
extern "C" int uvx_debug_gemm_override(int cfg, int splits) {
  forced = cfg;
  forced_splits = splits;
  return UVX_OK;
}

// tuning hook: 1 = run the 1-SM kernel without its MMAs, 2 = without its TMA loads (results are garbage; timing only)
extern "C" int uvx_debug_gemm_mode(int mode) {
  uvx::g_gemm_dbg = mode;
  return UVX_OK;
}

// tuning hook: 1 = TMA-store epilogue where eligible (default), 0 = transposing epilogue everywhere
extern "C" int uvx_debug_gemm_tma_store(int on) {
  uvx::g_gemm_tma_store = on;
  return UVX_OK;
}

// tuning hook: device buffer [grid][8] int64 the DIAG twin (cfg 7xxx / 8xxx) fills with per-CTA phase timestamps (NULL = off)
extern "C" int uvx_debug_gemm_times(void* dev_buf) {
  uvx::g_gemm_times = (long long*)dev_buf;
  return UVX_OK;
}

// tuning hook: cap the shared-memory ring depth (0 = as deep as fits)
extern "C" int uvx_debug_gemm_stages(int n) {
  uvx::g_gemm_stage_cap = n;
  return UVX_OK;
}

// tuning hook: L2 prefetch distance of the weight stream in k-blocks (0 = off, < 0 = default / UVX_GEMM_PF)
extern "C" int uvx_debug_gemm_pf(int pf) {
  uvx::g_gemm_pf = pf;
  return UVX_OK;
}

// tuning hook: force the thread-block cluster shape (cm row tiles x cn column tiles, 0 = heuristic)
extern "C" int uvx_debug_gemm_cluster(int cm, int cn) {
  forced_cm = cm;
  forced_cn = cn;
  return UVX_OK;
}

static void read_forced() {
  if (forced < 0) {
    const char* e = getenv("UVX_GEMM_CFG");
    forced = e ? atoi(e) : 0;
    const char* s = getenv("UVX_GEMM_SPLITS");
    forced_splits = s ? atoi(s) : 0;
  }
}

static void pick_cfg(int64_t rows, int64_t batch, int64_t N, int64_t K, int* cfg, int* splits) {
  read_forced();
  const int sms = 148;
  const int num_kb = (int)((K + 63) / 64);
  int mt, bn;
  if (rows > 128 && rows <= 256 && batch == 1) {
    // weight-streaming regime (LLM prefill at B=1, projector): one CTA tile spans every row, W is read once
    mt = 2;
    bn = (N % 256 == 0 && N / 256 >= sms / 2) ? 256 : (N % 128 == 0 ? 128 : 64);
    // 208-wide tiles (ragged last tile) when they spread the weight stream over more SMs per wave: N = 28672 is 112
    // tiles of 256 (76 % of the SMs) but 138 tiles of 208 (93 %)
    if (bn == 256) {
      const int64_t t256 = N / 256, t208 = (N + 207) / 208;
      if (((t208 + sms - 1) / sms) * 208 < ((t256 + sms - 1) / sms) * 256) bn = 208;
    }
    if (bn == 128 && 2 * (N / 128) >= (sms * 3) / 5 && num_kb <= 80) mt = 1;  // enough 128x128 tiles: skip split-K + reduce
  } else {
    // tensor-bound regime (encoder, training): 128-row tiles; 256-wide when that still gives >= ~1.3 waves of tiles
    mt = 1;
    const int64_t m_tiles = (rows + 127) / 128 * batch;
    if (N % 256 == 0 && m_tiles * (N / 256) >= (sms * 4) / 3) bn = 256;
    else bn = (N % 128 == 0 && m_tiles * (N / 128) >= sms / 2) ? 128 : 64;
  }
  int variant = 0;  // tuning variants: 6xxx = (2, xxx) with 8 epilogue warps, 7xxx = (2, xxx) diagnostic twin, 8xxx = both
  if (forced > 0 && (N % (forced % 1000) == 0 || forced % 1000 == 208 || forced / 1000 == 5)) {
    mt = forced / 1000;
    bn = forced % 1000;
    if (mt >= 6) {
      variant = mt;
      mt = 2;
    }
  }
  const int64_t tiles = ((rows + mt * 128 - 1) / (mt * 128)) * batch * ((N + bn - 1) / bn);
  int sp = 1;
  if (tiles * 2 <= sms && num_kb >= 16) {  // too few tiles to occupy the SMs: split K (>= 8 k-blocks per split)
    sp = (int)(sms / tiles);
    if (sp > num_kb / 8) sp = num_kb / 8;
    if (sp > 16) sp = 16;
    if (sp < 1) sp = 1;
  }
  if (forced_splits > 0) sp = forced_splits;
  *cfg = (variant ? variant : mt) * 1000 + bn;
  *splits = sp;
}

extern "C" int uvx_gemm_bf16(const uvx_gemm_args* a, uvx_stream_t stream_) {
  using namespace uvx;
  cudaStream_t stream = (cudaStream_t)stream_;
  UVX_REQUIRE(a && a->A && a->W && a->C, "uvx_gemm_bf16: null pointer");
  UVX_REQUIRE(a->a_batch >= 1 && a->a_rows >= 1 && a->K >= 8 && a->N >= 64, "uvx_gemm_bf16: empty problem");
  UVX_REQUIRE(a->K % 8 == 0 && a->N % 64 == 0, "uvx_gemm_bf16: K %% 8 and N %% 64 required (K=%lld N=%lld)",
              (long long)a->K, (long long)a->N);
  UVX_REQUIRE(a->a_row_stride % 8 == 0 && a->w_row_stride % 8 == 0 && (a->a_batch == 1 || a->a_batch_stride % 8 == 0),
              "uvx_gemm_bf16: strides must be multiples of 8 elements");
  UVX_REQUIRE(((uintptr_t)a->A % 16 == 0) && ((uintptr_t)a->W % 16 == 0) && ((uintptr_t)a->C % 16 == 0),
              "uvx_gemm_bf16: 16-byte aligned bases required");
  UVX_REQUIRE(a->c_row_stride % 8 == 0 && (!a->R || (a->r_row_stride % 8 == 0 && a->r_batch_stride % 8 == 0)),
              "uvx_gemm_bf16: output / residual strides must be multiples of 8");
  UVX_REQUIRE(a->a_rows < (1ll << 31) && a->K < (1ll << 31) && a->N < (1ll << 31), "uvx_gemm_bf16: dimension too large");
  UVX_REQUIRE(!a->workspace || (uintptr_t)a->workspace % 256 == 0, "uvx_gemm_bf16: workspace must be 256-byte aligned");
  UVX_REQUIRE(!(a->norm_w && a->norm_out) || (a->out_dtype == UVX_DT_BF16 && !a->c_row_map && a->c_row_offset == 0 &&
                                              (a->a_batch == 1 || a->c_batch_rows == a->a_rows)),
              "uvx_gemm_bf16: fused RMSNorm needs a plain bf16 output");
  // rows <= 256 in one batch: the weight-streaming form (tokens on the UMMA N dimension, stream-K; gemm_ws.cu)
  // (a forced tile configuration / split count / cluster shape / isolation mode addresses gemm_tc_kernel: those calls stay there)
  read_forced();
  if (forced <= 0 && forced_splits <= 0 && forced_cm == 0 && g_gemm_dbg == 0 && gemm_ws_eligible(a)) return launch_gemm_ws(a, stream);
  UVX_REQUIRE(a->w_perm == 0, "uvx_gemm_bf16: a permuted weight image (w_perm) is read by the weight-streaming form only (UVX_GEMM_WS=1, rows <= 256)");
  int cfg, splits, cm = 1, cn = 1;
  pick_cfg(a->a_rows, a->a_batch, a->N, a->K, &cfg, &splits);
  if (forced_cm > 0 && forced_cn > 0) {
    cm = forced_cm;
    cn = forced_cn;
  }
  // Deep-K, at most one wave of 128 x 128 tiles (the Whisper fc2: 1500 x 1280 x 5120 = 120 tiles): the 2-SM pair kernel (256 x 128
  // pair tiles, each CTA stages its 128 rows of A and HALF of the weight tile, 8 epilogue warps) moves 25 % fewer bytes into shared
  // memory per k-block and the main loop dominates - 22.8 vs 26.5 us (profiles/r2_enc_cfg_sweep2.txt).  Shallow-K shapes stay on the
  // 1-SM kernel (their exposed epilogue is shorter there).
  static int pair_fc2 = -1;
  if (pair_fc2 < 0) {
    const char* e = getenv("UVX_PAIR_DEEPK");
    pair_fc2 = e ? atoi(e) : 1;
  }
  if (pair_fc2 && forced <= 0 && forced_splits <= 0 && forced_cm == 0 && g_gemm_dbg == 0 && a->a_batch == 1 && a->a_rows > 256 &&
      a->N % 128 == 0 && a->K >= 4096 && ((a->a_rows + 127) / 128) * (a->N / 128) <= 148 && !a->w_tiled && !a->rope_cos &&
      a->act != UVX_ACT_SWIGLU && !a->norm_w && !a->c_row_map && a->c_row_offset == 0 && a->out_dtype == UVX_DT_BF16)
    cfg = 9128;
  UVX_REQUIRE(a->act != UVX_ACT_SWIGLU || (a->w_tiled == 208 && !a->bias && !a->R && !a->norm_w && a->out_dtype == UVX_DT_BF16 && a->N % 16 == 0),
              "uvx_gemm_bf16: UVX_ACT_SWIGLU needs the 208-row interleaved gate|up image, bf16 output, no bias / residual / norm");
  UVX_REQUIRE(!a->rope_cos || (a->rope_sin && a->a_batch == 1 && a->alpha == 1.0f && !a->bias && !a->R && a->act == UVX_ACT_NONE &&
                               a->rope_cols % 128 == 0 && (a->w_tiled == 0 || a->w_tiled == 128) && a->N % 128 == 0),
              "uvx_gemm_bf16: fused RoPE needs head_dim 128 tiles, a_batch 1, no bias / residual / activation");
  if (a->w_tiled) {
    // the image was cut for one tile width: that width is the configuration (row sub-tiles still follow the row count)
    UVX_REQUIRE(a->w_tiled == 64 || a->w_tiled == 128 || a->w_tiled == 208 || a->w_tiled == 256, "uvx_gemm_bf16: w_tiled must be 64 / 128 / 208 / 256");
    int mt = cfg / 1000;
    const int variant = mt >= 6 ? mt : 0;
    if (mt != 1 && mt != 2) mt = (a->a_rows > 128 && a->a_rows <= 256 && a->a_batch == 1) ? 2 : 1;
    if (a->rope_cos) mt = 1;
    cfg = (variant && mt == 2 ? variant : mt) * 1000 + (int)a->w_tiled;
    if (a->act == UVX_ACT_SWIGLU && mt == 2 && !variant) {
      // the fused SwiGLU epilogue is arithmetic on the exposed tail of a one-tile CTA: UVX_SWIGLU_EW=8 spreads it over eight warps
      static int sw_ew = -1;
      if (sw_ew < 0) {
        const char* e = getenv("UVX_SWIGLU_EW");
        sw_ew = e ? atoi(e) : 4;
      }
      if (sw_ew == 8) cfg = 6000 + (int)a->w_tiled;
    }
    cm = cn = 1;
    const int64_t tiles = ((a->a_rows + mt * 128 - 1) / (mt * 128)) * a->a_batch * ((a->N + a->w_tiled - 1) / a->w_tiled);
    const int num_kb = (int)((a->K + 63) / 64);
    int sp = 1;
    if (tiles * 2 <= 148 && num_kb >= 16) {
      sp = (int)(148 / tiles);
      if (sp > num_kb / 8) sp = num_kb / 8;
      if (sp > 16) sp = 16;
      if (sp < 1) sp = 1;
    }
    if (forced_splits > 0) sp = forced_splits;
    splits = sp;
  } else if (a->rope_cos) {
    cfg = 1128;   // one 128-wide head per tile, both rotation halves in the same accumulator row
  }
  switch (cfg) {
    case 1064: return launch_gemm<1, 64>(a, splits, cm, cn, stream);
    case 1128: return launch_gemm<1, 128>(a, splits, cm, cn, stream);
    case 1256: return launch_gemm<1, 256>(a, splits, cm, cn, stream);
    case 2064: return launch_gemm<2, 64>(a, splits, cm, cn, stream);
    case 2128: return launch_gemm<2, 128>(a, splits, cm, cn, stream);
    case 2256: return launch_gemm<2, 256>(a, splits, cm, cn, stream);
    case 1208: return launch_gemm<1, 208>(a, splits, cm, cn, stream);
    case 2208: return launch_gemm<2, 208>(a, splits, cm, cn, stream);
    case 4128: return launch_gemm_2sm<128, 1, 4>(a, stream);
    case 9128: return launch_gemm_2sm<128, 1, 8>(a, stream);   // (tuning: 8 epilogue warps per CTA)
    case 9256: return launch_gemm_2sm<256, 1, 8>(a, stream);
    case 4256: return launch_gemm_2sm<256, 1, 4>(a, stream);
    case 5416: return launch_gemm_2sm<208, 2, 4>(a, stream);   // 256 x 416 pair tiles (weight-streaming regime)
    case 5512: return launch_gemm_2sm<256, 2, 8>(a, stream);   // 256 x 512 pair tiles (tensor-bound regime)
    case 6208: return launch_gemm<2, 208, 8, false>(a, splits, 1, 1, stream);   // tuning variants (uvx_debug_gemm_override)
    case 6128: return launch_gemm<2, 128, 8, false>(a, splits, 1, 1, stream);
    case 7208: return launch_gemm<2, 208, 0, true>(a, splits, 1, 1, stream);
    case 7128: return launch_gemm<2, 128, 0, true>(a, splits, 1, 1, stream);
    case 8208: return launch_gemm<2, 208, 8, true>(a, splits, 1, 1, stream);
    default: return launch_gemm<1, 64>(a, splits, cm, cn, stream);
  }
}
