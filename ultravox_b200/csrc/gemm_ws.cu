// uvx_gemm_bf16, weight-streaming form (rows <= 256, one batch: the Llama prefill at B = 1, the projector, decode batches):
//
//   C[m, n] = act(alpha * sum_k A[m,k] W[n,k] + bias[n]) + R[m,n]
//
// At M = 201 tokens the prefill GEMMs read every weight byte once and are bound by the HBM stream - unless the tensor pipe is
// slower than the stream.  With the tokens on the 128-row UMMA M dimension (gemm_tc.cu, MT = 2) the 201 rows are padded to 256
// and the MMA pipeline alone takes as long as the weight stream (profiles/r2_ws_diag_v1.txt).  Here the operands are SWAPPED:
//
//   UMMA M = 128 weight rows (output features) of one [128 x 64] W box        (A operand, shared memory, K-major, 128B swizzle)
//   UMMA N = round16(M) tokens (208 for 201): the [N x 64] activation box       (B operand, shared memory, K-major, 128B swizzle)
//   accumulator [128 features (TMEM lanes)] x [tokens (TMEM columns)], fp32, two of them (double buffered)
//
// so the padding is 208 / 201 instead of 256 / 201 and one 128 x 208 x 16 MMA (104 cycles, profiles/r2_probe_ws.txt) covers 128
// weight rows: 0.81 cycles per weight row and k-step instead of 1.00.
//
// Work decomposition is stream-K: the (feature tile, k-block) units [0, n_tiles * K/64) are cut into gridDim.x EQUAL contiguous
// ranges, one per CTA (one CTA per SM), whatever the tile count - 224 gate|up tiles or 32 down tiles load all 148 SMs evenly, and
// the weight stream of a CTA is one contiguous run of the pre-tiled image.  A range covers at most one tile tail (its first
// segment), whole tiles, and one tile head (its last segment).  The CTA holding k-block 0 of a tile OWNS it: every other CTA
// with a segment of that tile parks its fp32 partial accumulator in its workspace slot and raises its flag; the owner - whose
// segment of the tile is the LAST thing it computes, while the contributors' segments are the FIRST thing they compute - adds the
// slots in CTA order (deterministic) and runs the fused epilogue.  All CTAs are co-resident (grid <= #SMs, 1 CTA / SM), and a
// contributor never waits for anybody, so the flag wait cannot deadlock.
//
//   warp 0      TMA producer: {W box 16 KB, activation box N x 128 B} per unit into a 5-deep ring (N = 208); under programmatic
//               dependent launch the W boxes of the first ring round are requested before griddepcontrol.wait
//   warp 1      TMEM allocator (512 columns) + tcgen05.mma issuer (4 x UMMA 128 x N x 16 per unit), elect.sync convergent loops
//   warps 2-9   epilogue: TMEM lane quarter q = warp % 4 (32 features), even / odd 32-token chunks per warp of a quarter.  A thread
//               owns ONE output feature and reads 32 tokens of it per tcgen05.ld, so a warp stores 32 consecutive features of one
//               token = 64 contiguous bytes per store instruction.  Fused: alpha, bias, GELU, residual; SwiGLU (tile = 64 gate rows
//               | 64 up rows of the same features) and RoPE (tile = one 128-wide head, columns j and j + 64 rotate together) swap
//               the partner values through a 16 KB shared-memory exchange buffer.
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "tc_ptx.cuh"

namespace uvx {

static constexpr int kWsSmemTotal = 227 * 1024;
static constexpr int kWsBarOff = kWsSmemTotal - 256;
static constexpr int kWsXchBytes = 16384;               // 2 chunk parities x [32 tokens][128 features] bf16
static constexpr int kWsXchOff = kWsBarOff - kWsXchBytes;
static constexpr int kWsMaxStages = 8;
static constexpr int kWsWBytes = kBM * kBK * 2;         // 16 KB weight box
static constexpr int kWsThreads = 64 + 8 * 32;
static constexpr int kWsAccStride = 256;                // TMEM columns between the two accumulators

struct WsParams {
  int M, N, K;             // tokens, output features, reduction
  bf16* C;
  int64_t c_row_stride;
  const bf16* bias;
  const bf16* R;
  int64_t r_row_stride;
  float alpha;
  int act;                 // UVX_ACT_NONE / GELU / SWIGLU
  int bnt;                 // UMMA N = round16(M)
  int num_kb, n_tiles;
  int units;               // stream-K units: dp_r * num_kb
  int dp_w, dp_r;          // whole tiles per CTA after the stream-K share; tiles [0, dp_r) are the stream-K tiles
  int stage_bytes, stages;
  int tiled;               // W is the pre-tiled image [n_tiles][K/64][128][64]
  int pairs;               // the image is pair-permuted (UVX_TILE_ROPE_PAIRS): per 32 rows, 16 rows of columns d | the 16 rows of d + 64
  float* ws_partial;       // [gridDim.x][bnt][128] fp32 partial accumulators (token-major: a warp writes 128 contiguous bytes)
  int* flags;              // [gridDim.x], 0 between launches: slot c is complete
  const float* rope_cos;
  const float* rope_sin;
  const int32_t* rope_pos;
  int64_t rope_rows_per_seq, rope_pos_offset;
  int rope_cols;
  int pdl;
  int dbg_mode;            // tuning only, low 2 bits: 1 = loads only, 2 = MMAs only, 3 = no epilogue; + 8 = contributors skip the slot
                           // stores, 16 = owners skip the slot reads, 32 = no flag traffic at all, 64 = owners skip the output stores
  long long* dbg_times;    // tuning only: [grid][16] clock64 stamps {entry, set-up done, first stage landed, last MMA issued,
                           // 3 x {accumulators complete, flags passed, segment done}, -, exit, globaltimer entry, exit}
};

__device__ __forceinline__ void ws_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ int ws_ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ws_ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// 8 floats = 16-byte chunks ch0, ch0 + 1 (ch0 even) of a pad row, chunk index XOR-swizzled with the token parity
__device__ __forceinline__ void pad_read8(const float* row, int tok, int ch0, float* o) {
  const float4 a = *reinterpret_cast<const float4*>(row + ((ch0 ^ (tok & 1)) << 2));
  const float4 b = *reinterpret_cast<const float4*>(row + (((ch0 + 1) ^ (tok & 1)) << 2));
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
  o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
// global -> shared bulk copy (TMA engine, no tensor map), completion on an mbarrier
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void ws_st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

__global__ void __launch_bounds__(kWsThreads, 1)
gemm_ws_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const WsParams p) {
  pdl_trigger();
  const long long t_entry = p.dbg_times ? clock64() : 0;
  unsigned long long gt_entry = 0;
  if (p.dbg_times) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_entry));
  const int mode = p.dbg_mode & 3;
  long long* const dbg = p.dbg_times ? p.dbg_times + (size_t)blockIdx.x * 16 : nullptr;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full_bar = (uint64_t*)(smem + kWsBarOff);
  uint64_t* empty_bar = full_bar + kWsMaxStages;
  uint64_t* tmem_full = empty_bar + kWsMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* fix_bar = tmem_empty + 2;                  // bulk copies of the contributors' slots (owner fix-up)
  uint32_t* tmem_slot = (uint32_t*)(fix_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = p.num_kb;
  const int stages = p.stages;
  const int G = (int)gridDim.x;
  // Work of this CTA, in this order: (1) its share [u0, u1) of the stream-K units - the (tile, k-block) pairs of the first dp_r
  // tiles cut into gridDim.x equal contiguous ranges - then (2) dp_w whole tiles (tile dp_r + blockIdx.x + k * gridDim.x).  The split
  // tiles come FIRST: their partial-accumulator traffic and the owners' fix-up run in the shadow of the whole tiles' main loops, and
  // the only exposed epilogue is the plain one of the last whole tile.  (Fewer tiles than CTAs: dp_w = 0, stream-K only.)
  const int u0 = (int)((int64_t)p.units * blockIdx.x / G);
  const int u1 = (int)((int64_t)p.units * (blockIdx.x + 1) / G);
  const int dp_w = p.dp_w, dp_r = p.dp_r;
  // next segment of the sequence: (tile, first k-block, k-blocks); `u` and `k` are the cursor
#define WS_NEXT_SEGMENT(u, k, tile, kb0, len, more)                                        \
  {                                                                                       \
    more = true;                                                                          \
    if (u < u1) {                                                                         \
      tile = u / KB;                                                                      \
      kb0 = u - tile * KB;                                                                \
      len = (KB - kb0) < (u1 - u) ? (KB - kb0) : (u1 - u);                                \
      u += len;                                                                           \
    } else if (k < dp_w) {                                                                \
      tile = dp_r + (int)blockIdx.x + k * G;                                              \
      kb0 = 0;                                                                            \
      len = KB;                                                                           \
      ++k;                                                                                \
    } else {                                                                              \
      more = false;                                                                       \
    }                                                                                     \
  }

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 8);  // one arrival per epilogue warp
    }
    mbar_init(fix_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp != 0) pdl_wait();
  if (dbg && threadIdx.x == 64) {
    dbg[0] = t_entry;
    dbg[1] = clock64();
    dbg[14] = (long long)gt_entry;
  }

  if (warp == 0) {
    // ---- TMA producer (convergent warp, elect.sync issues)
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    }
    __syncwarp();
    const int tiled = p.tiled;
    int pre = 0;
    int cu = u0, ck = 0, tile = 0, kb0 = 0, len = 0;
    bool more;
    WS_NEXT_SEGMENT(cu, ck, tile, kb0, len, more);
    if (p.pdl && mode != 2 && more) {
      // weights do not depend on the previous kernel: the first W boxes of the first segment go out before griddepcontrol.wait
      pre = len < stages ? len : stages;
      for (int i = 0; i < pre; ++i) {
        mbar_expect_tx_e(&full_bar[i], (uint32_t)p.stage_bytes);
        tma_load_2d_e(smem + i * p.stage_bytes, &tmW, tiled ? 0 : (kb0 + i) * kBK, tiled ? (tile * KB + kb0 + i) * kBM : tile * kBM, &full_bar[i]);
      }
    }
    pdl_wait();
    int s = 0;
    uint32_t ph = 0;
    while (more) {
      for (int kb = kb0; kb < kb0 + len; ++kb) {
        uint8_t* sa = smem + s * p.stage_bytes;
        if (pre > 0) {
          --pre;
          tma_load_2d_e(sa + kWsWBytes, &tmX, kb * kBK, 0, &full_bar[s]);
        } else {
          mbar_wait(&empty_bar[s], ph ^ 1u);
          if (mode == 2) {   // MMAs only: hand the (never loaded) slot over
            if (lane == 0) mbar_arrive(&full_bar[s]);
            __syncwarp();
          } else {
            mbar_expect_tx_e(&full_bar[s], (uint32_t)p.stage_bytes);
            tma_load_2d_e(sa, &tmW, tiled ? 0 : kb * kBK, tiled ? (tile * KB + kb) * kBM : tile * kBM, &full_bar[s]);
            tma_load_2d_e(sa + kWsWBytes, &tmX, kb * kBK, 0, &full_bar[s]);
          }
        }
        if (++s == stages) { s = 0; ph ^= 1u; }
      }
      WS_NEXT_SEGMENT(cu, ck, tile, kb0, len, more);
    }
  } else if (warp == 1) {
    // ---- MMA issuer (convergent warp, elect.sync issues)
    const uint32_t idesc = make_idesc(p.bnt);
    int s = 0;
    uint32_t ph = 0, seg = 0;
    int cu = u0, ck = 0, tile = 0, kb0 = 0, len = 0;
    bool more;
    WS_NEXT_SEGMENT(cu, ck, tile, kb0, len, more);
    for (; more; ++seg) {
      const uint32_t acc = seg & 1u, aph = (seg >> 1) & 1u;
      mbar_wait(&tmem_empty[acc], aph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kWsAccStride;
      for (int i = 0; i < len; ++i) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (dbg && lane == 0 && seg == 0 && i == 0) dbg[2] = clock64();
        if (mode == 1) {   // loads only: free the slot at once
          if (lane == 0) mbar_arrive(&empty_bar[s]);
          __syncwarp();
        } else {
          const uint32_t sa = smem_u32(smem + s * p.stage_bytes);
          const uint64_t dw = make_smem_desc(sa);
          const uint64_t dx = make_smem_desc(sa + kWsWBytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) umma_f16_e(d_tmem, dw + (uint64_t)(2 * k), dx + (uint64_t)(2 * k), idesc, (i > 0 || k > 0) ? 1u : 0u);
          umma_commit_e(&empty_bar[s]);
        }
        if (++s == stages) { s = 0; ph ^= 1u; }
      }
      if (mode == 1) {
        if (lane == 0) mbar_arrive(&tmem_full[acc]);
        __syncwarp();
      } else {
        umma_commit_e(&tmem_full[acc]);
      }
      WS_NEXT_SEGMENT(cu, ck, tile, kb0, len, more);
    }
    (void)tile;
    if (dbg && lane == 0) dbg[3] = clock64();
  } else {
    // ---- epilogue
    const int q = warp & 3;              // TMEM lane quarter of this warp
    const int cpar = (warp - 2) >> 2;    // this warp takes the 32-token chunks of this parity
    const int fl = q * 32 + lane;        // accumulator row (tile row) of this thread
    const int nchunks = (p.bnt + 31) >> 5;
    // Per-warp transpose pad, [16 tokens][32 tile rows] fp32 (2 KB): a thread owns one tile row (TMEM lane) x 32 tokens, the output
    // wants one token x consecutive features.  Half a chunk goes through the pad and comes back as 8-feature pieces (16-byte
    // residual loads and output stores).  No cross-warp traffic: the pre-tiled images put partner rows (gate / up, RoPE pairs)
    // into the same 32-row quarter.  16-byte chunks of a pad row are XOR-swizzled with the token parity (conflict-free both ways).
    float* const pad = reinterpret_cast<float*>(smem + kWsXchOff) + (warp - 2) * 512;
    const int wofs0 = lane, wofs1 = lane ^ 4;
    const bool swiglu = p.act == UVX_ACT_SWIGLU;
    const bool gelu = p.act == UVX_ACT_GELU;
    const bool store_out = !(p.dbg_mode & 64);
    uint32_t seg = 0, fix_ph = 0;
    int cu = u0, ck = 0, tile = 0, kb0 = 0, len = 0, ntile = 0, nkb0 = 0, nlen = 0;
    bool more, nmore;
    WS_NEXT_SEGMENT(cu, ck, tile, kb0, len, more);
    for (; more; ++seg, tile = ntile, kb0 = nkb0, len = nlen, more = nmore) {
      WS_NEXT_SEGMENT(cu, ck, ntile, nkb0, nlen, nmore);    // (nmore == false: this is the CTA's last segment)
      const uint32_t acc = seg & 1u, aph = (seg >> 1) & 1u;
      mbar_wait(&tmem_full[acc], aph);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kWsAccStride;
      const bool stamp = dbg && threadIdx.x == 64 && seg < 3;
      if (stamp) dbg[4 + 3 * seg] = clock64();
      if (mode == 1 || mode == 3) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        continue;
      }
      if (kb0 != 0) {
        // ---- contributor: park the fp32 partial accumulator in this CTA's slot ([token][128 tile rows]: a warp writes 128
        // contiguous bytes per token; token rows past M hold exact zeros - their activation rows are TMA zero fill), then raise
        // the flag: bar.sync orders every thread's stores before the one release at gpu scope
        float* slot = p.ws_partial + (size_t)blockIdx.x * (size_t)(kBM * p.bnt) + fl;
        for (int c = cpar; c < nchunks; c += 2) {
          uint32_t raw[32];
          tmem_ld32(tbase + (uint32_t)(c * 32), raw);
          tmem_ld_wait();
          float* dst = slot + (size_t)c * 32 * kBM;
          if (!(p.dbg_mode & 8)) {
            if (c * 32 + 32 <= p.bnt) {
#pragma unroll
              for (int i = 0; i < 32; ++i) dst[i * kBM] = __uint_as_float(raw[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) dst[i * kBM] = __uint_as_float(raw[i]);   // (bnt is a multiple of 16)
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        if (!(p.dbg_mode & 32)) {
          ws_bar_sync(1, 256);
          if (warp == 2 && lane == 0) ws_st_release(p.flags + blockIdx.x, 1);
        }
        if (stamp) dbg[4 + 3 * seg + 2] = clock64();
        continue;
      }
      // ---- owner: add the contributors' slots (CTAs blockIdx.x + 1 .. + nc, in that order), fused epilogue, store
      int nc = 0;
      if (kb0 + len < KB) {
        const int tile_end = (tile + 1) * KB;
        for (int c2 = (int)blockIdx.x + 1; c2 < G && (int)((int64_t)p.units * c2 / G) < tile_end; ++c2) ++nc;
        if (p.dbg_mode & 32) nc = 0;
      }
      const int nread = (p.dbg_mode & 16) ? 0 : nc;
      // When the owner segment is the LAST thing this CTA computes its operand ring is idle (every stage was consumed before the
      // accumulator completed): the contributors' slots come in by bulk copies - 16 KB per (chunk, slot), as many chunks per pass
      // as the ring holds - instead of 4-byte loads whose L2 latency the eight epilogue warps cannot cover on the exposed tail.
      // With whole tiles still to come the ring is busy and the fix-up is hidden under their main loops: plain loads then.
      const bool bulk = nread > 0 && !nmore;
      const int ring_chunks = (stages * p.stage_bytes) / 16384;
      int per_pass = bulk ? (ring_chunks / nread < nchunks ? ring_chunks / nread : nchunks) : nchunks;
      if (per_pass < 1) per_pass = 1;   // (the host keeps tiles to <= 9 members; 13 chunk buffers fit the ring at 208 tokens)
      if (nc > 0) {
        if (bulk) {
          if (warp == 2 && lane == 0) {
            for (int j = 1; j <= nc; ++j)
              while (ws_ld_relaxed(p.flags + blockIdx.x + j) == 0) {}
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            asm volatile("fence.proxy.async;" ::: "memory");   // slots written through the generic proxy, bulk copies read through the async proxy
          }
        } else {
          if (lane == 0) {
            for (int j = 1; j <= nc; ++j)
              while (ws_ld_relaxed(p.flags + blockIdx.x + j) == 0) {}
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
          }
          __syncwarp();
        }
      }
      if (stamp) dbg[4 + 3 * seg + 1] = clock64();
      // tile row -> output feature of this thread's 8-lane piece k: plain rows, or the pair-permuted image (16 rows d | 16 rows d + 64)
      const int64_t tile_f0 = (int64_t)tile * kBM;
      const bool rope = p.rope_cos != nullptr && tile_f0 < p.rope_cols;
      for (int c_lo = 0; c_lo < nchunks; c_lo += per_pass) {
        const int c_hi = c_lo + per_pass < nchunks ? c_lo + per_pass : nchunks;
        if (bulk) {
          if (c_lo > 0) ws_bar_sync(1, 256);   // every warp is done with the previous pass's buffers
          if (warp == 2 && lane == 0) {
            uint32_t bytes = 0;
            for (int c = c_lo; c < c_hi; ++c) bytes += (uint32_t)((c * 32 + 32 <= p.bnt ? 32 : 16) * 512) * (uint32_t)nread;
            mbar_expect_tx(fix_bar, bytes);
            for (int c = c_lo; c < c_hi; ++c) {
              const uint32_t cb = (uint32_t)((c * 32 + 32 <= p.bnt ? 32 : 16) * 512);
              for (int j = 1; j <= nread; ++j) {
                const float* src = p.ws_partial + (size_t)(blockIdx.x + j) * (size_t)(kBM * p.bnt) + (size_t)c * 32 * kBM;
                bulk_load(smem + ((c - c_lo) * nread + (j - 1)) * 16384, src, cb, fix_bar);
              }
            }
          }
          mbar_wait(fix_bar, fix_ph);
          fix_ph ^= 1u;
        }
        for (int c = c_lo + ((cpar - c_lo) & 1); c < c_hi; c += 2) {
          uint32_t raw[32];
          tmem_ld32(tbase + (uint32_t)(c * 32), raw);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
          const bool full = c * 32 + 32 <= p.bnt;       // (else 16 tokens: bnt is a multiple of 16)
          for (int j = 0; j < nread; ++j) {
            const float* src = bulk ? reinterpret_cast<const float*>(smem + ((c - c_lo) * nread + j) * 16384) + fl
                                    : p.ws_partial + (size_t)(blockIdx.x + 1 + j) * (size_t)(kBM * p.bnt) + (size_t)c * 32 * kBM + fl;
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += src[i * kBM];
            if (full) {
#pragma unroll
              for (int i = 16; i < 32; ++i) v[i] += src[i * kBM];
            }
          }
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            if (half == 1 && !full) break;
#pragma unroll
            for (int i = 0; i < 16; ++i) pad[i * 32 + ((i & 1) ? wofs1 : wofs0)] = v[half * 16 + i];
            __syncwarp();
            const int tok0 = c * 32 + half * 16;
            if (swiglu) {
              // quarter rows 0..15 = gate rows, 16..31 = the up rows of the same 16 features: 16 tokens x 16 outputs = one 8-feature
              // piece per lane.  Rounding order of the unfused path: gate / up rounded to bf16, silu rounded to bf16, product rounded.
              const int tok = lane >> 1, h = lane & 1;
              const int t = tok0 + tok;
              float gg[8], uu[8], o[8];
              pad_read8(pad + tok * 32, tok, 2 * h, gg);
              pad_read8(pad + tok * 32, tok, 4 + 2 * h, uu);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float gate = __bfloat162float(__float2bfloat16_rn(gg[e] * p.alpha));
                const float up = __bfloat162float(__float2bfloat16_rn(uu[e] * p.alpha));
                o[e] = __bfloat162float(__float2bfloat16_rn(silu_fast(gate))) * up;
              }
              const int64_t fo = (int64_t)tile * 64 + q * 16 + h * 8;
              if (t < p.M && fo < p.N / 2 && store_out) *reinterpret_cast<bf16x8*>(p.C + (int64_t)t * p.c_row_stride + fo) = pack8(o);
            } else if (rope) {
              // pair-permuted image: quarter rows 0..15 = head columns d = 16 q + r, rows 16..31 = columns d + 64, which rotate together
              // (hf:modeling_llama.py:124-168).  The projection is rounded to bf16 first, the rotation runs in fp32 on those values
              // (rope_pair: same bits as uvx_rope).
              const int tok = lane >> 1, h = lane & 1;
              const int t = tok0 + tok;
              if (t < p.M) {
                float x1[8], x2[8], o1[8], o2[8];
                pad_read8(pad + tok * 32, tok, 2 * h, x1);
                pad_read8(pad + tok * 32, tok, 4 + 2 * h, x2);
                const int d8 = q * 16 + h * 8;
                const int64_t pos = p.rope_pos ? (int64_t)p.rope_pos[t] : p.rope_pos_offset + (t % p.rope_rows_per_seq);
                const float4* c4 = reinterpret_cast<const float4*>(p.rope_cos + pos * 64 + d8);
                const float4* s4 = reinterpret_cast<const float4*>(p.rope_sin + pos * 64 + d8);
                const float4 ca = c4[0], cb = c4[1], sa = s4[0], sb = s4[1];
                const float cs[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w}, sn[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const float a = __bfloat162float(__float2bfloat16_rn(x1[e] * p.alpha));
                  const float b = __bfloat162float(__float2bfloat16_rn(x2[e] * p.alpha));
                  rope_pair(a, b, cs[e], sn[e], o1[e], o2[e]);
                }
                if (store_out) {
                  bf16* crow = p.C + (int64_t)t * p.c_row_stride + tile_f0;
                  *reinterpret_cast<bf16x8*>(crow + d8) = pack8(o1);
                  *reinterpret_cast<bf16x8*>(crow + 64 + d8) = pack8(o2);
                }
              }
            } else {
              // 16 tokens x 32 tile rows = two 8-feature pieces per lane
#pragma unroll
              for (int k2 = 0; k2 < 2; ++k2) {
                const int idx = lane + 32 * k2;
                const int tok = idx >> 2, k = idx & 3;
                const int t = tok0 + tok;
                const int64_t fg = tile_f0 + (p.pairs ? (k < 2 ? q * 16 + k * 8 : 64 + q * 16 + (k - 2) * 8) : q * 32 + k * 8);
                if (t < p.M && fg < p.N) {
                  float o[8];
                  pad_read8(pad + tok * 32, tok, 2 * k, o);
                  if (p.bias) {
                    float bb[8];
                    unpack8(*reinterpret_cast<const bf16x8*>(p.bias + fg), bb);
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] = o[e] * p.alpha + bb[e];
                  } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] *= p.alpha;
                  }
                  if (gelu) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] = gelu_fast(o[e]);
                  }
                  if (p.R) {
                    float r[8];
                    unpack8(*reinterpret_cast<const bf16x8*>(p.R + (int64_t)t * p.r_row_stride + fg), r);
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] += r[e];
                  }
                  if (store_out) *reinterpret_cast<bf16x8*>(p.C + (int64_t)t * p.c_row_stride + fg) = pack8(o);
                }
              }
            }
            __syncwarp();   // the pad is rewritten by the next half chunk
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (nc > 0) {
        // every epilogue warp has read the slots: lower the flags for the next launch
        ws_bar_sync(1, 256);
        if (warp == 2 && lane == 0)
          for (int j = 1; j <= nc; ++j) p.flags[blockIdx.x + j] = 0;
      }
      if (stamp) dbg[4 + 3 * seg + 2] = clock64();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (dbg && threadIdx.x == 64) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    dbg[13] = clock64();
    dbg[15] = (long long)gt;
  }
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------- host side
typedef CUresult (*WsEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static WsEncodeFn ws_get_encode() {
  static WsEncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (WsEncodeFn)p;
  }
  return fn;
}

static int ws_encode_2d(CUtensorMap* tm, const void* base, uint64_t d0, uint64_t d1, uint64_t stride1_bytes, uint32_t b0, uint32_t b1,
                        CUtensorMapL2promotion promo) {
  WsEncodeFn enc = ws_get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return UVX_ERR_CUDA;
  }
  cuuint64_t gd[2] = {d0, d1};
  cuuint64_t gs[1] = {stride1_bytes};
  cuuint32_t bx[2] = {b0, b1}, es[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (weight-streaming GEMM) failed with CUresult %d (dims %llu %llu stride %llu box %u %u)", (int)r,
              (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)stride1_bytes, b0, b1);
    return UVX_ERR_CUDA;
  }
  return UVX_OK;
}

static int g_ws_enable = -1;   // -1: UVX_GEMM_WS env or 2.  0 = never, 1 = every eligible call (rows <= 256), 2 = auto: rows <= 32 only.
                               // Measured (profiles/r2_ws2_stream_k.md, r2_decode_sweep_v2.txt): at 2-8 decode streams (16-token UMMA N,
                               // tiny accumulators) this form is 5-10 % faster per step than gemm_tc_kernel + split-K reduce and needs
                               // 64 launches fewer; at the 201-token prefill it is at parity on gate|up and behind on the 32-48-tile
                               // projections (exposed fix-up + epilogue passes), so those stay on gemm_tc_kernel unless UVX_GEMM_WS=1.
static int g_ws_dbg = 0;
static int g_ws_grid = 0;      // tuning: force the grid (0 = #SMs)
static long long* g_ws_times = nullptr;

int gemm_ws_enabled() {
  if (g_ws_enable < 0) {
    const char* e = getenv("UVX_GEMM_WS");
    g_ws_enable = e ? atoi(e) : 2;
  }
  return g_ws_enable;
}

// The last kWsFlagBytes of the caller's workspace hold the slot flags; they must be zero between launches.  The kernel lowers
// every flag it raised, so the region is cleared once, the first time a workspace pointer is seen.
static constexpr size_t kWsFlagBytes = 1024;

static int ws_flags_ready(void* flags, cudaStream_t stream) {
  static std::mutex mu;
  static std::vector<void*> seen;
  std::lock_guard<std::mutex> lock(mu);
  for (void* s : seen)
    if (s == flags) return UVX_OK;
  cudaError_t e = cudaMemsetAsync(flags, 0, kWsFlagBytes, stream);
  if (e != cudaSuccess) {
    set_error("gemm_ws: cudaMemsetAsync(flags): %s", cudaGetErrorString(e));
    return UVX_ERR_CUDA;
  }
  seen.push_back(flags);
  return UVX_OK;
}

// which calls take this form (everything else stays on gemm_tc_kernel)
bool gemm_ws_eligible(const uvx_gemm_args* a) {
  const int mode = gemm_ws_enabled();
  if (!mode || (a->flags & 2)) return false;
  if (mode == 2 && a->a_rows > 32) return false;
  if (a->a_batch != 1 || a->a_rows > 256 || a->out_dtype != UVX_DT_BF16 || a->c_row_map || a->c_row_offset != 0) return false;
  if (a->K % 8 != 0 || a->K < 64 || a->N < 128) return false;
  if (a->w_tiled && (a->w_tiled != 128 || a->K % kBK != 0)) return false;
  if (a->act == UVX_ACT_SWIGLU && (a->w_tiled != 128 || a->N % 256 != 0 || a->bias || a->R)) return false;
  if (a->w_perm != 0 && (a->w_perm != 1 || a->w_tiled != 128 || a->N % 128 != 0 || a->act == UVX_ACT_SWIGLU)) return false;
  // fused RoPE rotates columns d and d + 64 of a head: they share a 32-row quarter only in the pair-permuted image
  if (a->rope_cos && (a->w_perm != 1 || a->rope_cols % 128 != 0 || a->bias || a->R || a->act != UVX_ACT_NONE || a->alpha != 1.0f)) return false;
  const int bnt = (int)((a->a_rows + 15) / 16 * 16);
  const int64_t n_tiles = (a->N + kBM - 1) / kBM, num_kb = (a->K + kBK - 1) / kBK;
  if (n_tiles * num_kb >= (1ll << 30)) return false;
  const size_t need = (size_t)148 * kBM * (size_t)bnt * 4 + kWsFlagBytes;
  if (!a->workspace || (size_t)a->workspace_bytes < need) return false;
  return true;
}

int launch_gemm_ws(const uvx_gemm_args* a, cudaStream_t stream) {
  WsParams p;
  p.M = (int)a->a_rows;
  p.N = (int)a->N;
  p.K = (int)a->K;
  p.C = (bf16*)a->C;
  p.c_row_stride = a->c_row_stride;
  p.bias = (const bf16*)a->bias;
  p.R = (const bf16*)a->R;
  p.r_row_stride = a->r_row_stride;
  p.alpha = a->alpha;
  p.act = a->act;
  p.bnt = (p.M + 15) / 16 * 16;
  p.num_kb = (p.K + kBK - 1) / kBK;
  p.n_tiles = (p.N + kBM - 1) / kBM;
  p.stage_bytes = kWsWBytes + p.bnt * kBK * 2;
  p.stages = kWsXchOff / p.stage_bytes;
  if (p.stages > kWsMaxStages) p.stages = kWsMaxStages;
  p.tiled = a->w_tiled ? 1 : 0;
  p.pairs = a->w_perm == 1 ? 1 : 0;
  p.rope_cos = a->rope_cos;
  p.rope_sin = a->rope_sin;
  p.rope_pos = a->rope_positions;
  p.rope_rows_per_seq = a->rope_rows_per_seq > 0 ? a->rope_rows_per_seq : 1;
  p.rope_pos_offset = a->rope_pos_offset;
  p.rope_cols = a->rope_cos ? (int)a->rope_cols : 0;
  p.pdl = pdl_enabled() ? 1 : 0;
  p.dbg_mode = g_ws_dbg;
  p.dbg_times = g_ws_times;
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0 || sms > 148) sms = 148;
  int grid = g_ws_grid > 0 && g_ws_grid <= sms ? g_ws_grid : sms;
  // at least four k-blocks per CTA, and at most ~9 CTAs per tile (the owner's fix-up stages every contributor's chunk in the ring)
  const int min_units = (p.num_kb + 7) / 8 > 4 ? (p.num_kb + 7) / 8 : 4;
  if (p.n_tiles >= grid) {
    // more tiles than CTAs: whole tiles round-robin, the remainder (plus one round, if the remainder alone is too small to cut
    // into `grid` useful shares) as stream-K units that every CTA works off FIRST
    p.dp_w = p.n_tiles / grid;
    p.dp_r = p.n_tiles - p.dp_w * grid;
    if (p.dp_r > 0 && p.dp_r * p.num_kb < grid * min_units) {
      p.dp_w -= 1;
      p.dp_r += grid;
    }
  } else {
    p.dp_w = 0;
    p.dp_r = p.n_tiles;
    if (grid > p.dp_r * p.num_kb / min_units) grid = p.dp_r * p.num_kb / min_units;
    if (grid < 1) grid = 1;
  }
  p.units = p.dp_r * p.num_kb;
  p.ws_partial = (float*)a->workspace;
  p.flags = (int*)((char*)a->workspace + (size_t)a->workspace_bytes - kWsFlagBytes);
  int rc = ws_flags_ready(p.flags, stream);
  if (rc) return rc;
  CUtensorMap tmX, tmW;
  rc = ws_encode_2d(&tmX, a->A, (uint64_t)a->K, (uint64_t)a->a_rows, (uint64_t)a->a_row_stride * 2, kBK, (uint32_t)p.bnt, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  if (rc) return rc;
  if (p.tiled)
    rc = ws_encode_2d(&tmW, a->W, kBK, (uint64_t)p.n_tiles * (uint64_t)p.num_kb * kBM, (uint64_t)kBK * 2, kBK, kBM, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
  else
    rc = ws_encode_2d(&tmW, a->W, (uint64_t)a->K, (uint64_t)a->N, (uint64_t)a->w_row_stride * 2, kBK, kBM, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWsSmemTotal);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(gemm_ws_kernel, smem %d): %s", kWsSmemTotal, cudaGetErrorString(e));
      return UVX_ERR_CUDA;
    }
    attr_set = true;
  }
  launch_k(gemm_ws_kernel, dim3((unsigned)grid), dim3(kWsThreads), kWsSmemTotal, stream, tmX, tmW, p);
  rc = check_launch("gemm_ws_kernel");
  if (rc) return rc;
  if (a->norm_w && a->norm_out)
    return uvx_rmsnorm(a->C, a->norm_w, a->norm_out, a->a_rows, a->N, a->c_row_stride, 0, 0, 0, a->norm_eps, (uvx_stream_t)stream);
  return UVX_OK;
}

}  // namespace uvx

// tuning hooks: enable (-1 = re-read UVX_GEMM_WS), isolation mode, forced grid
extern "C" int uvx_debug_gemm_ws(int enable, int mode, int grid) {
  uvx::g_ws_enable = enable;
  uvx::g_ws_dbg = mode;
  uvx::g_ws_grid = grid;
  return UVX_OK;
}
extern "C" int uvx_debug_gemm_ws_times(void* dev_buf) {
  uvx::g_ws_times = (long long*)dev_buf;
  return UVX_OK;
}
