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
  int units;               // n_tiles * num_kb
  int stage_bytes, stages;
  int tiled;               // W is the pre-tiled image [n_tiles][K/64][128][64]
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
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = p.num_kb;
  const int stages = p.stages;
  const int G = (int)gridDim.x;
  // equal contiguous unit ranges (stream-K)
  const int u0 = (int)((int64_t)p.units * blockIdx.x / G);
  const int u1 = (int)((int64_t)p.units * (blockIdx.x + 1) / G);

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
    if (p.pdl && mode != 2) {
      // weights do not depend on the previous kernel: the W boxes of the first ring round go out before griddepcontrol.wait
      pre = (u1 - u0) < stages ? (u1 - u0) : stages;
      for (int i = 0; i < pre; ++i) {
        const int u = u0 + i;
        const int tile = u / KB, kb = u - tile * KB;
        mbar_expect_tx_e(&full_bar[i], (uint32_t)p.stage_bytes);
        tma_load_2d_e(smem + i * p.stage_bytes, &tmW, tiled ? 0 : kb * kBK, tiled ? u * kBM : tile * kBM, &full_bar[i]);
      }
    }
    pdl_wait();
    int s = 0;
    uint32_t ph = 0;
    int tile = u0 / KB, kb = u0 - tile * KB;
    for (int u = u0; u < u1; ++u) {
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
          tma_load_2d_e(sa, &tmW, tiled ? 0 : kb * kBK, tiled ? u * kBM : tile * kBM, &full_bar[s]);
          tma_load_2d_e(sa + kWsWBytes, &tmX, kb * kBK, 0, &full_bar[s]);
        }
      }
      if (++s == stages) { s = 0; ph ^= 1u; }
      if (++kb == KB) { kb = 0; ++tile; }
    }
  } else if (warp == 1) {
    // ---- MMA issuer (convergent warp, elect.sync issues)
    const uint32_t idesc = make_idesc(p.bnt);
    int s = 0;
    uint32_t ph = 0, seg = 0;
    for (int u = u0; u < u1; ++seg) {
      const int kb0 = u % KB;
      const int len = (KB - kb0) < (u1 - u) ? (KB - kb0) : (u1 - u);
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
      u += len;
    }
    if (dbg && lane == 0) dbg[3] = clock64();
  } else {
    // ---- epilogue
    const int q = warp & 3;              // TMEM lane quarter of this warp
    const int cpar = (warp - 2) >> 2;    // this warp takes the 32-token chunks of this parity
    const int fl = q * 32 + lane;        // feature (accumulator row) inside the tile = index of this thread in its 4-warp group
    const int nchunks = (p.bnt + 31) >> 5;
    // transpose buffer of the group: [16 tokens][128 features] fp32.  A thread owns one feature (TMEM lane) x 32 tokens; the output
    // rows want one token x consecutive features, so half a chunk at a time goes through shared memory and comes back as
    // 8-feature pieces: 16-byte loads of the residual, 16-byte stores of the output.
    float* const stg = reinterpret_cast<float*>(smem + kWsXchOff + cpar * (kWsXchBytes / 2));
    const bool swiglu = p.act == UVX_ACT_SWIGLU;
    const bool gelu = p.act == UVX_ACT_GELU;
    const bool store_out = !(p.dbg_mode & 64);
    uint32_t seg = 0;
    for (int u = u0; u < u1; ++seg) {
      const int tile = u / KB, kb0 = u - tile * KB;
      const int len = (KB - kb0) < (u1 - u) ? (KB - kb0) : (u1 - u);
      u += len;
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
        // ---- contributor: park the fp32 partial accumulator in this CTA's slot ([token][128 features]: a warp writes 128
        // contiguous bytes per token; the token rows past M hold exact zeros - their activation rows are TMA zero fill), then
        // raise the flag: bar.sync orders every thread's stores before the one release at gpu scope
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
        if (lane == 0) {
          for (int j = 1; j <= nc; ++j)
            while (ws_ld_relaxed(p.flags + blockIdx.x + j) == 0) {}
          asm volatile("fence.acq_rel.gpu;" ::: "memory");
        }
        __syncwarp();
      }
      if (stamp) dbg[4 + 3 * seg + 1] = clock64();
      const int nread = (p.dbg_mode & 16) ? 0 : nc;
      const int f = tile * kBM + fl;                 // W row of this thread
      const bool rope = p.rope_cos != nullptr && tile * kBM < p.rope_cols;
      const float bias = (p.bias && f < p.N) ? __bfloat162float(p.bias[f]) : 0.f;
      for (int c = cpar; c < nchunks; c += 2) {
        uint32_t raw[32];
        tmem_ld32(tbase + (uint32_t)(c * 32), raw);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
        const bool full = c * 32 + 32 <= p.bnt;       // (else 16 tokens: bnt is a multiple of 16)
        for (int j = 1; j <= nread; ++j) {
          const float* src = p.ws_partial + (size_t)(blockIdx.x + j) * (size_t)(kBM * p.bnt) + (size_t)c * 32 * kBM + fl;
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += src[i * kBM];
          if (full) {
#pragma unroll
            for (int i = 16; i < 32; ++i) v[i] += src[i * kBM];
          }
        }
        if (!(swiglu || rope)) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float x = v[i] * p.alpha + bias;
            if (gelu) x = gelu_fast(x);
            v[i] = x;
          }
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (half == 1 && !full) break;
#pragma unroll
          for (int i = 0; i < 16; ++i) stg[i * kBM + fl] = v[half * 16 + i];
          ws_bar_sync(2 + cpar, 128);
          const int tbase_tok = c * 32 + half * 16;
          if (swiglu) {
            // tile rows 0..63 are gate rows, 64..127 the up rows of the same 64 features: 16 tokens x 64 outputs = one 8-feature piece
            // per thread.  Rounding order of the unfused path: gate / up rounded to bf16, silu rounded to bf16, product rounded.
            const int tok = fl >> 3, f8 = (fl & 7) * 8;
            const int t = tbase_tok + tok;
            const float4* g4 = reinterpret_cast<const float4*>(stg + tok * kBM + f8);
            const float4* u4 = reinterpret_cast<const float4*>(stg + tok * kBM + 64 + f8);
            const float4 ga = g4[0], gb = g4[1], ua = u4[0], ub = u4[1];
            const float gg[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w}, uu[8] = {ua.x, ua.y, ua.z, ua.w, ub.x, ub.y, ub.z, ub.w};
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float gate = __bfloat162float(__float2bfloat16_rn(gg[e] * p.alpha));
              const float up = __bfloat162float(__float2bfloat16_rn(uu[e] * p.alpha));
              o[e] = __bfloat162float(__float2bfloat16_rn(silu_fast(gate))) * up;
            }
            const int64_t fo = (int64_t)tile * 64 + f8;
            if (t < p.M && fo < p.N / 2 && store_out) *reinterpret_cast<bf16x8*>(p.C + (int64_t)t * p.c_row_stride + fo) = pack8(o);
          } else if (rope) {
            // tile = one 128-wide head: features d and d + 64 rotate together (hf:modeling_llama.py:124-168).  The projection is
            // rounded to bf16 first, the rotation runs in fp32 on those values (rope_pair: same bits as uvx_rope).
            const int tok = fl >> 3, d8 = (fl & 7) * 8;
            const int t = tbase_tok + tok;
            if (t < p.M) {
              const float4* a4 = reinterpret_cast<const float4*>(stg + tok * kBM + d8);
              const float4* b4 = reinterpret_cast<const float4*>(stg + tok * kBM + 64 + d8);
              const float4 xa = a4[0], xb = a4[1], ya = b4[0], yb = b4[1];
              const float x1[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w}, x2[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
              const int64_t pos = p.rope_pos ? (int64_t)p.rope_pos[t] : p.rope_pos_offset + (t % p.rope_rows_per_seq);
              const float4* c4 = reinterpret_cast<const float4*>(p.rope_cos + pos * 64 + d8);
              const float4* s4 = reinterpret_cast<const float4*>(p.rope_sin + pos * 64 + d8);
              const float4 ca = c4[0], cb = c4[1], sa = s4[0], sb = s4[1];
              const float cs[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w}, sn[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
              float o1[8], o2[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float a = __bfloat162float(__float2bfloat16_rn(x1[e] * p.alpha));
                const float b = __bfloat162float(__float2bfloat16_rn(x2[e] * p.alpha));
                rope_pair(a, b, cs[e], sn[e], o1[e], o2[e]);
              }
              if (store_out) {
                bf16* crow = p.C + (int64_t)t * p.c_row_stride + (int64_t)tile * kBM;
                *reinterpret_cast<bf16x8*>(crow + d8) = pack8(o1);
                *reinterpret_cast<bf16x8*>(crow + 64 + d8) = pack8(o2);
              }
            }
          } else {
            // 16 tokens x 128 features = two 8-feature pieces per thread
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const int pc = fl + 128 * k;
              const int tok = pc >> 4, f8 = (pc & 15) * 8;
              const int t = tbase_tok + tok;
              const int64_t fg = (int64_t)tile * kBM + f8;
              if (t < p.M && fg < p.N) {
                const float4* s4 = reinterpret_cast<const float4*>(stg + tok * kBM + f8);
                const float4 a = s4[0], b = s4[1];
                float o[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
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
          ws_bar_sync(2 + cpar, 128);   // the buffer is rewritten by the next half chunk
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

static int g_ws_enable = -1;   // -1: UVX_GEMM_WS env or 1
static int g_ws_dbg = 0;
static int g_ws_grid = 0;      // tuning: force the grid (0 = #SMs)
static long long* g_ws_times = nullptr;

int gemm_ws_enabled() {
  if (g_ws_enable < 0) {
    const char* e = getenv("UVX_GEMM_WS");
    g_ws_enable = e ? atoi(e) : 1;
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
  if (!gemm_ws_enabled() || (a->flags & 2)) return false;
  if (a->a_batch != 1 || a->a_rows > 256 || a->out_dtype != UVX_DT_BF16 || a->c_row_map || a->c_row_offset != 0) return false;
  if (a->K % 8 != 0 || a->K < 64 || a->N < 128) return false;
  if (a->w_tiled && (a->w_tiled != 128 || a->K % kBK != 0)) return false;
  if (a->act == UVX_ACT_SWIGLU && (a->w_tiled != 128 || a->N % 256 != 0 || a->bias || a->R)) return false;
  if (a->rope_cos && (a->rope_cols % 128 != 0 || a->bias || a->R || a->act != UVX_ACT_NONE || a->alpha != 1.0f)) return false;
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
  p.units = p.n_tiles * p.num_kb;
  p.stage_bytes = kWsWBytes + p.bnt * kBK * 2;
  p.stages = kWsXchOff / p.stage_bytes;
  if (p.stages > kWsMaxStages) p.stages = kWsMaxStages;
  p.tiled = a->w_tiled ? 1 : 0;
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
  if (grid > p.units / 4) grid = p.units / 4;   // at least four k-blocks per CTA
  if (grid < 1) grid = 1;
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
