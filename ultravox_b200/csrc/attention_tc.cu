// Whisper-encoder self-attention on the 5th-gen tensor cores (tcgen05 + TMEM), head_dim 64, non-causal with
// key-length / block-causal masks from indices (hf:modeling_whisper.py:215-238; ref:ultravox_model.py:915-936).
//
// One CTA = 256 queries (two 128-row tiles) of one (clip, head); K/V stream through a TMA ring of 128-key tiles.
//   warp 0        TMA producer: Q once (2 x [128 x 64]), then K_j / V_j tiles (128B swizzle, zero-filled past the clip).
//   warp 1        single-thread MMA issuer:  S_t = Q_t K_j^T  (UMMA 128x128x16, fp32 in TMEM) and
//                 O_t += P_t V_j  (UMMA 128x64x16, A = P_t from TENSOR memory (TS form), B = V_j MN-major as TMA left it).
//   warps 2-5     softmax warpgroup of query tile 0; warps 6-9 of query tile 1: one thread per query row (= TMEM lane),
//                 so row max / sum need no shuffles.  One TMEM pass per tile (the 128 scores of a row live in registers),
//                 lazy rescaling (the reference max only moves when it grows by > 2^8), ex2.approx, bf16 P stored with tcgen05.st
//                 into its own TMEM columns (no shared-memory round trip); O rescaled in TMEM only when needed.
// The two query tiles ping-pong: while one warpgroup does its softmax the tensor core works for the other.
#include "uvx_common.cuh"

namespace uvx {

static constexpr int kQT = 128;         // rows per query tile
static constexpr int kKT = 128;         // keys per K/V tile
static constexpr int kHD = 64;          // head dim
static constexpr int kKvStages = 4;
static constexpr int kAtThreads = 320;  // 10 warps

struct AttnTcParams {
  bf16* o;
  int64_t o_rs, o_bs;
  const int32_t* kv_len;
  int Sq, Skv, q_col, k_col, v_col;  // column offsets of this head's q / k / v inside the fused row are added per head
  int block;
  float scale_log2;
};

__device__ __forceinline__ uint32_t at_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void at_mbar_init(uint64_t* b, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(at_smem(b)), "r"(c));
}
__device__ __forceinline__ void at_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(at_smem(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void at_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(at_smem(b)) : "memory");
}
__device__ __forceinline__ void at_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "AT_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra AT_DONE;\n"
      "bra AT_WAIT;\n"
      "AT_DONE:\n"
      "}\n" ::"r"(at_smem(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void at_tma_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          at_smem(dst)),
      "l"(tm), "r"(at_smem(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void at_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void at_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void at_mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
// A operand from tensor memory (TS form): P never touches shared memory
__device__ __forceinline__ void at_mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d),
      "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void at_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(at_smem(bar)) : "memory");
}
__device__ __forceinline__ void at_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void at_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::
          "r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void at_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void at_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major operand, 128B swizzle (rows 128 B apart, 8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t at_desc_kmajor(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// MN-major operand (V tile [keys][d], d contiguous = the MMA's N), 128B swizzle: 8-key groups 1024 B apart (SBO);
// LBO = distance between 64-element atoms along N (only one atom for d = 64)
__device__ __forceinline__ uint64_t at_desc_mnmajor(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((kKT * 128u) >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t at_idesc(int n, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

struct AtSmem {
  static constexpr int kQ = 0;                                   // 2 x 16 KB
  static constexpr int kK = kQ + 2 * kQT * kHD * 2;              // stages x 16 KB
  static constexpr int kV = kK + kKvStages * kKT * kHD * 2;      // stages x 16 KB
  static constexpr int kP = kV + kKvStages * kKT * kHD * 2;      // output staging: 2 query tiles x 16 KB
  static constexpr int kBar = kP + 2 * kQT * kHD * 2;
  static constexpr int kTotal = kBar + 256 + 1024;
};

__global__ void __launch_bounds__(kAtThreads, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnTcParams p) {
  pdl_trigger();
  extern __shared__ uint8_t at_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)at_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* q_full = (uint64_t*)(smem + AtSmem::kBar);
  uint64_t* kv_full = q_full + 1;             // [stages]
  uint64_t* kv_empty = kv_full + kKvStages;   // [stages]
  uint64_t* s_full = kv_empty + kKvStages;    // [2]  MMA -> softmax: S_t ready
  uint64_t* p_full = s_full + 2;              // [2]  softmax -> MMA: P_t written, S_t consumed, O_t rescaled
  uint64_t* pv_done = p_full + 2;             // [2]  MMA -> softmax: O_t += P_t V done (P_t smem and O_t reusable)
  uint64_t* s_free = pv_done + 2;             // [2]  softmax -> MMA: S_t is in registers, the TMEM buffer may be overwritten
  uint32_t* tmem_slot = (uint32_t*)(s_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * kQT), h = blockIdx.y, b = blockIdx.z;

  if (threadIdx.x == 0) {
    at_mbar_init(q_full, 1);
    for (int s = 0; s < kKvStages; ++s) {
      at_mbar_init(&kv_full[s], 1);
      at_mbar_init(&kv_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      at_mbar_init(&s_full[t], 1);
      at_mbar_init(&p_full[t], 128);
      at_mbar_init(&pv_done[t], 1);
      at_mbar_init(&s_free[t], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(at_smem(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  at_fence_before();
  __syncthreads();
  at_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  int kv_end = p.Skv;
  if (p.kv_len) kv_end = min(kv_end, max(p.kv_len[b], 0));
  if (p.block > 0) kv_end = min(kv_end, ((min(q0 + 2 * kQT, p.Sq) - 1) / p.block + 1) * p.block);
  const int n_tiles = (kv_end + kKT - 1) / kKT;
  // TMEM columns: S_0 [0,128), S_1 [128,256), O_0 [256,320), O_1 [320,384), P_0 [384,448), P_1 [448,512) (bf16 pairs)
  const uint32_t tS[2] = {tmem_base, tmem_base + 128};
  const uint32_t tO[2] = {tmem_base + 256, tmem_base + 320};
  const uint32_t tP[2] = {tmem_base + 384, tmem_base + 448};

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQKV) : "memory");
      at_expect_tx(q_full, 2 * kQT * kHD * 2);
      at_tma_3d(smem + AtSmem::kQ, &tmQKV, p.q_col + h * kHD, q0, b, q_full);
      at_tma_3d(smem + AtSmem::kQ + kQT * kHD * 2, &tmQKV, p.q_col + h * kHD, q0 + kQT, b, q_full);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % kKvStages;
        at_wait(&kv_empty[s], ((j / kKvStages) & 1) ^ 1);
        at_expect_tx(&kv_full[s], 2 * kKT * kHD * 2);
        at_tma_3d(smem + AtSmem::kK + s * kKT * kHD * 2, &tmQKV, p.k_col + h * kHD, j * kKT, b, &kv_full[s]);
        at_tma_3d(smem + AtSmem::kV + s * kKT * kHD * 2, &tmQKV, p.v_col + h * kHD, j * kKT, b, &kv_full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_tiles > 0) {
      constexpr uint32_t idS = at_idesc(kKT, false);
      constexpr uint32_t idO = at_idesc(kHD, true);
      const uint32_t sQ = at_smem(smem + AtSmem::kQ), sK = at_smem(smem + AtSmem::kK), sV = at_smem(smem + AtSmem::kV),
                     sP = 0;
      (void)sP;
      auto issue_S = [&](int t, int stage) {
        const uint64_t da = at_desc_kmajor(sQ + t * (kQT * kHD * 2));
        const uint64_t db = at_desc_kmajor(sK + stage * (kKT * kHD * 2));
#pragma unroll
        for (int k = 0; k < kHD / 16; ++k) at_mma(tS[t], da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idS, k > 0 ? 1u : 0u);
        at_commit(&s_full[t]);
      };
      at_wait(q_full, 0);
      at_wait(&kv_full[0], 0);
      at_fence_after();
      issue_S(0, 0);
      issue_S(1, 0);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % kKvStages;
        const bool more = j + 1 < n_tiles;
        if (more) {
          // next tile's scores as soon as the softmax warps have pulled S_t(j) into registers: this runs under their
          // exp / pack work instead of after it (otherwise the two warpgroups and the tensor core move in lockstep)
          at_wait(&kv_full[(j + 1) % kKvStages], ((j + 1) / kKvStages) & 1);
          for (int t = 0; t < 2; ++t) {
            at_wait(&s_free[t], j & 1);
            at_fence_after();
            issue_S(t, (j + 1) % kKvStages);
          }
        }
        for (int t = 0; t < 2; ++t) {
          at_wait(&p_full[t], j & 1);  // P_t(j) in smem, O_t rescaled
          at_fence_after();
          const uint64_t dv = at_desc_mnmajor(sV + s * (kKT * kHD * 2));
#pragma unroll
          for (int k = 0; k < kKT / 16; ++k) {
            // P_t sits in TMEM as packed bf16 pairs: 16 keys = 8 columns per k-step.  V: 16 keys = 16 rows of 128 B.
            at_mma_ts(tO[t], tP[t] + (uint32_t)(k * 8), dv + (uint64_t)(k * (16 * 128 >> 4)), idO, (j > 0 || k > 0) ? 1u : 0u);
          }
          at_commit(&pv_done[t]);
        }
        at_commit(&kv_empty[s]);  // K_j / V_j no longer needed once everything issued so far has completed
      }
    }
  } else {
    // ---- softmax warpgroups: t = query tile, one thread per query row ---------------------------------
    const int t = (warp - 2) >> 2;
    const int qd = warp & 3;                                 // TMEM lane quarter of this warp
    const int row = qd * 32 + lane;                          // row inside the query tile
    const int qi = q0 + t * kQT + row;                       // global query index
    const uint32_t lane_sel = (uint32_t)(qd * 32) << 16;
    uint8_t* sPt = smem + AtSmem::kP + t * (kQT * kHD * 2);  // output staging only (P lives in TMEM)
    // Running reference maximum m_ref (raw score units) and row sum l relative to it.  The reference only moves when the
    // tile maximum exceeds it by more than 2^8 in the exponent (lazy rescaling): P stays <= 256, well inside bf16 / fp32
    // range, and the TMEM read-modify-write of O is skipped for almost every tile.
    float m_ref = -INFINITY, l_run = 0.f;
    const float kLazy = 8.0f;
    for (int j = 0; j < n_tiles; ++j) {
      at_wait(&s_full[t], j & 1);
      at_fence_after();
      // one TMEM pass: the whole 128-key score row of this query lives in registers
      uint32_t raw[kKT];
#pragma unroll
      for (int c = 0; c < kKT / 32; ++c) at_ld32(tS[t] + lane_sel + c * 32, raw + c * 32);
      at_wait_ld();
      at_fence_before();
      at_arrive(&s_free[t]);  // the tensor core may start S_t(j+1) now
      const bool need_mask = (j * kKT + kKT > kv_end) || (p.block > 0);  // CTA-uniform
      if (need_mask) {
#pragma unroll
        for (int i = 0; i < kKT; ++i) {
          const int key = j * kKT + i;
          bool ok = key < kv_end;
          if (p.block > 0) ok = ok && (key / p.block <= qi / p.block);
          if (!ok) raw[i] = 0xff800000u;  // -inf
        }
      }
      float tmax = -INFINITY;
#pragma unroll
      for (int i = 0; i < kKT; ++i) tmax = fmaxf(tmax, __uint_as_float(raw[i]));
      // lazy reference update (warp-uniform decision so that the TMEM ld/st below stay converged)
      const bool grow = (tmax - m_ref) * p.scale_log2 > kLazy;  // also true for the first tile (m_ref = -inf)
      const bool any_grow = __any_sync(0xffffffffu, grow) != 0;
      float corr = 1.f;
      if (grow) {
        const float m_new = (tmax == -INFINITY) ? m_ref : tmax;
        corr = (m_ref == -INFINITY) ? 0.f : exp2f((m_ref - m_new) * p.scale_log2);
        m_ref = m_new;
      }
      const float mb = (m_ref == -INFINITY) ? 0.f : m_ref * p.scale_log2;
      // P_t(j-1) must have been consumed (and O_t updated) before P_t smem / O_t are touched again
      if (j > 0) {
        at_wait(&pv_done[t], (j - 1) & 1);
        at_fence_after();
      }
      float psum = 0.f;
#pragma unroll
      for (int h2i = 0; h2i < 2; ++h2i) {
        uint32_t packed[32];  // 64 keys -> 32 packed bf16 pairs -> 32 TMEM columns
#pragma unroll
        for (int i = 0; i < 64; i += 2) {
          float e0, e1;
          const float a0 = fmaf(__uint_as_float(raw[h2i * 64 + i]), p.scale_log2, -mb);
          const float a1 = fmaf(__uint_as_float(raw[h2i * 64 + i + 1]), p.scale_log2, -mb);
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
          psum += e0 + e1;
          __nv_bfloat162 h2 = __floats2bfloat162_rn(e0, e1);
          packed[i >> 1] = *reinterpret_cast<uint32_t*>(&h2);
        }
        at_st32(tP[t] + lane_sel + h2i * 32, packed);
      }
      l_run = l_run * corr + psum;
      // rescale the running output in TMEM only when some row of this warp moved its reference
      if (j > 0 && any_grow) {
#pragma unroll 1
        for (int c = 0; c < kHD / 32; ++c) {
          uint32_t o32[32];
          at_ld32(tO[t] + lane_sel + c * 32, o32);
          at_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o32[i] = __float_as_uint(__uint_as_float(o32[i]) * corr);
          at_st32(tO[t] + lane_sel + c * 32, o32);
        }
      }
      at_wait_st();  // P (and any O rescale) are in tensor memory
      at_fence_before();
      at_arrive(&p_full[t]);
    }
    // ---- epilogue: O / l -> bf16 -> global (staged through this tile's P buffer for 128-byte row stores)
    if (n_tiles > 0) {
      at_wait(&pv_done[t], (n_tiles - 1) & 1);
      at_fence_after();
    }
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    uint8_t* stg = sPt;  // [128 rows][128 B], 16-byte chunks XOR-swizzled by (row & 7)
#pragma unroll 1
    for (int c = 0; c < kHD / 32; ++c) {
      uint32_t raw[32];
      if (n_tiles > 0) {
        at_ld32(tO[t] + lane_sel + c * 32, raw);
        at_wait_ld();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) raw[i] = 0u;
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(raw[g * 8 + 2 * e]) * inv, __uint_as_float(raw[g * 8 + 2 * e + 1]) * inv);
          w[e] = *reinterpret_cast<uint32_t*>(&h2);
        }
        const int chunk = (c * 4 + g) ^ (row & 7);
        *reinterpret_cast<uint4*>(stg + row * 128 + chunk * 16) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    // the four warps of this warpgroup exchange rows through smem: named barrier per warpgroup
    asm volatile("bar.sync %0, 128;" ::"r"(t + 1) : "memory");
    bf16* ob = p.o + (int64_t)b * p.o_bs + (int64_t)h * kHD;
    const int tw = (warp - 2) & 3;
#pragma unroll 1
    for (int it = 0; it < 8; ++it) {
      const int r = tw * 32 + it * 4 + (lane >> 3);  // 4 rows per instruction, 8 lanes (128 B) per row
      const int ch = lane & 7;
      const int gq = q0 + t * kQT + r;
      if (gq < p.Sq) {
        const uint4 vv = *reinterpret_cast<const uint4*>(stg + r * 128 + ((ch ^ (r & 7)) * 16));
        *reinterpret_cast<uint4*>(ob + (int64_t)gq * p.o_rs + ch * 8) = vv;
      }
    }
  }
  at_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

typedef CUresult (*AtEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace uvx

// Fused-QKV entry: qkv [B*S, row_stride] with q at column q_col + h*64, k at k_col + h*64, v at v_col + h*64.
extern "C" int uvx_attention_enc_tc(const void* qkv, int64_t row_stride, int64_t B, int64_t S, int64_t H, int64_t q_col,
                                    int64_t k_col, int64_t v_col, void* o, int64_t o_rs, const int32_t* kv_len, int32_t block,
                                    float scale, uvx_stream_t stream) {
  using namespace uvx;
  UVX_REQUIRE(qkv && o, "uvx_attention_enc_tc: null pointer");
  UVX_REQUIRE(B >= 1 && B < 65536 && H >= 1 && H < 65536 && S >= 1, "uvx_attention_enc_tc: bad shape");
  UVX_REQUIRE(row_stride % 8 == 0 && o_rs % 8 == 0 && q_col % 8 == 0 && k_col % 8 == 0 && v_col % 8 == 0 &&
                  (uintptr_t)qkv % 16 == 0 && (uintptr_t)o % 16 == 0,
              "uvx_attention_enc_tc: alignment");
  static AtEncodeFn enc = nullptr;
  if (!enc) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled entry point not available");
      return UVX_ERR_CUDA;
    }
    enc = (AtEncodeFn)fp;
  }
  CUtensorMap tm;
  cuuint64_t dims[3] = {(cuuint64_t)row_stride, (cuuint64_t)S, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)row_stride * 2, (cuuint64_t)S * row_stride * 2};
  cuuint32_t box[3] = {kHD, kKT, 1}, es[3] = {1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(qkv), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("uvx_attention_enc_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return UVX_ERR_CUDA;
  }
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AtSmem::kTotal);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(attn_tc_kernel): %s", cudaGetErrorString(e));
      return UVX_ERR_CUDA;
    }
    attr = true;
  }
  AttnTcParams p;
  p.o = (bf16*)o;
  p.o_rs = o_rs;
  p.o_bs = S * o_rs;
  p.kv_len = kv_len;
  p.Sq = (int)S;
  p.Skv = (int)S;
  p.q_col = (int)q_col;
  p.k_col = (int)k_col;
  p.v_col = (int)v_col;
  p.block = block;
  p.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((unsigned)((S + 2 * kQT - 1) / (2 * kQT)), (unsigned)H, (unsigned)B);
  launch_k(attn_tc_kernel, dim3(grid), dim3(kAtThreads), AtSmem::kTotal, (cudaStream_t)stream, tm, p);
  return check_launch("attn_tc_kernel");
}
