// Causal grouped-query attention of the Llama prefill on the 5th-gen tensor cores (tcgen05 + TMEM), head_dim 128
// (hf:models/llama/modeling_llama.py:199-289).  Same structure as attention_tc.cu (the Whisper-encoder kernel):
//
// One CTA = 128 queries of one (sequence, query head); the K / V tiles of its KV head stream through a TMA ring of 128-key tiles.
//   warp 0      TMA producer: Q once (two 64-wide d panels of [128 x 128 B], 128B swizzle), then K_j / V_j (two panels each).
//   warp 1      MMA issuer:  S = Q K_j^T  (8 x UMMA 128x128x16 over d, fp32 in TMEM) and
//               O += P V_j  (8 x UMMA 128x128x16 over the 128 keys, A = P from TENSOR memory (TS form), B = V_j MN-major).
//   warps 2-5   softmax: one thread per query row (= TMEM lane), the 128 scores of the row in registers, causal / key-range
//               masks from indices, lazy rescaling (reference max moves only on > 2^8 growth), ex2.approx, bf16 P written to
//               TMEM with tcgen05.st; O rescaled in TMEM only when a row's reference moved.
// Key tiles entirely above the causal diagonal or outside [kv_start, kv_len) are never loaded.  Optional LSE output for the
// training backward.  Decode steps (Sq == 1) stay on the mma.sync kernel (attention.cu): one query row cannot fill a 128-row MMA.
#include "uvx_common.cuh"

namespace uvx {

static constexpr int kLQ = 128;   // queries per CTA
static constexpr int kLK = 128;   // keys per tile
static constexpr int kLD = 128;   // head dim
static constexpr int kLStages = 2;
static constexpr int kLThreads = 192;

struct AttnLlmParams {
  bf16* o;
  int64_t o_rs, o_bs;
  const int32_t* kv_len;
  const int32_t* kv_start;
  float* lse;
  int Sq, Skv, group, Hq;
  int causal;
  float scale_log2;
};

__device__ __forceinline__ uint32_t al_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void al_mbar_init(uint64_t* b, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(al_smem(b)), "r"(c));
}
__device__ __forceinline__ void al_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(al_smem(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void al_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(al_smem(b)) : "memory");
}
__device__ __forceinline__ void al_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "AL_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra AL_DONE;\n"
      "bra AL_WAIT;\n"
      "AL_DONE:\n"
      "}\n" ::"r"(al_smem(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void al_tma_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          al_smem(dst)),
      "l"(tm), "r"(al_smem(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void al_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void al_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void al_mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void al_mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d),
      "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void al_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(al_smem(bar)) : "memory");
}
__device__ __forceinline__ void al_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void al_st32(uint32_t taddr, const uint32_t* r) {
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
__device__ __forceinline__ void al_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void al_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major operand panel ([rows][64 elements = 128 B], 128B swizzle, 8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t al_desc_kmajor(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// MN-major operand (V tile [keys][d], d contiguous = the MMA's N = 128 = two 64-wide panels kLK*128 B apart (LBO); 8-key groups
// 1024 B apart (SBO)), 128B swizzle
__device__ __forceinline__ uint64_t al_desc_mnmajor(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((kLK * 128u) >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t al_idesc(int n, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

struct AlSmem {
  static constexpr int kPanel = kLQ * 128;                      // one [128 rows x 128 B] panel = 16 KB
  static constexpr int kQ = 0;                                  // 2 panels
  static constexpr int kK = kQ + 2 * kPanel;                    // stages x 2 panels
  static constexpr int kV = kK + kLStages * 2 * kPanel;         // stages x 2 panels
  static constexpr int kBar = kV + kLStages * 2 * kPanel;
  static constexpr int kTotal = kBar + 256 + 1024;
};

__global__ void __launch_bounds__(kLThreads, 1)
attn_llm_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const AttnLlmParams p) {
  pdl_trigger();
  extern __shared__ uint8_t al_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)al_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* q_full = (uint64_t*)(smem + AlSmem::kBar);
  uint64_t* kv_full = q_full + 1;            // [stages]
  uint64_t* kv_empty = kv_full + kLStages;   // [stages]
  uint64_t* s_full = kv_empty + kLStages;    // MMA -> softmax: S ready
  uint64_t* p_full = s_full + 1;             // softmax -> MMA: P written, O rescaled
  uint64_t* pv_done = p_full + 1;            // MMA -> softmax: O += P V done
  uint64_t* s_free = pv_done + 1;            // softmax -> MMA: S is in registers
  uint32_t* tmem_slot = (uint32_t*)(s_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kLQ, h = blockIdx.y, b = blockIdx.z;
  const int hk = h / p.group;

  if (threadIdx.x == 0) {
    al_mbar_init(q_full, 1);
    for (int s = 0; s < kLStages; ++s) {
      al_mbar_init(&kv_full[s], 1);
      al_mbar_init(&kv_empty[s], 1);
    }
    al_mbar_init(s_full, 1);
    al_mbar_init(p_full, 128);
    al_mbar_init(pv_done, 1);
    al_mbar_init(s_free, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(al_smem(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  al_fence_before();
  __syncthreads();
  al_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  // visible key range of this query tile: [kv_begin, kv_end)
  int kv_end = p.Skv;
  if (p.kv_len) kv_end = min(kv_end, max(p.kv_len[b], 0));
  const int shift = p.Skv - p.Sq;
  const int last_q = min(q0 + kLQ, p.Sq) - 1;
  if (p.causal) kv_end = min(kv_end, last_q + shift + 1);
  const int kv_begin = p.kv_start ? min(max(p.kv_start[b], 0), kv_end) : 0;
  const int tile0 = kv_begin / kLK;
  const int n_tiles = (kv_end + kLK - 1) / kLK;   // tiles tile0 .. n_tiles-1 are processed
  const int nt = n_tiles > tile0 ? n_tiles - tile0 : 0;
  // TMEM columns: S [0,128) fp32, O [128,256) fp32, P [256,320) packed bf16 pairs
  const uint32_t tS = tmem_base, tO = tmem_base + 128, tP = tmem_base + 256;

  if (warp == 0) {
    if (lane == 0 && nt > 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
      al_expect_tx(q_full, 2 * AlSmem::kPanel);
      al_tma_3d(smem + AlSmem::kQ, &tmQ, h * kLD, q0, b, q_full);
      al_tma_3d(smem + AlSmem::kQ + AlSmem::kPanel, &tmQ, h * kLD + 64, q0, b, q_full);
      for (int jj = 0; jj < nt; ++jj) {
        const int j = tile0 + jj;
        const int s = jj % kLStages;
        al_wait(&kv_empty[s], ((jj / kLStages) & 1) ^ 1);
        al_expect_tx(&kv_full[s], 4 * AlSmem::kPanel);
        uint8_t* sk = smem + AlSmem::kK + s * 2 * AlSmem::kPanel;
        uint8_t* sv = smem + AlSmem::kV + s * 2 * AlSmem::kPanel;
        al_tma_3d(sk, &tmK, hk * kLD, j * kLK, b, &kv_full[s]);
        al_tma_3d(sk + AlSmem::kPanel, &tmK, hk * kLD + 64, j * kLK, b, &kv_full[s]);
        al_tma_3d(sv, &tmV, hk * kLD, j * kLK, b, &kv_full[s]);
        al_tma_3d(sv + AlSmem::kPanel, &tmV, hk * kLD + 64, j * kLK, b, &kv_full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && nt > 0) {
      constexpr uint32_t idS = al_idesc(kLK, false);
      constexpr uint32_t idO = al_idesc(kLD, true);
      const uint32_t sQ = al_smem(smem + AlSmem::kQ), sK = al_smem(smem + AlSmem::kK), sV = al_smem(smem + AlSmem::kV);
      auto issue_S = [&](int stage) {
#pragma unroll
        for (int k = 0; k < kLD / 16; ++k) {
          // d = 16 k .. 16 k + 15: panel k / 4, 32 bytes per step inside the swizzle atom
          const uint64_t da = al_desc_kmajor(sQ + (k >> 2) * AlSmem::kPanel) + (uint64_t)(2 * (k & 3));
          const uint64_t db = al_desc_kmajor(sK + stage * 2 * AlSmem::kPanel + (k >> 2) * AlSmem::kPanel) + (uint64_t)(2 * (k & 3));
          al_mma(tS, da, db, idS, k > 0 ? 1u : 0u);
        }
        al_commit(s_full);
      };
      al_wait(q_full, 0);
      al_wait(&kv_full[0], 0);
      al_fence_after();
      issue_S(0);
      for (int jj = 0; jj < nt; ++jj) {
        const int s = jj % kLStages;
        if (jj + 1 < nt) {
          // the next tile's scores as soon as the softmax warps hold S(jj) in registers
          al_wait(&kv_full[(jj + 1) % kLStages], ((jj + 1) / kLStages) & 1);
          al_wait(s_free, jj & 1);
          al_fence_after();
          issue_S((jj + 1) % kLStages);
        }
        al_wait(p_full, jj & 1);   // P(jj) in tensor memory, O rescaled
        al_fence_after();
        const uint64_t dv = al_desc_mnmajor(sV + s * 2 * AlSmem::kPanel);
#pragma unroll
        for (int k = 0; k < kLK / 16; ++k) {
          // P: 16 keys = 8 packed columns per k-step; V: 16 keys = 16 rows of 128 B in each panel
          al_mma_ts(tO, tP + (uint32_t)(k * 8), dv + (uint64_t)(k * (16 * 128 >> 4)), idO, (jj > 0 || k > 0) ? 1u : 0u);
        }
        al_commit(pv_done);
        al_commit(&kv_empty[s]);
      }
    }
  } else {
    // ---- softmax warps: one thread per query row ---------------------------------------------------------
    const int qd = warp & 3;                  // TMEM lane quarter this warp may access
    const int row = qd * 32 + lane;
    const int qi = q0 + row;                  // global query index
    const uint32_t lane_sel = (uint32_t)(qd * 32) << 16;
    float m_ref = -INFINITY, l_run = 0.f;
    const float kLazy = 8.0f;
    for (int jj = 0; jj < nt; ++jj) {
      const int j = tile0 + jj;
      al_wait(s_full, jj & 1);
      al_fence_after();
      uint32_t raw[kLK];
#pragma unroll
      for (int c = 0; c < kLK / 32; ++c) al_ld32(tS + lane_sel + c * 32, raw + c * 32);
      al_wait_ld();
      al_fence_before();
      al_arrive(s_free);
      // masks from indices: key range of the sequence and the causal diagonal (CTA-uniform test whether this tile needs them)
      const bool need_mask = (j * kLK + kLK > kv_end) || (j * kLK < kv_begin) || (p.causal && j * kLK + kLK - 1 > q0 + shift);
      if (need_mask) {
#pragma unroll
        for (int i = 0; i < kLK; ++i) {
          const int key = j * kLK + i;
          bool ok = key < kv_end && key >= kv_begin;
          if (p.causal) ok = ok && (key <= qi + shift);
          if (!ok) raw[i] = 0xff800000u;  // -inf
        }
      }
      float tmax = -INFINITY;
#pragma unroll
      for (int i = 0; i < kLK; ++i) tmax = fmaxf(tmax, __uint_as_float(raw[i]));
      const bool grow = (tmax - m_ref) * p.scale_log2 > kLazy;   // also true for the first unmasked tile (m_ref = -inf)
      const bool any_grow = __any_sync(0xffffffffu, grow) != 0;
      float corr = 1.f;
      if (grow) {
        const float m_new = (tmax == -INFINITY) ? m_ref : tmax;
        corr = (m_ref == -INFINITY) ? 0.f : exp2f((m_ref - m_new) * p.scale_log2);
        m_ref = m_new;
      }
      const float mb = (m_ref == -INFINITY) ? 0.f : m_ref * p.scale_log2;
      if (jj > 0) {
        al_wait(pv_done, (jj - 1) & 1);   // P(jj-1) consumed, O updated
        al_fence_after();
      }
      float psum = 0.f;
#pragma unroll
      for (int h2i = 0; h2i < 2; ++h2i) {
        uint32_t packed[32];   // 64 keys -> 32 packed bf16 pairs -> 32 TMEM columns
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
        al_st32(tP + lane_sel + h2i * 32, packed);
      }
      l_run = l_run * corr + psum;
      if (jj > 0 && any_grow) {
#pragma unroll 1
        for (int c = 0; c < kLD / 32; ++c) {
          uint32_t o32[32];
          al_ld32(tO + lane_sel + c * 32, o32);
          al_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o32[i] = __float_as_uint(__uint_as_float(o32[i]) * corr);
          al_st32(tO + lane_sel + c * 32, o32);
        }
      }
      al_wait_st();
      al_fence_before();
      al_arrive(p_full);
    }
    // ---- epilogue: O / l -> bf16, each thread writes its own 256-byte row
    if (nt > 0) {
      al_wait(pv_done, (nt - 1) & 1);
      al_fence_after();
    }
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    if (p.lse && qi < p.Sq) {
      const float ln2 = 0.6931471805599453f;
      p.lse[((int64_t)b * p.Hq + h) * p.Sq + qi] = (m_ref == -INFINITY ? -INFINITY : m_ref * p.scale_log2 * ln2) + logf(l_run);
    }
    bf16* orow = p.o + (int64_t)b * p.o_bs + (int64_t)qi * p.o_rs + (int64_t)h * kLD;
#pragma unroll 1
    for (int c = 0; c < kLD / 32; ++c) {
      uint32_t raw[32];
      if (nt > 0) {
        al_ld32(tO + lane_sel + c * 32, raw);
        al_wait_ld();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) raw[i] = 0u;
      }
      if (qi < p.Sq) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float f8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f8[e] = __uint_as_float(raw[g * 8 + e]) * inv;
          *reinterpret_cast<bf16x8*>(orow + c * 32 + g * 8) = pack8(f8);
        }
      }
    }
  }
  al_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

typedef CUresult (*AlEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int al_map(AlEncodeFn enc, CUtensorMap* tm, const void* base, int64_t width, int64_t rows, int64_t B, int64_t rs, int64_t bs) {
  cuuint64_t dims[3] = {(cuuint64_t)width, (cuuint64_t)rows, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)rs * 2, (cuuint64_t)(B > 1 ? bs : rs) * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)kLK, 1}, es[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("uvx_attention (tcgen05): cuTensorMapEncodeTiled failed (%d)", (int)r);
    return UVX_ERR_CUDA;
  }
  return UVX_OK;
}

// Entry used by uvx_attention for head_dim 128 prefill shapes (returns UVX_OK after launching).
int launch_attn_llm_tc(const uvx_attn_args* a, cudaStream_t st) {
  static AlEncodeFn enc = nullptr;
  if (!enc) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled entry point not available");
      return UVX_ERR_CUDA;
    }
    enc = (AlEncodeFn)fp;
  }
  CUtensorMap tmQ, tmK, tmV;
  int rc = al_map(enc, &tmQ, a->q, a->Hq * kLD, a->Sq, a->B, a->q_rs, a->q_bs);
  if (!rc) rc = al_map(enc, &tmK, a->k, a->Hkv * kLD, a->Skv, a->B, a->k_rs, a->k_bs);
  if (!rc) rc = al_map(enc, &tmV, a->v, a->Hkv * kLD, a->Skv, a->B, a->v_rs, a->v_bs);
  if (rc) return rc;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_llm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AlSmem::kTotal);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(attn_llm_tc_kernel): %s", cudaGetErrorString(e));
      return UVX_ERR_CUDA;
    }
    attr = true;
  }
  AttnLlmParams p;
  p.o = (bf16*)a->o;
  p.o_rs = a->o_rs;
  p.o_bs = a->o_bs;
  p.kv_len = a->kv_len;
  p.kv_start = a->kv_start;
  p.lse = a->lse;
  p.Sq = (int)a->Sq;
  p.Skv = (int)a->Skv;
  p.group = (int)(a->Hq / a->Hkv);
  p.Hq = (int)a->Hq;
  p.causal = a->causal;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  dim3 grid((unsigned)((a->Sq + kLQ - 1) / kLQ), (unsigned)a->Hq, (unsigned)a->B);
  launch_k(attn_llm_tc_kernel, dim3(grid), dim3(kLThreads), AlSmem::kTotal, st, tmQ, tmK, tmV, p);
  return check_launch("attn_llm_tc_kernel");
}

}  // namespace uvx
