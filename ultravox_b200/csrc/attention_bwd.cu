// Backward of uvx_attention (data gradients only): dQ, dK, dV from dO, Q, K, V, O and the forward's log-sum-exp.
// Used by the adapter-training path for the frozen LLM's causal GQA attention.
//
//   delta_i = sum_d dO_id O_id                              (attn_delta_kernel)
//   P_ij = exp(scale * q_i.k_j - lse_i),  dP_ij = dO_i.v_j,  dS_ij = scale * P_ij (dP_ij - delta_i)
//   dQ_i = sum_j dS_ij k_j      (attn_bwd_dq_kernel:  CTA = 64 queries of one head, loops over key tiles)
//   dK_j = sum_i dS_ij q_i,  dV_j = sum_i P_ij dO_i   (attn_bwd_dkv_kernel: CTA = 64 keys of one KV head, loops over
//                                                      the query heads of the group x query tiles; no atomics)
// Same building blocks as the forward: bf16 mma.sync.m16n8k16 + ldmatrix, cp.async double buffering, masks from
// indices.  S is recomputed in both kernels (7 matmuls instead of 5) in exchange for atomic-free, deterministic sums.
#include "uvx_common.cuh"

namespace uvx {

static constexpr int kBM_ = 64;  // rows owned by a CTA (queries for dQ, keys for dK/dV)
static constexpr int kBThreads = 128;

struct AttnBwdParams {
  const bf16 *q, *k, *v, *o, *dout;
  bf16 *dq, *dk, *dv;
  int64_t q_rs, q_bs, k_rs, k_bs, v_rs, v_bs, o_rs, o_bs;        // dq/dk/dv/dout share the strides of q/k/v/o
  int64_t dq_rs, dq_bs, dk_rs, dk_bs, dv_rs, dv_bs;
  const float* lse;
  float* delta;
  const int32_t* kv_len;
  int Sq, Skv, Hq, Hkv, group, causal, block;
  float scale, scale_log2;
};

__device__ __forceinline__ void b_cp_async16(void* dst, const void* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void b_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void b_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void b_ldsm(uint32_t* r, const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void b_ldsm_t(uint32_t* r, const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a));
}
__device__ __forceinline__ void b_mma(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t b_pack(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// A-operand fragment (16 rows x 16 k) of a row-major smem tile: rows r0.., k columns k0..
template <int LD>
__device__ __forceinline__ void load_a(uint32_t* f, const bf16* tile, int r0, int k0, int lane) {
  b_ldsm(f, tile + (r0 + (lane & 15)) * LD + k0 + (lane >> 4) * 8);
}
// B fragments for two adjacent n-blocks (n0..n0+15) where B[k][n] = tile[n][k] (tile rows are the n index): K in QK^T
template <int LD>
__device__ __forceinline__ void load_b_rows(uint32_t* f, const bf16* tile, int n0, int k0, int lane) {
  b_ldsm(f, tile + (n0 + (lane >> 4) * 8 + (lane & 7)) * LD + k0 + ((lane >> 3) & 1) * 8);
}
// B fragments for two adjacent n-blocks (n0..n0+15) where B[k][n] = tile[k][n] (tile rows are the k index): V in PV
template <int LD>
__device__ __forceinline__ void load_b_cols(uint32_t* f, const bf16* tile, int k0, int n0, int lane) {
  b_ldsm_t(f, tile + (k0 + ((lane >> 3) & 1) * 8 + (lane & 7)) * LD + n0 + (lane >> 4) * 8);
}

__global__ void attn_delta_kernel(const AttnBwdParams p, int D) {
  const int lane = threadIdx.x & 31;
  const int64_t idx = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t total = (int64_t)gridDim.y * p.Hq * p.Sq;
  (void)total;
  const int b = blockIdx.y;
  if (idx >= (int64_t)p.Hq * p.Sq) return;
  const int h = (int)(idx / p.Sq), i = (int)(idx % p.Sq);
  const bf16* o = p.o + (int64_t)b * p.o_bs + (int64_t)i * p.o_rs + (int64_t)h * D;
  const bf16* d = p.dout + (int64_t)b * p.o_bs + (int64_t)i * p.o_rs + (int64_t)h * D;
  float acc = 0.f;
  for (int c = lane * 8; c < D; c += 256) {
    float a[8], g[8];
    unpack8(*reinterpret_cast<const bf16x8*>(o + c), a);
    unpack8(*reinterpret_cast<const bf16x8*>(d + c), g);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc += a[e] * g[e];
  }
  acc = warp_sum(acc);
  if (lane == 0) p.delta[((int64_t)b * p.Hq + h) * p.Sq + i] = acc;
}

template <int D, int TN>
__global__ void __launch_bounds__(kBThreads) attn_bwd_dq_kernel(const AttnBwdParams p) {
  constexpr int LD = D + 8;
  extern __shared__ __align__(16) uint8_t bsm[];
  bf16* sQ = reinterpret_cast<bf16*>(bsm);  // [64][LD]
  bf16* sdO = sQ + kBM_ * LD;               // [64][LD]
  bf16* sK = sdO + kBM_ * LD;               // 2 x [TN][LD]
  bf16* sV = sK + 2 * TN * LD;              // 2 x [TN][LD]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const int m0 = blockIdx.x * kBM_, h = blockIdx.y, b = blockIdx.z, hk = h / p.group;
  const bf16* qb = p.q + (int64_t)b * p.q_bs + (int64_t)h * D;
  const bf16* dob = p.dout + (int64_t)b * p.o_bs + (int64_t)h * D;
  const bf16* kb = p.k + (int64_t)b * p.k_bs + (int64_t)hk * D;
  const bf16* vb = p.v + (int64_t)b * p.v_bs + (int64_t)hk * D;
  int kv_end = p.Skv;
  if (p.kv_len) kv_end = min(kv_end, max(p.kv_len[b], 0));
  const int shift = p.Skv - p.Sq;
  const int last_q = min(m0 + kBM_, p.Sq) - 1;
  if (p.causal) kv_end = min(kv_end, last_q + shift + 1);
  if (p.block > 0) kv_end = min(kv_end, (last_q / p.block + 1) * p.block);
  const int n_tiles = (kv_end + TN - 1) / TN;
  constexpr int CH = D / 8;
  for (int i = tid; i < kBM_ * CH; i += kBThreads) {
    const int r = i / CH, c = i % CH;
    const bool ok = (m0 + r) < p.Sq;
    b_cp_async16(sQ + r * LD + c * 8, qb + (int64_t)(ok ? m0 + r : 0) * p.q_rs + c * 8, ok);
    b_cp_async16(sdO + r * LD + c * 8, dob + (int64_t)(ok ? m0 + r : 0) * p.o_rs + c * 8, ok);
  }
  auto load_kv = [&](int tile, int stage) {
    const int n0 = tile * TN;
    for (int i = tid; i < TN * CH; i += kBThreads) {
      const int r = i / CH, c = i % CH;
      const bool ok = (n0 + r) < p.Skv;
      const int64_t row = ok ? n0 + r : 0;
      b_cp_async16(sK + (stage * TN + r) * LD + c * 8, kb + row * p.k_rs + c * 8, ok);
      b_cp_async16(sV + (stage * TN + r) * LD + c * 8, vb + row * p.v_rs + c * 8, ok);
    }
  };
  if (n_tiles > 0) load_kv(0, 0);
  b_commit();

  const int qrow0 = m0 + warp * 16 + g;
  float lse2[2], dlt[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int qr = qrow0 + r * 8;
    const bool ok = qr < p.Sq;
    const int64_t off = ((int64_t)b * p.Hq + h) * p.Sq + (ok ? qr : 0);
    lse2[r] = ok ? p.lse[off] * 1.4426950408889634f : 0.f;
    dlt[r] = ok ? p.delta[off] : 0.f;
  }
  float dq_acc[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) dq_acc[i][0] = dq_acc[i][1] = dq_acc[i][2] = dq_acc[i][3] = 0.f;

  for (int tile = 0; tile < n_tiles; ++tile) {
    const int stage = tile & 1;
    if (tile + 1 < n_tiles) load_kv(tile + 1, stage ^ 1);
    b_commit();
    b_wait<1>();
    __syncthreads();
    const bf16* tk = sK + stage * TN * LD;
    const bf16* tv = sV + stage * TN * LD;
    float s[TN / 8][4], dp[TN / 8][4];
#pragma unroll
    for (int i = 0; i < TN / 8; ++i) {
      s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
      uint32_t qa[4], da[4];
      load_a<LD>(qa, sQ, warp * 16, kk * 16, lane);
      load_a<LD>(da, sdO, warp * 16, kk * 16, lane);
#pragma unroll
      for (int nb = 0; nb < TN / 16; ++nb) {
        uint32_t f[4];
        load_b_rows<LD>(f, tk, nb * 16, kk * 16, lane);
        b_mma(s[2 * nb], qa, f[0], f[1]);
        b_mma(s[2 * nb + 1], qa, f[2], f[3]);
        load_b_rows<LD>(f, tv, nb * 16, kk * 16, lane);
        b_mma(dp[2 * nb], da, f[0], f[1]);
        b_mma(dp[2 * nb + 1], da, f[2], f[3]);
      }
    }
    const int n0 = tile * TN;
#pragma unroll
    for (int nb = 0; nb < TN / 8; ++nb) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = n0 + nb * 8 + t4 * 2 + (e & 1);
        const int r = e >> 1;
        const int qr = qrow0 + r * 8;
        bool ok = key < kv_end && qr < p.Sq;
        if (p.causal) ok = ok && (key <= qr + shift);
        if (p.block > 0) ok = ok && (key / p.block <= qr / p.block);
        const float pv = ok ? exp2f(s[nb][e] * p.scale_log2 - lse2[r]) : 0.f;
        s[nb][e] = pv * (dp[nb][e] - dlt[r]) * p.scale;  // dS
      }
    }
#pragma unroll
    for (int ks = 0; ks < TN / 16; ++ks) {
      uint32_t a[4];
      a[0] = b_pack(s[2 * ks][0], s[2 * ks][1]);
      a[1] = b_pack(s[2 * ks][2], s[2 * ks][3]);
      a[2] = b_pack(s[2 * ks + 1][0], s[2 * ks + 1][1]);
      a[3] = b_pack(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
      for (int db = 0; db < D / 16; ++db) {
        uint32_t f[4];
        load_b_cols<LD>(f, tk, ks * 16, db * 16, lane);
        b_mma(dq_acc[2 * db], a, f[0], f[1]);
        b_mma(dq_acc[2 * db + 1], a, f[2], f[3]);
      }
    }
    __syncthreads();
  }
  b_wait<0>();
  bf16* dqb = p.dq + (int64_t)b * p.dq_bs + (int64_t)h * D;
#pragma unroll
  for (int i = 0; i < D / 8; ++i) {
    const int col = i * 8 + t4 * 2;
    if (qrow0 < p.Sq) *reinterpret_cast<uint32_t*>(dqb + (int64_t)qrow0 * p.dq_rs + col) = b_pack(dq_acc[i][0], dq_acc[i][1]);
    if (qrow0 + 8 < p.Sq)
      *reinterpret_cast<uint32_t*>(dqb + (int64_t)(qrow0 + 8) * p.dq_rs + col) = b_pack(dq_acc[i][2], dq_acc[i][3]);
  }
}

template <int D, int TN>
__global__ void __launch_bounds__(kBThreads) attn_bwd_dkv_kernel(const AttnBwdParams p) {
  constexpr int LD = D + 8;
  extern __shared__ __align__(16) uint8_t bsm[];
  bf16* sK = reinterpret_cast<bf16*>(bsm);  // [64][LD]
  bf16* sV = sK + kBM_ * LD;                // [64][LD]
  bf16* sQ = sV + kBM_ * LD;                // 2 x [TN][LD]
  bf16* sdO = sQ + 2 * TN * LD;             // 2 x [TN][LD]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const int k0 = blockIdx.x * kBM_, hk = blockIdx.y, b = blockIdx.z;
  const bf16* kb = p.k + (int64_t)b * p.k_bs + (int64_t)hk * D;
  const bf16* vb = p.v + (int64_t)b * p.v_bs + (int64_t)hk * D;
  int kv_end = p.Skv;
  if (p.kv_len) kv_end = min(kv_end, max(p.kv_len[b], 0));
  const int shift = p.Skv - p.Sq;
  constexpr int CH = D / 8;
  for (int i = tid; i < kBM_ * CH; i += kBThreads) {
    const int r = i / CH, c = i % CH;
    const bool ok = (k0 + r) < p.Skv;
    const int64_t row = ok ? k0 + r : 0;
    b_cp_async16(sK + r * LD + c * 8, kb + row * p.k_rs + c * 8, ok);
    b_cp_async16(sV + r * LD + c * 8, vb + row * p.v_rs + c * 8, ok);
  }
  // query tiles that can see this key block
  int q_begin = 0;
  if (p.causal) q_begin = max(0, k0 - shift);
  if (p.block > 0) q_begin = max(q_begin, (k0 / p.block) * p.block);
  const int t_begin = q_begin / TN;
  const int t_end = (k0 < kv_end) ? (p.Sq + TN - 1) / TN : t_begin;  // nothing to do if every key is masked
  const int tiles_per_head = max(t_end - t_begin, 0);
  const int n_iter = tiles_per_head * p.group;
  auto load_q = [&](int it, int stage) {
    const int hq = hk * p.group + it / tiles_per_head;
    const int q0 = (t_begin + it % tiles_per_head) * TN;
    const bf16* qb = p.q + (int64_t)b * p.q_bs + (int64_t)hq * D;
    const bf16* dob = p.dout + (int64_t)b * p.o_bs + (int64_t)hq * D;
    for (int i = tid; i < TN * CH; i += kBThreads) {
      const int r = i / CH, c = i % CH;
      const bool ok = (q0 + r) < p.Sq;
      const int64_t row = ok ? q0 + r : 0;
      b_cp_async16(sQ + (stage * TN + r) * LD + c * 8, qb + row * p.q_rs + c * 8, ok);
      b_cp_async16(sdO + (stage * TN + r) * LD + c * 8, dob + row * p.o_rs + c * 8, ok);
    }
  };
  if (n_iter > 0) load_q(0, 0);
  b_commit();

  const int krow0 = k0 + warp * 16 + g;
  float dk_acc[D / 8][4], dv_acc[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) {
    dk_acc[i][0] = dk_acc[i][1] = dk_acc[i][2] = dk_acc[i][3] = 0.f;
    dv_acc[i][0] = dv_acc[i][1] = dv_acc[i][2] = dv_acc[i][3] = 0.f;
  }
  for (int it = 0; it < n_iter; ++it) {
    const int stage = it & 1;
    if (it + 1 < n_iter) load_q(it + 1, stage ^ 1);
    b_commit();
    b_wait<1>();
    __syncthreads();
    const int hq = hk * p.group + it / tiles_per_head;
    const int q0 = (t_begin + it % tiles_per_head) * TN;
    const bf16* tq = sQ + stage * TN * LD;
    const bf16* tdo = sdO + stage * TN * LD;
    float st[TN / 8][4], dpt[TN / 8][4];
#pragma unroll
    for (int i = 0; i < TN / 8; ++i) {
      st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
      dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
      uint32_t ka[4], va[4];
      load_a<LD>(ka, sK, warp * 16, kk * 16, lane);
      load_a<LD>(va, sV, warp * 16, kk * 16, lane);
#pragma unroll
      for (int nb = 0; nb < TN / 16; ++nb) {
        uint32_t f[4];
        load_b_rows<LD>(f, tq, nb * 16, kk * 16, lane);   // S^T = K Q^T
        b_mma(st[2 * nb], ka, f[0], f[1]);
        b_mma(st[2 * nb + 1], ka, f[2], f[3]);
        load_b_rows<LD>(f, tdo, nb * 16, kk * 16, lane);  // dP^T = V dO^T
        b_mma(dpt[2 * nb], va, f[0], f[1]);
        b_mma(dpt[2 * nb + 1], va, f[2], f[3]);
      }
    }
    const float* lse_h = p.lse + ((int64_t)b * p.Hq + hq) * p.Sq;
    const float* dl_h = p.delta + ((int64_t)b * p.Hq + hq) * p.Sq;
#pragma unroll
    for (int nb = 0; nb < TN / 8; ++nb) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int qr = q0 + nb * 8 + t4 * 2 + c;
        const bool qok = qr < p.Sq;
        const float l2 = qok ? __ldg(lse_h + qr) * 1.4426950408889634f : 0.f;
        const float dl = qok ? __ldg(dl_h + qr) : 0.f;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int e = r * 2 + c;
          const int key = krow0 + r * 8;
          bool ok = qok && key < kv_end;
          if (p.causal) ok = ok && (key <= qr + shift);
          if (p.block > 0) ok = ok && (key / p.block <= qr / p.block);
          const float pv = ok ? exp2f(st[nb][e] * p.scale_log2 - l2) : 0.f;
          st[nb][e] = pv;                                      // P^T
          dpt[nb][e] = pv * (dpt[nb][e] - dl) * p.scale;       // dS^T
        }
      }
    }
#pragma unroll
    for (int ks = 0; ks < TN / 16; ++ks) {
      uint32_t pa[4], da[4];
      pa[0] = b_pack(st[2 * ks][0], st[2 * ks][1]);
      pa[1] = b_pack(st[2 * ks][2], st[2 * ks][3]);
      pa[2] = b_pack(st[2 * ks + 1][0], st[2 * ks + 1][1]);
      pa[3] = b_pack(st[2 * ks + 1][2], st[2 * ks + 1][3]);
      da[0] = b_pack(dpt[2 * ks][0], dpt[2 * ks][1]);
      da[1] = b_pack(dpt[2 * ks][2], dpt[2 * ks][3]);
      da[2] = b_pack(dpt[2 * ks + 1][0], dpt[2 * ks + 1][1]);
      da[3] = b_pack(dpt[2 * ks + 1][2], dpt[2 * ks + 1][3]);
#pragma unroll
      for (int db = 0; db < D / 16; ++db) {
        uint32_t f[4];
        load_b_cols<LD>(f, tdo, ks * 16, db * 16, lane);  // dV += P^T dO
        b_mma(dv_acc[2 * db], pa, f[0], f[1]);
        b_mma(dv_acc[2 * db + 1], pa, f[2], f[3]);
        load_b_cols<LD>(f, tq, ks * 16, db * 16, lane);   // dK += dS^T Q
        b_mma(dk_acc[2 * db], da, f[0], f[1]);
        b_mma(dk_acc[2 * db + 1], da, f[2], f[3]);
      }
    }
    __syncthreads();
  }
  b_wait<0>();
  bf16* dkb = p.dk + (int64_t)b * p.dk_bs + (int64_t)hk * D;
  bf16* dvb = p.dv + (int64_t)b * p.dv_bs + (int64_t)hk * D;
#pragma unroll
  for (int i = 0; i < D / 8; ++i) {
    const int col = i * 8 + t4 * 2;
    if (krow0 < p.Skv) {
      *reinterpret_cast<uint32_t*>(dkb + (int64_t)krow0 * p.dk_rs + col) = b_pack(dk_acc[i][0], dk_acc[i][1]);
      *reinterpret_cast<uint32_t*>(dvb + (int64_t)krow0 * p.dv_rs + col) = b_pack(dv_acc[i][0], dv_acc[i][1]);
    }
    if (krow0 + 8 < p.Skv) {
      *reinterpret_cast<uint32_t*>(dkb + (int64_t)(krow0 + 8) * p.dk_rs + col) = b_pack(dk_acc[i][2], dk_acc[i][3]);
      *reinterpret_cast<uint32_t*>(dvb + (int64_t)(krow0 + 8) * p.dv_rs + col) = b_pack(dv_acc[i][2], dv_acc[i][3]);
    }
  }
}

template <int D, int TN>
static int launch_bwd(const AttnBwdParams& p, int B, cudaStream_t st) {
  constexpr int LD = D + 8;
  const size_t smem = (size_t)(2 * kBM_ + 4 * TN) * LD * 2;
  static bool attr = false;
  if (!attr) {
    cudaError_t e1 = cudaFuncSetAttribute(attn_bwd_dq_kernel<D, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaError_t e2 = cudaFuncSetAttribute(attn_bwd_dkv_kernel<D, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
      set_error("cudaFuncSetAttribute(attn_bwd<%d>): %s", D, cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
      return UVX_ERR_CUDA;
    }
    attr = true;
  }
  dim3 gd((unsigned)(((int64_t)p.Hq * p.Sq + 7) / 8), (unsigned)B);
  attn_delta_kernel<<<gd, 256, 0, st>>>(p, D);
  int rc = check_launch("attn_delta_kernel");
  if (rc) return rc;
  dim3 gq((unsigned)((p.Sq + kBM_ - 1) / kBM_), (unsigned)p.Hq, (unsigned)B);
  attn_bwd_dq_kernel<D, TN><<<gq, kBThreads, smem, st>>>(p);
  rc = check_launch("attn_bwd_dq_kernel");
  if (rc) return rc;
  dim3 gk((unsigned)((p.Skv + kBM_ - 1) / kBM_), (unsigned)p.Hkv, (unsigned)B);
  attn_bwd_dkv_kernel<D, TN><<<gk, kBThreads, smem, st>>>(p);
  return check_launch("attn_bwd_dkv_kernel");
}

}  // namespace uvx

extern "C" int uvx_attention_bwd(const uvx_attn_args* a, const void* o, const void* dout, void* dq, void* dk, void* dv,
                                 int64_t dq_rs, int64_t dq_bs, int64_t dk_rs, int64_t dk_bs, int64_t dv_rs, int64_t dv_bs,
                                 float* delta_ws, uvx_stream_t stream) {
  using namespace uvx;
  UVX_REQUIRE(a && a->q && a->k && a->v && o && dout && dq && dk && dv && a->lse && delta_ws, "uvx_attention_bwd: null pointer");
  UVX_REQUIRE(!a->kv_start, "uvx_attention_bwd: left-padded batches (kv_start) are a forward-only feature");
  UVX_REQUIRE(a->D == 64 || a->D == 128, "uvx_attention_bwd: head_dim must be 64 or 128");
  UVX_REQUIRE(a->Hq % a->Hkv == 0 && a->B >= 1 && a->B < 65536 && a->Hq < 65536, "uvx_attention_bwd: bad shape");
  UVX_REQUIRE(a->q_rs % 8 == 0 && a->k_rs % 8 == 0 && a->v_rs % 8 == 0 && a->o_rs % 8 == 0 && dq_rs % 2 == 0 && dk_rs % 2 == 0 &&
                  dv_rs % 2 == 0,
              "uvx_attention_bwd: strides must keep alignment");
  AttnBwdParams p;
  p.q = (const bf16*)a->q; p.k = (const bf16*)a->k; p.v = (const bf16*)a->v; p.o = (const bf16*)o; p.dout = (const bf16*)dout;
  p.dq = (bf16*)dq; p.dk = (bf16*)dk; p.dv = (bf16*)dv;
  p.q_rs = a->q_rs; p.q_bs = a->q_bs; p.k_rs = a->k_rs; p.k_bs = a->k_bs; p.v_rs = a->v_rs; p.v_bs = a->v_bs;
  p.o_rs = a->o_rs; p.o_bs = a->o_bs;
  p.dq_rs = dq_rs; p.dq_bs = dq_bs; p.dk_rs = dk_rs; p.dk_bs = dk_bs; p.dv_rs = dv_rs; p.dv_bs = dv_bs;
  p.lse = a->lse; p.delta = delta_ws; p.kv_len = a->kv_len;
  p.Sq = (int)a->Sq; p.Skv = (int)a->Skv; p.Hq = (int)a->Hq; p.Hkv = (int)a->Hkv; p.group = (int)(a->Hq / a->Hkv);
  p.causal = a->causal; p.block = a->block; p.scale = a->scale; p.scale_log2 = a->scale * 1.4426950408889634f;
  return a->D == 64 ? launch_bwd<64, 64>(p, (int)a->B, (cudaStream_t)stream) : launch_bwd<128, 32>(p, (int)a->B, (cudaStream_t)stream);
}
