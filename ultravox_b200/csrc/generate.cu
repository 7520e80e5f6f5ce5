// Decode-loop helpers (a13): everything `GenerationMixin.generate` does between two LLM steps, as kernels that read their
// step state from device memory so the whole step can sit in one CUDA graph (no host sync per token):
//   uvx_kv_write            prefill: k / v sections of the fused projection -> the static KV cache (replaces two strided copies)
//   uvx_repetition_penalty  hf:generation/logits_process.py RepetitionPenaltyLogitsProcessor (ref ultravox_pipeline.py:95-113)
//   uvx_sample              temperature / top-k multinomial sampling (ref:ultravox/inference/infer.py:319-328: do_sample when
//                           temperature > 0; hf:generation/utils.py _sample: softmax(logits / T) -> multinomial)
//   uvx_token_finish        EOS / pad bookkeeping of the finished rows, append to `sequences`, advance positions
#include "uvx_common.cuh"

namespace uvx {

__global__ void kv_write_kernel(const bf16* __restrict__ qkv, int64_t row_stride, int k_col, int v_col, int kv_width,
                                bf16* __restrict__ k_cache, bf16* __restrict__ v_cache, int64_t cache_batch_stride, int64_t B,
                                int64_t S, int64_t past) {
  pdl_trigger();
  pdl_wait();
  const int vec = kv_width / 8;
  const int64_t per = B * S * vec;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < 2 * per; idx += (int64_t)gridDim.x * blockDim.x) {
    const int which = (int)(idx / per);
    const int64_t rem = idx % per;
    const int64_t row = rem / vec;  // b * S + s
    const int j = (int)(rem % vec);
    const int64_t b = row / S, s = row % S;
    const bf16* src = qkv + row * row_stride + (which ? v_col : k_col) + j * 8;
    bf16* dst = (which ? v_cache : k_cache) + b * cache_batch_stride + (past + s) * kv_width + j * 8;
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
  }
}

// One CTA per row.  HF gathers the ORIGINAL scores of every token already in the sequence, rescales them and scatters them
// back, so a token that occurs several times is penalised once: read phase, barrier, write phase (duplicates write the same value).
__global__ void __launch_bounds__(1024) rep_penalty_kernel(float* __restrict__ logits, int64_t V, const int64_t* __restrict__ seq,
                                                          int64_t seq_stride, const int32_t* __restrict__ cur_len, float penalty,
                                                          float* __restrict__ scratch) {
  pdl_trigger();
  pdl_wait();
  const int64_t b = blockIdx.x;
  const int n = *cur_len;
  float* row = logits + b * V;
  const int64_t* sq = seq + b * seq_stride;
  float* sc = scratch + b * seq_stride;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int64_t t = sq[i];
    sc[i] = (t >= 0 && t < V) ? row[t] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int64_t t = sq[i];
    if (t >= 0 && t < V) {
      const float x = sc[i];
      row[t] = x < 0.f ? x * penalty : x / penalty;
    }
  }
}

__device__ __forceinline__ uint32_t f2key(float f) {  // order-preserving float -> uint map
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// One CTA (1024 threads) per row: max, optional exact k-th-largest threshold by 4-pass radix select on the float keys,
// sum of exp((x - max) / T) over the kept entries, then the inverse-CDF pick for the uniform u: thread t owns the contiguous
// chunk [t*c, (t+1)*c), a block scan of the chunk sums finds the chunk, a serial walk finds the index.  Deterministic for a
// given u (the host draws u from a seeded torch generator).
__global__ void __launch_bounds__(1024) sample_kernel(const float* __restrict__ logits, int64_t V, float inv_temp, int top_k,
                                                     const float* __restrict__ u_all, const int32_t* __restrict__ step_idx,
                                                     int64_t u_stride, int64_t* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[32];
  __shared__ uint32_t hist[256];
  __shared__ uint32_t sel_prefix, sel_remaining;
  __shared__ float chunk_base[32];
  __shared__ int win_thread;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int64_t b = blockIdx.x;
  const float* row = logits + b * V;
  // ---- max
  float mx = -INFINITY;
  for (int64_t i = tid; i < V; i += 1024) mx = fmaxf(mx, row[i]);
  mx = warp_max(mx);
  if (lane == 0) red[w] = mx;
  __syncthreads();
  mx = warp_max(red[lane]);
  __syncthreads();
  // ---- top-k threshold (keys >= thr_key are kept); top_k <= 0 or >= V keeps everything
  uint32_t thr_key = 0u;
  if (top_k > 0 && (int64_t)top_k < V) {
    if (tid == 0) { sel_prefix = 0u; sel_remaining = (uint32_t)top_k; }
    for (int pass = 3; pass >= 0; --pass) {
      if (tid < 256) hist[tid] = 0u;
      __syncthreads();
      const uint32_t prefix = sel_prefix;
      const uint32_t hi_mask = pass == 3 ? 0u : (0xFFFFFFFFu << ((pass + 1) * 8));
      for (int64_t i = tid; i < V; i += 1024) {
        const uint32_t k = f2key(row[i]);
        if ((k & hi_mask) == (prefix & hi_mask)) atomicAdd(&hist[(k >> (pass * 8)) & 255u], 1u);
      }
      __syncthreads();
      if (tid == 0) {
        uint32_t rem = sel_remaining;
        int d = 255;
        for (; d > 0; --d) {
          if (hist[d] >= rem) break;
          rem -= hist[d];
        }
        sel_prefix = prefix | ((uint32_t)d << (pass * 8));
        sel_remaining = rem;
      }
      __syncthreads();
    }
    thr_key = sel_prefix;
  }
  // ---- chunk sums
  const int64_t c = (V + 1023) / 1024;
  const int64_t lo = (int64_t)tid * c, hi = lo + c < V ? lo + c : V;
  float local = 0.f;
  for (int64_t i = lo; i < hi; ++i) {
    const float x = row[i];
    if (f2key(x) >= thr_key) local += __expf((x - mx) * inv_temp);
  }
  // inclusive scan over the 1024 chunk sums
  float incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) red[w] = incl;
  __syncthreads();
  if (w == 0) {
    float v = red[lane], s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    chunk_base[lane] = s - v;  // exclusive prefix of the warp totals
    if (lane == 31) red[0] = s;  // grand total (red[] is dead: every warp has read it)
  }
  if (tid == 0) win_thread = -1;
  __syncthreads();
  const float total = red[0];
  const float excl = chunk_base[w] + incl - local;
  const float u = u_all[(int64_t)(step_idx ? *step_idx : 0) * u_stride + b];
  const float target = fminf(u, 0.99999994f) * total;
  if (local > 0.f && target >= excl && target < excl + local) atomicMax(&win_thread, tid);
  __syncthreads();
  int wt = win_thread;
  if (wt < 0) {
    // rounding at a chunk edge: fall back to the last chunk with mass at or before the target
    if (local > 0.f && excl <= target) atomicMax(&win_thread, tid);
    __syncthreads();
    wt = win_thread;
  }
  if (tid == (wt < 0 ? 0 : wt)) {
    float acc = excl;
    int64_t pick = -1, last = -1;
    for (int64_t i = lo; i < hi; ++i) {
      const float x = row[i];
      if (f2key(x) < thr_key) continue;
      last = i;
      acc += __expf((x - mx) * inv_temp);
      if (acc > target) { pick = i; break; }
    }
    if (pick < 0) pick = last >= 0 ? last : 0;
    out[b] = pick;
  }
}

// One CTA: rows that already produced an EOS emit pad_id (HF semantics), the token is appended to `sequences`, rows whose token
// is an EOS become finished, every counter in `bump` advances by one, all_done[0] = (every row finished).
__global__ void token_finish_kernel(int64_t* __restrict__ tok, int32_t* __restrict__ done, const int64_t* __restrict__ eos,
                                    int n_eos, int64_t pad_id, int64_t* __restrict__ seq, int64_t seq_stride,
                                    int32_t* __restrict__ cur_len, int32_t* __restrict__ step_idx, int32_t* __restrict__ bump0,
                                    int32_t* __restrict__ bump1, int32_t* __restrict__ bump2, int32_t* __restrict__ all_done,
                                    int64_t B) {
  pdl_trigger();
  pdl_wait();
  __shared__ int any_open;
  if (threadIdx.x == 0) any_open = 0;
  __syncthreads();
  const int n = *cur_len;
  for (int64_t b = threadIdx.x; b < B; b += blockDim.x) {
    int64_t t = tok[b];
    int d = done[b];
    if (d) t = pad_id;
    tok[b] = t;
    if (seq) seq[b * seq_stride + n] = t;
    for (int e = 0; e < n_eos; ++e) d |= (t == eos[e]);
    done[b] = d;
    if (!d) atomicOr(&any_open, 1);
    if (bump0) bump0[b] += 1;
    if (bump1) bump1[b] += 1;
    if (bump2) bump2[b] += 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    *cur_len = n + 1;
    if (step_idx) *step_idx += 1;
    if (all_done) *all_done = any_open ? 0 : 1;
  }
}

}  // namespace uvx

extern "C" int uvx_kv_write(const void* qkv, int64_t row_stride, int64_t k_col, int64_t v_col, int64_t kv_width, void* k_cache,
                            void* v_cache, int64_t cache_batch_stride, int64_t B, int64_t S, int64_t past, uvx_stream_t stream) {
  using namespace uvx;
  UVX_REQUIRE(qkv && k_cache && v_cache && B >= 1 && S >= 1 && past >= 0, "uvx_kv_write: bad arguments");
  UVX_REQUIRE(kv_width % 8 == 0 && row_stride % 8 == 0 && k_col % 8 == 0 && v_col % 8 == 0 && cache_batch_stride % 8 == 0,
              "uvx_kv_write: alignment");
  const int64_t total = 2 * B * S * (kv_width / 8);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  launch_k(kv_write_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, (const bf16*)qkv, row_stride, (int)k_col,
           (int)v_col, (int)kv_width, (bf16*)k_cache, (bf16*)v_cache, cache_batch_stride, B, S, past);
  return check_launch("kv_write_kernel");
}

extern "C" int uvx_repetition_penalty(float* logits, int64_t B, int64_t V, const int64_t* seq, int64_t seq_stride,
                                      const int32_t* cur_len, float penalty, float* scratch, uvx_stream_t stream) {
  using namespace uvx;
  UVX_REQUIRE(logits && seq && cur_len && scratch && B >= 1 && V >= 1 && penalty > 0.f, "uvx_repetition_penalty: bad arguments");
  launch_k(rep_penalty_kernel, dim3((unsigned)B), dim3(1024), 0, (cudaStream_t)stream, logits, V, seq, seq_stride, cur_len, penalty,
           scratch);
  return check_launch("rep_penalty_kernel");
}

extern "C" int uvx_sample(const float* logits, int64_t B, int64_t V, float temperature, int32_t top_k, const float* u,
                          const int32_t* step_idx, int64_t u_stride, int64_t* out_idx, uvx_stream_t stream) {
  using namespace uvx;
  UVX_REQUIRE(logits && u && out_idx && B >= 1 && V >= 1 && temperature > 0.f, "uvx_sample: bad arguments");
  launch_k(sample_kernel, dim3((unsigned)B), dim3(1024), 0, (cudaStream_t)stream, logits, V, 1.0f / temperature, (int)top_k, u,
           step_idx, u_stride, out_idx);
  return check_launch("sample_kernel");
}

extern "C" int uvx_token_finish(int64_t* tok, int32_t* done, const int64_t* eos_ids, int32_t n_eos, int64_t pad_id, int64_t* seq,
                                int64_t seq_stride, int32_t* cur_len, int32_t* step_idx, int32_t* bump0, int32_t* bump1,
                                int32_t* bump2, int32_t* all_done, int64_t B, uvx_stream_t stream) {
  using namespace uvx;
  UVX_REQUIRE(tok && done && cur_len && B >= 1 && (n_eos == 0 || eos_ids), "uvx_token_finish: bad arguments");
  launch_k(token_finish_kernel, dim3(1), dim3(128), 0, (cudaStream_t)stream, tok, done, eos_ids, (int)n_eos, pad_id, seq, seq_stride,
           cur_len, step_idx, bump0, bump1, bump2, all_done, B);
  return check_launch("token_finish_kernel");
}
