/* libuvx - C ABI of the B200-native Ultravox audio->LLM hot path.
 *
 * The reference (fixie-ai/ultravox @ 648efe7f) has NO native / FFI boundary: its hot path is Python over
 * transformers/torch.  This header is therefore the boundary a maintainer would bind from
 * `ultravox/model/ultravox_model.py` / `ultravox_processing.py` with ctypes; every entry
 * cites the reference (or third-party) function whose arithmetic it replaces.  `ref:` = /root/reference,
 * `hf:` = transformers (4.51.3 pinned by the reference; same formulas in 5.5.0).
 *
 * Conventions
 *   - plain pointers + sizes only; all pointers are DEVICE pointers unless named host_*;
 *   - bf16 tensors are passed as `const void*` / `void*` (2-byte elements, row-major);
 *   - every call is asynchronous on `stream` (a cudaStream_t), allocates nothing and performs no host
 *     synchronisation.  Process-wide state: immutable per-device tables (twiddles, mel filters) built once on
 *     first use, and the uvx_debug_* tuning hooks (tile / split / pipeline-isolation overrides - process
 *     globals, for benchmarking only: leave them at their defaults in production).  The split-K workspace is
 *     the CALLER's buffer (uvx_gemm_args.workspace): one workspace per stream if GEMMs run concurrently;
 *   - returns 0 on success, a negative UVX_ERR_* otherwise; `uvx_last_error()` gives the message
 *     (thread-local).  Argument validation that the reference does in Python stays in Python.
 */
#ifndef UVX_H_
#define UVX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UVX_ABI_VERSION 1

#define UVX_OK 0
#define UVX_ERR_ARG (-1)   /* bad shape / alignment / null pointer */
#define UVX_ERR_CUDA (-2)  /* CUDA runtime / driver error (launch failure, tensor-map encode, ...) */
#define UVX_ERR_WS (-3)    /* workspace too small */

typedef void* uvx_stream_t; /* cudaStream_t */

int uvx_abi_version(void);
const char* uvx_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py's "gpu_launches") */
int64_t uvx_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * a1 / K1-K3  log-mel front end.
 * Replaces WhisperFeatureExtractor._torch_extract_fbank_features (hf:models/whisper/
 * feature_extraction_whisper.py:135-164; call site ref:ultravox/model/ultravox_processing.py:295-303):
 * reflect-pad 200, 400-point periodic-hann STFT with hop 160, last frame dropped, |.|^2, slaney mel
 * filterbank (n_mels 80 or 128), log10(max(.,1e-10)), per-clip max-8 floor, (x+4)/4.
 *   wave      [B, L] fp32, L % 160 == 0 (zero-padded by the host like the extractor does)
 *   out_f32   [B, n_mels, T] fp32, T = L/160 (the reference's `audio_values` layout) or NULL
 *   out_tm    [B, T + 2, n_mels] bf16 time-major with one zero row before and after each clip (the
 *             layout the conv stem consumes, see uvx_gemm_bf16) or NULL
 *   workspace >= uvx_logmel_workspace(B, L, n_mels) bytes                                            */
size_t uvx_logmel_workspace(int64_t B, int64_t L, int n_mels);
int uvx_logmel(const float* wave, int64_t B, int64_t L, int n_mels, float* out_f32, void* out_tm,
               void* workspace, size_t workspace_bytes, uvx_stream_t stream);

/* host-only helper: writes the dense [201, n_mels] fp32 slaney filterbank the library uses (no GPU needed) */
int uvx_debug_mel_filters(int n_mels, float* host_dense);

/* `audio_values` [N, n_mels, T] fp32 (as produced by the reference processor / collator, any padding
 * content) -> bf16 time-major [N, T + 2, n_mels] with zero guard rows.  Replaces the
 * `audio_values.to(dtype)` cast of ref:ultravox/model/ultravox_model.py:382-385.                      */
int uvx_mel_to_timemajor(const float* mel, int64_t N, int n_mels, int64_t T, void* out_tm, uvx_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Dense contraction on tcgen05 tensor cores (TMA -> smem -> tcgen05.mma -> TMEM -> epilogue).
 *   C[row(b,m), n] = act( alpha * sum_k A[b,m,k] * W[n,k] + bias[n] ) + R[b,m,n]
 * A is addressed as A + b*a_batch_stride + m*a_row_stride + k (elements), which lets the same kernel run
 *   - every nn.Linear of the path (hf:models/whisper/modeling_whisper.py:279-335,403-408;
 *     hf:models/llama/modeling_llama.py:171-184,262-288; ref:ultravox/model/ultravox_model.py:793-799),
 *   - conv1 / conv2 of the Whisper stem as implicit GEMMs over the time-major padded activation
 *     (ref:ultravox/model/ultravox_model.py:893-894): row m of clip b is the 3*C contiguous elements
 *     starting at frame stride*m of the guard-padded buffer, W is the conv weight re-laid as [C_out, 3*C_in].
 * Output rows: c_row_map ? c_row_map[b*a_rows+m] (negative = drop) : b*c_batch_rows + m + c_row_offset.
 * R (optional) is addressed R + b*r_batch_stride + m*r_row_stride + n: residual stream (whisper/llama) or the
 * positional embedding added after GELU (ref :896-899, r_batch_stride = 0).
 * Requirements: K % 8 == 0, N % 64 == 0, strides % 8 == 0, 16-byte aligned bases.
 * Split-K partial sums are reduced in a fixed order by a second kernel: results are deterministic.
 * Calls with a_batch == 1, a_rows <= 32 (decode batches; UVX_GEMM_WS=1 / uvx_debug_gemm_ws: every call with a_rows <= 256, e.g. the LLM
 * prefill at B = 1; UVX_GEMM_WS=0: never), a plain bf16 output and a workspace run the WEIGHT-STREAMING form (csrc/gemm_ws.cu): 128 weight rows on the UMMA M dimension, the tokens on the UMMA N
 * dimension (round16(rows) instead of 256 padded rows), stream-K over (feature tile, k-block) units with the partial accumulators
 * of split tiles summed in CTA order by the tile's owner (deterministic for a given device).  Workspace contract for that form:
 * at least 148*128*round16(rows)*4 + 1024 bytes, and the LAST 1024 bytes (slot flags) are the library's: it zeroes them the first
 * time it sees the pointer and leaves them zero after every launch - do not write there.                                       */
enum { UVX_ACT_NONE = 0, UVX_ACT_GELU = 1, UVX_ACT_SWIGLU = 2 };
enum { UVX_DT_BF16 = 0, UVX_DT_F32 = 1 };
enum { UVX_TILE_PLAIN = 0, UVX_TILE_ROPE_PAIRS = 1, UVX_TILE_GATE_UP_8 = 8, UVX_TILE_GATE_UP_16 = 16 };   /* uvx_tile_weight interleave */

typedef struct uvx_gemm_args {
  const void* A;            /* bf16 */
  int64_t a_batch, a_rows, K;
  int64_t a_row_stride, a_batch_stride;
  const void* W;            /* bf16 [N, K] row-major (nn.Linear layout) */
  int64_t N, w_row_stride;
  void* C;                  /* bf16 or f32 */
  int64_t c_row_stride, c_batch_rows, c_row_offset;
  const int32_t* c_row_map; /* optional */
  const void* bias;         /* bf16 [N] or NULL */
  const void* R;            /* bf16 or NULL */
  int64_t r_row_stride, r_batch_stride;
  float alpha;
  int32_t act;              /* UVX_ACT_* */
  int32_t out_dtype;        /* UVX_DT_* */
  void* workspace;          /* optional, 256-byte aligned scratch: enables split-K when the tile count cannot fill   */
  int64_t workspace_bytes;  /* the SMs (fp32 partial sums [splits][rows][N]; contents on entry do not matter).       */
  const void* norm_w;       /* optional: also emit norm_out[row,:] = norm_w * bf16(C[row,:] * rsqrt(mean(C^2) + eps)),  */
  void* norm_out;           /* the LlamaRMSNorm that follows o_proj / down_proj (hf:modeling_llama.py:53-67, 321-329),  */
  float norm_eps;           /* fused into the split-K reduction when there is one.  bf16 [rows, N], plain row order.    */
  /* ---- round 2 (all optional, zero = off) -------------------------------------------------------------------------
   * w_tiled = R (64 / 128 / 208 / 256): W points at the pre-tiled image [ceil(N/R)][K/64][R][64] of the [N, K] weight (rows past
   *   N zero) instead of the row-major matrix, so every k-block of a tile is ONE contiguous R*128-byte run of DRAM (the weight
   *   stream of the LLM prefill is HBM-bound; see uvx_tile_weight).  The kernel then uses R-wide tiles.
   * act = UVX_ACT_SWIGLU (needs w_tiled = 208 and the image built with interleave = 8: 8 gate rows alternate with the 8 up rows of
   *   the same features - or, for the weight-streaming form, w_tiled = 128 with interleave = 16: 16 gate rows | 16 up rows per
   *   32-row quarter): the epilogue writes C[row, f] = silu(gate_f) * up_f for the N/2 features
   *   (LlamaMLP act_fn(gate_proj(x)) * up_proj(x), hf:modeling_llama.py:183) - the [rows, N] intermediate never reaches HBM.
   * rope_cos/rope_sin [max_pos, 64] fp32 (+ rope_positions / rope_rows_per_seq / rope_pos_offset as in uvx_rope): tiles whose
   *   first column is < rope_cols (= (Hq + Hkv) * 128) are rotated in the epilogue (hf:modeling_llama.py:124-168, head_dim 128).  */
  int32_t w_tiled;
  int32_t rope_cols;
  const float* rope_cos;
  const float* rope_sin;
  const int32_t* rope_positions;
  int64_t rope_rows_per_seq, rope_pos_offset;
  int32_t flags;            /* bit 0: run the round-1 kernel variant (no TMA-store epilogue / weight-stream producer) for this call;
                             * bit 1: never take the weight-streaming form for this call                                           */
  int32_t w_perm;           /* row order inside the tiles of a w_tiled = 128 image: 0 = plain, 1 = UVX_TILE_ROPE_PAIRS (needed by the fused
                             * RoPE of the weight-streaming form)                                                                 */
} uvx_gemm_args;

int uvx_gemm_bf16(const uvx_gemm_args* args, uvx_stream_t stream);
/* tuning hook: force tile config MT*1000+BN (0 = heuristic) and split-K count (0 = heuristic) for later calls */
int uvx_debug_gemm_override(int cfg, int splits);
/* tuning hook, bit mask: 1 = bf16 outputs in plain row order without residual are written by TMA stores (default), 2 = also with a
 * residual, 4 = fp32 split-K partials too; 0 = transposing epilogue everywhere */
int uvx_debug_gemm_tma_store(int on);
/* tuning hook: device buffer [grid][8] int64 filled with per-CTA phase timestamps by the diagnostic twin kernels (NULL = off) */
int uvx_debug_gemm_times(void* dev_buf);
/* tuning hook: cap the shared-memory ring depth of uvx_gemm_bf16 (0 = as deep as fits) */
int uvx_debug_gemm_stages(int n);
/* tuning hook: L2 prefetch distance of the weight stream in k-blocks (0 = off; < 0 = default) */
int uvx_debug_gemm_pf(int pf);
/* tuning hook of the weight-streaming form: enable (0 = never, 1 = rows <= 256, 2 = rows <= 32; -1 = UVX_GEMM_WS env, default 2), isolation mode (0 = off, 1 = loads
 * only, 2 = MMAs only, 3 = no epilogue; + 8 / 16 / 32 / 64 skip slot stores / slot reads / flag traffic / output stores), forced grid size (0 = one CTA per SM) */
int uvx_debug_gemm_ws(int enable, int mode, int grid);
/* tuning hook: device buffer [grid][16] int64 of per-CTA phase timestamps of the weight-streaming form (NULL = off) */
int uvx_debug_gemm_ws_times(void* dev_buf);
/* W [N, K] bf16 row-major (row stride w_row_stride) -> the pre-tiled image uvx_gemm_args.w_tiled = R reads:
 * out[t][kb][r][0:64] = W[row(t, r), kb*64 : kb*64+64], zero where row >= N.  interleave = 0: row(t, r) = t*R + r.
 * interleave = 8 (fused gate|up, N = 2*F, R % 16 == 0): r = 16*g + j -> gate feature t*R/2 + 8g + j (j < 8) = W row of that
 * feature, or the up row F + t*R/2 + 8g + (j-8) (j >= 8).  interleave = 16 (R = 128, weight-streaming form): r = 32q + j ->
 * gate feature 64t + 16q + j (j < 16) or the up row of feature 64t + 16q + (j - 16).  interleave = 1 (UVX_TILE_ROPE_PAIRS, R = 128 =
 * head_dim): r = 32q + j -> row 128t + 16q + j (j < 16) or its rotation partner 128t + 64 + 16q + (j - 16).  out: ceil(N/R) * (K/64) * R * 64 bf16 elements.                 */
int uvx_tile_weight(const void* W, int64_t N, int64_t K, int64_t w_row_stride, int32_t R, int32_t interleave, void* out,
                    uvx_stream_t stream);
/* tuning hook: force the thread-block cluster shape cm x cn (row tiles x column tiles sharing operand loads; 0 = heuristic) */
int uvx_debug_gemm_cluster(int cm, int cn);
/* tuning hook: 1 = 1-SM kernel without its MMAs (load pipeline alone), 2 = without its TMA loads; outputs are garbage */
int uvx_debug_gemm_mode(int mode);

/* ---------------------------------------------------------------------------------------------
 * Row-wise normalisations (fp32 statistics, bf16 in/out).
 * uvx_layernorm: nn.LayerNorm(eps 1e-5) of the Whisper encoder (hf:modeling_whisper.py:393,403;
 *                ref:ultravox_model.py:980).
 * uvx_rmsnorm:   LlamaRMSNorm (hf:models/llama/modeling_llama.py:53-67) and the projector RMSNorm
 *                (ref:ultravox_model.py:733-736): y = w * bf16(x * rsqrt(mean(x^2) + eps)).
 *                `valid_per_group`/`group_rows` implement StackAudioFrames (ref :722-730) without a copy:
 *                rows are grouped `group_rows` per clip; row t of a clip only has
 *                clamp(valid_elems - t*cols, 0, cols) real elements, the rest read as zero.
 *                Pass group_rows = 0 for a plain matrix.                                               */
int uvx_layernorm(const void* x, const void* w, const void* b, void* y, int64_t rows, int64_t cols,
                  int64_t x_row_stride, float eps, uvx_stream_t stream);
int uvx_rmsnorm(const void* x, const void* w, void* y, int64_t rows, int64_t cols, int64_t x_row_stride,
                int64_t group_rows, int64_t group_stride, int64_t valid_elems, float eps, uvx_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Fused softmax(QK^T*scale + mask)V, flash style, fp32 softmax, bf16 in/out.
 *   Whisper encoder self-attention (hf:modeling_whisper.py:215-238,305-349): non-causal, keys >= kv_len[b]
 *   masked (ref:ultravox_model.py:915-926), optional block-causal streaming mask (ref :834-863,928-936);
 *   Llama attention (hf:modeling_llama.py:199-289): causal, grouped-query.
 * q[b, i, h, :] = q + b*q_bs + i*q_rs + h*D  (same for k, v with h / (Hq/Hkv), and o).                  */
typedef struct uvx_attn_args {
  const void *q, *k, *v;
  void* o;
  int64_t B, Hq, Hkv, Sq, Skv, D;   /* D in {64, 128} */
  int64_t q_rs, q_bs, k_rs, k_bs, v_rs, v_bs, o_rs, o_bs;
  const int32_t* kv_len;            /* [B] or NULL (= Skv) */
  int32_t causal;                   /* query i sees keys j <= i + (Skv - Sq) */
  int32_t block;                    /* >0: block-causal, query i sees keys j with j/block <= i/block */
  float scale;
  float* lse;                       /* optional [B, Hq, Sq] fp32: log-sum-exp of the scaled scores (training) */
  const int32_t* kv_start;          /* optional [B]: keys j < kv_start[b] are masked - left-padded batches              */
                                    /* (ref collator ultravox_processing.py:53-63; hf masking_utils padding mask)       */
} uvx_attn_args;
int uvx_attention(const uvx_attn_args* args, uvx_stream_t stream);
/* head_dim 128 with a tile of queries per head (Llama prefill / training) runs on tcgen05 tensor cores with TMEM accumulators
 * (attention_llm_tc.cu); single-token decode steps and head_dim 64 on mma.sync.  Tuning hook: 0 forces mma.sync everywhere. */
int uvx_debug_attn_tc(int on);
/* Whisper-encoder specialisation on tcgen05 tensor cores (head_dim 64, Sq == Skv, non-causal + key-length / block-causal
 * masks): qkv is the fused projection [B*S, row_stride] with head h's q / k / v at columns q_col + 64h, k_col + 64h,
 * v_col + 64h; output o[b*S + i, 64h .. 64h+63] (row stride o_rs).  Same math as uvx_attention.                      */
int uvx_attention_enc_tc(const void* qkv, int64_t row_stride, int64_t B, int64_t S, int64_t H, int64_t q_col, int64_t k_col,
                         int64_t v_col, void* o, int64_t o_rs, const int32_t* kv_len, int32_t block, float scale,
                         uvx_stream_t stream);

/* RoPE on the q and k sections of a fused [rows, (Hq + 2*Hkv) * D] projection, in place
 * (hf:modeling_llama.py:124-168; cos/sin tables [max_pos, D/2] fp32 built by the host exactly like
 * LlamaRotaryEmbedding incl. llama3 scaling, hf:modeling_rope_utils.py:550-626).
 * position of row r = positions ? positions[r] : pos_offset + (r % rows_per_seq).                      */
int uvx_rope(void* qkv, int64_t rows, int64_t row_stride, int Hq, int Hkv, int D, const float* cos_tab,
             const float* sin_tab, const int32_t* positions, int64_t rows_per_seq, int64_t pos_offset,
             uvx_stream_t stream);

/* out[r, j] = silu(gate) * lin where for x[r, 0:2H]:
 *   gate_first = 0: lin = x[:, j], gate = x[:, H + j]   (ref SwiGLU, ultravox_model.py:739-742)
 *   gate_first = 1: gate = x[:, j], lin = x[:, H + j]   (llama MLP act(gate)*up, hf:modeling_llama.py:183) */
int uvx_swiglu(const void* x, void* out, int64_t rows, int64_t H, int64_t x_row_stride, int gate_first,
               uvx_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * a11 + a5 (K14 + K15): token-embedding gather fused with the audio splice, sync-free and bit-exact.
 * uvx_splice_plan builds src[b*S + s] = (row of audio_embeds) or -1 from the reference's index vectors
 *   (ref:ultravox_model.py:259-275,390-394: chunks in batch order, later chunks overwrite earlier ones);
 *   audio_embeds rows are a*tok_stride + j, j < audio_token_len[a].
 * uvx_embed_splice writes out[b, s, :] = src >= 0 ? audio_embeds[src] : embed_tokens[input_ids[b, s]].  */
int uvx_splice_plan(const int64_t* start_idx, const int32_t* tok_len, const int64_t* audio_batch_size,
                    int64_t n_chunks, int64_t B, int64_t S, int64_t tok_stride, int32_t* src,
                    uvx_stream_t stream);
int uvx_embed_splice(const int64_t* input_ids, const void* embed_tokens, int64_t vocab, const void* audio_embeds,
                     const int32_t* src, int64_t rows, int64_t d, void* out, uvx_stream_t stream);

/* Last-position LM head: logits[b, v] = sum_k h[b, k] * W[v, k] (bf16 x bf16 -> f32), HBM-streaming GEMV
 * (hf:modeling_llama.py:485-491 with logits_to_keep = 1), and greedy argmax (first maximal index, like
 * torch.argmax; ref:ultravox/inference/infer.py:319-328 greedy path).                                  */
int uvx_lm_head(const void* h, int64_t B, int64_t h_row_stride, const void* W, int64_t V, int64_t d,
                float* logits, uvx_stream_t stream);
int uvx_argmax(const float* logits, int64_t B, int64_t V, int64_t* out_idx, uvx_stream_t stream);

/* a13 decode step (ref:ultravox/model/ultravox_model.py:398-426 -> HF greedy generate): with one token per stream every
 * linear layer is a weight-streaming matrix-vector product, y[b, n] = sum_k x[b, k] W[n, k] (+ R[b, n]), 1 <= B <= 8.   */
int uvx_gemv_bf16(const void* x, int64_t B, int64_t x_row_stride, const void* W, int64_t w_row_stride, int64_t N, int64_t K,
                  const void* R, int64_t r_row_stride, void* out, int64_t o_row_stride, int out_f32, uvx_stream_t stream);
/* copy this step's k / v sections of the fused projection into the static KV cache at positions[b] (device index, so
 * the decode step is capturable in a CUDA graph); cache layout [B, S_max, kv_width]                                    */
int uvx_kv_append(const void* qkv, int64_t row_stride, int64_t k_col, int64_t v_col, int64_t kv_width, void* k_cache,
                  void* v_cache, int64_t cache_batch_stride, const int32_t* positions, int64_t B, uvx_stream_t stream);
/* Decode-step fusions (one CUDA graph per step, ultravox_b200/engine.py): uvx_gemv_bf16 with a fused prologue on the B activation rows -
 * norm_w != NULL: LlamaRMSNorm (bit-identical to uvx_rmsnorm followed by uvx_gemv_bf16), swiglu = 1: the row is [gate | up] of width
 * 2K and the kernel consumes act_fn(gate) * up (bit-identical to uvx_swiglu(gate_first = 1) + uvx_gemv_bf16) - and RoPE on q / k +
 * KV-cache append in one launch (same bits as uvx_rope + uvx_kv_append).                                                       */
int uvx_gemv_fused_bf16(const void* x, int64_t B, int64_t x_row_stride, const void* W, int64_t w_row_stride, int64_t N, int64_t K,
                        const void* R, int64_t r_row_stride, void* out, int64_t o_row_stride, int out_f32, const void* norm_w,
                        float norm_eps, int swiglu, uvx_stream_t stream);
int uvx_rope_kv_append(void* qkv, int64_t B, int64_t row_stride, int Hq, int Hkv, int D, const float* cos_tab, const float* sin_tab,
                       const int32_t* rope_positions, void* k_cache, void* v_cache, int64_t cache_batch_stride,
                       const int32_t* positions, uvx_stream_t stream);
/* a[i] += delta (and b[i] += delta when b != NULL): advances the device-side positions / lengths after each step     */
int uvx_add_i32(int32_t* a, int32_t* b, int64_t n, int32_t delta, uvx_stream_t stream);
/* prefill counterpart of uvx_kv_append: rows b*S + s of the fused projection -> cache[b, past + s] (k and v sections),
 * what DynamicCache.update does for the prompt (hf:cache_utils.py; call site hf:modeling_llama.py:262-273)           */
int uvx_kv_write(const void* qkv, int64_t row_stride, int64_t k_col, int64_t v_col, int64_t kv_width, void* k_cache,
                 void* v_cache, int64_t cache_batch_stride, int64_t B, int64_t S, int64_t past, uvx_stream_t stream);

/* The rest of one `GenerationMixin` step (ref:ultravox/model/ultravox_model.py:398-426 -> hf:generation/utils.py _sample),
 * with every per-step scalar read from DEVICE memory so that a whole decode step replays from one CUDA graph:
 * uvx_repetition_penalty  scores of tokens already in seq[b, 0:cur_len[0]] are divided (positive) / multiplied (negative)
 *                         by `penalty`, each distinct token once (hf:generation/logits_process.py
 *                         RepetitionPenaltyLogitsProcessor; the reference pipeline enables 1.1, ref ultravox_pipeline.py:95-113);
 *                         scratch: [B, seq_stride] fp32.
 * uvx_sample              out[b] ~ softmax(logits[b] / temperature) restricted to the top_k largest logits (top_k <= 0: all),
 *                         by inverse CDF on the uniform u[step_idx[0] * u_stride + b] (step_idx NULL = 0): the do_sample branch of
 *                         ref:ultravox/inference/infer.py:319-328.  Deterministic given u.
 * uvx_token_finish        tok[b] = done[b] ? pad_id : tok[b]; seq[b, cur_len[0]] = tok[b]; done[b] |= tok[b] in eos_ids;
 *                         cur_len[0]++, step_idx[0]++ (if given), bump{0,1,2}[b]++ (if given: cache slot / visible keys / RoPE
 *                         position); all_done[0] = all(done).                                                                 */
int uvx_repetition_penalty(float* logits, int64_t B, int64_t V, const int64_t* seq, int64_t seq_stride, const int32_t* cur_len,
                           float penalty, float* scratch, uvx_stream_t stream);
int uvx_sample(const float* logits, int64_t B, int64_t V, float temperature, int32_t top_k, const float* u, const int32_t* step_idx,
               int64_t u_stride, int64_t* out_idx, uvx_stream_t stream);
int uvx_token_finish(int64_t* tok, int32_t* done, const int64_t* eos_ids, int32_t n_eos, int64_t pad_id, int64_t* seq,
                     int64_t seq_stride, int32_t* cur_len, int32_t* step_idx, int32_t* bump0, int32_t* bump1, int32_t* bump2,
                     int32_t* all_done, int64_t B, uvx_stream_t stream);

/* Shifted causal-LM cross entropy (hf:loss/loss_utils.py:28-67; called through LlamaForCausalLM.forward(labels=)
 * from ref:ultravox/model/ultravox_model.py:328-334).  logits [B*S, V] fp32 (row_stride elements), labels [B, S]
 * un-shifted (the shift and the ignore_index padding happen inside).  row_loss/row_lse [B*S] are kept for the
 * backward; out_loss2[0] = mean loss over non-ignored positions, out_loss2[1] = their count.  shift = 1 is the
 * HF convention above; shift = 0 takes labels[row] as is (rows pre-gathered by the caller).               */
int uvx_ce_loss(const float* logits, int64_t row_stride, const int64_t* labels, int64_t B, int64_t S, int64_t V,
                int64_t ignore_index, int shift, float* row_loss, float* row_lse, float* out_loss2, uvx_stream_t stream);

/* a15: KL distillation loss of ref:ultravox/model/ultravox_model.py:202-257 (the reference's default training loss):
 * F.kl_div(log_softmax(student / T), softmax(teacher / T), "batchmean") on the prediction rows + eot_loss_weight x the
 * same on the EOT rows.  Rows are pre-gathered ([R, V] fp32 each) and carry a weight row_w[r] (1/#pred, plus
 * eot_weight/#eot on EOT rows): out_loss[0] = sum_r row_w[r] * KL_r.  uvx_kl_bwd gives d/d(student logits) in bf16.  */
int uvx_kl_loss(const float* student, const float* teacher, int64_t row_stride, int64_t R, int64_t V, float temperature,
                const float* row_w, float* row_kl, float* lse_s, float* lse_t, float* out_loss, uvx_stream_t stream);
int uvx_kl_bwd(const float* student, const float* teacher, int64_t row_stride, int64_t R, int64_t V, float temperature,
               const float* row_w, const float* lse_s, const float* lse_t, float grad_scale, void* dlogits,
               uvx_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * a14: adapter backward (encoder + LLM frozen: ref apply_lora r=0, ultravox_model.py:690-709).  These are the
 * pieces torch.autograd runs for the reference between `loss.backward()` and the projector weights; dense
 * contractions reuse uvx_gemm_bf16 (dgrad against pre-transposed weights, wgrad against transposed activations). */
int uvx_rope_bwd(void* dqkv, int64_t rows, int64_t row_stride, int Hq, int Hkv, int D, const float* cos_tab,
                 const float* sin_tab, const int32_t* positions, int64_t rows_per_seq, int64_t pos_offset,
                 uvx_stream_t stream);
/* dQ/dK/dV of uvx_attention; `a` is the forward's argument struct (with a->lse filled by the forward), o the
 * forward output, dout its gradient (same strides as o); delta_ws is [B, Hq, Sq] fp32 scratch.               */
int uvx_attention_bwd(const uvx_attn_args* a, const void* o, const void* dout, void* dq, void* dk, void* dv,
                      int64_t dq_rs, int64_t dq_bs, int64_t dk_rs, int64_t dk_bs, int64_t dv_rs, int64_t dv_bs,
                      float* delta_ws, uvx_stream_t stream);
int uvx_transpose_bf16(const void* in, int64_t rows, int64_t cols, int64_t in_row_stride, void* out,
                       int64_t out_row_stride, uvx_stream_t stream);
/* dx = dres + d(rmsnorm)/dx . dy (dx may be NULL), dw[cols] += sum_rows dy * xhat (fp32, dw may be NULL);
 * group_* as in uvx_rmsnorm (stack mode: no dx).                                                              */
int uvx_rmsnorm_bwd(const void* dy, const void* x, const void* w, const void* dres, void* dx, float* dw,
                    int64_t rows, int64_t cols, int64_t x_row_stride, int64_t group_rows, int64_t group_stride,
                    int64_t valid_elems, float eps, uvx_stream_t stream);
/* LayerNorm data gradient (+ optional residual-branch gradient dres): dx = LN'(x)^T dy + dres; the norm's weight / bias are frozen
 * (encoder backward of LoRA training, hf:modeling_whisper.py:403-440).  cols %% 8 == 0, <= 2048.                                */
int uvx_layernorm_bwd(const void* dy, const void* x, const void* w, const void* dres, void* dx, int64_t rows, int64_t cols,
                      float eps, uvx_stream_t stream);
/* erf-form GELU on bf16 (training keeps fc1's pre-activation) and its derivative: dx = dy * (Phi(x) + x phi(x)); n %% 8 == 0 */
int uvx_gelu(const void* x, void* y, int64_t n, uvx_stream_t stream);
int uvx_gelu_bwd(const void* x, const void* dy, void* dx, int64_t n, uvx_stream_t stream);
int uvx_swiglu_bwd(const void* x, const void* dout, void* dx, int64_t rows, int64_t H, int64_t x_row_stride,
                   int gate_first, uvx_stream_t stream);
/* dlogits (bf16 [B*S, V]) of uvx_ce_loss: (softmax - onehot) * grad_scale / count on valid rows, 0 elsewhere. */
int uvx_ce_bwd(const float* logits, int64_t row_stride, const int64_t* labels, int64_t B, int64_t S, int64_t V,
               int64_t ignore_index, int shift, const float* row_lse, const float* loss2, float grad_scale,
               void* dlogits, uvx_stream_t stream);
/* out[i, :] = idx[i] >= 0 ? src[idx[i], :] : 0  (gradient of the splice: rows of d(inputs_embeds) -> d(audio_embeds)) */
int uvx_gather_rows(const void* src, const int32_t* idx, int64_t rows, int64_t d, void* out, uvx_stream_t stream);
/* inverse of the splice table: inv[audio_row] = position (b*S+s) it was spliced to, or -1                     */
int uvx_splice_inverse(const int32_t* src, int64_t n_pos, int32_t* inv, int64_t n_audio_rows, uvx_stream_t stream);
/* AdamW on bf16 parameters with fp32 gradient / moments (torch.optim.AdamW semantics, ref meta_config.yaml:27) */
int uvx_adamw(void* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
              float weight_decay, int64_t step, float grad_scale, uvx_stream_t stream);
int uvx_cast_f32_bf16(const float* in, void* out, int64_t n, uvx_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* UVX_H_ */
