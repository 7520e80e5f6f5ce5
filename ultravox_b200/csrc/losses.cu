// Shifted causal-LM cross entropy over fp32 logits (hf:loss/loss_utils.py:28-67: labels padded with
// ignore_index and shifted left by one, mean over the non-ignored positions, fp32 math).
// One CTA per (b, s) row: online max / sum-exp with 128-bit loads, then one deterministic single-CTA mean.
#include "uvx_common.cuh"

namespace uvx {

__global__ void __launch_bounds__(512) ce_rows_kernel(const float* __restrict__ logits, int64_t row_stride,
                                                     const int64_t* __restrict__ labels, int64_t S, int64_t V,
                                                     int64_t ignore_index, int shift, float* __restrict__ row_loss,
                                                     float* __restrict__ row_lse) {
  __shared__ float red[32];
  const int64_t row = blockIdx.x;
  const int64_t b = row / S, s = row % S;
  const int64_t label = shift ? ((s + 1 < S) ? labels[b * S + s + 1] : ignore_index) : labels[row];
  const float* x = logits + row * row_stride;
  // pass 1: max
  float mx = -INFINITY;
  for (int64_t i = threadIdx.x * 4; i < V; i += blockDim.x * 4) {
    if (i + 3 < V) {
      const float4 v = *reinterpret_cast<const float4*>(x + i);
      mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    } else {
      for (int64_t j = i; j < V; ++j) mx = fmaxf(mx, x[j]);
    }
  }
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) mx = fmaxf(mx, red[i]);
  // pass 2: sum exp (row is L2-resident after pass 1)
  float sum = 0.f;
  for (int64_t i = threadIdx.x * 4; i < V; i += blockDim.x * 4) {
    if (i + 3 < V) {
      const float4 v = *reinterpret_cast<const float4*>(x + i);
      sum += expf(v.x - mx) + expf(v.y - mx) + expf(v.z - mx) + expf(v.w - mx);
    } else {
      for (int64_t j = i; j < V; ++j) sum += expf(x[j] - mx);
    }
  }
  sum = block_sum(sum, red);
  if (threadIdx.x == 0) {
    const float lse = mx + logf(sum);
    row_lse[row] = lse;
    const bool valid = label != ignore_index && label >= 0 && label < V;
    row_loss[row] = valid ? lse - x[label] : 0.f;
  }
}

__global__ void __launch_bounds__(1024) ce_mean_kernel(const float* __restrict__ row_loss, const int64_t* __restrict__ labels,
                                                      int64_t B, int64_t S, int64_t V, int64_t ignore_index, int shift,
                                                      float* __restrict__ out) {
  __shared__ float red[32];
  float sum = 0.f, cnt = 0.f;
  for (int64_t r = threadIdx.x; r < B * S; r += blockDim.x) {
    const int64_t s = r % S;
    const int64_t label = shift ? ((s + 1 < S) ? labels[r + 1] : ignore_index) : labels[r];
    if (label != ignore_index && label >= 0 && label < V) {
      sum += row_loss[r];
      cnt += 1.f;
    }
  }
  sum = block_sum(sum, red);
  cnt = block_sum(cnt, red);
  if (threadIdx.x == 0) {
    out[0] = sum / cnt;  // 0/0 = NaN like torch when every label is ignored
    out[1] = cnt;
  }
}

// ------------------------------------------------------------------------------------------- KL distillation
// ref:ultravox/model/ultravox_model.py:202-257: kl_div(log_softmax(student / T), softmax(teacher / T), "batchmean") on the
// prediction rows, plus eot_loss_weight x the same on the EOT rows.  Rows arrive pre-gathered with a per-row weight
// (1 / #pred rows, + eot_weight / #eot rows on EOT rows), so loss = sum_r w_r * KL_r.  One CTA per row.
__global__ void __launch_bounds__(512) kl_rows_kernel(const float* __restrict__ s, const float* __restrict__ t, int64_t row_stride,
                                                     int64_t V, float inv_T, const float* __restrict__ row_w,
                                                     float* __restrict__ row_kl, float* __restrict__ lse_s,
                                                     float* __restrict__ lse_t) {
  __shared__ float red[32];
  const int64_t row = blockIdx.x;
  const float* xs = s + row * row_stride;
  const float* xt = t + row * row_stride;
  float ms = -INFINITY, mt = -INFINITY;
  for (int64_t i = threadIdx.x; i < V; i += blockDim.x) {
    ms = fmaxf(ms, xs[i] * inv_T);
    mt = fmaxf(mt, xt[i] * inv_T);
  }
  ms = warp_max(ms);
  mt = warp_max(mt);
  __shared__ float red2[32];
  if ((threadIdx.x & 31) == 0) {
    red[threadIdx.x >> 5] = ms;
    red2[threadIdx.x >> 5] = mt;
  }
  __syncthreads();
  ms = red[0];
  mt = red2[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) {
    ms = fmaxf(ms, red[i]);
    mt = fmaxf(mt, red2[i]);
  }
  float ss = 0.f, st = 0.f;
  for (int64_t i = threadIdx.x; i < V; i += blockDim.x) {
    ss += expf(xs[i] * inv_T - ms);
    st += expf(xt[i] * inv_T - mt);
  }
  ss = block_sum(ss, red);
  st = block_sum(st, red);
  const float ls = ms + logf(ss), lt = mt + logf(st);
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < V; i += blockDim.x) {
    const float lpt = xt[i] * inv_T - lt;
    acc += expf(lpt) * (lpt - (xs[i] * inv_T - ls));
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) {
    row_kl[row] = acc * row_w[row];
    lse_s[row] = ls;
    lse_t[row] = lt;
  }
}

__global__ void __launch_bounds__(1024) sum_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[0] = s;
}

// d(loss)/d(student logits) = w_r * (softmax(s/T) - softmax(t/T)) / T
__global__ void __launch_bounds__(512) kl_bwd_kernel(const float* __restrict__ s, const float* __restrict__ t, int64_t row_stride,
                                                    int64_t V, float inv_T, const float* __restrict__ row_w,
                                                    const float* __restrict__ lse_s, const float* __restrict__ lse_t,
                                                    float grad_scale, bf16* __restrict__ d) {
  const int64_t row = blockIdx.x;
  const float* xs = s + row * row_stride;
  const float* xt = t + row * row_stride;
  const float ls = lse_s[row], lt = lse_t[row], w = row_w[row] * inv_T * grad_scale;
  bf16* dr = d + row * V;
  for (int64_t i = threadIdx.x; i < V; i += blockDim.x)
    dr[i] = __float2bfloat16_rn(w * (expf(xs[i] * inv_T - ls) - expf(xt[i] * inv_T - lt)));
}

}  // namespace uvx

extern "C" int uvx_kl_loss(const float* student, const float* teacher, int64_t row_stride, int64_t R, int64_t V, float temperature,
                           const float* row_w, float* row_kl, float* lse_s, float* lse_t, float* out_loss, uvx_stream_t stream) {
  using namespace uvx;
  UVX_REQUIRE(student && teacher && row_w && row_kl && lse_s && lse_t && out_loss, "uvx_kl_loss: null pointer");
  UVX_REQUIRE(R >= 1 && V >= 1 && temperature > 0.f, "uvx_kl_loss: bad shape");
  kl_rows_kernel<<<(unsigned)R, 512, 0, (cudaStream_t)stream>>>(student, teacher, row_stride, V, 1.f / temperature, row_w, row_kl,
                                                               lse_s, lse_t);
  int rc = check_launch("kl_rows_kernel");
  if (rc) return rc;
  sum_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(row_kl, R, out_loss);
  return check_launch("sum_kernel");
}

extern "C" int uvx_kl_bwd(const float* student, const float* teacher, int64_t row_stride, int64_t R, int64_t V, float temperature,
                          const float* row_w, const float* lse_s, const float* lse_t, float grad_scale, void* dlogits,
                          uvx_stream_t stream) {
  using namespace uvx;
  UVX_REQUIRE(student && teacher && row_w && lse_s && lse_t && dlogits, "uvx_kl_bwd: null pointer");
  kl_bwd_kernel<<<(unsigned)R, 512, 0, (cudaStream_t)stream>>>(student, teacher, row_stride, V, 1.f / temperature, row_w, lse_s, lse_t,
                                                              grad_scale, (bf16*)dlogits);
  return check_launch("kl_bwd_kernel");
}

extern "C" int uvx_ce_loss(const float* logits, int64_t row_stride, const int64_t* labels, int64_t B, int64_t S, int64_t V,
                           int64_t ignore_index, int shift, float* row_loss, float* row_lse, float* out_loss2,
                           uvx_stream_t stream) {
  using namespace uvx;
  UVX_REQUIRE(logits && labels && row_loss && row_lse && out_loss2, "uvx_ce_loss: null pointer");
  UVX_REQUIRE(B >= 1 && S >= 1 && V >= 1 && row_stride % 4 == 0 && (uintptr_t)logits % 16 == 0, "uvx_ce_loss: bad shape");
  ce_rows_kernel<<<(unsigned)(B * S), 512, 0, (cudaStream_t)stream>>>(logits, row_stride, labels, S, V, ignore_index, shift,
                                                                     row_loss, row_lse);
  int rc = check_launch("ce_rows_kernel");
  if (rc) return rc;
  ce_mean_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(row_loss, labels, B, S, V, ignore_index, shift, out_loss2);
  return check_launch("ce_mean_kernel");
}
