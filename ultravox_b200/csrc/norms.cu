// Row-wise LayerNorm / RMSNorm: HBM-bound, one CTA per row, 128-bit loads, fp32 statistics, warp shuffles.
#include "uvx_common.cuh"

namespace uvx {

static constexpr int kNormThreads = 256;
static constexpr int kMaxVec = 8;  // up to 256 * 8 * 8 = 16384 columns

// LayerNorm (nn.LayerNorm semantics: biased variance, two-pass in registers)
__global__ void __launch_bounds__(kNormThreads) layernorm_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                                 const bf16* __restrict__ b, bf16* __restrict__ y,
                                                                 int64_t cols, int64_t x_row_stride, float eps) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[32];
  const int64_t row = blockIdx.x;
  const bf16* xr = x + row * x_row_stride;
  const int nvec = (int)(cols / 8);
  float v[kMaxVec][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int j = threadIdx.x + i * kNormThreads;
    if (j < nvec) {
      unpack8(*reinterpret_cast<const bf16x8*>(xr + (int64_t)j * 8), v[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v[i][e];
    }
  }
  const float mean = block_sum(s, red) / (float)cols;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int j = threadIdx.x + i * kNormThreads;
    if (j < nvec) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = v[i][e] - mean;
        sq += d * d;
      }
    }
  }
  const float rstd = rsqrtf(block_sum(sq, red) / (float)cols + eps);
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int j = threadIdx.x + i * kNormThreads;
    if (j < nvec) {
      float wv[8], bv[8], o[8];
      unpack8(*reinterpret_cast<const bf16x8*>(w + (int64_t)j * 8), wv);
      unpack8(*reinterpret_cast<const bf16x8*>(b + (int64_t)j * 8), bv);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (v[i][e] - mean) * rstd * wv[e] + bv[e];
      *reinterpret_cast<bf16x8*>(y + row * cols + (int64_t)j * 8) = pack8(o);
    }
  }
}

// Warp-per-row variant for cols <= 2048 (the Whisper encoder's d_model): no block barriers, 4 rows per CTA.
static constexpr int kLnWarpVec = 8;
__global__ void __launch_bounds__(128) layernorm_warp_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                             const bf16* __restrict__ b, bf16* __restrict__ y, int64_t rows,
                                                             int64_t cols, int64_t x_row_stride, float eps) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const bf16* xr = x + row * x_row_stride;
  const int nvec = (int)(cols / 8);
  float v[kLnWarpVec][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kLnWarpVec; ++i) {
    const int j = lane + i * 32;
    if (j < nvec) {
      unpack8(*reinterpret_cast<const bf16x8*>(xr + (int64_t)j * 8), v[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v[i][e];
    }
  }
  const float mean = warp_sum(s) / (float)cols;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kLnWarpVec; ++i) {
    const int j = lane + i * 32;
    if (j < nvec) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = v[i][e] - mean;
        sq += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / (float)cols + eps);
#pragma unroll
  for (int i = 0; i < kLnWarpVec; ++i) {
    const int j = lane + i * 32;
    if (j < nvec) {
      float wv[8], bv[8], o[8];
      unpack8(*reinterpret_cast<const bf16x8*>(w + (int64_t)j * 8), wv);
      unpack8(*reinterpret_cast<const bf16x8*>(b + (int64_t)j * 8), bv);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (v[i][e] - mean) * rstd * wv[e] + bv[e];
      *reinterpret_cast<bf16x8*>(y + row * cols + (int64_t)j * 8) = pack8(o);
    }
  }
}

// RMSNorm with the exact rounding order of LlamaRMSNorm: y = w * bf16(x * rsqrt(mean(x^2) + eps)).
// Grouped mode implements StackAudioFrames: elements past `valid` in a row read as zero.
__global__ void __launch_bounds__(kNormThreads) rmsnorm_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                               bf16* __restrict__ y, int64_t cols, int64_t x_row_stride,
                                                               int64_t group_rows, int64_t group_stride,
                                                               int64_t valid_elems, float eps) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[32];
  const int64_t row = blockIdx.x;
  const bf16* xr;
  int64_t valid = cols;
  if (group_rows > 0) {
    const int64_t g = row / group_rows, t = row % group_rows;
    xr = x + g * group_stride + t * cols;
    valid = valid_elems - t * cols;
    valid = valid < 0 ? 0 : (valid > cols ? cols : valid);
  } else {
    xr = x + row * x_row_stride;
  }
  const int nvec = (int)(cols / 8);
  float v[kMaxVec][8];
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int j = threadIdx.x + i * kNormThreads;
    if (j < nvec) {
      if ((int64_t)j * 8 < valid) {  // valid is a multiple of the frame width (multiple of 8)
        unpack8(*reinterpret_cast<const bf16x8*>(xr + (int64_t)j * 8), v[i]);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[i][e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) sq += v[i][e] * v[i][e];
    }
  }
  const float rstd = rsqrtf(block_sum(sq, red) / (float)cols + eps);
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int j = threadIdx.x + i * kNormThreads;
    if (j < nvec) {
      float wv[8], o[8];
      unpack8(*reinterpret_cast<const bf16x8*>(w + (int64_t)j * 8), wv);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = wv[e] * __bfloat162float(__float2bfloat16_rn(v[i][e] * rstd));
      *reinterpret_cast<bf16x8*>(y + row * cols + (int64_t)j * 8) = pack8(o);
    }
  }
}

}  // namespace uvx

extern "C" int uvx_layernorm(const void* x, const void* w, const void* b, void* y, int64_t rows, int64_t cols,
                             int64_t x_row_stride, float eps, uvx_stream_t stream) {
  using namespace uvx;
  UVX_REQUIRE(x && w && b && y, "uvx_layernorm: null pointer");
  UVX_REQUIRE(cols % 8 == 0 && cols <= kNormThreads * kMaxVec * 8 && x_row_stride % 8 == 0,
              "uvx_layernorm: cols must be a multiple of 8 and <= %d", kNormThreads * kMaxVec * 8);
  if (rows == 0) return UVX_OK;
  if (cols <= 32 * kLnWarpVec * 8) {
    launch_k(layernorm_warp_kernel, dim3((unsigned)((rows + 3) / 4)), dim3(128), 0, (cudaStream_t)stream, (const bf16*)x, (const bf16*)w, (const bf16*)b,
                                                                                       (bf16*)y, rows, cols, x_row_stride, eps);
    return check_launch("layernorm_warp_kernel");
  }
  launch_k(layernorm_kernel, dim3((unsigned)rows), dim3(kNormThreads), 0, (cudaStream_t)stream, (const bf16*)x, (const bf16*)w, (const bf16*)b,
                                                                             (bf16*)y, cols, x_row_stride, eps);
  return check_launch("layernorm_kernel");
}

extern "C" int uvx_rmsnorm(const void* x, const void* w, void* y, int64_t rows, int64_t cols, int64_t x_row_stride,
                           int64_t group_rows, int64_t group_stride, int64_t valid_elems, float eps, uvx_stream_t stream) {
  using namespace uvx;
  UVX_REQUIRE(x && w && y, "uvx_rmsnorm: null pointer");
  UVX_REQUIRE(cols % 8 == 0 && cols <= kNormThreads * kMaxVec * 8 && x_row_stride % 8 == 0,
              "uvx_rmsnorm: cols must be a multiple of 8 and <= %d", kNormThreads * kMaxVec * 8);
  UVX_REQUIRE(group_rows == 0 || (valid_elems % 8 == 0 && group_stride % 8 == 0), "uvx_rmsnorm: group alignment");
  if (rows == 0) return UVX_OK;
  launch_k(rmsnorm_kernel, dim3((unsigned)rows), dim3(kNormThreads), 0, (cudaStream_t)stream, (const bf16*)x, (const bf16*)w, (bf16*)y, cols,
                                                                           x_row_stride, group_rows, group_stride,
                                                                           valid_elems, eps);
  return check_launch("rmsnorm_kernel");
}
