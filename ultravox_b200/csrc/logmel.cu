// uvx_logmel: Whisper log-mel front end on the GPU (fp32).
//
// Pass 1 (logmel_power_kernel): one CTA per kFT consecutive frames of one clip.  The reflect-padded, hann-
//   windowed samples of those frames are staged once in shared memory (coalesced 128-bit loads; frames
//   overlap by 240 samples so each sample is fetched from HBM once per CTA), the 201-bin power spectrum is a
//   400 = 16 x 25 factorised real DFT (16-point column DFTs using the conjugate symmetry of real input, twiddle,
//   25-point row DFTs for the 201 wanted bins only: 29 k FMA per frame instead of the 161 k of the round-1 direct
//   DFT) against a shared-memory twiddle table (exact table indexing - no angle accumulation error), the
//   slaney filterbank is applied from its sparse form (<= 16 taps per filter),
//   log10(max(.,1e-10)) is written time-major to the workspace and the per-clip maximum is folded with one
//   atomicMax per CTA.
// Pass 2 (logmel_finish_kernel): max(x, clipmax - 8), (x + 4) / 4, written as the reference's
//   [B, n_mels, T] fp32 layout (shared-memory transpose) and/or the bf16 time-major guard-padded layout the
//   conv stem consumes.
// HBM traffic: 4 B/sample in, 4 B x n_mels x T scratch out+in, outputs.
#include <math.h>

#include <mutex>

#include "uvx_common.cuh"

namespace uvx {

static constexpr int kNfft = 400;
static constexpr int kHop = 160;
static constexpr int kBins = 201;
static constexpr int kFT = 8;        // frames per CTA
static constexpr int kMaxTaps = 16;  // non-zeros per mel filter (128 mels: <= 9, 80 mels: <= 14)
static constexpr int kMaxMels = 128;
static constexpr int kThreads = 256;
static constexpr int kZStride = 25 * kFT + 2;   // float2 elements per k1 row of the 16 x 25 intermediate (+16 bytes: bank spread)

struct MelTables {
  float2* twiddle;  // [400] (cos, sin)(2 pi j / 400)
  float* window;    // [400] periodic hann
  int* tap_start;   // [n_mels]
  int* tap_count;   // [n_mels]
  float* tap_w;     // [n_mels][kMaxTaps]
};

static double hz_to_mel(double f) {
  return f >= 1000.0 ? 15.0 + log(f / 1000.0) * (27.0 / log(6.4)) : 3.0 * f / 200.0;
}
static double mel_to_hz(double m) { return m >= 15.0 ? 1000.0 * exp((log(6.4) / 27.0) * (m - 15.0)) : 200.0 * m / 3.0; }

// slaney-scale, slaney-normalised triangular filters (hf:audio_utils.py:453-545), built in double then cast
// to fp32 exactly like WhisperFeatureExtractor does before the matmul.
static void build_filters(int n_mels, int* start, int* count, float* w) {
  const double mel_min = hz_to_mel(0.0), mel_max = hz_to_mel(8000.0);
  double hz[kMaxMels + 2];
  const double step = (mel_max - mel_min) / (double)(n_mels + 1);  // numpy.linspace: arange * step + start, last = stop
  for (int i = 0; i < n_mels + 2; ++i) hz[i] = mel_to_hz(i == n_mels + 1 ? mel_max : (double)i * step + mel_min);
  for (int m = 0; m < n_mels; ++m) {
    const double enorm = 2.0 / (hz[m + 2] - hz[m]);
    int first = -1, n = 0;
    for (int k = 0; k < kBins; ++k) {
      const double f = 8000.0 * k / (double)(kBins - 1);
      const double down = (f - hz[m]) / (hz[m + 1] - hz[m]);
      const double up = (hz[m + 2] - f) / (hz[m + 2] - hz[m + 1]);
      const double v = fmax(0.0, fmin(down, up)) * enorm;
      const float vf = (float)v;
      if (vf != 0.0f) {
        if (first < 0) first = k;
        if (k - first < kMaxTaps) {
          w[m * kMaxTaps + (k - first)] = vf;
          n = k - first + 1;
        }
      }
    }
    start[m] = first < 0 ? 0 : first;
    count[m] = n;
  }
}

static MelTables* get_tables(int n_mels) {
  // immutable per-(device, n_mels) tables, built once
  static std::mutex mu;
  static MelTables cache[16][2];
  static bool ready[16][2] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  const int slot = n_mels == 80 ? 0 : 1;
  std::lock_guard<std::mutex> lock(mu);
  if (dev < 0 || dev >= 16) return nullptr;
  if (!ready[dev][slot]) {
    float2 tw[kNfft];
    float win[kNfft];
    for (int j = 0; j < kNfft; ++j) {
      const double a = 2.0 * M_PI * j / (double)kNfft;
      tw[j] = make_float2((float)cos(a), (float)sin(a));
      win[j] = (float)(0.5 - 0.5 * cos(a));  // torch.hann_window(400) (periodic)
    }
    static int start[kMaxMels], count[kMaxMels];
    static float w[kMaxMels * kMaxTaps];
    memset(w, 0, sizeof(w));
    build_filters(n_mels, start, count, w);
    MelTables t;
    if (cudaMalloc(&t.twiddle, sizeof(tw)) != cudaSuccess) return nullptr;
    cudaMalloc(&t.window, sizeof(win));
    cudaMalloc(&t.tap_start, sizeof(int) * n_mels);
    cudaMalloc(&t.tap_count, sizeof(int) * n_mels);
    cudaMalloc(&t.tap_w, sizeof(float) * n_mels * kMaxTaps);
    cudaMemcpy(t.twiddle, tw, sizeof(tw), cudaMemcpyHostToDevice);
    cudaMemcpy(t.window, win, sizeof(win), cudaMemcpyHostToDevice);
    cudaMemcpy(t.tap_start, start, sizeof(int) * n_mels, cudaMemcpyHostToDevice);
    cudaMemcpy(t.tap_count, count, sizeof(int) * n_mels, cudaMemcpyHostToDevice);
    cudaMemcpy(t.tap_w, w, sizeof(float) * n_mels * kMaxTaps, cudaMemcpyHostToDevice);
    cache[dev][slot] = t;
    ready[dev][slot] = true;
  }
  return &cache[dev][slot];
}

__device__ __forceinline__ int float_to_ordered(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

__global__ void logmel_init_kernel(int* clipmax, int64_t B) {
  pdl_trigger();
  pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) clipmax[i] = float_to_ordered(-INFINITY);
}

__global__ void __launch_bounds__(kThreads) logmel_power_kernel(const float* __restrict__ wave, int64_t L, int64_t T, int n_mels,
                                                               MelTables tab, float* __restrict__ scratch,
                                                               int* __restrict__ clipmax) {
  pdl_trigger();
  pdl_wait();
  __shared__ float2 s_tw[kNfft];
  __shared__ __align__(16) float s_x[kNfft][kFT];            // windowed samples, [n][frame]; reused as the power spectrum
  __shared__ __align__(16) float2 s_z[16][kZStride];         // stage-A/B output Z[k1][n2][frame] (row padded: conflict-free)
  __shared__ float s_red[kThreads / 32];
  float (*s_pow)[kBins + 3] = reinterpret_cast<float (*)[kBins + 3]>(&s_x[0][0]);   // [kFT][kBins + 3] after the DFT

  const int64_t b = blockIdx.y;
  const int64_t t0 = (int64_t)blockIdx.x * kFT;
  const float* w = wave + b * L;

  for (int j = threadIdx.x; j < kNfft; j += kThreads) s_tw[j] = tab.twiddle[j];
  // frame f covers padded samples [ (t0+f)*160 , +400 ) where padded index p maps to original p - 200 (reflect)
  for (int i = threadIdx.x; i < kNfft * kFT; i += kThreads) {
    const int f = i / kNfft, n = i % kNfft;  // consecutive threads -> consecutive samples (coalesced)
    const int64_t t = t0 + f;
    float v = 0.f;
    if (t < T) {
      int64_t p = t * kHop + n - kNfft / 2;
      if (p < 0) p = -p;
      if (p >= L) p = 2 * (L - 1) - p;
      v = w[p] * tab.window[n];
    }
    s_x[n][f] = v;
  }
  __syncthreads();

  // 400-point real DFT as 16 x 25 (Cooley-Tukey, n = 25 n1 + n2, k = k1 + 16 k2; twiddles e^{-i 2 pi j / 400} = (cos, -sin) of the
  // exact table): X[k] = sum_n2 W25^{n2 k2} * ( W400^{n2 k1} * sum_n1 x[25 n1 + n2] W16^{n1 k1} ) - 29 k FMA per frame instead of 161 k.
  // Stage A + B: thread (k1 <= 8, n2) - the input is real, so Y[16 - k1] = conj(Y[k1]) - for the kFT frames.
  if (threadIdx.x < 9 * 25) {
    const int k1 = threadIdx.x / 25, n2 = threadIdx.x % 25;
    float re[kFT], im[kFT];
#pragma unroll
    for (int f = 0; f < kFT; ++f) re[f] = im[f] = 0.f;
#pragma unroll 4
    for (int n1 = 0; n1 < 16; ++n1) {
      const float2 tw = s_tw[25 * ((n1 * k1) & 15)];          // W16^{n1 k1} = (cos, -sin)
      const float4 xa = *reinterpret_cast<const float4*>(&s_x[25 * n1 + n2][0]);
      const float4 xb = *reinterpret_cast<const float4*>(&s_x[25 * n1 + n2][4]);
      const float xs[kFT] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
      for (int f = 0; f < kFT; ++f) {
        re[f] = fmaf(xs[f], tw.x, re[f]);
        im[f] = fmaf(xs[f], -tw.y, im[f]);
      }
    }
    const float2 t1 = s_tw[(n2 * k1) % kNfft];               // W400^{n2 k1} = (c, -s)
#pragma unroll
    for (int f = 0; f < kFT; ++f)                             // (re + i im)(c - i s)
      s_z[k1][n2 * kFT + f] = make_float2(re[f] * t1.x + im[f] * t1.y, im[f] * t1.x - re[f] * t1.y);
    if (k1 >= 1 && k1 <= 7) {
      const int kc = 16 - k1;
      const float2 t2 = s_tw[(n2 * kc) % kNfft];
#pragma unroll
      for (int f = 0; f < kFT; ++f)                           // conj(Y) = (re - i im), times (c - i s)
        s_z[kc][n2 * kFT + f] = make_float2(re[f] * t2.x - im[f] * t2.y, -im[f] * t2.x - re[f] * t2.y);
    }
  }
  __syncthreads();
  // Stage C: thread k (< 201): X[k1 + 16 k2] = sum_n2 Z[k1][n2] W25^{n2 k2}; power = |X|^2 (the samples' buffer is free now)
  if (threadIdx.x < kBins) {
    const int k = threadIdx.x, k1 = k & 15, k2 = k >> 4;
    float re[kFT], im[kFT];
#pragma unroll
    for (int f = 0; f < kFT; ++f) re[f] = im[f] = 0.f;
    int idx = 0;                                              // 16 * ((n2 k2) mod 25)
#pragma unroll 5
    for (int n2 = 0; n2 < 25; ++n2) {
      const float2 tw = s_tw[idx];                            // W25^{n2 k2} = (c, -s)
      const float4* zp = reinterpret_cast<const float4*>(&s_z[k1][n2 * kFT]);
#pragma unroll
      for (int h = 0; h < kFT / 2; ++h) {
        const float4 z = zp[h];                               // two frames: (re0, im0, re1, im1)
        re[2 * h] = fmaf(z.x, tw.x, fmaf(z.y, tw.y, re[2 * h]));
        im[2 * h] = fmaf(z.y, tw.x, fmaf(-z.x, tw.y, im[2 * h]));
        re[2 * h + 1] = fmaf(z.z, tw.x, fmaf(z.w, tw.y, re[2 * h + 1]));
        im[2 * h + 1] = fmaf(z.w, tw.x, fmaf(-z.z, tw.y, im[2 * h + 1]));
      }
      idx += 16 * k2;
      while (idx >= kNfft) idx -= kNfft;
    }
    // (s_pow aliases s_x, which nobody has read since the barrier after stage A / B)
#pragma unroll
    for (int f = 0; f < kFT; ++f) s_pow[f][k] = re[f] * re[f] + im[f] * im[f];
  }
  __syncthreads();

  // sparse mel filterbank + log10; scratch is time-major [B, T, n_mels]
  float local_max = -INFINITY;
  for (int i = threadIdx.x; i < kFT * n_mels; i += kThreads) {
    const int f = i / n_mels, m = i % n_mels;
    const int64_t t = t0 + f;
    if (t < T) {
      const int s0 = tab.tap_start[m], cnt = tab.tap_count[m];
      float acc = 0.f;
      for (int j = 0; j < cnt; ++j) acc = fmaf(tab.tap_w[m * kMaxTaps + j], s_pow[f][s0 + j], acc);
      const float lg = log10f(fmaxf(acc, 1e-10f));
      scratch[(b * T + t) * n_mels + m] = lg;
      local_max = fmaxf(local_max, lg);
    }
  }
  local_max = warp_max(local_max);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = local_max;
  __syncthreads();
  if (threadIdx.x == 0) {
    float mx = s_red[0];
    for (int i = 1; i < kThreads / 32; ++i) mx = fmaxf(mx, s_red[i]);
    atomicMax(&clipmax[b], float_to_ordered(mx));
  }
}

// grid: (ceil(T/32), B); block 256.  Reads scratch [B,T,n_mels], writes [B,n_mels,T] f32 and/or [B,T+2,n_mels] bf16.
__global__ void __launch_bounds__(kThreads) logmel_finish_kernel(const float* __restrict__ scratch, const int* __restrict__ clipmax,
                                                                int64_t T, int n_mels, float* __restrict__ out_f32,
                                                                bf16* __restrict__ out_tm) {
  pdl_trigger();
  pdl_wait();
  __shared__ float tile[32][kMaxMels + 1];
  const int64_t b = blockIdx.y;
  const int64_t t0 = (int64_t)blockIdx.x * 32;
  const float floor_v = ordered_to_float(clipmax[b]) - 8.0f;
  for (int i = threadIdx.x; i < 32 * n_mels; i += kThreads) {
    const int f = i / n_mels, m = i % n_mels;
    const int64_t t = t0 + f;
    float v = 0.f;
    if (t < T) {
      v = (fmaxf(scratch[(b * T + t) * n_mels + m], floor_v) + 4.0f) / 4.0f;
      if (out_tm) out_tm[(b * (T + 2) + t + 1) * n_mels + m] = __float2bfloat16_rn(v);
    }
    tile[f][m] = v;
  }
  if (out_tm && blockIdx.x == 0) {
    for (int m = threadIdx.x; m < n_mels; m += kThreads) {
      out_tm[(b * (T + 2)) * n_mels + m] = __float2bfloat16_rn(0.f);
      out_tm[(b * (T + 2) + T + 1) * n_mels + m] = __float2bfloat16_rn(0.f);
    }
  }
  if (out_f32) {
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * n_mels; i += kThreads) {
      const int m = i / 32, f = i % 32;
      const int64_t t = t0 + f;
      if (t < T) out_f32[(b * n_mels + m) * T + t] = tile[f][m];
    }
  }
}

}  // namespace uvx

// host-only: the dense [201, n_mels] fp32 filterbank this library uses (for CPU-side verification)
extern "C" int uvx_debug_mel_filters(int n_mels, float* host_dense) {
  using namespace uvx;
  UVX_REQUIRE(host_dense && (n_mels == 80 || n_mels == 128), "uvx_debug_mel_filters: bad arguments");
  static int start[kMaxMels], count[kMaxMels];
  static float w[kMaxMels * kMaxTaps];
  memset(w, 0, sizeof(w));
  build_filters(n_mels, start, count, w);
  memset(host_dense, 0, sizeof(float) * kBins * n_mels);
  for (int m = 0; m < n_mels; ++m)
    for (int j = 0; j < count[m]; ++j) host_dense[(start[m] + j) * n_mels + m] = w[m * kMaxTaps + j];
  return UVX_OK;
}

extern "C" size_t uvx_logmel_workspace(int64_t B, int64_t L, int n_mels) {
  const int64_t T = L / uvx::kHop;
  return (size_t)(B * T * n_mels * 4 + ((B * 4 + 255) / 256) * 256);
}

extern "C" int uvx_logmel(const float* wave, int64_t B, int64_t L, int n_mels, float* out_f32, void* out_tm, void* workspace,
                          size_t workspace_bytes, uvx_stream_t stream) {
  using namespace uvx;
  cudaStream_t st = (cudaStream_t)stream;
  UVX_REQUIRE(wave && workspace && (out_f32 || out_tm), "uvx_logmel: null pointer");
  UVX_REQUIRE(n_mels == 80 || n_mels == 128, "uvx_logmel: n_mels must be 80 or 128 (got %d)", n_mels);
  UVX_REQUIRE(B >= 1 && B < 65536 && L >= 2 * kHop && L % kHop == 0, "uvx_logmel: L must be a multiple of 160 and >= 320 (got %lld)",
              (long long)L);
  UVX_REQUIRE((uintptr_t)workspace % 16 == 0, "uvx_logmel: workspace must be 16-byte aligned");
  if (workspace_bytes < uvx_logmel_workspace(B, L, n_mels)) {
    set_error("uvx_logmel: workspace too small (%zu < %zu)", workspace_bytes, uvx_logmel_workspace(B, L, n_mels));
    return UVX_ERR_WS;
  }
  MelTables* tab = get_tables(n_mels);
  if (!tab) {
    set_error("uvx_logmel: could not build device tables");
    return UVX_ERR_CUDA;
  }
  const int64_t T = L / kHop;
  int* clipmax = (int*)workspace;
  float* scratch = (float*)((uint8_t*)workspace + ((B * 4 + 255) / 256) * 256);
  launch_k(logmel_init_kernel, dim3((unsigned)((B + 255) / 256)), dim3(256), 0, st, clipmax, B);
  int rc = check_launch("logmel_init_kernel");
  if (rc) return rc;
  dim3 g1((unsigned)((T + kFT - 1) / kFT), (unsigned)B);
  launch_k(logmel_power_kernel, dim3(g1), dim3(kThreads), 0, st, wave, L, T, n_mels, *tab, scratch, clipmax);
  rc = check_launch("logmel_power_kernel");
  if (rc) return rc;
  dim3 g2((unsigned)((T + 31) / 32), (unsigned)B);
  launch_k(logmel_finish_kernel, dim3(g2), dim3(kThreads), 0, st, scratch, clipmax, T, n_mels, out_f32, (bf16*)out_tm);
  return check_launch("logmel_finish_kernel");
}
