// Round-2 design probes for the weight-streaming GEMM (run under gpurun; results in profiles/r2_probe_ws.txt).
//   A. tcgen05.mma issue / execution rate: cycles per MMA for cta_group::1 (M=128, M=64) and cta_group::2 (M=256) at several N,
//      back-to-back from one elected thread, shared memory holding arbitrary data (timing only).
//   B. TMA weight-stream rate with the consumer releasing every ring slot at once (no MMAs): canonical [N, K] row-major weights
//      (128-byte runs per row per k-block, rows 8 KB apart) vs a pre-tiled layout (each TMA box is one contiguous 16/32 KB
//      run), with and without the activation tile re-read from L2 beside it.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o probe_ws probe_ws.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------- probe A
// CG = 1: every CTA issues `iters` x 4 MMAs (M x N x 16) into NACC accumulators round-robin.  CG = 2: pairs, leader issues.
template <int CG>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int M, int N, int nacc, int iters, long long* out_cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // fill shared memory with small finite bf16 values
  for (int i = threadIdx.x; i < (160 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CG == 2) cluster_sync_all();
  else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;
  long long t0 = 0, t1 = 0;
  if (warp == 0 && rank == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t nstride = N <= 128 ? 128 : 256;
    const uint64_t da0 = make_desc(smem_u32(smem));
    const uint64_t db0 = make_desc(smem_u32(smem) + 64 * 1024);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tmem_base + (uint32_t)(it % nacc) * nstride;
      const uint64_t da = da0 + (uint64_t)((it & 3) * 1024);   // walk a few "stages" (16 KB apart)
      const uint64_t db = db0 + (uint64_t)((it & 3) * 1024);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (CG == 1) {
          asm volatile(
              "{\n.reg .pred p, q;\nelect.sync _|q, 0xffffffff;\nsetp.ne.b32 p, %4, 0;\n"
              "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
              "l"(da + (uint64_t)(2 * k)), "l"(db + (uint64_t)(2 * k)), "r"(idesc), "r"(1u)
              : "memory");
        } else {
          asm volatile(
              "{\n.reg .pred p, q;\nelect.sync _|q, 0xffffffff;\nsetp.ne.b32 p, %4, 0;\n"
              "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
              "l"(da + (uint64_t)(2 * k)), "l"(db + (uint64_t)(2 * k)), "r"(idesc), "r"(1u)
              : "memory");
        }
      }
    }
    if (CG == 1) {
      asm volatile("{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\n@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(
                       smem_u32(&bar))
                   : "memory");
    } else {
      asm volatile("{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\n@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n}\n" ::"r"(
                       smem_u32(&bar)),
                   "h"((uint16_t)1)
                   : "memory");
    }
    mbar_wait(&bar, 0);
    t1 = clock64();
    if ((threadIdx.x & 31) == 0) out_cycles[blockIdx.x / CG] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CG == 2) cluster_sync_all();
  else __syncthreads();
  if (warp == 0) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------- probe B
struct StreamParams {
  int stages, stage_bytes;
  int w_boxes;        // TMA boxes of W per stage
  int w_box_bytes;    // bytes of one W box
  int w_box_rows;     // rows of one W box (coordinate step)
  int tiled;          // 1: W is the 2-D [rows_total, 64] pre-tiled image (box i of the CTA -> row (base + i) * w_box_rows)
  int kb_total;       // canonical: k-blocks (of 64) per row tile
  int units_total;    // tiled: number of boxes in the whole matrix
  int x_bytes;        // bytes of the activation box per stage (0 = none)
  int x_rows;
  int x_kb;           // k-blocks of the activation matrix (wraps)
  int passes;         // the CTA walks its range this many times (L2-resident ceiling runs)
};

__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX,
                                                      const StreamParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[8], empty_bar[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  int n_iter, base;
  if (p.tiled) {
    const int per = p.units_total / p.w_boxes;  // stage-units
    const int lo = (int)((long long)per * blockIdx.x / gridDim.x), hi = (int)((long long)per * (blockIdx.x + 1) / gridDim.x);
    n_iter = hi - lo;
    base = lo * p.w_boxes;
  } else {
    n_iter = p.kb_total / p.w_boxes;
    base = blockIdx.x * p.w_box_rows;
  }
  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int ii = 0; ii < n_iter * p.passes; ++ii) {
        const int i = ii % n_iter;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        uint8_t* dst = smem + s * p.stage_bytes;
        mbar_expect_tx(&full_bar[s], (uint32_t)(p.w_boxes * p.w_box_bytes + p.x_bytes));
        for (int j = 0; j < p.w_boxes; ++j) {
          int c0, c1;
          if (p.tiled) { c0 = 0; c1 = (base + i * p.w_boxes + j) * p.w_box_rows; }
          else { c0 = (i * p.w_boxes + j) * 64; c1 = base; }
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                           smem_u32(dst + j * p.w_box_bytes)),
                       "l"(&tmW), "r"(smem_u32(&full_bar[s])), "r"(c0), "r"(c1)
                       : "memory");
        }
        if (p.x_bytes) {
          const int kb = (i + blockIdx.x * 7) % p.x_kb;
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                           smem_u32(dst + p.w_boxes * p.w_box_bytes)),
                       "l"(&tmX), "r"(smem_u32(&full_bar[s])), "r"(kb * 64), "r"(0)
                       : "memory");
        }
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < n_iter * p.passes; ++i) {
        mbar_wait(&full_bar[s], ph);
        mbar_arrive(&empty_bar[s]);
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_enc = nullptr;

static CUtensorMap map2d(void* base, uint64_t d0, uint64_t d1, uint64_t stride1_bytes, uint32_t b0, uint32_t b1, CUtensorMapL2promotion promo) {
  CUtensorMap tm;
  cuuint64_t gd[2] = {d0, d1};
  cuuint64_t gs[1] = {stride1_bytes};
  cuuint32_t bx[2] = {b0, b1}, es[2] = {1, 1};
  CUresult r = g_enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("cuTensorMapEncodeTiled failed %d\n", (int)r);
    exit(1);
  }
  return tm;
}

static double run_stream(const char* label, void* const* wbufs, int nbuf, void* xbuf, bool tiled, int w_box_rows, int w_boxes, int x_rows,
                         int grid, int N, int K, int passes = 1) {
  StreamParams p = {};
  p.tiled = tiled ? 1 : 0;
  p.w_box_rows = w_box_rows;
  p.w_box_bytes = w_box_rows * 128;
  p.w_boxes = w_boxes;
  p.x_rows = x_rows;
  p.x_bytes = x_rows * 128;
  p.x_kb = K / 64;
  p.passes = passes;
  p.kb_total = K / 64;
  p.units_total = (int)((long long)N * K * 2 / p.w_box_bytes);
  p.stage_bytes = (p.w_boxes * p.w_box_bytes + p.x_bytes + 1023) / 1024 * 1024;
  p.stages = std::min(8, (220 * 1024) / p.stage_bytes);
  std::vector<CUtensorMap> maps;
  for (int i = 0; i < nbuf; ++i) {
    if (tiled) maps.push_back(map2d(wbufs[i], 64, (uint64_t)N * K / 64, 128, 64, (uint32_t)w_box_rows, CU_TENSOR_MAP_L2_PROMOTION_L2_256B));
    else maps.push_back(map2d(wbufs[i], (uint64_t)K, (uint64_t)N, (uint64_t)K * 2, 64, (uint32_t)w_box_rows, CU_TENSOR_MAP_L2_PROMOTION_L2_256B));
  }
  CUtensorMap tmX = map2d(xbuf, (uint64_t)K, 208, (uint64_t)K * 2, 64, (uint32_t)(x_rows ? x_rows : 8), CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  const int smem_bytes = p.stages * p.stage_bytes;
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int reps = 12;
  for (int i = 0; i < 3; ++i) stream_kernel<<<grid, 64, smem_bytes>>>(maps[i % nbuf], tmX, p);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) stream_kernel<<<grid, 64, smem_bytes>>>(maps[i % nbuf], tmX, p);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double us = ms * 1e3 / reps;
  double bytes = (tiled ? (double)N * K * 2 : (double)grid * w_box_rows * K * 2) * passes;
  printf("B %-58s grid %3d stages %d stage %6d B : %7.2f us/launch  W %6.0f GB/s  (SM ingest %6.0f GB/s)\n", label, grid, p.stages, p.stage_bytes, us,
         bytes / us / 1e3, bytes * (1.0 + (double)p.x_bytes / (p.w_boxes * p.w_box_bytes)) / us / 1e3);
  return us;
}

int main() {
  CK(cudaSetDevice(0));
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  g_enc = (EncodeTiledFn)fp;
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("device clock attr %d kHz\n", clk);

  // ---------------- A: MMA rate
  long long* d_cyc;
  CK(cudaMalloc(&d_cyc, 148 * sizeof(long long)));
  std::vector<long long> h(148);
  const int iters = 512;
  struct Case { int cg, M, N, nacc; };
  const Case cases[] = {{1, 128, 64, 1},  {1, 128, 128, 1}, {1, 128, 208, 1}, {1, 128, 256, 1}, {1, 128, 208, 2}, {1, 128, 256, 2}, {1, 64, 208, 1},
                        {1, 64, 256, 1},  {2, 256, 128, 1}, {2, 256, 208, 1}, {2, 256, 256, 1}, {2, 256, 208, 2}, {2, 256, 256, 2}, {2, 128, 208, 1},
                        {2, 128, 256, 1}};
  CK(cudaFuncSetAttribute(mma_rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(mma_rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (const Case& c : cases) {
    for (int rep = 0; rep < 2; ++rep) {
      if (c.cg == 1) {
        mma_rate_kernel<1><<<148, 128, 200 * 1024>>>(c.M, c.N, c.nacc, iters, d_cyc);
      } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(148);
        cfg.blockDim = dim3(128);
        cfg.dynamicSmemBytes = 200 * 1024;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2;
        at[0].val.clusterDim.y = at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, mma_rate_kernel<2>, c.M, c.N, c.nacc, iters, d_cyc));
      }
      CK(cudaDeviceSynchronize());
    }
    const int n = c.cg == 1 ? 148 : 74;
    CK(cudaMemcpy(h.data(), d_cyc, n * sizeof(long long), cudaMemcpyDeviceToHost));
    std::sort(h.begin(), h.begin() + n);
    const double per = (double)h[n / 2] / (iters * 4);
    const double macs = (double)c.M * c.N * 16;
    printf("A cta_group %d  M %3d N %3d acc %d : %7.1f cycles/MMA (min %.1f max %.1f)  %6.0f MAC/cycle/SM\n", c.cg, c.M, c.N, c.nacc, per,
           (double)h[0] / (iters * 4), (double)h[n - 1] / (iters * 4), macs / per / c.cg);
  }

  // ---------------- B: stream rate
  const int N = 28672, K = 4096;
  const size_t wbytes = (size_t)N * K * 2;
  void* wb[3];
  for (int i = 0; i < 3; ++i) {
    CK(cudaMalloc(&wb[i], wbytes));
    CK(cudaMemset(wb[i], i + 1, wbytes));
  }
  void* xb;
  CK(cudaMalloc(&xb, 208 * K * 2));
  CK(cudaMemset(xb, 1, 208 * K * 2));
  // reference: plain device-to-device copy rate of the same box
  {
    void* dst;
    CK(cudaMalloc(&dst, wbytes));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) CK(cudaMemcpyAsync(dst, wb[i], wbytes, cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 6; ++i) CK(cudaMemcpyAsync(dst, wb[i % 3], wbytes, cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("B cudaMemcpy D2D %zu MB: %.1f us -> %.0f GB/s read+write\n", wbytes >> 20, ms * 1e3 / 6, 2.0 * wbytes * 6 / (ms * 1e-3) / 1e9);
    cudaFree(dst);
  }
  run_stream("canonical box 208x64 (BK=64), no X", wb, 3, xb, false, 208, 1, 0, 138, N, K);
  run_stream("canonical box 208x64 (BK=64), + X 208x64 from L2", wb, 3, xb, false, 208, 1, 208, 138, N, K);
  run_stream("canonical 2 boxes 208x64 per barrier (BK=128), no X", wb, 3, xb, false, 208, 2, 0, 138, N, K);
  run_stream("canonical 2 boxes 208x64 per barrier (BK=128), + X", wb, 3, xb, false, 208, 2, 208, 138, N, K);
  run_stream("canonical box 128x64, no X (grid 148, 1 tile each)", wb, 3, xb, false, 128, 1, 0, 148, N, K);
  run_stream("canonical box 256x64, no X (grid 112)", wb, 3, xb, false, 256, 1, 0, 112, N, K);
  run_stream("tiled 16 KB boxes (128 rows), no X", wb, 3, xb, true, 128, 1, 0, 148, N, K);
  run_stream("tiled 32 KB boxes (256 rows), no X", wb, 3, xb, true, 256, 1, 0, 148, N, K);
  run_stream("tiled 2 x 32 KB boxes per barrier, no X", wb, 3, xb, true, 256, 2, 0, 148, N, K);
  run_stream("tiled 32 KB boxes + X 208x64 from L2", wb, 3, xb, true, 256, 1, 208, 148, N, K);
  run_stream("tiled 32 KB boxes + X 104x64 from L2 (pair half)", wb, 3, xb, true, 256, 1, 104, 148, N, K);
  run_stream("tiled 16 KB boxes + X 208x64 from L2", wb, 3, xb, true, 128, 1, 208, 148, N, K);
  run_stream("tiled 16 KB boxes + X 104x64 from L2", wb, 3, xb, true, 128, 1, 104, 148, N, K);
  // same tiled stream, everything L2-resident (one small buffer): the L2 -> SM ceiling
  {
    void* small[1] = {wb[0]};
    run_stream("tiled 32 KB boxes, 29 MB matrix x16 passes (L2-resident), no X", small, 1, xb, true, 256, 1, 0, 148, N / 8, K, 16);
    run_stream("tiled 32 KB boxes, 29 MB matrix x16 passes (L2-resident), + X 208", small, 1, xb, true, 256, 1, 208, 148, N / 8, K, 16);
  }
  printf("done\n");
  return 0;
}
