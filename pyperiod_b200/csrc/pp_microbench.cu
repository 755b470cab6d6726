// pyperiod_b200 -- roofline denominators the driver's MEASURED_PEAKS.json does not record:
// shared-memory load bandwidth and FP64 add throughput of the whole chip, measured live
// ("FP64-pipe and shared-memory peaks must be micro-benchmarked on the box").
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pyperiod_b200.h"

#include "pp_host.cuh"

namespace pp {

// Every lane streams conflict-free 16-byte shared-memory loads (4 wavefronts per warp
// instruction = 512 B) and folds them with integer XORs, so only the LSU/crossbar is loaded.
__global__ void __launch_bounds__(1024) smem_bw_kernel(int iters, unsigned long long* sink, long long* cycles) {
  extern __shared__ __align__(16) unsigned char buf[];
  uint4* s = reinterpret_cast<uint4*>(buf);
  const int words = 4096;  // 64 KiB
  for (int i = threadIdx.x; i < words; i += blockDim.x) s[i] = make_uint4(i, i * 3, i * 5, i * 7);
  __syncthreads();
  uint4 acc = make_uint4(0, 0, 0, 0);
  int idx = threadIdx.x;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint4 v = s[(idx + u * 512) & (words - 1)];
      acc.x ^= v.x;
      acc.y ^= v.y;
      acc.z ^= v.z;
      acc.w ^= v.w;
    }
    idx = (idx + 32) & (words - 1);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345679u) *sink = acc.x;
}

// 8 independent DADD chains per thread: saturates the FP64 pipe.
__global__ void __launch_bounds__(1024) dadd_kernel(int iters, double seed, double* sink, long long* cycles) {
  double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6,
         a7 = seed + 7;
  const double inc = seed * 1e-9 + (double)threadIdx.x;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      a0 += inc; a1 += inc; a2 += inc; a3 += inc;
      a4 += inc; a5 += inc; a6 += inc; a7 += inc;
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  const double t = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (t == 0.123456789) *sink = t;
}
// 8 independent m8n8k4 FP64 tensor-core accumulation chains per warp (SASS DMMA): the denominator of the
// Ramanujan contraction's roofline.
__global__ void __launch_bounds__(1024) dmma_kernel(int iters, double seed, double* sink, long long* cycles) {
  double c[8][2];
#pragma unroll
  for (int u = 0; u < 8; ++u) c[u][0] = c[u][1] = seed * (u + 1);
  const double a = seed + 1e-9 * (double)threadIdx.x, b = 1.0 - 1e-9 * (double)(threadIdx.x & 7);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[u][0]), "+d"(c[u][1])
                   : "d"(a), "d"(b));
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  double t = 0.0;
#pragma unroll
  for (int u = 0; u < 8; ++u) t += c[u][0] + c[u][1];
  if (t == 0.123456789) *sink = t;
}
}  // namespace pp

using namespace pp;

// kind 0: shared-memory bandwidth  -> out_host[0] = bytes/s (whole chip)
// kind 1: FP64 add throughput      -> out_host[0] = adds/s  (whole chip)
// kind 2: FP64 tensor-core (DMMA m8n8k4) throughput -> out_host[0] = flop/s (whole chip)
// out_host[1] = SM clock in MHz during the run (device cycles / event time); out_host[2] = ms.
// Synchronous (host timing with CUDA events); call outside any timed region.
extern "C" int pp_microbench(int32_t kind, int32_t iters, double* out_host) {
  if (out_host == nullptr || iters < 1 || kind < 0 || kind > 2) return fail(-1, "bad microbench arguments%s", "");
  int dev = 0, sms = 0;
  if (int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  void* scratch = nullptr;
  if (int rc = check_cuda(cudaMalloc(&scratch, 64), "cudaMalloc")) return rc;
  cudaMemset(scratch, 0, 64);
  long long* cyc = reinterpret_cast<long long*>(scratch);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int threads = 1024, blocks = sms * 2;
  const size_t smem = 65536;
  if (kind == 0) cudaFuncSetAttribute(smem_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  float best_ms = 1e30f;
  long long best_cyc = 0;
  for (int rep = 0; rep < 4; ++rep) {  // first rep is warm-up
    cudaEventRecord(e0);
    if (kind == 0)
      smem_bw_kernel<<<blocks, threads, smem>>>(iters, reinterpret_cast<unsigned long long*>(scratch) + 2, cyc);
    else if (kind == 1)
      dadd_kernel<<<blocks, threads>>>(iters, 1.0, reinterpret_cast<double*>(scratch) + 2, cyc);
    else
      dmma_kernel<<<blocks, threads>>>(iters, 1.0, reinterpret_cast<double*>(scratch) + 2, cyc);
    cudaEventRecord(e1);
    if (int rc = check_cuda(cudaEventSynchronize(e1), "microbench kernel")) {
      cudaFree(scratch);
      return rc;
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long c = 0;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    if (rep > 0 && ms < best_ms) {
      best_ms = ms;
      best_cyc = c;
    }
  }
  const double secs = best_ms * 1e-3;
  // kind 2: 8 MMAs of 8*8*4*2 = 512 flop per warp and iteration = 128 flop per thread
  const double per_thread = kind == 0 ? (double)iters * 8 * 16 : (kind == 1 ? (double)iters * 32 : (double)iters * 128);
  out_host[0] = per_thread * threads * blocks / secs;
  // two CTAs share an SM, so one CTA's cycle count spans (about) the whole kernel
  out_host[1] = (double)best_cyc / secs * 1e-6;
  out_host[2] = best_ms;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(scratch);
  return 0;
}
