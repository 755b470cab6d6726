// Prototype / microbenchmark of the hierarchical top pass (development aid, not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o mb_fold mb_fold.cu
// Each CTA holds a random window in shared memory and its warps run all tops of one class (L) `iters`
// times; reports useful shared-memory bytes per clock per SM (128 = peak).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>

#ifndef OCC
#define OCC 2
#endif
constexpr int kThreads = 256, kWarps = 8;
extern __shared__ __align__(128) unsigned char smem_raw[];

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int J>
__device__ __forceinline__ double sum_sq(const double (&v)[J]) {
  if constexpr (J == 1) return v[0] * v[0];
  else {
    double lo = v[0] * v[0], hi = v[1] * v[1];
#pragma unroll
    for (int j = 2; j < J; j += 2) { lo = fma(v[j], v[j], lo); if (j + 1 < J) hi = fma(v[j + 1], v[j + 1], hi); }
    return lo + hi;
  }
}
template <int J>
__device__ __forceinline__ void add_row(double (&acc)[J], const double* __restrict__ row) {
  double t[J];
#pragma unroll
  for (int j = 0; j < J; ++j) t[j] = row[32 * j];
#pragma unroll
  for (int j = 0; j < J; ++j) acc[j] += t[j];
}

// levels LV..0 of one tile; sets with index >= sets - sB have one more term on every residue, set 0 one more
// on the lanes flagged in isA (tail row).
template <int L, int J, int LV>
struct levels {
  static __device__ __forceinline__ void run(double (&acc)[1 << L][J], int M0, const bool (&isA)[J], double (&T)[L + 1],
                                             double (&A)[L + 1]) {
    constexpr int sets = 1 << LV;
    const int sB = M0 & (sets - 1);
#pragma unroll
    for (int s = 0; s < sets; ++s) {
      const double q = sum_sq<J>(acc[s]);
      T[LV] += q;
      if (s >= sets - sB) A[LV] += q;  // warp-uniform
    }
#pragma unroll
    for (int j = 0; j < J; ++j)
      if (isA[j]) A[LV] = fma(acc[0][j], acc[0][j], A[LV]);
    if constexpr (LV > 0) {
      constexpr int half = sets >> 1;
#pragma unroll
      for (int s = 0; s < half; ++s)
#pragma unroll
        for (int j = 0; j < J; ++j) acc[s][j] += acc[s + half][j];
      levels<L, J, LV - 1>::run(acc, M0, isA, T, A);
    }
  }
};

// one tile: base residues ra + lane + 32 j (j < J) of a top q = g 2^L
template <int L, int J, bool MASK>
__device__ __forceinline__ void tile(const double* __restrict__ xs, int g, int ra, int M0, int rr, double (&T)[L + 1],
                                     double (&A)[L + 1]) {
  constexpr int S = 1 << L;
  const int lane = threadIdx.x & 31;
  const double* ptr = xs + ra + lane;
  double acc[S][J];
  if constexpr (S == 1) {
#pragma unroll
    for (int j = 0; j < J; ++j) acc[0][j] = ptr[32 * j];
    ptr += g;
#pragma unroll 2
    for (int k = 1; k < M0; ++k) { add_row<J>(acc[0], ptr); ptr += g; }
  } else {
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
      for (int j = 0; j < J; ++j) acc[s][j] = 0.0;
    const int sB = M0 & (S - 1);
#pragma unroll
    for (int s = 1; s < S; ++s)
      if (s >= S - sB) { add_row<J>(acc[s], ptr); ptr += g; }
#pragma unroll 1
    for (int i = M0 >> L; i > 0; --i) {
#pragma unroll
      for (int s = 0; s < S; ++s) { add_row<J>(acc[s], ptr); ptr += g; }
    }
  }
  // tail row (partial): residues < rr
  bool isA[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    isA[j] = ra + lane + 32 * j < rr;
    double t = 0.0;
    if (isA[j]) t = ptr[32 * j];
    acc[0][j] += t;
  }
  if (MASK) {
#pragma unroll
    for (int j = 0; j < J; ++j)
      if (ra + lane + 32 * j >= g) {
#pragma unroll
        for (int s = 0; s < S; ++s) acc[s][j] = 0.0;
      }
  }
  levels<L, J, L>::run(acc, M0, isA, T, A);
}

template <int L, int J>
__device__ __forceinline__ double top_pass_(const double* xs, int g, int N, const double* rcp, double* en) {
  const int M0 = N / g, rr = N - M0 * g;
  double T[L + 1], A[L + 1];
#pragma unroll
  for (int i = 0; i <= L; ++i) T[i] = A[i] = 0.0;
  int ra = 0;
  for (; ra + 32 * J <= g; ra += 32 * J) tile<L, J, false>(xs, g, ra, M0, rr, T, A);
  if constexpr (J > 2)
    for (; ra + 64 <= g; ra += 64) tile<L, 2, false>(xs, g, ra, M0, rr, T, A);
  for (; ra < g; ra += 32) tile<L, 1, true>(xs, g, ra, M0, rr, T, A);
  double best = 0.0;
#pragma unroll
  for (int i = 0; i <= L; ++i) {
    const int M = M0 >> i;
    const double w_lo = rcp[M], w_diff = rcp[M + 1] - w_lo;
    const double e = warp_sum(fma(w_diff, A[i], w_lo * T[i]));
    best = fmax(best, e);
    if (en != nullptr && (threadIdx.x & 31) == 0) en[g << i] = e;
  }
  return best;
}

template <int L, int J>
__global__ void __launch_bounds__(kThreads, OCC) bench_kernel(int N, int pmax, int iters, double* out, long long* cycles, double* en) {
  double* xs = reinterpret_cast<double*>(smem_raw);
  double* rcp = xs + N + 1024;
  __shared__ int counter;
  for (int i = threadIdx.x; i < N + 1024; i += kThreads) xs[i] = i < N ? sin(0.001 * i * (blockIdx.x + 1)) : 0.0;
  for (int i = threadIdx.x; i < 2048; i += kThreads) rcp[i] = i ? 1.0 / i : 0.0;
  if (threadIdx.x == 0) counter = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  constexpr int S = 1 << L;
  // tops of class L: q in (pmax/2, pmax] with min(ctz(q),3) == L
  const int top_lo = pmax / 2 + 1;
  double best = 0.0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    while (true) {
      int idx = 0;
      if (lane == 0) idx = atomicAdd(&counter, 1);
      idx = __shfl_sync(0xffffffffu, idx, 0);
      const int local = idx - it * 100000;
      // enumerate class-L tops: L=0 odd; L=1 2 mod 4; L=2 4 mod 8; L=3 0 mod 8
      int q;
      if (L == 0) q = (top_lo | 1) + 2 * local;
      else if (L == 1) q = ((top_lo + 1) & ~3) + 2 + 4 * local, q = q < top_lo ? q + 4 : q;
      else if (L == 2) q = ((top_lo + 3) & ~7) + 4 + 8 * local, q = q < top_lo ? q + 8 : q;
      else q = ((top_lo + 7) & ~7) + 8 * local;
      if (local < 0 || q > pmax) break;
      best = fmax(best, top_pass_<L, J>(xs, q / S, N, rcp, blockIdx.x == 0 ? en : nullptr));
    }
    __syncthreads();
    if (threadIdx.x == 0) counter = (it + 1) * 100000;
    __syncthreads();
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) { cycles[blockIdx.x] = t1 - t0; }
  if (lane == 0) out[blockIdx.x * kWarps + (threadIdx.x >> 5)] = best;
}

static double* g_en = nullptr;
template <int L, int J>
void run(int N, int pmax, int iters, int sms) {
  const int grid = sms * OCC;
  double* out; long long* cyc;
  cudaMalloc(&out, grid * kWarps * 8); cudaMalloc(&cyc, grid * 8);
  const size_t smem = (size_t)(N + 1024 + 2048) * 8;
  cudaFuncSetAttribute(bench_kernel<L, J>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench_kernel<L, J><<<grid, kThreads, smem>>>(N, pmax, 2, out, cyc, g_en);
  cudaEventRecord(e0);
  bench_kernel<L, J><<<grid, kThreads, smem>>>(N, pmax, iters, out, cyc, g_en);
  cudaEventRecord(e1);
  cudaError_t err = cudaEventSynchronize(e1);
  if (err != cudaSuccess) { printf("L=%d J=%d: %s\n", L, J, cudaGetErrorString(err)); exit(1); }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long* h = (long long*)malloc(grid * 8); cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
  double mean = 0; for (int i = 0; i < grid; ++i) mean += h[i]; mean /= grid;
  double hsum = 0; double* ho = (double*)malloc(grid * kWarps * 8); cudaMemcpy(ho, out, grid * kWarps * 8, cudaMemcpyDeviceToHost);
  for (int i = 0; i < grid * kWarps; ++i) hsum += ho[i];
  // tops of the class
  int ntops = 0; const int top_lo = pmax / 2 + 1;
  for (int q = top_lo; q <= pmax; ++q) { int c = __builtin_ctz(q); if (c > 3) c = 3; if (c == L) ++ntops; }
  const double bytes = (double)ntops * N * 8 * iters;           // useful bytes per CTA
  printf("L=%d J=%d tops=%3d  %.3f ms  cycles/CTA %.0f  useful B/clk/SM %.1f (%.1f%% of 128)  clk %.0f MHz  chk %.6g\n", L, J,
         ntops, ms, mean, OCC * bytes / mean, OCC * bytes / mean / 1.28, mean / ms * 1e-3, hsum);
  cudaFree(out); cudaFree(cyc); free(h); free(ho);
}

int main(int argc, char** argv) {
  int N = 4096, pmax = 1024, iters = 20, sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaMalloc(&g_en, 2048 * 8); cudaMemset(g_en, 0, 2048 * 8);
  if (argc > 3) {
    const int l = atoi(argv[2]), j = atoi(argv[3]);
    if (l == 0 && j == 8) run<0, 8>(N, pmax, iters, sms);
    if (l == 1 && j == 8) run<1, 8>(N, pmax, iters, sms);
    if (l == 2 && j == 4) run<2, 4>(N, pmax, iters, sms);
    if (l == 3 && j == 2) run<3, 2>(N, pmax, iters, sms);
    return 0;
  }
  run<0, 4>(N, pmax, iters, sms);
  run<0, 8>(N, pmax, iters, sms);
  run<0, 16>(N, pmax, iters, sms);
  run<1, 4>(N, pmax, iters, sms);
  run<1, 8>(N, pmax, iters, sms);
  run<2, 2>(N, pmax, iters, sms);
  run<2, 4>(N, pmax, iters, sms);
  run<3, 1>(N, pmax, iters, sms);
  run<3, 2>(N, pmax, iters, sms);
  { double* h = (double*)malloc(2048 * 8); cudaMemcpy(h, g_en, 2048 * 8, cudaMemcpyDeviceToHost);
    FILE* f = fopen(argc > 1 ? argv[1] : "energies.bin", "wb"); fwrite(h, 8, 2048, f); fclose(f); }
  return 0;
}
