// pyperiod_b200 -- shared device primitives for the residue-class fold kernels (sm_100a).
//
// Vocabulary: a *window* is N float64 samples; the *fold* of a window at period p is
// S_p[r] = sum_{n = r (mod p)} x[n]; a *sweep* evaluates one metric of the fold for every
// candidate period; a *slot* is one selected period with its single-period basis vector.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace pp {

#ifndef PP_THREADS
#define PP_THREADS 256
#endif
#ifndef PP_CTAS
#define PP_CTAS 2
#endif
constexpr int kThreads = PP_THREADS;    // threads per CTA
constexpr int kWarps = kThreads / 32;
#ifndef PP_RESBLOCK
#define PP_RESBLOCK 256
#endif
constexpr int kResBlock = PP_RESBLOCK;  // residues one register tile covers (32 lanes x 8 columns)
constexpr int kCtasPerSm = PP_CTAS;     // occupancy target of the sweep kernels (sets the register budget)
constexpr int kSweepPad = 192;          // zero samples after the window: the register tiles of the last,
                                        // partial row may read up to 63 samples past N
constexpr int kMaxFactors = 128;        // non-trivial divisors per period handled by M-best step 2

// The dynamic shared memory of every kernel in this library starts with the staged window (xs).
// Declaring the array at namespace scope lets out-of-line device functions address it as shared
// memory (LDS with immediate offsets) instead of through a generic pointer.
extern __shared__ __align__(128) unsigned char pp_smem[];
__device__ __forceinline__ const double* staged_window() { return reinterpret_cast<const double*>(pp_smem); }

// ------------------------------------------------------------------------------------------
// small PTX wrappers: mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "PP_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra PP_DONE_%=;\n"
      "bra PP_WAIT_%=;\n"
      "PP_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// generic-proxy writes to smem must be ordered before an async-proxy (TMA) overwrite
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------------------------------
// deterministic reductions
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Sum of squares of s[0..n) over the whole CTA; fixed order => run-to-run identical.
// `red` is kWarps doubles of shared scratch.  Result returned to every thread.
__device__ __forceinline__ double cta_sum_sq(const double* s, int n, double* red) {
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += kThreads) a = fma(s[i], s[i], a);
  a = warp_sum(a);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) t += red[w];
  return t;
}

// ------------------------------------------------------------------------------------------
// window staging: HBM -> shared memory (TMA bulk copy when 16-byte aligned, else LDG)
// ------------------------------------------------------------------------------------------
struct WindowLoader {
  uint64_t* bar;
  uint32_t parity;
  __device__ void init(uint64_t* b) {
    bar = b;
    parity = 0;
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
  }
  // all threads call; returns after xs[0..n) holds the window (includes a CTA barrier)
  __device__ void load(double* xs, const double* src, int n) {
    const bool tma_ok = ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) && ((n & 1) == 0);
    if (tma_ok) {
      __syncthreads();  // every generic access to xs of the previous window is done
      if (threadIdx.x == 0) {
        fence_proxy_async();
        const uint32_t bytes = static_cast<uint32_t>(n) * 8u;
        mbar_arrive_expect_tx(bar, bytes);
        for (uint32_t off = 0; off < bytes; off += 32768u) {
          const uint32_t chunk = (bytes - off < 32768u) ? (bytes - off) : 32768u;
          tma_bulk_g2s(reinterpret_cast<char*>(xs) + off, reinterpret_cast<const char*>(src) + off, chunk, bar);
        }
      }
      mbar_wait(bar, parity);
      parity ^= 1u;
    } else {
      __syncthreads();
      for (int i = threadIdx.x; i < n; i += kThreads) xs[i] = __ldg(src + i);
    }
    __syncthreads();
  }
};

// ------------------------------------------------------------------------------------------
// dynamic hand-out of windows to the CTAs of a persistent grid
// ------------------------------------------------------------------------------------------
// Per-window cost varies (accepted periods, rounds, dictionary sizes), and a static stride leaves a tail of
// idle CTAs at the end of a launch.  `counter` is a global int zeroed on the stream before the launch
// (nullptr => static stride).  Usage:  for (WindowQueue q(counter); q.b < B; q.next()) { ... q.b ... }
struct WindowQueue {
  int* counter;
  int b;
  __device__ explicit WindowQueue(int* c) : counter(c), b(blockIdx.x) {}
  __device__ void next() {  // all threads call; contains CTA barriers
    __shared__ int s_next;
    __syncthreads();  // everyone is done with window b
    if (threadIdx.x == 0) s_next = counter ? (int)gridDim.x + atomicAdd(counter, 1) : b + (int)gridDim.x;
    __syncthreads();
    b = s_next;
  }
};

// ------------------------------------------------------------------------------------------
// exact fold + mean: bit-for-bit the reference's numpy arithmetic (Periods.py:171-198)
// ------------------------------------------------------------------------------------------
// Source element n is src[n] (WRAP=false) or src[n mod P] (WRAP=true: the length-N tiling
// of a P-periodic vector, p < P).  Per residue the terms are added in increasing n, the
// zero padding of the reference's rectangle is added as +0.0, and the mean is an IEEE
// division by the reference's divisor (rows for r < N-(rows-1)p, rows-1 after; the number
// of complete rows in trunc mode).  Result: vout[0..p).  No barrier inside.
template <bool WRAP>
__device__ __forceinline__ void cta_fold_mean_exact(const double* src, int P, int N, int p, bool trunc,
                                                    double* __restrict__ vout) {
  const int M = N / p;            // complete rows
  const int r0 = N - M * p;       // residues r < r0 have one more term
  for (int r = threadIdx.x; r < p; r += kThreads) {
    const int terms = trunc ? M : (M + (r < r0 ? 1 : 0));
    double acc;
    if (terms == 0) {
      acc = 0.0;
    } else {
      int idx = r;
      acc = src[idx];
      int k = 1;
      if (!WRAP) {
        const double* ptr = src + r;
        for (; k + 4 <= terms; k += 4) {
          const double v0 = ptr[(size_t)(k)*p], v1 = ptr[(size_t)(k + 1) * p], v2 = ptr[(size_t)(k + 2) * p],
                       v3 = ptr[(size_t)(k + 3) * p];
          acc += v0;
          acc += v1;
          acc += v2;
          acc += v3;
        }
        for (; k < terms; ++k) acc += ptr[(size_t)k * p];
      } else {
        for (; k + 4 <= terms; k += 4) {
          int i0 = idx + p; if (i0 >= P) i0 -= P;
          int i1 = i0 + p;  if (i1 >= P) i1 -= P;
          int i2 = i1 + p;  if (i2 >= P) i2 -= P;
          int i3 = i2 + p;  if (i3 >= P) i3 -= P;
          const double v0 = src[i0], v1 = src[i1], v2 = src[i2], v3 = src[i3];
          acc += v0;
          acc += v1;
          acc += v2;
          acc += v3;
          idx = i3;
        }
        for (; k < terms; ++k) {
          idx += p;
          if (idx >= P) idx -= P;
          acc += src[idx];
        }
      }
    }
    double div;
    if (trunc) {
      div = (double)M;  // np.mean over the complete rows (Periods.py:178-184)
    } else {
      // zero padding of the last row takes part in np.sum (only visible on -0.0)
      if (r0 != 0 && r >= r0) acc += 0.0;
      div = (double)(M + (r < r0 ? 1 : 0));
    }
    vout[r] = (trunc && M == 0) ? __longlong_as_double(0x7ff8000000000000LL) : acc / div;
  }
}

// Exact project(): fold/mean, then the Muresan-Parks prime-cofactor chain (Periods.py:208-214):
// for each cofactor q (host-ordered): v[r] -= mean-fold_q(tile(v))[r mod q].  vout, utmp >= p doubles.
// Contains CTA barriers; all threads must call.  On return vout[0..p) is valid for all threads.
template <bool WRAP>
__device__ __forceinline__ void cta_project_exact(const double* src, int P, int N, int p, bool trunc,
                                                  const int32_t* __restrict__ chain, int chain_len,
                                                  double* __restrict__ vout, double* __restrict__ utmp) {
  cta_fold_mean_exact<WRAP>(src, P, N, p, trunc, vout);
  __syncthreads();
  for (int ci = 0; ci < chain_len; ++ci) {
    const int q = chain[ci];
    cta_fold_mean_exact<true>(vout, p, N, q, trunc, utmp);
    __syncthreads();
    int rq = threadIdx.x % q;
    const int step = kThreads % q;
    for (int r = threadIdx.x; r < p; r += kThreads) {
      vout[r] = vout[r] - utmp[rq];
      rq += step;
      if (rq >= q) rq -= q;
    }
    __syncthreads();
  }
}

// xs[n] -= v[n mod p] for n < N (Periods.py:537, 283, 340).  No barrier inside.
__device__ __forceinline__ void cta_subtract_tiled(double* xs, int N, const double* v, int p) {
  int r = threadIdx.x % p;
  const int step = kThreads % p;
  for (int n = threadIdx.x; n < N; n += kThreads) {
    xs[n] = xs[n] - v[r];
    r += step;
    if (r >= p) r -= p;
  }
}

// out[n] = v[n mod p] for n < len, coalesced streaming stores (bases are write-once): 128-bit stores of sample
// pairs when the destination is 16-byte aligned, 64-bit stores otherwise.
__device__ __forceinline__ void cta_store_tiled(double* __restrict__ out, int len, const double* v, int p) {
  if ((reinterpret_cast<uintptr_t>(out) & 15u) == 0 && p >= 2) {
    const int pairs = len >> 1;
    int r = (2 * threadIdx.x) % p;
    const int step = (2 * kThreads) % p;
    double2* out2 = reinterpret_cast<double2*>(out);
    for (int i = threadIdx.x; i < pairs; i += kThreads) {
      const int r1 = (r + 1 == p) ? 0 : r + 1;
      __stcs(out2 + i, make_double2(v[r], v[r1]));
      r += step;
      if (r >= p) r -= p;
    }
    if ((len & 1) && threadIdx.x == 0) __stcs(out + len - 1, v[(len - 1) % p]);
    return;
  }
  int r = threadIdx.x % p;
  const int step = kThreads % p;
  for (int n = threadIdx.x; n < len; n += kThreads) {
    __stcs(out + n, v[r]);
    r += step;
    if (r >= p) r -= p;
  }
}

// ------------------------------------------------------------------------------------------
// approximate (non-sequential) helpers used only to RANK candidates; <= a few ulp from exact
// ------------------------------------------------------------------------------------------
// In-place cofactor chain on a P-periodic vector v (warp-private, P doubles), using
// multiplicities instead of N sequential adds.  trunc: only the first floor(N/q)*q samples.
__device__ __forceinline__ void warp_orth_chain_approx(double* v, int P, int N, bool trunc,
                                                       const int32_t* __restrict__ chain, int chain_len) {
  const int lane = threadIdx.x & 31;
  for (int ci = 0; ci < chain_len; ++ci) {
    const int q = chain[ci];
    const int t = P / q;
    const int Mq = N / q, r0q = N - Mq * q;
    for (int r = lane; r < q; r += 32) {
      const int K = trunc ? Mq : (Mq + (r < r0q ? 1 : 0));
      const int base = K / t, rem = K - base * t;
      double s = 0.0;
      for (int j = 0; j < t; ++j) s = fma((double)(base + (j < rem ? 1 : 0)), v[r + j * q], s);
      const double u = s / (double)K;
      for (int j = 0; j < t; ++j) v[r + j * q] -= u;
    }
    __syncwarp();
  }
}

}  // namespace pp
