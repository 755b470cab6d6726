// pyperiod_b200 -- Ramanujan periodogram (pyPeriod/RamanujanPeriods.py:67-86, 124-169) on the B200.
//
// Reference, per period q: dictionary of q rows (c_q rolled by i, tiled to N), each rescaled by its
// maximum phi(q); projection  output = sum_i <x, r_i> r_i ;  norms[q] = sum_n output[n]^2.
// With S_q the residue-class fold of the window and C[i][m] = c_q((m - i) mod q) / phi(q):
//     y = C S_q,  z = C^T y,  norms[q] = sum_m cnt_q[m] z_m^2        
// and since the circular autocorrelation of a Ramanujan sum is q c_q, C^T C = (q / phi^2) circ(c_q):
//     z = (q / phi(q)^2) * H S_q,   H[m][l] = c_q((l - m) mod q)      (integer matrix).
//
// Three kernels:
//   cq_kernel        the dictionary: c_q(n) = mu(q/g) phi(q) / phi(q/g), g = gcd(n, q), exact integers
//   fold_all_kernel  S_q for every q of every window of a tile (4 windows per CTA staged in shared memory)
//   ram_gemm_kernel  the dense contraction  H (q x q) * S_q (q x Bt)  on the FP64 tensor cores
//                    (mma.sync m8n8k4 f64 = DMMA) with the cnt-weighted column norms fused in the epilogue.
// H is never materialised: its fragments are generated from the q-vector c_q held in shared memory.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pyperiod_b200.h"
#include "pp_common.cuh"
#include "pp_host.cuh"

namespace pp {

// ------------------------------------------------------------------------------------------
// dictionary: c_q(n) for all q in [0, qmax], concatenated at offset q (q - 1) / 2 (triangular layout)
// ------------------------------------------------------------------------------------------
__host__ __device__ inline size_t cq_offset(int q) { return (size_t)q * (q - 1) / 2; }

__global__ void cq_kernel(int qmin, int qmax, const int32_t* __restrict__ mu, const int32_t* __restrict__ phi,
                          double* __restrict__ cq) {
  const int q = qmin + blockIdx.x;
  if (q > qmax) return;
  double* out = cq + cq_offset(q);
  for (int n = threadIdx.x; n < q; n += blockDim.x) {
    int a = n, b = q;  // gcd(n, q), gcd(0, q) = q
    while (a) {
      const int t = b % a;
      b = a;
      a = t;
    }
    const int g = b, d = q / g;
    out[n] = (double)(mu[d] * (phi[q] / phi[d]));
  }
}

// ------------------------------------------------------------------------------------------
// fold of every period for a tile of windows: S[soff(q) + m * ldS + b]
// ------------------------------------------------------------------------------------------
constexpr int kFoldWin = 4;  // windows per CTA: one 32-byte sector of S per (q, m)

__host__ __device__ inline size_t s_offset(int q, int qmin, size_t ldS) {
  // rows of all periods below q: sum_{t=qmin}^{q-1} t
  return ((size_t)q * (q - 1) / 2 - (size_t)qmin * (qmin - 1) / 2) * ldS;
}

__global__ void __launch_bounds__(kThreads, 1)
fold_all_kernel(const double* __restrict__ x, int64_t ldx, int b_first, int b_count, int N, int qmin, int qmax,
                double* __restrict__ S, int ldS) {
  double* xs = reinterpret_cast<double*>(pp_smem);  // [kFoldWin][nstride]
  const int nstride = (N + 1) & ~1;
  __shared__ int counter;
  const int lane = threadIdx.x & 31;
  for (int g0 = blockIdx.x * kFoldWin; g0 < b_count; g0 += gridDim.x * kFoldWin) {
    __syncthreads();
    for (int w = 0; w < kFoldWin; ++w) {
      const bool live = g0 + w < b_count;
      const double* src = x + (size_t)(b_first + g0 + w) * ldx;
      for (int n = threadIdx.x; n < N; n += kThreads) xs[w * nstride + n] = live ? __ldg(src + n) : 0.0;
    }
    if (threadIdx.x == 0) counter = 0;
    __syncthreads();
    while (true) {
      int idx = 0;
      if (lane == 0) idx = atomicAdd(&counter, 1);
      idx = __shfl_sync(0xffffffffu, idx, 0);
      const int q = qmax - idx;  // large periods first
      if (q < qmin) break;
      double* out = S + s_offset(q, qmin, ldS) + g0;
      for (int m = lane; m < q; m += 32) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        for (int n = m; n < N; n += q) {  // terms in increasing n, as the reference's np.dot would see them folded
          a0 += xs[n];
          a1 += xs[nstride + n];
          a2 += xs[2 * nstride + n];
          a3 += xs[3 * nstride + n];
        }
        double* o = out + (size_t)m * ldS;
        if (g0 + 3 < b_count) {
          *reinterpret_cast<double2*>(o) = make_double2(a0, a1);
          *reinterpret_cast<double2*>(o + 2) = make_double2(a2, a3);
        } else {
          o[0] = a0;
          if (g0 + 1 < b_count) o[1] = a1;
          if (g0 + 2 < b_count) o[2] = a2;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// dense contraction on the FP64 tensor cores
// ------------------------------------------------------------------------------------------
constexpr int kGemmM = 128;   // rows of H per CTA pass
constexpr int kGemmN = 64;    // windows per CTA
constexpr int kGemmK = 32;    // k-chunk staged in shared memory
constexpr int kLdB = 68;      // padded leading dimension of the staged S chunk (68 mod 16 = 4: conflict-free frags)

// 16-byte asynchronous global -> shared copy; bytes past src_bytes are zero-filled (rows past q, windows past
// the tile)
__device__ __forceinline__ void cp_async_16(void* dst_smem, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void dmma_m8n8k4(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// grid: (ceil(Bt / 64), number of periods).  norms[(b_first + b) * ld_norms + q] = sum_m cnt_q[m] z_m^2.
__global__ void __launch_bounds__(kThreads, 2)
ram_gemm_kernel(const double* __restrict__ S, int ldS, int b_first, int b_count, int N, int qmin, int qmax,
                const double* __restrict__ cq_all, const int32_t* __restrict__ phi, double* __restrict__ norms,
                int ld_norms) {
  const int q = qmax - blockIdx.y;  // big periods first
  if (q < qmin) return;
  const int b0 = blockIdx.x * kGemmN;
  if (b0 >= b_count) return;
  double* cqs = reinterpret_cast<double*>(pp_smem);              // [q]
  double* Bs0 = cqs + ((q + 1) & ~1);                              // [2][kGemmK][kLdB]: double-buffered S chunk
  double* red = Bs0 + 2 * kGemmK * kLdB;                           // [4][kGemmN]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int wm = wid >> 1, wn = wid & 1;                           // warp tile: rows wm*32.., cols wn*32..
  const int lr = lane >> 2, lc = lane & 3;
  const double* cq = cq_all + cq_offset(q);
  for (int i = tid; i < q; i += kThreads) cqs[i] = cq[i];
  const double* Sq = S + s_offset(q, qmin, ldS) + b0;
  const int Mrows = N / q, r0 = N - Mrows * q;

  double colacc[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j) colacc[j][0] = colacc[j][1] = 0.0;

  for (int m0 = 0; m0 < q; m0 += kGemmM) {
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    // stage S_q[l0 .. l0+31][b0 .. b0+63] with cp.async (rows past q and windows past the tile are zero-filled);
    // the copy of chunk l0 + 32 is in flight while chunk l0 feeds the tensor cores
    auto stage = [&](double* dst, int l0) {
      for (int idx = tid; idx < kGemmK * kGemmN / 2; idx += kThreads) {
        const int kk = idx >> 5, n = (idx & 31) * 2;
        int bytes = 0;
        if (l0 + kk < q) bytes = min(max(b_count - (b0 + n), 0), 2) * 8;
        const double* src = bytes ? Sq + (size_t)(l0 + kk) * ldS + n : Sq;
        cp_async_16(dst + kk * kLdB + n, src, bytes);
      }
      cp_async_commit();
    };
    __syncthreads();  // the previous row block is done with both buffers
    stage(Bs0, 0);
    int buf = 0;
    for (int l0 = 0; l0 < q; l0 += kGemmK, buf ^= 1) {
      const double* Bs = Bs0 + buf * (kGemmK * kLdB);
      if (l0 + kGemmK < q) {
        stage(Bs0 + (buf ^ 1) * (kGemmK * kLdB), l0 + kGemmK);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
#pragma unroll
      for (int k4 = 0; k4 < kGemmK / 4; ++k4) {
        const int l = l0 + k4 * 4 + lc;
        double bf[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) bf[j] = Bs[(k4 * 4 + lc) * kLdB + wn * 32 + j * 8 + lr];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int m = m0 + wm * 32 + i * 8 + lr;
          double a = 0.0;
          if (m < q && l < q) {
            int d = l - m;
            if (d < 0) d += q;
            a = cqs[d];
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j], a, bf[j]);
        }
      }
      __syncthreads();  // everyone is done with this buffer before the next iteration restages it
    }
    // epilogue of this row block: cnt-weighted squares, accumulated per column
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + wm * 32 + i * 8 + lr;
      const double cnt = (m < q) ? (double)(Mrows + (m < r0 ? 1 : 0)) : 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        colacc[j][0] = fma(cnt * acc[i][j][0], acc[i][j][0], colacc[j][0]);
        colacc[j][1] = fma(cnt * acc[i][j][1], acc[i][j][1], colacc[j][1]);
      }
    }
  }
  // reduce over the 8 row-lanes of a column pair, then over the 4 row-warps (fixed order)
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      double v = colacc[j][e];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      colacc[j][e] = v;
    }
  __syncthreads();
  if (lr == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      red[wm * kGemmN + wn * 32 + j * 8 + 2 * lc] = colacc[j][0];
      red[wm * kGemmN + wn * 32 + j * 8 + 2 * lc + 1] = colacc[j][1];
    }
  }
  __syncthreads();
  if (tid < kGemmN && b0 + tid < b_count) {
    const double ph = (double)phi[q];
    const double scale = (double)q / (ph * ph);  // C^T C = (q / phi^2) circ(c_q)
    const double t = ((red[tid] + red[kGemmN + tid]) + red[2 * kGemmN + tid]) + red[3 * kGemmN + tid];
    norms[(size_t)(b_first + b0 + tid) * ld_norms + q] = scale * scale * t;
  }
}

// ------------------------------------------------------------------------------------------
// TF32 option: the same contraction on the 19-bit tensor cores (mma.sync m16n8k8, fp32 accumulate).
// c_q(n) is an integer with |c_q| <= phi(q) < 2^11, exact in TF32; the fold sums are split into
// hi + lo TF32 parts (two MMAs per tile), which leaves ~2^-21 relative error per product and the fp32
// accumulation of q terms: norms agree with the fp64 path to ~1e-5 relative.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void mma_tf32_m16n8k8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(kThreads, 2)
ram_gemm_tf32_kernel(const double* __restrict__ S, int ldS, int b_first, int b_count, int N, int qmin, int qmax,
                     const double* __restrict__ cq_all, const int32_t* __restrict__ phi, double* __restrict__ norms,
                     int ld_norms) {
  const int q = qmax - blockIdx.y;  // big periods first
  if (q < qmin) return;
  const int b0 = blockIdx.x * kGemmN;
  if (b0 >= b_count) return;
  float* cqs = reinterpret_cast<float*>(pp_smem);                       // [q] (TF32-exact integers)
  double* Bs0 = reinterpret_cast<double*>(pp_smem) + ((q + 3) / 2 & ~1);  // [2][kGemmK][kLdB] fp64 fold chunk
  double* red = Bs0 + 2 * kGemmK * kLdB;                                // [4][kGemmN]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int wm = wid >> 1, wn = wid & 1;   // warp tile: rows wm*32.., cols wn*32..
  const int gid = lane >> 2, tig = lane & 3;
  const double* cq = cq_all + cq_offset(q);
  for (int i = tid; i < q; i += kThreads) cqs[i] = (float)cq[i];
  const double* Sq = S + s_offset(q, qmin, ldS) + b0;
  const int Mrows = N / q, r0 = N - Mrows * q;

  double colacc[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j) colacc[j][0] = colacc[j][1] = 0.0;

  for (int m0 = 0; m0 < q; m0 += kGemmM) {
    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;
    auto stage = [&](double* dst, int l0) {
      for (int idx = tid; idx < kGemmK * kGemmN / 2; idx += kThreads) {
        const int kk = idx >> 5, n = (idx & 31) * 2;
        int bytes = 0;
        if (l0 + kk < q) bytes = min(max(b_count - (b0 + n), 0), 2) * 8;
        const double* src = bytes ? Sq + (size_t)(l0 + kk) * ldS + n : Sq;
        cp_async_16(dst + kk * kLdB + n, src, bytes);
      }
      cp_async_commit();
    };
    __syncthreads();
    stage(Bs0, 0);
    int buf = 0;
    for (int l0 = 0; l0 < q; l0 += kGemmK, buf ^= 1) {
      const double* Bs = Bs0 + buf * (kGemmK * kLdB);
      if (l0 + kGemmK < q) {
        stage(Bs0 + (buf ^ 1) * (kGemmK * kLdB), l0 + kGemmK);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
#pragma unroll
      for (int k8 = 0; k8 < kGemmK / 8; ++k8) {
        // B fragments (k x n = 8 x 8 per tile j): b0 = (k = tig, n = gid), b1 = (k = tig + 4, n = gid); hi + lo split
        uint32_t bh[4][2], bl[4][2];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const double v = Bs[(k8 * 8 + tig + 4 * h) * kLdB + wn * 32 + j * 8 + gid];
            const float f = (float)v;
            const uint32_t hi = to_tf32(f);
            bh[j][h] = hi;
            bl[j][h] = to_tf32((float)(v - (double)__uint_as_float(hi)));
          }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          // A fragment (m x k = 16 x 8): a0 = (gid, tig), a1 = (gid + 8, tig), a2 = (gid, tig + 4), a3 = (gid + 8, tig + 4)
          uint32_t a[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int m = m0 + wm * 32 + i * 16 + gid + 8 * (e & 1);
            const int l = l0 + k8 * 8 + tig + 4 * (e >> 1);
            float v = 0.f;
            if (m < q && l < q) {
              int d = l - m;
              if (d < 0) d += q;
              v = cqs[d];
            }
            a[e] = __float_as_uint(v);  // small integers: already TF32-exact
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            mma_tf32_m16n8k8(acc[i][j], a, bh[j][0], bh[j][1]);
            mma_tf32_m16n8k8(acc[i][j], a, bl[j][0], bl[j][1]);
          }
        }
      }
      __syncthreads();
    }
    // epilogue: c0,c1 = (row gid, cols 2 tig, 2 tig + 1), c2,c3 = (row gid + 8, same cols)
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int hrow = 0; hrow < 2; ++hrow) {
        const int m = m0 + wm * 32 + i * 16 + gid + 8 * hrow;
        const double cnt = (m < q) ? (double)(Mrows + (m < r0 ? 1 : 0)) : 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double z0 = (double)acc[i][j][2 * hrow], z1 = (double)acc[i][j][2 * hrow + 1];
          colacc[j][0] = fma(cnt * z0, z0, colacc[j][0]);
          colacc[j][1] = fma(cnt * z1, z1, colacc[j][1]);
        }
      }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      double v = colacc[j][e];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      colacc[j][e] = v;
    }
  __syncthreads();
  if (gid == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      red[wm * kGemmN + wn * 32 + j * 8 + 2 * tig] = colacc[j][0];
      red[wm * kGemmN + wn * 32 + j * 8 + 2 * tig + 1] = colacc[j][1];
    }
  }
  __syncthreads();
  if (tid < kGemmN && b0 + tid < b_count) {
    const double ph = (double)phi[q];
    const double scale = (double)q / (ph * ph);
    const double t = ((red[tid] + red[kGemmN + tid]) + red[2 * kGemmN + tid]) + red[3 * kGemmN + tid];
    norms[(size_t)(b_first + b0 + tid) * ld_norms + q] = scale * scale * t;
  }
}

// periods whose norm exceeds thresh * |max norm|, ascending (RamanujanPeriods.py:97-101); one warp per window
__global__ void select_kernel(const double* __restrict__ norms, int B, int ld_norms, int qlen, double thresh, int kmax,
                              int32_t* __restrict__ periods, int32_t* __restrict__ nper) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const double* nr = norms + (size_t)b * ld_norms;
  double mx = -1.0 / 0.0;
  for (int q = lane; q < qlen; q += 32) mx = fmax(mx, nr[q]);
  mx = warp_max(mx);
  const double den = fabs(mx);
  int count = 0;
  for (int q0 = 0; q0 < qlen; q0 += 32) {
    const int q = q0 + lane;
    const bool hit = q < qlen && (nr[q] / den > thresh);
    const unsigned mask = __ballot_sync(0xffffffffu, hit);
    if (hit) {
      const int pos = count + __popc(mask & ((1u << lane) - 1u));
      if (pos < kmax) periods[(size_t)b * kmax + pos] = q;
    }
    count += __popc(mask);
  }
  if (lane == 0) nper[b] = count;
}

}  // namespace pp

using namespace pp;

extern "C" {

size_t pp_ramanujan_workspace_bytes(int32_t N, int32_t qmin, int32_t qmax, int32_t tile_windows) {
  (void)N;
  const size_t ldS = ((size_t)tile_windows + 3) & ~(size_t)3;
  const size_t rows = (size_t)qmax * (qmax + 1) / 2 - (size_t)qmin * (qmin - 1) / 2;
  return 4096 + (cq_offset(qmax + 1) + 2) * 8 + rows * ldS * 8;
}

// norms[b, q] for q in [qmin, qmax] (other entries untouched; the caller zero-fills, RamanujanPeriods.py:71).
// mu / phi: device int32 tables for 0..table_qmax.  Windows are processed in tiles of `tile_windows`
// (workspace holds the folds of one tile).
static int ramanujan_norms_impl(const double* x, int64_t ldx, int32_t B, int32_t N, int32_t qmin, int32_t qmax,
                                const int32_t* mu, const int32_t* phi, int32_t table_qmax, int32_t tile_windows,
                                double* norms, int32_t ld_norms, void* workspace, size_t workspace_bytes, void* stream,
                                bool tf32) {
  if (B == 0) return 0;  // empty batch: nothing to validate or launch
  if (x == nullptr || norms == nullptr || B < 0 || N < 2 || ldx < 1) return fail(-1, "bad window arguments%s");
  if (qmin < 1 || qmax < qmin || qmax > N) return fail(-1, "need 1 <= qmin <= qmax <= N%s");
  if (mu == nullptr || phi == nullptr || table_qmax < qmax) return fail(-1, "mu/phi tables must cover qmax%s");
  if (ld_norms < qmax + 1 || tile_windows < 1) return fail(-1, "bad norms leading dimension or tile%s");
  if (B == 0) return 0;
  DeviceFacts f;
  if (int rc = device_facts(f)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int ldS = (tile_windows + 3) & ~3;
  const size_t rows = (size_t)qmax * (qmax + 1) / 2 - (size_t)qmin * (qmin - 1) / 2;
  size_t off = 0;
  double* cq = carve(workspace, workspace_bytes, off, (cq_offset(qmax + 1) + 2) * 8);
  double* S = carve(workspace, workspace_bytes, off, rows * (size_t)ldS * 8);
  if (!cq || !S) return fail(-3, "workspace too small (see pp_ramanujan_workspace_bytes)%s");
  cq_kernel<<<qmax - qmin + 1, 128, 0, st>>>(qmin, qmax, mu, phi, cq);
  const size_t fold_smem = (size_t)kFoldWin * ((N + 1) & ~1) * 8;
  if (int rc = prep_kernel(fold_all_kernel, fold_smem, f)) return rc;
  const size_t gemm_smem = (size_t)(((qmax + 1) & ~1) + 2 * kGemmK * kLdB + 4 * kGemmN) * 8;
  if (int rc = prep_kernel(ram_gemm_kernel, gemm_smem, f)) return rc;
  if (int rc = prep_kernel(ram_gemm_tf32_kernel, gemm_smem, f)) return rc;
  for (int b_first = 0; b_first < B; b_first += tile_windows) {
    const int b_count = (B - b_first < tile_windows) ? (B - b_first) : tile_windows;
    int fgrid = (b_count + kFoldWin - 1) / kFoldWin;
    if (fgrid > f.sm_count) fgrid = f.sm_count;
    fold_all_kernel<<<fgrid, kThreads, fold_smem, st>>>(x, ldx, b_first, b_count, N, qmin, qmax, S, ldS);
    dim3 grid((b_count + kGemmN - 1) / kGemmN, qmax - qmin + 1);
    if (tf32)
      ram_gemm_tf32_kernel<<<grid, kThreads, gemm_smem, st>>>(S, ldS, b_first, b_count, N, qmin, qmax, cq, phi, norms,
                                                              ld_norms);
    else
      ram_gemm_kernel<<<grid, kThreads, gemm_smem, st>>>(S, ldS, b_first, b_count, N, qmin, qmax, cq, phi, norms,
                                                         ld_norms);
  }
  return check_cuda(cudaGetLastError(), "ramanujan kernels launch");
}

int pp_ramanujan_norms(const double* x, int64_t ldx, int32_t B, int32_t N, int32_t qmin, int32_t qmax,
                       const int32_t* mu, const int32_t* phi, int32_t table_qmax, int32_t tile_windows,
                       double* norms, int32_t ld_norms, void* workspace, size_t workspace_bytes, void* stream) {
  return ramanujan_norms_impl(x, ldx, B, N, qmin, qmax, mu, phi, table_qmax, tile_windows, norms, ld_norms, workspace,
                              workspace_bytes, stream, false);
}

int pp_ramanujan_norms_tf32(const double* x, int64_t ldx, int32_t B, int32_t N, int32_t qmin, int32_t qmax,
                            const int32_t* mu, const int32_t* phi, int32_t table_qmax, int32_t tile_windows,
                            double* norms, int32_t ld_norms, void* workspace, size_t workspace_bytes, void* stream) {
  return ramanujan_norms_impl(x, ldx, B, N, qmin, qmax, mu, phi, table_qmax, tile_windows, norms, ld_norms, workspace,
                              workspace_bytes, stream, true);
}

// periods[b, 0:nper[b]] = ascending q in [0, qlen) with norms[b, q] / |max_q norms[b, q]| > thresh
// (RamanujanPeriods.py:97-101).  nper[b] may exceed kmax; only the first kmax are stored.
int pp_ramanujan_select(const double* norms, int32_t B, int32_t ld_norms, int32_t qlen, double thresh, int32_t kmax,
                        int32_t* periods, int32_t* nper, void* stream) {
  if (B == 0) return 0;  // empty batch: nothing to validate or launch
  if (!norms || !periods || !nper || B < 0 || qlen < 1 || ld_norms < qlen || kmax < 1)
    return fail(-1, "bad select arguments%s");
  if (B == 0) return 0;
  select_kernel<<<(B + 7) / 8, 256, 0, (cudaStream_t)stream>>>(norms, B, ld_norms, qlen, thresh, kmax, periods, nper);
  return check_cuda(cudaGetLastError(), "select_kernel launch");
}

}  // extern "C"
