// pyperiod_b200 -- QOPeriods.get_periods on the device, batched (pyPeriod/QOPeriods.py:719-741, 854-938).
//
// Per window (one CTA): the weights of the quadratic program are laid out as one zero-padded vector of sum(q)
// entries (concatenate_periods, :854-887); every pair of periods (a, b) with g = gcd(a, b) contributes g rows
// "+1 on the comb {s, s+g, ...} of b's segment, -1 on the same comb of a's segment" (stack_pairwise_gcd_subspaces,
// :889-938; np.roll by s < g never leaves a segment because g divides both lengths); the result is the vector minus
// its projection onto the row space of that matrix, cut back into one waveform per period (:736-741).  The
// reference first drops linearly dependent rows (reduce_rows, :86-94) or factors the matrix ("lu", "qr") or calls
// lstsq: the row space, hence the projection, is the same in every branch.  Here the projection is computed by
// conjugate gradients on the normal equations with the rows applied implicitly (pp_cg.cuh): no matrix, no rank
// decisions, no host round trip.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pyperiod_b200.h"
#include "pp_cg.cuh"
#include "pp_common.cuh"
#include "pp_host.cuh"

namespace pp {

struct GpPlan {
  int kmax, npairs, tmax, rmax;
  __host__ __device__ size_t off_u() const { return (size_t)tmax * 8; }
  __host__ __device__ size_t off_r() const { return off_u() + (size_t)tmax * 8; }
  __host__ __device__ size_t off_p() const { return off_r() + (size_t)rmax * 8; }
  __host__ __device__ size_t off_ap() const { return off_p() + (size_t)rmax * 8; }
  __host__ __device__ size_t off_red() const { return off_ap() + (size_t)rmax * 8; }
  __host__ __device__ size_t off_ints() const { return off_red() + 2 * kWarps * 8; }
  // ints: q[kmax] o[kmax+1] g[npairs] rho[npairs+1] misc[4]
  __host__ __device__ size_t bytes() const { return off_ints() + (size_t)(2 * kmax + 2 * npairs + 6) * 4 + 16; }
};

__host__ __device__ inline GpPlan make_gp_plan(int kmax, int tmax, int rmax) {
  GpPlan pl;
  pl.kmax = kmax;
  pl.npairs = kmax * (kmax - 1) / 2;
  pl.tmax = (tmax + 1) & ~1;
  pl.rmax = (rmax + 1) & ~1;
  return pl;
}

__host__ __device__ inline int gcd_int(int a, int b) {
  while (b) {
    const int t = a % b;
    a = b;
    b = t;
  }
  return a;
}

// the pairwise-GCD rows, applied implicitly
struct GpOp {
  int n, npairs, R, T;
  const int* q;     // period of segment k
  const int* o;     // first element of segment k, [n + 1]
  const int* g;     // gcd of pair pi (itertools.combinations order: (0,1), (0,2), ..., (1,2), ...)
  const int* rho;   // first row of pair pi, [npairs + 1]
  __device__ __forceinline__ int pair_index(int ka, int kb) const { return ka * n - ka * (ka + 1) / 2 + (kb - ka - 1); }

  // out[rho + s] = sum_{j = s (mod g)} u_b[j] - sum_{j = s (mod g)} u_a[j]: one warp per row
  __device__ void apply(const double* u, double* out) const {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int pi = 0, ka = 0, kb = 1;
    for (int row = wid; row < R; row += kWarps) {
      while (rho[pi + 1] <= row) {   // rows ascend: advance the pair cursor
        ++pi;
        if (++kb == n) {
          ++ka;
          kb = ka + 1;
        }
      }
      const int gg = g[pi], s = row - rho[pi];
      double acc = 0.0;
      const double* ub = u + o[kb];
      for (int j = s + gg * lane; j < q[kb]; j += gg * 32) acc += ub[j];
      const double* ua = u + o[ka];
      for (int j = s + gg * lane; j < q[ka]; j += gg * 32) acc -= ua[j];
      acc = warp_sum(acc);
      if (lane == 0) out[row] = acc;
    }
  }

  // out[o_k + j] = sum over pairs holding k of (+1 if k is the pair's second period, -1 if its first) * v[rho + j mod g]
  __device__ void apply_t(const double* v, double* out) const {
    for (int e = threadIdx.x; e < T; e += kThreads) {
      int k = 0;
      while (o[k + 1] <= e) ++k;
      const int j = e - o[k];
      double acc = 0.0;
      for (int k2 = 0; k2 < n; ++k2) {
        if (k2 == k) continue;
        const int pi = k2 < k ? pair_index(k2, k) : pair_index(k, k2);
        const double val = v[rho[pi] + j % g[pi]];
        acc += k2 < k ? val : -val;
      }
      out[e] = acc;
    }
  }
};

__global__ void __launch_bounds__(kThreads)
gp_kernel(int B, int kmax, int tmax, int rmax, const int32_t* __restrict__ dict_q, const int32_t* __restrict__ dict_keep,
          const int32_t* __restrict__ n_dict, const double* __restrict__ weights, int64_t ldw,
          const int64_t* __restrict__ weights_off, double* __restrict__ out, const int64_t* __restrict__ out_off,
          int32_t* __restrict__ iters, int32_t* __restrict__ status) {
  unsigned char* smem = pp_smem;
  const GpPlan pl = make_gp_plan(kmax, tmax, rmax);
  double* act = reinterpret_cast<double*>(smem);
  double* u = reinterpret_cast<double*>(smem + pl.off_u());
  double* r = reinterpret_cast<double*>(smem + pl.off_r());
  double* p = reinterpret_cast<double*>(smem + pl.off_p());
  double* ap = reinterpret_cast<double*>(smem + pl.off_ap());
  double* red = reinterpret_cast<double*>(smem + pl.off_red());
  int* q = reinterpret_cast<int*>(smem + pl.off_ints());
  int* o = q + kmax;
  int* g = o + kmax + 1;
  int* rho = g + pl.npairs;
  const int tid = threadIdx.x;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    const int n = min(max(n_dict[b], 0), kmax);
    if (tid == 0) {
      int off = 0;
      for (int k = 0; k < n; ++k) {
        q[k] = dict_q[(size_t)b * kmax + k];
        o[k] = off;
        off += q[k];
      }
      o[n] = off;
      int row = 0, pi = 0;
      for (int ka = 0; ka < n; ++ka)
        for (int kb = ka + 1; kb < n; ++kb) {
          g[pi] = gcd_int(q[ka], q[kb]);
          rho[pi] = row;
          row += g[pi];
          ++pi;
        }
      rho[pi] = row;
    }
    __syncthreads();
    const int T = o[n], np = n * (n - 1) / 2, R = np > 0 ? rho[np] : 0;
    double* dst = out + out_off[b];
    if (T > pl.tmax || R > pl.rmax) {
      if (tid == 0) {
        status[b] = PP_STATUS_TOO_LARGE;
        iters[b] = 0;
      }
      continue;
    }
    // concatenate_periods (:854-887): the first keep_k weights of each period, zero padded to q_k; a repeated
    // period (keep 0) takes no weights and does not advance the read position, as in the reference
    const double* wsrc = weights + (weights_off ? weights_off[b] : (int64_t)b * ldw);
    for (int e = tid; e < T; e += kThreads) act[e] = 0.0;
    __syncthreads();
    {
      int pos = 0;
      for (int k = 0; k < n; ++k) {
        const int keep = min(max(dict_keep[(size_t)b * kmax + k], 0), q[k]);
        for (int i = tid; i < keep; i += kThreads) act[o[k] + i] = wsrc[pos + i];
        pos += keep;
      }
    }
    __syncthreads();
    int it = 0;
    if (n == 1) {
      // a single period: the matrix is one row of ones (:935-936), the projection is the mean
      double s = 0.0;
      for (int e = tid; e < T; e += kThreads) s += act[e];
      s = warp_sum(s);
      if ((tid & 31) == 0) red[tid >> 5] = s;
      __syncthreads();
      double tot = 0.0;
      for (int w = 0; w < kWarps; ++w) tot += red[w];
      const double mean = tot / (double)T;
      for (int e = tid; e < T; e += kThreads) act[e] -= mean;
    } else if (n > 1) {
      GpOp op{n, np, R, T, q, o, g, rho};
      it = cta_cg_project(op, R, T, act, u, r, p, ap, nullptr, red, R + 16);
    }
    __syncthreads();
    for (int e = tid; e < T; e += kThreads) dst[e] = act[e];
    if (tid == 0) {
      iters[b] = it;
      status[b] = it < 0 ? PP_STATUS_GUARD : PP_STATUS_OK;
    }
  }
}

}  // namespace pp

using namespace pp;

extern "C" {

int pp_qo_get_periods(int32_t B, int32_t kmax, int32_t tmax, int32_t rmax, const int32_t* dict_q,
                      const int32_t* dict_keep, const int32_t* n_dict, const double* weights, int64_t ldw,
                      const int64_t* weights_off, double* out, const int64_t* out_off, int32_t* iters,
                      int32_t* status, void* stream) {
  if (B == 0) return 0;
  if (B < 0 || kmax < 1 || kmax > 64 || tmax < 1 || rmax < 0) return fail(-1, "bad arguments (1 <= kmax <= 64)%s");
  if (!dict_q || !dict_keep || !n_dict || !weights || !out || !out_off || !iters || !status)
    return fail(-1, "pointers are null%s");
  DeviceFacts f;
  if (int rc = device_facts(f)) return rc;
  const GpPlan pl = make_gp_plan(kmax, tmax, rmax < 2 ? 2 : rmax);
  if (pl.bytes() > (size_t)f.smem_optin)
    return fail(-2, "periods too long for on-chip extraction (2 * sum(q) + 3 * sum(gcd) doubles must fit in shared memory)%s");
  if (int rc = prep_kernel(gp_kernel, pl.bytes(), f)) return rc;
  const int grid = grid_for(f, pl.bytes(), B, 4);
  gp_kernel<<<grid, kThreads, pl.bytes(), (cudaStream_t)stream>>>(B, kmax, pl.tmax, pl.rmax, dict_q, dict_keep, n_dict,
                                                                  weights, ldw, weights_off, out, out_off, iters, status);
  return check_cuda(cudaGetLastError(), "gp_kernel launch");
}

}  // extern "C"
