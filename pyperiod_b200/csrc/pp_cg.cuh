// pyperiod_b200 -- orthogonal projection onto the row space of an implicit dictionary by conjugate gradients
// on the normal equations, one CTA per window.
//
// Two places of the reference solve (A A^T) w = A t with a dictionary A whose Gram matrix is SINGULAR but whose
// system is consistent, and use only the reconstruction A^T w (the projection of t onto the row space of A,
// which does not depend on which solution w is taken):
//   * QOPeriods.get_periods (pyPeriod/QOPeriods.py:719-741): A = +-comb rows of every pair's gcd, all shifts
//     (:889-938); the reference first drops dependent rows (reduce_rows, :86-94) -- same row space;
//   * QOPeriods(basis_type="ramanujan") (QOPeriods.py:970-971, 1005-1052): A = shifted Ramanujan sums, q rows of
//     a period whose subspace has dimension phi(q); np.linalg.solve runs on the rounding-perturbed singular matrix
//     and its reconstruction equals the projection to 1e-14 (measured on the reference).
// CG needs only the products A u and A^T v, which are folds and tilings here; it converges on a consistent
// semidefinite system to the minimum-norm solution, in as many steps as G has distinct eigenvalues (few: the rows
// are shifts of a handful of periodic patterns).  The projection is accumulated directly in the target space
// (t - A^T w is updated by alpha * A^T p each step), so the result does not depend on w at all.
#pragma once

#include <cuda_runtime.h>

#include "pp_common.cuh"

namespace pp {

// CTA-wide dot product of two global/shared vectors, fixed order.  red: 2 * kWarps doubles of shared scratch.
__device__ __forceinline__ double cta_dot(const double* a, const double* b, int n, double* red) {
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += kThreads) s = fma(a[i], b[i], s);
  s = warp_sum(s);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) t += red[w];
  return t;
}

// Op interface (all threads call, may contain barriers; results visible after the call returns):
//   void apply(const double* u /*T*/, double* out /*R*/);     out = A u
//   void apply_t(const double* v /*R*/, double* out /*T*/);   out = A^T v
//
// act[T]: the vector to project on entry, (vector - projection) on exit.  u[T]; r, p, ap [R]; w[R] nullable
// (zero on entry; receives the minimum-norm weights).  Returns the number of CG steps taken, or -1 when the true
// residual of the normal equations is still above tolerance after the restarts.
template <class Op>
__device__ int cta_cg_project(Op& op, int R, int T, double* act, double* u, double* r, double* p, double* ap,
                              double* w, double* red, int max_iter) {
  const int tid = threadIdx.x;
  constexpr double kTol2 = 1e-29;   // on |r|^2 relative to |A t|^2
  int steps = 0;
  double b2 = 0.0;
  for (int pass = 0; pass < 4; ++pass) {
    op.apply(act, r);               // true residual of the normal equations: A (t - A^T w)
    __syncthreads();
    double rs = cta_dot(r, r, R, red);
    if (pass == 0) b2 = rs;
    if (!(b2 > 0.0) || rs <= kTol2 * b2) return steps;
    for (int i = tid; i < R; i += kThreads) p[i] = r[i];
    __syncthreads();
    for (int it = 0; it < max_iter; ++it) {
      op.apply_t(p, u);
      __syncthreads();
      const double uu = cta_dot(u, u, T, red);    // p^T G p = |A^T p|^2
      if (!(uu > 0.0)) break;
      const double alpha = rs / uu;
      for (int n = tid; n < T; n += kThreads) act[n] = fma(-alpha, u[n], act[n]);
      if (w != nullptr)
        for (int i = tid; i < R; i += kThreads) w[i] = fma(alpha, p[i], w[i]);
      op.apply(u, ap);                            // G p
      __syncthreads();
      for (int i = tid; i < R; i += kThreads) r[i] = fma(-alpha, ap[i], r[i]);
      __syncthreads();
      const double rs_new = cta_dot(r, r, R, red);
      ++steps;
      if (rs_new <= kTol2 * b2) break;
      const double beta = rs_new / rs;
      for (int i = tid; i < R; i += kThreads) p[i] = fma(beta, p[i], r[i]);
      rs = rs_new;
      __syncthreads();
    }
    __syncthreads();
  }
  op.apply(act, r);
  __syncthreads();
  const double rs = cta_dot(r, r, R, red);
  return rs <= 1e-24 * b2 ? steps : -1;
}

}  // namespace pp
