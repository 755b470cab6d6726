// pyperiod_b200 -- QOPeriods on the B200: the quadratic-program residualisation of
// pyPeriod/QOPeriods.py:313-643, 743-852 as one persistent kernel.
//
// Per window (one CTA): up to `num` rounds of
//   gamma-norm sweep of the residual (QOPeriods.py:470-478; the sweep of pp_sweep.cuh)
//   -> dictionary layout: rows kept per period = sum of phi over newly seen divisors (:830-840)
//   -> normal equations  G w = W  with  G = A A^T (integer counts) and W = A x = fold sums of the
//      ORIGINAL data (:781-782), Cholesky in an L2-resident workspace, two triangular solves
//   -> reconstruction A^T w, residual = data - reconstruction (:517-522), stop test on
//      rms(reconstruction) (:391).
// The same solve stage serves RamanujanPeriods.find_periods_with_weights (RamanujanPeriods.py:106-112)
// through pp_qo_solve.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pyperiod_b200.h"
#include "pp_common.cuh"
#include "pp_sweep.cuh"
#include "pp_host.cuh"

namespace pp {

#ifndef PP_QO_CTAS
#define PP_QO_CTAS 2
#endif
constexpr int kQoCtasPerSm = PP_QO_CTAS;  // persistent CTAs per SM of the QO kernels
constexpr int kCholNb = 32;    // Cholesky block size
constexpr int kCholTile = 32;  // trailing-update tile: 32 x 32 outputs per warp, 4 x 8 per lane

struct QoPlan {
  int xs_len, n_even, rmax, num, seen_words, hier_len;
  __host__ __device__ size_t off_x0() const { return (size_t)xs_len * 8; }
  __host__ __device__ size_t off_wv() const { return off_x0() + (size_t)n_even * 8; }
  __host__ __device__ size_t off_chol() const { return off_wv() + (size_t)rmax * 8; }
  // Cholesky tiles and the hierarchical-sweep scratch are never live together
  __host__ __device__ size_t chol_bytes() const {
    const size_t a = (size_t)(2 * kCholNb * (kCholNb + 1) + kCholNb) * 8;
    const size_t b = (size_t)kWarps * hier_len * 8;
    return a > b ? a : b;
  }
  __host__ __device__ size_t off_red() const { return off_chol() + chol_bytes(); }
  __host__ __device__ size_t off_bar() const { return off_red() + 2 * kWarps * 8; }
  __host__ __device__ size_t off_sweep() const { return off_bar() + 16; }
  __host__ __device__ size_t off_ints() const { return off_sweep() + ((sizeof(SweepShared) + 15) & ~15); }
  // ints: found[num] dict_q[num] dict_keep[num] dict_rows[num] dict_off[num+1] prev_rows[num] seen[seen_words] misc[16]
  __host__ __device__ size_t bytes() const { return off_ints() + (size_t)(6 * num + 1 + seen_words + 16) * 4 + 16; }
  // leading dimension of the per-CTA Gram matrix: fixed for the whole window so that the factor of one round can
  // be extended in the next; not a power of two (column walks would hit one L2 set)
  __host__ __device__ int ldg() const { return rmax + 8; }
};

__host__ __device__ inline QoPlan make_qo_plan(int N, int pmax, int num, int rmax, bool hier) {
  QoPlan pl;
  pl.xs_len = (N + kSweepPad + 1) & ~1;
  pl.n_even = (N + 1) & ~1;
  pl.rmax = (rmax + 1) & ~1;
  pl.num = num;
  pl.seen_words = (pmax + 32) / 32;
  pl.hier_len = hier ? hier_scratch_len(pmax) : 0;
  return pl;
}

// ------------------------------------------------------------------------------------------
// blocked Cholesky of an R x R symmetric positive definite matrix (lower triangle, row-major, ld)
// in global memory, by one CTA.  Pt is a 32 x ld scratch that holds the current panel transposed.
// Returns false (uniformly) when a pivot is not positive: the reference raises LinAlgError there.
// ------------------------------------------------------------------------------------------
struct CholSmem {
  double* D;    // [32][33] diagonal block (factor L after step 1)
  double* rD;   // [32] column broadcast buffer
  double* Li;   // [32][33] inverse of the factored block
};

// row0 (a multiple of 32): rows below it already hold the factor of the leading row0 x row0 block (previous round);
// only the rows from row0 on are factored, with exactly the operations the full factorisation would apply to them.
__device__ bool cta_cholesky(double* __restrict__ A, int R, int ld, double* __restrict__ Pt, const CholSmem& cs,
                             int* flag, int row0 = 0, long long* tphase = nullptr) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) *flag = 0;
  long long tq = clock64();
  for (int kb = 0; kb < R; kb += kCholNb) {
    const int nb = min(kCholNb, R - kb);
    // (1) diagonal block -> shared memory, padded with the identity to 32 x 32
    for (int idx = tid; idx < kCholNb * kCholNb; idx += kThreads) {
      const int r = idx >> 5, c = idx & 31;
      double v = (r == c) ? 1.0 : 0.0;
      if (r < nb && c <= r) v = A[(size_t)(kb + r) * ld + kb + c];
      cs.D[r * (kCholNb + 1) + c] = v;
    }
    __syncthreads();
    const bool old_block = kb + nb <= row0;  // already factored: only its inverse is needed
    if (wid == 0) {
      // lane owns row `lane` of the block in registers; column k is broadcast through cs.rD each step
      double r[kCholNb];
#pragma unroll
      for (int c = 0; c < kCholNb; ++c) r[c] = cs.D[lane * (kCholNb + 1) + c];
      bool ok = true;
      if (!old_block) {  // warp-uniform
#pragma unroll
      for (int k = 0; k < kCholNb; ++k) {
        const double dkk = __shfl_sync(0xffffffffu, r[k], k);
        if (k < nb && !(dkk > 1e-8)) ok = false;   // uniform: every lane sees the same pivot
        const double d = sqrt(ok ? dkk : 1.0);
        const double l = (lane > k) ? r[k] / d : (lane == k ? d : 0.0);
        r[k] = l;
        cs.rD[lane] = l;
        __syncwarp();
#pragma unroll
        for (int j = k + 1; j < kCholNb; ++j) r[j] = fma(-l, cs.rD[j], r[j]);   // only j <= lane is ever used
        __syncwarp();
      }
      }
      if (!ok && lane == 0) *flag = 1;
#pragma unroll
      for (int c = 0; c < kCholNb; ++c) cs.D[lane * (kCholNb + 1) + c] = (c <= lane) ? r[c] : 0.0;
      __syncwarp();
      // inverse of the block, X = L^-1 (lower triangular): lane c solves column c by forward substitution
      double xcol[kCholNb];
#pragma unroll
      for (int rr = 0; rr < kCholNb; ++rr) {
        double v = (rr == lane) ? 1.0 : 0.0;
#pragma unroll
        for (int m = 0; m < rr; ++m) v = fma(-cs.D[rr * (kCholNb + 1) + m], xcol[m], v);
        xcol[rr] = (rr >= lane) ? v / cs.D[rr * (kCholNb + 1) + rr] : 0.0;
      }
      __syncwarp();
      // store transposed: cs.Li[c][m] = X[c][m] is what the panel needs (out[c] = sum_m row[m] X[c][m])
#pragma unroll
      for (int rr = 0; rr < kCholNb; ++rr) cs.Li[rr * (kCholNb + 1) + lane] = xcol[rr];
    }
    __syncthreads();
    if (*flag) return false;
    // write the factored diagonal block back
    if (!old_block) {
      for (int idx = tid; idx < nb * nb; idx += kThreads) {
        const int r = idx / nb, c = idx - r * nb;
        if (c <= r) A[(size_t)(kb + r) * ld + kb + c] = cs.D[r * (kCholNb + 1) + c];
      }
    }
    const int below = R - kb - nb;
    if (below <= 0) {
      __syncthreads();  // cs.D is restaged by the caller's next step: every write-back read must be done
      break;
    }
    // (2) panel: one thread per row, forward substitution against the diagonal block
    for (int i = kb + nb + tid; i < R; i += kThreads) {
      double row[kCholNb];
      double* a = A + (size_t)i * ld + kb;
#pragma unroll
      for (int c = 0; c < kCholNb; ++c) row[c] = (c < nb) ? a[c] : 0.0;
      if (i < row0) {  // finished row of the previous factor: only its transposed copy is needed
#pragma unroll
        for (int c = 0; c < kCholNb; ++c) Pt[(size_t)c * ld + i] = row[c];
        continue;
      }
      // out[c] = sum_{m <= c} row[m] Li[c][m] depends on the loaded row only: store as we go (no second array)
#pragma unroll
      for (int c = 0; c < kCholNb; ++c) {
        double v0 = 0.0, v1 = 0.0;
#pragma unroll
        for (int m = 0; m <= c; m += 2) {
          v0 = fma(row[m], cs.Li[c * (kCholNb + 1) + m], v0);
          if (m + 1 <= c) v1 = fma(row[m + 1], cs.Li[c * (kCholNb + 1) + m + 1], v1);
        }
        const double v = v0 + v1;
        if (c < nb) a[c] = v;
        Pt[(size_t)c * ld + i] = v;  // transposed copy: coalesced tile loads below
      }
    }
    __syncthreads();
    if (tphase) { const long long now = clock64(); tphase[0] += now - tq; tq = now; }
    // (3) trailing update  A[i][j] -= sum_k P[i][k] P[j][k]  (j <= i).  Each warp owns 32 x 32 output
    // tiles (round-robin over the lower-triangular tile pairs) and streams the transposed panel Pt
    // straight from L1/L2 into registers: no shared-memory staging, no CTA barrier inside the update.
    {
      const int base = kb + nb;
      const int nt = (R - base + kCholTile - 1) / kCholTile;
      const int npairs = nt * (nt + 1) / 2;
      const int ty = lane >> 2, tx = lane & 3;  // lane's 4 rows x 4 columns of each half tile
      for (int pr = wid; pr < npairs; pr += kWarps) {
        // pr -> (it, jt) with jt <= it
        int it = (int)((sqrtf(8.0f * (float)pr + 1.0f) - 1.0f) * 0.5f);
        while (it * (it + 1) / 2 > pr) --it;
        while ((it + 1) * (it + 2) / 2 <= pr) ++it;
        const int jt = pr - it * (it + 1) / 2;
        const int ti = base + it * kCholTile, tj = base + jt * kCholTile;
        if (ti + kCholTile <= row0) continue;  // rows of the previous factor (tiles are 32-aligned, as row0 is)
        // two 32 x 16 halves, each lane 4 rows x 4 columns: 16 accumulators (the 4 x 8 version spilled)
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          const int i0 = ti + ty * 4, j0 = tj + half * 16 + tx * 4;
          if (j0 > i0 + 3) continue;  // entirely above the diagonal
          double acc[4][4];
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
          const bool full = (ti + kCholTile <= R) && (tj + kCholTile <= R);
#pragma unroll 4
          for (int k = 0; k < kCholNb; ++k) {
            const double* pk = Pt + (size_t)k * ld;
            double av[4], bv[4];
            if (full) {
              const double2 a01 = *reinterpret_cast<const double2*>(pk + i0);
              const double2 a23 = *reinterpret_cast<const double2*>(pk + i0 + 2);
              const double2 b01 = *reinterpret_cast<const double2*>(pk + j0);
              const double2 b23 = *reinterpret_cast<const double2*>(pk + j0 + 2);
              av[0] = a01.x; av[1] = a01.y; av[2] = a23.x; av[3] = a23.y;
              bv[0] = b01.x; bv[1] = b01.y; bv[2] = b23.x; bv[3] = b23.y;
            } else {
#pragma unroll
              for (int r = 0; r < 4; ++r) av[r] = (i0 + r < R) ? pk[i0 + r] : 0.0;
#pragma unroll
              for (int c = 0; c < 4; ++c) bv[c] = (j0 + c < R) ? pk[j0 + c] : 0.0;
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
              for (int c = 0; c < 4; ++c) acc[r][c] = fma(av[r], bv[c], acc[r][c]);
          }
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int i = i0 + r;
            if (i < R) {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const int j = j0 + c;
                if (j <= i) A[(size_t)i * ld + j] -= acc[r][c];
              }
            }
          }
        }
      }
    }
    __syncthreads();
    if (tphase) { const long long now = clock64(); tphase[1] += now - tq; tq = now; }
  }
  return true;
}

// Solve L L^T w = b in place (b in shared memory, length R) with the factor from cta_cholesky.
// Each 32 x 32 diagonal block is staged in shared memory so the serial substitution of warp 0 never
// waits on global memory; the rectangular updates are spread over the CTA.
__device__ void cta_chol_solve(const double* __restrict__ A, int R, int ld, double* b, const CholSmem& cs) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  auto stage = [&](int kb, int nb) {
    for (int idx = tid; idx < kCholNb * kCholNb; idx += kThreads) {
      const int r = idx >> 5, c = idx & 31;
      double v = (r == c) ? 1.0 : 0.0;
      if (r < nb && c <= r) v = A[(size_t)(kb + r) * ld + kb + c];
      cs.D[r * (kCholNb + 1) + c] = v;
    }
    __syncthreads();
  };
  // forward: L y = b
  for (int kb = 0; kb < R; kb += kCholNb) {
    const int nb = min(kCholNb, R - kb);
    stage(kb, nb);
    if (wid == 0) {
      for (int k = 0; k < nb; ++k) {
        const double yk = b[kb + k] / cs.D[k * (kCholNb + 1) + k];
        __syncwarp();
        if (lane == k) b[kb + k] = yk;
        if (lane > k && lane < nb) b[kb + lane] -= cs.D[lane * (kCholNb + 1) + k] * yk;
        __syncwarp();
      }
    }
    __syncthreads();
    for (int i = kb + nb + tid; i < R; i += kThreads) {
      const double* a = A + (size_t)i * ld + kb;
      double s = 0.0;
      for (int c = 0; c < nb; ++c) s = fma(a[c], b[kb + c], s);
      b[i] -= s;
    }
    __syncthreads();
  }
  // backward: L^T w = y
  for (int kb = ((R - 1) / kCholNb) * kCholNb; kb >= 0; kb -= kCholNb) {
    const int nb = min(kCholNb, R - kb);
    stage(kb, nb);
    if (wid == 0) {
      for (int k = nb - 1; k >= 0; --k) {
        const double wk = b[kb + k] / cs.D[k * (kCholNb + 1) + k];
        __syncwarp();
        if (lane == k) b[kb + k] = wk;
        if (lane < k) b[kb + lane] -= cs.D[k * (kCholNb + 1) + lane] * wk;
        __syncwarp();
      }
    }
    __syncthreads();
    for (int j = tid; j < kb; j += kThreads) {
      double s = 0.0;
      for (int c = 0; c < nb; ++c) s = fma(A[(size_t)(kb + c) * ld + j], b[kb + c], s);
      b[j] -= s;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// dictionary layout + normal equations + solve + reconstruction for one window
// ------------------------------------------------------------------------------------------
struct QoCtx {
  int N, num, rmax;
  const int32_t* phi;   // device table, Euler phi for 0..table_pmax
  const double* x0;     // original data (shared)
  double* xs;           // residual out (shared, zero padded)
  double* wv;           // W -> weights (shared, rmax)
  double* red;
  int* found;           // periods in the order found (duplicates allowed)
  int* dict_q;          // dictionary: first-occurrence order
  int* dict_keep;       // value stored in the reference's basis_dictionary (0 for a repeated period)
  int* dict_rows;       // rows actually present in A (keep, or q when keep == 0: `if keep:` QOPeriods.py:972)
  int* dict_off;        // row offsets, [ndict+1]
  int* prev_rows;       // dict_rows of the round whose Cholesky factor is still in G (incremental factorisation)
  int ldg;              // leading dimension of G and Pt
  uint32_t* seen;       // bitmap of divisors already counted
  int seen_words;
  int* misc;            // [0]=ndict [1]=R [2]=flag [3]=layout scratch [8]=entries and [9]=rows of the factor held in G
  double* G;            // global, rmax*rmax
  double* Pt;           // global, 32*rmax
  CholSmem cs;
  long long* t;         // per-thread phase timers (development aid): [0] layout [1] build [2] cholesky [3] solve
};

// QOPeriods.get_subspaces (QOPeriods.py:830-840) for `found[0..nfound)`.  All threads call (contains
// barriers): the divisor scan of each period is spread over the CTA, the bookkeeping is thread 0's.
__device__ void qo_layout(const QoCtx& c, int nfound) {
  const int tid = threadIdx.x;
  for (int i = tid; i < c.seen_words; i += kThreads) c.seen[i] = 0u;
  if (tid == 0) c.misc[0] = 0;
  __syncthreads();
  for (int f = 0; f < nfound; ++f) {
    const int q = c.found[f];
    if (tid == 0) c.misc[3] = 0;
    __syncthreads();
    int fresh = 0;
    for (int d = 1 + tid; d <= q; d += kThreads) {
      if (q % d == 0 && !((c.seen[d >> 5] >> (d & 31)) & 1u)) {
        fresh += c.phi[d];
        atomicOr(&c.seen[d >> 5], 1u << (d & 31));
      }
    }
    if (fresh) atomicAdd(&c.misc[3], fresh);
    __syncthreads();
    if (tid == 0) {
      int ndict = c.misc[0], slot = -1;
      for (int k = 0; k < ndict; ++k)
        if (c.dict_q[k] == q) slot = k;
      if (slot < 0) slot = ndict++;
      c.dict_q[slot] = q;
      c.dict_keep[slot] = c.misc[3];   // a repeated period overwrites its entry with 0 new dimensions
      c.misc[0] = ndict;
    }
    __syncthreads();
  }
  if (tid == 0) {
    const int ndict = c.misc[0];
    int off = 0;
    for (int k = 0; k < ndict; ++k) {
      c.dict_rows[k] = c.dict_keep[k] ? c.dict_keep[k] : c.dict_q[k];
      c.dict_off[k] = off;
      off += c.dict_rows[k];
    }
    c.dict_off[ndict] = off;
    c.misc[1] = off;
  }
}

// Returns 0 ok, PP_STATUS_SINGULAR, PP_STATUS_TOO_LARGE.  On success wv holds the weights, xs the residual,
// and *e_recon the sum of squares of the reconstruction.  All threads call; contains barriers.
__device__ int cta_qo_solve(const QoCtx& c, int nfound, double* e_recon, bool explicit_layout = false) {
  const int tid = threadIdx.x;
  const int N = c.N;
  long long tm = clock64();
  // explicit_layout: the caller has filled dict_q / dict_rows / dict_off and misc[0] (entries), misc[1] (rows)
  if (!explicit_layout) qo_layout(c, nfound);
  __syncthreads();
  { const long long now = clock64(); c.t[0] += now - tm; tm = now; }
  const int ndict = c.misc[0], R = c.misc[1];
  if (R > c.rmax) return PP_STATUS_TOO_LARGE;
  if (R == 0) {
    for (int n = tid; n < N; n += kThreads) c.xs[n] = c.x0[n];
    __syncthreads();
    *e_recon = 0.0;
    return PP_STATUS_OK;
  }
  const int ld = c.ldg;
  // The dictionary only grows from round to round (unless a repeated period rewrites an entry), so the factor of
  // the previous round is the leading block of this round's: rebuild and factor only the rows from the last
  // 32-aligned boundary on (the partial block below it is rebuilt from its integer counts).
  if (tid == 0) {
    const int nd_prev = c.misc[8], r_prev = c.misc[9];
    bool grow = r_prev > 0 && ndict >= nd_prev;
    for (int k = 0; grow && k < nd_prev; ++k) grow = c.dict_rows[k] == c.prev_rows[k];
    c.misc[10] = grow ? (r_prev / kCholNb) * kCholNb : 0;
    c.misc[9] = 0;  // no valid factor until this round's succeeds
  }
  __syncthreads();
  const int row_lo = c.misc[10];
  // W = A x: fold sums of the original data, one thread per kept row, terms in increasing n
  for (int k = 0; k < ndict; ++k) {
    const int q = c.dict_q[k], rows = c.dict_rows[k], off = c.dict_off[k];
    for (int i = tid; i < rows; i += kThreads) {
      double s = 0.0;
      for (int n = i; n < N; n += q) s += c.x0[n];
      c.wv[off + i] = s;
    }
  }
  // G = A A^T, lower triangle: zero, diagonal counts, then co-occurrence counts of residue pairs
  for (int r = row_lo; r < R; ++r)
    for (int j = tid; j <= r; j += kThreads) c.G[(size_t)r * ld + j] = 0.0;
  __syncthreads();
  for (int a = 0; a < ndict; ++a) {
    const int qa = c.dict_q[a], ra = c.dict_rows[a], oa = c.dict_off[a];
    if (oa + ra <= row_lo) continue;  // rows of the factor that is kept
    for (int i = tid; i < ra; i += kThreads)
      if (oa + i >= row_lo) c.G[(size_t)(oa + i) * ld + oa + i] = (double)((N - 1 - i) / qa + 1);
    for (int b = 0; b < a; ++b) {
      const int qb = c.dict_q[b], rb = c.dict_rows[b], ob = c.dict_off[b];
      int i = tid % qa, j = tid % qb;
      const int si = kThreads % qa, sj = kThreads % qb;
      for (int n = tid; n < N; n += kThreads) {
        if (i < ra && j < rb && oa + i >= row_lo)
          atomicAdd(&c.G[(size_t)(oa + i) * ld + ob + j], 1.0);  // integer-valued: order-free
        i += si; if (i >= qa) i -= qa;
        j += sj; if (j >= qb) j -= qb;
      }
    }
  }
  __syncthreads();
  { const long long now = clock64(); c.t[1] += now - tm; tm = now; }
  if (!cta_cholesky(c.G, R, ld, c.Pt, c.cs, &c.misc[2], row_lo, c.t + 4)) return PP_STATUS_SINGULAR;
  if (tid == 0) {  // G now holds the factor of this dictionary
    c.misc[8] = ndict;
    c.misc[9] = R;
    for (int k = 0; k < ndict; ++k) c.prev_rows[k] = c.dict_rows[k];
  }
  { const long long now = clock64(); c.t[2] += now - tm; tm = now; }
  cta_chol_solve(c.G, R, ld, c.wv, c.cs);
  // reconstruction A^T w and residual
  double e = 0.0;
  for (int n = tid; n < N; n += kThreads) {
    double r = 0.0;
    for (int k = 0; k < ndict; ++k) {
      const int i = n % c.dict_q[k];
      if (i < c.dict_rows[k]) r += c.wv[c.dict_off[k] + i];
    }
    c.xs[n] = c.x0[n] - r;
    e = fma(r, r, e);
  }
  e = warp_sum(e);
  __syncthreads();
  if ((tid & 31) == 0) c.red[tid >> 5] = e;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < kWarps; ++w) t += c.red[w];
  __syncthreads();
  c.t[3] += clock64() - tm;
  *e_recon = t;
  return PP_STATUS_OK;
}

struct QoOut {
  uint32_t* periods;   // [B, num]  periods reported (found order)
  double* norms;       // [B, num]
  int32_t* n_periods;  // [B] number reported (may be one less than the dictionary holds, QOPeriods.py:585-588)
  int32_t* dict_q;     // [B, num]
  int32_t* dict_keep;  // [B, num]
  int32_t* n_dict;     // [B]
  int32_t* n_weights;  // [B]
  double* weights;     // [B, rmax]
  double* res;         // [B, N] (nullable)
  int32_t* status;     // [B]
};

__device__ void qo_commit(const QoCtx& c, const QoOut& o, int b, int nfound, const double* round_norms) {
  const int tid = threadIdx.x;
  const int ndict = c.misc[0], R = c.misc[1];
  for (int i = tid; i < c.num; i += kThreads) {
    o.periods[(size_t)b * c.num + i] = i < nfound ? (uint32_t)c.found[i] : 0u;
    o.norms[(size_t)b * c.num + i] = i < nfound ? round_norms[i] : 0.0;
    o.dict_q[(size_t)b * c.num + i] = i < ndict ? c.dict_q[i] : 0;
    o.dict_keep[(size_t)b * c.num + i] = i < ndict ? c.dict_keep[i] : 0;
  }
  for (int i = tid; i < R; i += kThreads) o.weights[(size_t)b * c.rmax + i] = c.wv[i];
  if (o.res)
    for (int n = tid; n < c.N; n += kThreads) o.res[(size_t)b * c.N + n] = c.xs[n];
  if (tid == 0) {
    o.n_periods[b] = nfound;
    o.n_dict[b] = ndict;
    o.n_weights[b] = R;
  }
}

__device__ __forceinline__ QoCtx make_ctx(unsigned char* smem, const QoPlan& pl, int N, int num, const int32_t* phi,
                                          double* G, double* Pt) {
  QoCtx c;
  c.N = N;
  c.num = num;
  c.rmax = pl.rmax;
  c.phi = phi;
  c.xs = reinterpret_cast<double*>(smem);
  c.x0 = reinterpret_cast<double*>(smem + pl.off_x0());
  c.wv = reinterpret_cast<double*>(smem + pl.off_wv());
  double* chol = reinterpret_cast<double*>(smem + pl.off_chol());
  c.cs.D = chol;
  c.cs.rD = chol + kCholNb * (kCholNb + 1);
  c.cs.Li = c.cs.rD + kCholNb;
  c.red = reinterpret_cast<double*>(smem + pl.off_red());
  int* ints = reinterpret_cast<int*>(smem + pl.off_ints());
  c.found = ints;
  c.dict_q = ints + num;
  c.dict_keep = ints + 2 * num;
  c.dict_rows = ints + 3 * num;
  c.dict_off = ints + 4 * num;
  c.prev_rows = ints + 5 * num + 1;
  c.ldg = pl.ldg();
  c.seen = reinterpret_cast<uint32_t*>(ints + 6 * num + 1);
  c.seen_words = pl.seen_words;
  c.misc = ints + 6 * num + 1 + pl.seen_words;
  c.G = G;
  c.Pt = Pt;
  c.t = nullptr;
  return c;
}

// ------------------------------------------------------------------------------------------
// QOPeriods.find_periods, default branch (QOPeriods.py:313-596)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 2)
qo_find_kernel(const double* __restrict__ x, int64_t ldx, int B, int N, int num, double thresh, int pmin, int pmax,
               int trunc, int hier, const int32_t* __restrict__ phi, int rmax, QoOut out, double* __restrict__ ws_G,
               double* __restrict__ ws_Pt, double* __restrict__ ws_norms, const uint2* __restrict__ tops, int ntops,
               unsigned long long* __restrict__ prof, int* __restrict__ next_window) {
  unsigned char* smem = pp_smem;
  const QoPlan pl = make_qo_plan(N, pmax, num, rmax, hier != 0);
  const size_t ldg = (size_t)pl.ldg();
  QoCtx c = make_ctx(smem, pl, N, num, phi, ws_G + (size_t)blockIdx.x * pl.rmax * ldg,
                     ws_Pt + (size_t)blockIdx.x * kCholNb * ldg);
  SweepShared* sweep = reinterpret_cast<SweepShared*>(smem + pl.off_sweep());
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + pl.off_bar());
  double* round_norms = ws_norms + (size_t)blockIdx.x * num;
  double* x0 = const_cast<double*>(c.x0);
  const double sqrtN = sqrt((double)N);
  const int tid = threadIdx.x;
  long long timers[6] = {0, 0, 0, 0, 0, 0}, t_sweep = 0;
  c.t = timers;

  WindowLoader loader;
  loader.init(bar);
  for (int i = N + tid; i < pl.xs_len; i += kThreads) c.xs[i] = 0.0;
  sweep_shared_init(sweep);

  for (WindowQueue wq(next_window); wq.b < B; wq.next()) {
    const int b = wq.b;
    loader.load(x0, x + (size_t)b * ldx, N);
    // zero-signal early out (QOPeriods.py:394-406): sum |x| <= 1e-16
    double sa = 0.0;
    for (int n = tid; n < N; n += kThreads) {
      sa += fabs(x0[n]);
      c.xs[n] = x0[n];
    }
    sa = warp_sum(sa);
    if ((tid & 31) == 0) c.red[tid >> 5] = sa;
    __syncthreads();
    double sum_abs = 0.0;
    for (int w = 0; w < kWarps; ++w) sum_abs += c.red[w];
    __syncthreads();
    const double e_data = cta_sum_sq(x0, N, c.red);
    // defaults: what the reference returns when nothing could be solved (empty lists, res = data)
    if (tid == 0) {
      c.misc[0] = 0;
      c.misc[1] = 0;
      c.misc[8] = 0;  // no Cholesky factor of this window in G yet
      c.misc[9] = 0;
    }
    __syncthreads();
    qo_commit(c, out, b, 0, round_norms);
    int status = PP_STATUS_OK;
    if (sum_abs <= 1e-16) {
      if (out.res)
        for (int n = tid; n < N; n += kThreads) out.res[(size_t)b * N + n] = 0.0;
      if (tid == 0) out.status[b] = PP_STATUS_ZERO_INPUT;
      __syncthreads();
      continue;
    }
    if (tid == 0) {
      SweepParams& sp = sweep->params;
      sp.N = N;
      sp.pmin = pmin;
      sp.pmax = pmax;
      sp.metric = PP_METRIC_GAMMA;
      sp.trunc = trunc;
      sp.orth = 0;  // QOPeriods.py:471-473 passes orthogonalize = False
      sp.chain_off = nullptr;
      sp.chain_q = nullptr;
      sp.warp_scr = nullptr;
      sp.pv = 0;
      sp.sqrtN = sqrtN;
      sp.e_res = 0.0;
      sp.data_norm = 1.0;
      sp.thresh = -1.0;
      sp.skip = nullptr;
      sp.nskip = 0;
      sp.metric_out = nullptr;
      sp.hier_scr = pl.hier_len ? reinterpret_cast<double*>(smem + pl.off_chol()) : nullptr;
      sp.hier_len = pl.hier_len;
      sp.rcp = sweep->rcp;
      sp.tops = tops;
      sp.ntops = ntops;
      sp.verify_keys = nullptr;
      sp.xf0_off = 0;
      sp.xf1_off = 0;
    }
    __syncthreads();
    int nfound = 0;
    int reported = 0;
    double e_recon = 0.0;
    const double rms_data = sqrt(e_data / (double)N);  // rms(), QOPeriods.py:78-79
    for (int i = 0; i < num; ++i) {
      if (i > 0) {
        // default test_function: rms(reconstruction) > rms(data) * thresh  (QOPeriods.py:391)
        const bool go = sqrt(e_recon / (double)N) > rms_data * thresh;
        if (!go) {
          // weights are re-solved with all periods (identical to what we hold) but the last period is
          // not reported (QOPeriods.py:560-594)
          reported = nfound - 1;
          break;
        }
      }
      const long long t0 = clock64();
      const SweepResult top = cta_sweep<kSweepHier | kSweepNoMetricOut>(sweep);
      t_sweep += clock64() - t0;
      if (tid == 0) {
        round_norms[i] = top.val;
        if (top.p > 0) c.found[nfound] = top.p;
      }
      if (top.p > 0) ++nfound;
      __syncthreads();
      const int rc = cta_qo_solve(c, nfound, &e_recon);
      if (rc != PP_STATUS_OK) {  // LinAlgError in the reference: keep the previous round's outputs (:552-559)
        status = rc;
        break;
      }
      // norms are indexed by round in the reference (norms[:len(found)]); rounds without a period only
      // happen once the residual is exactly zero, after which nothing changes
      qo_commit(c, out, b, nfound, round_norms);
      reported = nfound;
      __syncthreads();
    }
    if (tid == 0) {
      out.n_periods[b] = status == PP_STATUS_OK ? reported : out.n_periods[b];
      out.status[b] = status;
    }
    __syncthreads();
  }
  if (prof != nullptr && tid == 0) {
    atomicAdd(prof + 0, (unsigned long long)t_sweep);
    atomicAdd(prof + 1, (unsigned long long)timers[0]);
    atomicAdd(prof + 2, (unsigned long long)timers[1]);
    atomicAdd(prof + 3, (unsigned long long)timers[2]);
    atomicAdd(prof + 5, (unsigned long long)timers[3]);
    atomicAdd(prof + 6, (unsigned long long)timers[4]);
    atomicAdd(prof + 7, (unsigned long long)timers[5]);
    atomicAdd(prof + 4, (unsigned long long)((B - blockIdx.x + gridDim.x - 1) / gridDim.x));
  }
}

// ------------------------------------------------------------------------------------------
// solve stage alone, for given periods (RamanujanPeriods.find_periods_with_weights :106-112)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 2)
qo_solve_kernel(const double* __restrict__ x, int64_t ldx, int B, int N, int kmax, const int32_t* __restrict__ periods,
                const int32_t* __restrict__ nper, int pmax, const int32_t* __restrict__ phi, int rmax, QoOut out,
                double* __restrict__ ws_G, double* __restrict__ ws_Pt, int* __restrict__ next_window) {
  unsigned char* smem = pp_smem;
  const QoPlan pl = make_qo_plan(N, pmax, kmax, rmax, false);
  const size_t ldg = (size_t)pl.ldg();
  QoCtx c = make_ctx(smem, pl, N, kmax, phi, ws_G + (size_t)blockIdx.x * pl.rmax * ldg,
                     ws_Pt + (size_t)blockIdx.x * kCholNb * ldg);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + pl.off_bar());
  double* x0 = const_cast<double*>(c.x0);
  const int tid = threadIdx.x;
  long long timers[6] = {0, 0, 0, 0, 0, 0};
  c.t = timers;
  WindowLoader loader;
  loader.init(bar);
  for (WindowQueue wq(next_window); wq.b < B; wq.next()) {
    const int b = wq.b;
    loader.load(x0, x + (size_t)b * ldx, N);
    const int nfound = min(nper[b], kmax);
    for (int i = tid; i < nfound; i += kThreads) c.found[i] = periods[(size_t)b * kmax + i];
    if (tid == 0) c.misc[8] = c.misc[9] = 0;  // one solve per window: nothing to extend
    __syncthreads();
    double e_recon = 0.0;
    const int rc = cta_qo_solve(c, nfound, &e_recon);
    if (rc == PP_STATUS_OK) {
      // norms are the caller's (periodogram values); only layout, weights and residual are produced here
      const int ndict = c.misc[0], R = c.misc[1];
      for (int i = tid; i < kmax; i += kThreads) {
        out.dict_q[(size_t)b * kmax + i] = i < ndict ? c.dict_q[i] : 0;
        out.dict_keep[(size_t)b * kmax + i] = i < ndict ? c.dict_keep[i] : 0;
      }
      for (int i = tid; i < R; i += kThreads) out.weights[(size_t)b * c.rmax + i] = c.wv[i];
      if (out.res)
        for (int n = tid; n < N; n += kThreads) out.res[(size_t)b * N + n] = c.xs[n];
      if (tid == 0) {
        out.n_dict[b] = ndict;
        out.n_weights[b] = R;
      }
    } else if (tid == 0) {
      out.n_dict[b] = 0;
      out.n_weights[b] = 0;
    }
    if (tid == 0) out.status[b] = rc;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// solve stage for a caller-supplied dictionary layout (QOPeriodsWithGCDsExtracted.get_subspaces,
// QOPeriodsWithGCDsExtracted.py:98-143: the layout depends on CPython set order and is built on the host)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 2)
qo_solve_rows_kernel(const double* __restrict__ x, int64_t ldx, int B, int N, int kmax, const int32_t* __restrict__ dict_q,
                     const int32_t* __restrict__ dict_rows, const int32_t* __restrict__ n_dict, int pmax, int rmax,
                     int32_t* __restrict__ n_weights, double* __restrict__ weights, double* __restrict__ res,
                     int32_t* __restrict__ status, double* __restrict__ ws_G, double* __restrict__ ws_Pt,
                     int* __restrict__ next_window) {
  unsigned char* smem = pp_smem;
  const QoPlan pl = make_qo_plan(N, pmax, kmax, rmax, false);
  const size_t ldg = (size_t)pl.ldg();
  QoCtx c = make_ctx(smem, pl, N, kmax, nullptr, ws_G + (size_t)blockIdx.x * pl.rmax * ldg,
                     ws_Pt + (size_t)blockIdx.x * kCholNb * ldg);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + pl.off_bar());
  double* x0 = const_cast<double*>(c.x0);
  const int tid = threadIdx.x;
  long long timers[6] = {0, 0, 0, 0, 0, 0};
  c.t = timers;
  WindowLoader loader;
  loader.init(bar);
  for (WindowQueue wq(next_window); wq.b < B; wq.next()) {
    const int b = wq.b;
    loader.load(x0, x + (size_t)b * ldx, N);
    const int nd = min(max(n_dict[b], 0), kmax);
    if (tid == 0) {
      int off = 0;
      for (int k = 0; k < nd; ++k) {
        const int q = dict_q[(size_t)b * kmax + k];
        int rows = dict_rows[(size_t)b * kmax + k];
        rows = rows > 0 && rows <= q ? rows : q;   // 0 keeps every row (QOPeriods.py:972)
        c.dict_q[k] = q;
        c.dict_keep[k] = rows;
        c.dict_rows[k] = rows;
        c.dict_off[k] = off;
        off += rows;
      }
      c.dict_off[nd] = off;
      c.misc[0] = nd;
      c.misc[1] = off;
      c.misc[8] = c.misc[9] = 0;  // one solve per window: nothing to extend
    }
    __syncthreads();
    double e_recon = 0.0;
    const int rc = cta_qo_solve(c, nd, &e_recon, true);
    if (rc == PP_STATUS_OK) {
      const int R = c.misc[1];
      for (int i = tid; i < R; i += kThreads) weights[(size_t)b * c.rmax + i] = c.wv[i];
      if (res)
        for (int n = tid; n < N; n += kThreads) res[(size_t)b * N + n] = c.xs[n];
      if (tid == 0) n_weights[b] = R;
    } else if (tid == 0) {
      n_weights[b] = 0;
    }
    if (tid == 0) status[b] = rc;
  }
}

}  // namespace pp

using namespace pp;

extern "C" {

size_t pp_qo_workspace_bytes(int32_t N, int32_t pmax, int32_t num, int32_t rmax) {
  DeviceFacts f;
  if (device_facts(f)) return 0;
  const QoPlan pl = make_qo_plan(N, pmax, num, rmax, true);
  const size_t grid = (size_t)f.sm_count * 2;
  return 8192 + (size_t)(pmax + 2) * sizeof(uint2) +
         grid * ((size_t)pl.rmax * pl.ldg() + (size_t)kCholNb * pl.ldg() + (size_t)num + 64) * 8;
}

static int qo_check(const void* x, int64_t ldx, int B, int N, int num, int pmax, int rmax, const void* phi,
                    int table_pmax) {
  if (x == nullptr || B < 0 || N < 2 || ldx < 1) return fail(-1, "bad window arguments%s");
  if (num < 1 || num > 64) return fail(-1, "need 1 <= num <= 64%s");
  if (pmax < 2 || pmax > N) return fail(-1, "need 2 <= pmax <= N%s");
  if (rmax < 2) return fail(-1, "rmax must be >= 2%s");
  if (phi == nullptr || table_pmax < pmax) return fail(-1, "phi table must cover pmax%s");
  return 0;
}

int pp_qo_find_periods(const double* x, int64_t ldx, int32_t B, int32_t N, int32_t num, double thresh, int32_t pmin,
                       int32_t pmax, int32_t trunc, const int32_t* phi, int32_t table_pmax, int32_t rmax,
                       uint32_t* periods, double* norms, int32_t* n_periods, int32_t* dict_q, int32_t* dict_keep,
                       int32_t* n_dict, int32_t* n_weights, double* weights, double* res, int32_t* status,
                       void* workspace, size_t workspace_bytes, void* stream) {
  if (B == 0) return 0;  // empty batch: nothing to validate or launch
  if (int rc = qo_check(x, ldx, B, N, num, pmax, rmax, phi, table_pmax)) return rc;
  if (pmin < 1 || pmin > pmax) return fail(-1, "need 1 <= pmin <= pmax%s");
  if (!periods || !norms || !n_periods || !dict_q || !dict_keep || !n_dict || !n_weights || !weights || !status)
    return fail(-1, "output pointers are null%s");
  if (B == 0) return 0;
  DeviceFacts f;
  if (int rc = device_facts(f)) return rc;
  const int hier = (pp_get_fold_mode() != PP_FOLD_DIRECT && !trunc) ? 1 : 0;
  const QoPlan pl = make_qo_plan(N, pmax, num, rmax, hier != 0);
  if (int rc = prep_kernel(qo_find_kernel, pl.bytes(), f)) return rc;
  const int grid = grid_for(f, pl.bytes(), B, kQoCtasPerSm);
  size_t off = 0;
  double* G = carve(workspace, workspace_bytes, off, (size_t)grid * pl.rmax * pl.ldg() * 8);
  double* Pt = carve(workspace, workspace_bytes, off, (size_t)grid * kCholNb * pl.ldg() * 8);
  double* nr = carve(workspace, workspace_bytes, off, (size_t)grid * num * 8);
  if (!G || !Pt || !nr) return fail(-3, "workspace too small (see pp_qo_workspace_bytes)%s");
  uint2* tops = nullptr;
  int ntops = hier ? hier_top_count(pmin, pmax) : 0;
  if (ntops > 0) {
    tops = reinterpret_cast<uint2*>(carve(workspace, workspace_bytes, off, (size_t)ntops * sizeof(uint2)));
    if (!tops) return fail(-3, "workspace too small (see pp_qo_workspace_bytes)%s");
    ntops = build_hier_jobs(N, pmin, pmax, tops, (cudaStream_t)stream);
  }
  QoOut o{periods, norms, n_periods, dict_q, dict_keep, n_dict, n_weights, weights, res, status};
  int* next_window = carve_window_counter(workspace, workspace_bytes, off, (cudaStream_t)stream);
  qo_find_kernel<<<grid, kThreads, pl.bytes(), (cudaStream_t)stream>>>(x, ldx, B, N, num, thresh, pmin, pmax, trunc,
                                                                       hier, phi, pl.rmax, o, G, Pt, nr, tops, ntops,
                                                                       reinterpret_cast<unsigned long long*>(pp_get_profile_buffer()),
                                                                       next_window);
  return check_cuda(cudaGetLastError(), "qo_find_kernel launch");
}

int pp_qo_solve(const double* x, int64_t ldx, int32_t B, int32_t N, int32_t kmax, const int32_t* periods,
                const int32_t* nper, int32_t pmax, const int32_t* phi, int32_t table_pmax, int32_t rmax,
                int32_t* dict_q, int32_t* dict_keep, int32_t* n_dict, int32_t* n_weights, double* weights,
                double* res, int32_t* status, void* workspace, size_t workspace_bytes, void* stream) {
  if (B == 0) return 0;  // empty batch: nothing to validate or launch
  if (int rc = qo_check(x, ldx, B, N, kmax, pmax, rmax, phi, table_pmax)) return rc;
  if (!periods || !nper || !dict_q || !dict_keep || !n_dict || !n_weights || !weights || !status)
    return fail(-1, "pointers are null%s");
  if (B == 0) return 0;
  DeviceFacts f;
  if (int rc = device_facts(f)) return rc;
  const QoPlan pl = make_qo_plan(N, pmax, kmax, rmax, false);
  if (int rc = prep_kernel(qo_solve_kernel, pl.bytes(), f)) return rc;
  const int grid = grid_for(f, pl.bytes(), B, kQoCtasPerSm);
  size_t off = 0;
  double* G = carve(workspace, workspace_bytes, off, (size_t)grid * pl.rmax * pl.ldg() * 8);
  double* Pt = carve(workspace, workspace_bytes, off, (size_t)grid * kCholNb * pl.ldg() * 8);
  if (!G || !Pt) return fail(-3, "workspace too small (see pp_qo_workspace_bytes)%s");
  QoOut o{nullptr, nullptr, nullptr, dict_q, dict_keep, n_dict, n_weights, weights, res, status};
  int* next_window = carve_window_counter(workspace, workspace_bytes, off, (cudaStream_t)stream);
  qo_solve_kernel<<<grid, kThreads, pl.bytes(), (cudaStream_t)stream>>>(x, ldx, B, N, kmax, periods, nper, pmax, phi,
                                                                        pl.rmax, o, G, Pt, next_window);
  return check_cuda(cudaGetLastError(), "qo_solve_kernel launch");
}

int pp_qo_solve_rows(const double* x, int64_t ldx, int32_t B, int32_t N, int32_t kmax, const int32_t* dict_q,
                     const int32_t* dict_rows, const int32_t* n_dict, int32_t pmax, int32_t rmax, int32_t* n_weights,
                     double* weights, double* res, int32_t* status, void* workspace, size_t workspace_bytes,
                     void* stream) {
  if (B == 0) return 0;  // empty batch: nothing to validate or launch
  if (x == nullptr || B < 0 || N < 2 || ldx < 1) return fail(-1, "bad window arguments%s");
  if (kmax < 1 || kmax > 256) return fail(-1, "need 1 <= kmax <= 256%s");
  if (pmax < 1 || pmax > N) return fail(-1, "need 1 <= pmax <= N%s");
  if (rmax < 2) return fail(-1, "rmax must be >= 2%s");
  if (!dict_q || !dict_rows || !n_dict || !n_weights || !weights || !status) return fail(-1, "pointers are null%s");
  DeviceFacts f;
  if (int rc = device_facts(f)) return rc;
  const QoPlan pl = make_qo_plan(N, pmax, kmax, rmax, false);
  if (int rc = prep_kernel(qo_solve_rows_kernel, pl.bytes(), f)) return rc;
  const int grid = grid_for(f, pl.bytes(), B, kQoCtasPerSm);
  size_t off = 0;
  double* G = carve(workspace, workspace_bytes, off, (size_t)grid * pl.rmax * pl.ldg() * 8);
  double* Pt = carve(workspace, workspace_bytes, off, (size_t)grid * kCholNb * pl.ldg() * 8);
  if (!G || !Pt) return fail(-3, "workspace too small (see pp_qo_workspace_bytes)%s");
  int* next_window = carve_window_counter(workspace, workspace_bytes, off, (cudaStream_t)stream);
  qo_solve_rows_kernel<<<grid, kThreads, pl.bytes(), (cudaStream_t)stream>>>(x, ldx, B, N, kmax, dict_q, dict_rows, n_dict,
                                                                             pmax, pl.rmax, n_weights, weights, res,
                                                                             status, G, Pt, next_window);
  return check_cuda(cudaGetLastError(), "qo_solve_rows_kernel launch");
}

}  // extern "C"
