// pyperiod_b200 -- the period sweep: one warp per candidate period, residue sums in registers.
//
// Direct fold: for period p a warp keeps S_p[r] for r = rb + lane + 32 j (j < J <= 16) in registers
// and walks the rows k of the (rows x p) rectangle.  Every shared-memory read is 32 consecutive
// doubles (conflict-free, 2 wavefronts) and feeds exactly one DADD, so the sweep sits on the
// shared-memory roofline (8 B per add).  Sums are sequential in n -- the order numpy
// uses -- so MAXABS metrics are bit-exact; energies agree with the reference's BLAS norm to a few ulp.
//
// Hierarchical fold (ranking sweeps only): only the "top" periods q in (pmax/2, pmax] are folded
// from the window; S_p for every smaller candidate follows from S_2p[r] + S_2p[r + p].  Half the
// shared-memory traffic; sums differ from the sequential ones by rounding only.
//
// Candidates are ranked by the squared metric (energy, or energy / p compared by cross
// multiplication); the square root and divisions are taken once, for the winner.
#pragma once

#include "pp_common.cuh"
#include "../../include/pyperiod_b200.h"

namespace pp {

enum PassMode { kPassEnergy = 0, kPassEnergyTail = 1, kPassMaxAbs = 2, kPassStore = 3 };

// Lives in shared memory (one per CTA), written by thread 0 before a sweep.
struct SweepParams {
  int N;
  int pmin, pmax;     // inclusive candidate range
  int metric;         // PP_METRIC_*
  int trunc;
  int orth;
  const int32_t* chain_off;  // device tables (orth only)
  const int32_t* chain_q;
  double* warp_scr;   // per-warp global scratch, 2*pv doubles (orth only)
  int pv;
  double sqrtN;
  double e_res;       // ||residual||^2   (IMPOSED)
  double data_norm;   // ||x||/sqrt(N)    (IMPOSED)
  double thresh;      // IMPOSED early-stop threshold; <0 disables first-hit mode
  const uint32_t* skip;  // bitmap of periods to ignore (M-best), nullable
  double* metric_out;    // global [pmax+1], nullable
  double* hier_scr;      // shared scratch of the hierarchical sweep: kWarps * hier_len doubles (nullable => direct)
  int hier_len;
  const double* rcp;     // shared table rcp[m] = 1.0 / m for m < kRcpTab (energy weights without a division)
};

constexpr int kRcpTab = 256;

// 1 / m: table for the small row counts every large period has, a real division otherwise
__device__ __forceinline__ double rcp_of(const double* rcp, int m) {
  return m < kRcpTab ? rcp[m] : 1.0 / (double)m;
}

struct SweepResult {
  double val;  // metric value of the winner (norm / gamma norm / max |S| / imposed norm)
  int p;
};

// Ranking key: energy for NORM and GAMMA (GAMMA compares energy / p by cross multiplication),
// the metric value itself for MAXABS and IMPOSED.
struct Best {
  double key;
  int p;
};

__device__ __forceinline__ bool better(int metric, double key, int p, const Best& b) {
  if (b.p == 0) return key > 0.0;
  if (metric == PP_METRIC_GAMMA) {
    const double l = key * (double)b.p, r = b.key * (double)p;
    return l > r || (l == r && p < b.p);
  }
  return key > b.key || (key == b.key && p < b.p);
}

__device__ __forceinline__ double key_to_value(int metric, double key, int p, double sqrtN) {
  if (metric == PP_METRIC_MAXABS || metric == PP_METRIC_IMPOSED) return key;
  double v = sqrt(key) / sqrtN;                            // periodic_norm, Periods.py:241
  if (metric == PP_METRIC_GAMMA) v = v / sqrt((double)p);  // Periods.py:239
  return v;
}

// loop-invariant pieces of SweepParams a warp keeps in registers while it ranks candidates
struct RankCtx {
  const uint32_t* skip;
  double* metric_out;
  double sqrtN;
  int metric;
};
__device__ __forceinline__ RankCtx rank_ctx(const SweepParams* sp) {
  return RankCtx{sp->skip, sp->metric_out, sp->sqrtN, sp->metric};
}

__device__ __forceinline__ void consider(const RankCtx& rc, double key, int p, Best& best) {
  const int metric = rc.metric;
  if (rc.metric_out != nullptr && (threadIdx.x & 31) == 0) rc.metric_out[p] = key_to_value(metric, key, p, rc.sqrtN);
  const uint32_t* skip = rc.skip;
  if (skip != nullptr && ((skip[p >> 5] >> (p & 31)) & 1u)) return;
  if (better(metric, key, p, best)) {
    best.key = key;
    best.p = p;
  }
}

// a += v*v on the lanes where x < y: one ISETP and one predicated DFMA
__device__ __forceinline__ void fma_sq_if_lt(double& a, double v, int x, int y) {
  asm("{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %2, %3;\n\t@p fma.rn.f64 %0, %1, %1, %0;\n\t}"
      : "+d"(a)
      : "d"(v), "r"(x), "r"(y));
}

// acc[j] += row[32 j] for j < J, with the loads of a batch issued together before the adds
// (written out so the compiler keeps several shared-memory loads in flight per warp)
template <int J>
__device__ __forceinline__ void add_row(double (&acc)[J], const double* __restrict__ row) {
  constexpr int BATCH = 8;
#pragma unroll
  for (int j0 = 0; j0 < J; j0 += BATCH) {
    double t[BATCH];
#pragma unroll
    for (int u = 0; u < BATCH; ++u)
      if (j0 + u < J) t[u] = row[32 * (j0 + u)];
#pragma unroll
    for (int u = 0; u < BATCH; ++u)
      if (j0 + u < J) acc[j0 + u] += t[u];
  }
}

// ------------------------------------------------------------------------------------------
// direct fold of one residue block
// ------------------------------------------------------------------------------------------
// T = sum of S^2 over valid residues, A = the same over residues < r0 (those with one more term),
// C = sum of S_r * x[M p + r] (trunc + IMPOSED only); for MAXABS T carries max |S_r|.
template <int J, int MODE>
__device__ __forceinline__ void block_pass(const double* __restrict__ xs, int p, int rb, int rows, int r0, int M,
                                           bool trunc, double& T, double& A, double& C, double* __restrict__ gsS,
                                           double* __restrict__ gsV) {
  const int lane = threadIdx.x & 31;
  const double* ptr = xs + rb + lane;
  double acc[J];
#pragma unroll
  for (int j = 0; j < J; ++j) acc[j] = ptr[32 * j];
  if (J <= 4) {
#pragma unroll 4
    for (int k = 1; k < rows; ++k) {
      ptr += p;
      add_row<J>(acc, ptr);
    }
  } else {
#pragma unroll 1
    for (int k = 1; k < rows; ++k) {
      ptr += p;
      add_row<J>(acc, ptr);
    }
  }
  const int tail_off = M * p;
  if (MODE == kPassEnergy || MODE == kPassEnergyTail) {
    // lanes past p hold sums of the next row's samples; only the last two registers can (J gap <= 2)
#pragma unroll
    for (int j = (J > 2 ? J - 2 : 0); j < J; ++j)
      if (rb + lane + 32 * j >= p) acc[j] = 0.0;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const double s = acc[j];
      T = fma(s, s, T);
      fma_sq_if_lt(A, s, lane, r0 - (rb + 32 * j));
      if (MODE == kPassEnergyTail) C = fma(s, xs[tail_off + rb + lane + 32 * j], C);  // zero pad beyond N
    }
  } else if (MODE == kPassMaxAbs) {
#pragma unroll
    for (int j = 0; j < J; ++j)
      if (rb + lane + 32 * j < p) T = fmax(T, fabs(acc[j]));
  } else {  // kPassStore: full-N sums and (approximate) means to the warp scratch
    const double invHi = 1.0 / (double)(M + 1), invLo = 1.0 / (double)M;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int r = rb + lane + 32 * j;
      if (r < p) {
        const double s = acc[j];
        if (trunc) {
          gsS[r] = s + xs[tail_off + r];
          gsV[r] = s * invLo;
        } else {
          gsS[r] = s;
          gsV[r] = s * (r < r0 ? invHi : invLo);
        }
      }
    }
  }
}

#define PP_J_DISPATCH(MODE, JN, ...)                              \
  do {                                                            \
    switch (JN) {                                                 \
      case 1: block_pass<1, MODE>(__VA_ARGS__); break;            \
      case 2: block_pass<2, MODE>(__VA_ARGS__); break;            \
      case 3: block_pass<3, MODE>(__VA_ARGS__); break;            \
      case 4: block_pass<4, MODE>(__VA_ARGS__); break;            \
      case 5: block_pass<5, MODE>(__VA_ARGS__); break;            \
      case 6: block_pass<6, MODE>(__VA_ARGS__); break;            \
      case 7: block_pass<7, MODE>(__VA_ARGS__); break;            \
      case 8: block_pass<8, MODE>(__VA_ARGS__); break;            \
      case 9: case 10: block_pass<10, MODE>(__VA_ARGS__); break;  \
      case 11: case 12: block_pass<12, MODE>(__VA_ARGS__); break; \
      case 13: case 14: block_pass<14, MODE>(__VA_ARGS__); break; \
      default: block_pass<16, MODE>(__VA_ARGS__); break;          \
    }                                                             \
  } while (0)

// Ranking key of one period, computed by one warp; identical in all lanes.
template <int MODE>
__device__ __noinline__ double warp_period_key(const SweepParams* sp, int p) {
  const int lane = threadIdx.x & 31;
  const int N = sp->N;
  const bool trunc = sp->trunc != 0;
  const double* xs = staged_window();
  const int M = N / p, r0 = N - M * p;
  const int rows = trunc ? M : (M + (r0 ? 1 : 0));
  double T = 0.0, A = 0.0, C = 0.0;
  double* gsS = nullptr;
  double* gsV = nullptr;
  if (MODE == kPassStore) {
    gsS = sp->warp_scr + (size_t)(threadIdx.x >> 5) * 2 * sp->pv;
    gsV = gsS + sp->pv;
  }
  for (int rb = 0; rb < p; rb += kResBlock) {
    const int jn = (min(p - rb, kResBlock) + 31) >> 5;
    PP_J_DISPATCH(MODE, jn, xs, p, rb, rows, r0, M, trunc, T, A, C, gsS, gsV);
  }
  if (MODE == kPassMaxAbs) return warp_max(T);

  double energy, dot;
  if (MODE == kPassStore) {
    __syncwarp();
    warp_orth_chain_approx(gsV, p, N, trunc, sp->chain_q + sp->chain_off[p], sp->chain_off[p + 1] - sp->chain_off[p]);
    double e = 0.0, d = 0.0;
    for (int r = lane; r < p; r += 32) {
      const double v = gsV[r];
      e = fma((double)(M + (r < r0 ? 1 : 0)) * v, v, e);
      d = fma(v, gsS[r], d);
    }
    energy = warp_sum(e);
    dot = warp_sum(d);
    __syncwarp();
  } else {
    const double m = (double)M;
    if (trunc) {
      // mean = S / M on every residue, counts over all N:  sum cnt * mean^2 = (M T + A) / M^2
      energy = warp_sum(fma(m, T, A) / (m * m));
      dot = (MODE == kPassEnergyTail) ? warp_sum((T + C) / m) : energy;
    } else {
      const double w_lo = 1.0 / m, w_diff = 1.0 / (double)(M + 1) - w_lo;
      energy = warp_sum(fma(w_diff, A, w_lo * T));
      dot = energy;
    }
  }
  if (sp->metric == PP_METRIC_IMPOSED) {
    const double e_res = sp->e_res, sqrtN = sp->sqrtN;
    const double e_trial = fmax(e_res - 2.0 * dot + energy, 0.0);
    return (sqrt(e_res) / sqrtN - sqrt(e_trial) / sqrtN) / sp->data_norm;  // Periods.py:278-280
  }
  return energy;
}

__device__ __forceinline__ double warp_period_key_any(const SweepParams* sp, int p) {
  if (sp->metric == PP_METRIC_MAXABS) return warp_period_key<kPassMaxAbs>(sp, p);
  if (sp->orth) return warp_period_key<kPassStore>(sp, p);
  if (sp->metric == PP_METRIC_IMPOSED && sp->trunc) return warp_period_key<kPassEnergyTail>(sp, p);
  return warp_period_key<kPassEnergy>(sp, p);
}

// ------------------------------------------------------------------------------------------
// hierarchical fold
// ------------------------------------------------------------------------------------------
// A top q = g * 2^L (L <= 3) is folded at base period g with 2^L accumulator sets selected by
// (row mod 2^L): set s holds S_q[r + s g].  Pairwise in-register adds then give the sums of
// q/2, q/4, .. g.  If g is still even the chain continues through a small per-warp scratch.
__host__ __device__ inline int hier_scratch_len(int pmax) { return ((((pmax >> 3) + 2) * 3) / 2 + 3) & ~1; }

// compile-time recursion over the levels l = LV .. 0 (keeps every accumulator index static).
// Level energy = w_lo * T + w_diff * A (T over all residues, A over residues < r0).
template <int L, int J, int LV>
struct hier_levels {
  static __device__ __forceinline__ void run(double (&acc)[1 << L][J], int lane, int rb, int g,
                                             const int (&r0)[L + 1], double (&T)[L + 1], double (&A)[L + 1]) {
    constexpr int sets = 1 << LV;
#pragma unroll
    for (int s = 0; s < sets; ++s) {
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const double v = acc[s][j];
        T[LV] = fma(v, v, T[LV]);
        fma_sq_if_lt(A[LV], v, lane, r0[LV] - (rb + 32 * j + s * g));
      }
    }
    if constexpr (LV > 0) {
      constexpr int half = sets >> 1;
#pragma unroll
      for (int s = 0; s < half; ++s)
#pragma unroll
        for (int j = 0; j < J; ++j) acc[s][j] += acc[s + half][j];
      hier_levels<L, J, LV - 1>::run(acc, lane, rb, g, r0, T, A);
    }
  }
};

// One residue block [rb, rb + 32 J) of a top q = g * 2^L: 2^L accumulator sets of J registers.
// The caller picks the smallest available J >= needed with a gap of at most 2 registers, so only
// the last two registers of a set can hold lanes past g.
template <int L, int J>
__device__ __forceinline__ void hier_pass(const double* __restrict__ xs, int rb, int g, int rows,
                                          const int (&r0)[L + 1], double (&T)[L + 1], double (&A)[L + 1],
                                          double* scr) {
  constexpr int S = 1 << L;
  const int lane = threadIdx.x & 31;
  const double* ptr = xs + rb + lane;
  double acc[S][J];
  if (S == 1) {
#pragma unroll
    for (int j = 0; j < J; ++j) acc[0][j] = ptr[32 * j];
    if (J <= 4) {
#pragma unroll 4
      for (int k = 1; k < rows; ++k) {
        ptr += g;
        add_row<J>(acc[0], ptr);
      }
    } else {
#pragma unroll 1
      for (int k = 1; k < rows; ++k) {
        ptr += g;
        add_row<J>(acc[0], ptr);
      }
    }
  } else {
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
      for (int j = 0; j < J; ++j) acc[s][j] = 0.0;
    int k = 0;
#pragma unroll 1
    for (; k + S <= rows; k += S) {
#pragma unroll
      for (int s = 0; s < S; ++s) {
        add_row<J>(acc[s], ptr);
        ptr += g;
      }
    }
#pragma unroll
    for (int s = 0; s < S - 1; ++s) {
      if (k + s < rows) {
        add_row<J>(acc[s], ptr);
        ptr += g;
      }
    }
  }
#pragma unroll
  for (int j = (J > 2 ? J - 2 : 0); j < J; ++j) {
    if (rb + lane + 32 * j >= g) {
#pragma unroll
      for (int s = 0; s < S; ++s) acc[s][j] = 0.0;
    }
  }
  hier_levels<L, J, L>::run(acc, lane, rb, g, r0, T, A);
  if (scr != nullptr) {
#pragma unroll
    for (int j = 0; j < J; ++j)
      if (rb + lane + 32 * j < g) scr[rb + lane + 32 * j] = acc[0][j];
  }
}

#define PP_HIER_CASE(L, J) \
  case J: hier_pass<L, J>(xs, rb, g, rows, r0, T, A, scr); break;

template <int L>
__device__ __forceinline__ void hier_top_fold(const double* xs, int g, int rows, const int (&r0)[L + 1],
                                              double (&T)[L + 1], double (&A)[L + 1], double* scr) {
  constexpr int JMAX = (kResBlock / 32) >> L;  // 16, 8, 4, 2 registers per accumulator set
  for (int rb = 0; rb < g; rb += 32 * JMAX) {
    int jn = (min(g - rb, 32 * JMAX) + 31) >> 5;
    if (jn > 8) jn = (jn + 1) & ~1;  // even sizes only above 8
    if constexpr (JMAX == 2) {
      switch (jn) {
        PP_HIER_CASE(L, 1)
        default: hier_pass<L, 2>(xs, rb, g, rows, r0, T, A, scr); break;
      }
    } else if constexpr (JMAX == 4) {
      switch (jn) {
        PP_HIER_CASE(L, 1) PP_HIER_CASE(L, 2) PP_HIER_CASE(L, 3)
        default: hier_pass<L, 4>(xs, rb, g, rows, r0, T, A, scr); break;
      }
    } else if constexpr (JMAX == 8) {
      switch (jn) {
        PP_HIER_CASE(L, 1) PP_HIER_CASE(L, 2) PP_HIER_CASE(L, 3) PP_HIER_CASE(L, 4)
        PP_HIER_CASE(L, 5) PP_HIER_CASE(L, 6) PP_HIER_CASE(L, 7)
        default: hier_pass<L, 8>(xs, rb, g, rows, r0, T, A, scr); break;
      }
    } else {
      switch (jn) {
        PP_HIER_CASE(L, 1) PP_HIER_CASE(L, 2) PP_HIER_CASE(L, 3) PP_HIER_CASE(L, 4)
        PP_HIER_CASE(L, 5) PP_HIER_CASE(L, 6) PP_HIER_CASE(L, 7) PP_HIER_CASE(L, 8)
        PP_HIER_CASE(L, 10) PP_HIER_CASE(L, 12) PP_HIER_CASE(L, 14)
        default: hier_pass<L, 16>(xs, rb, g, rows, r0, T, A, scr); break;
      }
    }
  }
}

// One top period q = g * 2^L and every candidate q / 2^k below it.  Non-trunc, non-orth, NORM / GAMMA.
template <int L>
__device__ __noinline__ Best warp_hier_top_L(const SweepParams* sp, int g, double* scr, Best best) {
  const int lane = threadIdx.x & 31;
  const int N = sp->N;
  const int pmin = sp->pmin;
  const RankCtx rc = rank_ctx(sp);
  const double* rcp = sp->rcp;
  // rows of the level periods: floor(N / (g 2^i)) = floor(N / g) >> i
  const int M0 = N / g;
  int r0[L + 1];
#pragma unroll
  for (int i = 0; i <= L; ++i) r0[i] = N - (M0 >> i) * (g << i);
  const int rows = M0 + (r0[0] ? 1 : 0);
  const bool chain = (L == 3) && !(g & 1) && (g >> 1) >= pmin;
  double T[L + 1], A[L + 1];
#pragma unroll
  for (int i = 0; i <= L; ++i) T[i] = A[i] = 0.0;
  hier_top_fold<L>(staged_window(), g, rows, r0, T, A, chain ? scr : nullptr);
#pragma unroll
  for (int i = 0; i <= L; ++i) {
    const int M = M0 >> i;
    const double w_lo = rcp_of(rcp, M), w_diff = rcp_of(rcp, M + 1) - w_lo;
    consider(rc, warp_sum(fma(w_diff, A[i], w_lo * T[i])), g << i, best);
  }
  if (chain) {
    __syncwarp();
    double* src = scr;
    double* dst = scr + ((g + 1) & ~1);
    int h = g;
    while (!(h & 1) && (h >> 1) >= pmin) {
      const int h2 = h >> 1;
      const int M = N / h2, r0h = N - M * h2;
      const double w_lo = rcp_of(rcp, M), w_diff = rcp_of(rcp, M + 1) - w_lo;
      double t = 0.0, a = 0.0;
      for (int r = lane; r < h2; r += 32) {
        const double v = src[r] + src[r + h2];
        dst[r] = v;
        t = fma(v, v, t);
        if (r < r0h) a = fma(v, v, a);
      }
      __syncwarp();
      consider(rc, warp_sum(fma(w_diff, a, w_lo * t)), h2, best);
      double* swp = src;
      src = dst;
      dst = swp;
      h = h2;
    }
    __syncwarp();
  }
  return best;
}

__device__ __forceinline__ void warp_hier_top(const SweepParams* sp, int q, double* scr, Best& best) {
  int L = min(__ffs(q) - 1, 3);
  while (L > 0 && (q >> L) < sp->pmin) --L;
  const int g = q >> L;
  switch (L) {
    case 0: best = warp_hier_top_L<0>(sp, g, scr, best); break;
    case 1: best = warp_hier_top_L<1>(sp, g, scr, best); break;
    case 2: best = warp_hier_top_L<2>(sp, g, scr, best); break;
    default: best = warp_hier_top_L<3>(sp, g, scr, best); break;
  }
}

// ------------------------------------------------------------------------------------------
// CTA-level sweep
// ------------------------------------------------------------------------------------------
struct SweepShared {
  double rcp[kRcpTab];  // rcp[m] = 1 / m (rcp[0] unused); filled once per CTA by sweep_shared_init
  SweepParams params;
  int counter;  // next candidate index
  int hit_p;    // first-hit mode: lowest period over threshold so far
  double wkey[kWarps];
  int wp[kWarps];
};

__device__ __forceinline__ void sweep_shared_init(SweepShared* sh) {
  for (int m = threadIdx.x; m < kRcpTab; m += kThreads) sh->rcp[m] = m ? 1.0 / (double)m : 0.0;
}

// Sweep all candidates with dynamic (atomic-counter) distribution over the CTA's warps.
//   argmax mode (thresh < 0): strict '>' from 0, lowest p on ties, periods in `skip` ignored
//                             (Periods.py:512-515).
//   first-hit mode (thresh >= 0): lowest p whose metric > thresh; warps stop once their next
//                             candidate lies above the current hit (Periods.py:273-286).
// sh->params must have been written and a CTA barrier passed.  All threads call; the result is valid
// in all threads.  Contains CTA barriers.  Kept out of line so the caller's live state does not
// compete with the fold's registers.
__device__ __noinline__ SweepResult cta_sweep(SweepShared* sh) {
  const SweepParams* sp = &sh->params;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    sh->counter = 0;
    sh->hit_p = 0x7fffffff;
  }
  __syncthreads();  // also publishes params written by thread 0 just before the call
  const int metric = sp->metric;
  const int pmin = sp->pmin, pmax = sp->pmax;
  const double thresh = sp->thresh;
  const bool first_hit = thresh >= 0.0;
  Best best{0.0, 0};
  const bool hier = sp->hier_scr != nullptr && !first_hit && !sp->orth && !sp->trunc &&
                    (metric == PP_METRIC_NORM || metric == PP_METRIC_GAMMA);
  if (hier) {
    // tops: candidates p with 2p > pmax; everything else is derived from exactly one of them.
    // They are handed out grouped by L = min(ctz(q), 3) so that all warps of the SM run the same
    // specialisation at the same time (the per-(L, J) code does not fit the instruction cache together).
    const int top_lo = max(pmin, (pmax >> 1) + 1);
    int first[4], cnt[4], total = 0;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const int mod = (l < 3) ? (2 << l) : 8, rem = (l < 3) ? (1 << l) : 0;
      const int f = top_lo + ((rem - top_lo) % mod + mod) % mod;  // smallest q >= top_lo, q = rem (mod mod)
      first[l] = f;
      cnt[l] = f <= pmax ? (pmax - f) / mod + 1 : 0;
      total += cnt[l];
    }
    double* scr = sp->hier_scr + (size_t)wid * sp->hier_len;
    while (true) {
      int idx = 0;
      if (lane == 0) idx = atomicAdd(&sh->counter, 1);
      idx = __shfl_sync(0xffffffffu, idx, 0);
      if (idx >= total) break;
      int q;
      if (idx < cnt[0]) q = first[0] + idx * 2;
      else if ((idx -= cnt[0]) < cnt[1]) q = first[1] + idx * 4;
      else if ((idx -= cnt[1]) < cnt[2]) q = first[2] + idx * 8;
      else q = first[3] + (idx - cnt[2]) * 8;
      warp_hier_top(sp, q, scr, best);
    }
  } else {
    const RankCtx rc = rank_ctx(sp);
    const int ncand = pmax - pmin + 1;
    while (true) {
      int idx = 0;
      if (lane == 0) idx = atomicAdd(&sh->counter, 1);
      idx = __shfl_sync(0xffffffffu, idx, 0);
      if (idx >= ncand) break;
      const int p = pmin + idx;
      if (first_hit) {
        const int hp = *reinterpret_cast<volatile int*>(&sh->hit_p);
        if (p > hp) break;
      }
      const double key = warp_period_key_any(sp, p);
      if (first_hit) {
        if (sp->metric_out != nullptr && lane == 0) sp->metric_out[p] = key;
        if (key > thresh) {
          if (best.p == 0 || p < best.p) {
            best.p = p;
            best.key = key;
          }
          if (lane == 0) atomicMin(&sh->hit_p, p);
        }
      } else {
        consider(rc, key, p, best);
      }
    }
  }
  if (lane == 0) {
    sh->wkey[wid] = best.key;
    sh->wp[wid] = best.p;
  }
  __syncthreads();
  Best res{0.0, 0};
#pragma unroll
  for (int w = 0; w < kWarps; ++w) {
    const double k = sh->wkey[w];
    const int q = sh->wp[w];
    if (q == 0) continue;
    if (first_hit) {
      if (res.p == 0 || q < res.p) res = Best{k, q};
    } else if (better(metric, k, q, res)) {
      res = Best{k, q};
    }
  }
  SweepResult out{0.0, res.p};
  if (res.p != 0) out.val = key_to_value(metric, res.key, res.p, sp->sqrtN);
  __syncthreads();  // wkey/wp may be rewritten by the next sweep
  return out;
}

}  // namespace pp
