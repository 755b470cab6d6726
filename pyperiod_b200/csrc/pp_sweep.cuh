// pyperiod_b200 -- the period sweep: one warp per candidate period, residue sums in registers.
//
// For period p a warp keeps S_p[r] for r = rb + lane + 32 j (j < J <= 32) in registers and walks
// the rows k of the (rows x p) rectangle: every shared-memory read is 32 consecutive doubles
// (conflict-free, 2 wavefronts) and feeds exactly one DADD, so the sweep sits on the
// shared-memory roofline (8 B per add).  Sums are sequential in n, i.e. the same
// order numpy uses, so MAXABS metrics are bit-exact; energies agree with the reference's BLAS
// norm to a few ulp.
#pragma once

#include "pp_common.cuh"
#include "../../include/pyperiod_b200.h"

namespace pp {

// what one residue-block pass does with the sums it holds
enum PassMode { kPassEnergy = 0, kPassEnergyTail = 1, kPassMaxAbs = 2, kPassStore = 3 };

struct SweepParams {
  const double* xs;   // staged window, zero padded to N + pmax + kSweepPad
  int N;
  int pmin, pmax;     // inclusive candidate range
  int metric;         // PP_METRIC_*
  bool trunc;
  bool orth;
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
  double* hier_scr;      // shared scratch for the hierarchical sweep: kWarps * hier_len doubles (nullable => direct)
  int hier_len;
};

struct SweepResult {
  double val;
  int p;
};

template <int J, int MODE>
__device__ __forceinline__ void block_pass(const double* __restrict__ xs, int p, int rb, int rows, int r0, int M,
                                           int tail_off, bool trunc, double& a0, double& a1, double& a2,
                                           double* __restrict__ gsS, double* __restrict__ gsV) {
  const int lane = threadIdx.x & 31;
  const double* ptr = xs + rb + lane;
  double acc[J];
#pragma unroll
  for (int j = 0; j < J; ++j) acc[j] = ptr[32 * j];
  if (J <= 4) {
#pragma unroll 4
    for (int k = 1; k < rows; ++k) {
      ptr += p;
#pragma unroll
      for (int j = 0; j < J; ++j) acc[j] += ptr[32 * j];
    }
  } else {
#pragma unroll 2
    for (int k = 1; k < rows; ++k) {
      ptr += p;
#pragma unroll
      for (int j = 0; j < J; ++j) acc[j] += ptr[32 * j];
    }
  }
  const double invHi = 1.0 / (double)(M + 1), invLo = 1.0 / (double)M;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int r = rb + lane + 32 * j;
    const double s = acc[j];
    if (MODE == kPassEnergy || MODE == kPassEnergyTail) {
      const double s2 = s * s;
      if (r < r0) a0 += s2;
      else if (r < p) a1 += s2;
      if (MODE == kPassEnergyTail) {
        if (r < p) a2 = fma(s, xs[tail_off + r], a2);  // samples past M*p (zero pad beyond N)
      }
    } else if (MODE == kPassMaxAbs) {
      if (r < p) a0 = fmax(a0, fabs(s));
    } else {  // kPassStore: full-N sums and (approximate) means to the warp scratch
      if (r < p) {
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

#define PP_J_DISPATCH(MODE, JN, ...)                                        \
  do {                                                                      \
    if ((JN) <= 1) block_pass<1, MODE>(__VA_ARGS__);                        \
    else if ((JN) <= 2) block_pass<2, MODE>(__VA_ARGS__);                   \
    else if ((JN) <= 3) block_pass<3, MODE>(__VA_ARGS__);                   \
    else if ((JN) <= 4) block_pass<4, MODE>(__VA_ARGS__);                   \
    else if ((JN) <= 5) block_pass<5, MODE>(__VA_ARGS__);                   \
    else if ((JN) <= 6) block_pass<6, MODE>(__VA_ARGS__);                   \
    else if ((JN) <= 7) block_pass<7, MODE>(__VA_ARGS__);                   \
    else if ((JN) <= 8) block_pass<8, MODE>(__VA_ARGS__);                   \
    else if ((JN) <= 10) block_pass<10, MODE>(__VA_ARGS__);                 \
    else if ((JN) <= 12) block_pass<12, MODE>(__VA_ARGS__);                 \
    else if ((JN) <= 14) block_pass<14, MODE>(__VA_ARGS__);                 \
    else if ((JN) <= 16) block_pass<16, MODE>(__VA_ARGS__);                 \
    else if ((JN) <= 20) block_pass<20, MODE>(__VA_ARGS__);                 \
    else if ((JN) <= 24) block_pass<24, MODE>(__VA_ARGS__);                 \
    else if ((JN) <= 28) block_pass<28, MODE>(__VA_ARGS__);                 \
    else block_pass<32, MODE>(__VA_ARGS__);                                 \
  } while (0)

// Metric of one period, computed by one warp; result identical in all lanes.
template <int MODE>
__device__ __forceinline__ double warp_period_metric(const SweepParams& sp, int p) {
  const int lane = threadIdx.x & 31;
  const int N = sp.N;
  const int M = N / p, r0 = N - M * p;
  const int rows = sp.trunc ? M : (M + (r0 ? 1 : 0));
  const int tail_off = M * p;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  double* gsS = nullptr;
  double* gsV = nullptr;
  if (MODE == kPassStore) {
    gsS = sp.warp_scr + (size_t)(threadIdx.x >> 5) * 2 * sp.pv;
    gsV = gsS + sp.pv;
  }
  for (int rb = 0; rb < p; rb += kResBlock) {
    const int nres = min(p - rb, kResBlock);
    const int jn = (nres + 31) >> 5;
    PP_J_DISPATCH(MODE, jn, sp.xs, p, rb, rows, r0, M, tail_off, sp.trunc, a0, a1, a2, gsS, gsV);
  }
  if (MODE == kPassMaxAbs) return warp_max(a0);

  double energy, dot;
  if (MODE == kPassStore) {
    __syncwarp();
    warp_orth_chain_approx(gsV, p, N, sp.trunc, sp.chain_q + sp.chain_off[p], sp.chain_off[p + 1] - sp.chain_off[p]);
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
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    if (sp.trunc) {
      const double m = (double)M;
      energy = ((double)(M + 1) * a0 + m * a1) / (m * m);
      dot = (MODE == kPassEnergyTail) ? (a0 + a1 + warp_sum(a2)) / m : energy;
    } else {
      energy = a0 / (double)(M + 1) + a1 / (double)M;
      dot = energy;
    }
  }
  if (sp.metric == PP_METRIC_IMPOSED) {
    const double e_trial = fmax(sp.e_res - 2.0 * dot + energy, 0.0);
    return (sqrt(sp.e_res) / sp.sqrtN - sqrt(e_trial) / sp.sqrtN) / sp.data_norm;
  }
  double val = sqrt(energy) / sp.sqrtN;
  if (sp.metric == PP_METRIC_GAMMA) val = val / sqrt((double)p);
  return val;
}

__device__ __forceinline__ double warp_period_metric_any(const SweepParams& sp, int p) {
  if (sp.metric == PP_METRIC_MAXABS) return warp_period_metric<kPassMaxAbs>(sp, p);
  if (sp.orth) return warp_period_metric<kPassStore>(sp, p);
  if (sp.metric == PP_METRIC_IMPOSED && sp.trunc) return warp_period_metric<kPassEnergyTail>(sp, p);
  return warp_period_metric<kPassEnergy>(sp, p);
}


// ------------------------------------------------------------------------------------------
// hierarchical sweep (ranking only): fold only the "top" periods q in (pmax/2, pmax] from the
// window; every smaller candidate is q / 2^k of exactly one top and its sums follow from
// S_{p}[r] = S_{2p}[r] + S_{2p}[r + p].  Halves the shared-memory traffic of a sweep.  The
// derived sums differ from the sequential ones by rounding only (<= a few ulp), which is the
// same order as the reference's own BLAS norm noise; exact projections of the winners are
// always recomputed sequentially (cta_project_exact).
//
// A top q = g * 2^L (L <= 3) is folded at base period g with 2^L accumulator sets selected by
// (row mod 2^L): set s holds S_q[r + s g].  Pairwise in-register adds then give the sums of
// q/2, q/4, .. g.  If g is still even the chain continues through a small per-warp scratch.
// ------------------------------------------------------------------------------------------
__host__ __device__ inline int hier_scratch_len(int pmax) { return ((((pmax >> 3) + 2) * 3) / 2 + 3) & ~1; }

struct HierLevels {
  int r0[4];
  double inv_hi[4], inv_lo[4];
};

// compile-time recursion over the levels l = LV .. 0 (keeps every accumulator index static)
template <int L, int W, int LV>
struct hier_levels {
  static __device__ __forceinline__ void run(double (&acc)[1 << L][W], int r_base, int g, const HierLevels& lv,
                                             double (&E)[4]) {
    constexpr int sets = 1 << LV;
#pragma unroll
    for (int s = 0; s < sets; ++s) {
#pragma unroll
      for (int w = 0; w < W; ++w) {
        const double v = acc[s][w];
        const int rho = r_base + 32 * w + s * g;
        E[LV] = fma(v * v, rho < lv.r0[LV] ? lv.inv_hi[LV] : lv.inv_lo[LV], E[LV]);
      }
    }
    if constexpr (LV > 0) {
      constexpr int half = sets >> 1;
#pragma unroll
      for (int s = 0; s < half; ++s)
#pragma unroll
        for (int w = 0; w < W; ++w) acc[s][w] += acc[s + half][w];
      hier_levels<L, W, LV - 1>::run(acc, r_base, g, lv, E);
    }
  }
};

template <int L, int W>
__device__ __forceinline__ void hier_chunk(const double* __restrict__ ptr, int cb, int g, int rows,
                                           const HierLevels& lv, double (&E)[4], double* scr) {
  constexpr int S = 1 << L;
  const int lane = threadIdx.x & 31;
  double acc[S][W];
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int w = 0; w < W; ++w) acc[s][w] = 0.0;
  int k = 0;
  if (S == 1) {
#pragma unroll 4
    for (; k < rows; ++k) {
#pragma unroll
      for (int w = 0; w < W; ++w) acc[0][w] += ptr[32 * w];
      ptr += g;
    }
  } else {
#pragma unroll 1
    for (; k + S <= rows; k += S) {
#pragma unroll
      for (int s = 0; s < S; ++s) {
#pragma unroll
        for (int w = 0; w < W; ++w) acc[s][w] += ptr[32 * w];
        ptr += g;
      }
    }
#pragma unroll
    for (int s = 0; s < S - 1; ++s) {
      if (k + s < rows) {
#pragma unroll
        for (int w = 0; w < W; ++w) acc[s][w] += ptr[32 * w];
        ptr += g;
      }
    }
  }
  // lanes past g hold sums of the next row's samples: drop them once
#pragma unroll
  for (int w = 0; w < W; ++w) {
    if (cb + lane + 32 * w >= g) {
#pragma unroll
      for (int s = 0; s < S; ++s) acc[s][w] = 0.0;
    }
  }
  hier_levels<L, W, L>::run(acc, cb + lane, g, lv, E);
  if (scr != nullptr) {
#pragma unroll
    for (int w = 0; w < W; ++w)
      if (cb + lane + 32 * w < g) scr[cb + lane + 32 * w] = acc[0][w];
  }
}

template <int L>
__device__ __forceinline__ void hier_top_fold(const double* xs, int g, int rows, const HierLevels& lv,
                                              double (&E)[4], double* scr) {
  constexpr int WMAX = (L == 3) ? 4 : 8;
  const int lane = threadIdx.x & 31;
  for (int cb = 0; cb < g; cb += 32 * WMAX) {
    const int w = (min(g - cb, 32 * WMAX) + 31) >> 5;
    const double* ptr = xs + cb + lane;
    switch (w) {
      case 1: hier_chunk<L, 1>(ptr, cb, g, rows, lv, E, scr); break;
      case 2: hier_chunk<L, 2>(ptr, cb, g, rows, lv, E, scr); break;
      case 3: hier_chunk<L, 3>(ptr, cb, g, rows, lv, E, scr); break;
      case 4: hier_chunk<L, 4>(ptr, cb, g, rows, lv, E, scr); break;
      default:
        if (WMAX == 8) {
          switch (w) {
            case 5: hier_chunk<L, (WMAX == 8 ? 5 : 1)>(ptr, cb, g, rows, lv, E, scr); break;
            case 6: hier_chunk<L, (WMAX == 8 ? 6 : 1)>(ptr, cb, g, rows, lv, E, scr); break;
            case 7: hier_chunk<L, (WMAX == 8 ? 7 : 1)>(ptr, cb, g, rows, lv, E, scr); break;
            default: hier_chunk<L, (WMAX == 8 ? 8 : 1)>(ptr, cb, g, rows, lv, E, scr); break;
          }
        }
        break;
    }
  }
}

// keeps the running strict-'>' argmax with lowest-p tie-break, honouring the skip bitmap
__device__ __forceinline__ void consider(const SweepParams& sp, double val, int p, double& bestv, int& bestp) {
  if (sp.metric_out != nullptr && (threadIdx.x & 31) == 0) sp.metric_out[p] = val;
  const bool skipped = sp.skip != nullptr && ((sp.skip[p >> 5] >> (p & 31)) & 1u);
  if (!skipped && (val > bestv || (val == bestv && bestp != 0 && p < bestp))) {
    bestv = val;
    bestp = p;
  }
}

// One top period q and every candidate q / 2^k below it.  Non-trunc, non-orth, NORM / GAMMA.
__device__ __forceinline__ void warp_hier_top(const SweepParams& sp, int q, double* scr, double& bestv, int& bestp) {
  const int lane = threadIdx.x & 31;
  const int N = sp.N;
  int L = min(__ffs(q) - 1, 3);
  while (L > 0 && (q >> L) < sp.pmin) --L;
  const int g = q >> L;
  // per-level constants, one level per lane, then broadcast
  HierLevels lv;
  {
    const int l = lane & 3;
    const int P = g << l;
    const int M = N / P;
    const int my_r0 = N - M * P;
    const double my_hi = 1.0 / (double)(M + 1), my_lo = 1.0 / (double)M;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      lv.r0[i] = __shfl_sync(0xffffffffu, my_r0, i);
      lv.inv_hi[i] = __shfl_sync(0xffffffffu, my_hi, i);
      lv.inv_lo[i] = __shfl_sync(0xffffffffu, my_lo, i);
    }
  }
  const int rows = (N + g - 1) / g;
  const bool chain = (L == 3) && !(g & 1) && (g >> 1) >= sp.pmin;
  double E[4] = {0.0, 0.0, 0.0, 0.0};
  switch (L) {
    case 0: hier_top_fold<0>(sp.xs, g, rows, lv, E, nullptr); break;
    case 1: hier_top_fold<1>(sp.xs, g, rows, lv, E, nullptr); break;
    case 2: hier_top_fold<2>(sp.xs, g, rows, lv, E, nullptr); break;
    default: hier_top_fold<3>(sp.xs, g, rows, lv, E, chain ? scr : nullptr); break;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) E[i] = warp_sum(E[i]);
  {
    const int l = lane & 3;
    const double myE = l == 0 ? E[0] : (l == 1 ? E[1] : (l == 2 ? E[2] : E[3]));
    double val = sqrt(myE) / sp.sqrtN;
    if (sp.metric == PP_METRIC_GAMMA) val = val / sqrt((double)(g << l));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double vi = __shfl_sync(0xffffffffu, val, i);
      if (i <= L) consider(sp, vi, g << i, bestv, bestp);
    }
  }
  if (chain) {
    __syncwarp();
    double* src = scr;
    double* dst = scr + ((g + 1) & ~1);
    int h = g;
    while (!(h & 1) && (h >> 1) >= sp.pmin) {
      const int h2 = h >> 1;
      const int M = N / h2, r0 = N - M * h2;
      const double hi = 1.0 / (double)(M + 1), lo = 1.0 / (double)M;
      double e = 0.0;
      for (int r = lane; r < h2; r += 32) {
        const double v = src[r] + src[r + h2];
        dst[r] = v;
        e = fma(v * v, r < r0 ? hi : lo, e);
      }
      __syncwarp();
      e = warp_sum(e);
      double val = sqrt(e) / sp.sqrtN;
      if (sp.metric == PP_METRIC_GAMMA) val = val / sqrt((double)h2);
      consider(sp, val, h2, bestv, bestp);
      double* t = src;
      src = dst;
      dst = t;
      h = h2;
    }
    __syncwarp();
  }
}

// Shared scratch the sweep needs (one per CTA).
struct SweepShared {
  int counter;            // next candidate index
  int hit_p;              // first-hit mode: lowest period over threshold so far
  double wval[kWarps];
  int wp[kWarps];
};

// Sweep all candidates with dynamic (atomic-counter) distribution over the CTA's warps.
//   argmax mode (thresh < 0): strict '>' from 0, lowest p on ties, periods in `skip` ignored
//                             (Periods.py:512-515).
//   first-hit mode (thresh >= 0): lowest p whose metric > thresh; warps stop once their next
//                             candidate lies above the current hit (Periods.py:273-286).
// All threads call; result valid in all threads.  Contains CTA barriers.
__device__ __forceinline__ SweepResult cta_sweep(const SweepParams& sp, SweepShared* sh) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const bool first_hit = sp.thresh >= 0.0;
  if (threadIdx.x == 0) {
    sh->counter = 0;
    sh->hit_p = 0x7fffffff;
  }
  __syncthreads();
  double bestv = 0.0;
  int bestp = 0;
  const bool hier = sp.hier_scr != nullptr && !first_hit && !sp.orth && !sp.trunc &&
                    (sp.metric == PP_METRIC_NORM || sp.metric == PP_METRIC_GAMMA);
  if (hier) {
    // tops: candidates p with 2p > pmax; everything else is derived from exactly one of them
    const int top_lo = max(sp.pmin, (sp.pmax >> 1) + 1);
    const int ntops = sp.pmax - top_lo + 1;
    double* scr = sp.hier_scr + (size_t)wid * sp.hier_len;
    while (true) {
      int idx = 0;
      if (lane == 0) idx = atomicAdd(&sh->counter, 1);
      idx = __shfl_sync(0xffffffffu, idx, 0);
      if (idx >= ntops) break;
      warp_hier_top(sp, sp.pmax - idx, scr, bestv, bestp);
    }
  }
  const int ncand = hier ? 0 : sp.pmax - sp.pmin + 1;
  while (true) {
    int idx = 0;
    if (lane == 0) idx = atomicAdd(&sh->counter, 1);
    idx = __shfl_sync(0xffffffffu, idx, 0);
    if (idx >= ncand) break;
    const int p = sp.pmin + idx;
    if (first_hit) {
      const int hp = *reinterpret_cast<volatile int*>(&sh->hit_p);
      if (p > hp) break;
    }
    const double val = warp_period_metric_any(sp, p);
    if (sp.metric_out != nullptr && lane == 0) sp.metric_out[p] = val;
    if (first_hit) {
      if (val > sp.thresh) {
        if (bestp == 0 || p < bestp) {
          bestp = p;
          bestv = val;
        }
        if (lane == 0) atomicMin(&sh->hit_p, p);
      }
    } else {
      const bool skipped = sp.skip != nullptr && ((sp.skip[p >> 5] >> (p & 31)) & 1u);
      if (!skipped && (val > bestv || (val == bestv && bestp != 0 && p < bestp))) {
        bestv = val;
        bestp = p;
      }
    }
  }
  if (lane == 0) {
    sh->wval[wid] = bestv;
    sh->wp[wid] = bestp;
  }
  __syncthreads();
  SweepResult res{0.0, 0};
#pragma unroll
  for (int w = 0; w < kWarps; ++w) {
    const double v = sh->wval[w];
    const int q = sh->wp[w];
    if (q == 0) continue;
    if (first_hit) {
      if (res.p == 0 || q < res.p) res = SweepResult{v, q};
    } else if (v > res.val || (v == res.val && res.p != 0 && q < res.p)) {
      res = SweepResult{v, q};
    }
  }
  __syncthreads();  // wval/wp may be rewritten by the next sweep
  return res;
}

}  // namespace pp
