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
  const int ncand = sp.pmax - sp.pmin + 1;
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
