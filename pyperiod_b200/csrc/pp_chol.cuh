// pyperiod_b200 -- normal equations of a periodic dictionary on the FP64 tensor cores.
//
// The dictionary A of QOPeriods.get_subspaces (pyPeriod/QOPeriods.py:807-852) stacks, per period q, the first
// rows_q indicator rows 1[n = i (mod q)].  Its Gram matrix G = A A^T (QOPeriods.py:781) is integer valued and never
// stored: G[(a,i),(b,j)] = #{n < N : n = i (mod q_a), n = j (mod q_b)} has a closed form (Chinese remainder
// theorem), evaluated where the factorisation consumes it.  What is stored is the Cholesky factor L only, as
// 32 x 32 blocks of 8 KB each (block row by block row, row major inside a block): the 16 rows x 32 columns a warp
// needs per k-chunk are 4 KB of CONTIGUOUS memory and the 32 x 32 block of block row j that every warp shares is one
// 8 KB segment -- the factors of the concurrent windows do not fit L2, and with whole rows stored contiguously
// (23 KB apart at R = 3000) every chunk opened a new DRAM page per row for 256 bytes.  Each diagonal block is
// replaced by its INVERSE (the panel below a diagonal block, the forward substitution and the back substitution
// all multiply by it; the block itself is never needed again).
//
// Factorisation: left-looking by block columns.  For block column j (32 columns) the CTA computes
//     P = G[rows, j] - L[rows, 0:j) L[j, 0:j)^T      (DMMA m8n8k4, FP64 tensor cores)
// for all rows below (16 m-tiles of 8 rows per pass, two per warp), factors the 32 x 32 diagonal block with one
// warp, and multiplies the rest by the inverse block -- a second DMMA whose A operand is the accumulator of the
// first (the C fragment of m8n8k4 is an A fragment with a permuted k index, so nothing moves).  The A operand of
// the big product comes straight from L2 in its own fragment layout (four 16-byte loads per thread and 32-k
// chunk, fully used sectors); only the 32 rows of block row j, shared by every warp, are staged in shared
// memory (cp.async, double buffered).  The right-hand side rides along as one more row ("row R"), which makes the
// forward substitution L y = W part of the factorisation.  Rows below `row_lo` are final from an earlier call
// (QOPeriods re-solves a growing dictionary every round): only the new rows are computed.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "pp_common.cuh"

namespace pp {

constexpr int kCb = 32;        // Cholesky block size
constexpr int kLdStage = 40;   // doubles per row of a staged 32 x 32 block: 16-byte fragment reads are conflict free
constexpr int kLdD = 33;       // doubles per row of the diagonal-block scratch

// packed factor: offset of the 32 x 32 block (block row bi, block column bk <= bi), of element (r, k), and the total
// length for R rows
__host__ __device__ inline size_t chol_blk(int bi, int bk) {
  return (size_t)1024 * ((size_t)bi * (size_t)(bi + 1) / 2 + (size_t)bk);
}
__host__ __device__ inline size_t chol_at(int r, int k) {
  return chol_blk(r >> 5, k >> 5) + (size_t)((r & 31) * 32 + (k & 31));
}
__host__ __device__ inline size_t chol_packed_len(int R) {
  const size_t nb = (size_t)((R + 31) >> 5);
  return 512 * nb * (nb + 1);
}

// ------------------------------------------------------------------------------------------
// the dictionary, as the Gram matrix sees it
// ------------------------------------------------------------------------------------------
struct DictView {
  const int* q;      // shared: period of entry a
  const int* rows;   // shared: rows of entry a present in A
  const int* off;    // shared: first row of entry a, [n + 1]
  int n;             // entries
  int N;             // window length
};

__device__ __forceinline__ int dict_entry_of(const DictView& dv, int r) {
  int a = 0;
  while (a + 1 < dv.n && dv.off[a + 1] <= r) ++a;
  return a;
}

// G[r][r] = number of samples n < N in the residue class of row r
__device__ __forceinline__ double gram_diagonal(const DictView& dv, int r) {
  const int a = dict_entry_of(dv, r);
  return (double)((dv.N - 1 - (r - dv.off[a])) / dv.q[a] + 1);
}

constexpr int kLdTile = 40;   // ints per row of the Gram tile (128 rows x 32 columns of counts)

// Gram entries of 128 rows x 32 columns, as integer counts in shared memory: tile[slot][c - j0] = G[row0 + slot][c].
// G[(a,i),(b,j)] = #{n < N : n = i (mod q_a), n = j (mod q_b)}: two threads walk the residue class of a row
// (n = i, i + q_a, ...) and, for each dictionary entry that owns columns of the block, step the residue of n modulo
// q_b along with it (one add and one conditional subtract per sample; no division, no table of modular inverses) and
// count the samples whose residue falls on a column of the block.  `tile` must be zero on entry.  No barrier inside.
__device__ __forceinline__ void gram_tile_count(const DictView& dv, int R, int row0, int j0, int* tile) {
  const int slot = threadIdx.x >> 1, half = threadIdx.x & 1;
  const int r = row0 + slot;
  if (r >= R) return;
  const int a = dict_entry_of(dv, r);
  const int i = r - dv.off[a], qa = dv.q[a];
  const int terms = (dv.N - 1 - i) / qa + 1;
  const int k0 = half ? terms >> 1 : 0, k1 = half ? terms : terms >> 1;
  if (k1 <= k0) return;
  const int c_end = min(j0 + kCb, R);
  int* trow = tile + slot * kLdTile;
  for (int b = dict_entry_of(dv, j0); b < dv.n && dv.off[b] < c_end; ++b) {
    const int qb = dv.q[b], ob = dv.off[b];
    const int lo = max(j0, ob) - ob, hi = min(c_end, dv.off[b + 1]) - ob;   // residues of entry b inside the block
    const int shift = ob - j0;                                               // column (relative to j0) = shift + residue
    int j = (i + qa * k0) % qb;
    const int step = qa % qb;
    for (int k = k0; k < k1; ++k) {
      if (j >= lo && j < hi) atomicAdd(trow + shift + j, 1);
      j += step;
      if (j >= qb) j -= qb;
    }
  }
}

// ------------------------------------------------------------------------------------------
// staging + tensor-core helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void chol_cp_async_16(void* dst_smem, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void chol_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void chol_cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void chol_dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

struct CholStage {   // shared memory, carved by the caller
  double* Bs;    // [2][32 * kLdStage]  block row j of the factor, 32-k chunks
  double* D;     // [32 * kLdD]         diagonal block
  double* Li;    // [32 * kLdD]         its inverse
  double* rD;    // [32]
  int* flag;
};
constexpr size_t kCholStageBs = (size_t)2 * kCb * kLdStage * 8;             // bytes of Bs
constexpr size_t kCholStageD = (size_t)(2 * kCb * kLdD + kCb) * 8 + 16;     // bytes of D, Li, rD, flag

// Cholesky of the 32 x 32 block in st.D (lower triangle; rows >= nb are identity padding) by warp 0:
// st.D <- factor, st.Li <- inverse of the factor (row major).  thr: pivots must exceed it (lane k: row k).
static __device__ __noinline__ void warp_factor_block(const CholStage st, double thr) {
  const int lane = threadIdx.x & 31;
  // lane owns row `lane` of the block in registers; the scaled column k travels by shuffle (no shared-memory
  // round trip, no warp barrier inside the 32 steps)
  double r[kCb];
#pragma unroll
  for (int c = 0; c < kCb; ++c) r[c] = st.D[lane * kLdD + c];
  bool ok = true;
  double rdiag = 1.0;   // 1 / L[lane][lane]
#pragma unroll
  for (int k = 0; k < kCb; ++k) {
    const double dkk = __shfl_sync(0xffffffffu, r[k], k);
    const double tk = __shfl_sync(0xffffffffu, thr, k);
    if (!(dkk > tk)) ok = false;   // uniform: every lane sees the same pivot
    const double rs = rsqrt(ok ? dkk : 1.0);
    const double l = (lane > k) ? r[k] * rs : (lane == k ? (ok ? dkk : 1.0) * rs : 0.0);
    r[k] = l;
    if (lane == k) rdiag = rs;
#pragma unroll
    for (int j = k + 1; j < kCb; ++j) {
      const double lj = __shfl_sync(0xffffffffu, l, j);   // L[j][k]
      r[j] = fma(-l, lj, r[j]);                           // only j <= lane is ever used
    }
  }
  if (!ok && lane == 0) *st.flag = 1;
#pragma unroll
  for (int c = 0; c < kCb; ++c) st.D[lane * kLdD + c] = (c <= lane) ? r[c] : 0.0;
  __syncwarp();
  // X = L^-1 (lower triangular), row by row: X[i][:] = (e_i - sum_{m < i} L[i][m] X[m][:]) / L[i][i].  Lane c owns
  // column c of X; L[i][m] is a broadcast read of the factor just written, the sum over m runs on two accumulators.
  double xcol[kCb];
#pragma unroll
  for (int i = 0; i < kCb; ++i) {
    double v0 = (i == lane) ? 1.0 : 0.0, v1 = 0.0;
#pragma unroll
    for (int m = 0; m < i; m += 2) {
      v0 = fma(-st.D[i * kLdD + m], xcol[m], v0);
      if (m + 1 < i) v1 = fma(-st.D[i * kLdD + m + 1], xcol[m + 1], v1);
    }
    const double di = __shfl_sync(0xffffffffu, rdiag, i);
    xcol[i] = (i >= lane) ? (v0 + v1) * di : 0.0;
  }
#pragma unroll
  for (int rr = 0; rr < kCb; ++rr) st.Li[rr * kLdD + lane] = xcol[rr];
}

// ------------------------------------------------------------------------------------------
// factorisation + forward substitution
// ------------------------------------------------------------------------------------------
// L: packed factor (global).  R rows.  Block rows below row_lo (a multiple of 32) already hold the factor of the
// same leading rows.  y (shared, >= 32 * ceil(R / 32) + 32 doubles): right-hand side W on entry (entries past R
// must be finite), L^-1 W on exit.  Returns false (uniformly) when a pivot is not positive: the reference's
// np.linalg.solve raises LinAlgError on a singular matrix (QOPeriods.py:794).  All threads call.
static __device__ __noinline__ bool cta_chol_factor(double* __restrict__ L, int R, int row_lo, const DictView dv,
                                                    double* y, const CholStage st, long long* tph) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  const int nbk = (R + kCb - 1) / kCb;
  const int Rv = nbk * kCb;  // index of the virtual row that carries the right-hand side
  if (tid == 0) *st.flag = 0;
  for (int jb = 0; jb < nbk; ++jb) {
    const int j0 = jb * kCb;
    const bool diag_new = j0 >= row_lo;
    const int first = diag_new ? j0 : row_lo;         // first row computed at this block column
    const int ntile = ((Rv - first) >> 3) + 1;        // m-tiles of 8 rows, the last one holds the virtual row
    const double* Lj = L + chol_blk(jb, 0);           // block row j: jb + 1 blocks of 32 x 32, one after the other
    if (!diag_new) {  // inverse of a diagonal block factored by an earlier call
      for (int idx = tid; idx < kCb * kCb; idx += kThreads) {
        const int r = idx >> 5, c = idx & 31;
        st.Li[r * kLdD + c] = Lj[(size_t)jb * 1024 + r * kCb + c];
      }
    }
    __syncthreads();
    for (int t0 = 0; t0 < ntile; t0 += 2 * kWarps) {
      const int mt0 = t0 + 2 * wid;
      // rows of this warp's two m-tiles: 0 = nothing (padding / past the end), 1 = factor row, 2 = right-hand side
      int kind[2], row[2], kstr[2];
      const double* rp[2];   // element (row, 32 kc + c) of the A operand is rp[kc * kstr + c]
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int r = first + 8 * (mt0 + mt) + lr;
        row[mt] = r;
        kind[mt] = (mt0 + mt < ntile) ? (r < R ? 1 : (r == Rv ? 2 : 0)) : 0;
        rp[mt] = kind[mt] == 1 ? L + chol_blk(r >> 5, 0) + (r & 31) * kCb : y;
        kstr[mt] = kind[mt] == 1 ? 1024 : kCb;
      }
      double acc[2][4][2];
      const long long t_init = clock64();
      {
        // Gram entries of the pass (integer counts, built in the staging area), the right-hand side on the virtual
        // row, identity padding past the last row of the last diagonal block
        int* tile = reinterpret_cast<int*>(st.Bs);
        for (int idx = tid; idx < 16 * kWarps * kLdTile; idx += kThreads) tile[idx] = 0;
        __syncthreads();
        gram_tile_count(dv, R, first + 8 * t0, j0, tile);
        __syncthreads();
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int r = row[mt];
          const int* trow = tile + (16 * wid + 8 * mt + lr) * kLdTile + 2 * lc;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int c = j0 + 8 * nt + 2 * lc + e;
              double v = 0.0;
              if (kind[mt] == 1) {
                v = (double)trow[8 * nt + e];   // columns past R were never counted: 0
              } else if (kind[mt] == 2) {
                if (c < R) v = y[c];
              } else if (mt0 + mt < ntile && r == c) {
                v = 1.0;  // identity padding of the last diagonal block
              }
              acc[mt][nt][e] = v;
            }
        }
        __syncthreads();  // the tile lives where the factor blocks are staged next
      }
      if (tph) tph[0] += clock64() - t_init;
      // P -= L[rows, 0:j0) L[j, 0:j0)^T, 32 columns of k at a time
      auto stage = [&](int buf, int kc) {
        for (int idx = tid; idx < kCb * 16; idx += kThreads) {
          const int n = idx >> 4, piece = idx & 15;
          const bool live = j0 + n < R;
          const double* src = live ? Lj + (size_t)kc * 1024 + n * kCb + 2 * piece : Lj;
          chol_cp_async_16(st.Bs + buf * (kCb * kLdStage) + n * kLdStage + 2 * piece, src, live ? 16 : 0);
        }
        chol_cp_async_commit();
      };
      const bool busy = mt0 < ntile;   // warp-uniform: this warp owns at least one m-tile of the pass
      if (jb > 0) stage(0, 0);
      for (int kc = 0; kc < jb; ++kc) {
        const int buf = kc & 1;
        if (kc + 1 < jb) {
          stage(buf ^ 1, kc + 1);
          chol_cp_async_wait<1>();
        } else {
          chol_cp_async_wait<0>();
        }
        __syncthreads();
        if (busy) {
        double a[2][4][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          if (kind[mt] != 0) {
            const double2* p = reinterpret_cast<const double2*>(rp[mt] + (size_t)kc * kstr[mt] + 2 * lc);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              const double2 d2 = p[4 * v];
              a[mt][v][0] = -d2.x;
              a[mt][v][1] = -d2.y;
            }
          } else {
#pragma unroll
            for (int v = 0; v < 4; ++v) a[mt][v][0] = a[mt][v][1] = 0.0;
          }
        }
        const double* Bs = st.Bs + buf * (kCb * kLdStage);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          double2 b[4];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
            b[nt] = *reinterpret_cast<const double2*>(Bs + (8 * nt + lr) * kLdStage + 8 * v + 2 * lc);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              chol_dmma(acc[mt][nt], a[mt][v][0], b[nt].x);
              chol_dmma(acc[mt][nt], a[mt][v][1], b[nt].y);
            }
        }
        }
        __syncthreads();  // everyone is done with this buffer before it is restaged
      }
      const bool diag_pass = diag_new && t0 == 0;
      if (diag_pass) {
        // tiles 0..3 are the diagonal block (warps 0 and 1): factor it, keep its inverse
        if (mt0 < 4) {
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
              for (int e = 0; e < 2; ++e)
                st.D[(8 * (mt0 + mt) + lr) * kLdD + 8 * nt + 2 * lc + e] = acc[mt][nt][e];
        }
        const long long t_diag = clock64();
        __syncthreads();
        if (wid == 0) {
          // a pivot at rounding level (relative to its diagonal count) means the dictionary is rank deficient
          const int r = j0 + lane;
          const double thr = r < R ? 1e-14 * gram_diagonal(dv, r) : 0.5;
          warp_factor_block(st, thr);
        }
        __syncthreads();
        if (tph) tph[1] += clock64() - t_diag;
        if (*st.flag) return false;
        double* Ljw = L + chol_blk(jb, jb);
        for (int idx = tid; idx < kCb * kCb; idx += kThreads) {
          const int r = idx >> 5, c = idx & 31;
          Ljw[idx] = st.Li[r * kLdD + c];
        }
      }
      // rows below the diagonal block: multiply by the inverse block, out = P Li^T.  The accumulator is the A
      // fragment: thread (lr, lc) holds P[lr][8 nt + 2 lc + e], i.e. k index 8 nt + 2 lc + e at step (nt, e).
      if (busy && !(diag_pass && mt0 < 4)) {
        double out[2][4][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) out[mt][nt][0] = out[mt][nt][1] = 0.0;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int nt2 = nt; nt2 < 4; ++nt2) {   // Li is lower triangular: k-group nt only reaches columns >= 8 nt
              const double bv = st.Li[(8 * nt2 + lr) * kLdD + 8 * nt + 2 * lc + e];
#pragma unroll
              for (int mt = 0; mt < 2; ++mt) chol_dmma(out[mt][nt2], acc[mt][nt][e], bv);
            }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          if (kind[mt] == 1) {
            double* p = L + chol_blk(row[mt] >> 5, jb) + (row[mt] & 31) * kCb + 2 * lc;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
              *reinterpret_cast<double2*>(p + 8 * nt) = make_double2(out[mt][nt][0], out[mt][nt][1]);
          } else if (kind[mt] == 2) {
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int c = j0 + 8 * nt + 2 * lc + e;
                if (c < R) y[c] = out[mt][nt][e];
              }
          }
        }
      }
    }
    __syncthreads();  // st.Li and y are rewritten by the next block column
  }
  return true;
}

// ------------------------------------------------------------------------------------------
// triangular solves with the packed factor (diagonal blocks hold their inverses)
// ------------------------------------------------------------------------------------------
// In place: v <- L^-T v.  v: shared or global, entries [R, 32 * ceil(R / 32)) must be finite.  All threads call.
static __device__ __noinline__ void cta_chol_backward(const double* L, int R, double* v,
                                                      const CholStage st) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nbk = (R + kCb - 1) / kCb;
  for (int jb = nbk - 1; jb >= 0; --jb) {
    const int j0 = jb * kCb;
    const double* Lj = L + chol_blk(jb, 0);
    for (int idx = tid; idx < kCb * kCb; idx += kThreads) {
      const int r = idx >> 5, c = idx & 31;
      st.Li[r * kLdD + c] = Lj[(size_t)jb * 1024 + idx];
    }
    __syncthreads();
    if (wid == 0) {  // w_block = Li^T t
      double s = 0.0;
#pragma unroll 8
      for (int m = 0; m < kCb; ++m) {
        const double t = (j0 + m < R) ? v[j0 + m] : 0.0;
        s = fma(st.Li[m * kLdD + lane], t, s);   // Li[m][lane] = 0 for m < lane
      }
      __syncwarp();
      if (j0 + lane < R) v[j0 + lane] = s;
      st.rD[lane] = (j0 + lane < R) ? s : 0.0;
    }
    __syncthreads();
    // v[k] -= sum_c L[j0 + c][k] w[c] for every k left of the block (row access: coalesced over k)
    for (int k = tid; k < j0; k += kThreads) {
      const double* col = Lj + (size_t)(k >> 5) * 1024 + (k & 31);   // L[j0 + c][k] = col[32 c]
      double s0 = 0.0, s1 = 0.0;
#pragma unroll 8
      for (int c = 0; c < kCb; c += 2) {
        s0 = fma(col[c * kCb], st.rD[c], s0);
        s1 = fma(col[(c + 1) * kCb], st.rD[c + 1], s1);
      }
      v[k] -= s0 + s1;
    }
    __syncthreads();
  }
}

// In place: v <- L^-1 v (used by the refinement step; the first right-hand side is solved inside the
// factorisation).  All threads call.
static __device__ __noinline__ void cta_chol_forward(const double* L, int R, double* v,
                                                     const CholStage st) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nbk = (R + kCb - 1) / kCb;
  for (int jb = 0; jb < nbk; ++jb) {
    const int j0 = jb * kCb;
    const double* Lj = L + chol_blk(jb, 0);
    for (int idx = tid; idx < kCb * kCb; idx += kThreads) {
      const int r = idx >> 5, c = idx & 31;
      st.Li[r * kLdD + c] = Lj[(size_t)jb * 1024 + idx];
    }
    // t[c] = v[j0 + c] - sum_{k < j0} L[j0 + c][k] v[k]: one warp per row, lanes over k
    for (int c = wid; c < kCb; c += kWarps) {
      double s = 0.0;
      if (j0 + c < R)
        for (int k = lane; k < j0; k += 32) s = fma(Lj[(size_t)(k >> 5) * 1024 + c * kCb + (k & 31)], v[k], s);
      s = warp_sum(s);
      if (lane == 0) st.rD[c] = (j0 + c < R) ? v[j0 + c] - s : 0.0;
    }
    __syncthreads();
    if (wid == 0) {  // block = Li t
      double s = 0.0;
#pragma unroll 8
      for (int m = 0; m < kCb; ++m) s = fma(st.Li[lane * kLdD + m], st.rD[m], s);   // Li[lane][m] = 0 for m > lane
      if (j0 + lane < R) v[j0 + lane] = s;
    }
    __syncthreads();
  }
}

}  // namespace pp
