// Cyclic two-sided Jacobi eigen-solver for a symmetric PSD matrix held in shared memory.
// This is the "batched Jacobi/eigen kernel held in shared memory" of the north star: it yields every
// eigenvalue (= squared singular value of the matrix the Gram was built from), so the trailing ones
// are summed directly instead of being obtained as 1 - top4/total (splitp/phylogenetics.py:293-300
// computes the same quantity from LAPACK singular values).
#pragma once
#include <stdint.h>

namespace spb {

constexpr int kJacobiMaxK = 128;
constexpr int kJacobiMaxSweeps = 40;

struct JacobiScratch {  // lives in shared memory
  double c[kJacobiMaxK / 2], s[kJacobiMaxK / 2], npp[kJacobiMaxK / 2], nqq[kJacobiMaxK / 2];
  int p[kJacobiMaxK / 2], q[kJacobiMaxK / 2];
  int rotated;
  double trace0;
};

// All threads of the CTA call this.  A: k x k symmetric, row stride lda (odd lda avoids bank
// conflicts).  If V != nullptr it must hold the identity on entry and receives the eigenvectors as
// columns (A_in = V diag V^T).  On return the eigenvalues are on the diagonal of A (unsorted).
__device__ inline void jacobi_eig_smem(double* A, int lda, int k, double* V, int ldv, JacobiScratch* js) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int m = (k + 1) & ~1;  // players in the round-robin (one dummy when k is odd)
  const int np = m >> 1;
  if (tid == 0) {
    double tr = 0.0;
    for (int i = 0; i < k; ++i) tr += fabs(A[i * lda + i]);
    js->trace0 = tr;
  }
  __syncthreads();
  if (k < 2) return;
  const double abs_floor = 1e-19 * js->trace0;
  for (int sweep = 0; sweep < kJacobiMaxSweeps; ++sweep) {
    if (tid == 0) js->rotated = 0;
    __syncthreads();
    for (int r = 0; r < m - 1; ++r) {
      // 1. rotation parameters for the np disjoint pairs of this round
      if (tid < np) {
        int p, q;
        if (tid == 0) { p = r; q = m - 1; }
        else { p = (r + tid) % (m - 1); q = (r - tid + m - 1) % (m - 1); }
        if (p > q) { int t = p; p = q; q = t; }
        double c = 1.0, s = 0.0, npp = 0.0, nqq = 0.0;
        if (q < k) {
          double app = A[p * lda + p], aqq = A[q * lda + q], apq = A[p * lda + q];
          npp = app; nqq = aqq;
          double thr = fmax(2.0e-16 * sqrt(fabs(app * aqq)), abs_floor);
          if (fabs(apq) > thr) {
            double tau = (aqq - app) / (2.0 * apq);
            double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + t * t);
            s = t * c;
            npp = app - t * apq;
            nqq = aqq + t * apq;
            js->rotated = 1;
          } else if (apq != 0.0) {
            // below threshold: drop the entry so it cannot accumulate
            A[p * lda + q] = 0.0; A[q * lda + p] = 0.0;
          }
        } else {
          q = -1;  // dummy pair
        }
        js->p[tid] = p; js->q[tid] = q; js->c[tid] = c; js->s[tid] = s; js->npp[tid] = npp; js->nqq[tid] = nqq;
      }
      __syncthreads();
      // 2. columns: A <- A J  (and V <- V J)
      for (int idx = tid; idx < np * k; idx += nt) {
        int t = idx / k, row = idx - t * k;
        int q = js->q[t];
        double s = js->s[t];
        if (q < 0 || s == 0.0) continue;
        int p = js->p[t];
        double c = js->c[t];
        double x = A[row * lda + p], y = A[row * lda + q];
        A[row * lda + p] = c * x - s * y;
        A[row * lda + q] = s * x + c * y;
        if (V) {
          double vx = V[row * ldv + p], vy = V[row * ldv + q];
          V[row * ldv + p] = c * vx - s * vy;
          V[row * ldv + q] = s * vx + c * vy;
        }
      }
      __syncthreads();
      // 3. rows: A <- J^T A, with the rotated 2x2 block written in closed form
      for (int idx = tid; idx < np * k; idx += nt) {
        int t = idx / k, col = idx - t * k;
        int q = js->q[t];
        double s = js->s[t];
        if (q < 0 || s == 0.0) continue;
        int p = js->p[t];
        double c = js->c[t];
        if (col == p) { A[p * lda + p] = js->npp[t]; A[q * lda + p] = 0.0; }
        else if (col == q) { A[p * lda + q] = 0.0; A[q * lda + q] = js->nqq[t]; }
        else {
          double x = A[p * lda + col], y = A[q * lda + col];
          A[p * lda + col] = c * x - s * y;
          A[q * lda + col] = s * x + c * y;
        }
      }
      __syncthreads();
    }
    if (!js->rotated) break;
    __syncthreads();
  }
}

// Descending rank sort of the diagonal of A into out[0..k).  All threads call; scratch `tmp` k doubles.
__device__ inline void sort_diag_desc(const double* A, int lda, int k, double* tmp, double* out) {
  for (int i = threadIdx.x; i < k; i += blockDim.x) tmp[i] = A[i * lda + i];
  __syncthreads();
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    double v = tmp[i];
    int rank = 0;
    for (int j = 0; j < k; ++j) {
      double w = tmp[j];
      rank += (w > v || (w == v && j < i)) ? 1 : 0;
    }
    out[rank] = v;
  }
  __syncthreads();
}

// score = sqrt( sum_{i>=4} lam_i / sum_i lam_i ), eigenvalues sorted descending, negatives (rounding of
// exact zeros) clamped.  K = 4 is hard-coded in the reference (phylogenetics.py:293-300).
__device__ inline double score_from_sorted(const double* lam, int k) {
  double top = 0.0, tail = 0.0;
  for (int i = 0; i < k && i < 4; ++i) top += fmax(lam[i], 0.0);
  for (int i = k - 1; i >= 4; --i) tail += fmax(lam[i], 0.0);  // small to large
  double tot = top + tail;
  if (!(tot > 0.0)) return nan("");  // 0/0 in the reference
  return sqrt(tail / tot);
}

}  // namespace spb
