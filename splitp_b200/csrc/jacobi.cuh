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

template <int MAXK>
struct JacobiScratchT {  // lives in shared memory; MAXK = largest matrix dimension the owner kernel solves
  static constexpr int kPairs = (MAXK + 1) / 2;
  static constexpr int kBlocks = kPairs * (kPairs + 1) / 2;
  double c[kPairs], s[kPairs], npp[kPairs], nqq[kPairs];
  int p[kPairs], q[kPairs];
  int rotated;
  double trace0;
  uint8_t bt1[kBlocks], bt2[kBlocks];  // (t1 <= t2) of every 2x2 block: an even deal over the threads
};
using JacobiScratch = JacobiScratchT<kJacobiMaxK>;

// Storage contract: the solver works on an even dimension m = jacobi_dim(k) (one zero row/column of padding when
// k is odd), row stride lda >= m (jacobi_ld(k) = m | 1 is odd, which keeps the 2x2-block accesses spread over the
// banks).  The CALLER zero-fills row/column k when k is odd.
__host__ __device__ constexpr int jacobi_dim(int k) { return (k + 1) & ~1; }
__host__ __device__ constexpr int jacobi_ld(int k) { return jacobi_dim(k) | 1; }

// All threads of the CTA call this.  A: k x k symmetric in the storage described above.  If V != nullptr it must
// hold the identity (jacobi_dim(k) rows) on entry and receives the eigenvectors as columns (A_in = V diag V^T).
// On return the eigenvalues are on the diagonal of A (unsorted).
//
// One round = the m/2 disjoint pairs of a round-robin tournament rotated at once:
//   (1) m/2 threads compute the rotation (c, s) of their pair from the current 2x2 diagonal block,
//   (2) every 2x2 block B(t1, t2) = A[{p1,q1}][{p2,q2}], t1 <= t2, is replaced by J1^T B J2 in ONE pass (the block
//       and its mirror image are written by the same thread, so the update is in place with a single barrier);
//       the diagonal blocks are written in closed form (npp, 0; 0, nqq).
template <class Scratch>
__device__ inline void jacobi_eig_smem(double* A, int lda, int k, double* V, int ldv, Scratch* js) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;  // used by the eigenvector update
  const int m = jacobi_dim(k);
  const int np = m >> 1;
  const int nblocks = np * (np + 1) / 2;
  if (tid == 0) {
    double tr = 0.0;
    for (int i = 0; i < k; ++i) tr += fabs(A[i * lda + i]);
    js->trace0 = tr;
  }
  for (int t1 = tid; t1 < np; t1 += nt) {  // row t1 of the block triangle starts at t1 * np - t1 (t1 - 1) / 2
    const int base = t1 * np - (t1 * (t1 - 1)) / 2;
    for (int t2 = t1; t2 < np; ++t2) { js->bt1[base + t2 - t1] = (uint8_t)t1; js->bt2[base + t2 - t1] = (uint8_t)t2; }
  }
  __syncthreads();
  if (k < 2) return;
  const double abs_floor = 1e-19 * js->trace0;
  for (int sweep = 0; sweep < kJacobiMaxSweeps; ++sweep) {
    if (tid == 0) js->rotated = 0;
    __syncthreads();
    for (int r = 0; r < m - 1; ++r) {
      // (1) rotation parameters of the np disjoint pairs of this round
      if (tid < np) {
        int p, q;
        if (tid == 0) { p = r; q = m - 1; }
        else { p = (r + tid) % (m - 1); q = (r - tid + m - 1) % (m - 1); }
        if (p > q) { int t = p; p = q; q = t; }
        double app = A[p * lda + p], aqq = A[q * lda + q], apq = A[p * lda + q];
        double c = 1.0, s = 0.0, npp = app, nqq = aqq;
        // rotate when |apq| > max(2e-16 sqrt|app aqq|, floor), compared on squares (no square root on this path)
        const double apq2 = apq * apq;
        if (apq2 > fmax(4.0e-32 * fabs(app * aqq), abs_floor * abs_floor)) {
          // t = tan(theta) of the smaller rotation angle: 2 apq / (d + sign(d) sqrt(d^2 + 4 apq^2)), d = aqq - app
          // (the usual sign(tau) / (|tau| + sqrt(1 + tau^2)) with one square root and one division)
          const double d = aqq - app;
          const double h = sqrt(fma(d, d, 4.0 * apq2));
          const double t = (2.0 * apq) / (d >= 0.0 ? d + h : d - h);
          c = rsqrt(fma(t, t, 1.0));
          s = t * c;
          npp = app - t * apq;
          nqq = aqq + t * apq;
          js->rotated = 1;
        } else if (apq != 0.0) {
          // below threshold: drop the entry so it cannot accumulate
          A[p * lda + q] = 0.0; A[q * lda + p] = 0.0;
        }
        js->p[tid] = p; js->q[tid] = q; js->c[tid] = c; js->s[tid] = s; js->npp[tid] = npp; js->nqq[tid] = nqq;
      }
      __syncthreads();
      // (2) fused two-sided update, one 2x2 block per thread iteration (blocks dealt evenly over all threads)
      for (int b = tid; b < nblocks; b += nt) {
        const int t1 = js->bt1[b], t2 = js->bt2[b];
        const double c1 = js->c[t1], s1 = js->s[t1], c2 = js->c[t2], s2 = js->s[t2];
        if (s1 == 0.0 && s2 == 0.0) continue;
        const int p1 = js->p[t1], q1 = js->q[t1];
        if (t1 == t2) {
          A[p1 * lda + p1] = js->npp[t1]; A[q1 * lda + q1] = js->nqq[t1];
          A[p1 * lda + q1] = 0.0; A[q1 * lda + p1] = 0.0;
          continue;
        }
        const int p2 = js->p[t2], q2 = js->q[t2];
        double a = A[p1 * lda + p2], bb = A[p1 * lda + q2], cc = A[q1 * lda + p2], d = A[q1 * lda + q2];
        double a1 = c1 * a - s1 * cc, b1 = c1 * bb - s1 * d, cc1 = s1 * a + c1 * cc, d1 = s1 * bb + c1 * d;
        double a2 = c2 * a1 - s2 * b1, b2 = s2 * a1 + c2 * b1, cc2 = c2 * cc1 - s2 * d1, d2 = s2 * cc1 + c2 * d1;
        A[p1 * lda + p2] = a2; A[p1 * lda + q2] = b2; A[q1 * lda + p2] = cc2; A[q1 * lda + q2] = d2;
        A[p2 * lda + p1] = a2; A[q2 * lda + p1] = b2; A[p2 * lda + q1] = cc2; A[q2 * lda + q1] = d2;
      }
      if (V) {
        for (int t = warp; t < np; t += nwarps) {
          const double c = js->c[t], s = js->s[t];
          if (s == 0.0) continue;
          const int p = js->p[t], q = js->q[t];
          for (int row = lane; row < m; row += 32) {
            double vx = V[row * ldv + p], vy = V[row * ldv + q];
            V[row * ldv + p] = c * vx - s * vy;
            V[row * ldv + q] = s * vx + c * vy;
          }
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
