// jacobi_coop.cuh -- Hermitian eigen-decomposition of a small complex matrix by
// a group of threads of one warp (parallel-order cyclic Jacobi).
//
// The thread-per-matrix Jacobi of hermitian_solve.cuh keeps A and V in local
// memory and runs one long dependent chain per thread; with only B*F matrices
// that leaves a B200 mostly idle for C > 4.  Here a matrix lives in shared
// memory and GS = 2/4/8/16 threads (>= C, one per row/column) work on it:
//   round-robin tournament, CP - 1 rounds per sweep, CP/2 disjoint (p, q) per round
//   phase 1  thread m computes the rotation of pair m              (c, s, e^{-i phi})
//   phase 2  thread r updates its ROW r of  A <- A J  and  V <- V J (all pairs)
//   phase 3  thread k updates its COLUMN k of  A <- J^H A           (all pairs)
// separated by __syncwarp().  Same rotation formulas and stopping rule as
// jacobi_eigh(); the rotation order differs, so results agree to rounding.
// Replaces numpy.linalg.eigh at cluster.py:104 (and beamformer.py:45 for C > 4).
#pragma once
#include "hermitian_solve.cuh"

namespace setk {

template <int C>
struct Coop {
  static constexpr int GS = C <= 2 ? 2 : (C <= 4 ? 4 : (C <= 8 ? 8 : 16));   // threads per matrix
  static constexpr int LD = C + 1;            // row pitch in cd: rows start in different banks
  static constexpr int CP = (C + 1) & ~1;     // tournament positions (a dummy when C is odd)
  static constexpr int MAT = C * LD;          // cd per matrix
  static constexpr int ROT = 4 * (GS / 2 > 0 ? GS / 2 : 1);   // doubles of rotation parameters
};

// pair m of a round (circle method): position CP-1 stays, the others rotate
__device__ __forceinline__ void coop_pair(int CP, int round, int m, int& p, int& q) {
  const int n1 = CP - 1;
  int a, b;
  if (m == 0) { a = n1; b = round; }
  else { a = (round + m) % n1; b = (round - m + n1) % n1; }
  p = a < b ? a : b;
  q = a < b ? b : a;
}

// A, V: this group's matrices in shared memory, [C][LD] cd; A holds the
// Hermitian input (both triangles) and ends diagonal, V its eigenvectors in
// columns.  rot: Coop<C>::ROT doubles of scratch.  r: thread index in the group.
// `active` false: the group only keeps the warp's barriers company.
// EVERY lane of the warp must call.  Returns false when the sweep limit was hit.
template <int C>
__device__ inline bool jacobi_coop(cd* A, cd* V, double* rot, int r, bool active) {
  using K = Coop<C>;
  constexpr int LD = K::LD, CP = K::CP, GS = K::GS;
  const unsigned full = 0xffffffffu;
  const bool row = active && r < C;
  if (row) {
    for (int j = 0; j < C; ++j) V[r * LD + j] = cd_make(r == j ? 1.0 : 0.0, 0.0);
  }
  __syncwarp();
  const int kMaxSweeps = 40;
  bool conv = !active || C == 1;
  bool limit_ok = true;
  for (int sweep = 0;; ++sweep) {
    double off = 0.0, dg = 0.0;
    if (row) {
      for (int j = 0; j < C; ++j) if (j != r) off += cd_abs2(A[r * LD + j]);
      dg = A[r * LD + r].x * A[r * LD + r].x;
    }
    for (int o = GS / 2; o > 0; o >>= 1) {
      off += __shfl_xor_sync(full, off, o);
      dg += __shfl_xor_sync(full, dg, o);
    }
    if (active && (off <= 2e-30 * dg || off == 0.0)) conv = true;     // off counts both triangles
    if (sweep == kMaxSweeps && !conv) { limit_ok = off <= 2e-20 * dg; conv = true; }
    int all = conv ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) all &= __shfl_xor_sync(full, all, o);
    if (all) break;
    for (int round = 0; round < CP - 1; ++round) {
      if (active && !conv && r < CP / 2) {
        int p, q;
        coop_pair(CP, round, r, p, q);
        double c = 1.0, s = 0.0;
        cd em = cd_make(1.0, 0.0);
        if (q < C) {
          const cd apq = A[p * LD + q];
          const double mag2 = cd_abs2(apq);
          if (mag2 != 0.0) {
            const double mag = sqrt(mag2);
            const double app = A[p * LD + p].x, aqq = A[q * LD + q].x;
            if (mag <= 1e-19 * (fabs(app) + fabs(aqq))) {
              A[p * LD + q] = cd_make(0.0, 0.0);
              A[q * LD + p] = cd_make(0.0, 0.0);
            } else {
              em = cd_make(apq.x / mag, -apq.y / mag);
              const double theta = (aqq - app) / (2.0 * mag);
              const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
              c = 1.0 / sqrt(t * t + 1.0);
              s = t * c;
            }
          }
        }
        rot[4 * r + 0] = c; rot[4 * r + 1] = s; rot[4 * r + 2] = em.x; rot[4 * r + 3] = em.y;
      }
      __syncwarp();
      if (row && !conv) {                      // A <- A J, V <- V J : my row
        for (int m = 0; m < CP / 2; ++m) {
          const double s = rot[4 * m + 1];
          if (s == 0.0) continue;
          int p, q;
          coop_pair(CP, round, m, p, q);
          const double c = rot[4 * m];
          const cd em = cd_make(rot[4 * m + 2], rot[4 * m + 3]);
          const cd Jqp = cd_scale(em, -s), Jqq = cd_scale(em, c);
          const cd akp = A[r * LD + p], akq = A[r * LD + q];
          A[r * LD + p] = cd_add(cd_scale(akp, c), cd_mul(akq, Jqp));
          A[r * LD + q] = cd_add(cd_scale(akp, s), cd_mul(akq, Jqq));
          const cd vkp = V[r * LD + p], vkq = V[r * LD + q];
          V[r * LD + p] = cd_add(cd_scale(vkp, c), cd_mul(vkq, Jqp));
          V[r * LD + q] = cd_add(cd_scale(vkp, s), cd_mul(vkq, Jqq));
        }
      }
      __syncwarp();
      if (row && !conv) {                      // A <- J^H A : my column
        for (int m = 0; m < CP / 2; ++m) {
          const double s = rot[4 * m + 1];
          if (s == 0.0) continue;
          int p, q;
          coop_pair(CP, round, m, p, q);
          const double c = rot[4 * m];
          const cd em = cd_make(rot[4 * m + 2], rot[4 * m + 3]);
          const cd Jqp = cd_scale(em, -s), Jqq = cd_scale(em, c);
          const cd apk = A[p * LD + r], aqk = A[q * LD + r];
          cd np_ = cd_add(cd_scale(apk, c), cd_mul(cd_conj(Jqp), aqk));
          cd nq_ = cd_add(cd_scale(apk, s), cd_mul(cd_conj(Jqq), aqk));
          if (r == q) { np_ = cd_make(0.0, 0.0); nq_.y = 0.0; }
          if (r == p) { nq_ = cd_make(0.0, 0.0); np_.y = 0.0; }
          A[p * LD + r] = np_;
          A[q * LD + r] = nq_;
        }
      }
      __syncwarp();
    }
  }
  return limit_ok;
}

}  // namespace setk
