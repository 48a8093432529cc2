// hermitian_solve.cuh -- fp64 per-bin dense kernels for C x C complex matrices,
// C <= 16, one problem per thread.
//
// Replaces, per frequency bin, the LAPACK calls the reference makes through
// numpy/scipy (and through include/cblas-cpl-wrappers.h:150-227 in the C++
// twin):
//   jacobi_eigh     np.linalg.eigh           (beamformer.py:45)  / cheev (Hed)
//   cholesky_lower  potrf inside scipy.linalg.eigh(a, b)   (beamformer.py:53)
//   gev_principal   hegvd                    (beamformer.py:53)  / chegv (Hged)
//   lu_solve        np.linalg.solve (gesv)   (beamformer.py:536,568,646)
//
// Plain C++ on purpose (no CUDA intrinsics): the same source is compiled for
// the device by nvcc and for the CPU test tier by g++ (tests/emu).
#pragma once
#include "common.cuh"
#include "compat.cuh"
#include <math.h>

// Loops over the matrix dimension are fully unrolled for C <= 4 (everything in
// registers) and left rolled above that (arrays in local memory, seconds to
// compile instead of minutes).
#ifdef SETK_EMU
#define SETK_UNROLL_C
#define SETK_NOUNROLL
#else
#define SETK_UNROLL_C _Pragma("unroll (C <= 4 ? 16 : 1)")
#define SETK_NOUNROLL _Pragma("unroll 1")
#endif

namespace setk {

struct cd { double x, y; };

__device__ __forceinline__ cd cd_make(double x, double y) { cd r; r.x = x; r.y = y; return r; }
__device__ __forceinline__ cd cd_add(cd a, cd b) { return cd_make(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cd cd_sub(cd a, cd b) { return cd_make(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cd cd_mul(cd a, cd b) { return cd_make(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// a * conj(b)
__device__ __forceinline__ cd cd_mulc(cd a, cd b) { return cd_make(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
__device__ __forceinline__ cd cd_conj(cd a) { return cd_make(a.x, -a.y); }
__device__ __forceinline__ cd cd_scale(cd a, double s) { return cd_make(a.x * s, a.y * s); }
__device__ __forceinline__ double cd_abs2(cd a) { return a.x * a.x + a.y * a.y; }
__device__ __forceinline__ cd cd_div(cd a, cd b) {
  // Smith's algorithm (what numpy's complex division does)
  if (fabs(b.x) >= fabs(b.y)) {
    double r = b.y / b.x, d = b.x + b.y * r;
    return cd_make((a.x + a.y * r) / d, (a.y - a.x * r) / d);
  } else {
    double r = b.x / b.y, d = b.x * r + b.y;
    return cd_make((a.x * r + a.y) / d, (a.y * r - a.x) / d);
  }
}

template <int C>
struct CMat {
  cd a[C][C];
};
template <int C>
struct CVec {
  cd v[C];
};

// ---------------------------------------------------------------------------
// Cyclic Jacobi for a complex Hermitian matrix.  On return A is (numerically)
// diagonal, V holds the eigenvectors in its columns (A_in = V diag V^H).
// Returns the number of sweeps used, or -1 if the sweep limit was reached.
// ---------------------------------------------------------------------------
template <int C>
__device__ inline int jacobi_eigh(CMat<C>& A, CMat<C>& V) {
  SETK_UNROLL_C
  for (int i = 0; i < C; ++i)
    SETK_UNROLL_C
    for (int j = 0; j < C; ++j) V.a[i][j] = cd_make(i == j ? 1.0 : 0.0, 0.0);
  if (C == 1) return 0;
  const int kMaxSweeps = 40;
  SETK_NOUNROLL
  for (int sweep = 0; sweep < kMaxSweeps; ++sweep) {
    double off = 0.0, diag = 0.0;
    SETK_UNROLL_C
    for (int i = 0; i < C; ++i) {
      diag += A.a[i][i].x * A.a[i][i].x;
      SETK_UNROLL_C
      for (int j = i + 1; j < C; ++j) off += cd_abs2(A.a[i][j]);
    }
    if (off <= 1e-30 * diag || off == 0.0) return sweep;
    SETK_UNROLL_C
    for (int p = 0; p < C - 1; ++p) {
      SETK_UNROLL_C
      for (int q = p + 1; q < C; ++q) {
        cd apq = A.a[p][q];
        double mag2 = cd_abs2(apq);
        if (mag2 == 0.0) continue;
        double mag = sqrt(mag2);
        double app = A.a[p][p].x, aqq = A.a[q][q].x;
        // skip rotations that cannot change the diagonal in double precision
        if (mag <= 1e-19 * (fabs(app) + fabs(aqq))) {
          A.a[p][q] = cd_make(0.0, 0.0);
          A.a[q][p] = cd_make(0.0, 0.0);
          continue;
        }
        // e^{-i phi} with a_pq = |a_pq| e^{i phi}
        cd em = cd_make(apq.x / mag, -apq.y / mag);
        double theta = (aqq - app) / (2.0 * mag);
        double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0);
        double s = t * c;
        // J = P R : J_pp = c, J_pq = s, J_qp = -s e^{-i phi}, J_qq = c e^{-i phi}
        cd Jqp = cd_scale(em, -s), Jqq = cd_scale(em, c);
        // A <- A J  (columns p, q)
        SETK_UNROLL_C
        for (int k = 0; k < C; ++k) {
          cd akp = A.a[k][p], akq = A.a[k][q];
          A.a[k][p] = cd_add(cd_scale(akp, c), cd_mul(akq, Jqp));
          A.a[k][q] = cd_add(cd_scale(akp, s), cd_mul(akq, Jqq));
        }
        // A <- J^H A  (rows p, q)
        SETK_UNROLL_C
        for (int k = 0; k < C; ++k) {
          cd apk = A.a[p][k], aqk = A.a[q][k];
          A.a[p][k] = cd_add(cd_scale(apk, c), cd_mul(cd_conj(Jqp), aqk));
          A.a[q][k] = cd_add(cd_scale(apk, s), cd_mul(cd_conj(Jqq), aqk));
        }
        A.a[p][q] = cd_make(0.0, 0.0);
        A.a[q][p] = cd_make(0.0, 0.0);
        A.a[p][p].y = 0.0;
        A.a[q][q].y = 0.0;
        // V <- V J
        SETK_UNROLL_C
        for (int k = 0; k < C; ++k) {
          cd vkp = V.a[k][p], vkq = V.a[k][q];
          V.a[k][p] = cd_add(cd_scale(vkp, c), cd_mul(vkq, Jqp));
          V.a[k][q] = cd_add(cd_scale(vkp, s), cd_mul(vkq, Jqq));
        }
      }
    }
  }
  // sweep limit: accept if the off-diagonal mass is below 1e-10 relative
  double off = 0.0, diag = 0.0;
  SETK_UNROLL_C
  for (int i = 0; i < C; ++i) {
    diag += A.a[i][i].x * A.a[i][i].x;
    SETK_UNROLL_C
    for (int j = i + 1; j < C; ++j) off += cd_abs2(A.a[i][j]);
  }
  return (off <= 1e-20 * diag) ? kMaxSweeps : -1;
}

// Normalise v to unit 2-norm with component 0 real and >= 0 (the library's
// eigenvector convention; LAPACK's differs by a per-bin sign).
template <int C>
__device__ inline void canonical_phase(CVec<C>& v) {
  double n2 = 0.0;
  SETK_UNROLL_C
  for (int i = 0; i < C; ++i) n2 += cd_abs2(v.v[i]);
  double inv = n2 > 0.0 ? 1.0 / sqrt(n2) : 0.0;
  double m0 = sqrt(cd_abs2(v.v[0]));
  cd ph = m0 > 0.0 ? cd_make(v.v[0].x / m0, -v.v[0].y / m0) : cd_make(1.0, 0.0);
  SETK_UNROLL_C
  for (int i = 0; i < C; ++i) v.v[i] = cd_scale(cd_mul(v.v[i], ph), inv);
  if (m0 > 0.0) v.v[0].y = 0.0;
}

// Principal eigenvector of a Hermitian matrix (A is destroyed).
// Returns false when Jacobi did not converge.
template <int C>
__device__ inline bool principal_eigvec(CMat<C>& A, CVec<C>& out) {
  // enforce exact Hermitian symmetry from the lower triangle, as LAPACK
  // zheevd(UPLO='L') reads only that triangle (np.linalg.eigh default)
  SETK_UNROLL_C
  for (int i = 0; i < C; ++i) {
    A.a[i][i].y = 0.0;
    SETK_UNROLL_C
    for (int j = i + 1; j < C; ++j) A.a[i][j] = cd_conj(A.a[j][i]);
  }
  CMat<C> V;
  int sweeps = jacobi_eigh<C>(A, V);
  int best = 0;
  double lam = A.a[0][0].x;
  SETK_UNROLL_C
  for (int i = 1; i < C; ++i)
    if (A.a[i][i].x > lam) { lam = A.a[i][i].x; best = i; }
  SETK_UNROLL_C
  for (int i = 0; i < C; ++i) out.v[i] = V.a[i][best];
  canonical_phase<C>(out);
  return sweeps >= 0;
}

// Lower Cholesky factor of a Hermitian positive definite matrix, in place in the
// lower triangle (upper triangle zeroed).  Returns false if not PD.
template <int C>
__device__ inline bool cholesky_lower(CMat<C>& A) {
  bool ok = true;
  SETK_UNROLL_C
  for (int j = 0; j < C; ++j) {
    double d = A.a[j][j].x;
    SETK_UNROLL_C
    for (int k = 0; k < j; ++k) d -= cd_abs2(A.a[j][k]);
    if (!(d > 0.0)) { ok = false; d = 1.0; }
    double l = sqrt(d);
    A.a[j][j] = cd_make(l, 0.0);
    SETK_UNROLL_C
    for (int i = j + 1; i < C; ++i) {
      cd s = A.a[i][j];
      SETK_UNROLL_C
      for (int k = 0; k < j; ++k) s = cd_sub(s, cd_mulc(A.a[i][k], A.a[j][k]));
      A.a[i][j] = cd_scale(s, 1.0 / l);
    }
    SETK_UNROLL_C
    for (int i = 0; i < j; ++i) A.a[i][j] = cd_make(0.0, 0.0);
  }
  return ok;
}

// Solve L X = B in place (L lower triangular with real positive diagonal).
template <int C>
__device__ inline void forward_subst(const CMat<C>& L, CMat<C>& B) {
  SETK_UNROLL_C
  for (int col = 0; col < C; ++col)
    SETK_UNROLL_C
    for (int i = 0; i < C; ++i) {
      cd s = B.a[i][col];
      SETK_UNROLL_C
      for (int k = 0; k < i; ++k) s = cd_sub(s, cd_mul(L.a[i][k], B.a[k][col]));
      B.a[i][col] = cd_scale(s, 1.0 / L.a[i][i].x);
    }
}

// Principal generalised eigenvector of (Rs, Rn): Rn = L L^H, M = L^-1 Rs L^-H,
// y = principal(M) (canonical phase), w = L^-H y  =>  w^H Rn w = 1.
// Returns status bits (SETK_ST_NOT_PD / SETK_ST_NO_CONVERGE) -- Rs, Rn destroyed.
template <int C>
__device__ inline unsigned gev_principal(CMat<C>& Rs, CMat<C>& Rn, CVec<C>& w) {
  unsigned st = 0;
  // read only the lower triangle of Rn (scipy.linalg.eigh default lower=True).
  // When Rn is numerically not positive definite (e.g. a near-binary mask leaves
  // fewer noise frames than channels) the reference falls back to a non-Hermitian
  // scipy.linalg.eig (beamformer.py:55-58) whose answer on a singular pencil is
  // arbitrary; here the diagonal is loaded by tr(Rn)/C * 1e-10, 1e-8, ... and the
  // SETK_ST_REGULARIZED warning bit is set.  Only a matrix that stays non-PD
  // (zero / negative trace) reports SETK_ST_NOT_PD.
  {
    double tr = 0.0;
    SETK_UNROLL_C
    for (int i = 0; i < C; ++i) tr += Rn.a[i][i].x;
    CMat<C> L = Rn;
    bool ok = cholesky_lower<C>(L);
    double load = 1e-10;
    SETK_NOUNROLL
    for (int attempt = 0; attempt < 5 && !ok && tr > 0.0; ++attempt, load *= 100.0) {
      L = Rn;
      SETK_UNROLL_C
      for (int i = 0; i < C; ++i) L.a[i][i].x += load * tr / C;
      ok = cholesky_lower<C>(L);
      st |= SETK_ST_REGULARIZED;
    }
    if (!ok) st |= SETK_ST_NOT_PD;
    Rn = L;
  }
  // Rs from its lower triangle as well
  SETK_UNROLL_C
  for (int i = 0; i < C; ++i) {
    Rs.a[i][i].y = 0.0;
    SETK_UNROLL_C
    for (int j = i + 1; j < C; ++j) Rs.a[i][j] = cd_conj(Rs.a[j][i]);
  }
  forward_subst<C>(Rn, Rs);                 // Rs <- X = L^-1 Rs
  CMat<C> Xh;                               // X^H
  SETK_UNROLL_C
  for (int i = 0; i < C; ++i)
    SETK_UNROLL_C
    for (int j = 0; j < C; ++j) Xh.a[i][j] = cd_conj(Rs.a[j][i]);
  forward_subst<C>(Rn, Xh);                 // Y = L^-1 X^H ; M = Y^H
  CMat<C> M;
  SETK_UNROLL_C
  for (int i = 0; i < C; ++i)
    SETK_UNROLL_C
    for (int j = 0; j <= i; ++j) {
      // symmetrise: M_ij = (Y^H_ij + conj(Y^H_ji)) / 2 = (conj(Y_ji) + Y_ij) / 2
      cd m = cd_scale(cd_add(cd_conj(Xh.a[j][i]), Xh.a[i][j]), 0.5);
      M.a[i][j] = m;
      M.a[j][i] = cd_conj(m);
    }
  CVec<C> y;
  if (!principal_eigvec<C>(M, y)) st |= SETK_ST_NO_CONVERGE;
  // back substitution L^H w = y
  SETK_UNROLL_C
  for (int i = C - 1; i >= 0; --i) {
    cd s = y.v[i];
    SETK_UNROLL_C
    for (int k = i + 1; k < C; ++k) s = cd_sub(s, cd_mul(cd_conj(Rn.a[k][i]), w.v[k]));
    w.v[i] = cd_scale(s, 1.0 / Rn.a[i][i].x);
  }
  return st;
}

// LU factorisation with partial pivoting (pivot by |re|+|im| like LAPACK's
// izamax), in place; perm[i] = row swapped into position i.
// Returns false on an exactly zero pivot (numpy: LinAlgError "Singular matrix").
template <int C>
__device__ inline bool lu_factor(CMat<C>& A, int* perm) {
  bool ok = true;
  SETK_UNROLL_C
  for (int k = 0; k < C; ++k) {
    int piv = k;
    double best = fabs(A.a[k][k].x) + fabs(A.a[k][k].y);
    SETK_UNROLL_C
    for (int i = k + 1; i < C; ++i) {
      double m = fabs(A.a[i][k].x) + fabs(A.a[i][k].y);
      if (m > best) { best = m; piv = i; }
    }
    perm[k] = piv;
    if (piv != k)
      SETK_UNROLL_C
      for (int j = 0; j < C; ++j) { cd t = A.a[k][j]; A.a[k][j] = A.a[piv][j]; A.a[piv][j] = t; }
    if (best == 0.0) { ok = false; continue; }
    cd inv = cd_div(cd_make(1.0, 0.0), A.a[k][k]);
    SETK_UNROLL_C
    for (int i = k + 1; i < C; ++i) {
      cd l = cd_mul(A.a[i][k], inv);
      A.a[i][k] = l;
      SETK_UNROLL_C
      for (int j = k + 1; j < C; ++j) A.a[i][j] = cd_sub(A.a[i][j], cd_mul(l, A.a[k][j]));
    }
  }
  return ok;
}

template <int C>
__device__ inline void lu_solve_vec(const CMat<C>& LU, const int* perm, CVec<C>& b) {
  SETK_UNROLL_C
  for (int k = 0; k < C; ++k) {
    int p = perm[k];
    if (p != k) { cd t = b.v[k]; b.v[k] = b.v[p]; b.v[p] = t; }
  }
  SETK_UNROLL_C
  for (int i = 0; i < C; ++i) {
    cd s = b.v[i];
    SETK_UNROLL_C
    for (int k = 0; k < i; ++k) s = cd_sub(s, cd_mul(LU.a[i][k], b.v[k]));
    b.v[i] = s;
  }
  SETK_UNROLL_C
  for (int i = C - 1; i >= 0; --i) {
    cd s = b.v[i];
    SETK_UNROLL_C
    for (int k = i + 1; k < C; ++k) s = cd_sub(s, cd_mul(LU.a[i][k], b.v[k]));
    b.v[i] = cd_div(s, LU.a[i][i]);
  }
}

template <int C>
__device__ inline cd dotc(const CVec<C>& a, const CVec<C>& b) {  // a^H b
  cd s = cd_make(0.0, 0.0);
  SETK_UNROLL_C
  for (int i = 0; i < C; ++i) s = cd_add(s, cd_mul(cd_conj(a.v[i]), b.v[i]));
  return s;
}

template <int C>
__device__ inline void matvec(const CMat<C>& A, const CVec<C>& x, CVec<C>& y) {
  SETK_UNROLL_C
  for (int i = 0; i < C; ++i) {
    cd s = cd_make(0.0, 0.0);
    SETK_UNROLL_C
    for (int j = 0; j < C; ++j) s = cd_add(s, cd_mul(A.a[i][j], x.v[j]));
    y.v[i] = s;
  }
}

}  // namespace setk
