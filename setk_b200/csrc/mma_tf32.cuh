// mma_tf32.cuh -- warp-level TF32 tensor-core MMA (mma.sync.m16n8k8, SASS HMMA.1688.F32.TF32)
// with fp32 accumulation, and the 3xTF32 operand split that keeps fp32 accuracy:
//   a = a_hi + a_lo,  a_hi = tf32(a),  a_lo = tf32(a - a_hi)
//   a b ~= a_hi b_hi + a_hi b_lo + a_lo b_hi          (the lo x lo term is below fp32's ulp)
// Fragment ownership (PTX ISA, m16n8k8 .tf32), g = lane >> 2, q = lane & 3:
//   A (16 x 8, row):  a0 (g, q)   a1 (g + 8, q)   a2 (g, q + 4)   a3 (g + 8, q + 4)
//   B (8 x 8, col):   b0 (k = q, n = g)           b1 (k = q + 4, n = g)
//   C / D (16 x 8):   c0 (g, 2q)  c1 (g, 2q + 1)  c2 (g + 8, 2q)  c3 (g + 8, 2q + 1)
// Under SETK_EMU the same contract is modelled through the warp's exchange slots.
#pragma once
#include "compat.cuh"

namespace setk {

#ifdef SETK_EMU
__device__ inline unsigned tf32_rna(float x) {            // round to nearest, ties away, 10-bit mantissa
  unsigned u = __float_as_uint(x);
  if ((u & 0x7f800000u) == 0x7f800000u) return u;         // inf / nan
  u += 0x1000u;
  return u & 0xffffe000u;
}
__device__ inline void mma_tf32_16x8x8(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  // every lane publishes its fragments; then each lane gathers the row / column it needs
  float mine[6], all[6 * 32];
  // the tensor core reads the upper 19 bits of a .tf32 operand
  for (int i = 0; i < 4; ++i) mine[i] = __uint_as_float(a[i] & 0xffffe000u);
  for (int i = 0; i < 2; ++i) mine[4 + i] = __uint_as_float(b[i] & 0xffffe000u);
  emu::warp_allgather(mine, 6, all);
  const int lane = emu::g_tid & 31, g = lane >> 2, q = lane & 3;
  auto A = [&](int r, int k) { return all[((r >= 8 ? 1 : 0) + (k >= 4 ? 2 : 0)) * 32 + (r & 7) * 4 + (k & 3)]; };
  auto B = [&](int k, int n) { return all[(4 + (k >= 4 ? 1 : 0)) * 32 + n * 4 + (k & 3)]; };
  const int rows[4] = {g, g, g + 8, g + 8}, cols[4] = {2 * q, 2 * q + 1, 2 * q, 2 * q + 1};
  for (int i = 0; i < 4; ++i) {
    float s = d[i];
    for (int k = 0; k < 8; ++k) s += A(rows[i], k) * B(k, cols[i]);
    d[i] = s;
  }
}
#else
__device__ __forceinline__ unsigned tf32_rna(float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
#endif

// x = hi + lo in TF32.  hi = x with its low 13 mantissa bits cleared (one LOP3), lo = x - hi (exact,
// one FADD) handed over as it is: the tensor core reads the upper 19 bits of a .tf32 operand, i.e.
// truncates lo itself (|error| <= 2^-10 |lo| <= 2^-20 |x|, the size of the lo x lo term the scheme
// drops anyway).  cvt.rna.tf32.f32 is not one instruction on sm_100a: ptxas expands it into an
// integer add, a mask and a NaN test -- with two of them per operand, 217 M of the covariance
// kernel's 308 M instructions were conversions (profiles/r2_cov_mma8_rna_ncu.txt).
__device__ __forceinline__ void tf32_split(float x, unsigned& hi, unsigned& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

// d += A B with both operands split: three tensor-core passes
__device__ __forceinline__ void mma_3xtf32(float (&d)[4], const unsigned (&ah)[4], const unsigned (&al)[4],
                                           const unsigned (&bh)[2], const unsigned (&bl)[2]) {
  mma_tf32_16x8x8(d, al, bh);
  mma_tf32_16x8x8(d, ah, bl);
  mma_tf32_16x8x8(d, ah, bh);
}

}  // namespace setk
