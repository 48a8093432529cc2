// weights_args.cuh -- argument block shared by the per-bin weight kernels
// (weights.cu: one thread per bin; weights_coop.cu: one thread group per bin).
#pragma once
#include "common.cuh"

namespace setk {

struct WeightsArgs {
  int kind, rank1, ban, ref_channel;
  double beta;
  const void* Rs; const void* Rn; const void* Ry;
  int r_dtype;  // SETK_C64 / SETK_C128
  int B, F;
  void* w; int w_dtype;
  unsigned* status;
  int* ref_used;
  // PMWF automatic reference selection
  double* Wfull;   // [B][F][C][C] complex128 (interleaved)
  double* pows;    // [B][F][C][2]  (Re w^H Rs w, Re w^H Rn w)
};

}  // namespace setk
