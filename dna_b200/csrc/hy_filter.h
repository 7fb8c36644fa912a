// hyena-b200: device-side argument block shared by the implicit-filter kernels (hy_filter.cu: CUDA-core kernels,
// hy_filter_tc.cu: legacy mma.sync last-Linear backward, hy_filter_tc05.cu: tcgen05 / TMEM kernels).
#pragma once
#include "hy_host.h"

namespace hy {

constexpr int kFT = 64;    // positions per tile of the CUDA-core kernels
constexpr int kFO = 64;    // padded MLP width

struct FilterDev {
  int L, D, order, emb_dim, n_inner;
  const float* z; int ldz;
  const float* t;
  const float* w_in; const float* b_in;
  const float* w_h; const float* b_h;
  const float* w_out;
  const float* freq;
  const float* deltas;
  float shift; int modulate;
  float* hsave; int ldh;     // optional: last hidden activation [L][ldh] for the backward (fast forward kernel only)
  float* asave; int lda;     // optional: the trunk's pre-activations a_l[j][t] as [layer][kFO][lda] (forward: written;
                             // saved-trunk backward: read) — lda a multiple of kFT
};

// tcgen05 forward (hy_filter_tc05.cu): k [D][ldk] (+ optional h_last / a_save through FilterDev). Returns HY_OK, or
// HY_ERR_UNSUPPORTED (without setting the error text) when the shape is outside its range and the caller should take
// the CUDA-core kernel.
int filter_fwd_tc05(const FilterDev& a, float* k, int ldk, void* stream);
bool filter_tc05_enabled();

}  // namespace hy
