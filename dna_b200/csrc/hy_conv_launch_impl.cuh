// hyena-b200: dtype-templated launch wrappers (included by hy_conv_f32.cu / hy_conv_bf16.cu).
#pragma once
#include "hy_conv_launch.h"

namespace hy {

#ifndef HY_CONV_ODD_TU
template <class DT, int S, int MODE>
static int fused_fwd_s(const ConvArgs& a, void* stream) {
  constexpr int NB = 4096 / S;
  auto kern = k_fused_fwd<DT, S, NB, kNT, MODE>;
  const size_t smem = sizeof(float4) * Plan<S>::tw_slots() + sizeof(float2) * NB * RowSmem<S>::kRow;
  const int grid = (a.nrows + NB - 1) / NB;
  HY_LAUNCH(kern, grid, kNT, smem, stream, a);
  return check_launch("k_fused_fwd");
}

template <class DT>
int launch_fused_fwd(const ConvArgs& a, int S, int mode, void* stream) {
#define HY_CASE(SS)                                                               \
  case SS:                                                                        \
    return mode == HY_PW_SPEC ? fused_fwd_s<DT, SS, HY_PW_SPEC>(a, stream)        \
                              : fused_fwd_s<DT, SS, HY_PW_CONV>(a, stream);
  switch (S) {
    HY_CASE(256)
    HY_CASE(512)
    HY_CASE(1024)
    HY_CASE(2048)
    HY_CASE(4096)
  }
#undef HY_CASE
  return fail(HY_ERR_UNSUPPORTED, "fused forward: unsupported transform length %d", S);
}

template <class DT, int S>
static int fused_bwd_s(const ConvArgs& a, void* stream) {
  constexpr int NB = 4096 / S;
  constexpr int TOTAL = (S / Plan<S>::radix(0)) * NB;
  // 79 KB of shared memory admit two CTAs per SM: with HY_FUSED_BWD_NT = 512 they carry 32 warps instead of 16 (the kernel
  // is bound by exposed load latency: long_scoreboard 6.5-7.3 stalls per issue at 16 warps, profiles/r02g_ncu_fused_regime_*)
  constexpr int NTB = HY_FUSED_BWD_NT;
  auto kern = k_fused_bwd<DT, S, NB, NTB>;
  const size_t smem = sizeof(float4) * Plan<S>::tw_slots() + sizeof(float2) * 2 * NB * RowSmem<S>::kRow + sizeof(float) * TOTAL;
  const int grid = (a.nrows + NB - 1) / NB;
  HY_LAUNCH(kern, grid, NTB, smem, stream, a);
  return check_launch("k_fused_bwd");
}

template <class DT>
int launch_fused_bwd(const ConvArgs& a, int S, void* stream) {
  switch (S) {
    case 256: return fused_bwd_s<DT, 256>(a, stream);
    case 512: return fused_bwd_s<DT, 512>(a, stream);
    case 1024: return fused_bwd_s<DT, 1024>(a, stream);
    case 2048: return fused_bwd_s<DT, 2048>(a, stream);
    case 4096: return fused_bwd_s<DT, 4096>(a, stream);
  }
  return fail(HY_ERR_UNSUPPORTED, "fused backward: unsupported transform length %d", S);
}

template <class DT, int S>
static int fused_bwdg_s(const ConvArgs& a, void* stream) {
  constexpr int NB = 4096 / S;
  auto kern = k_fused_bwdg<DT, S, NB, kNT>;
  const size_t smem = sizeof(float4) * Plan<S>::tw_slots() + sizeof(float2) * NB * RowSmem<S>::kRow;
  const int grid = (a.nrows + NB - 1) / NB;
  HY_LAUNCH(kern, grid, kNT, smem, stream, a);
  return check_launch("k_fused_bwdg");
}

template <class DT>
int launch_fused_bwdg(const ConvArgs& a, int S, void* stream) {
  switch (S) {
    case 256: return fused_bwdg_s<DT, 256>(a, stream);
    case 512: return fused_bwdg_s<DT, 512>(a, stream);
    case 1024: return fused_bwdg_s<DT, 1024>(a, stream);
    case 2048: return fused_bwdg_s<DT, 2048>(a, stream);
    case 4096: return fused_bwdg_s<DT, 4096>(a, stream);
  }
  return fail(HY_ERR_UNSUPPORTED, "fused backward (saved spectrum): unsupported transform length %d", S);
}

#endif  // !HY_CONV_ODD_TU

template <int M1, int T2, int NSEQ>
constexpr size_t col_smem_bytes() {
  using P = Plan<M1>;
  return sizeof(float4) * P::tw_slots() + sizeof(float2) * (M1 + (P::NS > 1 ? NSEQ * M1 * T2 : 0)) + sizeof(float) * 32;
}
// bytes of one staged bf16 source-row tile ([M1/2][2*T2 + 8] elements)
template <int M1, int T2>
constexpr size_t stage_tile_bytes() {
  return sizeof(unsigned short) * (size_t)(M1 / 2 > 0 ? M1 / 2 : 1) * (2 * T2 + 8);
}

template <class DT, int M1, int NSEQ, bool DYO = false>
static int col_fwd_m(const ConvArgs& a0, void* stream) {
  constexpr int T2 = col_T2(M1);
  ConvArgs a = a0;
  a.twV = twV_table(M1, a.S, T2);
  if (!a.twV) return HY_ERR_CUDA;
  if (a.S % T2 != 0) return fail(HY_ERR_UNSUPPORTED, "row length %d not a multiple of the column tile %d", a.S, T2);
  constexpr int NT = col_nt<M1, NSEQ>();
  auto kern = k_col_fwd<DT, M1, T2, NT, NSEQ, DYO>;
  const size_t smem = col_smem_bytes<M1, T2, NSEQ>() + ((DT::kBf16 && NSEQ == 1 && a.stage_ok) ? 2 * stage_tile_bytes<M1, T2>() : 0);
  HY_LAUNCH(kern, dim3(a.S / T2, a.nrows), NT, smem, stream, a);
  return check_launch("k_col_fwd");
}

#define HY_COL_FWD_CASE(MM) \
  case MM:          \
    return nseq == 2 ? col_fwd_m<DT, MM, 2>(a, stream)                \
                     : (dyo ? col_fwd_m<DT, MM, 1, true>(a, stream) : col_fwd_m<DT, MM, 1>(a, stream));
#ifndef HY_CONV_ODD_TU
template <class DT>
int launch_col_fwd(const ConvArgs& a, int M1, int S, int nseq, int dyo, void* stream) {
  if (dyo && nseq != 1) return fail(HY_ERR_ARG, "col_fwd: the dy-only phase carries one sequence");
  switch (M1) {
    HY_COLS_POW2(HY_COL_FWD_CASE)
  }
  return (M1 % 5 == 0) ? launch_col_fwd_odd<DT, 5>(a, M1, S, nseq, dyo, stream) : launch_col_fwd_odd<DT, 3>(a, M1, S, nseq, dyo, stream);
}
#else
template <class DT, int FAMILY>
int launch_col_fwd_odd(const ConvArgs& a, int M1, int S, int nseq, int dyo, void* stream) {
  (void)S;
  switch (M1) {
#if HY_CONV_ODD_TU == 5
    HY_COLS_ODD5(HY_COL_FWD_CASE)
#else
    HY_COLS_ODD3(HY_COL_FWD_CASE)
#endif
  }
  return fail(HY_ERR_UNSUPPORTED, "four-step: unsupported column length %d", M1);
}
#endif
#undef HY_COL_FWD_CASE

template <class DT, int M1, int NSEQ, int EPI>
static int col_inv_m(const ConvArgs& a0, void* stream) {
  constexpr int T2 = col_T2(M1);
  ConvArgs a = a0;
  a.twV = twV_table(M1, a.S, T2);
  if (!a.twV) return HY_ERR_CUDA;
  constexpr int NT = col_base_nt();
  auto kern = k_col_inv<DT, M1, T2, NT, NSEQ, EPI>;
  const size_t smem = col_smem_bytes<M1, T2, 1>() + ((DT::kBf16 && a.stage_ok) ? (EPI == 0 ? 1 : 2) * stage_tile_bytes<M1, T2>() : 0);
  HY_LAUNCH(kern, dim3(a.S / T2, a.nrows), NT, smem, stream, a);
  return check_launch("k_col_inv");
}

#define HY_COL_INV_CASE(MM) \
  case MM:          \
    return epi == 1 ? (nseq == 2 ? col_inv_m<DT, MM, 2, 1>(a, stream) : col_inv_m<DT, MM, 1, 1>(a, stream)) \
                    : col_inv_m<DT, MM, 1, 0>(a, stream);
#ifndef HY_CONV_ODD_TU
template <class DT>
int launch_col_inv(const ConvArgs& a, int M1, int S, int nseq, int epi, void* stream) {
  if (epi == 0 && nseq != 1) return fail(HY_ERR_ARG, "col_inv: epilogue/sequence mismatch");
  switch (M1) {
    HY_COLS_POW2(HY_COL_INV_CASE)
  }
  return (M1 % 5 == 0) ? launch_col_inv_odd<DT, 5>(a, M1, S, nseq, epi, stream) : launch_col_inv_odd<DT, 3>(a, M1, S, nseq, epi, stream);
}
#else
template <class DT, int FAMILY>
int launch_col_inv_odd(const ConvArgs& a, int M1, int S, int nseq, int epi, void* stream) {
  (void)S;
  switch (M1) {
#if HY_CONV_ODD_TU == 5
    HY_COLS_ODD5(HY_COL_INV_CASE)
#else
    HY_COLS_ODD3(HY_COL_INV_CASE)
#endif
  }
  return fail(HY_ERR_UNSUPPORTED, "four-step: unsupported column length %d", M1);
}
#endif
#undef HY_COL_INV_CASE

}  // namespace hy
