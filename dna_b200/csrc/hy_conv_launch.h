// hyena-b200: launch wrappers for the long-convolution kernels, split over several translation
// units (one per activation dtype) so that nvcc compiles them in parallel.
#pragma once
#include "hy_host.h"
#include "hy_conv.cuh"

namespace hy {

constexpr int kNT = 256;

// columns per CTA tile of the four-step column transforms
#ifndef HY_COL_SMALL
#define HY_COL_SMALL 0
#endif
// Column lengths with an odd factor (M1 = 5 * 2^a, 3 * 2^a): pass 0 has 5 (3) x 2^j butterflies per column, so the tile
// must be wide enough that 256 threads are busy for several full rounds (M1 = 40 with 64 columns leaves 320 butterflies,
// 1.25 rounds; with 128 columns 640): M1 * T2 = 2560 .. 6144 points per tile.  Measured, long-conv family per 128 rows
// (profiles/r02m_sweep_lengths_{narrow,wide}_tiles.txt): M1 = 40: 64 -> 128 columns 1.25 -> 1.09 ms; M1 = 80: 32 -> 64 columns 2.37 -> 2.12 ms;
// M1 = 96: 3.21 -> 2.73 ms; M1 = 20 / 24: 128 -> 256 columns 0.66 -> 0.60 / 0.79 -> 0.70 ms.
HY_HD constexpr int col_T2_odd(int M1) {
  return M1 <= 24 ? 256 : (M1 <= 48 ? 128 : (M1 <= 96 ? 64 : (M1 <= 192 ? 32 : 16)));
}
HY_HD constexpr int col_T2(int M1) {
  return (M1 % 3 == 0 || M1 % 5 == 0) ? col_T2_odd(M1)
       : HY_COL_SMALL ? (M1 <= 16 ? 128 : (M1 <= 32 ? 32 : 16))
                      : (M1 <= 16 ? 256 : (M1 == 32 ? 64 : (M1 == 512 ? 16 : 32)));
}
HY_HD constexpr int col_base_nt() { return HY_COL_SMALL ? 128 : 256; }

// threads per CTA of the column kernels: 512 when the backward (two sequences) has that many butterflies
template <int M1, int NSEQ>
HY_HD constexpr int col_nt() {
  return (NSEQ == 2 && col_T2(M1) * (M1 / Plan<M1>::radix(0)) >= 2 * col_base_nt()) ? 2 * col_base_nt() : col_base_nt();
}
constexpr int kNTRowBwd = 512;

inline bool valid_block(int S) { return S == 256 || S == 512 || S == 1024 || S == 2048 || S == 4096; }
// column lengths with an instance: 2^a (a = 1..9) in hy_conv_{f32,bf16}.cu, 5 * 2^a and 3 * 2^a in hy_conv_odd{5,3}_{f32,bf16}.cu
inline bool pow2_cols(int M1) { return M1 >= 2 && M1 <= 512 && (M1 & (M1 - 1)) == 0; }
inline bool odd_cols(int M1) {
#define HY_CASE(MM) if (M1 == MM) return true;
  HY_COLS_ODD(HY_CASE)
#undef HY_CASE
  return false;
}
inline bool valid_cols(int M1) { return pow2_cols(M1) || odd_cols(M1); }

// dtype-dependent kernels (defined in hy_conv_f32.cu / hy_conv_bf16.cu via hy_conv_launch_impl.cuh)
template <class DT> int launch_fused_fwd(const ConvArgs& a, int S, int mode, void* stream);
template <class DT> int launch_fused_bwd(const ConvArgs& a, int S, void* stream);
template <class DT> int launch_fused_bwdg(const ConvArgs& a, int S, void* stream);   // with ConvArgs::gsave (dy only)
template <class DT> int launch_col_fwd(const ConvArgs& a, int M1, int S, int nseq, int dyo, void* stream);
template <class DT> int launch_col_inv(const ConvArgs& a, int M1, int S, int nseq, int epi, void* stream);
// the same for the column lengths with an odd factor (hy_conv_odd{5,3}_{f32,bf16}.cu)
// FAMILY = 5: M1 = 5 * 2^a (hy_conv_odd5_*.cu), FAMILY = 3: M1 = 3 * 2^a (hy_conv_odd3_*.cu) — one translation unit per
// family and dtype so that the build stays parallel
template <class DT, int FAMILY> int launch_col_fwd_odd(const ConvArgs& a, int M1, int S, int nseq, int dyo, void* stream);
template <class DT, int FAMILY> int launch_col_inv_odd(const ConvArgs& a, int M1, int S, int nseq, int epi, void* stream);
// persistent A/B/C pipeline over a ring of row buffers (hy_conv_pipe.cuh; hy_conv_pipe_f32.cu / _bf16.cu): returns
// HY_ERR_UNSUPPORTED without touching the error text when the geometry has no instance
template <class DT> int launch_conv_pipe(const ConvArgs& a, int M1, int S, int kind, float2* ring, unsigned* ctl, void* stream);
// dtype-independent kernels (hy_conv_rows.cu)
int launch_fused_dk(const ConvArgs& a, int S, void* stream);
int launch_row_conv(const ConvArgs& a, int M1, int S, int mode, void* stream);

}  // namespace hy
