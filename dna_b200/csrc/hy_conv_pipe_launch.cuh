// hyena-b200: launch wrapper of the persistent four-step pipeline (included by hy_conv_pipe_f32.cu / _bf16.cu).
#pragma once
#include "hy_conv_launch.h"
#include "hy_conv_launch_impl.cuh"
#include "hy_conv_pipe.cuh"

namespace hy {

unsigned long long* pipe_stats_buffer();   // hy_conv_api.cu: nullptr unless hy_debug_pipe_stats enabled collection
int pipe_lag_override();                    // hy_conv_api.cu: 0 = pipe_lag(items per step), 1..4 = forced (experiments)

template <class DT, int M1, int S, int KIND>
static int conv_pipe_m(const ConvArgs& a0, float2* ring, unsigned* ctl, void* stream) {
  constexpr int T2 = col_T2(M1);
  constexpr int NT = kNT;
  static_assert(col_nt<M1, 1>() == NT && col_base_nt() == NT, "the pipeline runs every phase with 256 threads");
  ConvArgs a = a0;
  a.twV = twV_table(M1, a.S, T2);
  if (!a.twV) return HY_ERR_CUDA;
  if (a.S != S || a.S % T2 != 0) return HY_ERR_UNSUPPORTED;
  constexpr size_t smem_col = col_smem_bytes<M1, T2, 1>() + (DT::kBf16 ? 2 * stage_tile_bytes<M1, T2>() : 0);
  constexpr size_t smem_row = sizeof(float4) * Plan<S>::tw_slots() + sizeof(float2) * 2 * RowSmem<S>::kRow;
  constexpr size_t smem = smem_col > smem_row ? smem_col : smem_row;
  auto kern = k_conv_pipe<DT, M1, T2, S, NT, KIND>;
  constexpr int per = (pipe_has_a(KIND) ? S / T2 : 0) + M1 / 2 + (pipe_has_c(KIND) ? S / T2 : 0);
  int lag = pipe_lag_override();
  if (lag < 1 || lag > (kPipeRingMax - 2) / 2) lag = pipe_lag(per);
  const long long total = (long long)(a.nrows + 2 * lag) * per;
#ifdef HY_EMU_BUILD
  memset(ctl, 0, pipe_ctl_bytes(a.nrows));
  const int grid = (int)(total < 6 ? total : 6);
#else
  static thread_local int cached_dev = -1, cached_grid = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  hy_set_smem(kern, smem);
  if (dev != cached_dev) {
    int nb = 0, sms = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, NT, smem) != cudaSuccess || nb < 1) nb = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cached_grid = nb * (sms > 0 ? sms : 1);
    cached_dev = dev;
  }
  const int grid = (int)(total < cached_grid ? total : cached_grid);
  if (cudaMemsetAsync(ctl, 0, pipe_ctl_bytes(a.nrows), (cudaStream_t)stream) != cudaSuccess)
    return fail(HY_ERR_CUDA, "conv pipeline: clearing the control block failed");
#endif
  HY_LAUNCH(kern, grid, NT, smem, stream, a, ring, ctl, pipe_stats_buffer(), lag);
  return check_launch("k_conv_pipe");
}

template <class DT, int S>
static int conv_pipe_s(const ConvArgs& a, int M1, int kind, float2* ring, unsigned* ctl, void* stream) {
#define HY_CASE_K(MM, KK) \
  case KK: return conv_pipe_m<DT, MM, S, KK>(a, ring, ctl, stream);
#define HY_CASE(MM)                                                         \
  case MM:                                                                  \
    switch (kind) {                                                         \
      HY_CASE_K(MM, HY_PIPE_FWD)                                            \
      HY_CASE_K(MM, HY_PIPE_BWDG)                                           \
      default:                                                              \
        if constexpr (!DT::kBf16) {                                         \
          switch (kind) {                                                   \
            HY_CASE_K(MM, HY_PIPE_SPEC)                                     \
            HY_CASE_K(MM, HY_PIPE_DK)                                       \
          }                                                                 \
        }                                                                   \
    }                                                                       \
    return HY_ERR_UNSUPPORTED;
  switch (M1) {
#ifdef HY_EMU_BUILD
    HY_CASE(4)
    HY_CASE(16)
    HY_CASE(32)
#endif
    HY_CASE(64)
    HY_CASE(128)
    HY_CASE(256)
    HY_CASE(512)
  }
#undef HY_CASE
#undef HY_CASE_K
  return HY_ERR_UNSUPPORTED;
}

// HY_ERR_UNSUPPORTED (error text untouched) = this geometry has no pipeline instance: take the per-phase launches
template <class DT>
int launch_conv_pipe(const ConvArgs& a, int M1, int S, int kind, float2* ring, unsigned* ctl, void* stream) {
  switch (S) {
    case 4096: return conv_pipe_s<DT, 4096>(a, M1, kind, ring, ctl, stream);
#ifdef HY_EMU_BUILD
    case 256: return conv_pipe_s<DT, 256>(a, M1, kind, ring, ctl, stream);
#endif
  }
  return HY_ERR_UNSUPPORTED;
}

}  // namespace hy
