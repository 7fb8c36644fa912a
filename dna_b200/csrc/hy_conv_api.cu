// hyena-b200: C-ABI entry points of the fused long convolution + dtype-independent kernels.
#include "hy_conv_launch.h"
#include "hy_conv_pipe.cuh"
#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>

namespace hy {

template <int S>
static int fused_dk_s(const ConvArgs& a, void* stream) {
  constexpr int NB = 4096 / S;
  auto kern = k_fused_dk<S, NB, kNT>;
  const size_t smem = sizeof(float4) * Plan<S>::tw_slots() + sizeof(float2) * NB * RowSmem<S>::kRow;
  HY_LAUNCH(kern, (a.nrows + NB - 1) / NB, kNT, smem, stream, a);
  return check_launch("k_fused_dk");
}
int launch_fused_dk(const ConvArgs& a, int S, void* stream) {
  switch (S) {
    case 256: return fused_dk_s<256>(a, stream);
    case 512: return fused_dk_s<512>(a, stream);
    case 1024: return fused_dk_s<1024>(a, stream);
    case 2048: return fused_dk_s<2048>(a, stream);
    case 4096: return fused_dk_s<4096>(a, stream);
  }
  return fail(HY_ERR_UNSUPPORTED, "dk: unsupported transform length %d", S);
}

template <int S, int MODE>
static int row_conv_s(const ConvArgs& a, void* stream) {
  constexpr int NSEQ = (MODE == HY_PW_BWD) ? 2 : 1;
  constexpr int NT = (MODE == HY_PW_BWD) ? kNTRowBwd : kNT;
  auto kern = k_row_conv<S, NT, MODE>;
  const size_t smem = sizeof(float4) * Plan<S>::tw_slots() + sizeof(float2) * 2 * NSEQ * RowSmem<S>::kRow;
  const int npair = a.M1 / 2;
  HY_LAUNCH(kern, dim3(npair, a.nrows), NT, smem, stream, a);
  return check_launch("k_row_conv");
}
template <int S>
static int row_conv_mode(const ConvArgs& a, int mode, void* stream) {
  switch (mode) {
    case HY_PW_CONV: return row_conv_s<S, HY_PW_CONV>(a, stream);
    case HY_PW_SPEC: return row_conv_s<S, HY_PW_SPEC>(a, stream);
    case HY_PW_BWD: return row_conv_s<S, HY_PW_BWD>(a, stream);
    case HY_PW_REPACK: return row_conv_s<S, HY_PW_REPACK>(a, stream);
    case HY_PW_BWDG: return row_conv_s<S, HY_PW_BWDG>(a, stream);
  }
  return fail(HY_ERR_ARG, "row_conv: bad mode %d", mode);
}
int launch_row_conv(const ConvArgs& a, int M1, int S, int mode, void* stream) {
  (void)M1;
  switch (S) {
    case 4096: return row_conv_mode<4096>(a, mode, stream);
#ifdef HY_EMU_BUILD
    case 256: return row_conv_mode<256>(a, mode, stream);
#endif
  }
  return fail(HY_ERR_UNSUPPORTED, "four-step: unsupported row length %d", S);
}

// ---- geometry -----------------------------------------------------------------------------------
struct Geo {
  int M;       // complex transform length
  int S;       // row length (== M in the fused regime)
  int M1;      // column length (1 in the fused regime)
  bool fused;
};
static int fft_len(int L) {
  int M = 256;
  while (M < L) M <<= 1;
  return M;
}
// Transform lengths 5 * 2^k and 3 * 2^k (four-step regime only: column lengths M1 = 10, 12, 20, 24, ...): the reference
// transforms exactly 2L points (hyena.py:61-62), rounding L up to a power of two can almost double the work
// (L = 160 000: M = 40 x 4096 instead of 2^18).  HYENA_B200_POW2_ONLY=1 / hy_debug_set_odd_lengths(0) turn them off.
static int g_odd_lengths = -1;
static bool odd_lengths_enabled() {
  if (g_odd_lengths < 0) {
    const char* e = getenv("HYENA_B200_POW2_ONLY");
    g_odd_lengths = (e && e[0] == '1') ? 0 : 1;
  }
  return g_odd_lengths != 0;
}
static bool geometry(int L, Geo* g) {
  if (L < 1 || L > (1 << 21)) return false;
  g->M = fft_len(L);
  int blk = 4096;
#ifdef HY_EMU_BUILD
  if (g_debug_block) blk = g_debug_block;
#endif
  if (g->M <= blk) {
    g->S = g->M;
    g->M1 = 1;
    g->fused = true;
  } else {
    g->S = blk;
    g->M1 = g->M / blk;
    g->fused = false;
    if (!valid_cols(g->M1)) return false;
    if (odd_lengths_enabled()) {
      const int need = (L + blk - 1) / blk;   // smallest column length that holds L points
#define HY_CASE(MM) if (MM >= need && MM < g->M1) g->M1 = MM;
      HY_COLS_ODD(HY_CASE)
#undef HY_CASE
      g->M = g->M1 * blk;
    }
  }
  return true;
}
// rows per group: the scratch budget is shared by the g_nstream groups in flight
static long long group_rows(const Geo& g, int nseq, long long rows, size_t ws_bytes) {
  const size_t per_row = sizeof(float2) * (size_t)g.M * nseq;
  const int ns = std::max(1, g_nstream);
  long long by_budget = std::max<long long>(1, (long long)(g_scratch_budget / per_row / ns));
  long long by_ws = (long long)(ws_bytes / per_row / ns);
  if (by_ws < 1) by_ws = (long long)(ws_bytes / per_row);   // tiny workspace: single region
  // the row kernels launch dim3(pairs, rows x nseq): gridDim.y is limited to 65535
  const long long by_grid = 65535 / std::max(1, nseq);
  return std::min(std::min(rows, by_grid), std::min(by_budget, by_ws));
}

// ---- group pipeline: groups round-robin over internal streams, forked from / joined into the caller's stream --
#ifndef HY_EMU_BUILD
struct SideStreams {
  cudaStream_t s[4];
  cudaEvent_t fork, join[4];
  bool ok = false;
  std::mutex use;      // the fork event and the streams are shared by every host thread using this device
};
static SideStreams* side_streams() {
  static std::mutex mu;
  static std::map<int, SideStreams*> per_dev;
  std::lock_guard<std::mutex> lk(mu);
  int dev = 0;
  cudaGetDevice(&dev);
  auto it = per_dev.find(dev);
  if (it != per_dev.end()) return it->second;
  SideStreams* ss = new SideStreams();
  ss->ok = true;
  for (int i = 0; i < 4; ++i) {
    if (cudaStreamCreateWithFlags(&ss->s[i], cudaStreamNonBlocking) != cudaSuccess) ss->ok = false;
    if (cudaEventCreateWithFlags(&ss->join[i], cudaEventDisableTiming) != cudaSuccess) ss->ok = false;
  }
  if (cudaEventCreateWithFlags(&ss->fork, cudaEventDisableTiming) != cudaSuccess) ss->ok = false;
  per_dev[dev] = ss;
  return ss;
}
#endif

// launch_group(row0, nrows, scratch, stream) -> status
template <class F>
static int run_groups(void* caller_stream, long long rbeg, long long rend, long long G, float2* ws, size_t ws_bytes,
                      size_t per_row_elems, F&& launch_group) {
  const long long ngroups = (rend - rbeg + G - 1) / G;
  int ns = 1;
#ifndef HY_EMU_BUILD
  SideStreams* ss = nullptr;
  if (g_nstream > 1 && ngroups > 1) {
    ss = side_streams();
    const long long regions = (long long)(ws_bytes / (sizeof(float2) * per_row_elems * (size_t)G));
    ns = (int)std::min<long long>(std::min<long long>(g_nstream, ngroups), regions);
    if (!ss->ok || ns < 2) ns = 1;
  }
  if (ns > 1) {
    // one fork / launch / join sequence at a time per device: two host threads must not wait on each other's fork record
    std::lock_guard<std::mutex> lk(ss->use);
    cudaStream_t cs = (cudaStream_t)caller_stream;
    if (cudaEventRecord(ss->fork, cs) != cudaSuccess) return fail(HY_ERR_CUDA, "pipeline fork failed");
    for (int i = 0; i < ns; ++i)
      if (cudaStreamWaitEvent(ss->s[i], ss->fork, 0) != cudaSuccess) return fail(HY_ERR_CUDA, "pipeline fork wait failed");
    if (g_persist_l2) {
      for (int i = 0; i < ns; ++i) {
        cudaStreamAttrValue v;
        memset(&v, 0, sizeof(v));
        v.accessPolicyWindow.base_ptr = (void*)(ws + (size_t)i * (size_t)G * per_row_elems);
        v.accessPolicyWindow.num_bytes = sizeof(float2) * (size_t)G * per_row_elems;
        v.accessPolicyWindow.hitRatio = 1.0f;
        v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cudaStreamSetAttribute(ss->s[i], cudaStreamAttributeAccessPolicyWindow, &v);
      }
    }
    long long gi = 0;
    for (long long r0 = rbeg; r0 < rend; r0 += G, ++gi) {
      const int si = (int)(gi % ns);
      int rc = launch_group(r0, std::min(G, rend - r0), ws + (size_t)si * (size_t)G * per_row_elems, (void*)ss->s[si]);
      if (rc != HY_OK) return rc;
    }
    int jrc = HY_OK;
    for (int i = 0; i < ns; ++i) {
      if (cudaEventRecord(ss->join[i], ss->s[i]) != cudaSuccess || cudaStreamWaitEvent(cs, ss->join[i], 0) != cudaSuccess)
        jrc = HY_ERR_CUDA;
      if (g_persist_l2) {      // drop the access-policy window again: the side streams outlive this call
        cudaStreamAttrValue v;
        memset(&v, 0, sizeof(v));
        cudaStreamSetAttribute(ss->s[i], cudaStreamAttributeAccessPolicyWindow, &v);
      }
    }
    if (jrc != HY_OK) return fail(HY_ERR_CUDA, "pipeline join failed");
    return check_launch("pipeline join");
  }
#endif
  for (long long r0 = rbeg; r0 < rend; r0 += G) {
    int rc = launch_group(r0, std::min(G, rend - r0), ws, caller_stream);
    if (rc != HY_OK) return rc;
  }
  return HY_OK;
}

// ---- persistent pipeline (hy_conv_pipe.cuh): one launch for all rows, scratch = a ring of kPipeRingMax row buffers ------
static int g_pipe = -1;
static bool pipe_enabled() {
  if (g_pipe < 0) {
    // opt-in: on B200 the pipeline cuts the family's DRAM traffic (47.9 -> 29.0 GB per layer at 1 M) but runs 20-25 %
    // slower than the per-phase launches — the phase bodies are issue/barrier bound, not DRAM bound, and the per-item
    // fence + arrival (about 3000 cycles) is pure overhead (profiles/r02c_conv_pipeline_vs_per_phase_lag_sweep_negative.txt)
    const char* e = getenv("HYENA_B200_CONV_PIPE");
    g_pipe = (e && e[0] == '1') ? 1 : 0;
  }
  return g_pipe != 0;
}
// debug statistics of the pipeline schedule (hy_debug_pipe_stats): 16 u64 counters in device memory, see hy_conv_pipe.cuh
static unsigned long long* g_pipe_stats = nullptr;
static bool g_pipe_stats_on = false;
unsigned long long* pipe_stats_buffer() { return g_pipe_stats_on ? g_pipe_stats : nullptr; }
static int g_pipe_lag = 0;
int pipe_lag_override() { return g_pipe_lag; }
static size_t pipe_ws_bytes(const Geo& g, long long rows) {
  return sizeof(float2) * (size_t)g.M * kPipeRingMax + 256 + pipe_ctl_bytes(rows);
}
// HY_ERR_UNSUPPORTED: not applicable here (disabled, workspace too small, no instance) -> per-phase launches
template <class DT>
static int try_pipe(const ConvArgs& a, const Geo& g, int kind, long long rbeg, long long rend, void* ws, size_t ws_bytes,
                    void* stream) {
  if (!pipe_enabled() || g.fused || !ws || ws_bytes < pipe_ws_bytes(g, rend - rbeg)) return HY_ERR_UNSUPPORTED;
  if (DT::kBf16 ? !a.stage_ok : !a.vec_all) return HY_ERR_UNSUPPORTED;
  ConvArgs b = a;
  b.row_begin = (int)rbeg;
  b.nrows = (int)(rend - rbeg);
  b.scratch = reinterpret_cast<float2*>(ws);
  const size_t ring = (sizeof(float2) * (size_t)g.M * kPipeRingMax + 255) & ~(size_t)255;
  unsigned* ctl = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(ws) + ring);
  return launch_conv_pipe<DT>(b, g.M1, g.S, kind, b.scratch, ctl, stream);
}

template <class T>
static bool aligned2(const void* p) { return (reinterpret_cast<uintptr_t>(p) % (2 * sizeof(T))) == 0; }
static bool vec_ok(int dtype, std::initializer_list<const void*> ptrs, long long bs, int ld) {
  if ((ld & 1) || (bs & 1)) return false;
  for (const void* p : ptrs) {
    if (!p) continue;
    if (dtype == HY_BF16 ? !aligned2<unsigned short>(p) : !aligned2<float>(p)) return false;
  }
  return true;
}

// cp.async staging of the SHORTCONV source rows: bf16, 16-byte aligned base, strides multiple of 8 elements
static int stage_ok(int dtype, int in_mode, const void* u, long long bs, int ld) {
  return dtype == HY_BF16 && in_mode == HY_IN_SHORTCONV && (ld % 8 == 0) && (bs % 8 == 0) &&
         (reinterpret_cast<uintptr_t>(u) % 16 == 0);
}

static int check_modes(int in_mode, int out_mode) {
  if (in_mode < 0 || in_mode > 2 || out_mode < 0 || out_mode > 2) return fail(HY_ERR_ARG, "bad gating mode");
  if ((in_mode == HY_IN_SHORTCONV) != (out_mode == HY_OUT_SHORTCONV))
    return fail(HY_ERR_ARG, "SHORTCONV must be selected for both the input and the output gate");
  return HY_OK;
}

template <class DT>
static int conv_fwd_t(const hy_conv_fwd_args* p, void* stream) {
  Geo g;
  if (!geometry(p->L, &g)) return fail(HY_ERR_UNSUPPORTED, "unsupported sequence length %d", p->L);
  const Tables* tb = tables();
  if (!tb) return HY_ERR_CUDA;
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  a.u = p->u; a.pre = p->pre; a.post = p->post; a.out = p->out; a.ysave = p->ysave;
  a.u_bs = p->u_bs; a.ldu = p->ldu; a.out_bs = p->out_bs; a.ldo = p->ldo; a.post_bs = p->post_bs; a.ldpost = p->ldpost;
  a.sw = p->sw; a.sb = p->sb; a.pb = p->pb;
  a.Kf = reinterpret_cast<const float2*>(p->Kf);
  a.tw = tb->tw; a.twpos = tb->twpos[hy_ilog2(g.S)];
  a.B = p->B; a.H = p->H; a.L = p->L; a.M1 = g.M1; a.S = g.S;
  a.in_mode = p->in_mode; a.out_mode = p->out_mode;
  a.vec_all = vec_ok(p->dtype, {p->u, p->pre}, p->u_bs, p->ldu) && vec_ok(p->dtype, {p->out, p->ysave}, p->out_bs, p->ldo) &&
              vec_ok(p->dtype, {p->post}, p->post_bs, p->ldpost);
  a.stage_ok = a.vec_all && stage_ok(p->dtype, p->in_mode, p->u, p->u_bs, p->ldu);
  a.vec8_out = a.stage_ok && (p->ldo % 8 == 0) && (p->out_bs % 8 == 0) && (reinterpret_cast<uintptr_t>(p->out) % 16 == 0) &&
               (reinterpret_cast<uintptr_t>(p->ysave) % 16 == 0);
  if (p->dtype == HY_F32) {
    a.vec8_out = p->out_mode == HY_OUT_PLAIN && (p->ldo % 4 == 0) && (p->out_bs % 4 == 0) && (reinterpret_cast<uintptr_t>(p->out) % 16 == 0);
    a.vec16_in = p->in_mode == HY_IN_PLAIN && (p->ldu % 4 == 0) && (p->u_bs % 4 == 0) && (reinterpret_cast<uintptr_t>(p->u) % 16 == 0);
  }
  a.scratch = reinterpret_cast<float2*>(p->ws);
  const long long rows = (long long)p->B * p->H;
  a.gsave = reinterpret_cast<float2*>(p->gsave);
  if (g.fused) {
    a.row_begin = 0;
    a.nrows = (int)rows;
    return launch_fused_fwd<DT>(a, g.S, HY_PW_CONV, stream);
  }
  {
    const int rc = try_pipe<DT>(a, g, HY_PIPE_FWD, 0, rows, p->ws, p->ws_bytes, stream);
    if (rc != HY_ERR_UNSUPPORTED) return rc;
  }
  const long long G = group_rows(g, 1, rows, p->ws_bytes);
  if (G < 1 || !p->ws) return fail(HY_ERR_WORKSPACE, "hy_conv_fwd: workspace too small (%zu bytes)", p->ws_bytes);
  return run_groups(stream, 0, rows, G, a.scratch, p->ws_bytes, (size_t)g.M, [&](long long r0, long long n, float2* scr, void* st) {
    ConvArgs b = a;
    b.row_begin = (int)r0;
    b.nrows = (int)n;
    b.scratch = scr;
    int rc;
    if ((rc = launch_col_fwd<DT>(b, g.M1, g.S, 1, 0, st)) != HY_OK) return rc;
    if ((rc = launch_row_conv(b, g.M1, g.S, HY_PW_CONV, st)) != HY_OK) return rc;
    return launch_col_inv<DT>(b, g.M1, g.S, 1, 0, st);
  });
}

template <class DT>
static int conv_bwd_t(const hy_conv_bwd_args* p, void* stream) {
  Geo g;
  if (!geometry(p->L, &g)) return fail(HY_ERR_UNSUPPORTED, "unsupported sequence length %d", p->L);
  const Tables* tb = tables();
  if (!tb) return HY_ERR_CUDA;
  if (p->nslot < 1 || p->nslot > p->B) return fail(HY_ERR_ARG, "hy_conv_bwd: nslot must be in [1, B]");
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  a.u = p->u; a.pre = p->pre; a.post = p->post; a.dout = p->dout; a.ysave_in = p->ysave;
  a.du = p->du; a.dpre = p->dpre; a.dpost = p->dpost;
  a.u_bs = p->u_bs; a.ldu = p->ldu; a.out_bs = p->out_bs; a.ldo = p->ldo; a.post_bs = p->post_bs; a.ldpost = p->ldpost;
  a.ys_bs = p->ys_bs; a.ldys = p->ldys;
  a.sw = p->sw; a.sb = p->sb; a.pb = p->pb;
  a.Kf = reinterpret_cast<const float2*>(p->Kf);
  a.dKacc = reinterpret_cast<float2*>(p->dKacc);
  a.dDpart = p->dDpart;
  a.ndpart = hy_conv_ndpart(p->L);
  a.tw = tb->tw; a.twpos = tb->twpos[hy_ilog2(g.S)];
  a.B = p->B; a.H = p->H; a.L = p->L; a.M1 = g.M1; a.S = g.S;
  a.in_mode = p->in_mode; a.out_mode = p->out_mode;
  a.vec_all = vec_ok(p->dtype, {p->u, p->pre, p->du, p->dpre}, p->u_bs, p->ldu) && vec_ok(p->dtype, {p->dout}, p->out_bs, p->ldo) &&
              vec_ok(p->dtype, {p->ysave}, p->ys_bs, p->ldys) && vec_ok(p->dtype, {p->post, p->dpost}, p->post_bs, p->ldpost);
  a.stage_ok = a.vec_all && stage_ok(p->dtype, p->in_mode, p->u, p->u_bs, p->ldu);
  a.stage_dz_ok = a.stage_ok && (p->ldo % 8 == 0) && (p->out_bs % 8 == 0) && (reinterpret_cast<uintptr_t>(p->dout) % 16 == 0);
  a.vec8_out = a.stage_ok && (reinterpret_cast<uintptr_t>(p->du) % 16 == 0);   // du has the strides of u
  a.scratch = reinterpret_cast<float2*>(p->ws);
  a.defer_dx0 = (p->defer_dx0 && p->out_mode == HY_OUT_SHORTCONV) ? 1 : 0;
  if (p->out_mode != HY_OUT_PLAIN && !p->ysave && !a.defer_dx0) return fail(HY_ERR_ARG, "hy_conv_bwd: gated output modes need ysave");
  for (int b0 = 0; b0 < p->B; b0 += p->nslot) {
    const int b1 = std::min(p->B, b0 + p->nslot);
    a.slot_b0 = b0;
    a.accumulate = b0 > 0;
    const long long rbeg = (long long)b0 * p->H, rend = (long long)b1 * p->H;
    if (g.fused) {
      a.row_begin = (int)rbeg;
      a.nrows = (int)(rend - rbeg);
      a.gsave = reinterpret_cast<float2*>(const_cast<void*>(p->gsave));
      int rc = p->gsave ? launch_fused_bwdg<DT>(a, g.S, stream) : launch_fused_bwd<DT>(a, g.S, stream);
      if (rc != HY_OK) return rc;
      continue;
    }
    // with the forward's saved spectrum of g only dy is transformed: one sequence per row
    const int nseq = p->gsave ? 1 : 2;
    a.gsave = reinterpret_cast<float2*>(const_cast<void*>(p->gsave));
    if (nseq == 1) {
      const int rcp = try_pipe<DT>(a, g, HY_PIPE_BWDG, rbeg, rend, p->ws, p->ws_bytes, stream);
      if (rcp == HY_OK) continue;
      if (rcp != HY_ERR_UNSUPPORTED) return rcp;
    }
    const long long G = group_rows(g, nseq, rend - rbeg, p->ws_bytes);
    if (G < 1 || !p->ws) return fail(HY_ERR_WORKSPACE, "hy_conv_bwd: workspace too small (%zu bytes)", p->ws_bytes);
    int rc = run_groups(stream, rbeg, rend, G, a.scratch, p->ws_bytes, nseq * (size_t)g.M, [&](long long r0, long long n, float2* scr, void* st) {
      ConvArgs b = a;
      b.row_begin = (int)r0;
      b.nrows = (int)n;
      b.scratch = scr;
      int rc2;
      if ((rc2 = launch_col_fwd<DT>(b, g.M1, g.S, nseq, nseq == 1, st)) != HY_OK) return rc2;
      if ((rc2 = launch_row_conv(b, g.M1, g.S, nseq == 1 ? HY_PW_BWDG : HY_PW_BWD, st)) != HY_OK) return rc2;
      return launch_col_inv<DT>(b, g.M1, g.S, nseq, 1, st);
    });
    if (rc != HY_OK) return rc;
  }
  return HY_OK;
}

}  // namespace hy

using namespace hy;

extern "C" {

// experiment / test hook (not in the public header): 0 = per-phase launches over row groups, 1 = persistent pipeline
int hy_debug_set_conv_pipe(int on) {
  hy::g_pipe = on ? 1 : 0;
  return hy::g_pipe;
}

// experiment hook: rows the B items trail the A items (and C the B items); 0 = automatic
int hy_debug_set_pipe_lag(int lag) {
  hy::g_pipe_lag = lag;
  return lag;
}

// experiment hook: mode 1 = start collecting (zeroes the counters), 0 = stop; out16 (host, nullable) receives the counters
int hy_debug_pipe_stats(int mode, unsigned long long* out16) {
#ifndef HY_EMU_BUILD
  if (!hy::g_pipe_stats) {
    hy::g_pipe_stats = reinterpret_cast<unsigned long long*>(hy::table_alloc(16 * sizeof(unsigned long long)));
    if (!hy::g_pipe_stats) return HY_ERR_CUDA;
    cudaMemset(hy::g_pipe_stats, 0, 16 * sizeof(unsigned long long));
  }
  cudaDeviceSynchronize();
  if (out16) cudaMemcpy(out16, hy::g_pipe_stats, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  if (mode == 1) cudaMemset(hy::g_pipe_stats, 0, 16 * sizeof(unsigned long long));
  hy::g_pipe_stats_on = mode == 1;
#else
  (void)mode; (void)out16;
#endif
  return HY_OK;
}

int hy_debug_set_odd_lengths(int on) {
  g_odd_lengths = on ? 1 : 0;
  return g_odd_lengths;
}

int hy_fft_len(int L) {
  Geo g;
  if (!geometry(L, &g)) return -1;
  return g.M;
}

size_t hy_conv_gsave_bytes(int B, int H, int L) {
  Geo g;
  if (!geometry(L, &g) || B < 1 || H < 1) return 0;
  // single-kernel regime: the saved spectrum pays from M = 2048 on (backward 3.13 -> 2.16 ms at L = 4096, 3.63 -> 2.62 at
  // 2048; at 1024 the extra 8 B per point written and read back cost more than the transform they save)
  if (g.fused && g.M < 2048) return 0;
  return sizeof(float2) * (size_t)g.M * (size_t)B * (size_t)H;
}

int hy_conv_ndpart(int L) {
  Geo g;
  if (!geometry(L, &g)) return -1;
  return g.fused ? 1 : g.S / col_T2(g.M1);
}

size_t hy_conv_workspace_bytes(int B, int H, int L, int nseq) {
  Geo g;
  if (!geometry(L, &g)) return 0;
  if (g.fused) return 0;
  const size_t per_row = sizeof(float2) * (size_t)g.M * (size_t)nseq;
  const int ns = std::max(1, g_nstream);
  long long G = std::max<long long>(1, (long long)(g_scratch_budget / per_row / ns));
  G = std::min<long long>(G, (long long)B * H);
  long long regions = std::min<long long>(ns, ((long long)B * H + G - 1) / G);
  // enough for the per-phase launches over row groups AND for the persistent pipeline (ring + arrival counters)
  return std::max(per_row * (size_t)G * (size_t)regions, pipe_ws_bytes(g, (long long)B * H));
}

int hy_filter_spectrum(const float* k, int ldk, const float* D, void* Kf, int H, int L, void* ws, size_t ws_bytes,
                       void* stream) {
  Geo g;
  if (!k || !Kf || H < 1) return fail(HY_ERR_ARG, "hy_filter_spectrum: bad argument");
  if (!geometry(L, &g)) return fail(HY_ERR_UNSUPPORTED, "unsupported sequence length %d", L);
  const Tables* tb = tables();
  if (!tb) return HY_ERR_CUDA;
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  a.u = k; a.ldu = ldk; a.u_bs = 0;
  a.Kf_out = reinterpret_cast<float2*>(Kf);
  a.skipD = D;
  a.scale = 1.0f / (float)g.M;
  a.tw = tb->tw; a.twpos = tb->twpos[hy_ilog2(g.S)];
  a.B = 1; a.H = H; a.L = L; a.M1 = g.M1; a.S = g.S;
  a.in_mode = HY_IN_PLAIN; a.out_mode = HY_OUT_PLAIN;
  a.vec_all = vec_ok(HY_F32, {k}, 0, ldk);
  a.vec16_in = (ldk % 4 == 0) && (reinterpret_cast<uintptr_t>(k) % 16 == 0);
  a.scratch = reinterpret_cast<float2*>(ws);
  if (g.fused) {
    a.row_begin = 0; a.nrows = H;
    return launch_fused_fwd<DT_F32>(a, g.S, HY_PW_SPEC, stream);
  }
  {
    const int rc = try_pipe<DT_F32>(a, g, HY_PIPE_SPEC, 0, H, ws, ws_bytes, stream);
    if (rc != HY_ERR_UNSUPPORTED) return rc;
  }
  const long long G = group_rows(g, 1, H, ws_bytes);
  if (G < 1 || !ws) return fail(HY_ERR_WORKSPACE, "hy_filter_spectrum: workspace too small (%zu bytes)", ws_bytes);
  return run_groups(stream, 0, H, G, a.scratch, ws_bytes, (size_t)g.M, [&](long long r0, long long n, float2* scr, void* st) {
    ConvArgs b = a;
    b.row_begin = (int)r0;
    b.nrows = (int)n;
    b.scratch = scr;
    int rc;
    if ((rc = launch_col_fwd<DT_F32>(b, g.M1, g.S, 1, 0, st)) != HY_OK) return rc;
    return launch_row_conv(b, g.M1, g.S, HY_PW_SPEC, st);
  });
}

int hy_conv_fwd(const hy_conv_fwd_args* p, void* stream) {
  if (!p || !p->u || !p->out || !p->Kf || p->B < 1 || p->H < 1) return fail(HY_ERR_ARG, "hy_conv_fwd: bad argument");
  int rc = check_modes(p->in_mode, p->out_mode);
  if (rc != HY_OK) return rc;
  if (p->in_mode == HY_IN_PREGATE && !p->pre) return fail(HY_ERR_ARG, "hy_conv_fwd: PREGATE needs pre");
  if (p->out_mode == HY_OUT_POSTGATE && !p->post) return fail(HY_ERR_ARG, "hy_conv_fwd: POSTGATE needs post");
  if (p->in_mode == HY_IN_SHORTCONV && (!p->sw || !p->sb)) return fail(HY_ERR_ARG, "hy_conv_fwd: SHORTCONV needs sw and sb");
  if (p->dtype == HY_F32) return conv_fwd_t<DT_F32>(p, stream);
  if (p->dtype == HY_BF16) return conv_fwd_t<DT_BF16>(p, stream);
  return fail(HY_ERR_UNSUPPORTED, "hy_conv_fwd: unsupported dtype %d", p->dtype);
}

int hy_conv_bwd(const hy_conv_bwd_args* p, void* stream) {
  if (!p || !p->u || !p->dout || !p->Kf || !p->du || !p->dKacc || (!p->dDpart && !p->gsave) || p->B < 1 || p->H < 1)
    return fail(HY_ERR_ARG, "hy_conv_bwd: bad argument");
  int rc = check_modes(p->in_mode, p->out_mode);
  if (rc != HY_OK) return rc;
  if (p->in_mode == HY_IN_PREGATE && (!p->pre || !p->dpre)) return fail(HY_ERR_ARG, "hy_conv_bwd: PREGATE needs pre and dpre");
  if (p->out_mode == HY_OUT_POSTGATE && (!p->post || !p->dpost)) return fail(HY_ERR_ARG, "hy_conv_bwd: POSTGATE needs post and dpost");
  if (p->in_mode == HY_IN_SHORTCONV && (!p->sw || !p->sb)) return fail(HY_ERR_ARG, "hy_conv_bwd: SHORTCONV needs sw and sb");
  if (p->dtype == HY_F32) return conv_bwd_t<DT_F32>(p, stream);
  if (p->dtype == HY_BF16) return conv_bwd_t<DT_BF16>(p, stream);
  return fail(HY_ERR_UNSUPPORTED, "hy_conv_bwd: unsupported dtype %d", p->dtype);
}

int hy_conv_dk(const void* dKacc, int nslot, float* dk, int lddk, int H, int L, void* ws, size_t ws_bytes, void* stream) {
  Geo g;
  if (!dKacc || !dk || nslot < 1 || H < 1) return fail(HY_ERR_ARG, "hy_conv_dk: bad argument");
  if (!geometry(L, &g)) return fail(HY_ERR_UNSUPPORTED, "unsupported sequence length %d", L);
  const Tables* tb = tables();
  if (!tb) return HY_ERR_CUDA;
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  a.dKacc = reinterpret_cast<float2*>(const_cast<void*>(dKacc));
  a.nslot = nslot;
  a.scale = 1.0f / (float)g.M;
  a.out = dk; a.ldo = lddk; a.out_bs = 0;
  a.tw = tb->tw; a.twpos = tb->twpos[hy_ilog2(g.S)];
  a.B = 1; a.H = H; a.L = L; a.M1 = g.M1; a.S = g.S;
  a.in_mode = HY_IN_PLAIN; a.out_mode = HY_OUT_PLAIN;
  a.vec_all = vec_ok(HY_F32, {dk}, 0, lddk);
  a.vec8_out = (lddk % 4 == 0) && (reinterpret_cast<uintptr_t>(dk) % 16 == 0);
  a.scratch = reinterpret_cast<float2*>(ws);
  if (g.fused) {
    a.row_begin = 0; a.nrows = H;
    return launch_fused_dk(a, g.S, stream);
  }
  {
    const int rc = try_pipe<DT_F32>(a, g, HY_PIPE_DK, 0, H, ws, ws_bytes, stream);
    if (rc != HY_ERR_UNSUPPORTED) return rc;
  }
  const long long G = group_rows(g, 1, H, ws_bytes);
  if (G < 1 || !ws) return fail(HY_ERR_WORKSPACE, "hy_conv_dk: workspace too small (%zu bytes)", ws_bytes);
  return run_groups(stream, 0, H, G, a.scratch, ws_bytes, (size_t)g.M, [&](long long r0, long long n, float2* scr, void* st) {
    ConvArgs b = a;
    b.row_begin = (int)r0;
    b.nrows = (int)n;
    b.scratch = scr;
    int rc;
    if ((rc = launch_row_conv(b, g.M1, g.S, HY_PW_REPACK, st)) != HY_OK) return rc;
    return launch_col_inv<DT_F32>(b, g.M1, g.S, 1, 0, st);
  });
}

}  // extern "C"
