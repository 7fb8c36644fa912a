// hyena-b200: short depthwise causal convolution (HyenaOperator.short_filter,
// /root/reference/src/models/sequence/hyena.py:407-413 and its use at :444):
//     xc[b][ch][t] = sb[ch] + sum_j sw[ch][j] * x[b][ch][t-2+j],   x = uT + pb,  x[<0] = 0
// The forward of this op is fused into the long-conv kernels (hy_conv.cuh, ShortConvRow); this file
// holds its backward (the long-conv backward leaves dX = (dx0|dx1|dv) in uT layout) and a
// standalone forward used by tests and by callers that want the conv alone.
#include "hy_host.h"

namespace hy {

constexpr int kScThreads = 256;
constexpr int kScPairs = 16;                                 // element pairs per thread (8192 elements of a row per CTA:
                                                             // the 5-value block reduction is paid once per 32 elements per thread)
constexpr int kScChunk = kScThreads * kScPairs * 2;          // elements of one row per CTA

template <class DT>
struct ScRow {
  const typename DT::elem* p;
  float pb;
  bool has_pb, vec;
  int L;
  HY_DEVICE float fix(float r) const {
    if (has_pb) {
      r += pb;
      if (DT::kBf16) r = round_to_bf16(r);
    }
    return r;
  }
  // x[t], x[t+1] for even t (zero outside [0, L))
  HY_DEVICE float2 pair(int t) const {
    if (t < 0 || t >= L) return make_float2(0.f, 0.f);
    if (t + 1 < L) {
      float2 r = ld2<DT>(p + t, vec);
      return make_float2(fix(r.x), fix(r.y));
    }
    return make_float2(fix(ld1<DT>(p + t)), 0.f);
  }
};

template <class DT>
HY_DEVICE float2 ld_pair0(const typename DT::elem* p, int t, int L, bool vec) {
  if (t < 0 || t >= L) return make_float2(0.f, 0.f);
  if (t + 1 < L) return ld2<DT>(p + t, vec);
  return make_float2(ld1<DT>(p + t), 0.f);
}

HY_DEVICE float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// grid: (ceil(L / kScChunk), H3, B)
template <class DT>
__global__ void __launch_bounds__(kScThreads) k_shortconv_bwd(const typename DT::elem* uT, const typename DT::elem* dX,
                                                            typename DT::elem* duT, long long bs, int ld,
                                                            const float* sw, const float* pb, float* dwpart,
                                                            float* dpbpart, int H3, int L, int vec) {
  HY_STATIC_SMEM(float, red, 5 * (kScThreads / 32));
  const int ch = blockIdx.y, b = blockIdx.z;
  const long long roff = (long long)b * bs + (long long)ch * ld;
  ScRow<DT> x;
  x.p = uT + roff;
  x.has_pb = pb != nullptr;
  x.pb = x.has_pb ? pb[ch] : 0.f;
  x.vec = vec != 0;
  x.L = L;
  const typename DT::elem* g = dX + roff;
  typename DT::elem* o = duT + roff;
  const float w0 = sw[ch * 3 + 0], w1 = sw[ch * 3 + 1], w2 = sw[ch * 3 + 2];
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, ab = 0.f, ap = 0.f;
  const int t0 = blockIdx.x * kScChunk;
#pragma unroll
  for (int i = 0; i < kScPairs; ++i) {
    const int t = t0 + 2 * (i * kScThreads + threadIdx.x);
    if (t >= L) continue;
    // gradient of the conv input: du[t] = w2 g[t] + w1 g[t+1] + w0 g[t+2]
    const float2 gc = ld_pair0<DT>(g, t, L, vec != 0);
    const float2 gn = ld_pair0<DT>(g, t + 2, L, vec != 0);
    float d0 = fmaf(w0, gn.x, fmaf(w1, gc.y, w2 * gc.x));
    float d1 = fmaf(w0, gn.y, fmaf(w1, gn.x, w2 * gc.y));
    if (t + 1 >= L) d1 = 0.f;
    if (DT::kBf16) {
      d0 = round_to_bf16(d0);
      d1 = round_to_bf16(d1);
    }
    if (t + 1 < L) st2<DT>(o + t, make_float2(d0, d1), vec != 0);
    else st1<DT>(o + t, d0);
    ap += d0 + d1;
    // weight gradients: dw_j = sum_t g[t] x[t-2+j]
    const float2 xp = x.pair(t - 2), xc = x.pair(t);
    a0 += gc.x * xp.x + gc.y * xp.y;
    a1 += gc.x * xp.y + gc.y * xc.x;
    a2 += gc.x * xc.x + gc.y * xc.y;
    ab += gc.x + gc.y;
  }
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); ab = warp_sum(ab); ap = warp_sum(ap);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[warp * 5 + 0] = a0; red[warp * 5 + 1] = a1; red[warp * 5 + 2] = a2; red[warp * 5 + 3] = ab; red[warp * 5 + 4] = ap;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    float s = 0.f;
    for (int w = 0; w < kScThreads / 32; ++w) s += red[w * 5 + threadIdx.x];
    const long long chunk = (long long)b * gridDim.x + blockIdx.x;
    if (threadIdx.x < 4) dwpart[(chunk * H3 + ch) * 4 + threadIdx.x] = s;
    else dpbpart[chunk * H3 + ch] = s;
  }
}

// 8 consecutive elements at p (16-byte aligned for bf16, 2 x 16 bytes for fp32)
template <class DT>
HY_DEVICE void ld8(const typename DT::elem* p, float* v) {   // v[0..7]
  if (DT::kBf16) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const unsigned w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  } else {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
}
template <class DT>
HY_DEVICE void st8(typename DT::elem* p, const float (&v)[8]) {
  if (DT::kBf16) {
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = u;
  } else {
    float* q = reinterpret_cast<float*>(p);
    *reinterpret_cast<float4*>(q) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(q + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// Vectorised backward: every thread owns runs of 8 consecutive samples (one 16-byte load of dX and of uT,
// one 16-byte store of duT for bf16) — the kernel is pure streaming (3 reads... 2 rows in, 1 row out).
// grid: (ceil(L / kScChunk), H3, B); rows must be 16-byte aligned with a stride that is a multiple of 8.
// gate operands of the x0 group when dx0 = dout * ysave is formed here (hy_shortconv_bwd_gate); dz == nullptr: off
template <class DT>
struct ScGate {
  const typename DT::elem* dz;
  const typename DT::elem* ys;
  long long dz_bs, ys_bs;
  int lddz, ldys;
};

// a[0..N) rounded to bf16 in place, two values per cvt.rn.bf16x2.f32 (N even)
template <int N>
HY_DEVICE void round_pairs_to_bf16(float (&a)[N]) {
#pragma unroll
  for (int i = 0; i < N; i += 2) {
    const float2 r = round2_to_bf16(make_float2(a[i], a[i + 1]));
    a[i] = r.x; a[i + 1] = r.y;
  }
}

template <class DT>
__global__ void __launch_bounds__(kScThreads) k_shortconv_bwd_v8(const typename DT::elem* uT, const typename DT::elem* dX,
                                                               typename DT::elem* duT, long long bs, int ld,
                                                               const float* sw, const float* pb, float* dwpart,
                                                               float* dpbpart, int H3, int L, ScGate<DT> gt) {
  HY_STATIC_SMEM(float, red, 5 * (kScThreads / 32));
  const int ch = blockIdx.y, b = blockIdx.z;
  const long long roff = (long long)b * bs + (long long)ch * ld;
  const typename DT::elem* xrow = uT + roff;
  const typename DT::elem* grow = dX + roff;
  typename DT::elem* orow = duT + roff;
  const bool gated = gt.dz != nullptr && ch < H3 / 3;      // uniform per CTA
  const typename DT::elem* zrow = gated ? gt.dz + (long long)b * gt.dz_bs + (long long)ch * gt.lddz : nullptr;
  const typename DT::elem* yrow = gated ? gt.ys + (long long)b * gt.ys_bs + (long long)ch * gt.ldys : nullptr;
  const bool has_pb = pb != nullptr;
  const float pbv = has_pb ? pb[ch] : 0.f;
  const float w0 = sw[ch * 3 + 0], w1 = sw[ch * 3 + 1], w2 = sw[ch * 3 + 2];
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, ab = 0.f, ap = 0.f;
  const int c0 = blockIdx.x * kScChunk;
#pragma unroll
  for (int it = 0; it < kScChunk / (kScThreads * 8); ++it) {
    const int t0 = c0 + (it * kScThreads + threadIdx.x) * 8;
    if (t0 >= L) continue;
    float g[10], x[10];
    if (gated) {
      // dx0 = dout * y, rounded like the store the long-conv backward would have done
      float zz[10], yy[10];
      if (t0 + 8 <= L) {
        ld8<DT>(zrow + t0, zz);
        ld8<DT>(yrow + t0, yy);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          zz[i] = (t0 + i < L) ? ld1<DT>(zrow + t0 + i) : 0.f;
          yy[i] = (t0 + i < L) ? ld1<DT>(yrow + t0 + i) : 0.f;
        }
      }
#pragma unroll
      for (int i = 8; i < 10; ++i) {
        zz[i] = (t0 + i < L) ? ld1<DT>(zrow + t0 + i) : 0.f;
        yy[i] = (t0 + i < L) ? ld1<DT>(yrow + t0 + i) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 10; ++i) g[i] = zz[i] * yy[i];
      if (DT::kBf16) round_pairs_to_bf16(g);
    } else if (t0 + 8 <= L) {
      ld8<DT>(grow + t0, g);
      g[8] = (t0 + 8 < L) ? ld1<DT>(grow + t0 + 8) : 0.f;
      g[9] = (t0 + 9 < L) ? ld1<DT>(grow + t0 + 9) : 0.f;
    } else {
#pragma unroll
      for (int i = 0; i < 10; ++i) g[i] = (t0 + i < L) ? ld1<DT>(grow + t0 + i) : 0.f;
    }
    if (t0 + 8 <= L) {
      ld8<DT>(xrow + t0, x + 2);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i + 2] = (t0 + i < L) ? ld1<DT>(xrow + t0 + i) : 0.f;
    }
    x[0] = (t0 >= 2) ? ld1<DT>(xrow + t0 - 2) : 0.f;
    x[1] = (t0 >= 1) ? ld1<DT>(xrow + t0 - 1) : 0.f;
    const bool interior = t0 >= 2 && t0 + 8 <= L;   // no position of x[0..10) / d[0..8) falls outside the row
    if (has_pb) {
#pragma unroll
      for (int i = 0; i < 10; ++i) x[i] += pbv;
      if (DT::kBf16) round_pairs_to_bf16(x);
      if (!interior) {
#pragma unroll
        for (int i = 0; i < 10; ++i) {
          const int tt = t0 - 2 + i;
          if (tt < 0 || tt >= L) x[i] = 0.f;
        }
      }
    }
    float d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = fmaf(w0, g[i + 2], fmaf(w1, g[i + 1], w2 * g[i]));
    if (DT::kBf16) round_pairs_to_bf16(d);
    if (!interior) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (t0 + i >= L) d[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      ap += d[i];
      a0 = fmaf(g[i], x[i], a0);
      a1 = fmaf(g[i], x[i + 1], a1);
      a2 = fmaf(g[i], x[i + 2], a2);
      ab += g[i];
    }
    if (t0 + 8 <= L) {
      st8<DT>(orow + t0, d);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (t0 + i < L) st1<DT>(orow + t0 + i, d[i]);
    }
  }
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); ab = warp_sum(ab); ap = warp_sum(ap);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[warp * 5 + 0] = a0; red[warp * 5 + 1] = a1; red[warp * 5 + 2] = a2; red[warp * 5 + 3] = ab; red[warp * 5 + 4] = ap;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    float s = 0.f;
    for (int w = 0; w < kScThreads / 32; ++w) s += red[w * 5 + threadIdx.x];
    const long long chunk = (long long)b * gridDim.x + blockIdx.x;
    if (threadIdx.x < 4) dwpart[(chunk * H3 + ch) * 4 + threadIdx.x] = s;
    else dpbpart[chunk * H3 + ch] = s;
  }
}

template <class DT>
__global__ void __launch_bounds__(kScThreads) k_shortconv_fwd(const typename DT::elem* uT, typename DT::elem* xc,
                                                            long long bs, int ld, const float* sw, const float* sb,
                                                            const float* pb, int L, int vec) {
  const int ch = blockIdx.y, b = blockIdx.z;
  const long long roff = (long long)b * bs + (long long)ch * ld;
  ScRow<DT> x;
  x.p = uT + roff;
  x.has_pb = pb != nullptr;
  x.pb = x.has_pb ? pb[ch] : 0.f;
  x.vec = vec != 0;
  x.L = L;
  typename DT::elem* o = xc + roff;
  const float w0 = sw[ch * 3 + 0], w1 = sw[ch * 3 + 1], w2 = sw[ch * 3 + 2], bias = sb[ch];
  const int t0 = blockIdx.x * kScChunk;
#pragma unroll
  for (int i = 0; i < kScPairs; ++i) {
    const int t = t0 + 2 * (i * kScThreads + threadIdx.x);
    if (t >= L) continue;
    const float2 p = x.pair(t - 2), c = x.pair(t);
    const float o0 = fmaf(w2, c.x, fmaf(w1, p.y, fmaf(w0, p.x, bias)));
    const float o1 = fmaf(w2, c.y, fmaf(w1, c.x, fmaf(w0, p.y, bias)));
    if (t + 1 < L) st2<DT>(o + t, make_float2(o0, o1), vec != 0);
    else st1<DT>(o + t, o0);
  }
}

static bool sc_vec(int dtype, std::initializer_list<const void*> ptrs, long long bs, int ld) {
  if ((ld & 1) || (bs & 1)) return false;
  const size_t al = dtype == HY_BF16 ? 4 : 8;
  for (const void* p : ptrs)
    if (p && (reinterpret_cast<uintptr_t>(p) % al)) return false;
  return true;
}

}  // namespace hy

using namespace hy;

extern "C" {

int hy_shortconv_nchunk(int B, int L) {
  if (B < 1 || L < 1) return -1;
  return B * ((L + kScChunk - 1) / kScChunk);
}

int hy_shortconv_bwd_gate(int dtype, const void* uT, const void* dX, void* duT, long long bs, int ld, const float* sw,
                          const float* pb, float* dwpart, float* dpbpart, int B, int H3, int L, const void* dout,
                          long long dout_bs, int lddout, const void* ysave, long long ys_bs, int ldys, void* stream) {
  if (!uT || !dX || !duT || !sw || !dwpart || !dpbpart || B < 1 || H3 < 1 || L < 1)
    return fail(HY_ERR_ARG, "hy_shortconv_bwd: bad argument");
  if ((dout == nullptr) != (ysave == nullptr)) return fail(HY_ERR_ARG, "hy_shortconv_bwd_gate: dout and ysave go together");
  if (dout && H3 % 3 != 0) return fail(HY_ERR_ARG, "hy_shortconv_bwd_gate: H3 must be 3 * H");
  const dim3 grid((L + kScChunk - 1) / kScChunk, H3, B);
  const int vec = sc_vec(dtype, {uT, dX, duT}, bs, ld);
  bool vec16 = (ld % 8 == 0) && (bs % 8 == 0);
  for (const void* p : {uT, dX, (const void*)duT}) vec16 = vec16 && (reinterpret_cast<uintptr_t>(p) % 16 == 0);
  if (dout) {
    vec16 = vec16 && (lddout % 8 == 0) && (dout_bs % 8 == 0) && (ldys % 8 == 0) && (ys_bs % 8 == 0) &&
            (reinterpret_cast<uintptr_t>(dout) % 16 == 0) && (reinterpret_cast<uintptr_t>(ysave) % 16 == 0);
    if (!vec16 || (dtype != HY_F32 && dtype != HY_BF16))
      return fail(HY_ERR_UNSUPPORTED, "hy_shortconv_bwd_gate: needs 16-byte aligned rows with strides %% 8 == 0");
  }
  if (vec16 && dtype == HY_F32) {
    auto kern = k_shortconv_bwd_v8<DT_F32>;
    ScGate<DT_F32> gt{(const float*)dout, (const float*)ysave, dout_bs, ys_bs, lddout, ldys};
    HY_LAUNCH(kern, grid, kScThreads, 0, stream, (const float*)uT, (const float*)dX, (float*)duT, bs, ld, sw, pb, dwpart, dpbpart,
              H3, L, gt);
    return check_launch("k_shortconv_bwd_v8");
  }
  if (vec16 && dtype == HY_BF16) {
    auto kern = k_shortconv_bwd_v8<DT_BF16>;
    ScGate<DT_BF16> gt{(const unsigned short*)dout, (const unsigned short*)ysave, dout_bs, ys_bs, lddout, ldys};
    HY_LAUNCH(kern, grid, kScThreads, 0, stream, (const unsigned short*)uT, (const unsigned short*)dX, (unsigned short*)duT, bs, ld,
              sw, pb, dwpart, dpbpart, H3, L, gt);
    return check_launch("k_shortconv_bwd_v8");
  }
  if (dtype == HY_F32) {
    auto kern = k_shortconv_bwd<DT_F32>;
    HY_LAUNCH(kern, grid, kScThreads, 0, stream, (const float*)uT, (const float*)dX, (float*)duT, bs, ld, sw, pb, dwpart,
              dpbpart, H3, L, vec);
  } else if (dtype == HY_BF16) {
    auto kern = k_shortconv_bwd<DT_BF16>;
    HY_LAUNCH(kern, grid, kScThreads, 0, stream, (const unsigned short*)uT, (const unsigned short*)dX, (unsigned short*)duT,
              bs, ld, sw, pb, dwpart, dpbpart, H3, L, vec);
  } else {
    return fail(HY_ERR_UNSUPPORTED, "hy_shortconv_bwd: unsupported dtype %d", dtype);
  }
  return check_launch("k_shortconv_bwd");
}

int hy_shortconv_bwd(int dtype, const void* uT, const void* dX, void* duT, long long bs, int ld, const float* sw,
                     const float* pb, float* dwpart, float* dpbpart, int B, int H3, int L, void* stream) {
  return hy_shortconv_bwd_gate(dtype, uT, dX, duT, bs, ld, sw, pb, dwpart, dpbpart, B, H3, L, nullptr, 0, 0, nullptr, 0, 0,
                               stream);
}

int hy_shortconv_fwd(int dtype, const void* uT, void* xc, long long bs, int ld, const float* sw, const float* sb,
                     const float* pb, int B, int H3, int L, void* stream) {
  if (!uT || !xc || !sw || !sb || B < 1 || H3 < 1 || L < 1) return fail(HY_ERR_ARG, "hy_shortconv_fwd: bad argument");
  const dim3 grid((L + kScChunk - 1) / kScChunk, H3, B);
  const int vec = sc_vec(dtype, {uT, xc}, bs, ld);
  if (dtype == HY_F32) {
    auto kern = k_shortconv_fwd<DT_F32>;
    HY_LAUNCH(kern, grid, kScThreads, 0, stream, (const float*)uT, (float*)xc, bs, ld, sw, sb, pb, L, vec);
  } else if (dtype == HY_BF16) {
    auto kern = k_shortconv_fwd<DT_BF16>;
    HY_LAUNCH(kern, grid, kScThreads, 0, stream, (const unsigned short*)uT, (unsigned short*)xc, bs, ld, sw, sb, pb, L, vec);
  } else {
    return fail(HY_ERR_UNSUPPORTED, "hy_shortconv_fwd: unsupported dtype %d", dtype);
  }
  return check_launch("k_shortconv_fwd");
}

}  // extern "C"
