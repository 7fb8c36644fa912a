// hyena-b200: fused long-convolution kernels (forward, backward, filter spectrum, dk).
//
// Replaces, for one (batch, channel) row of length L (reference: src/models/sequence/hyena.py:60-92
// `fftconv_ref`, standalone_hyenadna.py:45-60 `fftconv`, and the gating around it in
// HyenaOperator.forward, hyena.py:444-503):
//     x0,x1,v = short_filter(in_proj(u))        (depthwise causal 3-tap, hyena.py:407-413,444)
//     g = v * x1                                (gate #1, hyena.py:481)
//     y = irfft(rfft(g, 2L) * rfft(k, 2L)/2L)[:L] + D * g      (hyena.py:60-92)
//     z = y * x0                                (gate #2, hyena.py:496-503)
// The real length-N (N = 2M, M = pow2 >= L) sequence is packed into M complex points; the
// (k, M-k) pair algebra in hy_fft.cuh turns the packed spectrum into the true one and back, so
// one M-point complex transform does the work of a 2M-point real one.  D is folded into the
// spectrum (K + D) so the skip costs nothing.
//
// Two regimes:
//   * M <= 4096 ("fused"): one CTA owns NB = 4096/M rows end-to-end in shared memory.
//   * M  > 4096 ("four-step"): M = M1 x 4096.  Phase A (column transforms over n1 + twiddle),
//     phase B (row transforms, spectrum multiply, inverse row transforms) and phase C (inverse
//     column transforms + epilogue) run as three launches over a small group of rows whose
//     complex scratch stays resident in the 126 MB L2.
#pragma once
#include "hy_fft.cuh"

enum { HY_IN_PLAIN = 0, HY_IN_PREGATE = 1, HY_IN_SHORTCONV = 2 };
enum { HY_OUT_PLAIN = 0, HY_OUT_POSTGATE = 1, HY_OUT_SHORTCONV = 2 };

struct ConvArgs {
  // forward inputs
  const void* u;       // PLAIN/PREGATE: [B][H][ldu];  SHORTCONV: uT [B][3H][ldu] (x0 | x1 | v channel groups)
  const void* pre;     // PREGATE: second factor of g, same strides as u
  const void* post;    // POSTGATE: q, strides (post_bs, ldpost)
  const void* ysave_in;  // backward: y saved by the forward (same strides as out)
  const void* dout;    // backward: gradient wrt out, strides (out_bs, ldo)
  void* out;           // forward: [B][H][ldo]
  void* ysave;         // forward: optional pre-gate y, same strides as out
  void* du;            // backward: PLAIN/PREGATE: grad wrt u (strides of u); SHORTCONV: dX [B][3H][ldu]
  void* dpre;          // backward PREGATE: grad wrt pre
  void* dpost;         // backward POSTGATE: grad wrt q (strides of post)
  long long u_bs, out_bs, post_bs, ys_bs;  // batch strides in elements
  int ldu, ldo, ldpost, ldys;
  const float* sw;     // short filter weight [3H][3]
  const float* sb;     // short filter bias [3H]
  const float* pb;     // in_proj bias [3H] (nullable)
  const float2* Kf;    // [H][M] spectrum of (k + D delta)/M in position layout
  float2* Kf_out;      // spectrum kernels: output
  const float* skipD;  // spectrum kernels: D [H] (nullable)
  float2* dKacc;       // backward: [nslot][H][M] true-spectrum products DY*conj(G)
  float* dDpart;       // backward: [B*H][ndpart] partial sums of dy*g
  float2* scratch;     // four-step: [rows][nseq][M] complex
  const float2* tw;    // W_8192 table
  const float2* twpos; // [S] W_{2S}^{freq(p)}
  const float2* twV;   // four-step: [M1][T2] W_M^{l*k1(pos1)}
  int B, H, L, M1;
  int S;               // four-step: row length (M = M1 * S)
  int row_begin, nrows;  // global (b*H + c) row range handled by this launch; scratch is indexed by local row
  int slot_b0;         // backward: dKacc slot of batch b is (b - slot_b0)
  int vec_u, vec_o, vec_q, vec_y;  // 2-element vector access allowed on the u / out / post / ysave_in families
  int in_mode, out_mode;
  int accumulate;      // backward: dKacc += instead of =
  int nslot;           // dk finalize: number of slots to sum
  int ndpart;          // row stride of dDpart
  float scale;         // spectrum / dk scaling (1/M)
};

// ---- short depthwise causal conv (hyena.py:407-413,444): out[t] = b + w0 x[t-2] + w1 x[t-1] + w2 x[t]
template <class DT>
struct ShortConvRow {
  const typename DT::elem* p;
  float w0, w1, w2, bias, pb;
  bool has_pb, vec;
  int L;
  HY_DEVICE void init(const ConvArgs& a, int b, int ch) {
    p = reinterpret_cast<const typename DT::elem*>(a.u) + (long long)b * a.u_bs + (long long)ch * a.ldu;
    w0 = a.sw[ch * 3 + 0];
    w1 = a.sw[ch * 3 + 1];
    w2 = a.sw[ch * 3 + 2];
    bias = a.sb[ch];
    has_pb = a.pb != nullptr;
    pb = has_pb ? a.pb[ch] : 0.f;
    L = a.L;
    vec = a.vec_u != 0;
  }
  HY_DEVICE float fix(float r) const {
    if (has_pb) {
      r += pb;
      if (DT::kBf16) r = round_to_bf16(r);
    }
    return r;
  }
  // raw (projected, bias-added) inputs at t and t+1 (t even, t < L)
  HY_DEVICE float2 raw_pair(int t) const {
    float x0 = 0.f, x1 = 0.f;
    if (t + 1 < L) {
      float2 r = ld2<DT>(p + t, vec);
      x0 = fix(r.x);
      x1 = fix(r.y);
    } else if (t < L) {
      x0 = fix(ld1<DT>(p + t));
    }
    return make_float2(x0, x1);
  }
  // conv outputs at t and t+1 (t even, t < L); the t+1 value is garbage-free but only valid if t+1 < L
  HY_DEVICE float2 pair(int t) const {
    float xm2 = 0.f, xm1 = 0.f;
    if (t >= 2) {
      float2 r = ld2<DT>(p + t - 2, vec);
      xm2 = fix(r.x);
      xm1 = fix(r.y);
    }
    float2 c = raw_pair(t);
    float o0 = fmaf(w2, c.x, fmaf(w1, xm1, fmaf(w0, xm2, bias)));
    float o1 = fmaf(w2, c.y, fmaf(w1, c.x, fmaf(w0, xm1, bias)));
    if (DT::kBf16) {
      o0 = round_to_bf16(o0);
      o1 = round_to_bf16(o1);
    }
    return make_float2(o0, o1);
  }
};

template <class DT>
HY_DEVICE float2 ld_pair_bounded(const typename DT::elem* p, int t, int L, bool vec) {
  if (t + 1 < L) return ld2<DT>(p + t, vec);
  if (t < L) return make_float2(ld1<DT>(p + t), 0.f);
  return make_float2(0.f, 0.f);
}
template <class DT>
HY_DEVICE void st_pair_bounded(typename DT::elem* p, int t, int L, bool vec, float2 v) {
  if (t + 1 < L) st2<DT>(p + t, v, vec);
  else if (t < L) st1<DT>(p + t, v.x);
}

// ---- per-row signal access --------------------------------------------------------------------
// Produces g (and, for the backward, dy) as packed complex samples z[n] = (s[2n], s[2n+1]) and
// consumes results.  One instance per thread; set_row() is called once per butterfly.
template <class DT>
struct RowIO {
  typedef typename DT::elem elem;
  const ConvArgs& a;
  bool valid;
  int b, c;
  // forward sources
  const elem* pu;
  const elem* ppre;
  ShortConvRow<DT> s0, s1, sv;  // x0, x1, v rows (SHORTCONV)
  bool vec_u, vec_o, vec_q;
  // sinks / gates
  const elem* pq;
  elem* pout;
  elem* pys;
  const elem* pys_in;
  const elem* pdout;
  elem *pdu, *pdpre, *pdq;

  HY_DEVICE explicit RowIO(const ConvArgs& a_) : a(a_) {}

  HY_DEVICE void set_row(int row) {
    valid = row < a.nrows;
    if (!valid) return;
    const int grow = a.row_begin + row;
    b = grow / a.H;
    c = grow - b * a.H;
    vec_u = a.vec_u != 0;
    vec_o = a.vec_o != 0;
    vec_q = a.vec_q != 0;
    const long long uoff = (long long)b * a.u_bs;
    if (a.in_mode == HY_IN_SHORTCONV || a.out_mode == HY_OUT_SHORTCONV) {
      s0.init(a, b, c);
      s1.init(a, b, a.H + c);
      sv.init(a, b, 2 * a.H + c);
      pu = nullptr;
      ppre = nullptr;
    } else {
      pu = reinterpret_cast<const elem*>(a.u) + uoff + (long long)c * a.ldu;
      ppre = a.pre ? reinterpret_cast<const elem*>(a.pre) + uoff + (long long)c * a.ldu : nullptr;
    }
    const long long ooff = (long long)b * a.out_bs + (long long)c * a.ldo;
    pout = a.out ? reinterpret_cast<elem*>(a.out) + ooff : nullptr;
    pys = a.ysave ? reinterpret_cast<elem*>(a.ysave) + ooff : nullptr;
    pys_in = a.ysave_in ? reinterpret_cast<const elem*>(a.ysave_in) + (long long)b * a.ys_bs + (long long)c * a.ldys : nullptr;
    pdout = a.dout ? reinterpret_cast<const elem*>(a.dout) + ooff : nullptr;
    const long long qoff = (long long)b * a.post_bs + (long long)c * a.ldpost;
    pq = a.post ? reinterpret_cast<const elem*>(a.post) + qoff : nullptr;
    pdq = a.dpost ? reinterpret_cast<elem*>(a.dpost) + qoff : nullptr;
    if (a.in_mode == HY_IN_SHORTCONV) {
      pdu = a.du ? reinterpret_cast<elem*>(a.du) + uoff : nullptr;  // channel offset added at use
      pdpre = nullptr;
    } else {
      pdu = a.du ? reinterpret_cast<elem*>(a.du) + uoff + (long long)c * a.ldu : nullptr;
      pdpre = a.dpre ? reinterpret_cast<elem*>(a.dpre) + uoff + (long long)c * a.ldu : nullptr;
    }
  }

  // g at complex index n (reals 2n, 2n+1); zero beyond L
  HY_DEVICE float2 load_g(int n) const {
    const int t = 2 * n;
    if (!valid || t >= a.L) return make_float2(0.f, 0.f);
    float2 g;
    if (a.in_mode == HY_IN_SHORTCONV) {
      float2 x1 = s1.pair(t), v = sv.pair(t);
      g = make_float2(v.x * x1.x, v.y * x1.y);
      if (DT::kBf16) g = make_float2(round_to_bf16(g.x), round_to_bf16(g.y));
    } else if (a.in_mode == HY_IN_PREGATE) {
      float2 u = ld_pair_bounded<DT>(pu, t, a.L, vec_u), p = ld_pair_bounded<DT>(ppre, t, a.L, vec_u);
      g = make_float2(u.x * p.x, u.y * p.y);
      if (DT::kBf16) g = make_float2(round_to_bf16(g.x), round_to_bf16(g.y));
    } else {
      g = ld_pair_bounded<DT>(pu, t, a.L, vec_u);
    }
    if (t + 1 >= a.L) g.y = 0.f;
    return g;
  }

  // forward epilogue: y pair at complex index n
  HY_DEVICE void store_out(int n, float2 y) const {
    const int t = 2 * n;
    if (!valid || t >= a.L) return;
    if (a.out_mode == HY_OUT_SHORTCONV) {
      if (DT::kBf16) y = make_float2(round_to_bf16(y.x), round_to_bf16(y.y));
      if (pys) st_pair_bounded<DT>(pys, t, a.L, vec_o, y);
      float2 x0 = s0.pair(t);
      st_pair_bounded<DT>(pout, t, a.L, vec_o, make_float2(y.x * x0.x, y.y * x0.y));
    } else if (a.out_mode == HY_OUT_POSTGATE) {
      if (pys) st_pair_bounded<DT>(pys, t, a.L, vec_o, y);
      float2 q = ld_pair_bounded<DT>(pq, t, a.L, vec_q);
      st_pair_bounded<DT>(pout, t, a.L, vec_o, make_float2(y.x * q.x, y.y * q.y));
    } else {
      st_pair_bounded<DT>(pout, t, a.L, vec_o, y);
    }
  }

  // backward prologue: dy pair at n; also emits the gate gradient (dx0 / dq) and returns dy*g
  // contribution through `dot`.
  HY_DEVICE float2 load_dy(int n, float2 g, float& dot) const {
    const int t = 2 * n;
    if (!valid || t >= a.L) return make_float2(0.f, 0.f);
    float2 dz = ld_pair_bounded<DT>(pdout, t, a.L, vec_o);
    float2 dy = dz;
    if (a.out_mode == HY_OUT_SHORTCONV) {
      float2 x0 = s0.pair(t);
      dy = make_float2(dz.x * x0.x, dz.y * x0.y);
      float2 ys = ld_pair_bounded<DT>(pys_in, t, a.L, a.vec_y != 0);
      // dx0 = dz * y  -> dX channel group 0
      st_pair_bounded<DT>(pdu + (long long)c * a.ldu, t, a.L, vec_u, make_float2(dz.x * ys.x, dz.y * ys.y));
      if (DT::kBf16) dy = make_float2(round_to_bf16(dy.x), round_to_bf16(dy.y));
    } else if (a.out_mode == HY_OUT_POSTGATE) {
      float2 q = ld_pair_bounded<DT>(pq, t, a.L, vec_q);
      dy = make_float2(dz.x * q.x, dz.y * q.y);
      float2 ys = ld_pair_bounded<DT>(pys_in, t, a.L, a.vec_y != 0);
      st_pair_bounded<DT>(pdq, t, a.L, vec_q, make_float2(dz.x * ys.x, dz.y * ys.y));
    }
    if (t + 1 >= a.L) dy.y = 0.f;
    dot += dy.x * g.x + dy.y * g.y;
    return dy;
  }

  // backward epilogue: dg pair at n -> gradients of the pre-gate factors
  HY_DEVICE void store_dg(int n, float2 dg) const {
    const int t = 2 * n;
    if (!valid || t >= a.L) return;
    if (a.in_mode == HY_IN_SHORTCONV) {
      if (DT::kBf16) dg = make_float2(round_to_bf16(dg.x), round_to_bf16(dg.y));
      float2 x1 = s1.pair(t), v = sv.pair(t);
      st_pair_bounded<DT>(pdu + (long long)(a.H + c) * a.ldu, t, a.L, vec_u, make_float2(dg.x * v.x, dg.y * v.y));
      st_pair_bounded<DT>(pdu + (long long)(2 * a.H + c) * a.ldu, t, a.L, vec_u, make_float2(dg.x * x1.x, dg.y * x1.y));
    } else if (a.in_mode == HY_IN_PREGATE) {
      if (DT::kBf16) dg = make_float2(round_to_bf16(dg.x), round_to_bf16(dg.y));
      float2 u = ld_pair_bounded<DT>(pu, t, a.L, vec_u), p = ld_pair_bounded<DT>(ppre, t, a.L, vec_u);
      st_pair_bounded<DT>(pdu, t, a.L, vec_u, make_float2(dg.x * p.x, dg.y * p.y));
      st_pair_bounded<DT>(pdpre, t, a.L, vec_u, make_float2(dg.x * u.x, dg.y * u.y));
    } else {
      st_pair_bounded<DT>(pdu, t, a.L, vec_u, dg);
    }
  }
};

// ---- pointwise stage ---------------------------------------------------------------------------
enum { HY_PW_CONV = 0, HY_PW_CONVCONJ = 1, HY_PW_SPEC = 2, HY_PW_BWD = 3, HY_PW_REPACK = 4 };

// One (k, M-k) pair.  (za, zb) packed spectra of sequence 0 at the two positions, (ga, gb) of
// sequence 1 (backward only).  ka/kb index the spectrum arrays.
struct PairCtx {
  const float2* K;   // spectrum row base for this channel ([M] complex, position layout)
  float2* Kout;      // PW_SPEC
  float2* dK;        // PW_BWD: product destination (slot, channel)
  const float2* dKin;  // PW_REPACK: slot 0 base; slots strided by slot_stride
  long long slot_stride;
  int nslot;
  int accumulate;
  float scale;
  float skip;
};

template <int MODE>
HY_DEVICE void pair_op(const PairCtx& cx, long long ia, long long ib, float2 w, float2& za, float2& zb, float2 ga, float2 gb) {
  if (MODE == HY_PW_CONV || MODE == HY_PW_CONVCONJ) {
    float2 xa, xb;
    unpack_pair(za, zb, w, xa, xb);
    float2 ka = __ldg(cx.K + ia), kb = __ldg(cx.K + ib);
    float2 ya = (MODE == HY_PW_CONV) ? cmul(xa, ka) : cmulc(xa, ka);
    float2 yb = (MODE == HY_PW_CONV) ? cmul(xb, kb) : cmulc(xb, kb);
    repack_pair(ya, yb, w, za, zb);
  } else if (MODE == HY_PW_SPEC) {
    float2 xa, xb;
    unpack_pair(za, zb, w, xa, xb);
    cx.Kout[ia] = make_float2(xa.x * cx.scale + cx.skip, xa.y * cx.scale);
    cx.Kout[ib] = make_float2(xb.x * cx.scale + cx.skip, xb.y * cx.scale);
  } else if (MODE == HY_PW_BWD) {
    float2 xa, xb, ha, hb;
    unpack_pair(za, zb, w, xa, xb);   // DY
    unpack_pair(ga, gb, w, ha, hb);   // G
    float2 pa = cmulc(xa, ha), pb = cmulc(xb, hb);
    if (cx.accumulate) {
      float2 oa = cx.dK[ia], ob = cx.dK[ib];
      pa = cadd(pa, oa);
      pb = cadd(pb, ob);
    }
    cx.dK[ia] = pa;
    cx.dK[ib] = pb;
    float2 ka = __ldg(cx.K + ia), kb = __ldg(cx.K + ib);
    repack_pair(cmulc(xa, ka), cmulc(xb, kb), w, za, zb);
  } else {  // HY_PW_REPACK: true spectrum summed over slots -> packed
    float2 ya = make_float2(0.f, 0.f), yb = make_float2(0.f, 0.f);
    for (int s = 0; s < cx.nslot; ++s) {
      ya = cadd(ya, cx.dKin[s * cx.slot_stride + ia]);
      yb = cadd(yb, cx.dKin[s * cx.slot_stride + ib]);
    }
    repack_pair(cscale(ya, cx.scale), cscale(yb, cx.scale), w, za, zb);
  }
}

// The k = 0 slot carries (X[0], X[M]) (both real) as (x, y).
template <int MODE>
HY_DEVICE void dc_op(const PairCtx& cx, long long i0, float2& z0, float2 g0) {
  if (MODE == HY_PW_CONV || MODE == HY_PW_CONVCONJ) {
    float2 k0 = __ldg(cx.K + i0);
    float y0 = (z0.x + z0.y) * k0.x, ym = (z0.x - z0.y) * k0.y;
    z0 = make_float2(0.5f * (y0 + ym), 0.5f * (y0 - ym));
  } else if (MODE == HY_PW_SPEC) {
    cx.Kout[i0] = make_float2((z0.x + z0.y) * cx.scale + cx.skip, (z0.x - z0.y) * cx.scale + cx.skip);
  } else if (MODE == HY_PW_BWD) {
    float d0 = z0.x + z0.y, dm = z0.x - z0.y, h0 = g0.x + g0.y, hm = g0.x - g0.y;
    float2 p = make_float2(d0 * h0, dm * hm);
    if (cx.accumulate) p = cadd(p, cx.dK[i0]);
    cx.dK[i0] = p;
    float2 k0 = __ldg(cx.K + i0);
    float y0 = d0 * k0.x, ym = dm * k0.y;
    z0 = make_float2(0.5f * (y0 + ym), 0.5f * (y0 - ym));
  } else {
    float2 y = make_float2(0.f, 0.f);
    for (int s = 0; s < cx.nslot; ++s) y = cadd(y, cx.dKin[s * cx.slot_stride + i0]);
    y = cscale(y, cx.scale);
    z0 = make_float2(0.5f * (y.x + y.y), 0.5f * (y.x - y.y));
  }
}

// Row holding frequencies k = M1 * f (the k1 = 0 row; the only row when M1 == 1): f pairs with S - f.
// seq0 at sm0 (in/out), seq1 at sm1 (PW_BWD only).  rowoff = offset of this row inside the [M] spectrum.
template <int S, int MODE>
HY_DEVICE void pointwise_row0(float2* sm0, const float2* sm1, const PairCtx& cx, long long rowoff,
                              const float2* __restrict__ twpos, int tid, int nt) {
  using P = Plan<S>;
  constexpr int RL = P::radix(P::NS - 1);
  constexpr int HL = RL / 2;
  for (int i = tid; i < S / 2; i += nt) {
    const int p = (i / HL) * RL + (i % HL);
    const int f = freq_of_pos<S>(p);
    if (f == 0) {
      // DC / Nyquist slot at p = 0, and the self-paired f = S/2 at p = HL
      float2 z0 = (MODE == HY_PW_REPACK) ? make_float2(0.f, 0.f) : sm0[0];
      float2 g0 = (MODE == HY_PW_BWD) ? sm1[0] : make_float2(0.f, 0.f);
      dc_op<MODE>(cx, rowoff, z0, g0);
      if (MODE != HY_PW_SPEC) sm0[0] = z0;
      const int pm = HL + (HL >> 4);
      float2 zm = (MODE == HY_PW_REPACK) ? make_float2(0.f, 0.f) : sm0[pm];
      float2 gm = (MODE == HY_PW_BWD) ? sm1[pm] : make_float2(0.f, 0.f);
      float2 za = zm, zb = zm;
      pair_op<MODE>(cx, rowoff + HL, rowoff + HL, __ldg(twpos + HL), za, zb, gm, gm);
      if (MODE != HY_PW_SPEC) sm0[pm] = za;
    } else {
      const int pp = pos_of_freq<S>(S - f);
      const int ia = p + (p >> 4), ib = pp + (pp >> 4);
      float2 za = (MODE == HY_PW_REPACK) ? make_float2(0.f, 0.f) : sm0[ia];
      float2 zb = (MODE == HY_PW_REPACK) ? make_float2(0.f, 0.f) : sm0[ib];
      float2 ga = make_float2(0.f, 0.f), gb = ga;
      if (MODE == HY_PW_BWD) {
        ga = sm1[ia];
        gb = sm1[ib];
      }
      pair_op<MODE>(cx, rowoff + p, rowoff + pp, __ldg(twpos + p), za, zb, ga, gb);
      if (MODE != HY_PW_SPEC) {
        sm0[ia] = za;
        sm0[ib] = zb;
      }
    }
  }
}

// Rows k1 (A) and M1 - k1 (B), 1 <= k1 < M1/2: (A, p) pairs with (B, S-1-p).  cw = W_N^{k1}.
template <int S, int MODE>
HY_DEVICE void pointwise_rows(float2* smA, float2* smB, const float2* gA, const float2* gB, const PairCtx& cx,
                              long long offA, long long offB, float2 cw, const float2* __restrict__ twpos,
                              int tid, int nt) {
  for (int p = tid; p < S; p += nt) {
    const int pp = S - 1 - p;
    const int ia = p + (p >> 4), ib = pp + (pp >> 4);
    float2 za = (MODE == HY_PW_REPACK) ? make_float2(0.f, 0.f) : smA[ia];
    float2 zb = (MODE == HY_PW_REPACK) ? make_float2(0.f, 0.f) : smB[ib];
    float2 ga = make_float2(0.f, 0.f), gb = ga;
    if (MODE == HY_PW_BWD) {
      ga = gA[ia];
      gb = gB[ib];
    }
    float2 w = cmul(cw, __ldg(twpos + p));
    pair_op<MODE>(cx, offA + p, offB + pp, w, za, zb, ga, gb);
    if (MODE != HY_PW_SPEC) {
      smA[ia] = za;
      smB[ib] = zb;
    }
  }
}

// Row k1 = M1/2 (self-paired): p pairs with S-1-p inside the row.
template <int S, int MODE>
HY_DEVICE void pointwise_rowmid(float2* sm0, const float2* sm1, const PairCtx& cx, long long rowoff, float2 cw,
                                const float2* __restrict__ twpos, int tid, int nt) {
  for (int p = tid; p < S / 2; p += nt) {
    const int pp = S - 1 - p;
    const int ia = p + (p >> 4), ib = pp + (pp >> 4);
    float2 za = (MODE == HY_PW_REPACK) ? make_float2(0.f, 0.f) : sm0[ia];
    float2 zb = (MODE == HY_PW_REPACK) ? make_float2(0.f, 0.f) : sm0[ib];
    float2 ga = make_float2(0.f, 0.f), gb = ga;
    if (MODE == HY_PW_BWD) {
      ga = sm1[ia];
      gb = sm1[ib];
    }
    float2 w = cmul(cw, __ldg(twpos + p));
    pair_op<MODE>(cx, rowoff + p, rowoff + pp, w, za, zb, ga, gb);
    if (MODE != HY_PW_SPEC) {
      sm0[ia] = za;
      sm0[ib] = zb;
    }
  }
}

// =================================================================================================
//  Fused regime: M = S <= 4096, NB rows per CTA
// =================================================================================================
template <class DT, int S>
struct FusedLoadG {
  RowIO<DT>& io;
  int row0;
  HY_DEVICE FusedLoadG(RowIO<DT>& io_, int r0) : io(io_), row0(r0) {}
  HY_DEVICE void set_batch(int b) { io.set_row(row0 + b); }
  HY_DEVICE float2 ld(int e) const { return io.load_g(e); }
};
template <class DT, int S>
struct FusedStoreOut {
  RowIO<DT>& io;
  int row0;
  HY_DEVICE FusedStoreOut(RowIO<DT>& io_, int r0) : io(io_), row0(r0) {}
  HY_DEVICE void set_batch(int b) { io.set_row(row0 + b); }
  HY_DEVICE void st(int e, float2 v) const { io.store_out(e, v); }
};
template <class DT, int S>
struct FusedStoreDg {
  RowIO<DT>& io;
  int row0;
  HY_DEVICE FusedStoreDg(RowIO<DT>& io_, int r0) : io(io_), row0(r0) {}
  HY_DEVICE void set_batch(int b) { io.set_row(row0 + b); }
  HY_DEVICE void st(int e, float2 v) const { io.store_dg(e, v); }
};

template <int S>
HY_DEVICE PairCtx make_pair_ctx(const ConvArgs& a, int b, int c) {
  PairCtx cx;
  const long long M = (long long)a.M1 * S;
  cx.K = a.Kf ? a.Kf + (long long)c * M : nullptr;
  cx.Kout = a.Kf_out ? a.Kf_out + (long long)c * M : nullptr;
  cx.dK = a.dKacc ? a.dKacc + ((long long)(b - a.slot_b0) * a.H + c) * M : nullptr;
  cx.dKin = a.dKacc ? a.dKacc + (long long)c * M : nullptr;
  cx.slot_stride = (long long)a.H * M;
  cx.nslot = a.nslot;
  cx.accumulate = a.accumulate;
  cx.scale = a.scale;
  cx.skip = (a.skipD != nullptr) ? a.skipD[c] * a.scale : 0.f;
  return cx;
}

// forward (MODE = HY_PW_CONV) and filter spectrum (MODE = HY_PW_SPEC; DT = F32, rows = channels)
template <class DT, int S, int NB, int NT, int MODE>
__global__ void __launch_bounds__(NT) k_fused_fwd(ConvArgs a) {
  HY_DYN_SMEM(float2, sm);
  using P = Plan<S>;
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * NB;
  RowIO<DT> io(a);
  {
    FusedLoadG<DT, S> ld(io, row0);
    SmemRows<S> st(sm);
    fft_pass<S, NB, NT, 0, false, false, true, false>(a.tw, tid, ld, st);
  }
  __syncthreads();
  row_fwd_smem<S, NB, NT, 1, P::NS - 1>(sm, a.tw, tid);
  for (int r = 0; r < NB; ++r) {
    const int row = row0 + r;
    if (row < a.nrows) {
      const int grow = a.row_begin + row;
      PairCtx cx = make_pair_ctx<S>(a, grow / a.H, grow % a.H);
      pointwise_row0<S, MODE>(sm + r * RowSmem<S>::kRow, nullptr, cx, 0, a.twpos, tid, NT);
    }
  }
  if (MODE == HY_PW_SPEC) return;
  __syncthreads();
  row_inv_smem<S, NB, NT, P::NS - 1, 1>(sm, a.tw, tid);
  {
    SmemRows<S> ld(sm);
    FusedStoreOut<DT, S> st(io, row0);
    fft_pass<S, NB, NT, 0, true, false, false, true>(a.tw, tid, ld, st);
  }
}

// backward: sequences dy (seq 0) and g (seq 1)
template <class DT, int S, int NB, int NT>
__global__ void __launch_bounds__(NT) k_fused_bwd(ConvArgs a) {
  HY_DYN_SMEM(float2, sm);
  using P = Plan<S>;
  constexpr int R0 = P::radix(0);
  constexpr int SUB0 = S / R0;
  constexpr int NBF = S / R0;
  constexpr int TOTAL = NBF * NB;
  float2* sm_dy = sm;
  float2* sm_g = sm + NB * RowSmem<S>::kRow;
  float* part = reinterpret_cast<float*>(sm_g + NB * RowSmem<S>::kRow);  // [TOTAL]
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * NB;
  RowIO<DT> io(a);
  // pass 0 of both transforms, straight from global memory (upper half of the inputs is zero)
  for (int bid = tid; bid < TOTAL; bid += NT) {
    const int w = bid % NBF, batch = bid / NBF;
    io.set_row(row0 + batch);
    float2 xg[R0], xd[R0];
    float dot = 0.f;
#pragma unroll
    for (int m = 0; m < R0; ++m) {
      if (m >= R0 / 2) {
        xg[m] = make_float2(0.f, 0.f);
        xd[m] = make_float2(0.f, 0.f);
      } else {
        xg[m] = io.load_g(w + m * SUB0);
        xd[m] = io.load_dy(w + m * SUB0, xg[m], dot);
      }
    }
    part[bid] = dot;
    RegFFT<R0, false>::run(xg);
    RegFFT<R0, false>::run(xd);
    if (SUB0 > 1) {
      apply_twiddles<R0, false>(xg, a.tw, w * (HY_TWN / S));
      apply_twiddles<R0, false>(xd, a.tw, w * (HY_TWN / S));
    }
    const int off = batch * RowSmem<S>::kRow;
#pragma unroll
    for (int q = 0; q < R0; ++q) {
      const int e = w + q * SUB0;
      sm_g[off + e + (e >> 4)] = xg[q];
      sm_dy[off + e + (e >> 4)] = xd[q];
    }
  }
  __syncthreads();
  if (tid < NB && row0 + tid < a.nrows) {
    float s = 0.f;
    for (int i = 0; i < NBF; ++i) s += part[tid * NBF + i];
    a.dDpart[(long long)(a.row_begin + row0 + tid) * a.ndpart] = s;
  }
  row_fwd_smem<S, 2 * NB, NT, 1, P::NS - 1>(sm, a.tw, tid);  // dy rows then g rows are contiguous
  for (int r = 0; r < NB; ++r) {
    const int row = row0 + r;
    if (row < a.nrows) {
      const int grow = a.row_begin + row;
      PairCtx cx = make_pair_ctx<S>(a, grow / a.H, grow % a.H);
      pointwise_row0<S, HY_PW_BWD>(sm_dy + r * RowSmem<S>::kRow, sm_g + r * RowSmem<S>::kRow, cx, 0, a.twpos, tid, NT);
    }
  }
  __syncthreads();
  row_inv_smem<S, NB, NT, P::NS - 1, 1>(sm_dy, a.tw, tid);
  {
    SmemRows<S> ld(sm_dy);
    FusedStoreDg<DT, S> st(io, row0);
    fft_pass<S, NB, NT, 0, true, false, false, true>(a.tw, tid, ld, st);
  }
}

// dk finalize: dk[c][:L] = irfft(sum_slots dKacc)[ :L]  (DT = F32 rows = channels, OUT_PLAIN)
template <int S, int NB, int NT>
__global__ void __launch_bounds__(NT) k_fused_dk(ConvArgs a) {
  HY_DYN_SMEM(float2, sm);
  using P = Plan<S>;
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * NB;
  for (int r = 0; r < NB; ++r) {
    const int row = row0 + r;
    if (row < a.nrows) {
      PairCtx cx = make_pair_ctx<S>(a, a.slot_b0, a.row_begin + row);
      pointwise_row0<S, HY_PW_REPACK>(sm + r * RowSmem<S>::kRow, nullptr, cx, 0, a.twpos, tid, NT);
    }
  }
  __syncthreads();
  row_inv_smem<S, NB, NT, P::NS - 1, 1>(sm, a.tw, tid);
  RowIO<DT_F32> io(a);
  {
    SmemRows<S> ld(sm);
    FusedStoreOut<DT_F32, S> st(io, row0);
    fft_pass<S, NB, NT, 0, true, false, false, true>(a.tw, tid, ld, st);
  }
}

// =================================================================================================
//  Four-step regime: M = M1 x S
// =================================================================================================
// scratch layout: [row][seq][pos1][n2], row = (b - b0) * H + c, seq in [0, NSEQ)

template <int M1, int T2>
struct ColTile {
  float2* sm;
  int col;
  HY_DEVICE explicit ColTile(float2* s) : sm(s), col(0) {}
  HY_DEVICE void set_batch(int b) { col = b; }
  HY_DEVICE float2 ld(int e) const { return sm[e * T2 + col]; }
  HY_DEVICE void st(int e, float2 v) const { sm[e * T2 + col] = v; }
};

// big twiddle W_M^{n2 * k1}: U[pos1] (per CTA, shared) * V[pos1][l] (global table)
template <int M1, int T2>
HY_DEVICE void fill_U(float2* U, int n2_0, int M, int tid, int nt) {
  for (int i = tid; i < M1; i += nt) {
    const int k1 = freq_of_pos<M1>(i);
    const unsigned e = ((unsigned)n2_0 * (unsigned)k1) & (unsigned)(M - 1);
    float s, c;
    sincospif(2.0f * (float)e / (float)M, &s, &c);
    U[i] = make_float2(c, -s);
  }
}

// Phase A: NSEQ sequences per row (1: forward / spectrum, 2: backward dy + g).
template <class DT, int M1, int T2, int NT, int NSEQ>
__global__ void __launch_bounds__(NT) k_col_fwd(ConvArgs a) {
  HY_DYN_SMEM(float2, sm);
  using P = Plan<M1>;
  constexpr int NS = P::NS;
  const int S = a.S;
  const int M = M1 * S;
  float2* U = sm;                         // [M1]
  float2* tile = sm + M1;                 // [NSEQ][M1][T2] (only when NS > 1)
  float* part = reinterpret_cast<float*>(tile + (NS > 1 ? NSEQ * M1 * T2 : 0));  // [T2 * M1 / R0]
  const int tid = threadIdx.x;
  const int n2_0 = blockIdx.x * T2;
  const int row = blockIdx.y;
  fill_U<M1, T2>(U, n2_0, M, tid, NT);
  RowIO<DT> io(a);
  io.set_row(row);
  float2* out0 = a.scratch + ((long long)row * NSEQ) * M;
  constexpr int R0 = P::radix(0);
  constexpr int SUB0 = M1 / R0;
  constexpr int TOTAL0 = T2 * SUB0;
  __syncthreads();
  // pass 0 from global memory; n1 >= M1/2 is the zero padding
  for (int bid = tid; bid < TOTAL0; bid += NT) {
    const int col = bid % T2, w = bid / T2;
    float2 xg[R0], xd[R0];
    float dot = 0.f;
#pragma unroll
    for (int m = 0; m < R0; ++m) {
      const int n = (w + m * SUB0) * S + n2_0 + col;
      if (m >= R0 / 2 && R0 > 1) {
        xg[m] = make_float2(0.f, 0.f);
        if (NSEQ == 2) xd[m] = make_float2(0.f, 0.f);
      } else {
        xg[m] = io.load_g(n);
        if (NSEQ == 2) xd[m] = io.load_dy(n, xg[m], dot);
      }
    }
    if (NSEQ == 2) part[bid] = dot;
    RegFFT<R0, false>::run(xg);
    if (NSEQ == 2) RegFFT<R0, false>::run(xd);
    if (SUB0 > 1) {
      apply_twiddles<R0, false>(xg, a.tw, w * (HY_TWN / M1));
      if (NSEQ == 2) apply_twiddles<R0, false>(xd, a.tw, w * (HY_TWN / M1));
    }
#pragma unroll
    for (int q = 0; q < R0; ++q) {
      const int e = w + q * SUB0;  // pos1 when NS == 1
      if (NS == 1) {
        const float2 t = cmul(U[e], __ldg(a.twV + e * T2 + col));
        if (NSEQ == 2) {
          out0[(long long)e * S + n2_0 + col] = cmul(xd[q], t);
          out0[(long long)M + (long long)e * S + n2_0 + col] = cmul(xg[q], t);
        } else {
          out0[(long long)e * S + n2_0 + col] = cmul(xg[q], t);
        }
      } else {
        if (NSEQ == 2) {
          tile[e * T2 + col] = xd[q];
          tile[M1 * T2 + e * T2 + col] = xg[q];
        } else {
          tile[e * T2 + col] = xg[q];
        }
      }
    }
  }
  __syncthreads();
  if (NSEQ == 2 && tid == 0) {
    float s = 0.f;
    for (int i = 0; i < TOTAL0; ++i) s += part[i];
    a.dDpart[(long long)(a.row_begin + row) * a.ndpart + blockIdx.x] = s;
  }
  if constexpr (NS > 1) {
    for (int q = 0; q < NSEQ; ++q) {
      float2* t = tile + q * M1 * T2;
      // middle passes
      if constexpr (NS > 2) {
        ColTile<M1, T2> acc(t);
        fft_pass<M1, T2, NT, 1, false, true, false, false>(a.tw, tid, acc, acc);
      }
    }
    if constexpr (NS > 2) __syncthreads();
    static_assert(NS <= 3, "column transforms use at most 3 passes");
    // last pass: tile -> twiddle -> scratch
    for (int q = 0; q < NSEQ; ++q) {
      float2* t = tile + q * M1 * T2;
      float2* dst = out0 + (long long)q * M;
      struct Sink {
        float2* dst; const float2* U; const float2* V; int n2_0, col, S;
        HY_DEVICE void set_batch(int b) { col = b; }
        HY_DEVICE void st(int e, float2 v) const {
          const float2 t = cmul(U[e], __ldg(V + e * T2 + col));
          dst[(long long)e * S + n2_0 + col] = cmul(v, t);
        }
      } sink{dst, U, a.twV, n2_0, 0, S};
      ColTile<M1, T2> src(t);
      fft_pass<M1, T2, NT, NS - 1, false, true, false, false>(a.tw, tid, src, sink);
    }
  }
}

// Phase B: one CTA per pair of rows (k1, M1-k1) [CTA 0: rows k1 = 0 and k1 = M1/2] of one signal row.
template <int S, int NT, int MODE>
__global__ void __launch_bounds__(NT) k_row_conv(ConvArgs a) {
  HY_DYN_SMEM(float2, sm);
  using P = Plan<S>;
  constexpr int NSEQ = (MODE == HY_PW_BWD) ? 2 : 1;
  const int M1 = a.M1;
  const long long M = (long long)M1 * S;
  constexpr int NB = 2 * NSEQ;
  const int tid = threadIdx.x;
  const int pr = blockIdx.x;   // pair index
  const int row = blockIdx.y;
  const int kA = (pr == 0) ? 0 : pr;
  const int kB = (pr == 0) ? (M1 / 2) : (M1 - pr);
  const int pA = pos_of_freq_rt(M1, kA), pB = pos_of_freq_rt(M1, kB);
  float2* base = a.scratch + (long long)row * NSEQ * M;
  const int grow = a.row_begin + row;
  const int b = grow / a.H, c = grow % a.H;
  // smem rows: [seq0 A, seq0 B, seq1 A, seq1 B]
  struct Src {
    const float2* base; int pA, pB; long long M; const float2* p;
    HY_DEVICE void set_batch(int bb) { p = base + (long long)(bb >> 1) * M + (long long)((bb & 1) ? pB : pA) * S; }
    HY_DEVICE float2 ld(int e) const { return p[e]; }
  } src{base, pA, pB, M, nullptr};
  if (MODE != HY_PW_REPACK) {
    SmemRows<S> st(sm);
    fft_pass<S, NB, NT, 0, false, false, false, false>(a.tw, tid, src, st);
    __syncthreads();
    row_fwd_smem<S, NB, NT, 1, P::NS - 1>(sm, a.tw, tid);
  }
  PairCtx cx = make_pair_ctx<S>(a, (MODE == HY_PW_REPACK) ? a.slot_b0 : b, c);
  float2* s0A = sm;
  float2* s0B = sm + RowSmem<S>::kRow;
  float2* s1A = sm + 2 * RowSmem<S>::kRow;
  float2* s1B = sm + 3 * RowSmem<S>::kRow;
  const float N = 2.0f * (float)M;
  if (pr == 0) {
    pointwise_row0<S, MODE>(s0A, s1A, cx, (long long)pA * S, a.twpos, tid, NT);
    float sn, cs;
    sincospif(2.0f * (float)kB / N, &sn, &cs);
    if (M1 > 1) pointwise_rowmid<S, MODE>(s0B, s1B, cx, (long long)pB * S, make_float2(cs, -sn), a.twpos, tid, NT);
  } else {
    float sn, cs;
    sincospif(2.0f * (float)kA / N, &sn, &cs);
    pointwise_rows<S, MODE>(s0A, s0B, s1A, s1B, cx, (long long)pA * S, (long long)pB * S, make_float2(cs, -sn), a.twpos, tid, NT);
  }
  if (MODE == HY_PW_SPEC) return;
  __syncthreads();
  row_inv_smem<S, 2, NT, P::NS - 1, 1>(sm, a.tw, tid);
  struct Dst {
    float2* base; int pA, pB; float2* p;
    HY_DEVICE void set_batch(int bb) { p = base + (long long)((bb & 1) ? pB : pA) * S; }
    HY_DEVICE void st(int e, float2 v) const { p[e] = v; }
  } dst{base, pA, pB, nullptr};
  SmemRows<S> ld(sm);
  fft_pass<S, 2, NT, 0, true, false, false, false>(a.tw, tid, ld, dst);
}

// Phase C: inverse column transforms + epilogue.  EPI: 0 forward output, 1 backward dg.
template <class DT, int M1, int T2, int NT, int NSEQ, int EPI>
__global__ void __launch_bounds__(NT) k_col_inv(ConvArgs a) {
  HY_DYN_SMEM(float2, sm);
  using P = Plan<M1>;
  constexpr int NS = P::NS;
  const int S = a.S;
  const int M = M1 * S;
  float2* U = sm;
  float2* tile = sm + M1;
  const int tid = threadIdx.x;
  const int n2_0 = blockIdx.x * T2;
  const int row = blockIdx.y;
  fill_U<M1, T2>(U, n2_0, M, tid, NT);
  RowIO<DT> io(a);
  io.set_row(row);
  const float2* src0 = a.scratch + ((long long)row * NSEQ) * M;
  __syncthreads();
  struct Src {
    const float2* src; const float2* U; const float2* V; int n2_0, col, S;
    HY_DEVICE void set_batch(int b) { col = b; }
    HY_DEVICE float2 ld(int e) const {
      const float2 t = cmul(U[e], __ldg(V + e * T2 + col));
      return cmulc(src[(long long)e * S + n2_0 + col], t);
    }
  } src{src0, U, a.twV, n2_0, 0, S};
  struct Epi {
    const RowIO<DT>& io; int n2_0, col, S;
    HY_DEVICE void set_batch(int b) { col = b; }
    HY_DEVICE void st(int e, float2 v) const {
      const int n = e * S + n2_0 + col;
      if (EPI == 0) io.store_out(n, v);
      else io.store_dg(n, v);
    }
  } epi{io, n2_0, 0, S};
  if constexpr (NS == 1) {
    fft_pass<M1, T2, NT, 0, true, true, false, true>(a.tw, tid, src, epi);
  } else {
    ColTile<M1, T2> t(tile);
    fft_pass<M1, T2, NT, NS - 1, true, true, false, false>(a.tw, tid, src, t);
    __syncthreads();
    if constexpr (NS > 2) {
      fft_pass<M1, T2, NT, 1, true, true, false, false>(a.tw, tid, t, t);
      __syncthreads();
    }
    fft_pass<M1, T2, NT, 0, true, true, false, true>(a.tw, tid, t, epi);
  }
}
