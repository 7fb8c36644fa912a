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
//     column transforms + epilogue).  Two schedules: three launches per group of rows with the complex
//     scratch in HBM (groups sized by hy_set_scratch_budget, default 1 GB), or — hy_conv_pipe.cuh — ONE
//     persistent launch that deals A tiles of row r+1, B row pairs of row r and C tiles of row r-1 from
//     a single work queue over a ring of a few row buffers, so the A->B->C hand-off stays in L2.
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
  float2* gsave;       // four-step only: [B*H][M] packed row-transform of g; the forward writes it, the backward
                       // reads it INSTEAD of transforming g again (nullable: recompute)
  float* dDpart;       // backward: [B*H][ndpart] partial sums of dy*g
  float2* scratch;     // four-step: [rows][nseq][M] complex
  const float2* tw;    // W_8192 table
  const float2* twpos; // [S] W_{2S}^{freq(p)}
  const float2* twV;   // four-step: [M1][T2] W_M^{l*k1(pos1)}
  int B, H, L, M1;
  int S;               // four-step: row length (M = M1 * S)
  int row_begin, nrows;  // global (b*H + c) row range handled by this launch; scratch is indexed by local row
  int slot_b0;         // backward: dKacc slot of batch b is (b - slot_b0)
  int vec_all;         // every activation pointer / stride allows aligned 2-element vector access
  int stage_ok;        // SHORTCONV source rows are bf16, 16-byte aligned with stride % 8 == 0: cp.async staging allowed
  int stage_dz_ok;     // same property for the dout rows of the backward
  int vec8_out;        // the rows the column epilogue writes (out / ysave, or du) allow aligned 16-byte stores
  int vec16_in;        // fp32 PLAIN input rows allow aligned 16-byte loads (cp.async straight into the transform tile)
  int defer_dx0;       // backward, SHORTCONV: do not form dx0 = dout * y here (hy_shortconv_bwd_gate does): no ysave read
  int in_mode, out_mode;
  int accumulate;      // backward: dKacc += instead of =
  int nslot;           // dk finalize: number of slots to sum
  int ndpart;          // row stride of dDpart
  float scale;         // spectrum / dk scaling (1/M)
};

// ---- branch-free pair access -----------------------------------------------------------------------
// Every global access of the prologues/epilogues goes through these: the address is clamped into the
// row and the value masked afterwards, so an unrolled loop of them is ONE basic block and all its
// loads are in flight together (the first cut was long_scoreboard-bound at 4 loads in flight).
template <class DT, bool VEC>
struct PairRow {
  typedef typename DT::elem elem;
  const elem* p;
  int L, Lc;   // Lc = last even index < L
  HY_DEVICE void init(const elem* base, int L_) {
    p = base;
    L = L_;
    Lc = (L_ - 1) & ~1;
  }
  // elements (t, t+1) of the row for even t (any value; clamped). Entries outside [0, L) are garbage: mask!
  HY_DEVICE float2 ld(int t) const {
    const int tc = t < 0 ? 0 : (t > Lc ? Lc : t);
    if (VEC) return ld2<DT>(p + tc, true);
    const float x = ld1<DT>(p + tc);
    const float y = ld1<DT>(p + (tc + 1 < L ? tc + 1 : tc));
    return make_float2(x, y);
  }
  HY_DEVICE void st(elem* q, int t, float2 v) const {   // t even, t < L
    if (VEC) {
      st2<DT>(q + t, v, true);   // in vec mode the row stride is even and >= L+1 for odd L: slot t+1 exists
    } else {
      st1<DT>(q + t, v.x);
      if (t + 1 < L) st1<DT>(q + t + 1, v.y);
    }
  }
};

// Source row of the short filter: global memory, or (STG) a tile of it staged into shared memory by
// cp.async at kernel start — rows [n1][rs] holding elements 2*(n1*S + n2_0) - 8 .. + rs of the global row.
// OPT: the tile may be absent at run time (sp == nullptr -> global memory); costs a branch per access, so only the
// rows that need it ask for it.
template <class DT, bool VEC, bool STG, bool OPT = false>
struct ConvSrcRow {
  typedef typename DT::elem elem;
  PairRow<DT, VEC> g;
  const elem* sp;
  int lgS, n2_0, rs, tile_elems, L;
  HY_DEVICE void init(const elem* base, int L_) {
    g.init(base, L_);
    L = L_;
    sp = nullptr;
  }
  HY_DEVICE void attach(const elem* sp_, int lgS_, int n2_0_, int rs_, int tile_elems_) {
    sp = sp_; lgS = lgS_; n2_0 = n2_0_; rs = rs_; tile_elems = tile_elems_;
  }
  HY_DEVICE float2 ld(int t) const {
    if (!STG) return g.ld(t);
    if (OPT && sp == nullptr) return g.ld(t);
    const int tc = t < 0 ? 0 : (t > g.Lc ? g.Lc : t);
    // row n1 of the tile covers elements [2(n1 S + n2_0) - 8, 2(n1 S + n2_0 + T2)): the 8-element left margin
    // holds the causal halo, which for the tile's first column belongs to the previous n1's samples
    const int q = tc - 2 * n2_0 + 8;
    int idx = (q >> (lgS + 1)) * rs + (q & ((2 << lgS) - 1));
    idx = idx < 0 ? 0 : (idx > tile_elems - 2 ? tile_elems - 2 : idx);
    return ld2<DT>(sp + idx, true);
  }
};

// cooperative cp.async staging of one source-row tile (16-byte chunks; bf16 rows, 16-byte aligned, stride % 8 == 0)
template <class DT>
HY_DEVICE void stage_tile(typename DT::elem* sdst, const typename DT::elem* grow, int nrows, int rs, int S, int n2_0,
                          int ldrow, int tid, int nt) {
  const int cpr = rs / 8;
  for (int i = tid; i < nrows * cpr; i += nt) {
    const int n1 = i / cpr, ch = i - n1 * cpr;
    const long long t0 = 2LL * ((long long)n1 * S + n2_0) - 8 + 8 * ch;
    typename DT::elem* d = sdst + n1 * rs + 8 * ch;
    if (t0 >= 0 && t0 + 8 <= ldrow) hy_cp_async16(d, grow + t0);
    else *reinterpret_cast<uint4*>(d) = make_uint4(0u, 0u, 0u, 0u);
  }
}

// ---- short depthwise causal conv (hyena.py:407-413,444): out[t] = b + w0 x[t-2] + w1 x[t-1] + w2 x[t]
template <class DT, bool VEC, bool STG = false>
struct ShortConvRow {
  ConvSrcRow<DT, VEC, STG> row;
  float w0, w1, w2, bias, pb;
  bool has_pb;
  HY_DEVICE void init(const ConvArgs& a, int b, int ch) {
    row.init(reinterpret_cast<const typename DT::elem*>(a.u) + (long long)b * a.u_bs + (long long)ch * a.ldu, a.L);
    w0 = a.sw[ch * 3 + 0];
    w1 = a.sw[ch * 3 + 1];
    w2 = a.sw[ch * 3 + 2];
    bias = a.sb[ch];
    has_pb = a.pb != nullptr;
    pb = has_pb ? a.pb[ch] : 0.f;
  }
  HY_DEVICE float2 fix2(float2 r) const {
    if (has_pb) {
      r = make_float2(r.x + pb, r.y + pb);
      if (DT::kBf16) r = round2_to_bf16(r);
    }
    return r;
  }
  // conv outputs at (t, t+1) from the raw pairs at t-2 and t (t even, t < L)
  HY_DEVICE float2 conv(int t, float2 prev, float2 cur) const {
    const float2 fp = fix2(prev), fc = fix2(cur);
    const float xm2 = t >= 2 ? fp.x : 0.f;
    const float xm1 = t >= 2 ? fp.y : 0.f;
    const float x0 = fc.x;
    const float x1 = (t + 1 < row.L) ? fc.y : 0.f;
    float2 o = make_float2(fmaf(w2, x0, fmaf(w1, xm1, fmaf(w0, xm2, bias))), fmaf(w2, x1, fmaf(w1, x0, fmaf(w0, xm1, bias))));
    if (DT::kBf16) o = round2_to_bf16(o);
    return o;
  }
};


// ---- vectorised gate sweeps over staged bf16 tiles (four-step column kernels, interior chunks) -------------------
// The per-butterfly prologue / epilogue above resolves every sample on its own (clamps, masks, index algebra: ~50
// instructions per complex point).  When the short-filter source rows sit in shared memory as [n1][RS] bf16 tiles,
// a chunk of 8 consecutive samples that lies entirely inside [2, L) needs none of that: one 16-byte + one 4-byte
// shared load per source row, the 3-tap filter on ten values, and 16-byte accesses on the other side.
HY_DEVICE void unpack8_bf16(uint4 q, float (&r)[8]) {
  r[0] = __uint_as_float(q.x << 16); r[1] = __uint_as_float(q.x & 0xffff0000u);
  r[2] = __uint_as_float(q.y << 16); r[3] = __uint_as_float(q.y & 0xffff0000u);
  r[4] = __uint_as_float(q.z << 16); r[5] = __uint_as_float(q.z & 0xffff0000u);
  r[6] = __uint_as_float(q.w << 16); r[7] = __uint_as_float(q.w & 0xffff0000u);
}
HY_DEVICE uint4 pack8_bf16(const float (&r)[8]) {
  return make_uint4(pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]), pack_bf16x2(r[4], r[5]), pack_bf16x2(r[6], r[7]));
}
HY_DEVICE void round8_bf16(float (&r)[8]) {
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    const float2 t = round2_to_bf16(make_float2(r[i], r[i + 1]));
    r[i] = t.x;
    r[i + 1] = t.y;
  }
}
// eight short-filter outputs at samples t0 .. t0+7 (same arithmetic, same order as ShortConvRow::conv); p points at
// sample t0 of the staged row (16-byte aligned; the two samples before it are the causal halo)
template <class SC>
HY_DEVICE void conv8_staged(const SC& s, const unsigned short* p, float (&o)[8]) {
  const unsigned h = *reinterpret_cast<const unsigned*>(p - 2);
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  float r[10];
  r[0] = __uint_as_float(h << 16);
  r[1] = __uint_as_float(h & 0xffff0000u);
  float r8[8];
  unpack8_bf16(q, r8);
#pragma unroll
  for (int i = 0; i < 8; ++i) r[2 + i] = r8[i];
  if (s.has_pb) {
#pragma unroll
    for (int i = 0; i < 10; i += 2) {
      const float2 t = round2_to_bf16(make_float2(r[i] + s.pb, r[i + 1] + s.pb));
      r[i] = t.x;
      r[i + 1] = t.y;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = fmaf(s.w2, r[i + 2], fmaf(s.w1, r[i + 1], fmaf(s.w0, r[i], s.bias)));
  round8_bf16(o);
}

// raw operands of one packed sample: two source rows (x1 & v / u & pre), previous and current pair
struct GIn {
  float2 a_prev, a_cur, b_prev, b_cur;
};
// raw operands of the backward prologue
struct DIn {
  float2 c_prev, c_cur, dz, ys;
};

// ---- per-row signal access --------------------------------------------------------------------
// Produces g (and, for the backward, dy) as packed complex samples z[n] = (s[2n], s[2n+1]) and
// consumes results.  One instance per thread; set_row() is called once per row.
template <class DT, bool VEC, bool STG = false>
struct RowIO {
  typedef typename DT::elem elem;
  const ConvArgs& a;
  bool valid;
  int b, c;
  ShortConvRow<DT, VEC, STG> s0, s1, sv;   // x0, x1, v rows (SHORTCONV); STG: read from the staged tiles
  PairRow<DT, VEC> ru, rpre, rq, rys;
  ConvSrcRow<DT, VEC, STG, true> rdout;          // backward: dout row (STG: may be staged like the short-filter sources)
  elem* pout;
  elem* pys;
  elem *pdu, *pdpre, *pdq;

  HY_DEVICE explicit RowIO(const ConvArgs& a_) : a(a_) {}

  HY_DEVICE void set_row(int row) {
    valid = row < a.nrows;
    if (!valid) row = 0;            // keep every pointer valid: loads stay branch-free, results are masked
    const int grow = a.row_begin + row;
    b = grow / a.H;
    c = grow - b * a.H;
    const long long uoff = (long long)b * a.u_bs;
    if (a.in_mode == HY_IN_SHORTCONV) {
      s0.init(a, b, c);
      s1.init(a, b, a.H + c);
      sv.init(a, b, 2 * a.H + c);
    } else {
      ru.init(reinterpret_cast<const elem*>(a.u) + uoff + (long long)c * a.ldu, a.L);
      rpre.init(a.pre ? reinterpret_cast<const elem*>(a.pre) + uoff + (long long)c * a.ldu : ru.p, a.L);
    }
    const long long ooff = (long long)b * a.out_bs + (long long)c * a.ldo;
    pout = a.out ? reinterpret_cast<elem*>(a.out) + ooff : nullptr;
    pys = a.ysave ? reinterpret_cast<elem*>(a.ysave) + ooff : nullptr;
    rdout.init(a.dout ? reinterpret_cast<const elem*>(a.dout) + ooff : nullptr, a.L);
    if (a.ysave_in) rys.init(reinterpret_cast<const elem*>(a.ysave_in) + (long long)b * a.ys_bs + (long long)c * a.ldys, a.L);
    const long long qoff = (long long)b * a.post_bs + (long long)c * a.ldpost;
    if (a.post) rq.init(reinterpret_cast<const elem*>(a.post) + qoff, a.L);
    pdq = a.dpost ? reinterpret_cast<elem*>(a.dpost) + qoff : nullptr;
    if (a.in_mode == HY_IN_SHORTCONV) {
      pdu = a.du ? reinterpret_cast<elem*>(a.du) + uoff : nullptr;  // channel offset added at use
      pdpre = nullptr;
    } else {
      pdu = a.du ? reinterpret_cast<elem*>(a.du) + uoff + (long long)c * a.ldu : nullptr;
      pdpre = a.dpre ? reinterpret_cast<elem*>(a.dpre) + uoff + (long long)c * a.ldu : nullptr;
    }
  }

  // ---- g = (gated) input at complex index n: fetch (loads only) + make (arithmetic only) ----------
  HY_DEVICE GIn fetch_g(int n) const {
    const int t = 2 * n;
    GIn r;
    if (a.in_mode == HY_IN_SHORTCONV) {
      r.a_prev = s1.row.ld(t - 2);
      r.a_cur = s1.row.ld(t);
      r.b_prev = sv.row.ld(t - 2);
      r.b_cur = sv.row.ld(t);
    } else {
      r.a_cur = ru.ld(t);
      r.b_cur = rpre.ld(t);
      r.a_prev = r.a_cur;
      r.b_prev = r.b_cur;
    }
    return r;
  }
  HY_DEVICE float2 make_g(int n, const GIn& r) const {
    const int t = 2 * n;
    float2 g;
    if (a.in_mode == HY_IN_SHORTCONV) {
      const float2 x1 = s1.conv(t, r.a_prev, r.a_cur), v = sv.conv(t, r.b_prev, r.b_cur);
      g = make_float2(v.x * x1.x, v.y * x1.y);
      if (DT::kBf16) g = round2_to_bf16(g);
    } else if (a.in_mode == HY_IN_PREGATE) {
      g = make_float2(r.a_cur.x * r.b_cur.x, r.a_cur.y * r.b_cur.y);
      if (DT::kBf16) g = round2_to_bf16(g);
    } else {
      g = r.a_cur;
    }
    if (!valid || t >= a.L) g.x = 0.f;
    if (!valid || t + 1 >= a.L) g.y = 0.f;
    return g;
  }
  HY_DEVICE float2 load_g(int n) const { return make_g(n, fetch_g(n)); }

  // ---- forward epilogue: gate operands, then out = y * gate ---------------------------------------
  HY_DEVICE GIn fetch_gate(int n) const {
    const int t = 2 * n;
    GIn r;
    r.a_prev = r.a_cur = r.b_prev = r.b_cur = make_float2(0.f, 0.f);
    if (a.out_mode == HY_OUT_SHORTCONV) {
      r.a_prev = s0.row.ld(t - 2);
      r.a_cur = s0.row.ld(t);
    } else if (a.out_mode == HY_OUT_POSTGATE) {
      r.a_cur = rq.ld(t);
    }
    return r;
  }
  HY_DEVICE void store_out(int n, float2 y, const GIn& r) const {
    const int t = 2 * n;
    if (!valid || t >= a.L) return;
    PairRow<DT, VEC> w;
    w.init(nullptr, a.L);
    if (a.out_mode == HY_OUT_SHORTCONV) {
      if (DT::kBf16) y = round2_to_bf16(y);
      if (pys) w.st(pys, t, y);
      const float2 x0 = s0.conv(t, r.a_prev, r.a_cur);
      w.st(pout, t, make_float2(y.x * x0.x, y.y * x0.y));
    } else if (a.out_mode == HY_OUT_POSTGATE) {
      if (pys) w.st(pys, t, y);
      w.st(pout, t, make_float2(y.x * r.a_cur.x, y.y * r.a_cur.y));
    } else {
      w.st(pout, t, y);
    }
  }
  HY_DEVICE void store_out(int n, float2 y) const { store_out(n, y, fetch_gate(n)); }

  // ---- backward prologue: dy = dout * gate (also emits the gate gradient) -------------------------
  HY_DEVICE DIn fetch_dy(int n) const {
    const int t = 2 * n;
    DIn r;
    r.dz = rdout.ld(t);
    r.c_prev = r.c_cur = r.ys = r.dz;
    if (a.out_mode == HY_OUT_SHORTCONV) {
      r.c_prev = s0.row.ld(t - 2);
      r.c_cur = s0.row.ld(t);
      if (!a.defer_dx0) r.ys = rys.ld(t);
    } else if (a.out_mode == HY_OUT_POSTGATE) {
      r.c_cur = rq.ld(t);
      r.ys = rys.ld(t);
    }
    return r;
  }
  HY_DEVICE float2 make_dy(int n, const DIn& r, float2 g, float& dot) const {
    const int t = 2 * n;
    if (!valid || t >= a.L) return make_float2(0.f, 0.f);
    PairRow<DT, VEC> w;
    w.init(nullptr, a.L);
    float2 dy = r.dz;
    if (a.out_mode == HY_OUT_SHORTCONV) {
      const float2 x0 = s0.conv(t, r.c_prev, r.c_cur);
      dy = make_float2(r.dz.x * x0.x, r.dz.y * x0.y);
      if (!a.defer_dx0) w.st(pdu + (long long)c * a.ldu, t, make_float2(r.dz.x * r.ys.x, r.dz.y * r.ys.y));   // dx0 = dz * y
      if (DT::kBf16) dy = round2_to_bf16(dy);
    } else if (a.out_mode == HY_OUT_POSTGATE) {
      dy = make_float2(r.dz.x * r.c_cur.x, r.dz.y * r.c_cur.y);
      w.st(pdq, t, make_float2(r.dz.x * r.ys.x, r.dz.y * r.ys.y));                           // dq = dout * y
    }
    if (t + 1 >= a.L) dy.y = 0.f;
    dot += dy.x * g.x + dy.y * g.y;
    return dy;
  }

  // ---- backward epilogue: dg -> gradients of the pre-gate factors ---------------------------------
  HY_DEVICE void store_dg(int n, float2 dg, const GIn& r) const {
    const int t = 2 * n;
    if (!valid || t >= a.L) return;
    PairRow<DT, VEC> w;
    w.init(nullptr, a.L);
    if (a.in_mode == HY_IN_SHORTCONV) {
      if (DT::kBf16) dg = round2_to_bf16(dg);
      const float2 x1 = s1.conv(t, r.a_prev, r.a_cur), v = sv.conv(t, r.b_prev, r.b_cur);
      w.st(pdu + (long long)(a.H + c) * a.ldu, t, make_float2(dg.x * v.x, dg.y * v.y));
      w.st(pdu + (long long)(2 * a.H + c) * a.ldu, t, make_float2(dg.x * x1.x, dg.y * x1.y));
    } else if (a.in_mode == HY_IN_PREGATE) {
      if (DT::kBf16) dg = round2_to_bf16(dg);
      w.st(pdu, t, make_float2(dg.x * r.b_cur.x, dg.y * r.b_cur.y));
      w.st(pdpre, t, make_float2(dg.x * r.a_cur.x, dg.y * r.a_cur.y));
    } else {
      w.st(pdu, t, dg);
    }
  }
};

// ---- pointwise stage ---------------------------------------------------------------------------
enum { HY_PW_CONV = 0, HY_PW_CONVCONJ = 1, HY_PW_SPEC = 2, HY_PW_BWD = 3, HY_PW_REPACK = 4,
       HY_PW_BWDG = 5 /* backward with the packed spectrum of g read from ConvArgs::gsave */ };

// One (k, M-k) pair.  (za, zb) packed spectra of sequence 0 at the two positions, (ga, gb) of
// sequence 1 (backward only).  ka/kb index the spectrum arrays.
struct PairCtx {
  const float2* K;   // spectrum row base for this channel ([M] complex, position layout)
  float2* Kout;      // PW_SPEC
  float2* dK;        // PW_BWD: product destination (slot, channel)
  const float2* dKin;  // PW_REPACK: slot 0 base; slots strided by slot_stride
  float2* Gs;          // PW_CONV: where to save the packed spectrum of g (nullable); PW_BWDG: where to read it
  long long slot_stride;
  int nslot;
  int accumulate;
  float scale;
  float skip;
};

// operands a pair needs from global memory, fetched ahead of the arithmetic (batched by the callers)
struct PairK {
  float2 ka, kb, tw;
  float2 ga, gb;   // PW_BWDG: saved packed spectrum of g at the two positions
};
template <int MODE>
HY_DEVICE PairK pair_fetch(const PairCtx& cx, long long ia, long long ib, const float2* __restrict__ twpos, int p) {
  PairK r;
  r.tw = __ldg(twpos + p);
  r.ga = r.gb = make_float2(0.f, 0.f);
  if (MODE == HY_PW_CONV || MODE == HY_PW_CONVCONJ || MODE == HY_PW_BWD || MODE == HY_PW_BWDG) {
    r.ka = __ldg(cx.K + ia);
    r.kb = __ldg(cx.K + ib);
    if (MODE == HY_PW_BWDG) {
      r.ga = __ldg(cx.Gs + ia);
      r.gb = __ldg(cx.Gs + ib);
    }
  } else if (MODE == HY_PW_REPACK) {
    // slot sums are part of the batched fetch: every load of a batch of pairs is in flight together
    // slot 0 unconditionally (the loads of a batch of pairs then sit in one basic block), further slots in a loop
    float2 ya = __ldg(cx.dKin + ia), yb = __ldg(cx.dKin + ib);
    for (int s = 1; s < cx.nslot; ++s) {
      ya = cadd(ya, __ldg(cx.dKin + s * cx.slot_stride + ia));
      yb = cadd(yb, __ldg(cx.dKin + s * cx.slot_stride + ib));
    }
    r.ka = ya;
    r.kb = yb;
  } else {
    r.ka = r.kb = make_float2(0.f, 0.f);
  }
  return r;
}

template <int MODE>
HY_DEVICE void pair_op(const PairCtx& cx, long long ia, long long ib, float2 w, float2& za, float2& zb, float2 ga, float2 gb,
                       float2 ka, float2 kb) {
  if (MODE == HY_PW_CONV || MODE == HY_PW_CONVCONJ) {
    if (MODE == HY_PW_CONV && cx.Gs != nullptr) {
      cx.Gs[ia] = za;
      cx.Gs[ib] = zb;
    }
    float2 xa, xb;
    unpack_pair(za, zb, w, xa, xb);
    float2 ya = (MODE == HY_PW_CONV) ? cmul(xa, ka) : cmulc(xa, ka);
    float2 yb = (MODE == HY_PW_CONV) ? cmul(xb, kb) : cmulc(xb, kb);
    repack_pair(ya, yb, w, za, zb);
  } else if (MODE == HY_PW_SPEC) {
    float2 xa, xb;
    unpack_pair(za, zb, w, xa, xb);
    cx.Kout[ia] = make_float2(xa.x * cx.scale + cx.skip, xa.y * cx.scale);
    cx.Kout[ib] = make_float2(xb.x * cx.scale + cx.skip, xb.y * cx.scale);
  } else if (MODE == HY_PW_BWD || MODE == HY_PW_BWDG) {
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
    repack_pair(cmulc(xa, ka), cmulc(xb, kb), w, za, zb);
  } else {  // HY_PW_REPACK: true spectrum summed over slots (by pair_fetch) -> packed
    repack_pair(cscale(ka, cx.scale), cscale(kb, cx.scale), w, za, zb);
  }
}

// The k = 0 slot carries (X[0], X[M]) (both real) as (x, y).
template <int MODE>
HY_DEVICE void dc_op(const PairCtx& cx, long long i0, float2& z0, float2 g0) {
  if (MODE == HY_PW_CONV || MODE == HY_PW_CONVCONJ) {
    if (MODE == HY_PW_CONV && cx.Gs != nullptr) cx.Gs[i0] = z0;
    float2 k0 = __ldg(cx.K + i0);
    float y0 = (z0.x + z0.y) * k0.x, ym = (z0.x - z0.y) * k0.y;
    z0 = make_float2(0.5f * (y0 + ym), 0.5f * (y0 - ym));
  } else if (MODE == HY_PW_SPEC) {
    cx.Kout[i0] = make_float2((z0.x + z0.y) * cx.scale + cx.skip, (z0.x - z0.y) * cx.scale + cx.skip);
  } else if (MODE == HY_PW_BWD || MODE == HY_PW_BWDG) {
    if (MODE == HY_PW_BWDG) g0 = __ldg(cx.Gs + i0);
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
      const PairK kk = pair_fetch<MODE>(cx, rowoff + HL, rowoff + HL, twpos, HL);
      if (MODE == HY_PW_BWDG) gm = kk.ga;
      pair_op<MODE>(cx, rowoff + HL, rowoff + HL, kk.tw, za, zb, gm, gm, kk.ka, kk.kb);
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
      const PairK kk = pair_fetch<MODE>(cx, rowoff + p, rowoff + pp, twpos, p);
      if (MODE == HY_PW_BWDG) {
        ga = kk.ga;
        gb = kk.gb;
      }
      pair_op<MODE>(cx, rowoff + p, rowoff + pp, kk.tw, za, zb, ga, gb, kk.ka, kk.kb);
      if (MODE != HY_PW_SPEC) {
        sm0[ia] = za;
        sm0[ib] = zb;
      }
    }
  }
}

// Rows k1 (A) and M1 - k1 (B), 1 <= k1 < M1/2: (A, p) pairs with (B, S-1-p).  cw = W_N^{k1}.
// Pairs are processed four at a time: the spectrum / twiddle loads of a batch are issued together.
template <int S, int MODE>
HY_DEVICE void pointwise_rows(float2* smA, float2* smB, const float2* gA, const float2* gB, const PairCtx& cx,
                              long long offA, long long offB, float2 cw, const float2* __restrict__ twpos,
                              int tid, int nt) {
  // the dk repack has nothing before this stage to hide its loads behind: twice the pairs in flight per batch
  constexpr int U = (MODE == HY_PW_REPACK) ? 8 : 4;
  for (int p0 = tid; p0 < S; p0 += U * nt) {
    PairK kk[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u * nt;
      const int pc = p < S ? p : 0;
      kk[u] = pair_fetch<MODE>(cx, offA + pc, offB + (S - 1 - pc), twpos, pc);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u * nt;
      if (p < S) {
        const int pp = S - 1 - p;
        const int ia = p + (p >> 4), ib = pp + (pp >> 4);
        float2 za = (MODE == HY_PW_REPACK) ? make_float2(0.f, 0.f) : smA[ia];
        float2 zb = (MODE == HY_PW_REPACK) ? make_float2(0.f, 0.f) : smB[ib];
        float2 ga = make_float2(0.f, 0.f), gb = ga;
        if (MODE == HY_PW_BWD) {
          ga = gA[ia];
          gb = gB[ib];
        } else if (MODE == HY_PW_BWDG) {
          ga = kk[u].ga;
          gb = kk[u].gb;
        }
        pair_op<MODE>(cx, offA + p, offB + pp, cmul(cw, kk[u].tw), za, zb, ga, gb, kk[u].ka, kk[u].kb);
        if (MODE != HY_PW_SPEC) {
          smA[ia] = za;
          smB[ib] = zb;
        }
      }
    }
  }
}

// Row k1 = M1/2 (self-paired): p pairs with S-1-p inside the row.
template <int S, int MODE>
HY_DEVICE void pointwise_rowmid(float2* sm0, const float2* sm1, const PairCtx& cx, long long rowoff, float2 cw,
                                const float2* __restrict__ twpos, int tid, int nt) {
  constexpr int U = 4;
  for (int p0 = tid; p0 < S / 2; p0 += U * nt) {
    PairK kk[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u * nt;
      const int pc = p < S / 2 ? p : 0;
      kk[u] = pair_fetch<MODE>(cx, rowoff + pc, rowoff + (S - 1 - pc), twpos, pc);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u * nt;
      if (p < S / 2) {
        const int pp = S - 1 - p;
        const int ia = p + (p >> 4), ib = pp + (pp >> 4);
        float2 za = (MODE == HY_PW_REPACK) ? make_float2(0.f, 0.f) : sm0[ia];
        float2 zb = (MODE == HY_PW_REPACK) ? make_float2(0.f, 0.f) : sm0[ib];
        float2 ga = make_float2(0.f, 0.f), gb = ga;
        if (MODE == HY_PW_BWD) {
          ga = sm1[ia];
          gb = sm1[ib];
        } else if (MODE == HY_PW_BWDG) {
          ga = kk[u].ga;
          gb = kk[u].gb;
        }
        pair_op<MODE>(cx, rowoff + p, rowoff + pp, cmul(cw, kk[u].tw), za, zb, ga, gb, kk[u].ka, kk[u].kb);
        if (MODE != HY_PW_SPEC) {
          sm0[ia] = za;
          sm0[ib] = zb;
        }
      }
    }
  }
}

// =================================================================================================
//  Fused regime: M = S <= 4096, NB rows per CTA
// =================================================================================================
template <int S>
HY_DEVICE PairCtx make_pair_ctx(const ConvArgs& a, int b, int c) {
  PairCtx cx;
  const long long M = (long long)a.M1 * S;
  cx.K = a.Kf ? a.Kf + (long long)c * M : nullptr;
  cx.Kout = a.Kf_out ? a.Kf_out + (long long)c * M : nullptr;
  cx.dK = a.dKacc ? a.dKacc + ((long long)(b - a.slot_b0) * a.H + c) * M : nullptr;
  cx.dKin = a.dKacc ? a.dKacc + (long long)c * M : nullptr;
  cx.Gs = a.gsave ? a.gsave + ((long long)b * a.H + c) * M : nullptr;
  cx.slot_stride = (long long)a.H * M;
  cx.nslot = a.nslot;
  cx.accumulate = a.accumulate;
  cx.scale = a.scale;
  cx.skip = (a.skipD != nullptr) ? a.skipD[c] * a.scale : 0.f;
  return cx;
}

// deterministic block-wide sum of per-thread partials grouped by `groups` consecutive-thread groups
HY_DEVICE float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <class DT, bool VEC, int MAXR>
struct LoadG2P {   // two-phase loader of g for pass 0
  enum { kAffine = 0 };
  RowIO<DT, VEC>& io;
  int row0;
  GIn raw[MAXR];
  HY_DEVICE LoadG2P(RowIO<DT, VEC>& io_, int r0) : io(io_), row0(r0) {}
  HY_DEVICE void set_batch(int b) { io.set_row(row0 + b); }
  HY_DEVICE void fetch(int m, int e) { raw[m] = io.fetch_g(e); }
  HY_DEVICE float2 get(int m, int e) const { return io.make_g(e, raw[m]); }
  HY_DEVICE float2 ld(int e) const { return io.load_g(e); }
};
template <class DT, bool VEC, int MAXR>
struct LoadDY2P {   // two-phase loader of dy = dout * gate for pass 0 (backward with the saved spectrum of g)
  enum { kAffine = 0 };
  RowIO<DT, VEC>& io;
  int row0;
  DIn raw[MAXR];
  HY_DEVICE LoadDY2P(RowIO<DT, VEC>& io_, int r0) : io(io_), row0(r0) {}
  HY_DEVICE void set_batch(int b) { io.set_row(row0 + b); }
  HY_DEVICE void fetch(int m, int e) { raw[m] = io.fetch_dy(e); }
  HY_DEVICE float2 get(int m, int e) const {
    float dot = 0.f;
    return io.make_dy(e, raw[m], make_float2(0.f, 0.f), dot);
  }
  HY_DEVICE float2 ld(int e) const {
    float dot = 0.f;
    return io.make_dy(e, io.fetch_dy(e), make_float2(0.f, 0.f), dot);
  }
};
template <class DT, bool VEC, int MAXR, int EPI>
struct StoreEpi2P {   // two-phase sink: EPI 0 forward output, EPI 1 backward dg
  enum { kAffine = 0 };
  RowIO<DT, VEC>& io;
  int row0;
  GIn raw[4];
  HY_DEVICE StoreEpi2P(RowIO<DT, VEC>& io_, int r0) : io(io_), row0(r0) {}
  HY_DEVICE void set_batch(int b) { io.set_row(row0 + b); }
  HY_DEVICE void prefetch(int m, int e) { raw[m] = (EPI == 0) ? io.fetch_gate(e) : io.fetch_g(e); }
  HY_DEVICE void st_pref(int m, int e, float2 v) const {
    if (EPI == 0) io.store_out(e, v, raw[m]);
    else io.store_dg(e, v, raw[m]);
  }
};

// forward (MODE = HY_PW_CONV) and filter spectrum (MODE = HY_PW_SPEC; DT = F32, rows = channels)
// resident CTAs per SM the single-kernel regime is compiled for.  The forward is bound by exposed load latency
// (long_scoreboard 4.1 stalls per issue at 2 CTAs per SM, profiles/r02g_ncu_fused_regime_*): more resident CTAs win
// despite the spills of the register cap — at L = 4096, D = 256, B = 128: 1.47 ms at 2 CTAs (128 registers), 1.15 at 3
// (80), 1.05 at 4 (64).  The backward's 79 KB of shared memory admit 2 CTAs only: a register cap there just spills.
#ifndef HY_FUSED_BWD_CH
#define HY_FUSED_BWD_CH 4
#endif
#ifndef HY_FUSED_BWD_NT
#define HY_FUSED_BWD_NT 256
#endif
#ifndef HY_FUSED_FWD_MINB
#define HY_FUSED_FWD_MINB 4
#endif
template <class DT, int S, int NB, int NT, int MODE>
__global__ void __launch_bounds__(NT, HY_FUSED_FWD_MINB) k_fused_fwd(ConvArgs a) {
  if (a.vec_all) {
    constexpr bool VEC = true;
#include "hy_conv_fusedfwd_body.inc"
  } else {
    constexpr bool VEC = false;
#include "hy_conv_fusedfwd_body.inc"
  }
}

// backward: sequences dy (seq 0) and g (seq 1)
template <class DT, int S, int NB, int NT>
__global__ void __launch_bounds__(NT, (NT <= 256 ? 2 : 2)) k_fused_bwd(ConvArgs a) {
  if (a.vec_all) {
    constexpr bool VEC = true;
#include "hy_conv_fusedbwd_body.inc"
  } else {
    constexpr bool VEC = false;
#include "hy_conv_fusedbwd_body.inc"
  }
}

// backward with the forward's saved spectrum of g (ConvArgs::gsave): only dy is transformed, ONE row buffer per sequence —
// 43.5 KB of shared memory instead of 79, so four CTAs share an SM like the forward's (the two-sequence kernel above is
// bound by exposed load latency at its two: long_scoreboard 6.5-7.3 stalls per issue).  dD is dk[:, 0] (host).
#ifndef HY_FUSED_BWDG_MINB
#define HY_FUSED_BWDG_MINB 4
#endif
template <class DT, int S, int NB, int NT>
__global__ void __launch_bounds__(NT, HY_FUSED_BWDG_MINB) k_fused_bwdg(ConvArgs a) {
  if (a.vec_all) {
    constexpr bool VEC = true;
#include "hy_conv_fusedbwdg_body.inc"
  } else {
    constexpr bool VEC = false;
#include "hy_conv_fusedbwdg_body.inc"
  }
}

// dk finalize: dk[c][:L] = irfft(sum_slots dKacc)[ :L]  (DT = F32 rows = channels, OUT_PLAIN)
template <int S, int NB, int NT>
__global__ void __launch_bounds__(NT, 2) k_fused_dk(ConvArgs a) {
  using P = Plan<S>;
  HY_DYN_SMEM(float4, smem4);
  float4* twt = smem4;
  float2* sm = reinterpret_cast<float2*>(smem4 + P::tw_slots());
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * NB;
  build_tw_smem<S>(twt, a.tw, tid, NT);
  TwSmem<S> tw{twt};
  for (int r = 0; r < NB; ++r) {
    const int row = row0 + r;
    if (row < a.nrows) {
      PairCtx cx = make_pair_ctx<S>(a, a.slot_b0, a.row_begin + row);
      pointwise_row0<S, HY_PW_REPACK>(sm + r * RowSmem<S>::kRow, nullptr, cx, 0, a.twpos, tid, NT);
    }
  }
  __syncthreads();
  row_inv_smem<S, NB, NT, P::NS - 1, 1>(sm, tw, tid);
  RowIO<DT_F32, false> io(a);
  {
    SmemRows<S> ld(sm);
    StoreEpi2P<DT_F32, false, P::radix(0), 0> st(io, row0);
    fft_pass<S, NB, NT, 0, true, false, false, true, false, true>(tw, tid, ld, st);
  }
}

// =================================================================================================
//  Four-step regime: M = M1 x S
// =================================================================================================
// scratch layout: [row][seq][pos1][n2], row = local row of the launch, seq in [0, NSEQ)

template <int M1, int T2>
struct ColTile {
  enum { kAffine = 1 };
  float2* sm;
  int col;
  HY_DEVICE explicit ColTile(float2* s) : sm(s), col(0) {}
  HY_DEVICE void set_batch(int b) { col = b; }
  HY_DEVICE float2 ld(int e) const { return sm[e * T2 + col]; }
  HY_DEVICE void st(int e, float2 v) const { sm[e * T2 + col] = v; }
  HY_DEVICE int pbase(int base) const { return base * T2 + col; }
  HY_DEVICE float2 ldp(int pb, int K) const { return sm[pb + K * T2]; }
  HY_DEVICE void stp(int pb, int K, float2 v) const { sm[pb + K * T2] = v; }
};

// big twiddle W_M^{n2 * k1}: U[pos1] (per CTA, shared) * V[pos1][l] (global table)
template <int M1, int T2>
HY_DEVICE void fill_U(float2* U, int n2_0, int M, int tid, int nt) {
  for (int i = tid; i < M1; i += nt) {
    const int k1 = freq_of_pos<M1>(i);
    const unsigned e = ((unsigned)n2_0 * (unsigned)k1) % (unsigned)M;
    float s, c;
    sincospif(2.0f * (float)e / (float)M, &s, &c);
    U[i] = make_float2(c, -s);
  }
}

// shared-memory carve-up of the column kernels: [tw table][U (M1)][tile(s)][part]
template <int M1>
struct ColSmem {
  using P = Plan<M1>;
  static constexpr int kTw4 = P::tw_slots();                 // float4
  static constexpr int kHead = kTw4 * 2 + M1;                // float2 slots before the tiles
};

// Phase A: NSEQ sequences per row (1: forward / spectrum, 2: backward dy + g).  DYO (NSEQ == 1): the one sequence
// is dy = dout * gate (the backward when the spectrum of g was saved by the forward); dx0 / dq are emitted as usual.
// (bx, row) = (column tile, local row) — blockIdx for the per-phase launches, a work-queue item for the persistent
// pipeline (hy_conv_pipe.cuh); out0 = this row's scratch.
template <class DT, int M1, int T2, int NT, int NSEQ, bool VEC, bool STG = false, bool DYO = false>
HY_DEVICE void col_fwd_body(const ConvArgs& a, const int bx, const int row, float2* const out0) {
#include "hy_conv_colfwd_body.inc"
}
#ifndef HY_COL_MINB
#define HY_COL_MINB 2
#endif
// the fp32 phase A (filter spectrum) is the one column kernel that gains from 3 resident CTAs (spectrum 1.67 -> 1.55 ms
// per layer at 1 M; the bf16 and inverse kernels lose: measured with -DHY_COL_MINB=3)
#ifndef HY_COLF32_FWD_MINB
#define HY_COLF32_FWD_MINB 3
#endif
template <class DT, int M1, int T2, int NT, int NSEQ, bool DYO = false>
__global__ void __launch_bounds__(NT, (NT <= 128 ? 4 : (NT <= 256 ? (DT::kBf16 ? HY_COL_MINB : HY_COLF32_FWD_MINB) : 1))) k_col_fwd(ConvArgs a) {
  const int bx = blockIdx.x, row = blockIdx.y;
  float2* const out0 = a.scratch + ((long long)row * NSEQ) * ((long long)M1 * a.S);
  if constexpr (DT::kBf16 && NSEQ == 1) {
    if (a.stage_ok) {
      constexpr bool VEC = true, STG = true;
#include "hy_conv_colfwd_body.inc"
      return;
    }
  }
  if (a.vec_all) {
    constexpr bool VEC = true, STG = false;
#include "hy_conv_colfwd_body.inc"
  } else {
    constexpr bool VEC = false, STG = false;
#include "hy_conv_colfwd_body.inc"
  }
}

// Phase B: one CTA per pair of rows (k1, M1-k1) [CTA 0: rows k1 = 0 and k1 = M1/2] of one signal row.
// Phase B: one CTA per pair of scratch rows (k1, M1 - k1).  The body lives in hy_conv_row_body.inc (see there why).
// XCTA: the scratch was written by other CTAs of the SAME launch (persistent pipeline): read it at L2 (ld.global.cg).
template <int S, int NT, int MODE, bool XCTA = false>
HY_DEVICE void row_conv_body(const ConvArgs& a, const int pr, const int row, float2* const base) {
#include "hy_conv_row_body.inc"
}
template <int S, int NT, int MODE>
__global__ void __launch_bounds__(NT, (NT <= 256 ? 2 : 1)) k_row_conv(ConvArgs a) {
  constexpr bool XCTA = false;
  const int pr = blockIdx.x;   // pair index
  const int row = blockIdx.y;
  float2* const base = a.scratch + (long long)row * ((MODE == HY_PW_BWD) ? 2 : 1) * ((long long)a.M1 * S);
#include "hy_conv_row_body.inc"
}

// Phase C: inverse column transforms + epilogue.  EPI: 0 forward output, 1 backward dg.
template <class DT, int M1, int T2, int NT, int NSEQ, int EPI, bool VEC, bool STG = false, bool XCTA = false>
HY_DEVICE void col_inv_body(const ConvArgs& a, const int bx, const int row, const float2* const src0) {
#include "hy_conv_colinv_body.inc"
}
template <class DT, int M1, int T2, int NT, int NSEQ, int EPI>
__global__ void __launch_bounds__(NT, (NT <= 128 ? 4 : HY_COL_MINB)) k_col_inv(ConvArgs a) {
  const int bx = blockIdx.x, row = blockIdx.y;
  const float2* const src0 = a.scratch + ((long long)row * NSEQ) * ((long long)M1 * a.S);
  constexpr bool XCTA = false;
  if constexpr (DT::kBf16) {
    if (a.stage_ok) {
      constexpr bool VEC = true, STG = true;
#include "hy_conv_colinv_body.inc"
      return;
    }
  }
  if (a.vec_all) {
    constexpr bool VEC = true, STG = false;
#include "hy_conv_colinv_body.inc"
  } else {
    constexpr bool VEC = false, STG = false;
#include "hy_conv_colinv_body.inc"
  }
}
