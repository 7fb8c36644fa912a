// hyena-b200: FFT building blocks (fp32, CUDA cores).
//
// Design: every transform is an in-place decimation-in-frequency pass sequence
// (natural order in -> digit-reversed order out) whose inverse is the mirrored decimation-in-time
// sequence (digit-reversed in -> natural out).  Convolution only needs a consistent order in the
// frequency domain, so no bit-reversal pass is ever executed.  A pass = one radix-R butterfly per
// thread held entirely in registers (R in {2,4,8,16}); data moves between passes through shared
// memory (row transforms) or stays in a [M1][T2] column tile (column transforms of the four-step
// split).  Twiddles come from one 8192-entry table W_8192^i built in double precision at init.
#pragma once
#include "hy_common.cuh"

#define HY_TWN 8192  // twiddle table length: tw[i] = exp(-2*pi*i * i / 8192)

// ---- pass plan -------------------------------------------------------------------------------
// S = 2^LG * ODD with ODD in {1, 3, 5}: the power-of-two part is spread evenly over ceil(LG/4) passes (e.g. 4096 ->
// 16,16,16; 2048 -> 16,16,8; 512 -> 8,8,8; 32 -> 8,4) and the odd factor, when present, is the LAST forward pass
// (sub = 1: no twiddles after it; 40 -> 8,5; 320 -> 8,8,5; 48 -> 16,3).  Only column lengths of the four-step split
// carry an odd factor (transform lengths 3 * 2^k and 5 * 2^k: the reference's rfft(n = 2L) pads to exactly 2L,
// hyena.py:61-62, so L = 160 000 should not pay for 2^18 points); row lengths stay powers of two.
// column lengths M1 of the four-step split that have kernel instances (X-macro lists)
#define HY_COLS_POW2(X) X(2) X(4) X(8) X(16) X(32) X(64) X(128) X(256) X(512)
#define HY_COLS_ODD5(X) X(10) X(20) X(40) X(80) X(160) X(320)
#define HY_COLS_ODD3(X) X(12) X(24) X(48) X(96) X(192) X(384)
#define HY_COLS_ODD(X) HY_COLS_ODD5(X) HY_COLS_ODD3(X)
template <int S>
struct Plan {
  static constexpr int ODD = (S % 5 == 0) ? 5 : ((S % 3 == 0) ? 3 : 1);
  static constexpr int P2 = S / ODD;
  static_assert((P2 & (P2 - 1)) == 0, "transform length must be 2^a, 3 * 2^a or 5 * 2^a");
  static constexpr int LG = hy_ilog2(P2);
  static constexpr int NS2 = (LG + 3) / 4;          // power-of-two passes
  static constexpr int NS = NS2 + (ODD > 1 ? 1 : 0);
  HY_HD static constexpr int bits(int i) { return (NS2 == 0 || i >= NS2) ? 0 : (LG / NS2 + (i < LG % NS2 ? 1 : 0)); }
  HY_HD static constexpr int radix(int i) { return i < NS2 ? (1 << bits(i)) : ODD; }
  HY_HD static constexpr int shift_before(int i) {   // power-of-two plans only
    int s = 0;
    for (int m = 0; m < i; ++m) s += bits(m);
    return s;
  }
  // product of the radices of the passes before pass i = weight of pass i's digit in the frequency index
  HY_HD static constexpr int weight(int i) {
    int w = 1;
    for (int m = 0; m < i; ++m) w *= radix(m);
    return w;
  }
  HY_HD static constexpr int span(int i) { return S / weight(i); }
  HY_HD static constexpr int sub(int i) { return span(i) / radix(i); }
  HY_HD static constexpr int lgsub(int i) { return hy_ilog2(sub(i)); }   // power-of-two plans only
  // per-pass twiddle table (shared memory): passes with sub > 1 own 2*sub float4 slots
  HY_HD static constexpr int tw_off(int i) {
    int o = 0;
    for (int m = 0; m < i; ++m) o += (sub(m) > 1 ? 2 * sub(m) : 0);
    return o;
  }
  HY_HD static constexpr int tw_slots() { return tw_off(NS); }   // float4 count
};

// position p (after the forward passes) -> frequency index stored there, and back: pass i contributes the digit
// (p / sub(i)) % radix(i) of the position with weight(i) in the frequency
template <int S>
HY_DEVICE int freq_of_pos(int p) {
  using P = Plan<S>;
  int k = 0;
  if constexpr (P::ODD == 1) {
#pragma unroll
    for (int i = 0; i < P::NS; ++i) k |= ((p >> P::lgsub(i)) & (P::radix(i) - 1)) << P::shift_before(i);
  } else {
    const unsigned up = (unsigned)p;
#pragma unroll
    for (int i = 0; i < P::NS; ++i) k += (int)((up / (unsigned)P::sub(i)) % (unsigned)P::radix(i)) * P::weight(i);
  }
  return k;
}
template <int S>
HY_DEVICE int pos_of_freq(int k) {
  using P = Plan<S>;
  int p = 0;
  if constexpr (P::ODD == 1) {
#pragma unroll
    for (int i = 0; i < P::NS; ++i) p |= ((k >> P::shift_before(i)) & (P::radix(i) - 1)) << P::lgsub(i);
  } else {
    const unsigned uk = (unsigned)k;
#pragma unroll
    for (int i = 0; i < P::NS; ++i) p += (int)((uk / (unsigned)P::weight(i)) % (unsigned)P::radix(i)) * P::sub(i);
  }
  return p;
}

// runtime-length variant (same plan as Plan<S>), used where S is not a template parameter
HY_DEVICE int pos_of_freq_rt(int S, int k) {
  const int odd = (S % 5 == 0) ? 5 : ((S % 3 == 0) ? 3 : 1);
  const int p2 = S / odd;
  int lg = 0;
  while ((1 << lg) < p2) ++lg;
  const int ns = (lg + 3) / 4;
  int p = 0, w = 1;
  if (ns > 0) {
    const int base = lg / ns, extra = lg % ns;
    for (int i = 0; i < ns; ++i) {
      const int r = 1 << (base + (i < extra ? 1 : 0));
      p += ((k / w) % r) * (S / (w * r));
      w *= r;
    }
  }
  if (odd > 1) p += (k / w) % odd;
  return p;
}

// ---- register butterflies (natural order in, natural order out) ------------------------------
template <bool INV>
HY_DEVICE void fft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), d = csub(a1, a3);
  a0 = cadd(t0, t2);
  a2 = csub(t0, t2);
  // t1 +- (quarter turn of d): scalar on purpose — the turn is a free operand swap here, while a packed add would
  // need the swapped pair materialised in registers first
  if (INV) {   // +i d = (-d.y, d.x)
    a1 = make_float2(t1.x - d.y, t1.y + d.x);
    a3 = make_float2(t1.x + d.y, t1.y - d.x);
  } else {     // -i d = (d.y, -d.x)
    a1 = make_float2(t1.x + d.y, t1.y - d.x);
    a3 = make_float2(t1.x - d.y, t1.y + d.x);
  }
}

template <int R, bool INV>
struct RegFFT;

template <bool INV>
struct RegFFT<1, INV> {
  static HY_DEVICE void run(float2 (&)[1]) {}
};
template <bool INV>
struct RegFFT<2, INV> {
  static HY_DEVICE void run(float2 (&x)[2]) {
    float2 a = x[0], b = x[1];
    x[0] = cadd(a, b);
    x[1] = csub(a, b);
  }
};
// odd radices (last forward pass / first inverse pass of the 3 * 2^k and 5 * 2^k column plans)
template <bool INV>
struct RegFFT<3, INV> {
  static HY_DEVICE void run(float2 (&x)[3]) {
    const float s = 0.86602540378443864676f;
    const float2 t = cadd(x[1], x[2]), d = csub(x[1], x[2]);
    const float2 m = make_float2(x[0].x - 0.5f * t.x, x[0].y - 0.5f * t.y);
    const float2 n = make_float2(s * d.x, s * d.y);
    x[0] = cadd(x[0], t);
    if (!INV) {   // y1 = m - i n, y2 = m + i n
      x[1] = make_float2(m.x + n.y, m.y - n.x);
      x[2] = make_float2(m.x - n.y, m.y + n.x);
    } else {
      x[1] = make_float2(m.x - n.y, m.y + n.x);
      x[2] = make_float2(m.x + n.y, m.y - n.x);
    }
  }
};
template <bool INV>
struct RegFFT<5, INV> {
  static HY_DEVICE void run(float2 (&x)[5]) {
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;   // cos(2 pi/5), cos(4 pi/5)
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;    // sin(2 pi/5), sin(4 pi/5)
    const float2 t1 = cadd(x[1], x[4]), t2 = cadd(x[2], x[3]), t3 = csub(x[1], x[4]), t4 = csub(x[2], x[3]);
    const float2 m1 = make_float2(fmaf(c2, t2.x, fmaf(c1, t1.x, x[0].x)), fmaf(c2, t2.y, fmaf(c1, t1.y, x[0].y)));
    const float2 m2 = make_float2(fmaf(c1, t2.x, fmaf(c2, t1.x, x[0].x)), fmaf(c1, t2.y, fmaf(c2, t1.y, x[0].y)));
    const float2 n1 = make_float2(fmaf(s2, t4.x, s1 * t3.x), fmaf(s2, t4.y, s1 * t3.y));
    const float2 n2 = make_float2(fmaf(-s1, t4.x, s2 * t3.x), fmaf(-s1, t4.y, s2 * t3.y));
    x[0] = cadd(x[0], cadd(t1, t2));
    if (!INV) {   // y1 = m1 - i n1, y4 = m1 + i n1, y2 = m2 - i n2, y3 = m2 + i n2
      x[1] = make_float2(m1.x + n1.y, m1.y - n1.x);
      x[4] = make_float2(m1.x - n1.y, m1.y + n1.x);
      x[2] = make_float2(m2.x + n2.y, m2.y - n2.x);
      x[3] = make_float2(m2.x - n2.y, m2.y + n2.x);
    } else {
      x[1] = make_float2(m1.x - n1.y, m1.y + n1.x);
      x[4] = make_float2(m1.x + n1.y, m1.y - n1.x);
      x[2] = make_float2(m2.x - n2.y, m2.y + n2.x);
      x[3] = make_float2(m2.x + n2.y, m2.y - n2.x);
    }
  }
};
template <bool INV>
struct RegFFT<4, INV> {
  static HY_DEVICE void run(float2 (&x)[4]) { fft4<INV>(x[0], x[1], x[2], x[3]); }
};
template <bool INV>
struct RegFFT<8, INV> {
  static HY_DEVICE void run(float2 (&x)[8]) {
    const float h = 0.70710678118654752440f;
    fft4<INV>(x[0], x[2], x[4], x[6]);  // E[k1] -> x[2*k1]
    fft4<INV>(x[1], x[3], x[5], x[7]);  // O[k1] -> x[2*k1+1]
    float2 o1 = x[3], o3 = x[7];
    if (!INV) {
      x[3] = make_float2((o1.x + o1.y) * h, (o1.y - o1.x) * h);
      x[7] = make_float2((o3.y - o3.x) * h, (-o3.x - o3.y) * h);
    } else {
      x[3] = make_float2((o1.x - o1.y) * h, (o1.x + o1.y) * h);
      x[7] = make_float2((-o3.x - o3.y) * h, (o3.x - o3.y) * h);
    }
    x[5] = crot<INV>(x[5]);
    float2 r[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      r[k] = cadd(x[2 * k], x[2 * k + 1]);
      r[k + 4] = csub(x[2 * k], x[2 * k + 1]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = r[k];
  }
};
template <bool INV>
struct RegFFT<16, INV> {
  static HY_DEVICE void run(float2 (&x)[16]) {
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;
    const float h = 0.70710678118654752440f;
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) fft4<INV>(x[n2], x[4 + n2], x[8 + n2], x[12 + n2]);
    // x[4*k1 + n2] *= W16^(n2*k1)
    x[5] = ctw<INV>(x[5], c1, s1);     // e=1
    x[6] = ctw<INV>(x[6], h, h);       // e=2
    x[7] = ctw<INV>(x[7], s1, c1);     // e=3
    x[9] = ctw<INV>(x[9], h, h);       // e=2
    x[10] = crot<INV>(x[10]);          // e=4
    x[11] = ctw<INV>(x[11], -h, h);    // e=6
    x[13] = ctw<INV>(x[13], s1, c1);   // e=3
    x[14] = ctw<INV>(x[14], -h, h);    // e=6
    x[15] = ctw<INV>(x[15], -c1, -s1); // e=9
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) fft4<INV>(x[4 * k1], x[4 * k1 + 1], x[4 * k1 + 2], x[4 * k1 + 3]);
    float2 r[16];
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
      for (int k2 = 0; k2 < 4; ++k2) r[k1 + 4 * k2] = x[4 * k1 + k2];
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = r[k];
  }
};

// ---- per-pass twiddles -------------------------------------------------------------------------
// x[q] *= W^q for q = 1..R-1 given the table values W^1, W^2, W^4, W^8 (each rounded once from double);
// the remaining powers are products of at most three of them (<= 3.5 ulp).
template <int R, bool INV>
HY_DEVICE void apply_twiddles(float2 (&x)[R], float2 w1, float2 w2, float2 w4, float2 w8) {
  if constexpr (R == 3) {
    x[1] = cmul_dir<INV>(x[1], w1);
    x[2] = cmul_dir<INV>(x[2], w2);
  } else if constexpr (R == 5) {
    x[1] = cmul_dir<INV>(x[1], w1);
    x[2] = cmul_dir<INV>(x[2], w2);
    x[3] = cmul_dir<INV>(x[3], cmul(w1, w2));
    x[4] = cmul_dir<INV>(x[4], w4);
  } else if constexpr (R >= 2) {
    x[1] = cmul_dir<INV>(x[1], w1);
    if constexpr (R >= 4) {
      const float2 w3 = cmul(w1, w2);
      x[2] = cmul_dir<INV>(x[2], w2);
      x[3] = cmul_dir<INV>(x[3], w3);
      if constexpr (R >= 8) {
        const float2 w5 = cmul(w4, w1), w6 = cmul(w4, w2), w7 = cmul(w4, w3);
        x[4] = cmul_dir<INV>(x[4], w4);
        x[5] = cmul_dir<INV>(x[5], w5);
        x[6] = cmul_dir<INV>(x[6], w6);
        x[7] = cmul_dir<INV>(x[7], w7);
        if constexpr (R >= 16) {
          x[8] = cmul_dir<INV>(x[8], w8);
          x[9] = cmul_dir<INV>(x[9], cmul(w8, w1));
          x[10] = cmul_dir<INV>(x[10], cmul(w8, w2));
          x[11] = cmul_dir<INV>(x[11], cmul(w8, w3));
          x[12] = cmul_dir<INV>(x[12], cmul(w8, w4));
          x[13] = cmul_dir<INV>(x[13], cmul(w8, w5));
          x[14] = cmul_dir<INV>(x[14], cmul(w8, w6));
          x[15] = cmul_dir<INV>(x[15], cmul(w8, w7));
        }
      }
    }
  }
}

// Twiddle source living in shared memory: for pass i (sub > 1) slots [tw_off(i), +sub) hold (W^1, W^2)
// and the next sub slots hold (W^4, W^8) of W = W_span^j — consecutive j in consecutive float4s, so a
// warp's lookups are conflict-free 16-byte loads.
template <int S>
struct TwSmem {
  const float4* tab;
  template <int STAGE>
  HY_DEVICE void get(int j, float2& w1, float2& w2, float2& w4, float2& w8) const {
    using P = Plan<S>;
    const float4 a = tab[P::tw_off(STAGE) + j];
    const float4 b = tab[P::tw_off(STAGE) + P::sub(STAGE) + j];
    w1 = make_float2(a.x, a.y);
    w2 = make_float2(a.z, a.w);
    w4 = make_float2(b.x, b.y);
    w8 = make_float2(b.z, b.w);
  }
};
// fill the table from the global W_8192 table (once per CTA)
template <int S>
HY_DEVICE void build_tw_smem(float4* tab, const float2* __restrict__ twg, int tid, int nt) {
  using P = Plan<S>;
#pragma unroll
  for (int i = 0; i < P::NS; ++i) {
    if (P::sub(i) > 1 && (HY_TWN % P::span(i)) != 0) {
      // span with an odd factor (3 * 2^k, 5 * 2^k columns): not on the W_8192 grid, evaluated directly
      for (int j = tid; j < P::sub(i); j += nt) {
        float2 w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int e = (int)(((long long)j << q) % P::span(i));
          float sn, cs;
          sincospif(2.0f * (float)e / (float)P::span(i), &sn, &cs);
          w[q] = make_float2(cs, -sn);
        }
        tab[P::tw_off(i) + j] = make_float4(w[0].x, w[0].y, w[1].x, w[1].y);
        tab[P::tw_off(i) + P::sub(i) + j] = make_float4(w[2].x, w[2].y, w[3].x, w[3].y);
      }
    } else if (P::sub(i) > 1) {
      const int stride = HY_TWN / P::span(i);
      for (int j = tid; j < P::sub(i); j += nt) {
        const int i1 = j * stride;
        const float2 w1 = __ldg(twg + i1), w2 = __ldg(twg + ((2 * i1) & (HY_TWN - 1)));
        const float2 w4 = __ldg(twg + ((4 * i1) & (HY_TWN - 1))), w8 = __ldg(twg + ((8 * i1) & (HY_TWN - 1)));
        tab[P::tw_off(i) + j] = make_float4(w1.x, w1.y, w2.x, w2.y);
        tab[P::tw_off(i) + P::sub(i) + j] = make_float4(w4.x, w4.y, w8.x, w8.y);
      }
    }
  }
}

// ---- one FFT pass over NB independent length-S transforms by NT threads ------------------------
//   LD: set_batch(int), float2 ld(int elem)      ST: set_batch(int), void st(int elem, float2 v)
//   BATCH_FAST: consecutive threads take consecutive transforms (column tiles) instead of
//               consecutive butterflies of one transform (rows).
//   ZERO_UPPER: forward pass 0 only - elements >= S/2 are known zero (zero-padded input).
//   HALF_OUT:   inverse pass 0 only - outputs >= S/2 are not needed (truncated output).
//   LD2P: the loader is two-phase — fetch(m, elem) for every input first (independent global loads,
//         all in flight together), then float2 get(m, elem) to finish each one.
//   STPF: the sink is two-phase — prefetch(m, elem) for every output (its gate operands), then
//         st_pref(m, elem, v).
template <int S, int NB, int NT, int STAGE, bool INV, bool BATCH_FAST, bool ZERO_UPPER, bool HALF_OUT,
          bool LD2P = false, bool STPF = false, class TW, class LD, class ST>
HY_DEVICE void fft_pass(const TW& tw, int tid, LD& ld, ST& st) {
  using P = Plan<S>;
  constexpr int R = P::radix(STAGE);
  constexpr int SPAN = P::span(STAGE);
  constexpr int SUB = SPAN / R;
  constexpr int NBF = S / R;
  constexpr int TOTAL = NBF * NB;
  constexpr int NIN = (ZERO_UPPER && R > 1) ? R / 2 : R;
  constexpr int NOUT = (HALF_OUT && R > 1) ? R / 2 : R;
#pragma unroll 2
  for (int bid = tid; bid < TOTAL; bid += NT) {
    int batch, w;
    if (BATCH_FAST) {
      batch = bid % NB;
      w = bid / NB;
    } else {
      w = bid % NBF;
      batch = bid / NBF;
    }
    const int blk = w / SUB, j = w % SUB;
    const int base = blk * SPAN + j;
    ld.set_batch(batch);
    float2 x[R];
    float2 w1, w2, w4, w8;
    if constexpr (SUB > 1) tw.template get<STAGE>(j, w1, w2, w4, w8);
    // accessors whose address is affine in the element index resolve `base` once; the R element offsets are
    // then compile-time immediates (no per-element shift/add for the padded shared-memory layout)
    if constexpr (!INV) {
      if constexpr (LD2P) {
#pragma unroll
        for (int m = 0; m < NIN; ++m) ld.fetch(m, base + m * SUB);
      }
      if constexpr (!LD2P && LD::kAffine) {
        const int pb = ld.pbase(base);
#pragma unroll
        for (int m = 0; m < R; ++m) x[m] = (m >= NIN) ? make_float2(0.f, 0.f) : ld.ldp(pb, m * SUB);
      } else {
#pragma unroll
        for (int m = 0; m < R; ++m) {
          if (m >= NIN) x[m] = make_float2(0.f, 0.f);
          else if constexpr (LD2P) x[m] = ld.get(m, base + m * SUB);
          else x[m] = ld.ld(base + m * SUB);
        }
      }
      RegFFT<R, false>::run(x);
      if (SUB > 1) apply_twiddles<R, false>(x, w1, w2, w4, w8);
      st.set_batch(batch);
      if constexpr (ST::kAffine) {
        const int pb = st.pbase(base);
#pragma unroll
        for (int q = 0; q < R; ++q) st.stp(pb, q * SUB, x[q]);
      } else {
#pragma unroll
        for (int q = 0; q < R; ++q) st.st(base + q * SUB, x[q]);
      }
    } else {
      if constexpr (LD::kAffine) {
        const int pb = ld.pbase(base);
#pragma unroll
        for (int q = 0; q < R; ++q) x[q] = ld.ldp(pb, q * SUB);
      } else {
#pragma unroll
        for (int q = 0; q < R; ++q) x[q] = ld.ld(base + q * SUB);
      }
      if (SUB > 1) apply_twiddles<R, true>(x, w1, w2, w4, w8);
      RegFFT<R, true>::run(x);
      st.set_batch(batch);
      if constexpr (STPF) {
        constexpr int CH = NOUT > 4 ? 4 : NOUT;   // gate operands in flight per chunk
#pragma unroll
        for (int m0 = 0; m0 < NOUT; m0 += CH) {
#pragma unroll
          for (int m = 0; m < CH; ++m) st.prefetch(m, base + (m0 + m) * SUB);
#pragma unroll
          for (int m = 0; m < CH; ++m) st.st_pref(m, base + (m0 + m) * SUB, x[m0 + m]);
        }
      } else if constexpr (ST::kAffine) {
        const int pb = st.pbase(base);
#pragma unroll
        for (int m = 0; m < NOUT; ++m) st.stp(pb, m * SUB, x[m]);
      } else {
#pragma unroll
        for (int m = 0; m < NOUT; ++m) st.st(base + m * SUB, x[m]);
      }
    }
  }
}

// ---- shared-memory row layout: one pad slot per 16 complex keeps every pass conflict-free ------
template <int S>
struct RowSmem {
  static constexpr int kRow = S + S / 16;  // float2 slots per row
};
template <int S>
struct SmemRows {
  enum { kAffine = 1 };
  float2* sm;
  int off;
  HY_DEVICE explicit SmemRows(float2* s) : sm(s), off(0) {}
  HY_DEVICE void set_batch(int b) { off = b * RowSmem<S>::kRow; }
  HY_DEVICE float2 ld(int e) const { return sm[off + e + (e >> 4)]; }
  HY_DEVICE void st(int e, float2 v) const { sm[off + e + (e >> 4)] = v; }
  // phys(base + K) = phys(base) + K + (K >> 4) for the element sets of a pass (base mod 16 + K mod 16 < 16)
  HY_DEVICE int pbase(int base) const { return off + base + (base >> 4); }
  HY_DEVICE float2 ldp(int pb, int K) const { return sm[pb + K + (K >> 4)]; }
  HY_DEVICE void stp(int pb, int K, float2 v) const { sm[pb + K + (K >> 4)] = v; }
};

// Run passes [FIRST, LAST] (forward order) of the forward transform on rows held in shared memory.
template <int S, int NB, int NT, int FIRST, int LAST, class TW>
HY_DEVICE void row_fwd_smem(float2* sm, const TW& tw, int tid) {
  if constexpr (FIRST <= LAST) {
    SmemRows<S> acc(sm);
    fft_pass<S, NB, NT, FIRST, false, false, false, false>(tw, tid, acc, acc);
    __syncthreads();
    row_fwd_smem<S, NB, NT, FIRST + 1, LAST>(sm, tw, tid);
  }
}
// Inverse passes from stage HI down to stage LO (inclusive), all in shared memory.
template <int S, int NB, int NT, int HI, int LO, class TW>
HY_DEVICE void row_inv_smem(float2* sm, const TW& tw, int tid) {
  if constexpr (HI >= LO) {
    SmemRows<S> acc(sm);
    fft_pass<S, NB, NT, HI, true, false, false, false>(tw, tid, acc, acc);
    __syncthreads();
    row_inv_smem<S, NB, NT, HI - 1, LO>(sm, tw, tid);
  }
}

// ---- packed real-sequence algebra on a (k, M-k) pair -------------------------------------------
// z[n] = x[2n] + i x[2n+1], Z = FFT_M(z).  With W = exp(-2 pi i k / N), N = 2M:
//   X[k]   = E - iT,  X[M-k] = conj(E + iT),   E = (Z[k] + conj Z[M-k])/2,  T = W (Z[k] - conj Z[M-k])/2
HY_DEVICE void unpack_pair(float2 zk, float2 zm, float2 w, float2& xk, float2& xm) {
  float2 e = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
  float2 o = make_float2(0.5f * (zk.x - zm.x), 0.5f * (zk.y + zm.y));
  float2 t = cmul(w, o);
  xk = make_float2(e.x + t.y, e.y - t.x);   // E - iT
  xm = make_float2(e.x - t.y, -e.y - t.x);  // conj(E + iT)
}
// inverse of unpack_pair: packed spectrum of the real sequence whose true spectrum is (Y[k], Y[M-k])
//   W[k] = E' + iT',  W[M-k] = conj(E' - iT'),  E' = (Y[k] + conj Y[M-k])/2,  T' = conj(W)(Y[k] - conj Y[M-k])/2
HY_DEVICE void repack_pair(float2 yk, float2 ym, float2 w, float2& wk, float2& wm) {
  float2 e = make_float2(0.5f * (yk.x + ym.x), 0.5f * (yk.y - ym.y));
  float2 o = make_float2(0.5f * (yk.x - ym.x), 0.5f * (yk.y + ym.y));
  float2 t = cmulc(o, w);                   // conj(W) * O
  wk = make_float2(e.x - t.y, e.y + t.x);   // E' + iT'
  wm = make_float2(e.x + t.y, -e.y + t.x);  // conj(E' - iT')
}
