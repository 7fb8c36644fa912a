// hyena-b200: character tokenizer / nucleotide encoder.
// Restates CharacterTokenizer (/root/reference/src/dataloaders/datasets/hg38_char_tokenizer.py:58-94;
// standalone twin standalone_hyenadna.py:1003-1018 prepends [CLS]) as invoked by HG38Dataset
// (hg38_dataset.py:194-199: truncation=True, padding="max_length", padding_side='left', :16) plus
// the post-processing at hg38_dataset.py:216-220 (N -> PAD) and :383-386 (id-7 clipped to 4).
#include "hy_host.h"

namespace hy {

constexpr int kTokThreads = 256;
#ifndef HY_TOK_PER
#define HY_TOK_PER 16
#endif
constexpr int kTokPer = HY_TOK_PER;  // ids per thread: HY_TOK_PER / 4 groups of 4 consecutive ids

HY_DEVICE int tok_lut(unsigned c) {
  // vocab: [CLS]0 [SEP]1 [BOS]2 [MASK]3 [PAD]4 [RESERVED]5 [UNK]6 A7 C8 G9 T10 N11 (case sensitive)
  return c == 'A' ? 7 : c == 'C' ? 8 : c == 'G' ? 9 : c == 'T' ? 10 : c == 'N' ? 11 : 6;
}

// The int64 ids are 8 of the 9 bytes per nucleotide.  A thread owns groups of 4 consecutive ids (two 16-byte stores,
// 32 contiguous bytes per lane); the groups are counted from the 32-byte boundary at or before the row start (`mis` = 0 .. 3 ids,
// e.g. the rows of a [B, 2^20 + 1] batch cycle through all four), so the vector stores stay aligned on every row and
// only a row's first / last group can be partly outside.  The 4 characters behind a group are
// fetched as two ALIGNED 32-bit words and funnel-shifted into place (their offset in the row is arbitrary: left padding,
// [CLS]); a word that is not entirely inside the caller's buffer takes byte loads.  A group that lies entirely inside
// the row's characters — all but a handful — maps its bytes through a 256-entry table in shared memory that already
// contains the N -> PAD and id - 7 post-processing (hg38_dataset.py:216-220, :383-386): 4 table loads + 2 stores; only
// the groups touching padding, [CLS] or [SEP] run the general per-id logic.
// History on B x 2^20 nt (tools/bench_tokenizer.py; a fill of the same ids runs 7.4 TB/s, ATen's uint8 -> int64 cast
// 5.8 TB/s of the same 9 B per id): 8 consecutive ids per thread, byte loads, compare chains, scalar stores on odd rows
// 2.75 TB/s; aligned pairs 3.9 TB/s; word loads 4.0 TB/s (the kernel was INSTRUCTION bound: 1 152 SASS instructions per
// 16 ids, ~31 per id); table + interior fast path 4.2 TB/s (ncu, profiles/r02o_ncu_full_tokenizer_before_256bit_stores.csv: L1 -> L2 store path, 32 half sectors per
// request); one 256-bit store per group 6.56 TB/s = the measured copy peak (profiles/r02o_tokenizer_vs_fill_copy_cast.txt).
// four int64 ids (0 .. 11) to a 32-byte aligned address: ONE 256-bit store (sm_100: STG.E.256) — a full 32-byte sector
// per lane, 1 KB contiguous per warp instruction.  (Two 16-byte stores per lane put 32 half-filled sectors into every
// request: ncu showed the L1 -> L2 store path as the limiter, 134 M sectors for 2.1 GB of ids.)
HY_DEVICE void st_ids4(long long* p, unsigned a, unsigned b, unsigned c, unsigned d) {
#ifdef HY_EMU_BUILD
  p[0] = a; p[1] = b; p[2] = c; p[3] = d;
#else
  asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"((unsigned long long)a), "l"((unsigned long long)b),
               "l"((unsigned long long)c), "l"((unsigned long long)d)
               : "memory");
#endif
}

__global__ void __launch_bounds__(kTokThreads) k_tokenize(const uint8_t* __restrict__ seqs, long long ld_in,
                                                        const int32_t* __restrict__ lens, int max_chars,
                                                        long long* __restrict__ ids, int max_length, int flags) {
  const int b = blockIdx.y;
  const int add_sep = flags & 1, add_cls = (flags >> 1) & 1, n_to_pad = (flags >> 2) & 1, nuc = (flags >> 3) & 1;
  const int n_special = add_sep + add_cls;
  int len = lens ? lens[b] : max_chars;
  if (len > max_chars) len = max_chars;
  if (len < 0) len = 0;
  int room = max_length - n_special;
  if (room < 0) room = 0;
  const int n = len < room ? len : room;          // truncation keeps the first `room` characters
  const int total = n + n_special;
  const int pad = max_length - total;             // left padding
  const uint8_t* src = seqs + (long long)b * ld_in;
  long long* dst = ids + (long long)b * max_length;
  const int mis = (int)((reinterpret_cast<uintptr_t>(dst) >> 3) & 3);   // ids between the 32-byte boundary and the row start
  const int off = pad + add_cls;                  // id j holds character j - off (when that is one)
  // aligned 32-bit words entirely inside [seqs, seqs + (B - 1) * ld_in + max_chars) may be read whole
  const uintptr_t w_lo = (reinterpret_cast<uintptr_t>(seqs) + 3) & ~(uintptr_t)3;
  const uintptr_t w_hi = reinterpret_cast<uintptr_t>(seqs) + (uintptr_t)((long long)(gridDim.y - 1) * ld_in + max_chars);
  HY_DYN_SMEM(uint8_t, lut);   // [256]
  {
    int id = tok_lut(threadIdx.x);
    if (n_to_pad && id == 11) id = 4;
    if (nuc) {
      id -= 7;
      if (id >= 4 || id < 0) id = 4;
    }
    lut[threadIdx.x] = (uint8_t)id;
  }
  __syncthreads();
  constexpr int NG = kTokPer / 4;
  int j0[NG];
  unsigned ch[NG];
#pragma unroll
  for (int c = 0; c < NG; ++c) {
    j0[c] = 4 * ((blockIdx.x * NG + c) * kTokThreads + threadIdx.x) - mis;
    const int ci0 = j0[c] - off;
    unsigned v = 0;
    if (ci0 + 3 >= 0 && ci0 < n) {
      const uintptr_t a = reinterpret_cast<uintptr_t>(src) + (long long)ci0;
      const uintptr_t a0 = a & ~(uintptr_t)3;
      if (a0 >= w_lo && a0 + 8 <= w_hi) {
        const unsigned w0 = __ldg(reinterpret_cast<const unsigned*>(a0));
        const unsigned w1 = __ldg(reinterpret_cast<const unsigned*>(a0 + 4));
        v = __funnelshift_r(w0, w1, 8u * (unsigned)(a & 3));
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (ci0 + e >= 0 && ci0 + e < n) v |= (unsigned)__ldg(src + ci0 + e) << (8 * e);
      }
    }
    ch[c] = v;
  }
#pragma unroll
  for (int c = 0; c < NG; ++c) {
    unsigned v[4];
    const int ci0 = j0[c] - off;
    if (ci0 >= 0 && ci0 + 3 < n) {   // four characters (then 0 <= j0 and j0 + 3 < max_length as well)
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = lut[(ch[c] >> (8 * e)) & 0xffu];
      st_ids4(dst + j0[c], v[0], v[1], v[2], v[3]);
      continue;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = j0[c] + e;
      int id = 4;
      if (j >= pad) {
        const int q = j - pad;
        const int ci = q - add_cls;
        if (add_cls && q == 0) id = 0;
        else if (ci < n) id = tok_lut((ch[c] >> (8 * e)) & 0xffu);
        else id = 1;  // the only remaining slot is the trailing [SEP]
      }
      if (n_to_pad && id == 11) id = 4;
      if (nuc) {
        id -= 7;
        if (id >= 4 || id < 0) id = 4;
      }
      v[e] = (unsigned)id;
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = j0[c] + 2 * h;
      if (j >= 0 && j + 1 < max_length) {
        *reinterpret_cast<uint4*>(dst + j) = make_uint4(v[2 * h], 0u, v[2 * h + 1], 0u);   // ids are 0 .. 11
      } else {
        if (j >= 0 && j < max_length) dst[j] = (long long)v[2 * h];
        if (j + 1 >= 0 && j + 1 < max_length) dst[j + 1] = (long long)v[2 * h + 1];
      }
    }
  }
}

// ---- reverse complement (string_reverse_complement, hg38_dataset.py:28-38; the rc_aug branch of FastaInterval,
// :118-119): out[i] = comp(in[len-1-i]) with A<->T, C<->G, a<->t, c<->g and every other byte (N, n, '.', ...) kept.
// Rows whose apply[b] == 0 (or len <= 1 with nothing to swap) are copied; bytes at and beyond len are copied as they are.
constexpr int kRcPer = 16;  // output bytes per thread (one 16-byte store when aligned)

HY_DEVICE unsigned rc_comp(unsigned c) {
  return c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : c == 'G' ? 'C'
       : c == 'a' ? 't' : c == 't' ? 'a' : c == 'c' ? 'g' : c == 'g' ? 'c' : c;
}

__global__ void __launch_bounds__(kTokThreads) k_reverse_complement(const uint8_t* __restrict__ seqs, long long ld_in,
                                                                  const int32_t* __restrict__ lens,
                                                                  const uint8_t* __restrict__ apply, uint8_t* __restrict__ out,
                                                                  long long ld_out, int max_chars) {
  const int b = blockIdx.y;
  int len = lens ? lens[b] : max_chars;
  len = len < 0 ? 0 : (len > max_chars ? max_chars : len);
  const bool rc = apply ? apply[b] != 0 : true;
  const uint8_t* src = seqs + (long long)b * ld_in;
  uint8_t* dst = out + (long long)b * ld_out;
  const int j0 = (blockIdx.x * kTokThreads + threadIdx.x) * kRcPer;
  if (j0 >= max_chars) return;
  unsigned v[kRcPer];
#pragma unroll
  for (int i = 0; i < kRcPer; ++i) {
    const int j = j0 + i;
    unsigned c = 0;
    if (j < max_chars) {
      if (rc && j < len) c = rc_comp(src[len - 1 - j]);
      else c = src[j];
    }
    v[i] = c;
  }
  if (j0 + kRcPer <= max_chars && ((reinterpret_cast<uintptr_t>(dst + j0) & 15) == 0)) {
    uint4 w;
    w.x = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
    w.y = v[4] | (v[5] << 8) | (v[6] << 16) | (v[7] << 24);
    w.z = v[8] | (v[9] << 8) | (v[10] << 16) | (v[11] << 24);
    w.w = v[12] | (v[13] << 8) | (v[14] << 16) | (v[15] << 24);
    *reinterpret_cast<uint4*>(dst + j0) = w;
  } else {
#pragma unroll
    for (int i = 0; i < kRcPer; ++i)
      if (j0 + i < max_chars) dst[j0 + i] = (uint8_t)v[i];
  }
}

// ---- interval fetch (FastaInterval.__call__, hg38_dataset.py:72-124) over a chromosome resident in HBM --------------
// Per row: widen [start, end) symmetrically to max_length when shorter (:92-99), clip to the chromosome and remember the
// clipped amounts as left / right padding (:101-107), cut to max_length when longer (:110-111), optionally reverse-
// complement the fetched bytes (:118-119), then surround with '.' padding when pad_interval (:121-122).  Writes the
// bytes and the row's length; bytes beyond the length are '.' (never read by the tokenizer, which takes `lens`).
__global__ void __launch_bounds__(kTokThreads) k_fetch_intervals(const uint8_t* __restrict__ chrom, long long chrom_len,
                                                               const long long* __restrict__ starts,
                                                               const long long* __restrict__ ends,
                                                               const uint8_t* __restrict__ rc, int max_length, int pad_interval,
                                                               uint8_t* __restrict__ out, long long ld_out,
                                                               int32_t* __restrict__ lens, int width) {
  const int b = blockIdx.y;
  long long start = starts[b], end = ends[b];
  const long long interval = end - start;
  long long lpad = 0, rpad = 0;
  if (interval < max_length) {
    const long long extra = max_length - interval, left = extra / 2;
    start -= left;
    end += extra - left;
  }
  if (start < 0) { lpad = -start; start = 0; }
  if (end > chrom_len) { rpad = end - chrom_len; end = chrom_len; }
  if (interval > max_length) end = start + max_length;
  long long n = end - start;
  if (n < 0) n = 0;
  if (!pad_interval) lpad = rpad = 0;
  const long long total = lpad + n + rpad;
  if (blockIdx.x == 0 && threadIdx.x == 0) lens[b] = (int32_t)(total < width ? total : width);
  const bool flip = rc && rc[b];
  uint8_t* dst = out + (long long)b * ld_out;
  const int j0 = (blockIdx.x * kTokThreads + threadIdx.x) * kRcPer;
  if (j0 >= width) return;
  unsigned v[kRcPer];
#pragma unroll
  for (int i = 0; i < kRcPer; ++i) {
    const long long j = j0 + i;
    unsigned c = '.';
    const long long q = j - lpad;
    if (q >= 0 && q < n) c = flip ? rc_comp(chrom[start + (n - 1 - q)]) : chrom[start + q];
    v[i] = c;
  }
  if (j0 + kRcPer <= width && ((reinterpret_cast<uintptr_t>(dst + j0) & 15) == 0)) {
    uint4 w;
    w.x = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
    w.y = v[4] | (v[5] << 8) | (v[6] << 16) | (v[7] << 24);
    w.z = v[8] | (v[9] << 8) | (v[10] << 16) | (v[11] << 24);
    w.w = v[12] | (v[13] << 8) | (v[14] << 16) | (v[15] << 24);
    *reinterpret_cast<uint4*>(dst + j0) = w;
  } else {
#pragma unroll
    for (int i = 0; i < kRcPer; ++i)
      if (j0 + i < width) dst[j0 + i] = (uint8_t)v[i];
  }
}

// ---- BERT masking (bert_mask, hg38_dataset.py:238-286) given the random draws -------------------------------------------
// mask = (seq != pad) & (r_mask < mask_prob); labels = mask ? seq : -100;
// masked & r_kind < 1 - p_random - p_keep -> [MASK];  masked & r_kind in [1 - p_random - p_keep, 1 - p_keep) -> rand_tok;
// the rest unchanged.  One pass: 24 B read + 17 B written per token instead of the reference's ~10 elementwise kernels.
__global__ void __launch_bounds__(kTokThreads) k_bert_mask(const long long* __restrict__ seq, const float* __restrict__ r_mask,
                                                         const float* __restrict__ r_kind, const long long* __restrict__ rand_tok,
                                                         long long n, long long mask_id, long long pad_id, float mask_prob,
                                                         float th_mask, float th_random, long long* __restrict__ out,
                                                         uint8_t* __restrict__ mask, long long* __restrict__ labels) {
  for (long long i = (long long)blockIdx.x * kTokThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kTokThreads) {
    const long long s = seq[i];
    const bool m = (s != pad_id) && (r_mask[i] < mask_prob);
    const float r = r_kind[i];
    long long o = s;
    if (m && r < th_mask) o = mask_id;
    else if (m && r >= th_mask && r < th_random) o = rand_tok[i];
    out[i] = o;
    mask[i] = m ? 1 : 0;
    labels[i] = m ? s : -100;
  }
}

}  // namespace hy

using namespace hy;

extern "C" int hy_fetch_intervals(const uint8_t* chrom, long long chrom_len, const long long* starts, const long long* ends,
                                  const uint8_t* rc, int B, int max_length, int pad_interval, uint8_t* out, long long ld_out,
                                  int32_t* lens, int width, void* stream) {
  if (!chrom || !starts || !ends || !out || !lens || B < 1 || max_length < 1 || width < 1 || chrom_len < 0 || ld_out < width)
    return fail(HY_ERR_ARG, "hy_fetch_intervals: bad argument");
  const int per_cta = kTokThreads * kRcPer;
  const dim3 grid((width + per_cta - 1) / per_cta, B);
  HY_LAUNCH(k_fetch_intervals, grid, kTokThreads, 0, stream, chrom, chrom_len, starts, ends, rc, max_length, pad_interval, out,
            ld_out, lens, width);
  return check_launch("k_fetch_intervals");
}

extern "C" int hy_bert_mask(const int64_t* seq, const float* r_mask, const float* r_kind, const int64_t* rand_tok, long long n,
                            long long mask_id, long long pad_id, float mask_prob, float random_token_prob,
                            float unchanged_token_prob, int64_t* out, uint8_t* mask, int64_t* labels, void* stream) {
  if (!seq || !r_mask || !r_kind || !rand_tok || !out || !mask || !labels || n < 0) return fail(HY_ERR_ARG, "hy_bert_mask: bad argument");
  if (n == 0) return HY_OK;
  // the thresholds exactly as the reference forms them (python floats = double, compared with float32 draws)
  const float th_mask = (float)(1.0 - (double)random_token_prob - (double)unchanged_token_prob);
  const float th_random = (float)(1.0 - (double)unchanged_token_prob);
  long long blocks = (n + kTokThreads - 1) / kTokThreads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  HY_LAUNCH(k_bert_mask, (int)blocks, kTokThreads, 0, stream, (const long long*)seq, r_mask, r_kind, (const long long*)rand_tok, n,
            mask_id, pad_id, mask_prob, th_mask, th_random, (long long*)out, mask, (long long*)labels);
  return check_launch("k_bert_mask");
}

extern "C" int hy_reverse_complement(const uint8_t* seqs, long long ld_in, const int32_t* lens, const uint8_t* apply,
                                     uint8_t* out, long long ld_out, int B, int max_chars, void* stream) {
  if (!seqs || !out || B < 1 || max_chars < 0 || seqs == out) return fail(HY_ERR_ARG, "hy_reverse_complement: bad argument");
  if (max_chars == 0) return HY_OK;
  const int per_cta = kTokThreads * kRcPer;
  const dim3 grid((max_chars + per_cta - 1) / per_cta, B);
  HY_LAUNCH(k_reverse_complement, grid, kTokThreads, 0, stream, seqs, ld_in, lens, apply, out, ld_out, max_chars);
  return check_launch("k_reverse_complement");
}

extern "C" int hy_tokenize(const uint8_t* seqs, long long ld_in, const int32_t* lens, int max_chars, int64_t* ids, int B,
                           int max_length, int flags, void* stream) {
  if (!seqs || !ids || B < 1 || max_length < 1 || max_chars < 0) return fail(HY_ERR_ARG, "hy_tokenize: bad argument");
  const int per_cta = kTokThreads * kTokPer;
  const dim3 grid((max_length + 3 + per_cta - 1) / per_cta, B);   // + 3: a row may start up to 3 ids after a 32-byte boundary
  HY_LAUNCH(k_tokenize, grid, kTokThreads, 256, stream, seqs, ld_in, lens, max_chars, (long long*)ids, max_length, flags);
  return check_launch("k_tokenize");
}
