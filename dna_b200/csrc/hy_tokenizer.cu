// hyena-b200: character tokenizer / nucleotide encoder.
// Restates CharacterTokenizer (/root/reference/src/dataloaders/datasets/hg38_char_tokenizer.py:58-94;
// standalone twin standalone_hyenadna.py:1003-1018 prepends [CLS]) as invoked by HG38Dataset
// (hg38_dataset.py:194-199: truncation=True, padding="max_length", padding_side='left', :16) plus
// the post-processing at hg38_dataset.py:216-220 (N -> PAD) and :383-386 (id-7 clipped to 4).
#include "hy_host.h"

namespace hy {

constexpr int kTokThreads = 256;
constexpr int kTokPer = 8;  // ids per thread: 8 byte loads -> 4 x 16-byte stores

HY_DEVICE int tok_lut(unsigned c) {
  // vocab: [CLS]0 [SEP]1 [BOS]2 [MASK]3 [PAD]4 [RESERVED]5 [UNK]6 A7 C8 G9 T10 N11 (case sensitive)
  return c == 'A' ? 7 : c == 'C' ? 8 : c == 'G' ? 9 : c == 'T' ? 10 : c == 'N' ? 11 : 6;
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
  const int j0 = (blockIdx.x * kTokThreads + threadIdx.x) * kTokPer;
  long long v[kTokPer];
#pragma unroll
  for (int i = 0; i < kTokPer; ++i) {
    const int j = j0 + i;
    int id = 4;
    if (j < max_length && j >= pad) {
      const int q = j - pad;
      const int ci = q - add_cls;
      if (add_cls && q == 0) id = 0;
      else if (ci < n) id = tok_lut(src[ci]);
      else id = 1;  // the only remaining slot is the trailing [SEP]
    }
    if (n_to_pad && id == 11) id = 4;
    if (nuc) {
      id -= 7;
      if (id >= 4 || id < 0) id = 4;
    }
    v[i] = id;
  }
  if (j0 + kTokPer <= max_length && ((reinterpret_cast<uintptr_t>(dst + j0) & 15) == 0)) {
#pragma unroll
    for (int i = 0; i < kTokPer; i += 2) {
      // 16-byte store of two int64 ids
      uint4 w = make_uint4((unsigned)v[i], (unsigned)(v[i] >> 32), (unsigned)v[i + 1], (unsigned)(v[i + 1] >> 32));
      *reinterpret_cast<uint4*>(dst + j0 + i) = w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < kTokPer; ++i)
      if (j0 + i < max_length) dst[j0 + i] = v[i];
  }
}

}  // namespace hy

using namespace hy;

extern "C" int hy_tokenize(const uint8_t* seqs, long long ld_in, const int32_t* lens, int max_chars, int64_t* ids, int B,
                           int max_length, int flags, void* stream) {
  if (!seqs || !ids || B < 1 || max_length < 1 || max_chars < 0) return fail(HY_ERR_ARG, "hy_tokenize: bad argument");
  const int per_cta = kTokThreads * kTokPer;
  const dim3 grid((max_length + per_cta - 1) / per_cta, B);
  HY_LAUNCH(k_tokenize, grid, kTokThreads, 0, stream, seqs, ld_in, lens, max_chars, (long long*)ids, max_length, flags);
  return check_launch("k_tokenize");
}
