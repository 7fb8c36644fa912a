/* hyena_b200.h — C ABI of libhyena_b200.so (hand-written sm_100a CUDA for the HyenaDNA hot path).
 *
 * This is the boundary the reference binds at: the (absent) compiled module `fftconv` that
 * /root/reference/src/ops/fftconv.py:8 imports (`fftconv_fwd`, `fftconv_bwd`, called at :84 and :96-97),
 * plus the fused pieces of HyenaOperator.forward (src/models/sequence/hyena.py:436-508) and of the
 * data pipeline tokenizer (src/dataloaders/datasets/hg38_char_tokenizer.py:58-94).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless said otherwise;
 *   - the caller owns all memory (outputs and workspaces are pre-allocated by the caller);
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*), nothing synchronises;
 *   - every entry point returns 0 on success, a negative hy_status otherwise; the message is
 *     available from hy_last_error() (thread local);
 *   - there is no CPU fallback: without a usable CUDA device hy_init() fails.
 */
#ifndef HYENA_B200_H_
#define HYENA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  HY_OK = 0,
  HY_ERR_ARG = -1,        /* bad shape / mode / null pointer */
  HY_ERR_UNSUPPORTED = -2,/* length or dtype outside the supported set */
  HY_ERR_WORKSPACE = -3,  /* workspace too small */
  HY_ERR_CUDA = -4        /* a CUDA runtime call failed (message has the CUDA error string) */
} hy_status;

typedef enum { HY_F32 = 0, HY_BF16 = 1 } hy_dtype;

/* gating modes of the fused long convolution (mirrors the v/q arguments of fftconv_fwd,
 * src/ops/fftconv.py:84, with head_dim = 1, and HyenaOperator's gates, hyena.py:481,496) */
typedef enum {
  HY_INPUT_PLAIN = 0,     /* g = u                                        (fftconv_func(u,k,D))           */
  HY_INPUT_PREGATE = 1,   /* g = u * pre                                  (H3 `k*v`, ops/fftconv.py:41-42) */
  HY_INPUT_SHORTCONV = 2  /* g = sc(uT[2H+c]) * sc(uT[H+c]), sc = causal depthwise 3-tap (hyena.py:444-481) */
} hy_in_mode;
typedef enum {
  HY_OUTPUT_PLAIN = 0,    /* out = y                                                                 */
  HY_OUTPUT_POSTGATE = 1, /* out = y * post                              (H3 `* q`, ops/fftconv.py:55) */
  HY_OUTPUT_SHORTCONV = 2 /* out = y * sc(uT[c])                         (hyena.py:496-503)            */
} hy_out_mode;

/* ---- library state ---------------------------------------------------------------------- */
int hy_init(void);                 /* builds twiddle tables on the current device; idempotent */
const char* hy_last_error(void);
const char* hy_version(void);
/* complex transform length M used for sequence length L (real FFT size N = 2M >= 2L): a power of two up to 4096,
 * beyond that the smallest of 2^a, 3 * 2^a, 5 * 2^a that holds L (the reference transforms exactly 2L points,
 * src/models/sequence/hyena.py:61-62); replaces `fft_size = max(2 * 2**ceil(log2 L), 16)` of src/ops/fftconv.py:64 */
int hy_fft_len(int L);
/* bytes of scratch the long-conv entry points need (nseq = 1 forward/spectrum/dk, 2 backward) */
size_t hy_conv_workspace_bytes(int B, int H, int L, int nseq);
/* bytes of the optional saved-spectrum buffer of hy_conv_fwd / hy_conv_bwd (complex64 [B][H][M]) */
size_t hy_conv_gsave_bytes(int B, int H, int L);
/* number of partial-sum columns of the dD output of hy_conv_bwd for length L */
int hy_conv_ndpart(int L);
/* number of CUDA kernels this library has launched so far in this process (monotonic) */
unsigned long long hy_launch_count(void);
/* measurement aid: a 1-thread kernel that spins ~50 us and writes out2[0] = SM cycles, out2[1] = nanoseconds
 * (globaltimer) elapsed, i.e. the SM clock the surrounding kernels run at, without an NVML query */
int hy_clock_probe(unsigned long long* out2, void* stream);
/* scratch budget (bytes) used to size row groups of the four-step path; 0 restores the default */
int hy_set_scratch_budget(size_t bytes);
/* four-step path scheduling: row groups are issued round-robin on `nstream` internal streams (forked from and
 * joined back into the caller's stream with events, so the call stays stream-ordered and graph-capturable);
 * `scratch_bytes` (0 = keep) is the total scratch of the groups in flight. nstream = 1 uses only the caller's stream. */
int hy_set_pipeline(int nstream, size_t scratch_bytes);

/* ---- filter spectrum: replaces `k_f = torch.fft.rfft(k, n=fft_size)` (ops/fftconv.py:65) ----
 * Kf[h][M] (complex64, internal position order) = spectrum of (k[h] + D[h]*delta) / M.
 * k: fp32 [H][ldk]; D: fp32 [H] or NULL. */
int hy_filter_spectrum(const float* k, int ldk, const float* D, void* Kf, int H, int L,
                       void* ws, size_t ws_bytes, void* stream);

/* ---- fused long convolution, forward: replaces fftconv_fwd (ops/fftconv.py:84) ------------ */
typedef struct {
  int dtype;              /* hy_dtype of u/pre/post/out/ysave */
  int B, H, L;
  int in_mode, out_mode;  /* SHORTCONV must be used for both or neither */
  const void* u;          /* PLAIN/PREGATE: [B][H][ldu]; SHORTCONV: uT [B][3H][ldu] */
  const void* pre;        /* PREGATE only, strides of u */
  long long u_bs; int ldu;
  const void* post;       /* POSTGATE only: [B][H][ldpost] */
  long long post_bs; int ldpost;
  const float* sw;        /* SHORTCONV: short filter weight [3H][3]  (Conv1d weight [3H,1,3]) */
  const float* sb;        /* SHORTCONV: short filter bias [3H] */
  const float* pb;        /* SHORTCONV: in_proj bias [3H] added before the short filter, or NULL */
  const void* Kf;         /* from hy_filter_spectrum */
  void* out;              /* [B][H][ldo] */
  void* ysave;            /* optional: pre-gate y (needed by the backward of gated modes), strides of out */
  long long out_bs; int ldo;
  void* ws; size_t ws_bytes;
  void* gsave;            /* optional, complex64 [B][H][M], hy_conv_gsave_bytes(B, H, L) bytes (0: this L has no use for
                           * it, pass NULL): the row-transformed spectrum of g, which hy_conv_bwd can read back instead of
                           * transforming g a second time (8*M bytes per (b, h) row traded for ~1/3 of the backward) */
} hy_conv_fwd_args;
int hy_conv_fwd(const hy_conv_fwd_args* a, void* stream);

/* ---- fused long convolution, backward: replaces fftconv_bwd (ops/fftconv.py:96-97) --------
 * Produces du (PLAIN/PREGATE) or dX = (dx0 | dx1 | dv) in uT layout (SHORTCONV), dpre, dpost,
 * the per-(slot, channel) spectrum products dKacc (turned into dk by hy_conv_dk) and partial
 * sums of dD: dD[h] = sum_b sum_j dDpart[(b*H+h)*ndpart + j]. */
typedef struct {
  int dtype;
  int B, H, L;
  int in_mode, out_mode;
  const void* u; const void* pre; long long u_bs; int ldu;
  const void* post; long long post_bs; int ldpost;
  const float* sw; const float* sb; const float* pb;
  const void* Kf;
  const void* dout;       /* [B][H][ldo] */
  long long out_bs; int ldo;
  const void* ysave;      /* y saved by the forward (gated output modes): [B][H][ldys] */
  long long ys_bs; int ldys;
  void* du;               /* strides of u (SHORTCONV: [B][3H][ldu]) */
  void* dpre;             /* PREGATE */
  void* dpost;            /* POSTGATE, strides of post */
  void* dKacc;            /* complex64 [nslot][H][M] */
  int nslot;              /* 1 <= nslot <= B; batches b, b+nslot, ... accumulate into one slot */
  float* dDpart;          /* fp32 [B*H][ndpart] */
  void* ws; size_t ws_bytes;
  const void* gsave;      /* optional: what the forward wrote (same B, H, L, same inputs). When given (and
                           * hy_conv_gsave_bytes > 0) g is not re-transformed, the scratch need is that of nseq = 1, and
                           * dDpart is NOT written: dD[h] is then dk[h][0] of hy_conv_dk (the zero-lag correlation). */
  int defer_dx0;          /* SHORTCONV only: leave the x0 group of dX unwritten (ysave is not read); the caller forms
                           * dx0 = dout * ysave inside hy_shortconv_bwd_gate, which streams those rows anyway */
} hy_conv_bwd_args;
int hy_conv_bwd(const hy_conv_bwd_args* a, void* stream);
/* dk[h][:L] = irfft(sum_slot dKacc[slot][h])[:L] (fp32, row stride lddk) */
int hy_conv_dk(const void* dKacc, int nslot, float* dk, int lddk, int H, int L,
               void* ws, size_t ws_bytes, void* stream);

/* ---- short depthwise causal conv, backward (hyena.py:407-413,444) -------------------------
 * dX [B][3H][ld] -> duT [B][3H][ld] (gradient wrt the in_proj output), and partial sums
 * dwpart [nchunk][3H][4] = (dw0, dw1, dw2, dbias) over (batch, sequence-chunk); in_proj bias
 * gradient dpb[ch] = sum_t duT[ch][t] is returned as dpbpart [nchunk][3H]. */
int hy_shortconv_nchunk(int B, int L);
int hy_shortconv_bwd(int dtype, const void* uT, const void* dX, void* duT, long long bs, int ld,
                     const float* sw, const float* pb, float* dwpart, float* dpbpart,
                     int B, int H3, int L, void* stream);
/* standalone forward (tests / generic use): xc = short_filter(uT + pb)[..., :L] */
/* hy_shortconv_bwd whose x0 group of dX (channels [0, H3/3)) is formed on the fly as dout * ysave (rounded to the
 * activation dtype) -- the partner of hy_conv_bwd_args.defer_dx0. dout / ysave: [B][H3/3][ld*]; every tensor must be
 * 16-byte aligned with strides that are multiples of 8 elements (else HY_ERR_UNSUPPORTED: use the two-step form). */
int hy_shortconv_bwd_gate(int dtype, const void* uT, const void* dX, void* duT, long long bs, int ld,
                          const float* sw, const float* pb, float* dwpart, float* dpbpart, int B, int H3, int L,
                          const void* dout, long long dout_bs, int lddout, const void* ysave, long long ys_bs, int ldys,
                          void* stream);
int hy_shortconv_fwd(int dtype, const void* uT, void* xc, long long bs, int ld,
                     const float* sw, const float* sb, const float* pb, int B, int H3, int L, void* stream);

/* ---- implicit filter (HyenaFilter.filter, hyena.py:233-242): k[c][t] --------------------- */
typedef struct {
  int L, D, order, emb_dim, n_inner;  /* order = MLP width (<= 64), n_inner = number of hidden Linear(order,order) */
  const float* z; int ldz;            /* [L][ldz] positional features (first emb_dim columns used) */
  const float* t;                     /* [L] time axis */
  const float* w_in; const float* b_in;    /* [order][emb_dim], [order] */
  const float* w_h; const float* b_h;      /* [n_inner][order][order], [n_inner][order] */
  const float* w_out;                      /* [D][order] (no bias) */
  const float* freq;                       /* [order] (shared Sin) */
  const float* deltas;                     /* [D] */
  float shift; int modulate;
} hy_filter_args;
/* k: fp32 [D][ldk] (channel-major, the layout hy_filter_spectrum consumes, i.e. the reference's
 * `rearrange(k, 'l d -> d l')` of hyena.py:460 is free). */
int hy_filter_fwd(const hy_filter_args* a, float* k, int ldk, void* stream);
/* Same as hy_filter_fwd, and also stores the last hidden activation h_last[t][0..order) (fp32, row stride ldh >= order,
 * 16-byte aligned) so the backward needs no recompute of the MLP trunk. emb_dim <= 8, n_inner <= 2 only. */
int hy_filter_fwd_save(const hy_filter_args* a, float* k, int ldk, float* h_last, int ldh, void* stream);
/* Backward of the modulation + layout change: dh[t][c] (fp32 [L][ldh]) = dk[c][t] * (exp(-t[t] |deltas[c]|) + shift)
 * (modulate = 0: plain transpose). dh is the gradient wrt the output of the MLP's last Linear (hyena.py:219, 156-159);
 * the remaining MLP gradient is dense GEMM work done with cuBLAS by the host layer. */
int hy_filter_modulate_bwd(const float* dk, int lddk, const float* t, const float* deltas, float shift, int modulate,
                           float* dh, int ldh, int L, int D, void* stream);

/* Fused backward of the MLP trunk (every Linear+Sin before the last Linear): from dh_last [L][lddh] (gradient wrt
 * the last hidden activation) accumulate per-CTA partial sums of dW_in, db_in, dW_h[*], db_h[*], dfreq into
 * part[n_cta][stride] with the layout [dW_in order*emb][db_in order][per hidden layer: dW_h order*order, db_h order]
 * [dfreq order]; the caller sums over n_cta. Supports order <= 64, emb_dim <= 8, n_inner <= 2. */
int hy_filter_trunk_bwd_layout(const hy_filter_args* a, int* n_cta, int* stride);
int hy_filter_trunk_bwd(const hy_filter_args* a, const float* dh_last, int lddh, float* part, void* stream);
/* Saved-trunk variant: the forward also keeps the trunk's pre-activations a_l[j][t] (input of every Sin) as
 * a_save [1 + n_inner][64][lda] fp32 (lda = L rounded up to 64, 16-byte aligned; hy_filter_trunk_save_layout gives lda
 * and the element count: 768 MB per layer at L = 1 M), and the backward reads them instead of re-running the trunk's
 * Linear layers (a third of its arithmetic). Same outputs as hy_filter_fwd_save / hy_filter_trunk_bwd. */
int hy_filter_trunk_save_layout(const hy_filter_args* a, int* lda, long long* elems);
int hy_filter_fwd_save_trunk(const hy_filter_args* a, float* k, int ldk, float* h_last, int ldh, float* a_save, int lda,
                             void* stream);
int hy_filter_trunk_bwd_saved(const hy_filter_args* a, const float* dh_last, int lddh, const float* a_save, int lda,
                              float* part, void* stream);

/* Backward of the MLP's last Linear (implicit_filter[-1], hyena.py:219; no bias) fused with the modulation backward,
 * on the tensor cores (mma.sync TF32 with the 3xTF32 split: fp32-class accuracy). From dk [D][lddk] (channel-major, as
 * hy_conv_dk writes it) and h_last [L][ldh] (hy_filter_fwd_save):
 *   dh[t][c]       = dk[c][t] * (exp(-t[t] |deltas[c]|) + shift)      (never materialised)
 *   dh_last[t][o]  = sum_c dh[t][c] * w_out[c][o]                     -> [L][lddh], input of hy_filter_trunk_bwd
 *   dW_out[c][o]   = sum_t dh[t][c] * h_last[t][o]                    -> [D][order]
 * Replaces hy_filter_modulate_bwd + two cuBLAS SGEMMs (three passes over an [L][D] intermediate). Supported for
 * order == 64, D % 32 == 0, D <= 256 (hy_filter_out_bwd_supported); dk and h_last rows 16-byte aligned. The per-CTA
 * partial sums of dW_out live in the caller's workspace (hy_filter_out_bwd_workspace_bytes) and are reduced in a fixed
 * order (deterministic). */
int hy_filter_out_bwd_supported(int D, int order);
size_t hy_filter_out_bwd_workspace_bytes(int L);
int hy_filter_out_bwd(const float* dk, int lddk, const float* t, const float* deltas, float shift, int modulate,
                      const float* w_out, const float* h_last, int ldh, float* dh_last, int lddh, float* dW_out,
                      int D, int order, int L, void* ws, size_t ws_bytes, void* stream);

/* ---- character tokenizer (hg38_char_tokenizer.py:58-94, hg38_dataset.py:194-223,383-386) ---
 * seqs: uint8 [B][ld_in] ASCII; lens: int32 [B] (NULL = max_chars for every row).
 * ids: int64 [B][max_length]: LUT (A,C,G,T,N -> 7..11, else 6), truncation to
 * max_length - n_special, optional [CLS]=0 prefix / [SEP]=1 suffix, LEFT padding with [PAD]=4.
 * flags: bit0 add [SEP], bit1 add [CLS] (standalone variant), bit2 replace N(11) by PAD(4),
 *        bit3 emit `id-7 clipped to 4` nucleotide encoding (hg38_dataset.py:383-386). */
int hy_tokenize(const uint8_t* seqs, long long ld_in, const int32_t* lens, int max_chars,
                int64_t* ids, int B, int max_length, int flags, void* stream);

/* ---- reverse complement of a byte batch (string_reverse_complement, hg38_dataset.py:28-38, the rc_aug branch of
 * FastaInterval, :118-119): out[b][i] = comp(seqs[b][len-1-i]) for i < len with A<->T, C<->G, a<->t, c<->g, every
 * other byte kept; bytes in [len, max_chars) are copied. apply: uint8 [B] (NULL = all rows); rows with apply == 0 are
 * copied unchanged. Out of place (out != seqs). */
int hy_reverse_complement(const uint8_t* seqs, long long ld_in, const int32_t* lens, const uint8_t* apply,
                          uint8_t* out, long long ld_out, int B, int max_chars, void* stream);

/* ---- Block glue: residual add + LayerNorm in one pass (standalone_hyenadna.py:520-541; the src tree's
 * dropout_add_layer_norm hook, long_conv_lm.py:560-575, with dropout p = 0 as every HyenaDNA config has) ----
 *   r = x + res_in (rounded to res_dtype);  y = (r - mean r) * rsqrt(var r + eps) * gamma + beta
 * x, res_in: [rows][D] (either may be NULL, not both); y: [rows][D] of y_dtype; res_out: [rows][D] of res_dtype or
 * NULL (do not store r); mean, rstd: fp32 [rows] saved for the backward; gamma, beta fp32 [D].
 * D in {128, 256, 512, 1024} (hy_add_ln_supported); all tensors dense, 16-byte aligned. */
int hy_add_ln_supported(int D);
int hy_add_ln_fwd(const void* x, int x_dtype, const void* res_in, int res_dtype, const float* gamma,
                  const float* beta, float eps, void* y, int y_dtype, void* res_out, float* mean, float* rstd,
                  long long rows, int D, void* stream);
/* Backward: dr = LN'(dy) + dres_out, written to dx (x_dtype) and/or dres_in (res_dtype); dgamma, dbeta fp32 [D].
 * r is the residual stream the forward normalised (its res_out, or its x when res_out was NULL and res_in NULL).
 * part: fp32 scratch [hy_add_ln_bwd_parts(rows, D)][2][D]. */
int hy_add_ln_bwd_parts(long long rows, int D);
int hy_add_ln_bwd(const void* dy, int y_dtype, const void* dres_out, int res_dtype, const void* r,
                  const float* mean, const float* rstd, const float* gamma, void* dx, int x_dtype,
                  void* dres_in, float* part, float* dgamma, float* dbeta, long long rows, int D, void* stream);

/* The same with the Block's dropout applied to x first (standalone_hyenadna.py:521 `dropped = self.dropout1(hidden_states)`;
 * the training configs set embed_dropout = 0.1 on the first block, hg38_hyena.yaml:13): keep is the caller-drawn mask
 * uint8 [rows][D] (1 = keep), keep_scale = 1 / (1 - p);  r = round_xdtype(x * keep * keep_scale) + res_in.  The backward
 * applies the same mask to dx (dres_in is unaffected). */
int hy_add_ln_dropout_fwd(const void* x, int x_dtype, const unsigned char* keep, float keep_scale, const void* res_in,
                          int res_dtype, const float* gamma, const float* beta, float eps, void* y, int y_dtype,
                          void* res_out, float* mean, float* rstd, long long rows, int D, void* stream);
int hy_add_ln_dropout_bwd(const void* dy, int y_dtype, const void* dres_out, int res_dtype, const void* r,
                          const float* mean, const float* rstd, const float* gamma, void* dx, int x_dtype,
                          const unsigned char* keep, float keep_scale, void* dres_in, float* part, float* dgamma,
                          float* dbeta, long long rows, int D, void* stream);

/* ---- interval fetch (FastaInterval.__call__, hg38_dataset.py:72-124) from a chromosome resident in device memory ----
 * chrom: uint8 [chrom_len]; starts / ends: int64 [B] (BED interval, end exclusive); rc: uint8 [B] or NULL (reverse-
 * complement the fetched bytes, :118-119 — the caller draws the coin flips); max_length as passed by the dataset.
 * out: uint8 [B][ld_out], row b = '.' * left_padding + bases + '.' * right_padding (pad_interval) or the bases alone,
 * bytes up to `width` beyond the row's length are '.'; lens: int32 [B].  shift_augs are applied by the caller to
 * starts / ends before the call (:82-90: a uniform integer shift clipped to the chromosome). */
int hy_fetch_intervals(const uint8_t* chrom, long long chrom_len, const long long* starts, const long long* ends,
                       const uint8_t* rc, int B, int max_length, int pad_interval, uint8_t* out, long long ld_out,
                       int32_t* lens, int width, void* stream);

/* ---- BERT masking (bert_mask, hg38_dataset.py:238-286) given the random draws: r_mask, r_kind fp32 uniform [0,1) and
 * rand_tok int64 (replacement tokens, already free of special ids), all [n].  out / mask (uint8) / labels (-100 where
 * not masked) [n].  Bit-exact with the reference for the same draws. */
int hy_bert_mask(const int64_t* seq, const float* r_mask, const float* r_kind, const int64_t* rand_tok, long long n,
                 long long mask_id, long long pad_id, float mask_prob, float random_token_prob, float unchanged_token_prob,
                 int64_t* out, uint8_t* mask, int64_t* labels, void* stream);

/* ---- exchange step of the channel partition over NVLink peer memory (hy_exchange.cu) -------------------------------
 * ONE 1 M-nt sequence split over the GPUs of a box (BASELINE.json configs[3]): per-position
 * layers run on a rank's sequence chunk, the operator core on its channel slab; the reference has no counterpart (it
 * trains with Lightning DDP only, train.py:630-639) — the sites this sits between are in_proj hyena.py:441 and
 * out_proj :504.  Every rank owns one peer-mappable buffer [flag block | payload ...] (hy_peer_alloc), the others map
 * it through its 64-byte CUDA IPC handle (hy_peer_open).  hy_peer_pull copies, for every peer j, n_outer x n_inner
 * rows of row_bytes from  src[j] + src_base + o*src_outer + i*src_inner  to  dst + o*dst_outer + i*dst_inner +
 * j*dst_peer  (reduce = 1: dst rows = the fp32 sum over peers, dst_peer ignored).  `epoch` must increase by one per
 * call on every rank alike; a rank may overwrite the payload it exposed at epoch e only after hy_peer_wait_done(e). */
typedef struct hy_peer_pull_args {
  const void* src[8];          /* peer j's payload (this process's mapping; src[self] is the local buffer) */
  void* flags[8];              /* peer j's flag block (first hy_peer_flag_bytes() bytes of its allocation) */
  int G, self;
  unsigned epoch;
  int reduce;
  int n_outer, n_inner;
  long long row_bytes;
  long long src_base, src_outer, src_inner;
  long long dst_outer, dst_inner, dst_peer;
  void* dst;
  int max_ctas;                /* 0 = one CTA per SM; > 0: at most this many (exchanges overlapped with compute) */
} hy_peer_pull_args;
size_t hy_peer_flag_bytes(void);
int hy_peer_alloc(size_t bytes, void** ptr, void* handle64);
int hy_peer_open(const void* handle64, void** ptr);
int hy_peer_close(void* ptr);
int hy_peer_free(void* ptr);
int hy_peer_pull(const hy_peer_pull_args* p, void* stream);
int hy_peer_wait_done(void* my_flags, int G, int self, unsigned epoch, void* stream);
int hy_peer_error(void* my_flags);

#ifdef __cplusplus
}
#endif
#endif /* HYENA_B200_H_ */
