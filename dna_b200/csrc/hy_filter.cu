// hyena-b200: implicit long-filter generation, fused.
// Restates HyenaFilter.filter (/root/reference/src/models/sequence/hyena.py:233-242) with its parts
// PositionalEmbedding (:113-135, tables are inputs: they may be learned / loaded), the MLP
// Linear(emb,order) -> Sin -> [Linear(order,order) -> Sin] x n_inner -> Linear(order,D,bias=False)
// (:203-219; ONE shared Sin.freq, :100-110) and ExponentialModulation (:138-159):
//     k[c][t] = (W_out h_last(t))[c] * (exp(-t_t * |delta_c|) + shift)
// One CTA owns a tile of 64 positions; hidden activations never leave shared memory; the output is
// written channel-major [D][L] — the layout the spectrum kernel reads — so the reference's
// `rearrange(k, 'l d -> d l')` (hyena.py:460) costs nothing.
#include "hy_host.h"

namespace hy {

constexpr int kFT = 64;    // positions per tile
constexpr int kFO = 64;    // padded MLP width
constexpr int kFThreads = 256;

struct FilterDev {
  int L, D, order, emb_dim, n_inner;
  const float* z; int ldz;
  const float* t;
  const float* w_in; const float* b_in;
  const float* w_h; const float* b_h;
  const float* w_out;
  const float* freq;
  const float* deltas;
  float shift; int modulate;
};

// acc[a][b] += sum_i WT[i][4*jg + a] * h[i][4*pg + b]
HY_DEVICE void tile_gemm_4x4(const float* WT, const float* h, int n_in, int jg, int pg, float (&acc)[4][4]) {
  for (int i = 0; i < n_in; ++i) {
    const float4 w = *reinterpret_cast<const float4*>(WT + i * kFO + 4 * jg);
    const float4 x = *reinterpret_cast<const float4*>(h + i * kFT + 4 * pg);
    const float wv[4] = {w.x, w.y, w.z, w.w};
    const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(wv[a], xv[b], acc[a][b]);
  }
}

__global__ void __launch_bounds__(kFThreads) k_filter_fwd(FilterDev a, float* __restrict__ k, int ldk) {
  HY_DYN_SMEM(float, sm);
  float* WT = sm;                    // [kFO (in)][kFO (out)]
  float* hA = WT + kFO * kFO;        // [kFO][kFT]
  float* hB = hA + kFO * kFT;        // [kFO][kFT]
  float* bias = hB + kFO * kFT;      // [kFO]
  float* fr = bias + kFO;            // [kFO]
  float* tt = fr + kFO;              // [kFT]
  const int tid = threadIdx.x;
  const int pg = tid % 16, jg = tid / 16;
  const int t0 = blockIdx.x * kFT;
  const int O = a.order, E = a.emb_dim;

  // stage: z tile (as the "previous layer" activations), W_in^T, b_in, freq, t
  for (int i = tid; i < kFO * kFT; i += kFThreads) {
    const int e = i / kFT, p = i % kFT;
    hA[i] = (e < E && t0 + p < a.L) ? a.z[(long long)(t0 + p) * a.ldz + e] : 0.f;
  }
  for (int i = tid; i < kFO * kFO; i += kFThreads) {
    const int e = i / kFO, j = i % kFO;
    WT[i] = (e < E && j < O) ? a.w_in[j * E + e] : 0.f;
  }
  if (tid < kFO) {
    bias[tid] = tid < O ? a.b_in[tid] : 0.f;
    fr[tid] = tid < O ? a.freq[tid] : 0.f;
  }
  if (tid < kFT) tt[tid] = (t0 + tid < a.L) ? a.t[t0 + tid] : 0.f;
  __syncthreads();

  float* hin = hA;
  float* hout = hB;
  for (int layer = 0; layer <= a.n_inner; ++layer) {
    float acc[4][4];
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y) acc[x][y] = bias[4 * jg + x];
    tile_gemm_4x4(WT, hin, layer == 0 ? E : O, jg, pg, acc);
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const float f = fr[4 * jg + x];
      float4 v = make_float4(sinf(f * acc[x][0]), sinf(f * acc[x][1]), sinf(f * acc[x][2]), sinf(f * acc[x][3]));
      if (4 * jg + x >= O) v = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(hout + (4 * jg + x) * kFT + 4 * pg) = v;
    }
    __syncthreads();
    // next layer's weights
    if (layer < a.n_inner) {
      const float* W = a.w_h + (long long)layer * O * O;
      for (int i = tid; i < kFO * kFO; i += kFThreads) {
        const int e = i / kFO, j = i % kFO;
        WT[i] = (e < O && j < O) ? W[j * O + e] : 0.f;
      }
      if (tid < kFO) bias[tid] = tid < O ? a.b_h[layer * O + tid] : 0.f;
    }
    float* tmp = hin; hin = hout; hout = tmp;
    __syncthreads();
  }
  // hin now holds h_last [order][kFT]; output layer in chunks of kFO channels
  for (int c0 = 0; c0 < a.D; c0 += kFO) {
    for (int i = tid; i < kFO * kFO; i += kFThreads) {
      const int e = i / kFO, j = i % kFO;
      WT[i] = (e < O && c0 + j < a.D) ? a.w_out[(long long)(c0 + j) * O + e] : 0.f;
    }
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y) acc[x][y] = 0.f;
    tile_gemm_4x4(WT, hin, O, jg, pg, acc);
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const int c = c0 + 4 * jg + x;
      if (c >= a.D) continue;
      float r[4];
      if (a.modulate) {
        const float ad = fabsf(a.deltas[c]);
#pragma unroll
        for (int y = 0; y < 4; ++y) r[y] = acc[x][y] * (expf(-tt[4 * pg + y] * ad) + a.shift);
      } else {
#pragma unroll
        for (int y = 0; y < 4; ++y) r[y] = acc[x][y];
      }
      float* dst = k + (long long)c * ldk + t0 + 4 * pg;
      if (t0 + 4 * pg + 3 < a.L && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
        *reinterpret_cast<float4*>(dst) = make_float4(r[0], r[1], r[2], r[3]);
      } else {
#pragma unroll
        for (int y = 0; y < 4; ++y)
          if (t0 + 4 * pg + y < a.L) dst[y] = r[y];
      }
    }
    __syncthreads();
  }
}

// ---- fused backward of the MLP trunk -------------------------------------------------------------
// Given dh_last [L][order] (gradient wrt the last hidden activation, produced by one cuBLAS GEMM from the
// kernel below), recompute the trunk per 64-position tile and accumulate, per CTA and in a fixed order,
// dW_in, db_in, dW_h[*], db_h[*], dfreq.  Nothing is read back from the forward; activations live in shared
// memory in both i-major (for the W h products) and position-major (for the outer products) layouts.
constexpr int kFE = 8;       // max emb_dim of the fused backward
constexpr int kFL = 3;       // max number of Linear+Sin layers (1 + n_inner): HyenaDNA uses 3
constexpr int kDS = 68;      // padded row stride of the staged dh tile (spreads the transposing stores over banks)

// acc[a][b] += sum_r A[r][4*ag + a] * B[r][4*bg + b]   (reduction index major, row strides lda / ldb floats)
HY_DEVICE void tile_gemm_rmajor(const float* A, int lda, const float* B, int ldb, int nred, int ag, int bg, float (&acc)[4][4]) {
  for (int r = 0; r < nred; ++r) {
    const float4 x = *reinterpret_cast<const float4*>(A + r * lda + 4 * ag);
    const float4 y = *reinterpret_cast<const float4*>(B + r * ldb + 4 * bg);
    const float xv[4] = {x.x, x.y, x.z, x.w};
    const float yv[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(xv[a], yv[b], acc[a][b]);
  }
}

HY_DEVICE float half_warp_sum(float v) {   // sum over the 16 lanes sharing tid / 16
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// part layout per CTA: [dW_in O*E][db_in O][for l in 1..n_inner: dW_h O*O, db_h O][dfreq O]
template <int NL>
__global__ void __launch_bounds__(kFThreads) k_filter_trunk_bwd(FilterDev a, const float* __restrict__ dh, int lddh,
                                                              float* __restrict__ part, int part_stride) {
  HY_DYN_SMEM(float, sm);
  float* Wb = sm;                                  // [kFO][kFO] weight staging (transposed or natural)
  float* z_i = Wb + kFO * kFO;                     // [kFE][kFT]   z, feature-major
  float* z_p = z_i + kFE * kFT;                    // [kFT][kFE]   z, position-major
  float* h_i = z_p + kFT * kFE;                    // [kFL-1][kFO][kFT] inputs of layers 1.. (feature-major)
  float* h_p = h_i + (kFL - 1) * kFO * kFT;        // [kFL-1][kFT][kFO] same, position-major
  float* a_s = h_p + (kFL - 1) * kFO * kFT;        // [kFL][kFO][kFT] pre-activations
  float* d_s = a_s + kFL * kFO * kFT;              // [kFO][kDS] gradient wrt the current activation (feature-major)
  float* da_j = d_s + kFO * kDS;                   // [kFO][kFT] gradient wrt pre-activation, feature-major
  float* da_p = da_j + kFO * kFT;                  // [kFT][kFO] same, position-major
  float* fr = da_p + kFO * kFT;                    // [kFO]
  float* bs = fr + kFO;                            // [kFL][kFO] biases
  const int tid = threadIdx.x;
  const int lo = tid % 16, hi = tid / 16;          // (pg | ig, jg)
  const int O = a.order, E = a.emb_dim;
  if (tid < kFO) fr[tid] = tid < O ? a.freq[tid] : 0.f;
  for (int i = tid; i < kFL * kFO; i += kFThreads) {
    const int l = i / kFO, j = i % kFO;
    float v = 0.f;
    if (j < O && l < NL) v = (l == 0) ? a.b_in[j] : a.b_h[(l - 1) * O + j];
    bs[i] = v;
  }
  // persistent accumulators
  float accW[kFL - 1][4][4];    // dW_h[l-1][4*hi + a][4*lo + b]
  float accWin[2];              // dW_in[j = tid / 4][e = (tid % 4) * 2 + {0, 1}]
  float accB[kFL][4];           // db_l[4*hi + a]  (valid on lanes with lo == 0)
  float accF[4];                // dfreq[4*hi + a]
#pragma unroll
  for (int l = 0; l < kFL - 1; ++l)
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y) accW[l][x][y] = 0.f;
  accWin[0] = accWin[1] = 0.f;
#pragma unroll
  for (int l = 0; l < kFL; ++l)
#pragma unroll
    for (int x = 0; x < 4; ++x) accB[l][x] = 0.f;
#pragma unroll
  for (int x = 0; x < 4; ++x) accF[x] = 0.f;
  __syncthreads();

  const int ntiles = (a.L + kFT - 1) / kFT;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int t0 = tile * kFT;
    // ---- stage z (both layouts) and dh (feature-major)
    for (int i = tid; i < kFE * kFT; i += kFThreads) {
      const int e = i / kFT, p = i % kFT;
      const float v = (e < E && t0 + p < a.L) ? a.z[(long long)(t0 + p) * a.ldz + e] : 0.f;
      z_i[e * kFT + p] = v;
      z_p[p * kFE + e] = v;
    }
    for (int i = tid; i < kFO * kFT; i += kFThreads) {
      const int p = i / kFO, j = i % kFO;     // coalesced along j
      d_s[j * kDS + p] = (j < O && t0 + p < a.L) ? dh[(long long)(t0 + p) * lddh + j] : 0.f;
    }
    // ---- forward recompute (NL is a template parameter: every accumulator index below is a compile-time constant)
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const float* W = (l == 0) ? a.w_in : a.w_h + (long long)(l - 1) * O * O;
      const int nin = (l == 0) ? E : O;
      __syncthreads();
      for (int i = tid; i < kFO * kFO; i += kFThreads) {   // Wb[i_in][j_out] = W[j_out][i_in]
        const int e = i / kFO, j = i % kFO;
        Wb[i] = (e < nin && j < O) ? W[j * nin + e] : 0.f;
      }
      __syncthreads();
      float acc[4][4];
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = bs[l * kFO + 4 * hi + x];
      const float* hin = (l == 0) ? z_i : h_i + (l - 1) * kFO * kFT;
      tile_gemm_rmajor(Wb, kFO, hin, kFT, nin, hi, lo, acc);     // acc[x][y] = a_l[4hi+x][4lo+y]
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const int j = 4 * hi + x;
        *reinterpret_cast<float4*>(a_s + (l * kFO + j) * kFT + 4 * lo) = make_float4(acc[x][0], acc[x][1], acc[x][2], acc[x][3]);
        if (l + 1 < NL) {
          const float f = fr[j];
          float hv[4];
#pragma unroll
          for (int y = 0; y < 4; ++y) hv[y] = (j < O) ? sinf(f * acc[x][y]) : 0.f;
          *reinterpret_cast<float4*>(h_i + (l * kFO + j) * kFT + 4 * lo) = make_float4(hv[0], hv[1], hv[2], hv[3]);
#pragma unroll
          for (int y = 0; y < 4; ++y) h_p[(l * kFT + 4 * lo + y) * kFO + j] = hv[y];
        }
      }
    }
    __syncthreads();
    // ---- backward
#pragma unroll
    for (int l = NL - 1; l >= 0; --l) {
      // da = d * f * cos(f a);  dfreq += d * cos(f a) * a;  db += da      (thread owns j = 4hi+x, p = 4lo+y)
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const int j = 4 * hi + x;
        const float f = fr[j];
        const float4 dv = *reinterpret_cast<const float4*>(d_s + j * kDS + 4 * lo);
        const float4 av = *reinterpret_cast<const float4*>(a_s + (l * kFO + j) * kFT + 4 * lo);
        const float dd[4] = {dv.x, dv.y, dv.z, dv.w};
        const float aa[4] = {av.x, av.y, av.z, av.w};
        float g[4], sb = 0.f, sf = 0.f;
#pragma unroll
        for (int y = 0; y < 4; ++y) {
          const float c = cosf(f * aa[y]);
          const float dc = (j < O && t0 + 4 * lo + y < a.L) ? dd[y] * c : 0.f;
          g[y] = dc * f;
          sb += g[y];
          sf += dc * aa[y];
        }
        *reinterpret_cast<float4*>(da_j + j * kFT + 4 * lo) = make_float4(g[0], g[1], g[2], g[3]);
#pragma unroll
        for (int y = 0; y < 4; ++y) da_p[(4 * lo + y) * kFO + j] = g[y];
        sb = half_warp_sum(sb);
        sf = half_warp_sum(sf);
        accB[l][x] += sb;
        accF[x] += sf;
      }
      __syncthreads();
      // dW_l[j][i] += sum_p da[j][p] * hin_l[i][p]   (position-major operands; thread owns j = 4hi+x, i = 4lo+y)
      if (l > 0) {
        tile_gemm_rmajor(da_p, kFO, h_p + (l - 1) * kFT * kFO, kFO, kFT, hi, lo, accW[l - 1]);
      } else {
        const int j = tid / 4, e0 = (tid % 4) * 2;
        float s0 = 0.f, s1 = 0.f;
        for (int p = 0; p < kFT; ++p) {
          const float g = da_p[p * kFO + j];
          s0 = fmaf(g, z_p[p * kFE + e0], s0);
          s1 = fmaf(g, z_p[p * kFE + e0 + 1], s1);
        }
        accWin[0] += s0;
        accWin[1] += s1;
      }
      // d_prev[i][p] = sum_j W_l[j][i] * da[j][p]    (reduction over j: natural W layout)
      if (l > 0) {
        const float* W = a.w_h + (long long)(l - 1) * O * O;
        for (int i = tid; i < kFO * kFO; i += kFThreads) {     // Wb[j][i] = W[j][i]
          const int j = i / kFO, ii = i % kFO;
          Wb[i] = (j < O && ii < O) ? W[j * O + ii] : 0.f;
        }
        __syncthreads();
        float acc[4][4];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
          for (int y = 0; y < 4; ++y) acc[x][y] = 0.f;
        tile_gemm_rmajor(Wb, kFO, da_j, kFT, O, hi, lo, acc);    // acc[x][y] = d_prev[i = 4hi+x][p = 4lo+y]
        __syncthreads();                                         // everyone is done reading d_s / da_*
#pragma unroll
        for (int x = 0; x < 4; ++x)
          *reinterpret_cast<float4*>(d_s + (4 * hi + x) * kDS + 4 * lo) = make_float4(acc[x][0], acc[x][1], acc[x][2], acc[x][3]);
        __syncthreads();
      }
    }
    __syncthreads();
  }
  // ---- write this CTA's partial sums
  float* out = part + (long long)blockIdx.x * part_stride;
  {
    const int j = tid / 4, e0 = (tid % 4) * 2;
    if (j < O && e0 < E) out[j * E + e0] = accWin[0];
    if (j < O && e0 + 1 < E) out[j * E + e0 + 1] = accWin[1];
  }
  int off = O * E;
  if (lo == 0) {
#pragma unroll
    for (int x = 0; x < 4; ++x)
      if (4 * hi + x < O) out[off + 4 * hi + x] = accB[0][x];
  }
  off += O;
#pragma unroll
  for (int l = 1; l < NL; ++l) {
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y) {
        const int j = 4 * hi + x, i = 4 * lo + y;
        if (j < O && i < O) out[off + j * O + i] = accW[l - 1][x][y];
      }
    off += O * O;
    if (lo == 0) {
#pragma unroll
      for (int x = 0; x < 4; ++x)
        if (4 * hi + x < O) out[off + 4 * hi + x] = accB[l][x];
    }
    off += O;
  }
  if (lo == 0) {
#pragma unroll
    for (int x = 0; x < 4; ++x)
      if (4 * hi + x < O) out[off + 4 * hi + x] = accF[x];
  }
}

// dh[t][c] = dk[c][t] * (exp(-t_t |delta_c|) + shift): the modulation's backward fused with the
// [D][L] -> [L][D] transpose the MLP's GEMMs want (32x32 tiles through shared memory, both sides coalesced).
__global__ void __launch_bounds__(256) k_filter_modulate_bwd(const float* __restrict__ dk, int lddk, const float* __restrict__ t,
                                                           const float* __restrict__ deltas, float shift, int modulate,
                                                           float* __restrict__ dh, int ldh, int L, int D) {
  HY_STATIC_SMEM(float, tile, 32 * 33);
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;   // 32 x 8
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, tt = t0 + tx;
    tile[(ty + 8 * i) * 33 + tx] = (c < D && tt < L) ? dk[(long long)c * lddk + tt] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int tt = t0 + ty + 8 * i, c = c0 + tx;
    if (tt < L && c < D) {
      float v = tile[tx * 33 + ty + 8 * i];
      if (modulate) v *= expf(-t[tt] * fabsf(deltas[c])) + shift;
      dh[(long long)tt * ldh + c] = v;
    }
  }
}

}  // namespace hy

using namespace hy;

extern "C" int hy_filter_modulate_bwd(const float* dk, int lddk, const float* t, const float* deltas, float shift, int modulate,
                                      float* dh, int ldh, int L, int D, void* stream) {
  if (!dk || !dh || !t || (modulate && !deltas) || L < 1 || D < 1) return fail(HY_ERR_ARG, "hy_filter_modulate_bwd: bad argument");
  const dim3 grid((L + 31) / 32, (D + 31) / 32);
  HY_LAUNCH(k_filter_modulate_bwd, grid, 256, 0, stream, dk, lddk, t, deltas, shift, modulate, dh, ldh, L, D);
  return check_launch("k_filter_modulate_bwd");
}

extern "C" int hy_filter_trunk_bwd_layout(const hy_filter_args* p, int* n_cta, int* stride) {
  if (!p || !n_cta || !stride) return fail(HY_ERR_ARG, "hy_filter_trunk_bwd_layout: bad argument");
  const int ntiles = (p->L + kFT - 1) / kFT;
  *n_cta = ntiles < 148 ? ntiles : 148;
  *stride = p->order * p->emb_dim + p->order + p->n_inner * (p->order * p->order + p->order) + p->order;
  return HY_OK;
}

extern "C" int hy_filter_trunk_bwd(const hy_filter_args* p, const float* dh, int lddh, float* part, void* stream) {
  if (!p || !dh || !part || !p->z || !p->w_in || !p->b_in || !p->freq || p->L < 1)
    return fail(HY_ERR_ARG, "hy_filter_trunk_bwd: bad argument");
  if (p->order < 1 || p->order > kFO || p->emb_dim < 1 || p->emb_dim > kFE || p->n_inner < 0 || p->n_inner > kFL - 1)
    return fail(HY_ERR_UNSUPPORTED, "hy_filter_trunk_bwd: order %d (<= %d) / emb_dim %d (<= %d) / n_inner %d (<= %d) outside the fused range",
                p->order, kFO, p->emb_dim, kFE, p->n_inner, kFL - 1);
  if (p->n_inner > 0 && (!p->w_h || !p->b_h)) return fail(HY_ERR_ARG, "hy_filter_trunk_bwd: hidden layers need w_h and b_h");
  FilterDev a;
  a.L = p->L; a.D = p->D; a.order = p->order; a.emb_dim = p->emb_dim; a.n_inner = p->n_inner;
  a.z = p->z; a.ldz = p->ldz; a.t = p->t;
  a.w_in = p->w_in; a.b_in = p->b_in; a.w_h = p->w_h; a.b_h = p->b_h; a.w_out = p->w_out;
  a.freq = p->freq; a.deltas = p->deltas; a.shift = p->shift; a.modulate = p->modulate;
  int n_cta = 0, stride = 0;
  hy_filter_trunk_bwd_layout(p, &n_cta, &stride);
  const size_t smem = sizeof(float) * (kFO * kFO + 2 * kFE * kFT + 2 * (kFL - 1) * kFO * kFT + kFL * kFO * kFT + kFO * kDS +
                                       2 * kFO * kFT + kFO + kFL * kFO);
  if (p->n_inner == 0) {
    auto kern = k_filter_trunk_bwd<1>;
    HY_LAUNCH(kern, n_cta, kFThreads, smem, stream, a, dh, lddh, part, stride);
  } else if (p->n_inner == 1) {
    auto kern = k_filter_trunk_bwd<2>;
    HY_LAUNCH(kern, n_cta, kFThreads, smem, stream, a, dh, lddh, part, stride);
  } else {
    auto kern = k_filter_trunk_bwd<3>;
    HY_LAUNCH(kern, n_cta, kFThreads, smem, stream, a, dh, lddh, part, stride);
  }
  return check_launch("k_filter_trunk_bwd");
}

extern "C" int hy_filter_fwd(const hy_filter_args* p, float* k, int ldk, void* stream) {
  if (!p || !k || !p->z || !p->t || !p->w_in || !p->b_in || !p->w_out || !p->freq || p->L < 1 || p->D < 1)
    return fail(HY_ERR_ARG, "hy_filter_fwd: bad argument");
  if (p->order < 1 || p->order > kFO || p->emb_dim < 1 || p->emb_dim > kFO)
    return fail(HY_ERR_UNSUPPORTED, "hy_filter_fwd: order %d / emb_dim %d outside [1, %d]", p->order, p->emb_dim, kFO);
  if (p->n_inner > 0 && (!p->w_h || !p->b_h)) return fail(HY_ERR_ARG, "hy_filter_fwd: hidden layers need w_h and b_h");
  if (p->modulate && !p->deltas) return fail(HY_ERR_ARG, "hy_filter_fwd: modulation needs deltas");
  FilterDev a;
  a.L = p->L; a.D = p->D; a.order = p->order; a.emb_dim = p->emb_dim; a.n_inner = p->n_inner;
  a.z = p->z; a.ldz = p->ldz; a.t = p->t;
  a.w_in = p->w_in; a.b_in = p->b_in; a.w_h = p->w_h; a.b_h = p->b_h; a.w_out = p->w_out;
  a.freq = p->freq; a.deltas = p->deltas; a.shift = p->shift; a.modulate = p->modulate;
  const size_t smem = sizeof(float) * (kFO * kFO + 2 * kFO * kFT + 2 * kFO + kFT);
  HY_LAUNCH(k_filter_fwd, (p->L + kFT - 1) / kFT, kFThreads, smem, stream, a, k, ldk);
  return check_launch("k_filter_fwd");
}
