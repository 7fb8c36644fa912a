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
