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
#include "hy_filter.h"

namespace hy {

constexpr int kFThreads = 256;

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

__global__ void __launch_bounds__(kFThreads) k_filter_fwd_generic(FilterDev a, float* __restrict__ k, int ldk) {
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
// dW_in, db_in, dW_h[*], db_h[*], dfreq.  Nothing is read back from the forward.  Persistent CTAs (one per SM):
// the weights are staged ONCE per CTA in their natural [out][in] layout, which serves both the forward products
// (rows read along `in`) and the d_prev = W^T da products (rows read along `out`); pre-activations stay in
// registers between the recompute and the backward (the thread owns the same (feature, position) block in
// both); the next tile's dh / z are fetched into registers while the current tile computes.
constexpr int kFE = 8;       // max emb_dim of the fused backward
constexpr int kFL = 3;       // max number of Linear+Sin layers (1 + n_inner): HyenaDNA uses 3
constexpr int kLDW = 68;     // row stride (floats) of the [64][64] shared-memory matrices: float4-aligned, and rows
                             // 4 apart land 16 banks apart so the two row groups of a warp do not collide

// row stride of the z tile: the dW_in products read rows e0 = 0, 2, 4, 6 (+1) at the same column from one warp — with a
// stride of kFT (256 B) all four hit the same banks (ncu: 4x excess wavefronts, a quarter of the kernel's shared-memory
// traffic); 68 floats puts them 16 B apart
constexpr int kZS = kFT + 4;
constexpr int kTrunkSmemFloats = (kFL - 1) * kFO * kLDW + kFO * kFE + kFE * kZS + (kFL - 1) * kFO * kLDW + 2 * kFO * kLDW +
                                 kFO + kFL * kFO;

HY_DEVICE float half_warp_sum(float v) {   // sum over the 16 lanes sharing tid / 16
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
HY_DEVICE float f4c(const float4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
HY_DEVICE float dot4(const float4& a, const float4& b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))); }

// dh[t][j0 .. j0+3] with zeros outside [0, L) x [0, O)
HY_DEVICE float4 trunk_ld_dh(const float* __restrict__ dh, int lddh, int L, int O, int t, int j0, bool vec) {
  if (t >= L) return make_float4(0.f, 0.f, 0.f, 0.f);
  const float* p = dh + (long long)t * lddh + j0;
  if (vec && j0 + 3 < O) return *reinterpret_cast<const float4*>(p);
  return make_float4(j0 < O ? p[0] : 0.f, j0 + 1 < O ? p[1] : 0.f, j0 + 2 < O ? p[2] : 0.f, j0 + 3 < O ? p[3] : 0.f);
}

// part layout per CTA: [dW_in O*E][db_in O][for l in 1..n_inner: dW_h O*O, db_h O][dfreq O]
// SAVED: the forward kept the pre-activations (FilterDev::asave): no recompute — the tile's a_l arrive by cp.async
// (double buffered, the next tile's in flight during this tile's backward), h_l = sin(f a_l) is rebuilt from them.
template <int NL, int MINB = 1, bool SAVED = false>
__global__ void __launch_bounds__(kFThreads, MINB) k_filter_trunk_bwd(FilterDev a, const float* __restrict__ dh, int lddh,
                                                                    float* __restrict__ part, int part_stride) {
  HY_DYN_SMEM(float, sm);
  float* Wn = sm;                                   // [kFL-1][kFO][kLDW]  W_h[l][out][in]
  float* Win = Wn + (kFL - 1) * kFO * kLDW;         // [kFO][kFE]          W_in[out][e]
  float* z_i = Win + kFO * kFE;                     // [kFE][kZS]          z tile, feature-major
  float* h_s = z_i + kFE * kZS;                     // [kFL-1][kFO][kLDW]  h_l = sin(f a_l), [feature][position]
  float* d_s = h_s + (kFL - 1) * kFO * kLDW;        // [kFO][kLDW]         gradient wrt h_l (layers below the top)
  float* da_s = d_s + kFO * kLDW;                   // [kFO][kLDW]         gradient wrt a_l
  float* fr = da_s + kFO * kLDW;                    // [kFO]
  float* bs = fr + kFO;                             // [kFL][kFO]
  float* abuf = bs + kFL * kFO;                     // SAVED: [2][kFL][kFO][kFT] pre-activation tiles
  const int tid = threadIdx.x;
  const int lo = tid % 16, hi = tid / 16;
  const int O = a.order, E = a.emb_dim;
  const int O4 = (O + 3) & ~3;
  constexpr int kATile = kFL * kFO * kFT;
  // cp.async of one tile's a_l rows: NL * kFO rows of kFT floats (lda is a multiple of kFT: every tile lies inside)
  auto fetch_a = [&](int tile, int buf) {
    const float* src = a.asave + (long long)tile * kFT;
    float* dst = abuf + buf * kATile;
    for (int i = tid; i < NL * kFO * (kFT / 4); i += kFThreads) {
      const int r = i / (kFT / 4), q = i % (kFT / 4);
      hy_cp_async16(dst + r * kFT + 4 * q, src + (long long)r * a.lda + 4 * q);
    }
  };
  for (int i = tid; i < (NL - 1) * kFO * kFO; i += kFThreads) {
    const int l = i / (kFO * kFO), j = (i / kFO) % kFO, ii = i % kFO;
    Wn[(l * kFO + j) * kLDW + ii] = (j < O && ii < O) ? a.w_h[(long long)l * O * O + j * O + ii] : 0.f;
  }
  for (int i = tid; i < kFO * kFE; i += kFThreads) {
    const int j = i / kFE, e = i % kFE;
    Win[i] = (j < O && e < E) ? a.w_in[j * E + e] : 0.f;
  }
  if (tid < kFO) fr[tid] = tid < O ? a.freq[tid] : 0.f;
  for (int i = tid; i < kFL * kFO; i += kFThreads) {
    const int l = i / kFO, j = i % kFO;
    float v = 0.f;
    if (j < O && l < NL) v = (l == 0) ? a.b_in[j] : a.b_h[(l - 1) * O + j];
    bs[i] = v;
  }
  // persistent per-thread partial sums (reduced over the half-warp / written once at the end)
  float accW[NL > 1 ? NL - 1 : 1][4][4];   // dW_h[l-1][4*hi + x][lo + 16*y]
  float accWin[2];                         // dW_in[tid / 4][(tid % 4) * 2 + {0, 1}]
  float accB[NL][4];                       // db_l[4*hi + x], this lane's 4 positions
  float accF[4];                           // dfreq[4*hi + x]
#pragma unroll
  for (int l = 0; l < (NL > 1 ? NL - 1 : 1); ++l)
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y) accW[l][x][y] = 0.f;
  accWin[0] = accWin[1] = 0.f;
#pragma unroll
  for (int l = 0; l < NL; ++l)
#pragma unroll
    for (int x = 0; x < 4; ++x) accB[l][x] = 0.f;
#pragma unroll
  for (int x = 0; x < 4; ++x) accF[x] = 0.f;

  const bool vec = (lddh % 4 == 0) && ((reinterpret_cast<uintptr_t>(dh) & 15) == 0);
  const int ntiles = (a.L + kFT - 1) / kFT;
  // register prefetch of the tile's inputs: dpre[y] = dh[t0 + 4lo + y][4hi .. 4hi+3]; zpre[k] = z element tid + 256k
  float4 dpre[4];
  float zpre[2];
  auto fetch = [&](int tile) {
    const int t0 = tile * kFT;
#pragma unroll
    for (int y = 0; y < 4; ++y) dpre[y] = trunk_ld_dh(dh, lddh, a.L, O, t0 + 4 * lo + y, 4 * hi, vec);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = tid + k * kFThreads, p = i / kFE, e = i % kFE;
      zpre[k] = (e < E && t0 + p < a.L) ? a.z[(long long)(t0 + p) * a.ldz + e] : 0.f;
    }
  };
  if ((int)blockIdx.x < ntiles) {
    fetch(blockIdx.x);
    if (SAVED) fetch_a(blockIdx.x, 0);
  }
  __syncthreads();

  int abuf_i = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, abuf_i ^= 1) {
    const float* acur = abuf + abuf_i * kATile;
    float4 dtop[4];
#pragma unroll
    for (int y = 0; y < 4; ++y) dtop[y] = dpre[y];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = tid + k * kFThreads, p = i / kFE, e = i % kFE;
      z_i[e * kZS + p] = zpre[k];
    }
    if (SAVED) hy_cp_async_wait_all();   // this tile's a_l (requested one tile ago)
    if (tile + (int)gridDim.x < ntiles) fetch(tile + gridDim.x);
    __syncthreads();
    if (SAVED && tile + (int)gridDim.x < ntiles) fetch_a(tile + gridDim.x, abuf_i ^ 1);   // lands during this tile

    // ---- forward recompute; a_l stays in registers: areg[l][x][y] = a_l[4hi + x][4lo + y]
    float areg[SAVED ? 1 : NL][4][4];
    float creg[(SAVED && NL > 1) ? NL - 1 : 1][4][4];   // SAVED: cos(f a_l) of the layers below the top
    if (SAVED) {
      // h_l = sin(f a_l) for the layers below the top, from the saved pre-activations; the cosine the backward of the
      // same layer needs comes out of the same range reduction
#pragma unroll
      for (int l = 0; l + 1 < NL; ++l) {
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          const float f = fr[4 * hi + x];
          const float4 av = *reinterpret_cast<const float4*>(acur + (l * kFO + 4 * hi + x) * kFT + 4 * lo);
          float sn[4];
          const float fa[4] = {f * av.x, f * av.y, f * av.z, f * av.w};
          hy_sincos4(fa, sn, creg[l][x]);
          *reinterpret_cast<float4*>(h_s + (l * kFO + 4 * hi + x) * kLDW + 4 * lo) = make_float4(sn[0], sn[1], sn[2], sn[3]);
        }
      }
      if (NL > 1) __syncthreads();
    }
#pragma unroll
    for (int l = 0; l < (SAVED ? 0 : NL); ++l) {
      float acc[4][4];
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = bs[l * kFO + 4 * hi + x];
      if (l == 0) {
        float4 zz[kFE];
#pragma unroll
        for (int e = 0; e < kFE; ++e) zz[e] = *reinterpret_cast<const float4*>(z_i + e * kZS + 4 * lo);
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          const float4 wa = *reinterpret_cast<const float4*>(Win + (4 * hi + x) * kFE);
          const float4 wb = *reinterpret_cast<const float4*>(Win + (4 * hi + x) * kFE + 4);
#pragma unroll
          for (int e = 0; e < kFE; ++e) {
            const float w = e < 4 ? f4c(wa, e) : f4c(wb, e - 4);
            fma4s(acc[x], w, zz[e]);
          }
        }
      } else {
        const float* W = Wn + (l - 1) * kFO * kLDW;
        const float* H = h_s + (l - 1) * kFO * kLDW;
#pragma unroll 2
        for (int i4 = 0; i4 < O4; i4 += 4) {
          float4 w[4], h[4];
#pragma unroll
          for (int x = 0; x < 4; ++x) w[x] = *reinterpret_cast<const float4*>(W + (4 * hi + x) * kLDW + i4);
#pragma unroll
          for (int r = 0; r < 4; ++r) h[r] = *reinterpret_cast<const float4*>(H + (i4 + r) * kLDW + 4 * lo);
#pragma unroll
          for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int r = 0; r < 4; ++r)
              fma4s(acc[x], f4c(w[x], r), h[r]);
        }
      }
#pragma unroll
      for (int x = 0; x < 4; ++x) {
#pragma unroll
        for (int y = 0; y < 4; ++y) areg[SAVED ? 0 : l][x][y] = acc[x][y];
        if (l + 1 < NL) {
          const float f = fr[4 * hi + x];
          const float fa[4] = {f * acc[x][0], f * acc[x][1], f * acc[x][2], f * acc[x][3]};
          float sn[4];
          hy_sin4(fa, sn);
          *reinterpret_cast<float4*>(h_s + (l * kFO + 4 * hi + x) * kLDW + 4 * lo) = make_float4(sn[0], sn[1], sn[2], sn[3]);
        }
      }
      if (l + 1 < NL) __syncthreads();
    }

    // ---- backward
#pragma unroll
    for (int l = NL - 1; l >= 0; --l) {
      // da = d * f * cos(f a);  dfreq += d * cos(f a) * a;  db += da      (thread owns j = 4hi+x, p = 4lo+y)
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const int j = 4 * hi + x;
        const float f = fr[j];
        float dd[4];
        if (l == NL - 1) {
#pragma unroll
          for (int y = 0; y < 4; ++y) dd[y] = f4c(dtop[y], x);
        } else {
          const float4 dv = *reinterpret_cast<const float4*>(d_s + j * kLDW + 4 * lo);
          dd[0] = dv.x; dd[1] = dv.y; dd[2] = dv.z; dd[3] = dv.w;
        }
        float g[4];
        float4 asv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (SAVED) asv = *reinterpret_cast<const float4*>(acur + (l * kFO + j) * kFT + 4 * lo);
        float avv[4], cvv[4];
#pragma unroll
        for (int y = 0; y < 4; ++y) avv[y] = SAVED ? f4c(asv, y) : areg[SAVED ? 0 : l][x][y];
        if (SAVED && l + 1 < NL) {
#pragma unroll
          for (int y = 0; y < 4; ++y) cvv[y] = creg[(SAVED && l + 1 < NL) ? l : 0][x][y];
        } else {
          const float fa[4] = {f * avv[0], f * avv[1], f * avv[2], f * avv[3]};
          hy_cos4(fa, cvv);
        }
#pragma unroll
        for (int y = 0; y < 4; ++y) {
          const float av = avv[y];
          const float cv = cvv[y];
          const float dc = dd[y] * cv;
          g[y] = dc * f;
          accB[l][x] += g[y];
          accF[x] = fmaf(dc, av, accF[x]);
        }
        *reinterpret_cast<float4*>(da_s + j * kLDW + 4 * lo) = make_float4(g[0], g[1], g[2], g[3]);
      }
      __syncthreads();
      if (l > 0) {
        // dW_l[j][i] += sum_p da[j][p] * h_{l-1}[i][p]    (thread owns j = 4hi+x, i = lo + 16y: conflict-free rows)
        const float* H = h_s + (l - 1) * kFO * kLDW;
#pragma unroll 2
        for (int p4 = 0; p4 < kFT; p4 += 4) {
          float4 A[4], B[4];
#pragma unroll
          for (int x = 0; x < 4; ++x) A[x] = *reinterpret_cast<const float4*>(da_s + (4 * hi + x) * kLDW + p4);
#pragma unroll
          for (int y = 0; y < 4; ++y) B[y] = *reinterpret_cast<const float4*>(H + (lo + 16 * y) * kLDW + p4);
#pragma unroll
          for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y)   // four FMAs into the accumulator (dot4 + add is five operations)
              accW[l - 1][x][y] = fmaf(A[x].w, B[y].w, fmaf(A[x].z, B[y].z, fmaf(A[x].y, B[y].y, fmaf(A[x].x, B[y].x, accW[l - 1][x][y]))));
        }
        // d_prev[i][p] = sum_j W_l[j][i] * da[j][p]       (thread owns i = 4hi+x, p = 4lo+y)
        const float* W = Wn + (l - 1) * kFO * kLDW;
        float acc[4][4];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
          for (int y = 0; y < 4; ++y) acc[x][y] = 0.f;
#pragma unroll 2
        for (int j4 = 0; j4 < O4; j4 += 4) {
          float4 wv[4], dv[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            wv[r] = *reinterpret_cast<const float4*>(W + (j4 + r) * kLDW + 4 * hi);
            dv[r] = *reinterpret_cast<const float4*>(da_s + (j4 + r) * kLDW + 4 * lo);
          }
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int x = 0; x < 4; ++x)
              fma4s(acc[x], f4c(wv[r], x), dv[r]);
        }
        // every read of d_s for this layer happened before the barrier above
#pragma unroll
        for (int x = 0; x < 4; ++x)
          *reinterpret_cast<float4*>(d_s + (4 * hi + x) * kLDW + 4 * lo) = make_float4(acc[x][0], acc[x][1], acc[x][2], acc[x][3]);
      } else {
        // dW_in[j][e] += sum_p da[j][p] * z[e][p]
        const int j = tid / 4, e0 = (tid % 4) * 2;
        float s0 = 0.f, s1 = 0.f;
#pragma unroll 4
        for (int p4 = 0; p4 < kFT; p4 += 4) {
          const float4 g = *reinterpret_cast<const float4*>(da_s + j * kLDW + p4);
          s0 += dot4(g, *reinterpret_cast<const float4*>(z_i + e0 * kZS + p4));
          s1 += dot4(g, *reinterpret_cast<const float4*>(z_i + (e0 + 1) * kZS + p4));
        }
        accWin[0] += s0;
        accWin[1] += s1;
      }
      __syncthreads();   // d_s visible / da_s, h_s, z_i free for the next layer or tile
    }
  }
  // ---- write this CTA's partial sums
  float* out = part + (long long)blockIdx.x * part_stride;
  {
    const int j = tid / 4, e0 = (tid % 4) * 2;
    if (j < O && e0 < E) out[j * E + e0] = accWin[0];
    if (j < O && e0 + 1 < E) out[j * E + e0 + 1] = accWin[1];
  }
  int off = O * E;
#pragma unroll
  for (int l = 0; l < NL; ++l) {
    if (l > 0) {
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) {
          const int j = 4 * hi + x, i = lo + 16 * y;
          if (j < O && i < O) out[off + j * O + i] = accW[l - 1][x][y];
        }
      off += O * O;
    }
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const float v = half_warp_sum(accB[l][x]);
      if (lo == 0 && 4 * hi + x < O) out[off + 4 * hi + x] = v;
    }
    off += O;
  }
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const float v = half_warp_sum(accF[x]);
    if (lo == 0 && 4 * hi + x < O) out[off + 4 * hi + x] = v;
  }
}

// ---- forward, persistent variant for the HyenaDNA shapes (emb_dim <= 8, at most kFL Linear+Sin layers) -------------
// One CTA per SM, 512 threads = two 64-position half-tiles sharing ONE resident copy of every weight matrix (natural
// [out][in] layout, staged once per CTA): per tile only z and t are fetched (register prefetch of the next tile), the
// hidden activations ping-pong between two shared-memory buffers, and the output layer streams 64 channels at a time
// from the resident W_out straight into the modulation and the channel-major store.  gridDim.y splits D into slabs
// of kFwdCh channels (the trunk is recomputed per slab: 1/3 of the work at D = 256, so one slab whenever D <= 256).
constexpr int kFwdCh = 256;                 // channels of W_out resident per CTA
constexpr int kFwdThreads = 2 * kFThreads;
constexpr int kFwdSmemFloats = (kFL - 1) * kFO * kLDW + kFO * kFE + kFwdCh * kLDW + 2 * (kFE * kFT + 2 * kFO * kLDW + kFT) +
                               kFO + kFL * kFO + kFwdCh;

// TF = features per thread (register tile TF x 4 positions): 4 -> 256 threads per half-tile, 8 -> 128.  The kernel is
// bound by shared-memory wavefronts (ncu: LSU data pipe 74 % of peak with TF = 4, FMA pipe 55 %); TF = 8 feeds every
// activation float4 to twice the features — measured SLOWER (2.55 vs 1.59 ms at L = 1 M: 8 warps per SM instead of 16 lose
// more latency hiding than the saved wavefronts buy), so 4 stays the default; hy_debug_set_filter_fwd_tf(8) selects it.
template <int NL, int TF = 4>
__global__ void __launch_bounds__(2 * (kFO * kFT / (TF * 4)), 1) k_filter_fwd_fast(FilterDev a, float* __restrict__ k, int ldk) {
  constexpr int HT = kFO * kFT / (TF * 4);     // threads per half-tile
  constexpr int NTH = 2 * HT;                  // threads per CTA
  constexpr int ZQ = kFE * kFT / HT;           // z elements per thread per tile
  HY_DYN_SMEM(float, sm);
  float* Wn = sm;                                   // [kFL-1][kFO][kLDW]
  float* Win = Wn + (kFL - 1) * kFO * kLDW;         // [kFO][kFE]
  float* Wo = Win + kFO * kFE;                      // [kFwdCh][kLDW]     W_out[c][in]
  float* half0 = Wo + kFwdCh * kLDW;                // per half-tile: z_i [kFE][kFT], hA, hB [kFO][kLDW], tt [kFT]
  constexpr int kHalfFloats = kFE * kFT + 2 * kFO * kLDW + kFT;
  float* fr = half0 + 2 * kHalfFloats;              // [kFO]
  float* bs = fr + kFO;                             // [kFL][kFO]
  float* ad = bs + kFL * kFO;                       // [kFwdCh] |delta_c|
  const int tid = threadIdx.x, half = tid / HT, ht = tid % HT;
  const int lo = ht % 16, hi = ht / 16;
  float* z_i = half0 + half * kHalfFloats;
  float* hbuf[2] = {z_i + kFE * kFT, z_i + kFE * kFT + kFO * kLDW};
  float* tt = z_i + kFE * kFT + 2 * kFO * kLDW;
  const int O = a.order, E = a.emb_dim;
  const int O4 = (O + 3) & ~3;
  const int cbase = blockIdx.y * kFwdCh;
  const int Dc = (a.D - cbase) < kFwdCh ? (a.D - cbase) : kFwdCh;
  for (int i = tid; i < (NL - 1) * kFO * kFO; i += NTH) {
    const int l = i / (kFO * kFO), j = (i / kFO) % kFO, ii = i % kFO;
    Wn[(l * kFO + j) * kLDW + ii] = (j < O && ii < O) ? a.w_h[(long long)l * O * O + j * O + ii] : 0.f;
  }
  for (int i = tid; i < kFO * kFE; i += NTH) {
    const int j = i / kFE, e = i % kFE;
    Win[i] = (j < O && e < E) ? a.w_in[j * E + e] : 0.f;
  }
  for (int i = tid; i < kFwdCh * kFO; i += NTH) {
    const int c = i / kFO, ii = i % kFO;
    Wo[c * kLDW + ii] = (c < Dc && ii < O) ? a.w_out[(long long)(cbase + c) * O + ii] : 0.f;
  }
  if (tid < kFO) fr[tid] = tid < O ? a.freq[tid] : 0.f;
  for (int i = tid; i < kFL * kFO; i += NTH) {
    const int l = i / kFO, j = i % kFO;
    float v = 0.f;
    if (j < O && l < NL) v = (l == 0) ? a.b_in[j] : a.b_h[(l - 1) * O + j];
    bs[i] = v;
  }
  for (int i = tid; i < kFwdCh; i += NTH) ad[i] = (a.modulate && i < Dc) ? fabsf(a.deltas[cbase + i]) : 0.f;

  const int ntiles = (a.L + 2 * kFT - 1) / (2 * kFT);   // a tile = 128 positions, 64 per half
  float zpre[ZQ], tpre = 0.f;
  auto fetch = [&](int tile) {
    const int t0 = tile * 2 * kFT + half * kFT;
#pragma unroll
    for (int q = 0; q < ZQ; ++q) {
      const int i = ht + q * HT, p = i / kFE, e = i % kFE;
      zpre[q] = (e < E && t0 + p < a.L) ? a.z[(long long)(t0 + p) * a.ldz + e] : 0.f;
    }
    if (ht < kFT) tpre = (t0 + ht < a.L) ? a.t[t0 + ht] : 0.f;
  };
  if ((int)blockIdx.x < ntiles) fetch(blockIdx.x);

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int t0 = tile * 2 * kFT + half * kFT;
    __syncthreads();          // previous tile's readers of z_i / tt / hbuf are done (and the weight staging, first time)
#pragma unroll
    for (int q = 0; q < ZQ; ++q) {
      const int i = ht + q * HT, p = i / kFE, e = i % kFE;
      z_i[e * kFT + p] = zpre[q];
    }
    if (ht < kFT) tt[ht] = tpre;
    if (tile + (int)gridDim.x < ntiles) fetch(tile + gridDim.x);
    __syncthreads();
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      float acc[TF][4];
#pragma unroll
      for (int x = 0; x < TF; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = bs[l * kFO + TF * hi + x];
      if (l == 0) {
        float4 zz[kFE];
#pragma unroll
        for (int e = 0; e < kFE; ++e) zz[e] = *reinterpret_cast<const float4*>(z_i + e * kFT + 4 * lo);
#pragma unroll
        for (int x = 0; x < TF; ++x) {
          const float4 wa = *reinterpret_cast<const float4*>(Win + (TF * hi + x) * kFE);
          const float4 wb = *reinterpret_cast<const float4*>(Win + (TF * hi + x) * kFE + 4);
#pragma unroll
          for (int e = 0; e < kFE; ++e) {
            const float w = e < 4 ? f4c(wa, e) : f4c(wb, e - 4);
            fma4(acc[x], w, zz[e]);
          }
        }
      } else {
        const float* W = Wn + (l - 1) * kFO * kLDW;
        const float* H = hbuf[(l - 1) & 1];
#pragma unroll 2
        for (int i4 = 0; i4 < O4; i4 += 4) {
          float4 w[TF], h[4];
#pragma unroll
          for (int x = 0; x < TF; ++x) w[x] = *reinterpret_cast<const float4*>(W + (TF * hi + x) * kLDW + i4);
#pragma unroll
          for (int r = 0; r < 4; ++r) h[r] = *reinterpret_cast<const float4*>(H + (i4 + r) * kLDW + 4 * lo);
#pragma unroll
          for (int x = 0; x < TF; ++x)
#pragma unroll
            for (int r = 0; r < 4; ++r)
              fma4(acc[x], f4c(w[x], r), h[r]);
        }
      }
      float* Hout = hbuf[l & 1];
      if (a.asave != nullptr && blockIdx.y == 0 && t0 + 4 * lo < a.lda) {
        // pre-activations for the saved-trunk backward: [layer][feature][position], 16-byte stores
#pragma unroll
        for (int x = 0; x < TF; ++x)
          *reinterpret_cast<float4*>(a.asave + (long long)(l * kFO + TF * hi + x) * a.lda + t0 + 4 * lo) =
              make_float4(acc[x][0], acc[x][1], acc[x][2], acc[x][3]);
      }
#pragma unroll
      for (int x = 0; x < TF; ++x) {
        const float f = fr[TF * hi + x];
        const float fa[4] = {f * acc[x][0], f * acc[x][1], f * acc[x][2], f * acc[x][3]};
        hy_sin4(fa, acc[x]);
        *reinterpret_cast<float4*>(Hout + (TF * hi + x) * kLDW + 4 * lo) = make_float4(acc[x][0], acc[x][1], acc[x][2], acc[x][3]);
      }
      if (l == NL - 1 && a.hsave != nullptr && blockIdx.y == 0) {
        // h_last[t][j] for the backward: position-major, 4 features per store
#pragma unroll
        for (int y = 0; y < 4; ++y) {
          const int t = t0 + 4 * lo + y;
          if (t >= a.L) continue;
          float* dst = a.hsave + (long long)t * a.ldh + TF * hi;
          if (TF * hi + TF - 1 < O && (a.ldh & 3) == 0) {
#pragma unroll
            for (int x4 = 0; x4 < TF; x4 += 4)
              *reinterpret_cast<float4*>(dst + x4) = make_float4(acc[x4][y], acc[x4 + 1][y], acc[x4 + 2][y], acc[x4 + 3][y]);
          } else {
#pragma unroll
            for (int x = 0; x < TF; ++x)
              if (TF * hi + x < O) dst[x] = acc[x][y];
          }
        }
      }
      __syncthreads();
    }
    // output layer + modulation, 64 channels per pass, straight from the resident W_out
    const float* H = hbuf[(NL - 1) & 1];
    const float4 tv = *reinterpret_cast<const float4*>(tt + 4 * lo);
    for (int c0 = 0; c0 < Dc; c0 += kFO) {
      float acc[TF][4];
#pragma unroll
      for (int x = 0; x < TF; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = 0.f;
      const float* W = Wo + c0 * kLDW;
#pragma unroll 2
      for (int i4 = 0; i4 < O4; i4 += 4) {
        float4 w[TF], h[4];
#pragma unroll
        for (int x = 0; x < TF; ++x) w[x] = *reinterpret_cast<const float4*>(W + (TF * hi + x) * kLDW + i4);
#pragma unroll
        for (int r = 0; r < 4; ++r) h[r] = *reinterpret_cast<const float4*>(H + (i4 + r) * kLDW + 4 * lo);
#pragma unroll
        for (int x = 0; x < TF; ++x)
#pragma unroll
          for (int r = 0; r < 4; ++r)
            fma4(acc[x], f4c(w[x], r), h[r]);
      }
#pragma unroll
      for (int x = 0; x < TF; ++x) {
        const int cl = c0 + TF * hi + x;
        if (cl >= Dc) continue;
        float r[4];
        if (a.modulate) {
          const float adc = ad[cl];
#pragma unroll
          for (int y = 0; y < 4; ++y) r[y] = acc[x][y] * (expf(-f4c(tv, y) * adc) + a.shift);
        } else {
#pragma unroll
          for (int y = 0; y < 4; ++y) r[y] = acc[x][y];
        }
        float* dst = k + (long long)(cbase + cl) * ldk + t0 + 4 * lo;
        if (t0 + 4 * lo + 3 < a.L && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
          *reinterpret_cast<float4*>(dst) = make_float4(r[0], r[1], r[2], r[3]);
        } else {
#pragma unroll
          for (int y = 0; y < 4; ++y)
            if (t0 + 4 * lo + y < a.L) dst[y] = r[y];
        }
      }
    }
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

static int g_trunk_minb = 1;   // experiment: resident CTAs per SM of the trunk-backward kernel (1 or 2)
extern "C" int hy_debug_set_trunk_minb(int v) {
  g_trunk_minb = (v == 2) ? 2 : 1;
  return g_trunk_minb;
}

extern "C" int hy_filter_trunk_bwd_layout(const hy_filter_args* p, int* n_cta, int* stride) {
  if (!p || !n_cta || !stride) return fail(HY_ERR_ARG, "hy_filter_trunk_bwd_layout: bad argument");
  const int ntiles = (p->L + kFT - 1) / kFT;
  const int cap = 148 * g_trunk_minb;
  *n_cta = ntiles < cap ? ntiles : cap;
  *stride = p->order * p->emb_dim + p->order + p->n_inner * (p->order * p->order + p->order) + p->order;
  return HY_OK;
}

static int trunk_bwd_impl(const hy_filter_args* p, const float* dh, int lddh, const float* a_save, int lda, float* part,
                          void* stream);
extern "C" int hy_filter_trunk_bwd(const hy_filter_args* p, const float* dh, int lddh, float* part, void* stream) {
  return trunk_bwd_impl(p, dh, lddh, nullptr, 0, part, stream);
}
extern "C" int hy_filter_trunk_save_layout(const hy_filter_args* p, int* lda, long long* elems) {
  if (!p || !lda || !elems || p->L < 1) return fail(HY_ERR_ARG, "hy_filter_trunk_save_layout: bad argument");
  *lda = (p->L + kFT - 1) / kFT * kFT;
  *elems = (long long)(1 + p->n_inner) * kFO * *lda;
  return HY_OK;
}
extern "C" int hy_filter_trunk_bwd_saved(const hy_filter_args* p, const float* dh, int lddh, const float* a_save, int lda,
                                         float* part, void* stream) {
  if (!a_save || !p || lda < p->L || lda % kFT || (reinterpret_cast<uintptr_t>(a_save) & 15))
    return fail(HY_ERR_ARG, "hy_filter_trunk_bwd_saved: a_save must be the 16-byte aligned buffer hy_filter_fwd_save_trunk filled");
  return trunk_bwd_impl(p, dh, lddh, a_save, lda, part, stream);
}
static int trunk_bwd_impl(const hy_filter_args* p, const float* dh, int lddh, const float* a_save, int lda, float* part,
                          void* stream) {
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
  a.hsave = nullptr; a.ldh = 0;
  a.asave = const_cast<float*>(a_save); a.lda = lda;
  int n_cta = 0, stride = 0;
  hy_filter_trunk_bwd_layout(p, &n_cta, &stride);
  if (a_save) {
    const size_t smem_s = sizeof(float) * (kTrunkSmemFloats + 2 * kFL * kFO * kFT);
    if (p->n_inner == 0) {
      auto kern = k_filter_trunk_bwd<1, 1, true>;
      HY_LAUNCH(kern, n_cta, kFThreads, smem_s, stream, a, dh, lddh, part, stride);
    } else if (p->n_inner == 1) {
      auto kern = k_filter_trunk_bwd<2, 1, true>;
      HY_LAUNCH(kern, n_cta, kFThreads, smem_s, stream, a, dh, lddh, part, stride);
    } else {
      auto kern = k_filter_trunk_bwd<3, 1, true>;
      HY_LAUNCH(kern, n_cta, kFThreads, smem_s, stream, a, dh, lddh, part, stride);
    }
    return check_launch("k_filter_trunk_bwd(saved)");
  }
  const size_t smem = sizeof(float) * kTrunkSmemFloats;
  if (p->n_inner == 0) {
    auto kern = k_filter_trunk_bwd<1>;
    HY_LAUNCH(kern, n_cta, kFThreads, smem, stream, a, dh, lddh, part, stride);
  } else if (p->n_inner == 1) {
    auto kern = k_filter_trunk_bwd<2>;
    HY_LAUNCH(kern, n_cta, kFThreads, smem, stream, a, dh, lddh, part, stride);
  } else if (g_trunk_minb == 2) {
    auto kern = k_filter_trunk_bwd<3, 2>;
    HY_LAUNCH(kern, n_cta, kFThreads, smem, stream, a, dh, lddh, part, stride);
  } else {
    auto kern = k_filter_trunk_bwd<3>;
    HY_LAUNCH(kern, n_cta, kFThreads, smem, stream, a, dh, lddh, part, stride);
  }
  return check_launch("k_filter_trunk_bwd");
}

static int g_fwd_tf = 4;   // features per thread of the fast forward kernel (4 or 8), see k_filter_fwd_fast: 8 measured slower
extern "C" int hy_debug_set_filter_fwd_tf(int v) {
  g_fwd_tf = (v == 8) ? 8 : 4;
  return g_fwd_tf;
}
static int filter_fwd_impl(const hy_filter_args* p, float* k, int ldk, float* h_last, int ldh, void* stream,
                           float* a_save = nullptr, int lda = 0);
extern "C" int hy_filter_fwd(const hy_filter_args* p, float* k, int ldk, void* stream) {
  return filter_fwd_impl(p, k, ldk, nullptr, 0, stream);
}
extern "C" int hy_filter_fwd_save_trunk(const hy_filter_args* p, float* k, int ldk, float* h_last, int ldh, float* a_save,
                                        int lda, void* stream) {
  if (!h_last || !a_save || !p || ldh < p->order || (reinterpret_cast<uintptr_t>(h_last) & 15) ||
      (reinterpret_cast<uintptr_t>(a_save) & 15) || lda < p->L || lda % kFT)
    return fail(HY_ERR_ARG, "hy_filter_fwd_save_trunk: h_last [L][ldh >= order] and a_save [1 + n_inner][64][lda %% 64 == 0] must be 16-byte aligned");
  if (p->emb_dim > kFE || p->n_inner > kFL - 1)
    return fail(HY_ERR_UNSUPPORTED, "hy_filter_fwd_save_trunk: emb_dim %d (<= %d) / n_inner %d (<= %d) outside the fused range",
                p->emb_dim, kFE, p->n_inner, kFL - 1);
  return filter_fwd_impl(p, k, ldk, h_last, ldh, stream, a_save, lda);
}
extern "C" int hy_filter_fwd_save(const hy_filter_args* p, float* k, int ldk, float* h_last, int ldh, void* stream) {
  if (!h_last || !p || ldh < p->order || (reinterpret_cast<uintptr_t>(h_last) & 15))
    return fail(HY_ERR_ARG, "hy_filter_fwd_save: h_last must be a 16-byte aligned [L][ldh >= order] buffer");
  if (p->emb_dim > kFE || p->n_inner > kFL - 1)
    return fail(HY_ERR_UNSUPPORTED, "hy_filter_fwd_save: emb_dim %d (<= %d) / n_inner %d (<= %d) outside the fused range",
                p->emb_dim, kFE, p->n_inner, kFL - 1);
  return filter_fwd_impl(p, k, ldk, h_last, ldh, stream);
}
static int filter_fwd_impl(const hy_filter_args* p, float* k, int ldk, float* h_last, int ldh, void* stream, float* a_save,
                           int lda) {
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
  a.hsave = h_last; a.ldh = ldh;
  a.asave = a_save; a.lda = lda;
#ifndef HY_EMU_BUILD
  if (filter_tc05_enabled()) {
    // tensor-core forward (tcgen05 / TMEM, hy_filter_tc05.cu); shapes outside its range fall through to the FFMA kernels
    const int rc = filter_fwd_tc05(a, k, ldk, stream);
    if (rc != HY_ERR_UNSUPPORTED) return rc;
  }
#endif
  if (p->emb_dim <= kFE && p->n_inner <= kFL - 1) {
    const int ntiles = (p->L + 2 * kFT - 1) / (2 * kFT);
    const int nslab = (p->D + kFwdCh - 1) / kFwdCh;
    int per = 148 / nslab;
    if (per < 1) per = 1;
    const dim3 grid(ntiles < per ? ntiles : per, nslab);
    const size_t smem = sizeof(float) * kFwdSmemFloats;
    if (p->n_inner == 0) {
      auto kern = k_filter_fwd_fast<1>;
      HY_LAUNCH(kern, grid, kFwdThreads, smem, stream, a, k, ldk);
    } else if (p->n_inner == 1) {
      auto kern = k_filter_fwd_fast<2>;
      HY_LAUNCH(kern, grid, kFwdThreads, smem, stream, a, k, ldk);
    } else if (g_fwd_tf == 8) {
      auto kern = k_filter_fwd_fast<3, 8>;
      HY_LAUNCH(kern, grid, kFwdThreads / 2, smem, stream, a, k, ldk);
    } else {
      auto kern = k_filter_fwd_fast<3>;
      HY_LAUNCH(kern, grid, kFwdThreads, smem, stream, a, k, ldk);
    }
    return check_launch("k_filter_fwd_fast");
  }
  const size_t smem = sizeof(float) * (kFO * kFO + 2 * kFO * kFT + 2 * kFO + kFT);
  HY_LAUNCH(k_filter_fwd_generic, (p->L + kFT - 1) / kFT, kFThreads, smem, stream, a, k, ldk);
  return check_launch("k_filter_fwd_generic");
}
