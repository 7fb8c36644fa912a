// hyena-b200: backward of the filter MLP's last Linear on the tensor cores.
//
// Reference (src/models/sequence/hyena.py:219-242): k[t][c] = (h_last[t] . W_out[c]) * (exp(-t_t |delta_c|) + shift).
// Given dk [D][L] (channel-major, as the long conv produces it) the gradient wrt the Linear's output is
// dh[t][c] = dk[c][t] * (exp(-t_t |delta_c|) + shift), and the Linear's own backward is two dense contractions:
//     dh_last [L][O] = dh   @ W_out          (K = D)
//     dW_out  [D][O] = dh^T @ h_last         (K = L)
// These ARE matrix products (no FFT structure), 65 GFLOP per layer at L = 1M, D = 256, O = 64 — the one place on the
// path where the tensor pipe is the right unit.  The first cut ran them as a transpose kernel plus two fp32 SIMT
// GEMMs (2.2 ms per layer, three passes over a 1 GB [L][D] intermediate).  Here ONE kernel streams dk once,
// applies the modulation while staging a [D][64-position] tile in shared memory (dh is never materialised) and
// feeds both products to mma.sync.m16n8k8 TF32 with the 3xTF32 split (x = big + small, three MMAs per product:
// small*big + big*small + big*big, fp32 accumulate): the filter gradient keeps fp32-class accuracy (sin(10 x)
// amplifies rounding, so single-pass TF32/bf16 is not acceptable here), the arithmetic runs on the tensor pipe.
// dW_out is accumulated in registers over all the tiles of a persistent CTA and reduced over CTAs in a fixed order
// (deterministic).
#include "hy_host.h"
#include <algorithm>

namespace hy {

constexpr int kTcT = 64;      // positions per tile
constexpr int kTcO = 64;      // MLP width (order) this kernel is built for
constexpr int kTcLD = 72;     // shared-memory row stride (floats): fragment loads with 4 rows x 8 columns hit 32 banks
constexpr int kTcThreads = 256;
constexpr int kTcMaxD = 256;

#ifndef HY_EMU_BUILD

// x = big + small with big on the TF32 grid.  cvt.rna.tf32.f32 is emulated on sm_100 (add, infinity test, select:
// measured 37 % of this kernel's instructions), so the rounding is done by hand: adding half a TF32 ulp to the bit
// pattern rounds the magnitude to nearest once the low 13 bits are dropped, and the MMA unit ignores those bits of a
// .tf32 operand.  The residual x - big is exact in fp32 and goes in as it is (truncated by the unit: 2^-21 relative).
__device__ __forceinline__ void split_tf32(float x, unsigned& big, unsigned& small) {
  big = __float_as_uint(x) + 0x1000u;
  small = __float_as_uint(x - __uint_as_float(big & 0xffffe000u));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// d += A * B with fp32-class accuracy is three MMAs per product: small*big, big*small, big*big (small terms first; the
// small*small term, 2^-22 relative, is dropped).  The callers issue each term for all their independent accumulators
// before the next term, so that consecutive MMAs never depend on each other.

// One persistent CTA per SM; tile = 64 positions x D channels.
//   A_s [D][72]      dh tile, A_s[c][p] = dk[c][p0+p] * mod(c, p0+p)
//   W_s [D][72]      W_out, resident
//   H_s [2][64][72]  h_last tiles (cp.async, double buffered)
// The next tile's dk (64 floats per thread) is fetched into registers and its h_last rows by cp.async while the two
// products of the current tile run: the first cut loaded and computed in turn and spent 60 % of its time on the loads
// (ncu: long-scoreboard stalls, 8 warps per SM).
// GEMM1 (dh_last tile [64 p][64 o], K = D): warp w -> position block (w & 3) * 16, feature half (w >> 2) * 32.
// GEMM2 (dW_out [D][64 o] += ..., K = 64 p): warp w -> channels 32 w .. 32 w + 31, all 64 features; the running sums
//   live in registers across the CTA's tiles.
// The tensor unit accumulates with truncation, which biases a long chain (measured 2e-5 relative on dW_out at
// L = 300 k when the chain ran over all tiles): every chain inside the unit is kept to <= 24 MMAs from a zero
// accumulator and is added to the running sum with a rounded fp32 add.
__global__ void __launch_bounds__(kTcThreads, 1)
k_filter_out_bwd(const float* __restrict__ dk, int lddk, const float* __restrict__ t, const float* __restrict__ deltas,
                 float shift, int modulate, const float* __restrict__ w_out, const float* __restrict__ h_last, int ldh,
                 float* __restrict__ dh_last, int lddh, float* __restrict__ part, int D, int L) {
  extern __shared__ __align__(16) float smf[];
  float* A_s = smf;                       // [kTcMaxD][kTcLD]
  float* W_s = A_s + kTcMaxD * kTcLD;     // [kTcMaxD][kTcLD]
  float* H_s = W_s + kTcMaxD * kTcLD;     // [2][kTcT][kTcLD]
  float* ad_s = H_s + 2 * kTcT * kTcLD;   // [kTcMaxD] |delta_c|
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tg = lane & 3;
  for (int i = tid; i < kTcMaxD * kTcO; i += kTcThreads) {
    const int c = i / kTcO, o = i % kTcO;
    W_s[c * kTcLD + o] = c < D ? w_out[(long long)c * kTcO + o] : 0.f;
  }
  for (int i = tid; i < kTcMaxD; i += kTcThreads) ad_s[i] = (modulate && i < D) ? fabsf(deltas[i]) : 0.f;
  float accW[2][8][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int r = 0; r < 4; ++r) accW[mt][nt][r] = 0.f;
  const int ntiles = (L + kTcT - 1) / kTcT;
  const bool wactive = warp * 32 < D;     // GEMM2: this warp's channel block exists
  const int q = tid & 15;                 // float4 index inside the 64-position row segment
  const int c0 = tid >> 4;                // first channel of this thread's 16 (stride 16)
  constexpr int kRowsPerThread = kTcMaxD / (kTcThreads / 16);   // 16

  float4 pre[kRowsPerThread];             // the next tile's dk values
  float tv[4];                            // and its time stamps
  auto fetch = [&](int tile, int buf) {
    const int p0 = tile * kTcT;
    const int pp = p0 + 4 * q;
#pragma unroll
    for (int y = 0; y < 4; ++y) tv[y] = (pp + y < L) ? t[pp + y] : 0.f;
    if (pp + 3 < L) {
#pragma unroll
      for (int i = 0; i < kRowsPerThread; ++i) {
        const int c = c0 + 16 * i;
        pre[i] = c < D ? *reinterpret_cast<const float4*>(dk + (long long)c * lddk + pp) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
#pragma unroll
      for (int i = 0; i < kRowsPerThread; ++i) {
        const int c = c0 + 16 * i;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < D) {
          const float* src = dk + (long long)c * lddk + pp;
          if (pp < L) v.x = src[0];
          if (pp + 1 < L) v.y = src[1];
          if (pp + 2 < L) v.z = src[2];
        }
        pre[i] = v;
      }
    }
    float* Hb = H_s + buf * kTcT * kTcLD;
    for (int i = tid; i < kTcT * (kTcO / 4); i += kTcThreads) {
      const int r = i / (kTcO / 4), f4 = i % (kTcO / 4);
      float* dst = Hb + r * kTcLD + 4 * f4;
      if (p0 + r < L) hy_cp_async16(dst, h_last + (long long)(p0 + r) * ldh + 4 * f4);
      else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };

  int buf = 0;
  if ((int)blockIdx.x < ntiles) fetch(blockIdx.x, 0);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
    const int p0 = tile * kTcT;
    __syncthreads();   // previous tile's fragment reads are done (first time: W_s / ad_s staged)
    // ---- the modulated dk tile: registers -> A_s ----
#pragma unroll
    for (int i = 0; i < kRowsPerThread; ++i) {
      const int c = c0 + 16 * i;
      float4 v = pre[i];
      if (modulate) {
        const float adc = ad_s[c];
        v.x *= __expf(-tv[0] * adc) + shift;
        v.y *= __expf(-tv[1] * adc) + shift;
        v.z *= __expf(-tv[2] * adc) + shift;
        v.w *= __expf(-tv[3] * adc) + shift;
      }
      *reinterpret_cast<float4*>(A_s + c * kTcLD + 4 * q) = v;
    }
    hy_cp_async_wait_all();
    __syncthreads();
    if (tile + (int)gridDim.x < ntiles) fetch(tile + gridDim.x, buf ^ 1);   // in flight during the two products
    const float* Hb = H_s + buf * kTcT * kTcLD;
    // ---- GEMM1: dh_last[p][o] = sum_c A_s[c][p] * W_s[c][o] ----
    {
      const int m0 = (warp & 3) * 16, nb = (warp >> 2) * 32;
      float acc[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[nt][r] = 0.f;
#pragma unroll 1
      for (int ks = 0; ks < D; ks += 32) {
        float tmp[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int r = 0; r < 4; ++r) tmp[nt][r] = 0.f;
#pragma unroll
        for (int k0 = ks; k0 < ks + 32; k0 += 8) {
          unsigned ab[4], as[4], bb[4][2], bs[4][2];
          split_tf32(A_s[(k0 + tg) * kTcLD + m0 + g], ab[0], as[0]);
          split_tf32(A_s[(k0 + tg) * kTcLD + m0 + g + 8], ab[1], as[1]);
          split_tf32(A_s[(k0 + tg + 4) * kTcLD + m0 + g], ab[2], as[2]);
          split_tf32(A_s[(k0 + tg + 4) * kTcLD + m0 + g + 8], ab[3], as[3]);
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            split_tf32(W_s[(k0 + tg) * kTcLD + nb + 8 * nt + g], bb[nt][0], bs[nt][0]);
            split_tf32(W_s[(k0 + tg + 4) * kTcLD + nb + 8 * nt + g], bb[nt][1], bs[nt][1]);
          }
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma_tf32(tmp[nt], as, bb[nt]);
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma_tf32(tmp[nt], ab, bs[nt]);
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma_tf32(tmp[nt], ab, bb[nt]);
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int r = 0; r < 4; ++r) acc[nt][r] += tmp[nt][r];
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int col = nb + 8 * nt + 2 * tg;
        const int r0 = p0 + m0 + g, r1 = r0 + 8;
        if (r0 < L) *reinterpret_cast<float2*>(dh_last + (long long)r0 * lddh + col) = make_float2(acc[nt][0], acc[nt][1]);
        if (r1 < L) *reinterpret_cast<float2*>(dh_last + (long long)r1 * lddh + col) = make_float2(acc[nt][2], acc[nt][3]);
      }
    }
    // ---- GEMM2: dW_out[c][o] += sum_p A_s[c][p] * H_s[p][o], one half of the features at a time ----
    if (wactive) {
#pragma unroll
      for (int nh = 0; nh < 2; ++nh) {
        float tmp[2][4][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int r = 0; r < 4; ++r) tmp[mt][nt][r] = 0.f;
#pragma unroll 2
        for (int k0 = 0; k0 < kTcT; k0 += 8) {
          unsigned bb[4][2], bs[4][2];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            split_tf32(Hb[(k0 + tg) * kTcLD + 32 * nh + 8 * nt + g], bb[nt][0], bs[nt][0]);
            split_tf32(Hb[(k0 + tg + 4) * kTcLD + 32 * nh + 8 * nt + g], bb[nt][1], bs[nt][1]);
          }
          unsigned ab[2][4], as[2][4];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            const int m0 = warp * 32 + 16 * mt;
            split_tf32(A_s[(m0 + g) * kTcLD + k0 + tg], ab[mt][0], as[mt][0]);
            split_tf32(A_s[(m0 + g + 8) * kTcLD + k0 + tg], ab[mt][1], as[mt][1]);
            split_tf32(A_s[(m0 + g) * kTcLD + k0 + tg + 4], ab[mt][2], as[mt][2]);
            split_tf32(A_s[(m0 + g + 8) * kTcLD + k0 + tg + 4], ab[mt][3], as[mt][3]);
          }
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) mma_tf32(tmp[mt][nt], as[mt], bb[nt]);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) mma_tf32(tmp[mt][nt], ab[mt], bs[nt]);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) mma_tf32(tmp[mt][nt], ab[mt], bb[nt]);
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int r = 0; r < 4; ++r) accW[mt][4 * nh + nt][r] += tmp[mt][nt][r];
      }
    }
  }
  // per-CTA partial of dW_out: part[blockIdx.x][c][o]
  if (wactive) {
    float* dst = part + (long long)blockIdx.x * kTcMaxD * kTcO;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int cc = warp * 32 + 16 * mt + g, col = 8 * nt + 2 * tg;
        *reinterpret_cast<float2*>(dst + cc * kTcO + col) = make_float2(accW[mt][nt][0], accW[mt][nt][1]);
        *reinterpret_cast<float2*>(dst + (cc + 8) * kTcO + col) = make_float2(accW[mt][nt][2], accW[mt][nt][3]);
      }
  }
}

// dW_out[c][o] = sum over CTAs, fixed order
__global__ void k_filter_out_bwd_reduce(const float* __restrict__ part, int nparts, float* __restrict__ dW, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int b = 0; b < nparts; ++b) s += part[(long long)b * kTcMaxD * kTcO + i];
  dW[i] = s;
}

static int tc_grid(int L) {
  static thread_local int cached_dev = -1, cached_n = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 1;
    cached_dev = dev;
    cached_n = n;
  }
  const int ntiles = (L + kTcT - 1) / kTcT;
  return std::max(1, std::min(ntiles, cached_n));
}
#endif  // !HY_EMU_BUILD

}  // namespace hy

using namespace hy;

extern "C" int hy_filter_out_bwd_supported(int D, int order) {
#ifdef HY_EMU_BUILD
  (void)D; (void)order;
  return 0;   // mma.sync has no CPU emulation: the host layer keeps the three-kernel path in the test build
#else
  return (order == kTcO && D >= 32 && D <= kTcMaxD && D % 32 == 0) ? 1 : 0;
#endif
}

extern "C" size_t hy_filter_out_bwd_workspace_bytes(int L) {
#ifdef HY_EMU_BUILD
  (void)L;
  return 0;
#else
  return sizeof(float) * (size_t)tc_grid(L) * kTcMaxD * kTcO;
#endif
}

extern "C" int hy_filter_out_bwd(const float* dk, int lddk, const float* t, const float* deltas, float shift, int modulate,
                                 const float* w_out, const float* h_last, int ldh, float* dh_last, int lddh, float* dW_out,
                                 int D, int order, int L, void* ws, size_t ws_bytes, void* stream) {
#ifdef HY_EMU_BUILD
  (void)dk; (void)lddk; (void)t; (void)deltas; (void)shift; (void)modulate; (void)w_out; (void)h_last; (void)ldh;
  (void)dh_last; (void)lddh; (void)dW_out; (void)D; (void)order; (void)L; (void)ws; (void)ws_bytes; (void)stream;
  return fail(HY_ERR_UNSUPPORTED, "hy_filter_out_bwd: tensor-core kernel, not available in the CPU test build");
#else
  if (!dk || !t || !w_out || !h_last || !dh_last || !dW_out || L < 1) return fail(HY_ERR_ARG, "hy_filter_out_bwd: bad argument");
  if (modulate && !deltas) return fail(HY_ERR_ARG, "hy_filter_out_bwd: modulate needs deltas");
  if (!hy_filter_out_bwd_supported(D, order)) return fail(HY_ERR_UNSUPPORTED, "hy_filter_out_bwd: needs order == 64, D %% 32 == 0, D <= 256");
  if ((lddk & 3) || (ldh & 3) || (lddh & 1) || (reinterpret_cast<uintptr_t>(dk) & 15) || (reinterpret_cast<uintptr_t>(h_last) & 15) ||
      (reinterpret_cast<uintptr_t>(dh_last) & 7))
    return fail(HY_ERR_ARG, "hy_filter_out_bwd: dk / h_last rows must be 16-byte aligned (strides %% 4 == 0), dh_last 8-byte");
  const int grid = tc_grid(L);
  if (!ws || ws_bytes < sizeof(float) * (size_t)grid * kTcMaxD * kTcO) return fail(HY_ERR_WORKSPACE, "hy_filter_out_bwd: workspace too small");
  const size_t smem = sizeof(float) * ((size_t)2 * kTcMaxD * kTcLD + (size_t)2 * kTcT * kTcLD + kTcMaxD);
  float* part = reinterpret_cast<float*>(ws);
  HY_LAUNCH(k_filter_out_bwd, grid, kTcThreads, smem, stream, dk, lddk, t, deltas, shift, modulate, w_out, h_last, ldh, dh_last,
            lddh, part, D, L);
  int rc = check_launch("k_filter_out_bwd");
  if (rc != HY_OK) return rc;
  // CTAs whose warps own no channel (D < 256) never wrote their slab rows >= D: only the first D rows are read
  const int n = D * kTcO;
  HY_LAUNCH(k_filter_out_bwd_reduce, (n + 255) / 256, 256, 0, stream, part, grid, dW_out, n);
  return check_launch("k_filter_out_bwd_reduce");
#endif
}
