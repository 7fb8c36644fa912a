// hyena-b200: implicit-filter MLP on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators AND the
// activation operand in tensor memory).
//
// Reference: HyenaFilter.filter, /root/reference/src/models/sequence/hyena.py:233-242 with the MLP of :203-219
//     h0 = sin(f * (z W0^T + b0)); h_l = sin(f * (h_{l-1} W_l^T + b_l)); k[c][t] = (W_out h_last(t))[c] * (exp(-t|d_c|) + shift)
// The Linear layers [L,64]x[64,64] and [L,64]x[64,D] are the one plain dense contraction of the hot path,
// 50 kFLOP per position: on FFMA (k_filter_fwd_fast) they are 58 % of the FMA pipe and the kernel's bound.  Here:
//   * one persistent CTA per SM, two tile slots of 256 threads, each owning a stream of 128-position tiles: two threads
//     (one per column half) are position t0 + r through the whole network — TMEM lane r — so an activation never
//     crosses threads:
//         accumulator (TMEM, lane r) -> tcgen05.ld -> registers: + bias, sin(f a), 3xTF32 split -> tcgen05.st ->
//         A operand of the next layer (TMEM, lane r);
//   * the weights sit in shared memory once per CTA as pre-split (hi, lo) TF32 pairs in the K-major no-swizzle core
//     matrix layout (hy_tc05.cuh), 192 KB for W_1, W_2 and a 256-channel slab of W_out;
//   * every Linear is 3 x 8 MMAs of 128 x N x 8 (lo*hi, hi*lo, hi*hi: small terms first) issued by one thread and
//     tracked by an mbarrier (tcgen05.commit); the output layer runs as two N = 128 halves into the same 128 columns;
//   * TMEM per tile slot: columns [0,64) A_hi, [64,128) A_lo, [128,256) accumulator;
//   * the first layer (K = emb_dim <= 8) stays on the CUDA cores: 5 FMAs per feature.
// While one tile slot waits for its MMAs the other evaluates its sines, so the tensor pipe and the FMA/MUFU pipes
// overlap without any intra-tile software pipeline.
// 3xTF32: sin(10 x) amplifies operand rounding, one TF32 pass is not enough; hi carries 11 bits
// (rounded), lo the next 11 (rounded), the unit truncates below — measured 1e-6 relative (profiles/r02a_tc_probe.log).
#include "hy_filter.h"
#ifndef HY_EMU_BUILD
#include "hy_tc05.cuh"
#include <cstdlib>

namespace hy {

using namespace tc05;

constexpr int kTcThreads = 512;       // two tile slots x two column halves x 128 rows
constexpr int kTcM = 128;             // positions per tile (= TMEM lanes)
constexpr int kTcCh = 256;            // output channels resident per CTA (gridDim.y slabs beyond)
constexpr int kTcE = 8;               // padded emb_dim
constexpr int kTcMaxL = 3;            // Linear+Sin layers (1 + n_inner)

// shared memory carve-up (bytes)
constexpr int kOffWh = 0;                                   // [(kTcMaxL-1)][hi|lo][64 x 64] canonical, 16 KB each
constexpr int kOffWo = kOffWh + (kTcMaxL - 1) * 2 * 16384;  // [hi|lo][256 x 64] canonical, 64 KB each
constexpr int kOffWin = kOffWo + 2 * 65536;                 // [64][8] fp32
constexpr int kOffBias = kOffWin + 64 * kTcE * 4;           // [kTcMaxL][64]
constexpr int kOffFreq = kOffBias + kTcMaxL * 64 * 4;       // [64]
constexpr int kOffAd = kOffFreq + 64 * 4;                   // [256] |delta_c| * log2(e)
constexpr int kOffBar = kOffAd + kTcCh * 4;                 // 2 mbarriers + tmem slot
constexpr int kTcSmemBytes = kOffBar + 64;

// element (row n, k) of a K-major no-swizzle operand with K = 64: byte offset inside the operand
__device__ __forceinline__ int canon64(int n, int k) { return (n >> 3) * 2048 + (k >> 2) * 128 + (n & 7) * 16 + (k & 3) * 4; }

// hi / lo with BOTH halves rounded to TF32 (the unit truncates): x ~ hi + lo to 2^-22 relative, unbiased
__device__ __forceinline__ void split_tf32_rr(float x, uint32_t& hi, uint32_t& lo) {
  const uint32_t h = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  hi = h;
  lo = (__float_as_uint(x - __uint_as_float(h)) + 0x1000u) & 0xffffe000u;
}

__device__ __forceinline__ void wg_barrier(int wg) { asm volatile("bar.sync %0, 256;" ::"r"(wg + 1) : "memory"); }

// 3 x 8 MMAs of one Linear: D[128 x N] = A[128 x 64] * W[N x 64]^T with A = (hi, lo) in TMEM, W = (hi, lo) in smem
template <int N>
__device__ __forceinline__ void issue_linear(uint32_t tD, uint32_t tAhi, uint32_t tAlo, uint32_t sWhi, uint32_t sWlo) {
  constexpr uint32_t idesc = idesc_tf32(kTcM, N);
  const uint64_t dhi = smem_desc_noswizzle(sWhi, 128, 2048);
  const uint64_t dlo = smem_desc_noswizzle(sWlo, 128, 2048);
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) mma_tf32_ts(tD, tAlo + ks * 8, dhi + (uint64_t)(ks * 16), idesc, ks ? 1u : 0u);   // lo * hi
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) mma_tf32_ts(tD, tAhi + ks * 8, dlo + (uint64_t)(ks * 16), idesc, 1u);              // hi * lo
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) mma_tf32_ts(tD, tAhi + ks * 8, dhi + (uint64_t)(ks * 16), idesc, 1u);              // hi * hi
}

// sin of 16 values with ONE range test (the polynomial path of hy_common.cuh for |x| < 1e5, libm beyond): 16 independent
// evaluations for the scheduler to interleave
__device__ __forceinline__ void sin16(float (&x)[16]) {
  float m = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) m = fmaxf(m, fabsf(x[i]));
  if (m < 1.0e5f) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float c;
      hy_sincos_core(x[i], x[i], c);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = sinf(x[i]);
  }
}

// h[j] = sin(freq[j] * a[j]) for 16 features (in place), split and stored as the next layer's A operand (columns c0 .. c0+15)
__device__ __forceinline__ void act_store16(float (&a16)[16], const float* __restrict__ fr, int c0, uint32_t tAhi, uint32_t tAlo) {
#pragma unroll
  for (int q = 0; q < 16; q += 4) {
    const float4 f = *reinterpret_cast<const float4*>(fr + c0 + q);
    a16[q] *= f.x; a16[q + 1] *= f.y; a16[q + 2] *= f.z; a16[q + 3] *= f.w;
  }
  sin16(a16);
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) split_tf32_rr(a16[i], hi[i], lo[i]);
  tmem_st16(tAhi + c0, hi);
  tmem_st16(tAlo + c0, lo);
}

// features c0 .. c0+15 of one row of h_last [L][ldh] (ldh % 4 == 0, 16-byte aligned rows); only the first O exist
__device__ __forceinline__ void store_h16(float* __restrict__ row, int c0, int O, const float (&h16)[16]) {
  if (c0 + 16 <= O) {
#pragma unroll
    for (int q = 0; q < 16; q += 4)
      *reinterpret_cast<float4*>(row + c0 + q) = make_float4(h16[q], h16[q + 1], h16[q + 2], h16[q + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (c0 + i < O) row[c0 + i] = h16[i];
  }
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Thread layout: 512 threads = 2 tile slots x 2 column halves x 128 rows.  Thread (slot, half, r) owns row r of the
// slot's current tile (TMEM lane r: its warp index % 4 is r / 32 as the hardware requires) and half of the columns of
// every layer: features 32*half .. +31 of the trunk, channels 64*half .. +63 of each 128-channel output half.
template <int NL>
__global__ void __launch_bounds__(kTcThreads, 1) k_filter_fwd_tc05(FilterDev a, float* __restrict__ k, int ldk) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* Win = reinterpret_cast<float*>(smem + kOffWin);
  float* bias = reinterpret_cast<float*>(smem + kOffBias);
  float* fr = reinterpret_cast<float*>(smem + kOffFreq);
  float* ad2 = reinterpret_cast<float*>(smem + kOffAd);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + 32);
  const int tid = threadIdx.x, wg = tid >> 8, half = (tid >> 7) & 1, r = tid & 127, quarter = r >> 5;
  const int O = a.order, E = a.emb_dim;
  const int cbase = blockIdx.y * kTcCh;
  const int Dc = (a.D - cbase) < kTcCh ? (a.D - cbase) : kTcCh;

  // ---- one-time staging: weights as (hi, lo) TF32 pairs in the canonical layout, zero padded -------------------
  for (int i = tid; i < (NL - 1) * 64 * 64; i += kTcThreads) {
    const int l = i >> 12, n = (i >> 6) & 63, kk = i & 63;
    const float w = (n < O && kk < O) ? a.w_h[(long long)l * O * O + n * O + kk] : 0.f;
    uint32_t hi, lo;
    split_tf32_rr(w, hi, lo);
    *reinterpret_cast<uint32_t*>(smem + kOffWh + (2 * l) * 16384 + canon64(n, kk)) = hi;
    *reinterpret_cast<uint32_t*>(smem + kOffWh + (2 * l + 1) * 16384 + canon64(n, kk)) = lo;
  }
  for (int i = tid; i < kTcCh * 64; i += kTcThreads) {
    const int n = i >> 6, kk = i & 63;
    const float w = (n < Dc && kk < O) ? a.w_out[(long long)(cbase + n) * O + kk] : 0.f;
    uint32_t hi, lo;
    split_tf32_rr(w, hi, lo);
    *reinterpret_cast<uint32_t*>(smem + kOffWo + canon64(n, kk)) = hi;
    *reinterpret_cast<uint32_t*>(smem + kOffWo + 65536 + canon64(n, kk)) = lo;
  }
  for (int i = tid; i < 64 * kTcE; i += kTcThreads) {
    const int j = i / kTcE, e = i % kTcE;
    Win[i] = (j < O && e < E) ? a.w_in[j * E + e] : 0.f;
  }
  for (int i = tid; i < kTcMaxL * 64; i += kTcThreads) {
    const int l = i >> 6, j = i & 63;
    float v = 0.f;
    if (j < O && l < NL) v = (l == 0) ? a.b_in[j] : a.b_h[(l - 1) * O + j];
    bias[i] = v;
  }
  if (tid < 64) fr[tid] = tid < O ? a.freq[tid] : 0.f;
  for (int i = tid; i < kTcCh; i += kTcThreads)
    ad2[i] = (a.modulate && i < Dc) ? fabsf(a.deltas[cbase + i]) * 1.4426950408889634f : 0.f;
  if (tid < 32) tmem_alloc<512>(tmem_slot);
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  fence_async_smem();          // the weights were written by the generic proxy, the tensor core reads them through the async one
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase0 = *tmem_slot + wg * 256;                          // lane 0: MMA operand addresses
  const uint32_t tbase = tbase0 + ((uint32_t)(quarter * 32) << 16);       // this warp's lane quarter
  const uint32_t tAhi = tbase, tAlo = tbase + 64, tD = tbase + 128;
  uint64_t* bar = &bars[wg];
  uint32_t parity = 0;
  const uint32_t sWh = smem_u32(smem + kOffWh), sWo = smem_u32(smem + kOffWo);
  const bool save_a = a.asave != nullptr && blockIdx.y == 0;
  const bool save_h = a.hsave != nullptr && blockIdx.y == 0;
  const int f0 = 32 * half;                                                // first trunk feature of this thread
  const bool issuer = (r == 0 && half == 0);

  const int ntiles = (a.L + kTcM - 1) / kTcM;
  for (int tile = 2 * blockIdx.x + wg; tile < ntiles; tile += 2 * gridDim.x) {
    const int t = tile * kTcM + r;
    const bool live = t < a.L;
    const bool sa = save_a && t < a.lda;
    // ---- layer 0 on the CUDA cores: a0[j] = b0[j] + sum_e W0[j][e] z[t][e] ------------------------------------
    float z[kTcE];
#pragma unroll
    for (int e = 0; e < kTcE; ++e) z[e] = (live && e < E) ? __ldg(a.z + (long long)t * a.ldz + e) : 0.f;
    const float tt = live ? __ldg(a.t + t) : 0.f;
#pragma unroll 1
    for (int c0 = f0; c0 < f0 + 32; c0 += 16) {
      float a16[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const float4 w0 = *reinterpret_cast<const float4*>(Win + (c0 + q) * kTcE);
        const float4 w1 = *reinterpret_cast<const float4*>(Win + (c0 + q) * kTcE + 4);
        float s = bias[c0 + q];
        s = fmaf(w0.x, z[0], s); s = fmaf(w0.y, z[1], s); s = fmaf(w0.z, z[2], s); s = fmaf(w0.w, z[3], s);
        s = fmaf(w1.x, z[4], s); s = fmaf(w1.y, z[5], s); s = fmaf(w1.z, z[6], s); s = fmaf(w1.w, z[7], s);
        a16[q] = s;
      }
      if (sa) {
        float* p = a.asave + (long long)c0 * a.lda + t;
#pragma unroll
        for (int q = 0; q < 16; ++q, p += a.lda) *p = a16[q];
      }
      act_store16(a16, fr, c0, tAhi, tAlo);
      if (NL == 1 && save_h && live) store_h16(a.hsave + (long long)t * a.ldh, c0, O, a16);
    }
    // ---- hidden layers on the tensor cores -------------------------------------------------------------------
#pragma unroll
    for (int l = 1; l < NL; ++l) {
      tmem_st_wait();
      tc_fence_before_sync();
      wg_barrier(wg);
      if (issuer) {
        tc_fence_after_sync();
        issue_linear<64>(tbase0 + 128, tbase0, tbase0 + 64, sWh + (2 * (l - 1)) * 16384, sWh + (2 * (l - 1) + 1) * 16384);
        mma_commit(bar);
      }
      mbar_wait(bar, parity);
      parity ^= 1;
      tc_fence_after_sync();
      uint32_t raw[32];
      tmem_ld32(tD + f0, raw);
      tmem_ld_wait();
#pragma unroll
      for (int cc = 0; cc < 32; cc += 16) {
        const int c0 = f0 + cc;
        float a16[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) a16[q] = __uint_as_float(raw[cc + q]) + bias[l * 64 + c0 + q];
        if (sa) {
          float* p = a.asave + (long long)(l * 64 + c0) * a.lda + t;
#pragma unroll
          for (int q = 0; q < 16; ++q, p += a.lda) *p = a16[q];
        }
        act_store16(a16, fr, c0, tAhi, tAlo);
        if (l == NL - 1 && save_h && live) store_h16(a.hsave + (long long)t * a.ldh, c0, O, a16);
      }
    }
    // ---- output layer (two halves of 128 channels) + modulation + channel-major store ---------------------------
#pragma unroll 1
    for (int hh = 0; hh < 2; ++hh) {
      if (hh * 128 >= Dc) break;
      tmem_st_wait();
      tc_fence_before_sync();
      wg_barrier(wg);          // half 0: the A operand is complete; half 1: every thread has drained the accumulator
      if (issuer) {
        tc_fence_after_sync();
        issue_linear<128>(tbase0 + 128, tbase0, tbase0 + 64, sWo + hh * 32768, sWo + 65536 + hh * 32768);
        mma_commit(bar);
      }
      mbar_wait(bar, parity);
      parity ^= 1;
      tc_fence_after_sync();
      const int cb = hh * 128 + 64 * half;       // first channel (inside the slab) of this thread
      const float ntt = -tt;
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t raw[32];
        tmem_ld32(tD + 64 * half + c0, raw);
        tmem_ld_wait();
        if (live) {
          float* dst = k + (long long)(cbase + cb + c0) * ldk + t;
          const float* adp = ad2 + cb + c0;
          if (cb + c0 + 32 <= Dc) {
            if (a.modulate) {
#pragma unroll
              for (int q = 0; q < 32; ++q, dst += ldk) *dst = __uint_as_float(raw[q]) * (ex2_approx(ntt * adp[q]) + a.shift);
            } else {
#pragma unroll
              for (int q = 0; q < 32; ++q, dst += ldk) *dst = __uint_as_float(raw[q]);
            }
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q, dst += ldk) {
              if (cb + c0 + q < Dc) {
                float v = __uint_as_float(raw[q]);
                if (a.modulate) v *= ex2_approx(ntt * adp[q]) + a.shift;
                *dst = v;
              }
            }
          }
        }
      }
    }
    tc_fence_before_sync();
    wg_barrier(wg);            // the accumulator and A columns are free for this tile slot's next tile
  }
  tc_fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc<512>(*tmem_slot);
}

static int g_tc05 = -1;
bool filter_tc05_enabled() {
  if (g_tc05 < 0) {
    const char* e = getenv("HYENA_B200_FILTER_TC05");
    g_tc05 = (e && e[0] == '0') ? 0 : 1;
  }
  return g_tc05 != 0;
}
extern "C" int hy_debug_set_filter_tc05(int on) {
  g_tc05 = on ? 1 : 0;
  return g_tc05;
}

int filter_fwd_tc05(const FilterDev& a, float* k, int ldk, void* stream) {
  if (a.order > 64 || a.emb_dim > kTcE || a.n_inner > kTcMaxL - 1 || a.n_inner < 0) return HY_ERR_UNSUPPORTED;
  if (a.hsave && ((a.ldh & 3) || (reinterpret_cast<uintptr_t>(a.hsave) & 15))) return HY_ERR_UNSUPPORTED;
  const int ntiles = (a.L + kTcM - 1) / kTcM;
  const int nslab = (a.D + kTcCh - 1) / kTcCh;
  int per = 148 / nslab;
  if (per < 1) per = 1;
  const int gx = (ntiles + 1) / 2 < per ? (ntiles + 1) / 2 : per;
  const dim3 grid(gx, nslab);
  if (a.n_inner == 0) {
    auto kern = k_filter_fwd_tc05<1>;
    HY_LAUNCH(kern, grid, kTcThreads, kTcSmemBytes, stream, a, k, ldk);
  } else if (a.n_inner == 1) {
    auto kern = k_filter_fwd_tc05<2>;
    HY_LAUNCH(kern, grid, kTcThreads, kTcSmemBytes, stream, a, k, ldk);
  } else {
    auto kern = k_filter_fwd_tc05<3>;
    HY_LAUNCH(kern, grid, kTcThreads, kTcSmemBytes, stream, a, k, ldk);
  }
  return check_launch("k_filter_fwd_tc05");
}

}  // namespace hy
#else
namespace hy {
int filter_fwd_tc05(const FilterDev&, float*, int, void*) { return HY_ERR_UNSUPPORTED; }
bool filter_tc05_enabled() { return false; }
}  // namespace hy
extern "C" int hy_debug_set_filter_tc05(int) { return 0; }
#endif
