// Prototype: the 4096-point row transform of the long convolution's phase B as TWO radix-64 DFT-GEMM
// stages on tcgen05 — the measurement behind DESIGN.md 3.1 "why CUDA cores and not tcgen05 for the FFT".
//
//   x[64 n1 + n2]  --stage 1-->  Y[k1, n2] = sum_n1 W64^(n1 k1) x[n1, n2]      (GEMM: 128 x 128 real DFT block matrix x data)
//                  --twiddle-->  Y'[k1, n2] = Y[k1, n2] W4096^(n2 k1)           (CUDA cores, on the TMEM drain)
//                  --stage 2-->  X[k1 + 64 k2] = sum_n2 W64^(n2 k2) Y'[k1, n2]  (the same block matrix x Y')
// Complex arithmetic as real GEMMs: A[2k+c', 2n+c] = [[Fr, -Fi], [Fi, Fr]] (constant, shared memory), the data is the B
// operand (K-major: the (re, im) pair of a point is adjacent in K), accumulators in TMEM, 2 rows (N = 128) per group.
// fp32-class accuracy needs a split: kind::f16 with x = hi + lo (fp16 + fp16, 22 bits) and three products (lo*hi, hi*lo,
// hi*hi) — the fastest accurate option the unit offers (kind::tf32 x 3 runs at half the rate).
//
// Reports per 4096-point transform: total cycles, cycles the issuing thread waits for the MMAs, cycles of the CUDA-core
// sections (load + split + layout, drain + twiddle + split + layout, drain + store), the relative error against a
// double-precision DFT, and the time of the whole batch — to be set beside k_row_conv<4096,256,0> (2 transforms + the
// spectrum product per row pair item).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I dna_b200/csrc tools/proto_fft_tc.cu -o tools/bin/proto_fft_tc
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_fp16.h>
#include "hy_tc05.cuh"

using namespace tc05;

constexpr int kS = 4096, kR = 64;       // transform length, radix
constexpr int kRows = 2;                // rows per group (N = 128)
constexpr int kN = kRows * kR;          // MMA N
constexpr int kK = 2 * kR;              // MMA K = (n, re/im)

__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
  return (1u << 4)                      // c_format = F32
         | (0u << 7) | (0u << 10)       // a_format = b_format = F16
         | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}
// element (row, k) of a K-major no-swizzle 16-bit operand with K = 128: byte offset
__host__ __device__ inline int canon16(int row, int k) { return (row >> 3) * 2048 + (k >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2; }

// (re, im) -> packed fp16x2 hi and lo words
__device__ __forceinline__ void split_c(float re, float im, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(re, im);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(re - hf.x, im - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

struct Args {
  const float2* x;        // [nrows][4096]
  float2* X;              // [nrows][4096] natural order
  const __half* Ahi;      // canonical [128 x 128]
  const __half* Alo;
  const float2* tw;       // [64][64] W4096^(k1 n2)
  int nrows;
  unsigned long long* stats;   // [0] total cycles, [1] MMA wait cycles, [2] CUDA-core cycles, [3] groups   (thread 0 of CTA 0)
};

__device__ __forceinline__ void issue_stage(uint32_t tD, uint32_t sAhi, uint32_t sAlo, uint32_t sBhi, uint32_t sBlo) {
  constexpr uint32_t idesc = idesc_f16(128, kN);
  for (int pass = 0; pass < 3; ++pass) {      // lo*hi, hi*lo, hi*hi: small terms first
    const uint32_t sA = pass == 0 ? sAlo : sAhi, sB = pass == 1 ? sBlo : sBhi;
#pragma unroll
    for (int ks = 0; ks < kK / 16; ++ks)
      mma_f16_ss(tD, smem_desc_noswizzle(sA + ks * 256, 128, 2048), smem_desc_noswizzle(sB + ks * 256, 128, 2048), idesc,
                 (pass | ks) ? 1u : 0u);
  }
}

// Two tile slots of 128 threads per CTA (the structure of hy_filter_tc05.cu): while one slot waits for its MMAs the other
// runs its CUDA-core sections.  Each slot owns its B buffers, TMEM columns and mbarrier; the DFT matrix is shared.
__device__ __forceinline__ void slot_barrier(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory"); }

__global__ void __launch_bounds__(256, 1) k_fft4096_tc(Args a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int wg = threadIdx.x >> 7;
  unsigned char* sAhi = smem;                 // 32 KB each
  unsigned char* sAlo = smem + 32768;
  unsigned char* sBhi = smem + 65536 + wg * 65536;         // [N = 128 rows][K = 128] fp16: 32 KB each
  unsigned char* sBlo = sBhi + 32768;
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x & 127, warp = tid >> 5;
  uint64_t& bar = bars[wg];
  for (int i = threadIdx.x; i < 32768 / 16; i += 256) {
    reinterpret_cast<uint4*>(sAhi)[i] = reinterpret_cast<const uint4*>(a.Ahi)[i];
    reinterpret_cast<uint4*>(sAlo)[i] = reinterpret_cast<const uint4*>(a.Alo)[i];
  }
  if (threadIdx.x < 32) tmem_alloc<512>(&tmem_base_s);
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  fence_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = tmem_base_s + wg * 256;
  const uint32_t tD1 = tb, tD2 = tb + 128;
  const uint32_t lane_base = (uint32_t)(32 * warp) << 16;
  uint32_t parity = 0;
  const int kk = tid >> 1, c = tid & 1;        // TMEM lane = 2 * k + c: frequency index k of this stage, component c
  long long t_mma = 0, t_cuda = 0;
  const long long t_begin = clock64();
  int groups = 0;
  for (int g = 2 * blockIdx.x + wg; g * kRows < a.nrows; g += 2 * gridDim.x, ++groups) {
    long long t0 = clock64();
    // ---- load 2 rows, split, lay out as the B operand of stage 1: B1[n = (r, n2)][k = 2 n1 + c] --------------------------
    const float2* src = a.x + (long long)g * kRows * kS;
#pragma unroll 4
    for (int j = 0; j < kRows * kS / 128; ++j) {
      const int i = j * 128 + tid;
      const float2 v = src[i];
      const int r = i >> 12, nn = i & (kS - 1), n1 = nn >> 6, n2 = nn & 63;
      uint32_t hi, lo;
      split_c(v.x, v.y, hi, lo);
      const int off = canon16(r * 64 + n2, 2 * n1);
      *reinterpret_cast<uint32_t*>(sBhi + off) = hi;
      *reinterpret_cast<uint32_t*>(sBlo + off) = lo;
    }
    fence_async_smem();
    tc_fence_before_sync();
    slot_barrier(wg);
    long long t1 = clock64();
    t_cuda += t1 - t0;
    if (tid == 0) {
      tc_fence_after_sync();
      issue_stage(tD1, smem_u32(sAhi), smem_u32(sAlo), smem_u32(sBhi), smem_u32(sBlo));
      mma_commit(&bar);
    }
    mbar_wait(&bar, parity);
    parity ^= 1;
    tc_fence_after_sync();
    long long t2 = clock64();
    t_mma += t2 - t1;
    // ---- drain stage 1: lane (k1, c) holds Y_c[k1][(r, n2)]; pair exchange -> this thread owns n2 in [32 c, 32 c + 32) with
    // both components; twiddle; split; B operand of stage 2: B2[n = (r, k1)][k = 2 n2 + c'] ----------------------------------
    for (int r = 0; r < kRows; ++r) {
      uint32_t lo32[32], hi32[32];
      tmem_ld32(tD1 + lane_base + r * 64, lo32);          // n2 = 0 .. 31
      tmem_ld32(tD1 + lane_base + r * 64 + 32, hi32);     // n2 = 32 .. 63
      tmem_ld_wait();
      const int k1 = kk;
#pragma unroll
      for (int q = 0; q < 32; q += 4) {
        uint32_t whi[4], wlo[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          // even lane (c = 0, real parts) keeps n2 = q + u and needs the partner's imaginary part of it;
          // odd lane (c = 1) keeps n2 = 32 + q + u and needs the partner's real part of it
          const uint32_t send = c ? lo32[q + u] : hi32[q + u];
          const uint32_t got = __shfl_xor_sync(0xffffffffu, send, 1);
          const float yr = __uint_as_float(c ? got : lo32[q + u]);
          const float yi = __uint_as_float(c ? hi32[q + u] : got);
          const int n2 = 32 * c + q + u;
          const float2 t = a.tw[k1 * 64 + n2];
          split_c(yr * t.x - yi * t.y, yr * t.y + yi * t.x, whi[u], wlo[u]);
        }
        const int off = canon16(r * 64 + k1, 2 * (32 * c + q));       // 4 points = 8 halves = one 16-byte core-matrix row
        *reinterpret_cast<uint4*>(sBhi + off) = make_uint4(whi[0], whi[1], whi[2], whi[3]);
        *reinterpret_cast<uint4*>(sBlo + off) = make_uint4(wlo[0], wlo[1], wlo[2], wlo[3]);
      }
    }
    fence_async_smem();
    tc_fence_before_sync();
    slot_barrier(wg);
    long long t3 = clock64();
    t_cuda += t3 - t2;
    if (tid == 0) {
      tc_fence_after_sync();
      issue_stage(tD2, smem_u32(sAhi), smem_u32(sAlo), smem_u32(sBhi), smem_u32(sBlo));
      mma_commit(&bar);
    }
    mbar_wait(&bar, parity);
    parity ^= 1;
    tc_fence_after_sync();
    long long t4 = clock64();
    t_mma += t4 - t3;
    // ---- drain stage 2: lane (k2, c) holds X_c[(r, k1)] of frequency k1 + 64 k2; pair exchange, float2 stores ---------------
    float2* dst = a.X + (long long)g * kRows * kS;
    for (int r = 0; r < kRows; ++r) {
      uint32_t lo32[32], hi32[32];
      tmem_ld32(tD2 + lane_base + r * 64, lo32);
      tmem_ld32(tD2 + lane_base + r * 64 + 32, hi32);
      tmem_ld_wait();
      const int k2 = kk;
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const uint32_t send = c ? lo32[q] : hi32[q];
        const uint32_t got = __shfl_xor_sync(0xffffffffu, send, 1);
        const float xr = __uint_as_float(c ? got : lo32[q]);
        const float xi = __uint_as_float(c ? hi32[q] : got);
        dst[r * kS + 64 * k2 + 32 * c + q] = make_float2(xr, xi);
      }
    }
    tc_fence_before_sync();
    slot_barrier(wg);
    t_cuda += clock64() - t4;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.stats) {
    a.stats[0] = clock64() - t_begin;
    a.stats[1] = t_mma;
    a.stats[2] = t_cuda;
    a.stats[3] = groups;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem_base_s);
}

int main(int argc, char** argv) {
  const int nrows = argc > 1 ? atoi(argv[1]) : 32768;
  const double PI = 3.14159265358979323846;
  // DFT block matrix, split into fp16 hi / lo, canonical layout
  std::vector<__half> Ahi(128 * 128), Alo(128 * 128);
  for (int k = 0; k < 64; ++k)
    for (int n = 0; n < 64; ++n) {
      const double fr = cos(2 * PI * ((k * n) % 64) / 64.0), fi = -sin(2 * PI * ((k * n) % 64) / 64.0);
      const double e[2][2] = {{fr, -fi}, {fi, fr}};
      for (int cp = 0; cp < 2; ++cp)
        for (int cc = 0; cc < 2; ++cc) {
          const float v = (float)e[cp][cc];
          const __half h = __float2half_rn(v);
          const __half l = __float2half_rn(v - __half2float(h));
          const int off = canon16(2 * k + cp, 2 * n + cc) / 2;
          Ahi[off] = h;
          Alo[off] = l;
        }
    }
  std::vector<float2> tw(64 * 64);
  for (int k1 = 0; k1 < 64; ++k1)
    for (int n2 = 0; n2 < 64; ++n2) tw[k1 * 64 + n2] = make_float2((float)cos(2 * PI * k1 * n2 / 4096.0), (float)-sin(2 * PI * k1 * n2 / 4096.0));
  std::vector<float2> x((size_t)nrows * kS);
  srand(3);
  for (auto& v : x) v = make_float2((float)rand() / RAND_MAX * 2.f - 1.f, (float)rand() / RAND_MAX * 2.f - 1.f);
  float2 *dx, *dX, *dtw;
  __half *dAhi, *dAlo;
  unsigned long long* dstats;
  cudaMalloc(&dx, x.size() * 8); cudaMalloc(&dX, x.size() * 8); cudaMalloc(&dtw, tw.size() * 8);
  cudaMalloc(&dAhi, Ahi.size() * 2); cudaMalloc(&dAlo, Alo.size() * 2); cudaMalloc(&dstats, 64);
  cudaMemcpy(dx, x.data(), x.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dtw, tw.data(), tw.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dAhi, Ahi.data(), Ahi.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dAlo, Alo.data(), Alo.size() * 2, cudaMemcpyHostToDevice);
  const size_t smem = 6 * 32768;
  cudaFuncSetAttribute(k_fft4096_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  Args a{dx, dX, dAhi, dAlo, dtw, nrows, dstats};
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int it = 0; it < 3; ++it) k_fft4096_tc<<<sms, 256, smem>>>(a);
  cudaEventRecord(e0);
  const int iters = 10;
  for (int it = 0; it < iters; ++it) k_fft4096_tc<<<sms, 256, smem>>>(a);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return 1; }
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= iters;
  unsigned long long st[4];
  cudaMemcpy(st, dstats, 32, cudaMemcpyDeviceToHost);
  // accuracy: rows 0, 1 and the last one against a double-precision DFT
  std::vector<float2> X(x.size());
  cudaMemcpy(X.data(), dX, X.size() * 8, cudaMemcpyDeviceToHost);
  double max_err = 0, max_ref = 0;
  for (int row : {0, 1, nrows - 1}) {
    for (int k = 0; k < kS; k += 1) {
      double sr = 0, si = 0;
      for (int n = 0; n < kS; ++n) {
        const double ang = -2 * PI * ((long long)k * n % kS) / kS;
        const double cr = cos(ang), ci = sin(ang);
        const float2 v = x[(size_t)row * kS + n];
        sr += v.x * cr - v.y * ci;
        si += v.x * ci + v.y * cr;
      }
      const float2 got = X[(size_t)row * kS + k];
      max_err = fmax(max_err, fmax(fabs(got.x - sr), fabs(got.y - si)));
      max_ref = fmax(max_ref, fmax(fabs(sr), fabs(si)));
    }
  }
  const double per_row_ns = ms * 1e6 / nrows;
  printf("tcgen05 kind::f16 x3, two radix-64 DFT-GEMM stages, 4096-point complex rows: %d rows in %.3f ms = %.2f ns per transform "
         "(chip-wide), %.1f GB/s of row traffic\n", nrows, ms, per_row_ns, 2.0 * nrows * kS * 8 / ms / 1e6);
  printf("  max abs error %.3e against max |X| %.3e -> relative %.3e (fp32 FFT of this size: ~3e-7)\n", max_err, max_ref, max_err / max_ref);
  printf("  CTA 0, slot 0 (two slots of 128 threads alternate on the SM): %llu groups of 2 rows, cycles per transform of this slot: total %.0f, waiting for the MMAs %.0f, CUDA-core sections %.0f\n", st[3],
         (double)st[0] / (2.0 * st[3]), (double)st[1] / (2.0 * st[3]), (double)st[2] / (2.0 * st[3]));
  printf("  chip-wide: %.0f SM-cycles per transform at 1.9 GHz\n", per_row_ns * 1.9 * sms);
  return 0;
}
