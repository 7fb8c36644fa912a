// tcgen05 probe: pins down, on the real B200, the conventions hy_filter_tc05.cu relies on before the kernel is built
// on them — shared-memory descriptor (no-swizzle K-major: which of LBO / SBO is the K direction), A operand from
// TMEM, TMEM load/store lane mapping, TF32 input truncation, 3xTF32 accuracy, and the issue rate of the MMAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I dna_b200/csrc tools/tc_probe.cu -o tools/bin/tc_probe
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "hy_tc05.cuh"

using namespace tc05;

struct ProbeArgs {
  const float* A;   // [128][K]
  const float* B;   // [N][K]
  float* D;         // [128][N]
  int N, K;
  int mode;         // 0: A from TMEM (TS), 1: A from shared memory (SS)
  int swap;         // 1: exchange the roles of LBO and SBO in the descriptors
  int split3;       // 1: 3xTF32
  int reps;         // > 1: timing loop (result meaningless)
  int nacc;         // timing loop: number of accumulators the back-to-back MMAs alternate between (1..3)
  long long* cycles;
};

// element (r, k) of a K-major no-swizzle operand with `rows` rows and K columns -> byte offset
__device__ __host__ inline int canon_off(int r, int k, int K) { return (r / 8) * (128 * (K / 4)) + (k / 4) * 128 + (r % 8) * 16 + (k % 4) * 4; }

__global__ void __launch_bounds__(128, 1) k_probe(ProbeArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid / 32;
  const int N = a.N, K = a.K;
  unsigned char* sBhi = smem;
  unsigned char* sBlo = sBhi + N * K * 4;
  unsigned char* sAhi = sBlo + N * K * 4;
  unsigned char* sAlo = sAhi + 128 * K * 4;
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = tmem_base_s;
  const uint32_t tD = tb;            // columns [0, N)
  const uint32_t tAhi = tb + 256;    // columns [256, 256 + K)
  const uint32_t tAlo = tb + 256 + K;
  // B operand
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, k = i % K;
    uint32_t hi, lo;
    if (a.split3) split_tf32(a.B[i], hi, lo);
    else { hi = __float_as_uint(a.B[i]); lo = 0; }
    *reinterpret_cast<uint32_t*>(sBhi + canon_off(n, k, K)) = hi;
    *reinterpret_cast<uint32_t*>(sBlo + canon_off(n, k, K)) = lo;
  }
  // A operand: thread = row
  for (int k0 = 0; k0 < K; k0 += 16) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float x = a.A[tid * K + k0 + j];
      if (a.split3) split_tf32(x, hi[j], lo[j]);
      else { hi[j] = __float_as_uint(x); lo[j] = 0; }
    }
    if (a.mode == 0) {
      tmem_st16(tmem_addr(tAhi, 32 * warp, k0), hi);
      tmem_st16(tmem_addr(tAlo, 32 * warp, k0), lo);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        *reinterpret_cast<uint32_t*>(sAhi + canon_off(tid, k0 + j, K)) = hi[j];
        *reinterpret_cast<uint32_t*>(sAlo + canon_off(tid, k0 + j, K)) = lo[j];
      }
    }
  }
  if (a.mode == 0) tmem_st_wait();
  fence_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t idesc = idesc_tf32(128, N);
  const uint32_t lbo = a.swap ? 128u * (K / 4) : 128u, sbo = a.swap ? 128u : 128u * (K / 4);
  long long t0 = 0;
  if (tid == 0) {
    t0 = clock64();
    for (int rep = 0; rep < a.reps; ++rep) {
      const int npass = a.split3 ? 3 : 1;
      for (int pass = 0; pass < npass; ++pass) {
        // pass 0: hi*hi, 1: lo*hi, 2: hi*lo
        const uint32_t tA = (pass == 1) ? tAlo : tAhi;
        const unsigned char* sA = (pass == 1) ? sAlo : sAhi;
        const unsigned char* sB = (pass == 2) ? sBlo : sBhi;
        for (int ks = 0; ks < K / 8; ++ks) {
          const uint64_t db = smem_desc_noswizzle(smem_u32(sB) + ks * 256, lbo, sbo);
          const uint32_t acc = (pass | ks) ? 1u : 0u;
          // timing variant: consecutive MMAs go to different accumulators (columns [0,N), [N,2N), ...): separates the
          // latency of a dependent accumulation chain from the issue rate of the unit
          const uint32_t tDi = tD + (a.nacc > 1 ? ((pass * (K / 8) + ks) % a.nacc) * N : 0);
          if (a.mode == 0) mma_tf32_ts(tDi, tA + ks * 8, db, idesc, acc);
          else mma_tf32_ss(tDi, smem_desc_noswizzle(smem_u32(sA) + ks * 256, lbo, sbo), db, idesc, acc);
        }
      }
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  if (tid == 0 && a.cycles) *a.cycles = clock64() - t0;
  for (int n0 = 0; n0 < N; n0 += 32) {
    uint32_t r[32];
    tmem_ld32(tmem_addr(tD, 32 * warp, n0), r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) a.D[tid * N + n0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tb);
}

static float trunc_tf32(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xffffe000u;
  memcpy(&x, &u, 4);
  return x;
}

int main() {
  const int K = 64;
  int fails = 0;
  for (int N : {64, 128, 256}) {
    std::vector<float> A(128 * K), B(N * K), D(128 * N);
    srand(1);
    for (auto& x : A) x = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& x : B) x = (float)rand() / RAND_MAX * 2.f - 1.f;
    std::vector<double> ref(128 * N), reft(128 * N);
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        double s = 0, st = 0;
        for (int k = 0; k < K; ++k) {
          s += (double)A[m * K + k] * B[n * K + k];
          st += (double)trunc_tf32(A[m * K + k]) * trunc_tf32(B[n * K + k]);
        }
        ref[m * N + n] = s;
        reft[m * N + n] = st;
      }
    float *dA, *dB, *dD;
    long long* dC;
    cudaMalloc(&dA, A.size() * 4);
    cudaMalloc(&dB, B.size() * 4);
    cudaMalloc(&dD, D.size() * 4);
    cudaMalloc(&dC, 8);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = 2 * (size_t)N * K * 4 + 2 * 128 * K * 4;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int mode = 0; mode < 2; ++mode)
      for (int swap = 0; swap < 2; ++swap)
        for (int split3 = 0; split3 < 2; ++split3) {
          cudaMemset(dD, 0xff, D.size() * 4);
          ProbeArgs a{dA, dB, dD, N, K, mode, swap, split3, 1, 1, dC};
          k_probe<<<1, 128, smem>>>(a);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("N=%d mode=%d swap=%d split3=%d: CUDA error %s\n", N, mode, swap, split3, cudaGetErrorString(e));
            return 1;
          }
          cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
          double e_full = 0, e_trunc = 0;
          for (size_t i = 0; i < D.size(); ++i) {
            e_full = fmax(e_full, fabs(D[i] - ref[i]));
            e_trunc = fmax(e_trunc, fabs(D[i] - reft[i]));
          }
          printf("N=%3d %s swap=%d split3=%d  max|D-ref_fp64|=%.3e  max|D-ref_tf32trunc|=%.3e\n", N, mode ? "SS" : "TS", swap, split3,
                 e_full, e_trunc);
          if (swap == 0 && split3 == 1 && e_full > 2e-5) ++fails;
          if (swap == 0 && split3 == 0 && e_trunc > 2e-5) ++fails;
        }
    // issue rate: reps x (3 x 8) MMAs of 128 x N x 8
    for (int mode = 0; mode < 2; ++mode)
      for (int nacc = 1; nacc <= 3; ++nacc) {
        if (nacc * N > 256) continue;   // A lives in TMEM columns [256, 384)
        ProbeArgs a{dA, dB, dD, N, K, mode, 0, 1, 2000, nacc, dC};
        k_probe<<<1, 128, smem>>>(a);
        cudaDeviceSynchronize();
        long long cyc = 0;
        cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
        printf("N=%3d %s nacc=%d: %.1f cycles per 128x%dx8 tf32 MMA (2000 x 24 back to back)\n", N, mode ? "SS" : "TS", nacc,
               (double)cyc / (2000.0 * 24), N);
      }
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
  }
  printf(fails ? "PROBE FAILED (%d)\n" : "PROBE OK\n", fails);
  return fails ? 1 : 0;
}
