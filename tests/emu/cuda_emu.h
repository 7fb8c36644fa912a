// TEST INFRASTRUCTURE ONLY — never part of the shipped library.
//
// A tiny single-header CUDA execution-model emulator used by the "not gpu" test-suite to run the
// *same kernel source* that nvcc compiles for sm_100a on the CPU of the build container (which has
// no GPU).  Every CUDA thread of a block is a ucontext fiber; __syncthreads() / warp shuffles yield
// between fibers until the barrier condition holds.  Blocks run one after another (optionally
// spread over OS threads with OpenMP).  Nothing here is fast and nothing here is a fallback: the
// product loader (dna_b200/_lib.py) only ever loads the nvcc-built library.
#pragma once
#include <ucontext.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <vector>
#include <atomic>

#define HY_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct int2 { int x, y; };
struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
#define cudaSuccess 0

namespace emu {

struct Block {
  std::vector<ucontext_t> ctx;
  std::vector<char*> stacks;
  std::vector<char> done;
  ucontext_t sched;
  int nthreads = 0;
  int cur = -1;
  int live = 0;
  // block barrier
  int bar_arrived = 0;
  unsigned bar_gen = 0;
  // warp shuffle exchange: [warp][lane] 64-bit slots, double buffered by parity
  std::vector<uint64_t> shfl_slots;
  std::vector<int> warp_arrived;
  std::vector<unsigned> warp_gen;
  // shared memory
  std::vector<char> dyn_smem;
  std::map<int, std::vector<char>> static_smem;
  dim3 bdim, gdim, bidx;
  std::function<void()> body;
};

extern thread_local Block* g_blk;
extern thread_local uint3 g_tid;

inline void yield_to_sched() {
  Block* b = g_blk;
  int me = b->cur;
  swapcontext(&b->ctx[me], &b->sched);
}

inline void syncthreads() {
  Block* b = g_blk;
  unsigned gen = b->bar_gen;
  b->bar_arrived++;
  if (b->bar_arrived == b->live) {
    b->bar_arrived = 0;
    b->bar_gen++;
    return;
  }
  while (b->bar_gen == gen) yield_to_sched();
}

inline int linear_tid() {
  Block* b = g_blk;
  return g_tid.x + b->bdim.x * (g_tid.y + b->bdim.y * g_tid.z);
}

inline void warp_sync_internal(int warp, int nlanes) {
  Block* b = g_blk;
  unsigned gen = b->warp_gen[warp];
  b->warp_arrived[warp]++;
  if (b->warp_arrived[warp] == nlanes) {
    b->warp_arrived[warp] = 0;
    b->warp_gen[warp]++;
    return;
  }
  while (b->warp_gen[warp] == gen) yield_to_sched();
}

inline int warp_lanes(int warp) {
  Block* b = g_blk;
  int rem = b->nthreads - warp * 32;
  return rem < 32 ? rem : 32;
}

template <typename T>
inline T shfl_generic(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shfl payload");
  Block* b = g_blk;
  int t = linear_tid();
  int warp = t / 32, lane = t % 32;
  int nl = warp_lanes(warp);
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  b->shfl_slots[warp * 32 + lane] = raw;
  warp_sync_internal(warp, nl);
  uint64_t got = b->shfl_slots[warp * 32 + ((src_lane >= 0 && src_lane < nl) ? src_lane : lane)];
  warp_sync_internal(warp, nl);
  T out;
  memcpy(&out, &got, sizeof(T));
  return out;
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body);

inline void* dyn_smem() {
  uintptr_t p = (uintptr_t)g_blk->dyn_smem.data();
  p = (p + 127) & ~(uintptr_t)127;
  return (void*)p;
}
inline void* static_smem(int id, size_t bytes) {
  auto& v = g_blk->static_smem[id];
  if (v.size() < bytes) v.resize(bytes + 64);
  uintptr_t p = (uintptr_t)v.data();
  p = (p + 15) & ~(uintptr_t)15;
  return (void*)p;
}

}  // namespace emu

#define threadIdx (emu::g_tid)
#define blockIdx (emu::g_blk->bidx)
#define blockDim (emu::g_blk->bdim)
#define gridDim (emu::g_blk->gdim)

static inline void __syncthreads() { emu::syncthreads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) {
  int t = emu::linear_tid();
  emu::warp_sync_internal(t / 32, emu::warp_lanes(t / 32));
}
template <typename T>
static inline T __shfl_xor_sync(unsigned, T v, int lanemask) {
  return emu::shfl_generic(v, (emu::linear_tid() % 32) ^ lanemask);
}
template <typename T>
static inline T __shfl_down_sync(unsigned, T v, int d) {
  return emu::shfl_generic(v, (emu::linear_tid() % 32) + d);
}
template <typename T>
static inline T __shfl_up_sync(unsigned, T v, int d) {
  return emu::shfl_generic(v, (emu::linear_tid() % 32) - d);
}
template <typename T>
static inline T __shfl_sync(unsigned, T v, int src) {
  return emu::shfl_generic(v, src);
}
template <typename T>
static inline T __ldg(const T* p) { return *p; }
// funnel shift right: low 32 bits of ((hi:lo) >> (shift & 31))
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) {
  return (unsigned)(((((unsigned long long)hi) << 32) | lo) >> (shift & 31));
}

static inline float atomicAdd(float* p, float v) {
  std::atomic_ref<float> r(*p);
  return r.fetch_add(v);
}
static inline int atomicAdd(int* p, int v) {
  std::atomic_ref<int> r(*p);
  return r.fetch_add(v);
}
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
  std::atomic_ref<unsigned long long> r(*p);
  return r.fetch_add(v);
}
static inline unsigned atomicAdd(unsigned* p, unsigned v) {
  std::atomic_ref<unsigned> r(*p);
  return r.fetch_add(v);
}

static inline void sincospif(float x, float* s, float* c) {
  double a = (double)x * 3.14159265358979323846;
  *s = (float)std::sin(a);
  *c = (float)std::cos(a);
}
static inline void sincospi(double x, double* s, double* c) {
  double a = x * 3.14159265358979323846;
  *s = std::sin(a);
  *c = std::cos(a);
}
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline unsigned __brev(unsigned x) {
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i);
  return r;
}
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
