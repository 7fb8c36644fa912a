// hyena-b200: host-side plumbing shared by the .cu translation units (errors, launches, tables).
#pragma once
#include "hy_common.cuh"
#include "../../include/hyena_b200.h"
#include <string>
#include <cstdio>
#include <cstdarg>

namespace hy {

void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

#ifdef HY_EMU_BUILD
inline int check_launch(const char*) { return HY_OK; }
#else
int check_launch(const char* what);
#endif

// device allocation for library-owned constant tables only (never for user data)
void* table_alloc(size_t bytes);

struct Tables {
  const float2* tw;                 // W_8192^i
  const float2* twpos[14];          // index log2(S): W_{2S}^{freq_S(p)}
};
// returns nullptr (and sets the error) on failure
const Tables* tables();
// W_M^{l * k1(pos1)} table for the four-step split, M = M1 * S; cached per (M1, S, T2)
const float2* twV_table(int M1, int S, int T2);

extern size_t g_scratch_budget;
extern int g_persist_l2;  // experimental: pin the in-flight scratch regions in L2 (access-policy window)
extern int g_nstream;   // row groups of the four-step path in flight on internal streams (1 = caller's stream only)
extern unsigned long long g_launches;  // kernels launched by this library (all threads)
extern int g_debug_block;  // tests only: force the four-step path with this row length (0 = off)

}  // namespace hy

// ---- launch helper ---------------------------------------------------------------------------
#ifdef HY_EMU_BUILD
#define HY_LAUNCH(kern, grid, block, smem, stream, ...) \
  do {                                                    \
    __atomic_add_fetch(&hy::g_launches, 1ull, __ATOMIC_RELAXED); \
    emu::launch(dim3(grid), dim3(block), (smem), [=]() { kern(__VA_ARGS__); }); \
  } while (0)
#else
template <class K>
inline void hy_set_smem(K kern, size_t smem) {
  // one attribute call per (kernel, device, size): the attribute is sticky
  static thread_local K last_kern = nullptr;
  static thread_local size_t last_smem = 0;
  static thread_local int last_dev = -1;
  if (smem <= 48 * 1024) return;
  int dev = 0;
  cudaGetDevice(&dev);
  if (kern == last_kern && smem == last_smem && dev == last_dev) return;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  last_kern = kern; last_smem = smem; last_dev = dev;
}
#define HY_LAUNCH(kern, grid, block, smem, stream, ...)              \
  do {                                                               \
    hy_set_smem(kern, (smem));                                       \
    __atomic_add_fetch(&hy::g_launches, 1ull, __ATOMIC_RELAXED);     \
    kern<<<dim3(grid), dim3(block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__); \
  } while (0)
#endif
