// hyena-b200: library state — error reporting, twiddle tables, init.
#include "hy_host.h"
#include "hy_fft.cuh"
#include <map>
#include <mutex>
#include <tuple>
#include <vector>
#include <cstdlib>

namespace hy {

static thread_local std::string g_err;
constexpr size_t kDefaultScratch = 2048ull << 20;
size_t g_scratch_budget = kDefaultScratch;  // scratch budget of the four-step path. Measured on B200 (profiles/): the phase
                                     // kernels are latency/issue-bound, not L2-bound — 24 MB .. 2 GB is monotonically
                                     // faster (fewer, fuller launches), so the default is simply 'large': 2 GB = all 256
                                     // rows of a d_model-256 layer at M = 2^20 in one group (1 GB: 11.86 ms per layer,
                                     // 2 GB: 11.74, tools/prof_conv.py 1000000 256 1).
int g_debug_block = 0;
int g_nstream = 1;
int g_persist_l2 = 0;
unsigned long long g_launches = 0;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}
int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#ifndef HY_EMU_BUILD
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(HY_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return HY_OK;
}
#endif

void* table_alloc(size_t bytes) {
#ifdef HY_EMU_BUILD
  return malloc(bytes);
#else
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) {
    set_error("cudaMalloc(%zu) for twiddle tables failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  return p;
#endif
}

// ---- SM clock probe: (clock64, globaltimer) around a short spin, so the host can derive the SM frequency the
// kernels of a step actually ran at without an NVML query (NVML polling stalls kernel launches for 100s of ms).
__global__ void k_clock_probe(unsigned long long* out) {
#if defined(__CUDA_ARCH__)
  unsigned long long t0, t1, c0 = clock64(), c1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do { c1 = clock64(); } while (c1 - c0 < 100000ull);     // ~50 us at 2 GHz
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  out[0] = c1 - c0;
  out[1] = t1 - t0;
#else
  out[0] = 0; out[1] = 1;
#endif
}

// ---- table builders (double precision trig, rounded once to fp32) -----------------------------
__global__ void k_build_tw(float2* tw, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    double s, c;
    sincospi(2.0 * (double)i / (double)n, &s, &c);
    tw[i] = make_float2((float)c, (float)-s);
  }
}
template <int S>
__global__ void k_build_twpos(float2* t) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < S) {
    int f = freq_of_pos<S>(p);
    double s, c;
    sincospi((double)f / (double)S, &s, &c);  // W_{2S}^f = exp(-2 pi i f / (2S))
    t[p] = make_float2((float)c, (float)-s);
  }
}
template <int M1>
__global__ void k_build_twV(float2* t, int T2, int M) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M1 * T2) {
    int pos1 = i / T2, l = i % T2;
    int k1 = freq_of_pos<M1>(pos1);
    long long e = ((long long)l * k1) % M;
    double s, c;
    sincospi(2.0 * (double)e / (double)M, &s, &c);
    t[i] = make_float2((float)c, (float)-s);
  }
}

static std::mutex g_mu;
#ifdef HY_EMU_BUILD
static int current_device() { return 0; }
#else
static int current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return d;
}
#endif
static std::map<int, Tables*> g_tables;
static std::map<std::tuple<int, int, int, int>, const float2*> g_twV;

template <int S>
static bool build_twpos(Tables* t) {
  float2* p = (float2*)table_alloc(sizeof(float2) * S);
  if (!p) return false;
  auto kern = k_build_twpos<S>;
  HY_LAUNCH(kern, (S + 255) / 256, 256, 0, 0, p);
  t->twpos[hy_ilog2(S)] = p;
  return true;
}

const Tables* tables() {
  std::lock_guard<std::mutex> lk(g_mu);
  int dev = current_device();
  auto it = g_tables.find(dev);
  if (it != g_tables.end()) return it->second;
  Tables* t = new Tables();
  for (auto& p : t->twpos) p = nullptr;
  float2* tw = (float2*)table_alloc(sizeof(float2) * HY_TWN);
  if (!tw) { delete t; return nullptr; }
  HY_LAUNCH(k_build_tw, (HY_TWN + 255) / 256, 256, 0, 0, tw, HY_TWN);
  t->tw = tw;
  bool ok = build_twpos<256>(t) && build_twpos<512>(t) && build_twpos<1024>(t) && build_twpos<2048>(t) &&
            build_twpos<4096>(t);
  if (!ok) { delete t; return nullptr; }
#ifndef HY_EMU_BUILD
  if (check_launch("twiddle table build") != HY_OK) { delete t; return nullptr; }
  if (cudaDeviceSynchronize() != cudaSuccess) {
    set_error("twiddle table build: %s", cudaGetErrorString(cudaGetLastError()));
    delete t;
    return nullptr;
  }
#endif
  g_tables[dev] = t;
  return t;
}

template <int M1>
static void launch_twV(float2* p, int T2, int M) {
  auto kern = k_build_twV<M1>;
  HY_LAUNCH(kern, (M1 * T2 + 255) / 256, 256, 0, 0, p, T2, M);
}

const float2* twV_table(int M1, int S, int T2) {
  std::lock_guard<std::mutex> lk(g_mu);
  int dev = current_device();
  auto key = std::make_tuple(dev, M1, S, T2);
  auto it = g_twV.find(key);
  if (it != g_twV.end()) return it->second;
  float2* p = (float2*)table_alloc(sizeof(float2) * M1 * T2);
  if (!p) return nullptr;
  const int M = M1 * S;
  switch (M1) {
#define HY_CASE(MM) case MM: launch_twV<MM>(p, T2, M); break;
    HY_COLS_POW2(HY_CASE)
    HY_COLS_ODD(HY_CASE)
#undef HY_CASE
    default: set_error("unsupported column length %d", M1); return nullptr;
  }
#ifndef HY_EMU_BUILD
  // table builds run on the legacy default stream; make them visible to every other stream
  if (cudaDeviceSynchronize() != cudaSuccess) {
    set_error("twV table build: %s", cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
#endif
  g_twV[key] = p;
  return p;
}

}  // namespace hy

extern "C" {

int hy_init(void) {
#ifndef HY_EMU_BUILD
  {
    std::lock_guard<std::mutex> lk(hy::g_mu);
    int cur = 0;
    if (cudaGetDevice(&cur) == cudaSuccess && hy::g_tables.count(cur)) return HY_OK;   // already initialised here
  }
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
    return hy::fail(HY_ERR_CUDA, "hy_init: no CUDA device (this library has no CPU fallback)");
  cudaDeviceProp prop;
  int dev = 0;
  cudaGetDevice(&dev);
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return hy::fail(HY_ERR_CUDA, "hy_init: cannot query device");
  if (prop.major < 10) return hy::fail(HY_ERR_UNSUPPORTED, "hy_init: built for sm_100a, found sm_%d%d", prop.major, prop.minor);
#endif
  return hy::tables() ? HY_OK : HY_ERR_CUDA;
}

const char* hy_last_error(void) {
  static thread_local std::string copy;
  copy = hy::g_err;
  return copy.c_str();
}

const char* hy_version(void) {
#ifdef HY_EMU_BUILD
  return "hyena-b200 0.1 (cpu-emulation TEST build)";
#else
  return "hyena-b200 0.1 (sm_100a)";
#endif
}

int hy_clock_probe(unsigned long long* out2, void* stream) {
  if (!out2) return hy::fail(HY_ERR_ARG, "hy_clock_probe: null output");
  HY_LAUNCH(hy::k_clock_probe, 1, 1, 0, stream, out2);
  return hy::check_launch("k_clock_probe");
}

unsigned long long hy_launch_count(void) { return __atomic_load_n(&hy::g_launches, __ATOMIC_RELAXED); }

int hy_set_scratch_budget(size_t bytes) {
  hy::g_scratch_budget = bytes ? bytes : hy::kDefaultScratch;
  return HY_OK;
}

int hy_set_pipeline(int nstream, size_t scratch_bytes) {
  if (nstream < 1 || nstream > 4) return hy::fail(HY_ERR_ARG, "hy_set_pipeline: nstream must be in [1, 4]");
  hy::g_nstream = nstream;
  if (scratch_bytes) hy::g_scratch_budget = scratch_bytes;
  return HY_OK;
}

// experiment hook (not in the public header): persisting-L2 access window over the in-flight scratch regions
int hy_debug_set_persist(int on) {
  hy::g_persist_l2 = on;
#ifndef HY_EMU_BUILD
  if (on) {
    int dev = 0, maxp = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, dev);
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)maxp);
    return maxp >> 20;
  }
#endif
  return 0;
}

// tests only (not declared in the public header): force the four-step path with row length S
int hy_debug_set_block(int S) {
  hy::g_debug_block = S;
  return HY_OK;
}

}  // extern "C"
