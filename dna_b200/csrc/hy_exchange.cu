// hyena-b200: the exchange step of the channel partition over NVLink peer memory.
//
// One long sequence is split over the G GPUs of a box (BASELINE.json configs[3]): everything
// per-position runs on a rank's sequence chunk, the operator core (per-channel) on its channel slab.  Between the two
// sits a transposing all-to-all: rank r needs, from EVERY peer j, the rows of its slab restricted to j's chunk (or the
// inverse).  NCCL's all_to_all_single wants that packed — one strided copy before and one after, each a full pass over
// the tensor.  Here every rank keeps its data where the producer left it, in a buffer the peers have mapped (CUDA IPC,
// NVLink / NVSwitch), and ONE kernel per rank PULLS the rows it owns straight into their final layout with 16-byte
// loads over NVLink: no pack, no unpack, no staging.  The same kernel gathers (all-gather of the filter trunk's
// activations) and sums over peers (reduce-scatter of their gradients).
//
// Synchronisation is two epoch counters per peer pair in the mapped flag block, no host involvement:
//   ready[j] in rank r's block: peer j's source buffer holds epoch e   (written by j's pull kernel prologue — stream
//            order on j guarantees its producer has finished; r's blocks spin on it before the first remote load)
//   done[j]  in rank r's block: peer j has finished reading r's buffer of epoch e   (written by j's last block; r waits
//            for it — hy_peer_wait_done — before the buffer is overwritten two exchanges later)
// Spins are bounded (about 4 s): on expiry the kernel raises an error word instead of hanging the GPU.
#include "hy_host.h"
#include <cstdlib>
#ifndef HY_EMU_BUILD
#include "hy_tc05.cuh"
#endif

namespace hy {

constexpr int kPeerMax = 8;
// flag block layout (unsigned words): [0, 8) ready, [8, 16) done, [16] error, [17] block counter, [18] epoch whose handshake passed
constexpr int kFlagReady = 0, kFlagDone = kPeerMax, kFlagErr = 2 * kPeerMax, kFlagCount = 2 * kPeerMax + 1, kFlagOk = 2 * kPeerMax + 2;

struct PeerArgs {
  const char* src[kPeerMax];     // peer j's source buffer (this process's mapping; src[self] is local)
  unsigned* flags[kPeerMax];     // peer j's flag block
  int G, self;
  unsigned epoch;
  int reduce;                    // 0: dst[j-th slot] = src_j ; 1: dst = sum_j src_j (fp32)
  int n_outer, n_inner;
  long long row_bytes;
  long long src_base, src_outer, src_inner;          // bytes, inside every peer's buffer
  long long dst_outer, dst_inner, dst_peer;          // bytes
  char* dst;
};

#ifndef HY_EMU_BUILD

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// epoch counters wrap: a >= b in modular arithmetic
__device__ __forceinline__ bool epoch_reached(unsigned have, unsigned want) { return (int)(have - want) >= 0; }

__device__ bool spin_until(const unsigned* p, unsigned want, unsigned* err) {
  const long long t0 = clock64();
  while (!epoch_reached(ld_acquire_sys(p), want)) {
    __nanosleep(200);
    if (clock64() - t0 > 8000000000ll) {       // ~4 s at 2 GHz: a peer never arrived
      atomicExch(err, 1u);
      return false;
    }
  }
  return true;
}

template <int VEC>
struct VecT;
template <> struct VecT<16> { using type = uint4; };
template <> struct VecT<4> { using type = unsigned; };
template <> struct VecT<2> { using type = unsigned short; };

// The handshake runs as its own one-warp kernel in front of the copy: a CTA that spins must not hold the shared memory
// or the registers other streams' kernels need to reach THEIR handshake (two exchange lanes on two streams would
// otherwise wait on each other through the SM resources).  Stream order: this rank's producer is complete -> tell every
// peer its source of this epoch is readable, then wait until every peer has said the same.  ok[0] = 0 on a timeout.
__global__ void k_peer_handshake(PeerArgs a, unsigned* ok) {
  unsigned* mine = a.flags[a.self];
  bool good = true;
  if (threadIdx.x < a.G && (int)threadIdx.x != a.self) {
    __threadfence_system();
    st_release_sys(a.flags[threadIdx.x] + kFlagReady + a.self, a.epoch);
    good = spin_until(mine + kFlagReady + threadIdx.x, a.epoch, mine + kFlagErr);
  }
  good = __all_sync(0xffffffffu, good);
  if (threadIdx.x == 0) *ok = good ? a.epoch : a.epoch - 1;
}

template <int VEC>
__global__ void __launch_bounds__(512, 2) k_peer_pull(PeerArgs a) {
  using V = typename VecT<VEC>::type;
  unsigned* mine = a.flags[a.self];
  const bool s_ok = mine[kFlagOk] == a.epoch;       // written by this exchange's handshake kernel
  if (s_ok) {
    const long long vec_per_row = a.row_bytes / VEC;
    constexpr int SEG = 2048;                                   // vectors per work item (32 KB at 16 B)
    const long long segs = (vec_per_row + SEG - 1) / SEG;
    const long long rows = (long long)a.n_outer * a.n_inner;
    if (!a.reduce) {
      const long long items = rows * segs * a.G;
      for (long long it = blockIdx.x; it < items; it += gridDim.x) {
        // peers innermost: consecutive CTAs read from different peers, spreading the load over the switch
        const int j = (int)((it + a.self) % a.G);
        const long long rs = it / a.G;
        const long long row = rs / segs, seg = rs % segs;
        const long long o = row / a.n_inner, i = row % a.n_inner;
        const V* s = reinterpret_cast<const V*>(a.src[j] + a.src_base + o * a.src_outer + i * a.src_inner);
        V* d = reinterpret_cast<V*>(a.dst + o * a.dst_outer + i * a.dst_inner + j * a.dst_peer);
        const long long v0 = seg * SEG, v1 = (v0 + SEG < vec_per_row) ? v0 + SEG : vec_per_row;
        for (long long v = v0 + threadIdx.x; v < v1; v += 4 * blockDim.x) {
          V r[4];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (v + q * blockDim.x < v1) r[q] = s[v + q * blockDim.x];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (v + q * blockDim.x < v1) d[v + q * blockDim.x] = r[q];
        }
      }
    } else {
      // dst[x] = sum over peers of src_j[x], fp32 (row_bytes % 16 == 0 checked by the host for VEC == 16)
      const long long items = rows * segs;
      for (long long it = blockIdx.x; it < items; it += gridDim.x) {
        const long long row = it / segs, seg = it % segs;
        const long long o = row / a.n_inner, i = row % a.n_inner;
        const long long off = a.src_base + o * a.src_outer + i * a.src_inner;
        float4* d = reinterpret_cast<float4*>(a.dst + o * a.dst_outer + i * a.dst_inner);
        const long long v0 = seg * SEG, v1 = (v0 + SEG < vec_per_row) ? v0 + SEG : vec_per_row;
        for (long long v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          float4 r[kPeerMax];
#pragma unroll
          for (int j = 0; j < kPeerMax; ++j)
            if (j < a.G) r[j] = reinterpret_cast<const float4*>(a.src[j] + off)[v];
#pragma unroll
          for (int j = 0; j < kPeerMax; ++j)      // fixed peer order: every rank sums in the same order
            if (j < a.G) { acc.x += r[j].x; acc.y += r[j].y; acc.z += r[j].z; acc.w += r[j].w; }
          d[v] = acc;
        }
      }
    }
  }
  // last block out: tell every peer this rank no longer reads its buffer of this epoch
  __syncthreads();
  __shared__ unsigned s_last;
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(mine + kFlagCount, 1u) == gridDim.x - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) {
    if (threadIdx.x == 0) mine[kFlagCount] = 0;
    if (threadIdx.x < a.G && (int)threadIdx.x != a.self) st_release_sys(a.flags[threadIdx.x] + kFlagDone + a.self, a.epoch);
  }
}

// ---- the same copy driven by the TMA unit: ONE thread per CTA keeps kStages bulk loads (peer memory -> shared memory,
// completion on an mbarrier) and bulk stores (shared -> local global memory) in flight; no registers, no LSU traffic,
// and the 148 x 4 x 32 KB of outstanding requests cover the NVLink round trip.
constexpr int kBulkStages = 4;
constexpr int kBulkBytes = 32768;

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   tc05::smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(tc05::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(tc05::smem_u32(smem_src)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__global__ void __launch_bounds__(32, 1) k_peer_pull_bulk(PeerArgs a) {
  extern __shared__ __align__(128) unsigned char bulk_smem[];
  __shared__ __align__(8) uint64_t bars[kBulkStages];
  unsigned* mine = a.flags[a.self];
  const bool s_ok = mine[kFlagOk] == a.epoch;       // written by this exchange's handshake kernel
  if (threadIdx.x == 0) {
    for (int s = 0; s < kBulkStages; ++s) tc05::mbar_init(&bars[s], 1);
    tc05::mbar_fence_init();
  }
  __syncwarp();
  if (s_ok && threadIdx.x == 0) {
    const long long segs = (a.row_bytes + kBulkBytes - 1) / kBulkBytes;
    const long long rows = (long long)a.n_outer * a.n_inner;
    const long long items = rows * segs * a.G;
    const long long n_my = items > blockIdx.x ? (items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto locate = [&](long long i, const char*& src, char*& dst, unsigned& bytes) {
      const long long it = blockIdx.x + i * gridDim.x;
      const int j = (int)((it + a.self) % a.G);          // peers innermost: neighbouring CTAs read from different peers
      const long long rs = it / a.G;
      const long long row = rs / segs, seg = rs % segs;
      const long long o = row / a.n_inner, ii = row % a.n_inner;
      const long long b0 = seg * kBulkBytes;
      bytes = (unsigned)((a.row_bytes - b0 < kBulkBytes) ? a.row_bytes - b0 : kBulkBytes);
      src = a.src[j] + a.src_base + o * a.src_outer + ii * a.src_inner + b0;
      dst = a.dst + o * a.dst_outer + ii * a.dst_inner + j * a.dst_peer + b0;
    };
    for (long long i = 0; i < n_my + (kBulkStages - 1); ++i) {
      if (i < n_my) {
        const int s = (int)(i % kBulkStages);
        if (i >= kBulkStages) bulk_wait_read_all();       // the store that last used this stage has drained its smem
        const char* src; char* dst; unsigned bytes;
        locate(i, src, dst, bytes);
        tc05::mbar_arrive_expect_tx(&bars[s], bytes);
        bulk_load(bulk_smem + (size_t)s * kBulkBytes, src, bytes, &bars[s]);
      }
      const long long k = i - (kBulkStages - 1);
      if (k >= 0 && k < n_my) {
        const int s = (int)(k % kBulkStages);
        tc05::mbar_wait(&bars[s], (unsigned)((k / kBulkStages) & 1));
        const char* src; char* dst; unsigned bytes;
        locate(k, src, dst, bytes);
        bulk_store(dst, bulk_smem + (size_t)s * kBulkBytes, bytes);
      }
    }
    bulk_wait_all();                                       // writes complete before the arrival below
  }
  __syncwarp();
  __shared__ unsigned s_last;
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(mine + kFlagCount, 1u) == gridDim.x - 1 ? 1u : 0u;
  }
  __syncwarp();
  if (s_last) {
    if (threadIdx.x == 0) mine[kFlagCount] = 0;
    if (threadIdx.x < a.G && (int)threadIdx.x != a.self) st_release_sys(a.flags[threadIdx.x] + kFlagDone + a.self, a.epoch);
  }
}

__global__ void k_peer_wait_done(unsigned* mine, int G, int self, unsigned epoch) {
  if (threadIdx.x < G && (int)threadIdx.x != self) spin_until(mine + kFlagDone + threadIdx.x, epoch, mine + kFlagErr);
}

#endif  // !HY_EMU_BUILD

}  // namespace hy

namespace hy {
// 0 = LSU copy (k_peer_pull), 1 = TMA bulk copy (k_peer_pull_bulk); HYENA_B200_PEER_MODE / hy_debug_set_peer_mode
static int g_peer_mode = [] {
  const char* e = getenv("HYENA_B200_PEER_MODE");
  return e ? atoi(e) : 1;
}();
}  // namespace hy

using namespace hy;

extern "C" {

int hy_debug_set_peer_mode(int mode) {
  hy::g_peer_mode = mode;
  return mode;
}

size_t hy_peer_flag_bytes(void) { return 256; }

// cudaMalloc + zero + IPC handle (64 bytes).  Library-owned memory: the exchange buffers must outlive torch's caching
// allocator decisions and be mappable by the peers, so they are the one place this library allocates user-visible memory.
int hy_peer_alloc(size_t bytes, void** ptr, void* handle64) {
#ifdef HY_EMU_BUILD
  (void)bytes; (void)ptr; (void)handle64;
  return fail(HY_ERR_UNSUPPORTED, "hy_peer_alloc: peer memory needs CUDA devices");
#else
  if (!ptr || !handle64 || bytes == 0) return fail(HY_ERR_ARG, "hy_peer_alloc: bad argument");
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) return fail(HY_ERR_CUDA, "hy_peer_alloc: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
  if (cudaMemset(p, 0, bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
    cudaFree(p);
    return fail(HY_ERR_CUDA, "hy_peer_alloc: clearing the buffer failed");
  }
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) {
    cudaFree(p);
    return fail(HY_ERR_CUDA, "hy_peer_alloc: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  static_assert(sizeof(h) == 64, "IPC handle size");
  memcpy(handle64, &h, 64);
  *ptr = p;
  return HY_OK;
#endif
}

int hy_peer_open(const void* handle64, void** ptr) {
#ifdef HY_EMU_BUILD
  (void)handle64; (void)ptr;
  return fail(HY_ERR_UNSUPPORTED, "hy_peer_open: peer memory needs CUDA devices");
#else
  if (!ptr || !handle64) return fail(HY_ERR_ARG, "hy_peer_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
    return fail(HY_ERR_CUDA, "hy_peer_open: cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(cudaGetLastError()));
  *ptr = p;
  return HY_OK;
#endif
}

int hy_peer_close(void* ptr) {
#ifndef HY_EMU_BUILD
  if (ptr && cudaIpcCloseMemHandle(ptr) != cudaSuccess) return fail(HY_ERR_CUDA, "hy_peer_close failed");
#else
  (void)ptr;
#endif
  return HY_OK;
}

int hy_peer_free(void* ptr) {
#ifndef HY_EMU_BUILD
  if (ptr && cudaFree(ptr) != cudaSuccess) return fail(HY_ERR_CUDA, "hy_peer_free failed");
#else
  (void)ptr;
#endif
  return HY_OK;
}

// Pull (or sum) rows from the peers' mapped buffers into `dst`; see the header for the geometry.
int hy_peer_pull(const hy_peer_pull_args* p, void* stream) {
#ifdef HY_EMU_BUILD
  (void)p; (void)stream;
  return fail(HY_ERR_UNSUPPORTED, "hy_peer_pull: peer memory needs CUDA devices");
#else
  if (!p || p->G < 1 || p->G > kPeerMax || p->self < 0 || p->self >= p->G || !p->dst || p->row_bytes <= 0 || p->n_outer < 1 ||
      p->n_inner < 1)
    return fail(HY_ERR_ARG, "hy_peer_pull: bad argument");
  PeerArgs a;
  for (int j = 0; j < kPeerMax; ++j) {
    a.src[j] = j < p->G ? reinterpret_cast<const char*>(p->src[j]) : nullptr;
    a.flags[j] = j < p->G ? reinterpret_cast<unsigned*>(p->flags[j]) : nullptr;
    if (j < p->G && (!a.src[j] || !a.flags[j])) return fail(HY_ERR_ARG, "hy_peer_pull: null peer pointer");
  }
  a.G = p->G; a.self = p->self; a.epoch = p->epoch; a.reduce = p->reduce;
  a.n_outer = p->n_outer; a.n_inner = p->n_inner; a.row_bytes = p->row_bytes;
  a.src_base = p->src_base; a.src_outer = p->src_outer; a.src_inner = p->src_inner;
  a.dst_outer = p->dst_outer; a.dst_inner = p->dst_inner; a.dst_peer = p->dst_peer;
  a.dst = reinterpret_cast<char*>(p->dst);
  long long all = a.row_bytes | a.src_base | a.src_outer | a.src_inner | a.dst_outer | a.dst_inner | a.dst_peer |
                  (long long)reinterpret_cast<uintptr_t>(a.dst);
  for (int j = 0; j < a.G; ++j) all |= (long long)reinterpret_cast<uintptr_t>(a.src[j]);
  const int vec = (all % 16 == 0) ? 16 : ((all % 4 == 0) ? 4 : 2);
  if (all % 2) return fail(HY_ERR_ARG, "hy_peer_pull: rows must be 2-byte aligned");
  if (a.reduce && vec != 16) return fail(HY_ERR_ARG, "hy_peer_pull: the summing form needs 16-byte aligned fp32 rows");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // max_ctas: an exchange running beside compute kernels on another stream keeps to a few SMs (the link, not the SM
  // count, bounds it: 32 CTAs x 4 x 32 KB of requests in flight already cover the NVLink round trip)
  HY_LAUNCH(k_peer_handshake, 1, 32, 0, stream, a, a.flags[a.self] + kFlagOk);
  if (p->max_ctas > 0 && p->max_ctas < sms) sms = p->max_ctas;
  const int grid = sms * 2;
  if (vec == 16 && !a.reduce && g_peer_mode == 1) {
    const size_t smem = (size_t)kBulkStages * kBulkBytes;
    HY_LAUNCH(k_peer_pull_bulk, sms, 32, smem, stream, a);
    return check_launch("k_peer_pull_bulk");
  }
  if (vec == 16) HY_LAUNCH(k_peer_pull<16>, grid, 512, 0, stream, a);
  else if (vec == 4) HY_LAUNCH(k_peer_pull<4>, grid, 512, 0, stream, a);
  else HY_LAUNCH(k_peer_pull<2>, grid, 512, 0, stream, a);
  return check_launch("k_peer_pull");
#endif
}

// Block the stream until every peer has finished reading this rank's buffer of `epoch`.
int hy_peer_wait_done(void* my_flags, int G, int self, unsigned epoch, void* stream) {
#ifdef HY_EMU_BUILD
  (void)my_flags; (void)G; (void)self; (void)epoch; (void)stream;
  return fail(HY_ERR_UNSUPPORTED, "hy_peer_wait_done: peer memory needs CUDA devices");
#else
  if (!my_flags || G < 1 || G > kPeerMax) return fail(HY_ERR_ARG, "hy_peer_wait_done: bad argument");
  HY_LAUNCH(k_peer_wait_done, 1, 32, 0, stream, reinterpret_cast<unsigned*>(my_flags), G, self, epoch);
  return check_launch("k_peer_wait_done");
#endif
}

// 0 = no spin of this rank's exchange kernels has timed out so far (synchronises the device)
int hy_peer_error(void* my_flags) {
#ifdef HY_EMU_BUILD
  (void)my_flags;
  return 0;
#else
  unsigned e = 0;
  if (cudaMemcpy(&e, reinterpret_cast<unsigned*>(my_flags) + kFlagErr, sizeof(e), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int)e;
#endif
}

}  // extern "C"
