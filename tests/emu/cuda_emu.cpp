// TEST INFRASTRUCTURE ONLY — fiber scheduler behind tests/emu/cuda_emu.h (see that header).
#include "cuda_emu.h"

namespace emu {

thread_local Block* g_blk = nullptr;
thread_local uint3 g_tid = {0, 0, 0};

static const size_t kStack = 256 * 1024;

struct Tramp { Block* b; int t; };
static thread_local std::vector<Tramp>* g_tramps = nullptr;

static void fiber_entry(unsigned lo, unsigned hi) {
  uintptr_t p = ((uintptr_t)hi << 32) | (uintptr_t)lo;
  Tramp* tr = (Tramp*)p;
  Block* b = tr->b;
  b->body();
  b->done[tr->t] = 1;
  b->live--;
  // a finished thread no longer takes part in barriers: release a barrier that is now complete
  if (b->live > 0 && b->bar_arrived == b->live) {
    b->bar_arrived = 0;
    b->bar_gen++;
  }
  swapcontext(&b->ctx[tr->t], &b->sched);
}

static void run_block(Block& b) {
  int n = b.nthreads;
  b.ctx.resize(n);
  b.done.assign(n, 0);
  b.live = n;
  b.bar_arrived = 0;
  b.bar_gen = 0;
  int nwarps = (n + 31) / 32;
  b.shfl_slots.assign(nwarps * 32, 0);
  b.warp_arrived.assign(nwarps, 0);
  b.warp_gen.assign(nwarps, 0);
  std::vector<Tramp> tramps(n);
  if ((int)b.stacks.size() < n) {
    size_t old = b.stacks.size();
    b.stacks.resize(n);
    for (int i = (int)old; i < n; ++i) b.stacks[i] = (char*)malloc(kStack);
  }
  for (int t = 0; t < n; ++t) {
    tramps[t] = Tramp{&b, t};
    getcontext(&b.ctx[t]);
    b.ctx[t].uc_stack.ss_sp = b.stacks[t];
    b.ctx[t].uc_stack.ss_size = kStack;
    b.ctx[t].uc_link = &b.sched;
    uintptr_t p = (uintptr_t)&tramps[t];
    makecontext(&b.ctx[t], (void (*)())fiber_entry, 2, (unsigned)(p & 0xffffffffu), (unsigned)(p >> 32));
  }
  g_blk = &b;
  long spins = 0;
  while (b.live > 0) {
    for (int t = 0; t < n; ++t) {
      if (b.done[t]) continue;
      b.cur = t;
      g_tid.x = t % b.bdim.x;
      g_tid.y = (t / b.bdim.x) % b.bdim.y;
      g_tid.z = t / (b.bdim.x * b.bdim.y);
      swapcontext(&b.sched, &b.ctx[t]);
    }
    if (++spins > 50000000L) {
      fprintf(stderr, "cuda_emu: block (%u,%u,%u) appears deadlocked\n", b.bidx.x, b.bidx.y, b.bidx.z);
      abort();
    }
  }
  g_blk = nullptr;
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
  long nblocks = (long)grid.x * grid.y * grid.z;
#pragma omp parallel
  {
    Block b;
    b.bdim = block;
    b.gdim = grid;
    b.nthreads = block.x * block.y * block.z;
    b.body = body;
    b.dyn_smem.resize(smem_bytes + 1024);
#pragma omp for schedule(dynamic, 1)
    for (long i = 0; i < nblocks; ++i) {
      b.bidx.x = (unsigned)(i % grid.x);
      b.bidx.y = (unsigned)((i / grid.x) % grid.y);
      b.bidx.z = (unsigned)(i / ((long)grid.x * grid.y));
      b.static_smem.clear();
      run_block(b);
    }
    for (char* s : b.stacks) free(s);
  }
}

}  // namespace emu
