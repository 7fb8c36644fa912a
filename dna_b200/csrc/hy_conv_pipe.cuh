// hyena-b200: the four-step long convolution as ONE persistent launch per direction.
//
// The three phases of hy_conv.cuh (A: column transforms, B: row transforms + spectrum product, C: inverse column
// transforms + gating epilogue; reference math: src/models/sequence/hyena.py:60-92 and the backward of :44-57) hand a
// row's complex scratch (8 B x M, 8 MB at L = 1 M) from A to B to C.  As three launches over a 1 GB group of rows every
// hand-off is a DRAM round trip: 47.9 GB of measured traffic per layer against 8.7 GB of algorithmic bytes
// (profiles/r01p_*).  Small groups keep the hand-off in L2 but leave partial waves (profiles/r01f_*).  Here all CTAs
// stay resident and draw work items from one queue ordered so that, in steady state, step t offers
//       A tiles of row t,   B row pairs of row t-lag,   C tiles of row t-2 lag
// (384 items per step, lag 1 at M = 256 x 4096) over a ring of 2 lag + 2 row buffers: the live scratch is 3-4 rows (24-32 MB)
// and is read back while it is still in L2, and the machine always has a full mix of the three phases to run.
// Dependencies are per-row arrival counters: an item spins (thread 0, ld.acquire.gpu) until the producer phase of its
// row has arrived completely; producers publish with __threadfence + red.release after a CTA barrier.  Items are
// claimed in queue order and every dependency points to an EARLIER item, which a running CTA has already claimed, so
// the schedule cannot deadlock whatever the number of resident CTAs (the CPU emulator runs it with a handful).
#pragma once
#include "hy_conv.cuh"

enum { HY_PIPE_FWD = 0, HY_PIPE_BWDG = 1, HY_PIPE_SPEC = 2, HY_PIPE_DK = 3 };
constexpr int kPipeRingMax = 10;      // upper bound of 2 lag + 2 (lag <= 4): sizes the workspace
constexpr int kPipeCtlHead = 32;      // unsigned words before the counters (word 0: queue head)

HY_HD constexpr bool pipe_has_a(int kind) { return kind != HY_PIPE_DK; }
HY_HD constexpr bool pipe_has_c(int kind) { return kind != HY_PIPE_SPEC; }

// bytes of the control block for `nrows` rows: queue head + done[3][nrows]
HY_HD inline size_t pipe_ctl_bytes(long long nrows) { return sizeof(unsigned) * (kPipeCtlHead + 3 * (size_t)nrows); }

// spin until *cnt >= want; returns true when it had to wait at all
HY_DEVICE bool pipe_wait(const unsigned* cnt, unsigned want) {
  if (hy_ld_acquire(cnt) >= want) return false;
  do {
#if defined(__CUDA_ARCH__)
    __nanosleep(32);
#endif
  } while (hy_ld_acquire(cnt) < want);
  return true;
}
HY_DEVICE long long pipe_clock() {
#if defined(__CUDA_ARCH__)
  return clock64();
#else
  return 0;
#endif
}

// steps the B items of a row trail its A items (and the C items its B items): chosen so that a dependency lies at least
// ~380 queue positions back — more than the ~300 items in flight on 148 SMs x 2 CTAs — whatever the items per step
HY_HD constexpr int pipe_lag(int per) { return per >= 380 ? 1 : (per >= 190 ? 2 : (per >= 127 ? 3 : 4)); }
// row buffers in the ring: a buffer is reused once the C items of its previous row (2 lags back) have surely drained
HY_HD constexpr int pipe_ring(int per) { return 2 * pipe_lag(per) + 2; }

// stats (optional, device memory, 16 x u64): [2*phase] cycles thread 0 spent waiting on a dependency, [2*phase+1] items
// of that phase that had to wait, [6] items run, [7] cycles inside the phase bodies, [8] cycles in claim + publish
template <class DT, int M1, int T2, int S, int NT, int KIND, bool VEC, bool STG>
HY_DEVICE void conv_pipe_body(const ConvArgs& a, float2* ring, unsigned* ctl, unsigned long long* stats, const int LAG) {
  constexpr bool HAS_A = pipe_has_a(KIND), HAS_C = pipe_has_c(KIND);
  constexpr int nA = HAS_A ? S / T2 : 0, nB = M1 / 2, nC = HAS_C ? S / T2 : 0;
  constexpr int per = nA + nB + nC;
  const int RING = 2 * LAG + 2;   // a buffer is reused once the C items of its previous row (2 lags back) have drained
  constexpr int ROWMODE = KIND == HY_PIPE_FWD ? HY_PW_CONV : (KIND == HY_PIPE_BWDG ? HY_PW_BWDG : (KIND == HY_PIPE_SPEC ? HY_PW_SPEC : HY_PW_REPACK));
  const long long M = (long long)M1 * S;
  const int nrows = a.nrows;
  const unsigned total = (unsigned)(nrows + 2 * LAG) * per;
  unsigned* doneA = ctl + kPipeCtlHead;
  unsigned* doneB = doneA + nrows;
  unsigned* doneC = doneB + nrows;
  HY_STATIC_SMEM(unsigned, s_it, 4);
  const int tid = threadIdx.x;
  unsigned nxt = 0;
  if (tid == 0) nxt = atomicAdd(ctl, 1u);
  for (;;) {
    __syncthreads();                        // the previous item's shared memory (and s_it) is no longer in use
    if (tid == 0) s_it[0] = nxt;
    __syncthreads();
    const unsigned it = s_it[0];
    if (it >= total) break;
    // claim the following item now: the atomic's round trip hides behind this item's body (thread 0 only consumes it
    // at the top of the next iteration)
    if (tid == 0) nxt = atomicAdd(ctl, 1u);
    const int step = (int)(it / per), off = (int)(it % per);
    int phase, row, idx;
    if (off < nA) { phase = 0; row = step; idx = off; }
    else if (off < nA + nB) { phase = 1; row = step - LAG; idx = off - nA; }
    else { phase = 2; row = step - 2 * LAG; idx = off - nA - nB; }
    if (row < 0 || row >= nrows) continue;
    float2* scr = ring + (long long)(row % RING) * M;
    long long tw0 = 0;
    if (tid == 0) {
      tw0 = pipe_clock();
      bool waited = false;
      if (phase == 0) {                     // the ring buffer must have been drained by its previous row
        if (row >= RING) waited = pipe_wait((HAS_C ? doneC : doneB) + row - RING, HAS_C ? nC : nB);
      } else if (phase == 1) {
        if (HAS_A) waited = pipe_wait(doneA + row, nA);
        else if (row >= RING) waited = pipe_wait(doneC + row - RING, nC);
      } else {
        waited = pipe_wait(doneB + row, nB);
      }
      const long long tw1 = pipe_clock();
      if (stats != nullptr) {
        if (waited) {
          atomicAdd(stats + 2 * phase, (unsigned long long)(tw1 - tw0));
          atomicAdd(stats + 2 * phase + 1, 1ull);
        }
        atomicAdd(stats + 6, 1ull);
      }
      tw0 = tw1;
    }
    __syncthreads();
    if (phase == 0) {
      if constexpr (HAS_A) col_fwd_body<DT, M1, T2, NT, 1, VEC, STG, KIND == HY_PIPE_BWDG>(a, idx, row, scr);
    } else if (phase == 1) {
      row_conv_body<S, NT, ROWMODE, true>(a, idx, row, scr);
    } else {
      if constexpr (HAS_C) col_inv_body<DT, M1, T2, NT, 1, KIND == HY_PIPE_BWDG ? 1 : 0, VEC, STG, true>(a, idx, row, scr);
    }
    const long long tb1 = (tid == 0 && stats != nullptr) ? pipe_clock() : 0;
    // publish (the pattern of a cooperative-groups grid barrier): the CTA barrier orders every thread's writes before
    // thread 0's gpu-scope fence + release, which is cumulative over them
    __syncthreads();
    if (tid == 0) {
      hy_threadfence();
      hy_red_release((phase == 0 ? doneA : (phase == 1 ? doneB : doneC)) + row, 1u);
      if (stats != nullptr) {
        atomicAdd(stats + 7, (unsigned long long)(tb1 - tw0));
        atomicAdd(stats + 8, (unsigned long long)(pipe_clock() - tb1));
      }
    }
  }
}

template <class DT, int M1, int T2, int S, int NT, int KIND>
__global__ void __launch_bounds__(NT, 2) k_conv_pipe(ConvArgs a, float2* ring, unsigned* ctl, unsigned long long* stats, int lag) {
  if constexpr (DT::kBf16) {
    conv_pipe_body<DT, M1, T2, S, NT, KIND, true, true>(a, ring, ctl, stats, lag);     // host guarantees stage_ok
  } else {
    conv_pipe_body<DT, M1, T2, S, NT, KIND, true, false>(a, ring, ctl, stats, lag);    // host guarantees vec_all
  }
}
