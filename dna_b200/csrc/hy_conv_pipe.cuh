// hyena-b200: the four-step long convolution as ONE persistent launch per direction.
//
// The three phases of hy_conv.cuh (A: column transforms, B: row transforms + spectrum product, C: inverse column
// transforms + gating epilogue; reference math: src/models/sequence/hyena.py:60-92 and the backward of :44-57) hand a
// row's complex scratch (8 B x M, 8 MB at L = 1 M) from A to B to C.  As three launches over a 1 GB group of rows every
// hand-off is a DRAM round trip: 47.9 GB of measured traffic per layer against 8.7 GB of algorithmic bytes
// (profiles/r01p_*).  Small groups keep the hand-off in L2 but leave partial waves (profiles/r01f_*).  Here all CTAs
// stay resident and draw work items from one queue ordered so that, in steady state, step t offers
//       A tiles of row t,   B row pairs of row t-1,   C tiles of row t-2
// (384 items per step at M = 256 x 4096) over a ring of kRing row buffers: the live scratch is 3-4 rows (24-32 MB)
// and is read back while it is still in L2, and the machine always has a full mix of the three phases to run.
// Dependencies are per-row arrival counters: an item spins (thread 0, ld.acquire.gpu) until the producer phase of its
// row has arrived completely; producers publish with __threadfence + red.release after a CTA barrier.  Items are
// claimed in queue order and every dependency points to an EARLIER item, which a running CTA has already claimed, so
// the schedule cannot deadlock whatever the number of resident CTAs (the CPU emulator runs it with a handful).
#pragma once
#include "hy_conv.cuh"

enum { HY_PIPE_FWD = 0, HY_PIPE_BWDG = 1, HY_PIPE_SPEC = 2, HY_PIPE_DK = 3 };
constexpr int kPipeRing = 4;          // row buffers in the ring: A(t) reuses the buffer C(t-4) drained two steps ago
constexpr int kPipeCtlHead = 32;      // unsigned words before the counters (word 0: queue head)

HY_HD constexpr bool pipe_has_a(int kind) { return kind != HY_PIPE_DK; }
HY_HD constexpr bool pipe_has_c(int kind) { return kind != HY_PIPE_SPEC; }

// bytes of the control block for `nrows` rows: queue head + done[3][nrows]
HY_HD inline size_t pipe_ctl_bytes(long long nrows) { return sizeof(unsigned) * (kPipeCtlHead + 3 * (size_t)nrows); }

HY_DEVICE void pipe_wait(const unsigned* cnt, unsigned want) {
  while (hy_ld_acquire(cnt) < want) {
#if defined(__CUDA_ARCH__)
    __nanosleep(64);
#endif
  }
}

template <class DT, int M1, int T2, int S, int NT, int KIND, bool VEC, bool STG>
HY_DEVICE void conv_pipe_body(const ConvArgs& a, float2* ring, unsigned* ctl) {
  constexpr bool HAS_A = pipe_has_a(KIND), HAS_C = pipe_has_c(KIND);
  constexpr int nA = HAS_A ? S / T2 : 0, nB = M1 / 2, nC = HAS_C ? S / T2 : 0;
  constexpr int per = nA + nB + nC;
  constexpr int ROWMODE = KIND == HY_PIPE_FWD ? HY_PW_CONV : (KIND == HY_PIPE_BWDG ? HY_PW_BWDG : (KIND == HY_PIPE_SPEC ? HY_PW_SPEC : HY_PW_REPACK));
  const long long M = (long long)M1 * S;
  const int nrows = a.nrows;
  const unsigned total = (unsigned)(nrows + 2) * per;
  unsigned* doneA = ctl + kPipeCtlHead;
  unsigned* doneB = doneA + nrows;
  unsigned* doneC = doneB + nrows;
  HY_STATIC_SMEM(unsigned, s_it, 4);
  const int tid = threadIdx.x;
  for (;;) {
    __syncthreads();                        // the previous item's shared memory (and s_it) is no longer in use
    if (tid == 0) s_it[0] = atomicAdd(ctl, 1u);
    __syncthreads();
    const unsigned it = s_it[0];
    if (it >= total) break;
    const int step = (int)(it / per), off = (int)(it % per);
    int phase, row, idx;
    if (off < nA) { phase = 0; row = step; idx = off; }
    else if (off < nA + nB) { phase = 1; row = step - 1; idx = off - nA; }
    else { phase = 2; row = step - 2; idx = off - nA - nB; }
    if (row < 0 || row >= nrows) continue;
    float2* scr = ring + (long long)(row % kPipeRing) * M;
    if (tid == 0) {
      if (phase == 0) {                     // the ring buffer must have been drained by its previous row
        if (row >= kPipeRing) pipe_wait((HAS_C ? doneC : doneB) + row - kPipeRing, HAS_C ? nC : nB);
      } else if (phase == 1) {
        if (HAS_A) pipe_wait(doneA + row, nA);
        else if (row >= kPipeRing) pipe_wait(doneC + row - kPipeRing, nC);
      } else {
        pipe_wait(doneB + row, nB);
      }
    }
    __syncthreads();
    if (phase == 0) {
      if constexpr (HAS_A) col_fwd_body<DT, M1, T2, NT, 1, VEC, STG, KIND == HY_PIPE_BWDG>(a, idx, row, scr);
    } else if (phase == 1) {
      row_conv_body<S, NT, ROWMODE>(a, idx, row, scr);
    } else {
      if constexpr (HAS_C) col_inv_body<DT, M1, T2, NT, 1, KIND == HY_PIPE_BWDG ? 1 : 0, VEC, STG>(a, idx, row, scr);
    }
    hy_threadfence();                       // this thread's scratch / output writes are visible device-wide ...
    __syncthreads();                        // ... for every thread of the CTA, before the arrival is published
    if (tid == 0) hy_red_release((phase == 0 ? doneA : (phase == 1 ? doneB : doneC)) + row, 1u);
  }
}

template <class DT, int M1, int T2, int S, int NT, int KIND>
__global__ void __launch_bounds__(NT, 2) k_conv_pipe(ConvArgs a, float2* ring, unsigned* ctl) {
  if constexpr (DT::kBf16) {
    conv_pipe_body<DT, M1, T2, S, NT, KIND, true, true>(a, ring, ctl);     // host guarantees stage_ok
  } else {
    conv_pipe_body<DT, M1, T2, S, NT, KIND, true, false>(a, ring, ctl);    // host guarantees vec_all
  }
}
