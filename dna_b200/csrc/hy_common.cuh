// hyena-b200: common device/host helpers shared by all kernels.
//
// The same kernel source is compiled two ways:
//   * nvcc -gencode arch=compute_100a,code=sm_100a  -> the shipped library (libhyena_b200.so)
//   * g++ -DHY_EMU_BUILD with tests/emu/cuda_emu.h   -> a CPU execution-model emulation used ONLY by
//     the "not gpu" tests to exercise kernel index algebra where no GPU exists.  Not a fallback.
#pragma once

#ifdef HY_EMU_BUILD
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#endif
#include <stdint.h>

#define HY_DEVICE __device__ __forceinline__
#define HY_HD __host__ __device__

// ---- shared memory declaration that works in both builds -------------------------------------
#ifdef HY_EMU_BUILD
#define HY_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::dyn_smem())
#define HY_STATIC_SMEM(type, name, count) \
  type* name = reinterpret_cast<type*>(emu::static_smem(__LINE__, sizeof(type) * (count)))
#else
#define HY_DYN_SMEM(type, name)                                   \
  extern __shared__ __align__(128) unsigned char hy_dyn_smem_[]; \
  type* name = reinterpret_cast<type*>(hy_dyn_smem_)
#define HY_STATIC_SMEM(type, name, count) __shared__ __align__(16) type name[count]
#endif

// ---- complex helpers --------------------------------------------------------------------------
// sm_100 has packed fp32 arithmetic on aligned register pairs (FADD2 / FMUL2 / FFMA2, PTX add/mul/fma.f32x2): one issue
// slot per COMPLEX add instead of two.  The long-conv kernels are issue-bound, not FP32-pipe-bound, so the butterflies'
// adds and subtracts go through it; every lane is the same IEEE operation as the scalar form.
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000) && !defined(HY_NO_F32X2)
#define HY_F32X2 1
#else
#define HY_F32X2 0
#endif
HY_DEVICE float2 cadd(float2 a, float2 b) {
#if HY_F32X2
  return __fadd2_rn(a, b);
#else
  return make_float2(a.x + b.x, a.y + b.y);
#endif
}
HY_DEVICE float2 csub(float2 a, float2 b) {
#if HY_F32X2
  return __fadd2_rn(a, make_float2(-b.x, -b.y));   // the negation folds into the operand modifier
#else
  return make_float2(a.x - b.x, a.y - b.y);
#endif
}
HY_DEVICE float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
HY_DEVICE float2 cmulc(float2 a, float2 b) {
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
HY_DEVICE float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
HY_DEVICE float2 cscale(float2 a, float s) {
#if HY_F32X2
  return __fmul2_rn(a, make_float2(s, s));
#else
  return make_float2(a.x * s, a.y * s);
#endif
}
// multiply by -i (forward quarter turn) or +i (inverse quarter turn)
template <bool INV>
HY_DEVICE float2 crot(float2 a) {
  return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}
// a * (c - i s) forward, a * (c + i s) inverse, with (c, s) = (cos, sin) constants
template <bool INV>
HY_DEVICE float2 ctw(float2 a, float c, float s) {
  return INV ? make_float2(a.x * c - a.y * s, a.y * c + a.x * s)
             : make_float2(a.x * c + a.y * s, a.y * c - a.x * s);
}
// a * w (forward) or a * conj(w) (inverse) for a table twiddle w = exp(-i theta)
template <bool INV>
HY_DEVICE float2 cmul_dir(float2 a, float2 w) {
  return INV ? cmulc(a, w) : cmul(a, w);
}

// acc[0..3] += w * (h.x, h.y, h.z, h.w): the inner step of the register-tiled micro-GEMMs, two packed FMAs (FFMA2 with
// a broadcast scalar operand) instead of four scalar ones — each lane is the same fmaf.
HY_DEVICE void fma4(float (&acc)[4], float w, const float4& h) {
#if HY_F32X2
  const float2 ww = make_float2(w, w);
  const float2 a01 = __ffma2_rn(ww, make_float2(h.x, h.y), make_float2(acc[0], acc[1]));
  const float2 a23 = __ffma2_rn(ww, make_float2(h.z, h.w), make_float2(acc[2], acc[3]));
  acc[0] = a01.x; acc[1] = a01.y; acc[2] = a23.x; acc[3] = a23.y;
#else
  acc[0] = fmaf(w, h.x, acc[0]); acc[1] = fmaf(w, h.y, acc[1]);
  acc[2] = fmaf(w, h.z, acc[2]); acc[3] = fmaf(w, h.w, acc[3]);
#endif
}

// scalar form of fma4 (the trunk-backward kernel is register-bound: packing its accumulators costs 25 more registers
// and 27 % of its speed, measured)
HY_DEVICE void fma4s(float (&acc)[4], float w, const float4& h) {
  acc[0] = fmaf(w, h.x, acc[0]); acc[1] = fmaf(w, h.y, acc[1]);
  acc[2] = fmaf(w, h.z, acc[2]); acc[3] = fmaf(w, h.w, acc[3]);
}

// ---- sin / cos of four values at once ---------------------------------------------------------------------------
// The filter MLP evaluates sin(freq * a) for every (feature, position): a third of the forward kernel's instructions
// when done with sinf(), whose large-argument branch puts every call in its own reconvergence region, so the calls of
// a thread's register tile cannot interleave.  Below the same range as sinf's fast path (|x| < 1e5) this is the same
// scheme — three-term Cody-Waite reduction by pi/2 with FMAs, degree-7 / degree-8 polynomials on [-pi/4, pi/4] — but
// branch-free over the four values (max abs error 7e-8 against fp64 up to |x| = 1e5, the same as a float32 libm,
// tools/proto_sincos.py); larger arguments take sinf / cosf for all four.
HY_DEVICE void hy_sincos_core(float x, float& s, float& c) {
  const float j = rintf(x * 0.636619772367581343f);
  float r = fmaf(j, -1.5707963705062866f, x);
  r = fmaf(j, 4.371138828673793e-08f, r);
  r = fmaf(j, 1.7763568394002505e-15f, r);
  const int q = (int)j;
  const float r2 = r * r;
  float ps = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
  ps = fmaf(ps, r2, -1.6666654611e-1f);
  ps = fmaf(ps * r2, r, r);
  float pc = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
  pc = fmaf(pc, r2, 4.166664568298827e-2f);
  pc = fmaf(pc, r2, -0.5f);
  pc = fmaf(pc, r2, 1.0f);
  const bool sw = (q & 1) != 0;
  const float ss = sw ? pc : ps, cc = sw ? ps : pc;
  s = (q & 2) ? -ss : ss;
  c = ((q + 1) & 2) ? -cc : cc;
}
HY_DEVICE bool hy_trig_fast_range(const float (&x)[4]) {
  return fmaxf(fmaxf(fabsf(x[0]), fabsf(x[1])), fmaxf(fabsf(x[2]), fabsf(x[3]))) < 1.0e5f;
}
HY_DEVICE void hy_sincos4(const float (&x)[4], float (&s)[4], float (&c)[4]) {
  if (hy_trig_fast_range(x)) {
#pragma unroll
    for (int i = 0; i < 4; ++i) hy_sincos_core(x[i], s[i], c[i]);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) { s[i] = sinf(x[i]); c[i] = cosf(x[i]); }
  }
}
HY_DEVICE void hy_sin4(const float (&x)[4], float (&s)[4]) {
  if (hy_trig_fast_range(x)) {
    float c;
#pragma unroll
    for (int i = 0; i < 4; ++i) hy_sincos_core(x[i], s[i], c);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i] = sinf(x[i]);
  }
}
HY_DEVICE void hy_cos4(const float (&x)[4], float (&c)[4]) {
  if (hy_trig_fast_range(x)) {
    float s;
#pragma unroll
    for (int i = 0; i < 4; ++i) hy_sincos_core(x[i], s, c[i]);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] = cosf(x[i]);
  }
}

// ---- bf16 helpers (bit-level so that host emulation and device agree exactly) -----------------
HY_DEVICE float bf16_bits_to_float(unsigned short h) { return __uint_as_float(((unsigned)h) << 16); }
HY_DEVICE unsigned short float_to_bf16_bits(float f) {
#if defined(__CUDA_ARCH__)
  return __bfloat16_as_ushort(__float2bfloat16_rn(f));
#else
  unsigned u = __float_as_uint(f);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (unsigned short)((u >> 16) | 0x0040u);  // NaN
  unsigned lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return (unsigned short)(u >> 16);
#endif
}
HY_DEVICE float round_to_bf16(float f) { return bf16_bits_to_float(float_to_bf16_bits(f)); }
// (lo, hi) -> packed bf16x2 with round-to-nearest-even: ONE cvt.rn.bf16x2.f32 on the device
HY_DEVICE unsigned pack_bf16x2(float lo, float hi) {
#if defined(__CUDA_ARCH__)
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<unsigned*>(&h);
#else
  return (unsigned)float_to_bf16_bits(lo) | ((unsigned)float_to_bf16_bits(hi) << 16);
#endif
}
HY_DEVICE float2 round2_to_bf16(float2 v) {
  const unsigned u = pack_bf16x2(v.x, v.y);
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

// ---- asynchronous global -> shared copies (LDGSTS): no register staging, completion waited on later --------
HY_DEVICE void hy_cp_async16(void* smem_dst, const void* gsrc) {
#if defined(__CUDA_ARCH__)
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
#else
  const uint4* s = reinterpret_cast<const uint4*>(gsrc);
  *reinterpret_cast<uint4*>(smem_dst) = *s;
#endif
}
HY_DEVICE void hy_cp_async_wait_all() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}

// ask L2 for the 128-byte line holding p (no register, no stall): used to start the trip of operands a kernel
// will only consume after a compute phase
HY_DEVICE void hy_prefetch_l2(const void* p) {
#if defined(__CUDA_ARCH__)
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

// ---- cross-CTA hand-off inside ONE launch (persistent pipeline, hy_conv_pipe.cuh) ------------------------------------
// Data another CTA of the same launch wrote: read at L2 (ld.global.cg), never through this SM's L1, which no launch
// boundary has invalidated.
HY_DEVICE float2 hy_ldcg(const float2* p) {
#if defined(__CUDA_ARCH__)
  return __ldcg(p);
#else
  return *p;
#endif
}
HY_DEVICE unsigned hy_ld_acquire(const unsigned* p) {
#if defined(__CUDA_ARCH__)
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
#else
  return __atomic_load_n(p, __ATOMIC_ACQUIRE);
#endif
}
HY_DEVICE void hy_red_release(unsigned* p, unsigned v) {
#if defined(__CUDA_ARCH__)
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#else
  __atomic_fetch_add(p, v, __ATOMIC_RELEASE);
#endif
}
HY_DEVICE void hy_threadfence() {
#if defined(__CUDA_ARCH__)
  __threadfence();
#else
  __atomic_thread_fence(__ATOMIC_SEQ_CST);
#endif
}

// ---- dtype tags ------------------------------------------------------------------------------
// Activations cross the C-ABI as raw pointers plus a dtype enum (include/hyena_b200.h).
struct DT_F32 {
  typedef float elem;
  static constexpr bool kBf16 = false;
};
struct DT_BF16 {
  typedef unsigned short elem;
  static constexpr bool kBf16 = true;
};

template <class DT>
HY_DEVICE float ld1(const typename DT::elem* p) {
  if (DT::kBf16) return bf16_bits_to_float(*reinterpret_cast<const unsigned short*>(p));
  return *reinterpret_cast<const float*>(p);
}
template <class DT>
HY_DEVICE void st1(typename DT::elem* p, float v) {
  if (DT::kBf16) *reinterpret_cast<unsigned short*>(p) = float_to_bf16_bits(v);
  else *reinterpret_cast<float*>(p) = v;
}
// two adjacent elements starting at p; `vec` says p is aligned for one 2-element access
template <class DT>
HY_DEVICE float2 ld2(const typename DT::elem* p, bool vec) {
  if (DT::kBf16) {
    if (vec) {
      unsigned u = *reinterpret_cast<const unsigned*>(p);
      return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
    }
    const unsigned short* q = reinterpret_cast<const unsigned short*>(p);
    return make_float2(bf16_bits_to_float(q[0]), bf16_bits_to_float(q[1]));
  } else {
    const float* q = reinterpret_cast<const float*>(p);
    if (vec) return *reinterpret_cast<const float2*>(q);
    return make_float2(q[0], q[1]);
  }
}
template <class DT>
HY_DEVICE void st2(typename DT::elem* p, float2 v, bool vec) {
  if (DT::kBf16) {
    const unsigned u = pack_bf16x2(v.x, v.y);
    if (vec) *reinterpret_cast<unsigned*>(p) = u;
    else {
      unsigned short* q = reinterpret_cast<unsigned short*>(p);
      q[0] = (unsigned short)(u & 0xffffu); q[1] = (unsigned short)(u >> 16);
    }
  } else {
    float* q = reinterpret_cast<float*>(p);
    if (vec) *reinterpret_cast<float2*>(q) = v;
    else { q[0] = v.x; q[1] = v.y; }
  }
}

// constexpr helpers
// loop form on purpose: a recursive constexpr was being CALLED at run time (ptxas kept it as a real function)
HY_HD constexpr int hy_ilog2(int x) {
  int r = 0;
  while (x > 1) {
    x >>= 1;
    ++r;
  }
  return r;
}
HY_HD constexpr int hy_max(int a, int b) { return a > b ? a : b; }
HY_HD constexpr int hy_min(int a, int b) { return a < b ? a : b; }
