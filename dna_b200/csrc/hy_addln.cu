// hyena-b200: residual add + LayerNorm in one pass over [rows, D] -- the Block glue either side of the
// operator (/root/reference/standalone_hyenadna.py:520-541; the src tree has the same fusion hook as
// `dropout_add_layer_norm`, src/models/sequence/long_conv_lm.py:560-575).  Dropout p = 0 (every HyenaDNA
// config):
//     r = x + res_in                      (rounded to the residual dtype, as the reference's add is)
//     y = (r - mean(r)) * rsqrt(var(r) + eps) * gamma + beta
// One warp owns one row and keeps it in registers (D = 128 * K, K in {1, 2, 4, 8}); each element is read
// once and written once: 12 B/element forward for bf16 x/y + fp32 residual, against 24 B and three kernels
// for add -> layer_norm -> cast.  Measured at [1 M, 256] on B200: 6.2 TB/s forward and backward (95 % of the
// measured copy bandwidth) once each warp had two rows in flight.
#include "hy_host.h"

namespace hy {

constexpr int kLnThreads = 256;
constexpr int kLnWarps = kLnThreads / 32;

HY_DEVICE float ln_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// four adjacent elements at element index i (i % 4 == 0) of a row-major [rows, D] tensor of dtype dt
HY_DEVICE float4 ln_ld4(const void* base, int dt, long long i) {
  if (dt == HY_BF16) {
    const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned short*>(base) + i);
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                       __uint_as_float(u.y & 0xffff0000u));
  }
  return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + i);
}
HY_DEVICE void ln_st4(void* base, int dt, long long i, float4 v) {
  if (dt == HY_BF16) {
    *reinterpret_cast<uint2*>(reinterpret_cast<unsigned short*>(base) + i) =
        make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  } else {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + i) = v;
  }
}
HY_DEVICE float4 ln_round4(float4 v) {
  const float2 a = round2_to_bf16(make_float2(v.x, v.y)), b = round2_to_bf16(make_float2(v.z, v.w));
  return make_float4(a.x, a.y, b.x, b.y);
}

struct AddLnFwd {
  const void* x;      // [rows, D] or null
  const void* res;    // [rows, D] or null
  const float *gamma, *beta;
  void* y;
  void* res_out;      // null: do not write r
  float *mean, *rstd;
  long long rows;
  int D, xdt, rdt, ydt;
  float eps;
  const unsigned char* keep;   // [rows, D] dropout keep mask of x (1 = keep) or null
  float keep_scale;            // 1 / (1 - p)
};

// R rows per warp and iteration: the loads of all R rows are issued before the first reduction (memory-level
// parallelism without more resident warps).
template <int K>
struct LnRows {
  static constexpr int kR = (K <= 2) ? 2 : 1;   // wider rows already carry enough loads per lane (and the registers)
};

template <int K>
__global__ void __launch_bounds__(kLnThreads) k_add_ln_fwd(AddLnFwd a) {
  constexpr int kLnR = LnRows<K>::kR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 g[K], b[K];
#pragma unroll
  for (int c = 0; c < K; ++c) {
    const int col = (c * 32 + lane) * 4;
    g[c] = *reinterpret_cast<const float4*>(a.gamma + col);
    b[c] = *reinterpret_cast<const float4*>(a.beta + col);
  }
  const float invD = 1.f / (float)a.D;
  const long long stride = (long long)gridDim.x * kLnWarps;
  for (long long row0 = (long long)blockIdx.x * kLnWarps + warp; row0 < a.rows; row0 += stride * kLnR) {
    float4 v[kLnR][K];
#pragma unroll
    for (int r = 0; r < kLnR; ++r) {
      long long row = row0 + r * stride;
      if (row >= a.rows) row = row0;          // duplicate work, never stored
      const long long base = row * a.D;
#pragma unroll
      for (int c = 0; c < K; ++c) {
        const long long i = base + (c * 32 + lane) * 4;
        if (a.x != nullptr) {
          v[r][c] = ln_ld4(a.x, a.xdt, i);
          if (a.keep != nullptr) {
            // dropped = dropout(x): scaled where kept, in x's dtype (standalone_hyenadna.py:521) — before the add
            const unsigned m = *reinterpret_cast<const unsigned*>(a.keep + i);
            v[r][c].x = (m & 0xffu) ? v[r][c].x * a.keep_scale : 0.f;
            v[r][c].y = (m & 0xff00u) ? v[r][c].y * a.keep_scale : 0.f;
            v[r][c].z = (m & 0xff0000u) ? v[r][c].z * a.keep_scale : 0.f;
            v[r][c].w = (m & 0xff000000u) ? v[r][c].w * a.keep_scale : 0.f;
            if (a.xdt == HY_BF16) v[r][c] = ln_round4(v[r][c]);
          }
          if (a.res != nullptr) {
            const float4 q = ln_ld4(a.res, a.rdt, i);
            v[r][c].x += q.x; v[r][c].y += q.y; v[r][c].z += q.z; v[r][c].w += q.w;
          }
        } else {
          v[r][c] = ln_ld4(a.res, a.rdt, i);
        }
        if (a.rdt == HY_BF16) v[r][c] = ln_round4(v[r][c]);
      }
    }
#pragma unroll
    for (int r = 0; r < kLnR; ++r) {
      const long long row = row0 + r * stride;
      if (row >= a.rows) continue;            // warp-uniform
      const long long base = row * a.D;
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < K; ++c) s += (v[r][c].x + v[r][c].y) + (v[r][c].z + v[r][c].w);
      const float mu = ln_warp_sum(s) * invD;
      float q = 0.f;
#pragma unroll
      for (int c = 0; c < K; ++c) {
        const float dx = v[r][c].x - mu, dy = v[r][c].y - mu, dz = v[r][c].z - mu, dw = v[r][c].w - mu;
        q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
      }
      const float rs = 1.f / sqrtf(ln_warp_sum(q) * invD + a.eps);
#pragma unroll
      for (int c = 0; c < K; ++c) {
        const long long i = base + (c * 32 + lane) * 4;
        if (a.res_out != nullptr) ln_st4(a.res_out, a.rdt, i, v[r][c]);
        float4 o;
        o.x = (v[r][c].x - mu) * rs * g[c].x + b[c].x;
        o.y = (v[r][c].y - mu) * rs * g[c].y + b[c].y;
        o.z = (v[r][c].z - mu) * rs * g[c].z + b[c].z;
        o.w = (v[r][c].w - mu) * rs * g[c].w + b[c].w;
        ln_st4(a.y, a.ydt, i, o);
      }
      if (lane == 0) {
        a.mean[row] = mu;
        a.rstd[row] = rs;
      }
    }
  }
}

struct AddLnBwd {
  const void* dy;        // [rows, D]
  const void* dres_out;  // [rows, D] or null
  const void* r;         // [rows, D] the residual stream the forward normalised
  const float *mean, *rstd, *gamma;
  void* dx;              // null: not wanted
  void* dres_in;         // null: not wanted
  float* part;           // [gridDim.x][2][D]
  long long rows;
  int D, xdt, rdt, ydt;
  const unsigned char* keep;   // dropout keep mask of x (forward's) or null
  float keep_scale;
};

template <int K>
__global__ void __launch_bounds__(kLnThreads) k_add_ln_bwd(AddLnBwd a) {
  constexpr int kLnR = LnRows<K>::kR;
  HY_DYN_SMEM(float, sm);  // [kLnWarps][2][D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 g[K], dg[K], db[K];
#pragma unroll
  for (int c = 0; c < K; ++c) {
    g[c] = *reinterpret_cast<const float4*>(a.gamma + (c * 32 + lane) * 4);
    dg[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float invD = 1.f / (float)a.D;
  const long long stride = (long long)gridDim.x * kLnWarps;
  for (long long row0 = (long long)blockIdx.x * kLnWarps + warp; row0 < a.rows; row0 += stride * kLnR) {
    // all loads of the kLnR rows first
    float4 d[kLnR][K], r[kLnR][K], e[kLnR][K];
    float mu[kLnR], rs[kLnR];
#pragma unroll
    for (int q = 0; q < kLnR; ++q) {
      long long row = row0 + q * stride;
      if (row >= a.rows) row = row0;
      const long long base = row * a.D;
      mu[q] = a.mean[row];
      rs[q] = a.rstd[row];
#pragma unroll
      for (int c = 0; c < K; ++c) {
        const long long i = base + (c * 32 + lane) * 4;
        d[q][c] = ln_ld4(a.dy, a.ydt, i);
        r[q][c] = ln_ld4(a.r, a.rdt, i);
        e[q][c] = (a.dres_out != nullptr) ? ln_ld4(a.dres_out, a.rdt, i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int q = 0; q < kLnR; ++q) {
      const long long row = row0 + q * stride;
      if (row >= a.rows) continue;      // warp-uniform
      const long long base = row * a.D;
      float4 xh[K], gy[K];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int c = 0; c < K; ++c) {
        const float4 dd = d[q][c], rr = r[q][c];
        xh[c] = make_float4((rr.x - mu[q]) * rs[q], (rr.y - mu[q]) * rs[q], (rr.z - mu[q]) * rs[q], (rr.w - mu[q]) * rs[q]);
        gy[c] = make_float4(dd.x * g[c].x, dd.y * g[c].y, dd.z * g[c].z, dd.w * g[c].w);
        s1 += (gy[c].x + gy[c].y) + (gy[c].z + gy[c].w);
        s2 += (gy[c].x * xh[c].x + gy[c].y * xh[c].y) + (gy[c].z * xh[c].z + gy[c].w * xh[c].w);
        dg[c].x += dd.x * xh[c].x; dg[c].y += dd.y * xh[c].y; dg[c].z += dd.z * xh[c].z; dg[c].w += dd.w * xh[c].w;
        db[c].x += dd.x; db[c].y += dd.y; db[c].z += dd.z; db[c].w += dd.w;
      }
      const float c1 = ln_warp_sum(s1) * invD, c2 = ln_warp_sum(s2) * invD;
#pragma unroll
      for (int c = 0; c < K; ++c) {
        const long long i = base + (c * 32 + lane) * 4;
        float4 o;
        o.x = rs[q] * (gy[c].x - c1 - xh[c].x * c2) + e[q][c].x;
        o.y = rs[q] * (gy[c].y - c1 - xh[c].y * c2) + e[q][c].y;
        o.z = rs[q] * (gy[c].z - c1 - xh[c].z * c2) + e[q][c].z;
        o.w = rs[q] * (gy[c].w - c1 - xh[c].w * c2) + e[q][c].w;
        if (a.dres_in != nullptr) ln_st4(a.dres_in, a.rdt, i, o);
        if (a.dx != nullptr) {
          if (a.keep != nullptr) {
            const unsigned m = *reinterpret_cast<const unsigned*>(a.keep + i);
            o.x = (m & 0xffu) ? o.x * a.keep_scale : 0.f;
            o.y = (m & 0xff00u) ? o.y * a.keep_scale : 0.f;
            o.z = (m & 0xff0000u) ? o.z * a.keep_scale : 0.f;
            o.w = (m & 0xff000000u) ? o.w * a.keep_scale : 0.f;
          }
          ln_st4(a.dx, a.xdt, i, o);
        }
      }
    }
  }
  // per-CTA partial sums of dgamma / dbeta, fixed order (deterministic)
  float* mine = sm + (size_t)warp * 2 * a.D;
#pragma unroll
  for (int c = 0; c < K; ++c) {
    const int col = (c * 32 + lane) * 4;
    *reinterpret_cast<float4*>(mine + col) = dg[c];
    *reinterpret_cast<float4*>(mine + a.D + col) = db[c];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < 2 * a.D; j += kLnThreads) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kLnWarps; ++w) s += sm[(size_t)w * 2 * a.D + j];
    a.part[(size_t)blockIdx.x * 2 * a.D + j] = s;
  }
}

// dgamma[j] = sum_cta part[cta][0][j], dbeta[j] = sum_cta part[cta][1][j]; one warp per 32 columns
__global__ void __launch_bounds__(kLnThreads) k_add_ln_reduce(const float* part, int nparts, int D, float* dgamma,
                                                            float* dbeta) {
  HY_STATIC_SMEM(float, red, kLnThreads);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;  // column of the [2][D] pair, j in [0, 2D)
  float s = 0.f;
  if (j < 2 * D)
    for (int p = warp; p < nparts; p += kLnWarps) s += part[(size_t)p * 2 * D + j];
  red[threadIdx.x] = s;
  __syncthreads();
  if (warp == 0 && j < 2 * D) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kLnWarps; ++w) t += red[w * 32 + lane];
    if (j < D) dgamma[j] = t;
    else dbeta[j - D] = t;
  }
}

static int ln_grid(long long rows) {
  long long want = (rows + kLnWarps - 1) / kLnWarps;
  const long long cap = 148 * 8;  // 8 resident CTAs of 256 threads per SM
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

static bool ln_dt_ok(int dt) { return dt == HY_F32 || dt == HY_BF16; }
static bool ln_al(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace hy

using namespace hy;

extern "C" int hy_add_ln_supported(int D) { return (D == 128 || D == 256 || D == 512 || D == 1024) ? 1 : 0; }

extern "C" int hy_add_ln_bwd_parts(long long rows, int D) {
  (void)D;
  return ln_grid(rows);
}

extern "C" int hy_add_ln_dropout_fwd(const void* x, int x_dtype, const unsigned char* keep, float keep_scale,
                                     const void* res_in, int res_dtype, const float* gamma, const float* beta, float eps,
                                     void* y, int y_dtype, void* res_out, float* mean, float* rstd, long long rows, int D,
                                     void* stream);

extern "C" int hy_add_ln_fwd(const void* x, int x_dtype, const void* res_in, int res_dtype, const float* gamma,
                             const float* beta, float eps, void* y, int y_dtype, void* res_out, float* mean, float* rstd,
                             long long rows, int D, void* stream) {
  return hy_add_ln_dropout_fwd(x, x_dtype, nullptr, 1.f, res_in, res_dtype, gamma, beta, eps, y, y_dtype, res_out, mean, rstd,
                               rows, D, stream);
}

extern "C" int hy_add_ln_dropout_fwd(const void* x, int x_dtype, const unsigned char* keep, float keep_scale,
                                     const void* res_in, int res_dtype, const float* gamma, const float* beta, float eps,
                                     void* y, int y_dtype, void* res_out, float* mean, float* rstd, long long rows, int D,
                                     void* stream) {
  if (keep != nullptr && (x == nullptr || (reinterpret_cast<uintptr_t>(keep) & 3)))
    return fail(HY_ERR_ARG, "hy_add_ln_dropout_fwd: the keep mask needs x and 4-byte alignment");
  if (!hy_add_ln_supported(D)) return fail(HY_ERR_UNSUPPORTED, "hy_add_ln_fwd: D=%d (need 128, 256, 512 or 1024)", D);
  if (rows < 0) return fail(HY_ERR_ARG, "hy_add_ln_fwd: rows=%lld", rows);
  if (!ln_dt_ok(x_dtype) || !ln_dt_ok(res_dtype) || !ln_dt_ok(y_dtype)) return fail(HY_ERR_ARG, "hy_add_ln_fwd: dtype");
  if (rows == 0) return HY_OK;
  if ((x == nullptr && res_in == nullptr) || !gamma || !beta || !y || !mean || !rstd)
    return fail(HY_ERR_ARG, "hy_add_ln_fwd: null pointer");
  if (!ln_al(x) || !ln_al(res_in) || !ln_al(gamma) || !ln_al(beta) || !ln_al(y) || !ln_al(res_out))
    return fail(HY_ERR_ARG, "hy_add_ln_fwd: pointers must be 16-byte aligned");
  AddLnFwd a{x, res_in, gamma, beta, y, res_out, mean, rstd, rows, D, x_dtype, res_dtype, y_dtype, eps, keep, keep_scale};
  const int grid = ln_grid(rows);
  switch (D / 128) {
    case 1: HY_LAUNCH(k_add_ln_fwd<1>, grid, kLnThreads, 0, stream, a); break;
    case 2: HY_LAUNCH(k_add_ln_fwd<2>, grid, kLnThreads, 0, stream, a); break;
    case 4: HY_LAUNCH(k_add_ln_fwd<4>, grid, kLnThreads, 0, stream, a); break;
    default: HY_LAUNCH(k_add_ln_fwd<8>, grid, kLnThreads, 0, stream, a); break;
  }
  return check_launch("k_add_ln_fwd");
}

extern "C" int hy_add_ln_dropout_bwd(const void* dy, int y_dtype, const void* dres_out, int res_dtype, const void* r,
                                     const float* mean, const float* rstd, const float* gamma, void* dx, int x_dtype,
                                     const unsigned char* keep, float keep_scale, void* dres_in, float* part, float* dgamma,
                                     float* dbeta, long long rows, int D, void* stream);

extern "C" int hy_add_ln_bwd(const void* dy, int y_dtype, const void* dres_out, int res_dtype, const void* r,
                             const float* mean, const float* rstd, const float* gamma, void* dx, int x_dtype,
                             void* dres_in, float* part, float* dgamma, float* dbeta, long long rows, int D,
                             void* stream) {
  return hy_add_ln_dropout_bwd(dy, y_dtype, dres_out, res_dtype, r, mean, rstd, gamma, dx, x_dtype, nullptr, 1.f, dres_in, part,
                               dgamma, dbeta, rows, D, stream);
}

extern "C" int hy_add_ln_dropout_bwd(const void* dy, int y_dtype, const void* dres_out, int res_dtype, const void* r,
                                     const float* mean, const float* rstd, const float* gamma, void* dx, int x_dtype,
                                     const unsigned char* keep, float keep_scale, void* dres_in, float* part, float* dgamma,
                                     float* dbeta, long long rows, int D, void* stream) {
  if (!hy_add_ln_supported(D)) return fail(HY_ERR_UNSUPPORTED, "hy_add_ln_bwd: D=%d (need 128, 256, 512 or 1024)", D);
  if (rows <= 0) return fail(HY_ERR_ARG, "hy_add_ln_bwd: rows=%lld", rows);
  if (!ln_dt_ok(x_dtype) || !ln_dt_ok(res_dtype) || !ln_dt_ok(y_dtype)) return fail(HY_ERR_ARG, "hy_add_ln_bwd: dtype");
  if (!dy || !r || !mean || !rstd || !gamma || !part || !dgamma || !dbeta || (dx == nullptr && dres_in == nullptr))
    return fail(HY_ERR_ARG, "hy_add_ln_bwd: null pointer");
  if (!ln_al(dy) || !ln_al(dres_out) || !ln_al(r) || !ln_al(gamma) || !ln_al(dx) || !ln_al(dres_in))
    return fail(HY_ERR_ARG, "hy_add_ln_bwd: pointers must be 16-byte aligned");
  AddLnBwd a{dy, dres_out, r, mean, rstd, gamma, dx, dres_in, part, rows, D, x_dtype, res_dtype, y_dtype, keep, keep_scale};
  const int grid = ln_grid(rows);
  const size_t smem = (size_t)kLnWarps * 2 * D * sizeof(float);
  switch (D / 128) {
    case 1: HY_LAUNCH(k_add_ln_bwd<1>, grid, kLnThreads, smem, stream, a); break;
    case 2: HY_LAUNCH(k_add_ln_bwd<2>, grid, kLnThreads, smem, stream, a); break;
    case 4: HY_LAUNCH(k_add_ln_bwd<4>, grid, kLnThreads, smem, stream, a); break;
    default: HY_LAUNCH(k_add_ln_bwd<8>, grid, kLnThreads, smem, stream, a); break;
  }
  int rc = check_launch("k_add_ln_bwd");
  if (rc) return rc;
  HY_LAUNCH(k_add_ln_reduce, (2 * D + 31) / 32, kLnThreads, 0, stream, (const float*)part, grid, D, dgamma, dbeta);
  return check_launch("k_add_ln_reduce");
}
