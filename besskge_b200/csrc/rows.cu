// Per-row kernels: routed gather, score_triple fwd/bwd, query prologue fwd/bwd,
// PairRE candidate normalisation, BoxE relation-gradient finalize.
// One warp per row; HBM-bound streaming work.
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"
#include "families.cuh"

namespace bess {

template <>
struct Ld<__half> {
  static BESS_HD float f(const __half* p, int i) { return __half2float(p[i]); }
};
template <>
struct Ld<__nv_bfloat16> {
  static BESS_HD float f(const __nv_bfloat16* p, int i) { return __bfloat162float(p[i]); }
};

struct WarpCtx {
  static BESS_D int lane() { return threadIdx.x & 31; }
  static BESS_D int lanes() { return 32; }
  static BESS_D float sum(float v) { return warp_sum(v); }
  static BESS_D int all(int v) { return __all_sync(0xffffffffu, v); }
};

static inline FamCfg to_cfg(const bess_score_cfg_t* c) {
  FamCfg f;
  f.family = c->family; f.norm_p = c->norm_p; f.d = c->d; f.normalize = c->normalize;
  f.apply_tanh = c->apply_tanh; f.per_dim = c->per_dim; f.eps = c->eps; f.rel_u = c->rel_u;
  return f;
}

// ---------------------------------------------------------------------------
// Routed gather.  Rows are copied with 128-bit accesses; a warp moves
// kRowsPerWarp rows per iteration so every lane has several independent loads
// in flight (rows are 256 B - 4 KiB: one warp instruction covers 512 B).
// ---------------------------------------------------------------------------
struct RoutePtrs {
  void* dst[BESS_MAX_SHARD];
};

template <int kRowsPerWarp>
__global__ void __launch_bounds__(256) gather_route_kernel(
    const uint8_t* __restrict__ table, int64_t pitch_bytes, int row_bytes,
    const int32_t* __restrict__ idx, int n_local, int per_dst, int n_total, uint8_t* local_out,
    RoutePtrs route, int slot) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int warp_global = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int n_warps = gridDim.x * warps_per_block;
  const int vec_per_row = row_bytes >> 4;

  for (int base = warp_global * kRowsPerWarp; base < n_total; base += n_warps * kRowsPerWarp) {
    const uint4* src[kRowsPerWarp];
    uint4* dst[kRowsPerWarp];
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
      const int x = base + r;
      src[r] = nullptr; dst[r] = nullptr;
      if (x < n_total) {
        const int row = __ldg(idx + x);
        src[r] = reinterpret_cast<const uint4*>(table + (int64_t)row * pitch_bytes);
        if (x < n_local) {
          dst[r] = reinterpret_cast<uint4*>(local_out + (int64_t)x * row_bytes);
        } else {
          const int y = x - n_local;
          const int dsti = y / per_dst;
          const int i = y - dsti * per_dst;
          dst[r] = reinterpret_cast<uint4*>(static_cast<uint8_t*>(route.dst[dsti]) +
                                            ((int64_t)slot * per_dst + i) * row_bytes);
        }
      }
    }
    for (int v = lane; v < vec_per_row; v += 32) {
      uint4 val[kRowsPerWarp];
#pragma unroll
      for (int r = 0; r < kRowsPerWarp; ++r)
        if (src[r] != nullptr) val[r] = ld_stream(src[r] + v);
#pragma unroll
      for (int r = 0; r < kRowsPerWarp; ++r)
        if (dst[r] != nullptr) st_stream(dst[r] + v, val[r]);
    }
  }
}

// ---------------------------------------------------------------------------
// Python-surface helpers of the reference's utils.py.
// take_along_rows: out[i, j] = x[i or 0, index[i or 0, j]] (utils.py:10-33, a 2-D
// take-along-dim written there as a flat index_select).  Elements are opaque 1/2/4/8-byte words.
// complex_mul: rows hold [re | im] halves; out = v1 * v2 (utils.py:72-89) or, with
// rotate, v1 * (cos r + i sin r) for a row of e angles (utils.py:92-112, full-precision
// sin / cos: the fp16 shortcut there applies to IPU devices only).
// ---------------------------------------------------------------------------
template <typename U>
__global__ void __launch_bounds__(256) take_along_rows_kernel(const U* __restrict__ x, int a,
                                                               int64_t e,
                                                               const int32_t* __restrict__ index,
                                                               int b, int k, int n_out,
                                                               U* __restrict__ out) {
  const int64_t total = (int64_t)n_out * k;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(t / k), j = (int)(t - (int64_t)i * k);
    const int col = __ldg(index + (int64_t)(b == 1 ? 0 : i) * k + j);
    out[t] = x[(int64_t)(a == 1 ? 0 : i) * e + col];
  }
}

template <typename T>
__global__ void __launch_bounds__(256) complex_mul_kernel(const T* __restrict__ v1,
                                                           const T* __restrict__ v2, int n, int e,
                                                           int rotate, T* __restrict__ out) {
  const int64_t total = (int64_t)n * e;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(t / e), j = (int)(t - (int64_t)i * e);
    const float a_re = ldf(v1 + (int64_t)i * 2 * e + j), a_im = ldf(v1 + (int64_t)i * 2 * e + e + j);
    float b_re, b_im;
    if (rotate) {
      sincosf(ldf(v2 + (int64_t)i * e + j), &b_im, &b_re);
      // the reference rounds cos / sin to the tensor dtype before the product
      b_re = Elem<T>::to_f(Elem<T>::from_f(b_re));
      b_im = Elem<T>::to_f(Elem<T>::from_f(b_im));
    } else {
      b_re = ldf(v2 + (int64_t)i * 2 * e + j);
      b_im = ldf(v2 + (int64_t)i * 2 * e + e + j);
    }
    out[(int64_t)i * 2 * e + j] = Elem<T>::from_f(a_re * b_re - a_im * b_im);
    out[(int64_t)i * 2 * e + e + j] = Elem<T>::from_f(a_re * b_im + a_im * b_re);
  }
}

// ---------------------------------------------------------------------------
// DistMult on fp32 tables (BASELINE configs[1]): 128-bit versions of the three row functions.
// The generic family code reads one element per lane per instruction; these rows are 1 KiB and
// the kernels pure streams, so the instruction count decides the achieved HBM fraction.
// Same arithmetic, same per-lane summation order per 4-element group as the scalar code is
// NOT guaranteed (the dot product is re-associated), within fp32 rounding.
// ---------------------------------------------------------------------------
BESS_D bool aligned16(const void* a) { return (reinterpret_cast<uintptr_t>(a) & 15) == 0; }
BESS_D float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
BESS_D void st4(float* p, const float4& v, int add) {
  float4 o = v;
  if (add) { const float4 c = ld4(p); o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w; }
  *reinterpret_cast<float4*>(p) = o;
}
BESS_D float4 mul4(const float4& a, const float4& b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
BESS_D float4 scl4(float s, const float4& a) { return make_float4(s * a.x, s * a.y, s * a.z, s * a.w); }

// ---------------------------------------------------------------------------
// score_triple
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) triple_fwd_kernel(FamCfg cfg, bess_rows_t head,
                                                          bess_rows_t tail, const T* rel_table,
                                                          int rel_pitch, const int32_t* rel_id,
                                                          bess_rowmap_t rel_map, int n, float* score,
                                                          bess_rowmap_t score_map) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n) return;
  const T* h = static_cast<const T*>(head.base) + src_row(head, w) * head.pitch;
  const T* t = static_cast<const T*>(tail.base) + src_row(tail, w) * tail.pitch;
  const T* r = rel_table + (int64_t)__ldg(rel_id + map_row(rel_map, w)) * rel_pitch;
  if constexpr (sizeof(T) == 4) {
    if (cfg.family == FAM_DISTMULT && (cfg.d & 3) == 0 && aligned16(h) && aligned16(t) && aligned16(r)) {
      const int lane = threadIdx.x & 31;
      float acc = 0.f;
      for (int k = lane * 4; k < cfg.d; k += 128) {
        const float4 a = ld4((const float*)h + k), b = ld4((const float*)r + k), c = ld4((const float*)t + k);
        acc += a.x * b.x * c.x + a.y * b.y * c.y + a.z * b.z * c.z + a.w * b.w * c.w;
      }
      acc = warp_sum(acc);
      if (lane == 0) score[map_row(score_map, w)] = acc;
      return;
    }
  }
  const float s = triple_fwd<WarpCtx, T>(cfg, h, r, t);
  if ((threadIdx.x & 31) == 0) score[map_row(score_map, w)] = s;
}

template <typename T>
__global__ void __launch_bounds__(256) triple_bwd_kernel(
    FamCfg cfg, bess_rows_t head, bess_rows_t tail, const T* rel_table, int rel_pitch,
    const int32_t* rel_id, bess_rowmap_t rel_map, int n, const float* score, const float* d_score,
    bess_rowmap_t score_map, bess_rows_t d_head, bess_rows_t d_tail, float* d_rel, int rel_width,
    int add_h, int add_t, int add_r) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n) return;
  const T* h = static_cast<const T*>(head.base) + src_row(head, w) * head.pitch;
  const T* t = static_cast<const T*>(tail.base) + src_row(tail, w) * tail.pitch;
  const T* r = rel_table + (int64_t)__ldg(rel_id + map_row(rel_map, w)) * rel_pitch;
  float* dh = static_cast<float*>(const_cast<void*>(d_head.base)) + src_row(d_head, w) * d_head.pitch;
  float* dt = static_cast<float*>(const_cast<void*>(d_tail.base)) + src_row(d_tail, w) * d_tail.pitch;
  float* dr = d_rel + (int64_t)map_row(rel_map, w) * rel_width;
  const int sr = map_row(score_map, w);
  if constexpr (sizeof(T) == 4) {
    if (cfg.family == FAM_DISTMULT && (cfg.d & 3) == 0 && aligned16(h) && aligned16(t) && aligned16(r) &&
        aligned16(dh) && aligned16(dt) && aligned16(dr)) {
      const float g = d_score[sr];
      for (int k = (threadIdx.x & 31) * 4; k < cfg.d; k += 128) {
        const float4 a = ld4((const float*)h + k), b = ld4((const float*)r + k), c = ld4((const float*)t + k);
        st4(dh + k, scl4(g, mul4(b, c)), add_h);
        st4(dt + k, scl4(g, mul4(a, b)), add_t);
        st4(dr + k, scl4(g, mul4(a, c)), add_r);
      }
      return;
    }
  }
  triple_bwd<WarpCtx, T>(cfg, h, r, t, score[sr], d_score[sr], dh, dr, dt, add_h, add_r, add_t);
}

// ---------------------------------------------------------------------------
// query prologue
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) prologue_fwd_kernel(FamCfg cfg, int mode, bess_rows_t fixed,
                                                            const T* rel_table, int rel_pitch,
                                                            const int32_t* rel_id,
                                                            bess_rowmap_t rel_map, int n, float* qv,
                                                            int qv_row) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n) return;
  const T* x = static_cast<const T*>(fixed.base) + src_row(fixed, w) * fixed.pitch;
  const T* r = rel_table + (int64_t)__ldg(rel_id + map_row(rel_map, w)) * rel_pitch;
  if constexpr (sizeof(T) == 4) {
    float* q = qv + (int64_t)w * qv_row;
    if (cfg.family == FAM_DISTMULT && (cfg.d & 3) == 0 && aligned16(x) && aligned16(r) && aligned16(q)) {
      for (int k = (threadIdx.x & 31) * 4; k < cfg.d; k += 128)
        st4(q + k, mul4(ld4((const float*)x + k), ld4((const float*)r + k)), 0);
      return;
    }
  }
  prologue_fwd<WarpCtx, T>(cfg, mode, x, r, qv + (int64_t)w * qv_row);
}

template <typename T>
__global__ void __launch_bounds__(256) prologue_bwd_kernel(
    FamCfg cfg, int mode, bess_rows_t fixed, const T* rel_table, int rel_pitch,
    const int32_t* rel_id, bess_rowmap_t rel_map, int n, const float* d_qv, int qv_row,
    bess_rows_t d_fixed, float* d_rel, int rel_width, int add_x, int add_r) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n) return;
  const T* x = static_cast<const T*>(fixed.base) + src_row(fixed, w) * fixed.pitch;
  const int rrow = map_row(rel_map, w);
  const T* r = rel_table + (int64_t)__ldg(rel_id + rrow) * rel_pitch;
  float* dx = static_cast<float*>(const_cast<void*>(d_fixed.base)) + src_row(d_fixed, w) * d_fixed.pitch;
  if constexpr (sizeof(T) == 4) {
    const float* dq = d_qv + (int64_t)w * qv_row;
    float* dr = d_rel + (int64_t)rrow * rel_width;
    if (cfg.family == FAM_DISTMULT && (cfg.d & 3) == 0 && aligned16(x) && aligned16(r) && aligned16(dq) &&
        aligned16(dx) && aligned16(dr)) {
      for (int k = (threadIdx.x & 31) * 4; k < cfg.d; k += 128) {
        const float4 g = ld4(dq + k);
        st4(dx + k, mul4(g, ld4((const float*)r + k)), add_x);
        st4(dr + k, mul4(g, ld4((const float*)x + k)), add_r);
      }
      return;
    }
  }
  prologue_bwd<WarpCtx, T>(cfg, mode, x, r, d_qv + (int64_t)w * qv_row, dx,
                           d_rel + (int64_t)rrow * rel_width, add_x, add_r);
}

template <typename T>
__global__ void __launch_bounds__(256) boxe_finalize_kernel(FamCfg cfg, const T* rel_table,
                                                             int rel_pitch, const int32_t* rel_id,
                                                             int n, float* d_rel, int rel_width) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n) return;
  const T* r = rel_table + (int64_t)__ldg(rel_id + w) * rel_pitch;
  boxe_rel_finalize<WarpCtx, T>(cfg, r, d_rel + (int64_t)w * rel_width);
}

// ---------------------------------------------------------------------------
// PairRE candidate normalisation
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) cand_inv_norm_kernel(bess_rows_t cand, int n, int width,
                                                             float* inv_norm) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n) return;
  const T* c = static_cast<const T*>(cand.base) + src_row(cand, w) * cand.pitch;
  float a = 0.f;
  for (int k = threadIdx.x & 31; k < width; k += 32) { const float v = ldf(c + k); a += v * v; }
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) inv_norm[w] = 1.f / fmaxf(sqrtf(a), 1e-12f);
}

template <typename T>
__global__ void __launch_bounds__(256) cand_norm_bwd_kernel(bess_rows_t cand, int n, int width,
                                                             const float* inv_norm,
                                                             bess_rows_t d_cand) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n) return;
  const T* c = static_cast<const T*>(cand.base) + src_row(cand, w) * cand.pitch;
  float* g = static_cast<float*>(const_cast<void*>(d_cand.base)) + src_row(d_cand, w) * d_cand.pitch;
  const float inv = inv_norm[w];
  float proj = 0.f;
  for (int k = threadIdx.x & 31; k < width; k += 32) proj += ldf(c + k) * inv * g[k];
  proj = warp_sum(proj);
  const bool clamped = inv >= 1e12f;  // ||c|| <= eps: normalise is c / eps, Jacobian I / eps
  for (int k = threadIdx.x & 31; k < width; k += 32) {
    const float ch = ldf(c + k) * inv;
    g[k] = clamped ? g[k] * inv : (g[k] - ch * proj) * inv;
  }
}

}  // namespace bess

using namespace bess;

// ----------------------------------------------------------------------------
// error string
// ----------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void bess_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
unsigned long long g_bess_launches = 0;
extern "C" const char* bess_last_error(void) { return g_err; }
extern "C" int64_t bess_launch_count(void) { return (int64_t)g_bess_launches; }
extern "C" int bess_version(void) { return 100; }
extern "C" int bess_entity_width(const bess_score_cfg_t* cfg) { return ent_width(to_cfg(cfg)); }
extern "C" int bess_relation_width(const bess_score_cfg_t* cfg) { return rel_width(to_cfg(cfg)); }
extern "C" int bess_query_nvec(const bess_score_cfg_t* cfg) { return pair_nvec(to_cfg(cfg)); }

static inline int elem_size(int dtype) { return dtype == BESS_F32 ? 4 : 2; }

static int check_cfg(const bess_score_cfg_t* cfg) {
  BESS_CHECK_ARG(cfg != nullptr, "null score config");
  BESS_CHECK_ARG(cfg->family >= 0 && cfg->family <= BESS_TRANS, "unknown family %d", cfg->family);
  BESS_CHECK_ARG(cfg->d > 0, "embedding_size must be positive");
  if (cfg->family == BESS_TRANSE || cfg->family == BESS_ROTATE || cfg->family == BESS_PAIRRE ||
      cfg->family == BESS_BOXE || cfg->family == BESS_TRIPLERE || cfg->family == BESS_INTERHT ||
      cfg->family == BESS_TRANS)
    BESS_CHECK_ARG(cfg->norm_p == 1 || cfg->norm_p == 2, "scoring_norm %d not supported (1 or 2)",
                   cfg->norm_p);
  return BESS_OK;
}

extern "C" int bess_gather_route(const void* table, int64_t table_pitch, int dtype, int row_elems,
                                 const int32_t* idx, int n_local, int n_dst, int per_dst,
                                 void* local_out, void* const* dst_out, int slot, void* stream) {
  const int es = elem_size(dtype);
  const int row_bytes = row_elems * es;
  BESS_CHECK_ARG(row_bytes % 16 == 0, "row of %d bytes is not a multiple of 16", row_bytes);
  BESS_CHECK_ARG((table_pitch * es) % 16 == 0, "table pitch not 16-byte aligned");
  BESS_CHECK_ARG(n_dst >= 0 && n_dst <= BESS_MAX_SHARD, "n_dst %d out of range", n_dst);
  BESS_CHECK_ARG(((uintptr_t)table & 15) == 0 && ((uintptr_t)local_out & 15) == 0,
                 "buffers must be 16-byte aligned");
  const int n_total = n_local + n_dst * per_dst;
  if (n_total == 0) return BESS_OK;
  RoutePtrs route;
  for (int i = 0; i < BESS_MAX_SHARD; ++i) route.dst[i] = i < n_dst ? dst_out[i] : nullptr;
  constexpr int kRows = 4;
  const int warps = ceil_div(n_total, kRows);
  int blocks = ceil_div(warps, 8);
  const int max_blocks = kNumSM * 8;
  if (blocks > max_blocks) blocks = max_blocks;
  gather_route_kernel<kRows><<<blocks, 256, 0, (cudaStream_t)stream>>>(
      static_cast<const uint8_t*>(table), table_pitch * es, row_bytes, idx, n_local,
      per_dst > 0 ? per_dst : 1, n_total, static_cast<uint8_t*>(local_out), route, slot);
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

extern "C" int bess_gather_rows(const void* table, int64_t table_pitch, int dtype, int row_elems,
                                const int32_t* idx, int n_idx, void* out, void* stream) {
  return bess_gather_route(table, table_pitch, dtype, row_elems, idx, n_idx, 0, 0, out, nullptr, 0,
                           stream);
}

static inline int stream_blocks(int64_t total) {
  int64_t b = (total + 255) / 256;
  if (b > (int64_t)kNumSM * 8) b = (int64_t)kNumSM * 8;
  return b < 1 ? 1 : (int)b;
}

extern "C" int bess_take_along_rows(const void* x, int a, int64_t e, int elem_bytes,
                                    const int32_t* index, int b, int k, void* out, void* stream) {
  BESS_CHECK_ARG(a >= 1 && b >= 1 && (a == 1 || b == 1 || a == b),
                 "bess_take_along_rows: x has %d rows, index %d (need equal, or one of them 1)", a, b);
  const int n_out = a > b ? a : b;
  if (k == 0) return BESS_OK;
  const int blocks = stream_blocks((int64_t)n_out * k);
  cudaStream_t st = (cudaStream_t)stream;
  switch (elem_bytes) {
    case 1: take_along_rows_kernel<uint8_t><<<blocks, 256, 0, st>>>((const uint8_t*)x, a, e, index, b, k, n_out, (uint8_t*)out); break;
    case 2: take_along_rows_kernel<uint16_t><<<blocks, 256, 0, st>>>((const uint16_t*)x, a, e, index, b, k, n_out, (uint16_t*)out); break;
    case 4: take_along_rows_kernel<uint32_t><<<blocks, 256, 0, st>>>((const uint32_t*)x, a, e, index, b, k, n_out, (uint32_t*)out); break;
    case 8: take_along_rows_kernel<uint64_t><<<blocks, 256, 0, st>>>((const uint64_t*)x, a, e, index, b, k, n_out, (uint64_t*)out); break;
    default: bess_set_error("bess_take_along_rows: element size %d", elem_bytes); return BESS_ERR_INVALID_ARG;
  }
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

#define DISPATCH_DTYPE(dtype, ...)                                   \
  switch (dtype) {                                                   \
    case BESS_F32: { using T = float; __VA_ARGS__; break; }          \
    case BESS_F16: { using T = __half; __VA_ARGS__; break; }         \
    case BESS_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; } \
    default: bess_set_error("unknown dtype %d", dtype); return BESS_ERR_INVALID_ARG; \
  }

static inline dim3 warp_grid(int n_rows) { return dim3(ceil_div((int64_t)n_rows * 32, 256)); }

extern "C" int bess_complex_mul(int dtype, const void* v1, const void* v2, int n, int e, int rotate,
                                void* out, void* stream) {
  if (n == 0 || e == 0) return BESS_OK;
  const int blocks = stream_blocks((int64_t)n * e);
  DISPATCH_DTYPE(dtype, complex_mul_kernel<T><<<blocks, 256, 0, (cudaStream_t)stream>>>(
                            static_cast<const T*>(v1), static_cast<const T*>(v2), n, e, rotate,
                            static_cast<T*>(out)));
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

extern "C" int bess_score_triple_fwd(const bess_score_cfg_t* cfg, int dtype, bess_rows_t head,
                                     bess_rows_t tail, const void* rel_table,
                                     const int32_t* rel_id, bess_rowmap_t rel_map, int n,
                                     float* score, bess_rowmap_t score_map, void* stream) {
  if (int e = check_cfg(cfg)) return e;
  if (n == 0) return BESS_OK;
  const FamCfg f = to_cfg(cfg);
  DISPATCH_DTYPE(dtype, triple_fwd_kernel<T><<<warp_grid(n), 256, 0, (cudaStream_t)stream>>>(
                            f, head, tail, static_cast<const T*>(rel_table), rel_width(f), rel_id,
                            rel_map, n, score, score_map));
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

extern "C" int bess_score_triple_bwd(const bess_score_cfg_t* cfg, int dtype, bess_rows_t head,
                                     bess_rows_t tail, const void* rel_table,
                                     const int32_t* rel_id, bess_rowmap_t rel_map, int n,
                                     const float* score, const float* d_score,
                                     bess_rowmap_t score_map, bess_rows_t d_head,
                                     bess_rows_t d_tail, float* d_rel, int add_head, int add_tail,
                                     int add_rel, void* stream) {
  if (int e = check_cfg(cfg)) return e;
  if (n == 0) return BESS_OK;
  const FamCfg f = to_cfg(cfg);
  DISPATCH_DTYPE(dtype, triple_bwd_kernel<T><<<warp_grid(n), 256, 0, (cudaStream_t)stream>>>(
                            f, head, tail, static_cast<const T*>(rel_table), rel_width(f), rel_id,
                            rel_map, n, score, d_score, score_map, d_head, d_tail, d_rel,
                            rel_width(f), add_head, add_tail, add_rel));
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

extern "C" int bess_query_prologue_fwd(const bess_score_cfg_t* cfg, int dtype, int mode,
                                       bess_rows_t fixed, const void* rel_table,
                                       const int32_t* rel_id, bess_rowmap_t rel_map, int n,
                                       float* qv, void* stream) {
  if (int e = check_cfg(cfg)) return e;
  if (n == 0) return BESS_OK;
  const FamCfg f = to_cfg(cfg);
  const int qv_row = pair_nvec(f) * ent_width(f);
  DISPATCH_DTYPE(dtype, prologue_fwd_kernel<T><<<warp_grid(n), 256, 0, (cudaStream_t)stream>>>(
                            f, mode, fixed, static_cast<const T*>(rel_table), rel_width(f), rel_id,
                            rel_map, n, qv, qv_row));
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

extern "C" int bess_query_prologue_bwd(const bess_score_cfg_t* cfg, int dtype, int mode,
                                       bess_rows_t fixed, const void* rel_table,
                                       const int32_t* rel_id, bess_rowmap_t rel_map, int n,
                                       const float* d_qv, bess_rows_t d_fixed, float* d_rel,
                                       int add_fixed, int add_rel, void* stream) {
  if (int e = check_cfg(cfg)) return e;
  if (n == 0) return BESS_OK;
  const FamCfg f = to_cfg(cfg);
  const int qv_row = pair_nvec(f) * ent_width(f);
  DISPATCH_DTYPE(dtype, prologue_bwd_kernel<T><<<warp_grid(n), 256, 0, (cudaStream_t)stream>>>(
                            f, mode, fixed, static_cast<const T*>(rel_table), rel_width(f), rel_id,
                            rel_map, n, d_qv, qv_row, d_fixed, d_rel, rel_width(f), add_fixed,
                            add_rel));
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

extern "C" int bess_boxe_rel_finalize(const bess_score_cfg_t* cfg, int dtype,
                                      const void* rel_table, const int32_t* rel_id, int n,
                                      float* d_rel, void* stream) {
  if (int e = check_cfg(cfg)) return e;
  BESS_CHECK_ARG(cfg->family == BESS_BOXE, "bess_boxe_rel_finalize needs a BoxE config");
  if (n == 0) return BESS_OK;
  const FamCfg f = to_cfg(cfg);
  DISPATCH_DTYPE(dtype, boxe_finalize_kernel<T><<<warp_grid(n), 256, 0, (cudaStream_t)stream>>>(
                            f, static_cast<const T*>(rel_table), rel_width(f), rel_id, n, d_rel,
                            rel_width(f)));
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

extern "C" int bess_cand_inv_norm(int dtype, bess_rows_t cand, int n, int width, float* inv_norm,
                                  void* stream) {
  if (n == 0) return BESS_OK;
  DISPATCH_DTYPE(dtype, cand_inv_norm_kernel<T><<<warp_grid(n), 256, 0, (cudaStream_t)stream>>>(
                            cand, n, width, inv_norm));
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

extern "C" int bess_cand_norm_bwd(int dtype, bess_rows_t cand, int n, int width,
                                  const float* inv_norm, bess_rows_t d_cand, void* stream) {
  if (n == 0) return BESS_OK;
  DISPATCH_DTYPE(dtype, cand_norm_bwd_kernel<T><<<warp_grid(n), 256, 0, (cudaStream_t)stream>>>(
                            cand, n, width, inv_norm, d_cand));
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}
