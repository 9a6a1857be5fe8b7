// Shared device helpers for the BESS-KGE sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/besskge_b200.h"

#define BESS_HD __host__ __device__ __forceinline__
#define BESS_D __device__ __forceinline__

namespace bess {

constexpr int kWarp = 32;
constexpr int kNumSM = 148;  // B200

// ---------------------------------------------------------------------------
// Row addressing (see bess_rowmap_t).  Used for the [n, p + B*Nn, D]
// exchange-buffer layout (bess.py:348-360) without materialising views.
// ---------------------------------------------------------------------------
BESS_HD int map_row(const bess_rowmap_t& m, int x) {
  if (m.group <= 0) return x + m.offset;
  int r = m.offset;
  if (m.group1 > 0) {
    const int a = x / m.group1;
    r += a * m.stride1;
    x -= a * m.group1;
  }
  const int g = x / m.group;
  return r + g * m.stride + (x - g * m.group);
}

// Resolve the storage row of logical row x of a row source.
BESS_D int64_t src_row(const bess_rows_t& s, int x) {
  int r = map_row(s.map, x);
  if (s.idx != nullptr) r = __ldg(s.idx + r);
  return (int64_t)r;
}

// ---------------------------------------------------------------------------
// dtype traits: tables may be fp32 / fp16 / bf16; all arithmetic is fp32.
// ---------------------------------------------------------------------------
template <typename T>
struct Elem;
template <>
struct Elem<float> {
  static constexpr int kVec = 4;  // elements per 128-bit access
  static BESS_D void load_vec(const float* p, float (&out)[4]) {
    float4 v = *reinterpret_cast<const float4*>(p);
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
  }
  static BESS_D void unpack(const uint4& v, float (&out)[4]) {
    out[0] = __uint_as_float(v.x); out[1] = __uint_as_float(v.y);
    out[2] = __uint_as_float(v.z); out[3] = __uint_as_float(v.w);
  }
  static BESS_D float to_f(float v) { return v; }
  static BESS_D float from_f(float v) { return v; }
};
template <>
struct Elem<__half> {
  static constexpr int kVec = 8;
  static BESS_D void load_vec(const __half* p, float (&out)[8]) {
    uint4 v = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __half22float2(h[i]);
      out[2 * i] = f.x; out[2 * i + 1] = f.y;
    }
  }
  static BESS_D void unpack(const uint4& v, float (&out)[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __half22float2(h[i]);
      out[2 * i] = f.x; out[2 * i + 1] = f.y;
    }
  }
  static BESS_D float to_f(__half v) { return __half2float(v); }
  static BESS_D __half from_f(float v) { return __float2half_rn(v); }
};
template <>
struct Elem<__nv_bfloat16> {
  static constexpr int kVec = 8;
  static BESS_D void load_vec(const __nv_bfloat16* p, float (&out)[8]) {
    uint4 v = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      out[2 * i] = f.x; out[2 * i + 1] = f.y;
    }
  }
  static BESS_D void unpack(const uint4& v, float (&out)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      out[2 * i] = f.x; out[2 * i + 1] = f.y;
    }
  }
  static BESS_D float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static BESS_D __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};

template <typename T>
BESS_D float ldf(const T* p) { return Elem<T>::to_f(*p); }

// 128-bit streaming accesses (rows are read once: keep them out of L1)
BESS_D uint4 ld_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
BESS_D void st_stream(uint4* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w));
}

BESS_D float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
BESS_D float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide reductions in a FIXED order (deterministic): warp shuffles then a
// serial pass over the per-warp partials by warp 0.
template <int kThreads>
BESS_D float block_sum(float v, float* smem /* >= kThreads/32 floats */) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem[w] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < kThreads / 32; ++i) t += smem[i];
  return t;
}
template <int kThreads>
BESS_D float block_max(float v, float* smem) {
  v = warp_max(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem[w] = v;
  __syncthreads();
  float t = smem[0];
#pragma unroll
  for (int i = 1; i < kThreads / 32; ++i) t = fmaxf(t, smem[i]);
  return t;
}

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace bess

// thread-local last-error string, set by the C-ABI layer
void bess_set_error(const char* fmt, ...);

#define BESS_CHECK_ARG(cond, ...)      \
  do {                                 \
    if (!(cond)) {                     \
      bess_set_error(__VA_ARGS__);     \
      return BESS_ERR_INVALID_ARG;     \
    }                                  \
  } while (0)

// kernels launched by this library since load (host-side count; bess_launch_count())
extern unsigned long long g_bess_launches;
#define BESS_LAUNCHED(n) (g_bess_launches += (unsigned long long)(n))

#define BESS_CHECK_LAUNCH()                                             \
  do {                                                                  \
    BESS_LAUNCHED(1);                                                   \
    cudaError_t _e = cudaPeekAtLastError();                             \
    if (_e != cudaSuccess) {                                            \
      bess_set_error("CUDA launch failed: %s", cudaGetErrorString(_e)); \
      return BESS_ERR_CUDA;                                             \
    }                                                                   \
  } while (0)
