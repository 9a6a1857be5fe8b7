// Backward tail of the BESS step: deterministic segmented scatter-add of the
// gathered-row gradients into the shard, fused with the optimizer update.
//
// The reference has no such code: `entity_embedding[gather_idx]` (bess.py:333)
// is differentiated by (Pop)Torch autograd, which materialises a DENSE
// [Es, D] gradient (index_put_ accumulate) that the dense optimizer then
// consumes.  Here the gradient never becomes dense:
//   1. stable LSD radix sort of the step's gather indices (key = local row);
//   2. one warp per run of equal keys sums its fp32 gradient rows in sorted
//      (= gather) order -> bit-reproducible, no atomics;
//   3. plain SGD updates the touched rows in place (identical to the dense
//      update, untouched rows see g = 0); momentum / AdamW run a dense
//      streaming pass with the segment sums scattered in (dense semantics).
#include "common.cuh"

namespace bess {

// ---------------------------------------------------------------------------
// Stable LSD radix sort, 8 bits per pass, (key, position) pairs.
// ---------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // keys per block
constexpr int RS_RADIX = 256;

__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const int32_t* keys, int n, int shift,
                                                              int n_blocks, int32_t* block_hist) {
  __shared__ int hist[RS_RADIX];
  hist[threadIdx.x] = 0;
  __syncthreads();
  const int base = blockIdx.x * RS_TILE;
#pragma unroll
  for (int it = 0; it < RS_ITEMS; ++it) {
    const int i = base + it * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&hist[(keys[i] >> shift) & (RS_RADIX - 1)], 1);
  }
  __syncthreads();
  block_hist[threadIdx.x * n_blocks + blockIdx.x] = hist[threadIdx.x];  // digit-major
}

// exclusive scan of m ints in one CTA (m = 256 * n_blocks)
__global__ void __launch_bounds__(1024) rs_scan_kernel(int32_t* data, int m) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int base = 0; base < m; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < m ? data[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[w] = x;
    __syncthreads();
    if (w == 0) {
      int t = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      warp_tot[lane] = t;  // inclusive over warps
    }
    __syncthreads();
    const int carry = carry_s;
    const int warp_off = w == 0 ? 0 : warp_tot[w - 1];
    if (i < m) data[i] = carry + warp_off + x - v;  // exclusive
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + warp_off + x;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(
    const int32_t* keys_in, const int32_t* vals_in /* null: identity */, int n, int shift,
    int n_blocks, const int32_t* block_off, int32_t* keys_out, int32_t* vals_out) {
  __shared__ int warp_cnt[RS_THREADS / 32][RS_RADIX];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (RS_THREADS / 32) * RS_RADIX; i += RS_THREADS)
    (&warp_cnt[0][0])[i] = 0;
  __syncthreads();
  // warp w owns the contiguous keys [base + w*256, base + (w+1)*256), 32 at a time
  const int base = blockIdx.x * RS_TILE + w * (RS_ITEMS * 32);
  int key[RS_ITEMS], digit[RS_ITEMS];
#pragma unroll
  for (int it = 0; it < RS_ITEMS; ++it) {
    const int i = base + it * 32 + lane;
    key[it] = i < n ? keys_in[i] : 0;
    digit[it] = i < n ? ((key[it] >> shift) & (RS_RADIX - 1)) : RS_RADIX;  // RS_RADIX = inactive
  }
  // pass A: per-warp digit counts
#pragma unroll
  for (int it = 0; it < RS_ITEMS; ++it) {
    const unsigned peers = __match_any_sync(0xffffffffu, digit[it]);
    if (digit[it] < RS_RADIX && lane == __ffs(peers) - 1) warp_cnt[w][digit[it]] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  // pass B: thread d turns the counts of digit d into running global offsets
  {
    const int d = threadIdx.x;
    int run = block_off[d * n_blocks + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < RS_THREADS / 32; ++ww) {
      const int c = warp_cnt[ww][d];
      warp_cnt[ww][d] = run;
      run += c;
    }
  }
  __syncthreads();
  // pass C: stable placement
#pragma unroll
  for (int it = 0; it < RS_ITEMS; ++it) {
    const int i = base + it * 32 + lane;
    const unsigned peers = __match_any_sync(0xffffffffu, digit[it]);
    if (digit[it] < RS_RADIX) {
      const int rank = __popc(peers & ((1u << lane) - 1u));
      const int pos = warp_cnt[w][digit[it]] + rank;
      keys_out[pos] = key[it];
      vals_out[pos] = vals_in != nullptr ? vals_in[i] : i;
    }
    __syncwarp();
    if (digit[it] < RS_RADIX && lane == __ffs(peers) - 1) warp_cnt[w][digit[it]] += __popc(peers);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------
// Segmented reduce over sorted keys.  Warp per sorted position; only the
// first position of each run works.
// ---------------------------------------------------------------------------
struct GradSrc {
  const float* local;
  const float* dst;
  int n_local;
  int per_dst;
  int64_t dst_stride_rows;
  int row_elems;
};
BESS_D const float* grad_row(const GradSrc& g, int pos) {
  if (pos < g.n_local) return g.local + (int64_t)pos * g.row_elems;
  const int y = pos - g.n_local;
  const int d = y / g.per_dst;
  const int i = y - d * g.per_dst;
  return g.dst + ((int64_t)d * g.dst_stride_rows + i) * g.row_elems;
}

// w[0..3] -= lr * s as ONE 128-bit (fp32) or 64-bit (half) read-modify-write
template <typename T>
BESS_D void segment_apply_sgd(T* row, const float4& s, float lr);
template <>
BESS_D void segment_apply_sgd<float>(float* row, const float4& s, float lr) {
  float4 w = *reinterpret_cast<float4*>(row);
  w.x -= lr * s.x; w.y -= lr * s.y; w.z -= lr * s.z; w.w -= lr * s.w;
  *reinterpret_cast<float4*>(row) = w;
}
template <>
BESS_D void segment_apply_sgd<__half>(__half* row, const float4& s, float lr) {
  uint2 raw = *reinterpret_cast<uint2*>(row);
  __half2* h = reinterpret_cast<__half2*>(&raw);
  float2 a = __half22float2(h[0]), b = __half22float2(h[1]);
  h[0] = __halves2half2(__float2half_rn(a.x - lr * s.x), __float2half_rn(a.y - lr * s.y));
  h[1] = __halves2half2(__float2half_rn(b.x - lr * s.z), __float2half_rn(b.y - lr * s.w));
  *reinterpret_cast<uint2*>(row) = raw;
}
template <>
BESS_D void segment_apply_sgd<__nv_bfloat16>(__nv_bfloat16* row, const float4& s, float lr) {
  uint2 raw = *reinterpret_cast<uint2*>(row);
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
  float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  h[0] = __halves2bfloat162(__float2bfloat16_rn(a.x - lr * s.x), __float2bfloat16_rn(a.y - lr * s.y));
  h[1] = __halves2bfloat162(__float2bfloat16_rn(b.x - lr * s.z), __float2bfloat16_rn(b.y - lr * s.w));
  *reinterpret_cast<uint2*>(row) = raw;
}

// MODE 0: SGD update of the table row; MODE 1: store the segment sum; MODE 2: add the segment
// sum to row `key` of a dense fp32 accumulator (gradient accumulation over micro-batches: one
// warp owns a key, so the += needs no atomics and the result is bit-reproducible)
template <int MODE, typename T>
__global__ void __launch_bounds__(256) segment_kernel(T* table, int64_t pitch,
                                                       const int32_t* sorted_keys,
                                                       const int32_t* perm, int n, GradSrc g,
                                                       float lr, const float* hyper,
                                                       float* seg_grad, int32_t* row_to_seg) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= n) return;
  if (MODE == 0 && hyper != nullptr) lr = __ldg(hyper + BESS_HYPER_LR);
  const int lane = threadIdx.x & 31;
  const int key = sorted_keys[wid];
  if (wid > 0 && sorted_keys[wid - 1] == key) return;  // not a run head
  int end = wid + 1;
  while (end < n && sorted_keys[end] == key) ++end;
  const int W = g.row_elems;
  if (MODE == 1 && lane == 0) row_to_seg[key] = wid;
  // two 128-column chunks per pass: their gradient loads (and the table row's) are independent,
  // and the permutation entry is read once for both
  for (int k0 = lane * 4; k0 < W; k0 += 256) {
    const int k1 = k0 + 128;
    const bool two = k1 < W;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = wid; i < end; ++i) {
      const float* grow = grad_row(g, perm[i]);
      const float4 v = *reinterpret_cast<const float4*>(grow + k0);
      float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
      if (two) u = *reinterpret_cast<const float4*>(grow + k1);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    if (MODE == 0) {
      segment_apply_sgd<T>(table + (int64_t)key * pitch + k0, s, lr);
      if (two) segment_apply_sgd<T>(table + (int64_t)key * pitch + k1, t, lr);
    } else if (MODE == 2) {
      float4* a0 = reinterpret_cast<float4*>(seg_grad + (int64_t)key * W + k0);
      float4 c = *a0;
      c.x += s.x; c.y += s.y; c.z += s.z; c.w += s.w;
      *a0 = c;
      if (two) {
        float4* a1 = reinterpret_cast<float4*>(seg_grad + (int64_t)key * W + k1);
        float4 e = *a1;
        e.x += t.x; e.y += t.y; e.z += t.z; e.w += t.w;
        *a1 = e;
      }
    } else {
      *reinterpret_cast<float4*>(seg_grad + (int64_t)wid * W + k0) = s;
      if (two) *reinterpret_cast<float4*>(seg_grad + (int64_t)wid * W + k1) = t;
    }
  }
}

// Dense optimizer pass (torch.optim.SGD with momentum / AdamW semantics).
// Hyper-parameters come by value or, when `hyper` is non-null, from a small device array
// (BESS_HYPER_* slots) that the host rewrites before every replay of a captured step, so a
// learning-rate schedule and Adam's bias correction survive CUDA-graph capture.
struct OptHyper {
  float lr, momentum, dampening, beta1, beta2, eps, wd, bc1, bc2;
  int first_step;
};
BESS_D OptHyper load_hyper(OptHyper h, const float* hyper) {
  if (hyper != nullptr) {
    h.lr = __ldg(hyper + BESS_HYPER_LR); h.momentum = __ldg(hyper + BESS_HYPER_MOMENTUM);
    h.dampening = __ldg(hyper + BESS_HYPER_DAMPENING); h.beta1 = __ldg(hyper + BESS_HYPER_BETA1);
    h.beta2 = __ldg(hyper + BESS_HYPER_BETA2); h.eps = __ldg(hyper + BESS_HYPER_EPS);
    h.wd = __ldg(hyper + BESS_HYPER_WEIGHT_DECAY); h.bc1 = __ldg(hyper + BESS_HYPER_BC1);
    h.bc2 = __ldg(hyper + BESS_HYPER_BC2);
    h.first_step = __ldg(hyper + BESS_HYPER_FIRST_STEP) != 0.f;
  }
  return h;
}
BESS_D float opt_update(int kind, const OptHyper& h, float wv, float gval, float& s0, float& s1) {
  if (kind == BESS_OPT_SGD) {
    gval += h.wd * wv;
    wv -= h.lr * gval;
  } else if (kind == BESS_OPT_SGDM) {
    gval += h.wd * wv;
    const float b = h.first_step ? gval : h.momentum * s0 + (1.f - h.dampening) * gval;
    s0 = b;
    wv -= h.lr * b;
  } else {  // AdamW (decoupled weight decay), torch.optim.AdamW
    wv *= 1.f - h.lr * h.wd;
    const float m = h.beta1 * s0 + (1.f - h.beta1) * gval;
    const float v = h.beta2 * s1 + (1.f - h.beta2) * gval * gval;
    s0 = m; s1 = v;
    const float denom = sqrtf(v) / sqrtf(h.bc2) + h.eps;
    wv -= (h.lr / h.bc1) * (m / denom);
  }
  return wv;
}

template <typename T>
struct Vec4;  // 4 table elements as one 128-bit (fp32) / 64-bit (half) access
template <>
struct Vec4<float> {
  static BESS_D void load(const float* p, float (&w)[4]) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
  }
  static BESS_D void store(float* p, const float (&w)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(w[0], w[1], w[2], w[3]);
  }
};
template <>
struct Vec4<__half> {
  static BESS_D void load(const __half* p, float (&w)[4]) {
    const uint2 raw = *reinterpret_cast<const uint2*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&raw);
    const float2 a = __half22float2(h[0]), b = __half22float2(h[1]);
    w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y;
  }
  static BESS_D void store(__half* p, const float (&w)[4]) {
    uint2 raw;
    __half2* h = reinterpret_cast<__half2*>(&raw);
    h[0] = __halves2half2(__float2half_rn(w[0]), __float2half_rn(w[1]));
    h[1] = __halves2half2(__float2half_rn(w[2]), __float2half_rn(w[3]));
    *reinterpret_cast<uint2*>(p) = raw;
  }
};
template <>
struct Vec4<__nv_bfloat16> {
  static BESS_D void load(const __nv_bfloat16* p, float (&w)[4]) {
    const uint2 raw = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
    const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
    w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y;
  }
  static BESS_D void store(__nv_bfloat16* p, const float (&w)[4]) {
    uint2 raw;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
    h[0] = __halves2bfloat162(__float2bfloat16_rn(w[0]), __float2bfloat16_rn(w[1]));
    h[1] = __halves2bfloat162(__float2bfloat16_rn(w[2]), __float2bfloat16_rn(w[3]));
    *reinterpret_cast<uint2*>(p) = raw;
  }
};

// VEC: 4 elements per thread with 128-bit state / gradient accesses (rows a multiple of 4
// elements, 16-byte aligned); otherwise one element per thread.  Warp-coalesced along the
// row: a pure HBM stream of table (s) + state (4 / 8) bytes per element, read + write.
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) opt_dense_kernel(int kind, T* table, int64_t pitch,
                                                         int n_rows, int W, const float* seg_grad,
                                                         const int32_t* row_to_seg, float* state0,
                                                         float* state1, OptHyper hv,
                                                         const float* hyper, float grad_scale,
                                                         float* zero_grad) {
  const OptHyper h = load_hyper(hv, hyper);
  if (VEC) {
    const int W4 = W >> 2;
    const int64_t total = (int64_t)n_rows * W4;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
      const int row = (int)(t / W4), k = (int)(t - (int64_t)row * W4) << 2;
      const int seg = row_to_seg != nullptr ? __ldg(row_to_seg + row) : row;
      float g[4] = {0.f, 0.f, 0.f, 0.f}, w[4], s0[4] = {0.f, 0.f, 0.f, 0.f},
            s1[4] = {0.f, 0.f, 0.f, 0.f};
      if (seg >= 0) Vec4<float>::load(seg_grad + (int64_t)seg * W + k, g);
      if (zero_grad != nullptr) {  // dense accumulator: consumed, cleared for the next cycle
        const float z[4] = {0.f, 0.f, 0.f, 0.f};
        Vec4<float>::store(zero_grad + (int64_t)seg * W + k, z);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) g[i] *= grad_scale;
      T* pw = table + (int64_t)row * pitch + k;
      Vec4<T>::load(pw, w);
      const int64_t so = (int64_t)row * W + k;
      if (kind != BESS_OPT_SGD && !(kind == BESS_OPT_SGDM && h.first_step))
        Vec4<float>::load(state0 + so, s0);
      if (kind == BESS_OPT_ADAMW) Vec4<float>::load(state1 + so, s1);
#pragma unroll
      for (int i = 0; i < 4; ++i) w[i] = opt_update(kind, h, w[i], g[i], s0[i], s1[i]);
      Vec4<T>::store(pw, w);
      if (kind != BESS_OPT_SGD) Vec4<float>::store(state0 + so, s0);
      if (kind == BESS_OPT_ADAMW) Vec4<float>::store(state1 + so, s1);
    }
  } else {
    const int64_t total = (int64_t)n_rows * W;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
      const int row = (int)(t / W), k = (int)(t - (int64_t)row * W);
      const int seg = row_to_seg != nullptr ? row_to_seg[row] : row;
      const float gval = (seg >= 0 ? seg_grad[(int64_t)seg * W + k] : 0.f) * grad_scale;
      if (zero_grad != nullptr) zero_grad[(int64_t)seg * W + k] = 0.f;
      T* pw = table + (int64_t)row * pitch + k;
      float s0 = kind != BESS_OPT_SGD ? state0[t] : 0.f;
      float s1 = kind == BESS_OPT_ADAMW ? state1[t] : 0.f;
      const float wv = opt_update(kind, h, Elem<T>::to_f(*pw), gval, s0, s1);
      *pw = Elem<T>::from_f(wv);
      if (kind != BESS_OPT_SGD) state0[t] = s0;
      if (kind == BESS_OPT_ADAMW) state1[t] = s1;
    }
  }
}

struct HyperVals {
  float v[BESS_HYPER_COUNT];
};
__global__ void set_hyper_kernel(float* hyper, HyperVals h) {
  if (threadIdx.x < BESS_HYPER_COUNT) hyper[threadIdx.x] = h.v[threadIdx.x];
}

// Relation-table gradient: CTA (relation r, 128-column block) reduces the sorted run of
// per-query rows with RR_WARPS warps in a fixed interleaved order (warp w takes rows w,
// w + RR_WARPS, ... of the run; the per-warp sums are added in warp order): deterministic.
// The loop is a chain of dependent perm -> row loads (the rows sit anywhere in the [S, Wr]
// buffer), so latency rules: 16 warps x 8 rows in flight each (was 8 x 4: ~21 us per launch
// at biokg's 51 relations x 321 rows, now a few round trips).
constexpr int RR_WARPS = 16, RR_FLIGHT = 8;
__global__ void __launch_bounds__(RR_WARPS * 32) relation_reduce_kernel(
    const float* rows, int width, const int32_t* sorted_rel, const int32_t* perm, int n,
    float* d_table) {
  __shared__ float part[RR_WARPS][128];
  const int r = blockIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // run [lo, hi) of relation r by binary search (warp-uniform)
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (sorted_rel[mid] < r) lo = mid + 1; else hi = mid; }
  const int start = lo;
  hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (sorted_rel[mid] <= r) lo = mid + 1; else hi = mid; }
  const int end = lo;
  const int k0 = blockIdx.y * 128 + lane * 4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (k0 < width) {
    int i = start + w;
    if ((width & 3) == 0) {
      for (; i + (RR_FLIGHT - 1) * RR_WARPS < end; i += RR_FLIGHT * RR_WARPS) {
        float4 v[RR_FLIGHT];
#pragma unroll
        for (int u = 0; u < RR_FLIGHT; ++u)
          v[u] = *reinterpret_cast<const float4*>(rows + (int64_t)perm[i + u * RR_WARPS] * width + k0);
#pragma unroll
        for (int u = 0; u < RR_FLIGHT; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
      }
    }
    for (; i < end; i += RR_WARPS) {
      const float* row = rows + (int64_t)perm[i] * width + k0;
      if ((width & 3) == 0) {  // rows 16-byte aligned
        const float4 v = *reinterpret_cast<const float4*>(row);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      } else {
        s.x += row[0];
        if (k0 + 1 < width) s.y += row[1];
        if (k0 + 2 < width) s.z += row[2];
        if (k0 + 3 < width) s.w += row[3];
      }
    }
  }
  part[w][lane * 4 + 0] = s.x; part[w][lane * 4 + 1] = s.y;
  part[w][lane * 4 + 2] = s.z; part[w][lane * 4 + 3] = s.w;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int k = blockIdx.y * 128 + threadIdx.x;
    if (k < width) {
      float t = 0.f;
#pragma unroll
      for (int ww = 0; ww < RR_WARPS; ++ww) t += part[ww][threadIdx.x];
      d_table[(int64_t)r * width + k] = t;
    }
  }
}

}  // namespace bess

using namespace bess;

extern "C" int64_t bess_sort_workspace(int n) {
  const int n_blocks = ceil_div(n > 0 ? n : 1, RS_TILE);
  // block histogram + two ping-pong (key, value) buffers
  return (int64_t)RS_RADIX * n_blocks * 4 + 4 * (int64_t)(n > 0 ? n : 1) * 4 + 256;
}

extern "C" int bess_sort_keys(const int32_t* keys, int n, int key_bits, int32_t* keys_out,
                              int32_t* perm_out, void* workspace, void* stream) {
  if (n == 0) return BESS_OK;
  BESS_CHECK_ARG(key_bits >= 1 && key_bits <= 31, "key_bits %d", key_bits);
  BESS_CHECK_ARG(workspace != nullptr, "workspace required");
  cudaStream_t st = (cudaStream_t)stream;
  const int n_blocks = ceil_div(n, RS_TILE);
  int32_t* hist = static_cast<int32_t*>(workspace);
  int32_t* buf = hist + (((int64_t)RS_RADIX * n_blocks + 63) / 64) * 64;
  int32_t* kA = buf; int32_t* vA = buf + n; int32_t* kB = buf + 2 * (int64_t)n; int32_t* vB = buf + 3 * (int64_t)n;
  const int passes = (key_bits + 7) / 8;
  const int32_t* kin = keys; const int32_t* vin = nullptr;
  for (int p = 0; p < passes; ++p) {
    const bool last = p == passes - 1;
    int32_t* kout = last ? keys_out : ((p & 1) ? kB : kA);
    int32_t* vout = last ? perm_out : ((p & 1) ? vB : vA);
    rs_hist_kernel<<<n_blocks, RS_THREADS, 0, st>>>(kin, n, 8 * p, n_blocks, hist);
    rs_scan_kernel<<<1, 1024, 0, st>>>(hist, RS_RADIX * n_blocks);
    rs_scatter_kernel<<<n_blocks, RS_THREADS, 0, st>>>(kin, vin, n, 8 * p, n_blocks, hist, kout, vout);
    kin = kout; vin = vout;
  }
  BESS_LAUNCHED(3 * passes - 1);
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

static GradSrc make_src(int row_elems, int n_local, int per_dst, const float* grad_local,
                        const float* grad_dst, int64_t dst_stride_rows) {
  GradSrc g;
  g.local = grad_local; g.dst = grad_dst; g.n_local = n_local; g.per_dst = per_dst > 0 ? per_dst : 1;
  g.dst_stride_rows = dst_stride_rows; g.row_elems = row_elems;
  return g;
}

extern "C" int bess_scatter_sgd(void* table, int64_t table_pitch, int dtype, int row_elems,
                                const int32_t* sorted_keys, const int32_t* perm, int n, int n_local,
                                int per_dst, const float* grad_local, const float* grad_dst,
                                int64_t dst_stride_rows, float lr, const float* hyper,
                                void* stream) {
  if (n == 0) return BESS_OK;
  BESS_CHECK_ARG(row_elems % 4 == 0, "row_elems must be a multiple of 4");
  {  // the update is a 128-bit (fp32) / 64-bit (half) read-modify-write of 4 elements
    const int es = dtype == BESS_F32 ? 4 : 2;
    BESS_CHECK_ARG((reinterpret_cast<uintptr_t>(table) % (4 * es)) == 0 && (table_pitch % 4) == 0,
                   "bess_scatter_sgd: table rows must be aligned to 4 elements");
  }
  const GradSrc g = make_src(row_elems, n_local, per_dst, grad_local, grad_dst, dst_stride_rows);
  const int blocks = ceil_div((int64_t)n * 32, 256);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case BESS_F32: segment_kernel<0, float><<<blocks, 256, 0, st>>>((float*)table, table_pitch, sorted_keys, perm, n, g, lr, hyper, nullptr, nullptr); break;
    case BESS_F16: segment_kernel<0, __half><<<blocks, 256, 0, st>>>((__half*)table, table_pitch, sorted_keys, perm, n, g, lr, hyper, nullptr, nullptr); break;
    case BESS_BF16: segment_kernel<0, __nv_bfloat16><<<blocks, 256, 0, st>>>((__nv_bfloat16*)table, table_pitch, sorted_keys, perm, n, g, lr, hyper, nullptr, nullptr); break;
    default: bess_set_error("unknown dtype %d", dtype); return BESS_ERR_INVALID_ARG;
  }
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

extern "C" int bess_scatter_collect(int row_elems, const int32_t* sorted_keys, const int32_t* perm,
                                    int n, int n_local, int per_dst, const float* grad_local,
                                    const float* grad_dst, int64_t dst_stride_rows, float* seg_grad,
                                    int32_t* row_to_seg, void* stream) {
  if (n == 0) return BESS_OK;
  BESS_CHECK_ARG(row_elems % 4 == 0, "row_elems must be a multiple of 4");
  const GradSrc g = make_src(row_elems, n_local, per_dst, grad_local, grad_dst, dst_stride_rows);
  const int blocks = ceil_div((int64_t)n * 32, 256);
  segment_kernel<1, float><<<blocks, 256, 0, (cudaStream_t)stream>>>(
      nullptr, 0, sorted_keys, perm, n, g, 0.f, nullptr, seg_grad, row_to_seg);
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

extern "C" int bess_scatter_accumulate(int row_elems, const int32_t* sorted_keys,
                                       const int32_t* perm, int n, int n_local, int per_dst,
                                       const float* grad_local, const float* grad_dst,
                                       int64_t dst_stride_rows, float* acc, void* stream) {
  if (n == 0) return BESS_OK;
  BESS_CHECK_ARG(row_elems % 4 == 0 && (reinterpret_cast<uintptr_t>(acc) & 15) == 0,
                 "bess_scatter_accumulate: rows must be 16-byte aligned multiples of 4 elements");
  const GradSrc g = make_src(row_elems, n_local, per_dst, grad_local, grad_dst, dst_stride_rows);
  const int blocks = ceil_div((int64_t)n * 32, 256);
  segment_kernel<2, float><<<blocks, 256, 0, (cudaStream_t)stream>>>(
      nullptr, 0, sorted_keys, perm, n, g, 0.f, nullptr, acc, nullptr);
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

extern "C" int bess_opt_dense(int kind, void* table, int64_t table_pitch, int dtype, int n_rows,
                              int row_elems, const float* seg_grad, const int32_t* row_to_seg,
                              float* state0, float* state1, float lr, float momentum,
                              float dampening, float beta1, float beta2, float eps,
                              float weight_decay, int step, const float* hyper, float grad_scale,
                              int zero_grad, void* stream) {
  if (n_rows == 0) return BESS_OK;
  BESS_CHECK_ARG(!zero_grad || row_to_seg == nullptr,
                 "bess_opt_dense: zero_grad needs a dense gradient (row_to_seg == NULL)");
  float* zg = zero_grad ? const_cast<float*>(seg_grad) : nullptr;
  BESS_CHECK_ARG(kind >= BESS_OPT_SGD && kind <= BESS_OPT_ADAMW, "optimizer kind %d", kind);
  if (kind != BESS_OPT_SGD) BESS_CHECK_ARG(state0 != nullptr, "optimizer state missing");
  if (kind == BESS_OPT_ADAMW) BESS_CHECK_ARG(state1 != nullptr, "optimizer state missing");
  OptHyper h;
  h.lr = lr; h.momentum = momentum; h.dampening = dampening; h.beta1 = beta1; h.beta2 = beta2;
  h.eps = eps; h.wd = weight_decay;
  h.bc1 = 1.f - powf(beta1, (float)step); h.bc2 = 1.f - powf(beta2, (float)step);
  h.first_step = step <= 1 ? 1 : 0;
  const int es = dtype == BESS_F32 ? 4 : 2;
  const bool vec = row_elems % 4 == 0 && table_pitch % 4 == 0 &&
                   reinterpret_cast<uintptr_t>(table) % (4 * es) == 0 &&
                   reinterpret_cast<uintptr_t>(seg_grad) % 16 == 0 &&
                   reinterpret_cast<uintptr_t>(state0) % 16 == 0 &&
                   reinterpret_cast<uintptr_t>(state1) % 16 == 0;
  const int64_t total = (int64_t)n_rows * (vec ? row_elems / 4 : row_elems);
  int blocks = ceil_div(total, 256);
  if (blocks > kNumSM * 8) blocks = kNumSM * 8;
  cudaStream_t st = (cudaStream_t)stream;
#define BESS_OPT_LAUNCH(T, V)                                                                     \
  opt_dense_kernel<T, V><<<blocks, 256, 0, st>>>(kind, (T*)table, table_pitch, n_rows, row_elems, \
                                                 seg_grad, row_to_seg, state0, state1, h, hyper, \
                                                 grad_scale, zg)
  switch (dtype) {
    case BESS_F32: if (vec) BESS_OPT_LAUNCH(float, true); else BESS_OPT_LAUNCH(float, false); break;
    case BESS_F16: if (vec) BESS_OPT_LAUNCH(__half, true); else BESS_OPT_LAUNCH(__half, false); break;
    case BESS_BF16:
      if (vec) BESS_OPT_LAUNCH(__nv_bfloat16, true); else BESS_OPT_LAUNCH(__nv_bfloat16, false);
      break;
    default: bess_set_error("unknown dtype %d", dtype); return BESS_ERR_INVALID_ARG;
  }
#undef BESS_OPT_LAUNCH
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

extern "C" int bess_set_hyper(float* hyper, float lr, float momentum, float dampening, float beta1,
                              float beta2, float eps, float weight_decay, int step, void* stream) {
  BESS_CHECK_ARG(hyper != nullptr && step >= 1, "bess_set_hyper: hyper array and step >= 1 required");
  HyperVals h;
  for (int i = 0; i < BESS_HYPER_COUNT; ++i) h.v[i] = 0.f;
  h.v[BESS_HYPER_LR] = lr; h.v[BESS_HYPER_MOMENTUM] = momentum;
  h.v[BESS_HYPER_DAMPENING] = dampening; h.v[BESS_HYPER_BETA1] = beta1;
  h.v[BESS_HYPER_BETA2] = beta2; h.v[BESS_HYPER_EPS] = eps;
  h.v[BESS_HYPER_WEIGHT_DECAY] = weight_decay;
  h.v[BESS_HYPER_BC1] = 1.f - powf(beta1, (float)step);
  h.v[BESS_HYPER_BC2] = 1.f - powf(beta2, (float)step);
  h.v[BESS_HYPER_FIRST_STEP] = step <= 1 ? 1.f : 0.f;
  set_hyper_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(hyper, h);
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

extern "C" int bess_relation_grad_reduce(const float* d_rel_rows, int width,
                                         const int32_t* sorted_rel, const int32_t* perm, int n,
                                         int n_rel, float* d_table, void* stream) {
  if (n_rel == 0) return BESS_OK;
  dim3 grid(n_rel, ceil_div(width, 128));
  relation_reduce_kernel<<<grid, RR_WARPS * 32, 0, (cudaStream_t)stream>>>(d_rel_rows, width,
                                                                           sorted_rel, perm, n, d_table);
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}
