// Score post-processing: BAD_NEGATIVE_SCORE masks, fused loss forward +
// score-gradient, ranks, deterministic sums, small utilities.
// One CTA per query row of the [S, N] negative-score matrix.
#include <math_constants.h>
#include <stdlib.h>

#include <cstdlib>

#include "common.cuh"

namespace bess {

constexpr int L_THREADS = 256;

// logsigmoid(x) = min(x, 0) - log1p(exp(-|x|))
BESS_D float log_sigmoid(float x) { return fminf(x, 0.f) - log1pf(expf(-fabsf(x))); }
BESS_D float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(L_THREADS) mask_add_kernel(float* score, int n_row, int n_col,
                                                              int64_t ld, const uint8_t* mask,
                                                              int64_t ld_mask, int mask_rows,
                                                              int flag, float value) {
  const int r = blockIdx.x;
  const uint8_t* m = mask + (int64_t)(mask_rows == 1 ? 0 : r) * ld_mask;
  float* s = score + (int64_t)r * ld;
  for (int c = threadIdx.x; c < n_col; c += blockDim.x)
    if ((m[c] != 0) == (flag != 0)) s[c] += value;
}

// augment_negative: column hit for row r (bess.py:201-226).  Rows are grouped in
// blocks of `group` (= positive_per_partition for "ht", else all rows); inside a
// block the pattern of the first `half_group` rows repeats ("ht": both halves
// use the mask of the first half).  col = step * ((r / group) * half_group + (r % group) % half_group)
__global__ void mask_diag_kernel(float* score, int n_row, int64_t ld, int step, int half_group,
                                 int group, float value, int n_col) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_row) return;
  int base = r;
  if (group > 0) base = (r / group) * half_group + (r % group) % half_group;
  const int64_t col = (int64_t)step * base;
  if (col < n_col) score[(int64_t)r * ld + col] += value;
}

// ---------------------------------------------------------------------------
// Loss forward + gradient w.r.t. scores.  One CTA per query row.  The row is
// read ONCE with 128-bit loads into registers (up to L_CACHE * L_THREADS
// scores; longer rows re-read the tail from L2) and the gradient is written
// once, either as plain fp32 or directly in the operand form the tcgen05
// backward contractions consume (GRAD_TF32: hi = rna_tf32(g), lo =
// rna_tf32(g - hi); GRAD_BF16 / GRAD_F16: rounded halves), which removes a
// separate split pass over the [S, N] gradient.
// ---------------------------------------------------------------------------
constexpr int L_CACHE = 16;  // scores per thread kept in registers
enum GradOut { GRAD_F32 = 0, GRAD_TF32 = 1, GRAD_BF16 = 2, GRAD_F16 = 3, GRAD_F16X3 = 4 };

BESS_D float rna_tf32_(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

template <int OUT>
BESS_D void store_grad4(void* g_hi, void* g_lo, int64_t at, const float (&g)[4], int n_valid, bool vec,
                        float gscale) {
  if (OUT == GRAD_F16X3) {
    // scaled fp16 pair: hi = fp16(g * s), lo = fp16(g * s - hi)  (3xFP16 operand, gemm_tc.cu)
    __half h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float x = g[j] * gscale;
      h[j] = __float2half_rn(x);
      l[j] = __float2half_rn(x - __half2float(h[j]));
    }
    __half* ph = reinterpret_cast<__half*>(g_hi) + at;
    __half* pl = reinterpret_cast<__half*>(g_lo) + at;
    if (vec && n_valid == 4) {
      uint2 uh, ul;
      __half2 a = __halves2half2(h[0], h[1]), b = __halves2half2(h[2], h[3]);
      uh.x = *reinterpret_cast<uint32_t*>(&a); uh.y = *reinterpret_cast<uint32_t*>(&b);
      a = __halves2half2(l[0], l[1]); b = __halves2half2(l[2], l[3]);
      ul.x = *reinterpret_cast<uint32_t*>(&a); ul.y = *reinterpret_cast<uint32_t*>(&b);
      *reinterpret_cast<uint2*>(ph) = uh;
      *reinterpret_cast<uint2*>(pl) = ul;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < n_valid) { ph[j] = h[j]; pl[j] = l[j]; }
    }
  } else if (OUT == GRAD_F32 || OUT == GRAD_TF32) {
    float h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      h[j] = OUT == GRAD_TF32 ? rna_tf32_(g[j]) : g[j];
      l[j] = OUT == GRAD_TF32 ? rna_tf32_(g[j] - h[j]) : 0.f;
    }
    float* ph = reinterpret_cast<float*>(g_hi) + at;
    float* pl = reinterpret_cast<float*>(g_lo) + at;
    if (vec && n_valid == 4) {
      *reinterpret_cast<float4*>(ph) = make_float4(h[0], h[1], h[2], h[3]);
      if (OUT == GRAD_TF32) *reinterpret_cast<float4*>(pl) = make_float4(l[0], l[1], l[2], l[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < n_valid) {
          ph[j] = h[j];
          if (OUT == GRAD_TF32) pl[j] = l[j];
        }
    }
  } else if (OUT == GRAD_BF16) {
    __nv_bfloat16* ph = reinterpret_cast<__nv_bfloat16*>(g_hi) + at;
    if (vec && n_valid == 4) {
      __nv_bfloat162 a = __floats2bfloat162_rn(g[0], g[1]), b = __floats2bfloat162_rn(g[2], g[3]);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&a);
      u.y = *reinterpret_cast<uint32_t*>(&b);
      *reinterpret_cast<uint2*>(ph) = u;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < n_valid) ph[j] = __float2bfloat16_rn(g[j]);
    }
  } else {
    __half* ph = reinterpret_cast<__half*>(g_hi) + at;
    if (vec && n_valid == 4) {
      __half2 a = __floats2half2_rn(g[0], g[1]), b = __floats2half2_rn(g[2], g[3]);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&a);
      u.y = *reinterpret_cast<uint32_t*>(&b);
      *reinterpret_cast<uint2*>(ph) = u;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < n_valid) ph[j] = __float2half_rn(g[j]);
    }
  }
}

struct LossArgs {
  float margin;
  int adversarial;
  float adv_scale, loss_scale, ce_shift;
  const float* pos;
  float* neg;
  int n, n_neg;
  int64_t ld;
  const float* weight;
  int weight_n;
  float* row_loss;
  float* d_pos;
  void* g_hi;
  void* g_lo;
  int64_t ld_g;
  const float* g_scale;  // GRAD_F16X3: {scale, 1 / scale} of the gradient operand (device)
  int full_rows;         // n_neg == L_CACHE * TPB and every row 16-byte aligned in and out
};

// registers <- one row of scores: slot (i, j) <-> column (i * TPB + tid) * 4 + j
template <int TPB>
BESS_D void loss_load_row(const float* nrow, int n_neg, bool vec_in, float (&v)[L_CACHE]) {
#pragma unroll
  for (int i = 0; i < L_CACHE / 4; ++i) {
    const int c = (i * TPB + threadIdx.x) * 4;
    if (vec_in && c + 4 <= n_neg) {
      const float4 t = *reinterpret_cast<const float4*>(nrow + c);
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[4 * i + j] = c + j < n_neg ? nrow[c + j] : 0.f;
    }
  }
}

template <int KIND, int OUT, int TPB>
BESS_D void loss_row(const LossArgs& a, int r, float (&v)[L_CACHE], float* red) {
  constexpr int L_THREADS = TPB;
  const float margin = a.margin, adv_scale = a.adv_scale, loss_scale = a.loss_scale,
              ce_shift = a.ce_shift;
  const int adversarial = a.adversarial, n_neg = a.n_neg, weight_n = a.weight_n;
  const float* pos = a.pos;
  const float* weight = a.weight;
  float* neg = a.neg;
  float* row_loss = a.row_loss;
  float* d_pos = a.d_pos;
  void* g_hi = a.g_hi;
  void* g_lo = a.g_lo;
  const int64_t ld = a.ld, ld_g = a.ld_g;
  float* nrow = neg + (int64_t)r * ld;
  const int64_t grow = (int64_t)r * ld_g;
  const float w = weight_n == 1 ? weight[0] : weight[r];
  const float p = pos[r];
  const bool vec_in = (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(neg) & 15) == 0;
  constexpr int kGradBytes = (OUT == GRAD_BF16 || OUT == GRAD_F16 || OUT == GRAD_F16X3) ? 2 : 4;
  const bool vec_out = (ld_g & 3) == 0 && (reinterpret_cast<uintptr_t>(g_hi) & 15) == 0 &&
                       ((OUT != GRAD_TF32 && OUT != GRAD_F16X3) ||
                        (reinterpret_cast<uintptr_t>(g_lo) & 15) == 0);
  const float gscale = OUT == GRAD_F16X3 ? __ldg(a.g_scale) : 1.f;
  (void)kGradBytes;

  (void)vec_in;
  constexpr int kCached = L_CACHE * L_THREADS;  // columns beyond this are re-read

  if (KIND == BESS_LOSS_SOFTMAX_CE) {
    // scores adjusted in place by log(E-1) - log(N) (loss.py:233-237), then
    // cross entropy of [pos, neg...] against class 0
    float mx = p;
#pragma unroll
    for (int i = 0; i < L_CACHE / 4; ++i) {
      const int c = (i * L_THREADS + threadIdx.x) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c + j < n_neg) {
          v[4 * i + j] += ce_shift;
          nrow[c + j] = v[4 * i + j];
          mx = fmaxf(mx, v[4 * i + j]);
        }
    }
    for (int c = kCached + threadIdx.x; c < n_neg; c += L_THREADS) {
      const float t = nrow[c] + ce_shift;
      nrow[c] = t;
      mx = fmaxf(mx, t);
    }
    mx = block_max<L_THREADS>(mx, red);
    float se = 0.f;
#pragma unroll
    for (int i = 0; i < L_CACHE / 4; ++i) {
      const int c = (i * L_THREADS + threadIdx.x) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c + j < n_neg) se += expf(v[4 * i + j] - mx);
    }
    for (int c = kCached + threadIdx.x; c < n_neg; c += L_THREADS) se += expf(nrow[c] - mx);
    se = block_sum<L_THREADS>(se, red) + expf(p - mx);
    const float lse = mx + logf(se);
#pragma unroll
    for (int i = 0; i < L_CACHE / 4; ++i) {
      const int c = (i * L_THREADS + threadIdx.x) * 4;
      if (c < n_neg) {
        float g[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) g[j] = loss_scale * w * expf(v[4 * i + j] - lse);
        store_grad4<OUT>(g_hi, g_lo, grow + c, g, min(4, n_neg - c), vec_out, gscale);
      }
    }
    for (int c = kCached + threadIdx.x; c < n_neg; c += L_THREADS) {
      float g[4] = {loss_scale * w * expf(nrow[c] - lse), 0.f, 0.f, 0.f};
      store_grad4<OUT>(g_hi, g_lo, grow + c, g, 1, false, gscale);
    }
    if (threadIdx.x == 0) {
      row_loss[r] = loss_scale * w * (lse - p);
      d_pos[r] = loss_scale * w * (expf(p - lse) - 1.f);
    }
    return;
  }

  // negative weights: softmax(adv_scale * neg) (detached) or 1/N (loss.py:28-51)
  float mx = 0.f, inv_se = 1.f / (float)n_neg;
  float ev[L_CACHE];  // adversarial: exp(adv_scale * score - max) of the cached columns
  if (adversarial) {
    mx = -CUDART_INF_F;
#pragma unroll
    for (int i = 0; i < L_CACHE / 4; ++i) {
      const int c = (i * L_THREADS + threadIdx.x) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c + j < n_neg) mx = fmaxf(mx, adv_scale * v[4 * i + j]);
    }
    for (int c = kCached + threadIdx.x; c < n_neg; c += L_THREADS) mx = fmaxf(mx, adv_scale * nrow[c]);
    mx = block_max<L_THREADS>(mx, red);
    float se = 0.f;
#pragma unroll
    for (int i = 0; i < L_CACHE / 4; ++i) {
      const int c = (i * L_THREADS + threadIdx.x) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // the un-normalised softmax weight is kept for the second pass (one ex2 less per score)
        ev[4 * i + j] = c + j < n_neg ? __expf(adv_scale * v[4 * i + j] - mx) : 0.f;
        se += ev[4 * i + j];
      }
    }
    for (int c = kCached + threadIdx.x; c < n_neg; c += L_THREADS) se += __expf(adv_scale * nrow[c] - mx);
    se = block_sum<L_THREADS>(se, red);
    inv_se = 1.f / se;
  }
  float part = 0.f, dp = 0.f;
  // Per-element math uses the SFU intrinsics (ex2 / lg2 / rcp.approx, ~2 ulp): the kernel was
  // issue-bound on the libm expansions (ncu: 69 % issue active, DRAM 24 %).  Every term enters
  // the row loss / gradient through a weighted SUM with weights summing to 1, so the absolute
  // error stays ~1e-7 of a row loss of order 1..10 — far inside the 1e-5 bar (tests).
  // gradient of a column with score s (e = its un-normalised softmax weight); accumulates the loss
  auto one = [&](float s, float e) -> float {
    const float wj = adversarial ? e * inv_se : inv_se;
    if (KIND == BESS_LOSS_LOGSIGMOID) {
      // x = -s - m;  t = exp(-|x|);  logsigmoid(x) = min(x, 0) - log(1 + t);
      // sigmoid(s + m) = sigmoid(-x) = (x >= 0 ? t : 1) / (1 + t)
      const float x = -s - margin;
      const float t = __expf(-fabsf(x));
      part += wj * (fminf(x, 0.f) - __logf(1.f + t));
      // d/ds [-0.5 w wj logsigmoid(-s-m)] = 0.5 w wj sigmoid(s+m)
      return loss_scale * 0.5f * w * wj * __fdividef(x >= 0.f ? t : 1.f, 1.f + t);
    }
    // margin ranking: relu(s - pos + m)
    const float a = s - p + margin;
    part += wj * fmaxf(a, 0.f);
    const float g = loss_scale * w * wj * (a > 0.f ? 1.f : 0.f);
    dp -= g;
    return g;
  };
#pragma unroll
  for (int i = 0; i < L_CACHE / 4; ++i) {
    const int c = (i * L_THREADS + threadIdx.x) * 4;
    if (c < n_neg) {
      float g[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        g[j] = c + j < n_neg ? one(v[4 * i + j], adversarial ? ev[4 * i + j] : 0.f) : 0.f;
      store_grad4<OUT>(g_hi, g_lo, grow + c, g, min(4, n_neg - c), vec_out, gscale);
    }
  }
  for (int c = kCached + threadIdx.x; c < n_neg; c += L_THREADS) {
    const float sc = nrow[c];
    float g[4] = {one(sc, adversarial ? __expf(adv_scale * sc - mx) : 0.f), 0.f, 0.f, 0.f};
    store_grad4<OUT>(g_hi, g_lo, grow + c, g, 1, false, gscale);
  }
  part = block_sum<L_THREADS>(part, red);
  if (KIND == BESS_LOSS_MARGIN_RANKING) dp = block_sum<L_THREADS>(dp, red);
  if (threadIdx.x == 0) {
    if (KIND == BESS_LOSS_LOGSIGMOID) {
      row_loss[r] = loss_scale * (-0.5f) * w * (log_sigmoid(p + margin) + part);
      d_pos[r] = loss_scale * (-0.5f) * w * sigmoidf_(-(p + margin));
    } else {
      row_loss[r] = loss_scale * w * part;
      d_pos[r] = dp;
    }
  }
}

// ---------------------------------------------------------------------------
// Fast path of the LogSigmoid loss for FULL rows (n_neg == L_CACHE * TPB, 16-byte aligned in
// and out): the cfg-2 shape.  ncu on the generic path above showed ~63 executed instructions
// per score (ALU pipe 48 %, 16 warps / SM): bounds predicates on every element, three
// conversions per element for the fp16 pair, libm-style expansions.  Here: no predicates, all
// per-row constants folded into one multiplier, one ex2 for the softmax weight (reused), one
// ex2 + one lg2 + one rcp for logsigmoid / sigmoid (approx SFU ops, absolute error ~1e-7 of
// terms that enter weighted sums with weights summing to 1), and the fp16 (hi, lo) pair from
// integer rounding of the fp32 bits (hi = x rounded to 11 significant bits, exactly
// representable in fp16 unless it is subnormal there, i.e. < 2^-27 of the operand's largest
// element) + two packed conversions per pair of elements.
// ---------------------------------------------------------------------------
BESS_D float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
BESS_D float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
BESS_D float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
BESS_D float round11(float x) {  // round to 11 significant bits (half away from zero)
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

template <int OUT>
BESS_D void store_grad4_fast(void* g_hi, void* g_lo, int64_t at, const float (&g)[4]) {
  if (OUT == GRAD_F16X3) {  // g already carries the operand scale
    float h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = round11(g[j]);
    const __half2 a = __floats2half2_rn(h[0], h[1]), b = __floats2half2_rn(h[2], h[3]);
    const __half2 c = __floats2half2_rn(g[0] - h[0], g[1] - h[1]),
                  d = __floats2half2_rn(g[2] - h[2], g[3] - h[3]);
    uint2 uh, ul;
    uh.x = *reinterpret_cast<const uint32_t*>(&a); uh.y = *reinterpret_cast<const uint32_t*>(&b);
    ul.x = *reinterpret_cast<const uint32_t*>(&c); ul.y = *reinterpret_cast<const uint32_t*>(&d);
    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(g_hi) + at) = uh;
    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(g_lo) + at) = ul;
  } else {
    store_grad4<OUT>(g_hi, g_lo, at, g, 4, true, 1.f);
  }
}

template <int OUT, int TPB>
BESS_D void loss_row_logsigmoid_full(const LossArgs& a, int r, float (&v)[L_CACHE], float* red) {
  constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
  const float w = a.weight_n == 1 ? a.weight[0] : a.weight[r];
  const float p = a.pos[r];
  const float gscale = OUT == GRAD_F16X3 ? __ldg(a.g_scale) : 1.f;
  const int64_t grow = (int64_t)r * a.ld_g;
  float inv_se = 1.f / (float)a.n_neg;
  float ev[L_CACHE];
  if (a.adversarial) {
    float mx = -CUDART_INF_F;
#pragma unroll
    for (int i = 0; i < L_CACHE; ++i) mx = fmaxf(mx, a.adv_scale * v[i]);
    mx = block_max<TPB>(mx, red);
    const float sa = a.adv_scale * LOG2E, mb = mx * LOG2E;
    float se = 0.f;
#pragma unroll
    for (int i = 0; i < L_CACHE; ++i) {
      ev[i] = ex2_approx(fmaf(v[i], sa, -mb));
      se += ev[i];
    }
    se = block_sum<TPB>(se, red);
    inv_se = 1.f / se;
  }
  // dL/dscore_j = loss_scale * 0.5 * w * wj * sigmoid(s_j + m), wj = ev_j * inv_se (or 1 / N)
  const float gc = a.loss_scale * 0.5f * w * inv_se * gscale;
  const float nm = -a.margin;
  float part = 0.f;
#pragma unroll
  for (int i = 0; i < L_CACHE / 4; ++i) {
    float g[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float x = nm - v[4 * i + j];               // -s - m
      const float t = ex2_approx(-fabsf(x) * LOG2E);   // exp(-|x|)
      const float den = 1.f + t;
      const float ls = fmaf(-LN2, lg2_approx(den), fminf(x, 0.f));  // logsigmoid(x)
      const float sg = (x >= 0.f ? t : 1.f) * rcp_approx(den);      // sigmoid(s + m)
      if (a.adversarial) {
        part = fmaf(ev[4 * i + j], ls, part);
        g[j] = gc * ev[4 * i + j] * sg;
      } else {
        part += ls;
        g[j] = gc * sg;
      }
    }
    store_grad4_fast<OUT>(a.g_hi, a.g_lo, grow + (int64_t)(i * TPB + threadIdx.x) * 4, g);
  }
  part = block_sum<TPB>(part, red) * inv_se;
  if (threadIdx.x == 0) {
    a.row_loss[r] = a.loss_scale * (-0.5f) * w * (log_sigmoid(p + a.margin) + part);
    a.d_pos[r] = a.loss_scale * (-0.5f) * w * sigmoidf_(-(p + a.margin));
  }
}

// Persistent CTAs over rows; the next row's loads are issued before the
// current row is processed so every CTA always has one row in flight.
template <int KIND, int OUT, int TPB>
__global__ void __launch_bounds__(TPB) loss_kernel(const LossArgs a) {
  __shared__ float red[TPB / 32 + 1];
  const bool vec_in = (a.ld & 3) == 0 && (reinterpret_cast<uintptr_t>(a.neg) & 15) == 0;
  int r = blockIdx.x;
  if (r >= a.n) return;
  float v[L_CACHE], nxt[L_CACHE];
  loss_load_row<TPB>(a.neg + (int64_t)r * a.ld, a.n_neg, vec_in, nxt);
  for (; r < a.n; r += gridDim.x) {
#pragma unroll
    for (int i = 0; i < L_CACHE; ++i) v[i] = nxt[i];
    const int rn = r + gridDim.x;
    if (rn < a.n) loss_load_row<TPB>(a.neg + (int64_t)rn * a.ld, a.n_neg, vec_in, nxt);
    if (KIND == BESS_LOSS_LOGSIGMOID && a.full_rows) loss_row_logsigmoid_full<OUT, TPB>(a, r, v, red);
    else loss_row<KIND, OUT, TPB>(a, r, v, red);
  }
}

// single-CTA fixed-order sum
__global__ void __launch_bounds__(1024) sum_kernel(const float* x, int n, float* out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 1024) s += x[i];
  s = block_sum<1024>(s, red);
  if (threadIdx.x == 0) out[0] = s;
}

// ranks_from_scores (metric.py:129-183)
__global__ void __launch_bounds__(L_THREADS) rank_kernel(const float* pos, const float* neg, int n,
                                                          int n_neg, int64_t ld, int mode,
                                                          int worst_inf, float* rank) {
  __shared__ float red[L_THREADS / 32];
  const int r = blockIdx.x;
  float p = pos[r];
  if (isnan(p)) p = -CUDART_INF_F;  // pos_score.nan_to_num_(-inf)
  const float* row = neg + (int64_t)r * ld;
  float gt = 0.f, ge = 0.f;
  for (int c = threadIdx.x; c < n_neg; c += L_THREADS) {
    const float v = row[c];
    gt += v > p ? 1.f : 0.f;
    ge += v >= p ? 1.f : 0.f;
  }
  gt = block_sum<L_THREADS>(gt, red);
  ge = block_sum<L_THREADS>(ge, red);
  if (threadIdx.x == 0) {
    float better;
    bool worst;
    if (mode == 0) { better = gt; worst = gt == (float)n_neg; }
    else if (mode == 1) { better = ge; worst = ge == (float)n_neg; }
    else { better = 0.5f * (gt + ge); worst = gt == (float)n_neg || ge == (float)n_neg; }
    rank[r] = (worst_inf && worst) ? CUDART_INF_F : 1.f + better;
  }
}

__global__ void fill_f32_kernel(float* p, int64_t n, float v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void fill_i32_kernel(int32_t* p, int64_t n, int32_t v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
template <typename T>
__global__ void cast_kernel(const float* src, T* dst, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = Elem<T>::from_f(src[i]);
}

// running top-k merge (bess.py:807-814): one warp per query keeps its sorted
// best list in shared memory and inserts the window's scores one lane-batch at
// a time.  Ordering: score descending; ties keep the entry that came first
// (current list before the new window, lower window column first), which is
// the stable order of a descending sort of cat([window, current]) restricted
// to ... see DESIGN.md (tie order of torch.topk is unspecified).
constexpr int TK_MAXK = 64;
__global__ void __launch_bounds__(128) topk_merge_kernel(const float* win_score, int64_t ld,
                                                          int n_query, int n_win,
                                                          const int32_t* win_ids, int64_t ld_ids,
                                                          int win_id0, float* best_score,
                                                          int32_t* best_id, int k) {
  __shared__ float s_sc[4][TK_MAXK];
  __shared__ int32_t s_id[4][TK_MAXK];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int q = blockIdx.x * 4 + w;
  if (q >= n_query) return;
  float* sc = s_sc[w];
  int32_t* id = s_id[w];
  for (int i = lane; i < k; i += 32) {
    sc[i] = best_score[(int64_t)q * k + i];
    id[i] = best_id[(int64_t)q * k + i];
  }
  __syncwarp();
  const float* row = win_score + (int64_t)q * ld;
  for (int c0 = 0; c0 < n_win; c0 += 32) {
    const int c = c0 + lane;
    float v = c < n_win ? row[c] : -CUDART_INF_F;
    int32_t vid = 0;
    if (c < n_win) vid = win_ids != nullptr ? win_ids[(ld_ids == 0 ? 0 : (int64_t)q * ld_ids) + c] : win_id0 + c;
    // candidates that can enter the list (strictly better than the current worst)
    unsigned pending = __ballot_sync(0xffffffffu, c < n_win && v > sc[k - 1]);
    while (pending) {
      const int src = __ffs(pending) - 1;
      pending &= pending - 1;
      const float nv = __shfl_sync(0xffffffffu, v, src);
      const int32_t nid = __shfl_sync(0xffffffffu, vid, src);
      if (!(nv > sc[k - 1])) continue;  // list may have tightened since the ballot
      // insertion position: first i with sc[i] < nv (ties stay behind existing entries)
      int posn = k;
      for (int i0 = 0; i0 < k; i0 += 32) {
        const int i = i0 + lane;
        const unsigned m = __ballot_sync(0xffffffffu, i < k && sc[i] < nv);
        if (m) { posn = i0 + __ffs(m) - 1; break; }
      }
      // shift [posn, k-1) down by one (from the back), lanes cooperate in chunks
      for (int i0 = ((k - 1 - posn + 31) / 32 - 1) * 32; i0 >= 0; i0 -= 32) {
        const int i = posn + i0 + lane;  // element to move to i+1
        float t = 0.f; int32_t ti = 0;
        const bool act = i < k - 1;
        if (act) { t = sc[i]; ti = id[i]; }
        __syncwarp();
        if (act) { sc[i + 1] = t; id[i + 1] = ti; }
        __syncwarp();
      }
      if (lane == 0) { sc[posn] = nv; id[posn] = nid; }
      __syncwarp();
    }
  }
  for (int i = lane; i < k; i += 32) {
    best_score[(int64_t)q * k + i] = sc[i];
    best_id[(int64_t)q * k + i] = id[i];
  }
}

// Same merge for k <= 32 (the usual k + 1 = 11): the best list lives in
// registers (lane i = entry i), the window row is streamed with 128-bit loads,
// TKR_U of them in flight per lane, and a chunk is skipped with one vote when
// none of its TKR_U * 128 scores beats the current worst entry — the common case
// once the list has warmed up, which makes the kernel a plain streaming read of
// the [n_query, n_win] scores.  Entries are inserted in column order, so the
// result (ties included) is identical to topk_merge_kernel.
constexpr int TKR_WARPS = 8, TKR_U = 8;
__global__ void __launch_bounds__(TKR_WARPS * 32) topk_merge_reg_kernel(
    const float* __restrict__ win_score, int64_t ld, int n_query, int n_win,
    const int32_t* __restrict__ win_ids, int64_t ld_ids, int win_id0, float* best_score,
    int32_t* best_id, int k) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * TKR_WARPS + (threadIdx.x >> 5);
  if (q >= n_query) return;  // warp-uniform
  float sc = lane < k ? best_score[(int64_t)q * k + lane] : -CUDART_INF_F;
  int32_t id = lane < k ? best_id[(int64_t)q * k + lane] : 0;
  float thr = __shfl_sync(FULL, sc, k - 1);
  const float* row = win_score + (int64_t)q * ld;
  const int32_t* idrow =
      win_ids != nullptr ? win_ids + (ld_ids == 0 ? (int64_t)0 : (int64_t)q * ld_ids) : nullptr;

  // nv, c are warp-uniform
  auto insert = [&](float nv, int c) {
    if (!(nv > thr)) return;  // the list may have tightened since the vote
    // entries that stay in front: the sorted prefix with score >= nv (ties keep the older entry)
    const int pos = __popc(__ballot_sync(FULL, lane < k && sc >= nv));
    const int32_t nid = idrow != nullptr ? __ldg(idrow + c) : win_id0 + c;
    const float up_s = __shfl_up_sync(FULL, sc, 1);
    const int32_t up_i = __shfl_up_sync(FULL, id, 1);
    if (lane > pos) { sc = up_s; id = up_i; }
    else if (lane == pos) { sc = nv; id = nid; }
    thr = __shfl_sync(FULL, sc, k - 1);
  };

  int c0 = 0;
  if ((reinterpret_cast<uintptr_t>(row) & 15) == 0) {
    constexpr int CH = TKR_U * 128;
    for (; c0 + CH <= n_win; c0 += CH) {
      float v[TKR_U][4];
#pragma unroll
      for (int u = 0; u < TKR_U; ++u) {
        const uint4 r = ld_stream(reinterpret_cast<const uint4*>(row + c0 + u * 128) + lane);
        v[u][0] = __uint_as_float(r.x); v[u][1] = __uint_as_float(r.y);
        v[u][2] = __uint_as_float(r.z); v[u][3] = __uint_as_float(r.w);
      }
      float m = -CUDART_INF_F;
#pragma unroll
      for (int u = 0; u < TKR_U; ++u)
        m = fmaxf(m, fmaxf(fmaxf(v[u][0], v[u][1]), fmaxf(v[u][2], v[u][3])));
      if (!__any_sync(FULL, m > thr)) continue;
#pragma unroll
      for (int u = 0; u < TKR_U; ++u) {
        unsigned mj[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) mj[j] = __ballot_sync(FULL, v[u][j] > thr);
        unsigned any = mj[0] | mj[1] | mj[2] | mj[3];
        while (any) {  // lanes in order, then the lane's four columns in order = column order
          const int src = __ffs(any) - 1;
          any &= any - 1;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if ((mj[j] >> src) & 1u)
              insert(__shfl_sync(FULL, v[u][j], src), c0 + u * 128 + src * 4 + j);
        }
      }
    }
  }
  for (; c0 < n_win; c0 += 32) {  // tail of the window / rows that are not 16-byte aligned
    const int c = c0 + lane;
    const float v = c < n_win ? row[c] : -CUDART_INF_F;
    unsigned pending = __ballot_sync(FULL, v > thr);
    while (pending) {
      const int src = __ffs(pending) - 1;
      pending &= pending - 1;
      insert(__shfl_sync(FULL, v, src), c0 + src);
    }
  }
  if (lane < k) {
    best_score[(int64_t)q * k + lane] = sc;
    best_id[(int64_t)q * k + lane] = id;
  }
}

// Final merge of the per-shard best lists (bess.py:866-891): discard padding
// rows of each scoring shard (score += bad where idx >= shard_counts[j]), map
// local ids to global ids and select the k best of the n * kb entries.  Order:
// score descending, ties by position in the [shard j, entry e] concatenation —
// the stable order of the reference's flatten + topk.  One thread per query.
__global__ void topk_finalize_kernel(const float* __restrict__ score, const int32_t* __restrict__ idx,
                                     int n, int S, int kb, const int32_t* __restrict__ shard_counts,
                                     const int32_t* __restrict__ shard_idx_to_entity, int Es, int k,
                                     float bad, float* __restrict__ out_score,
                                     int32_t* __restrict__ out_id) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= S) return;
  const int total = n * kb;
  float last_sc = CUDART_INF_F;
  int last_pos = -1;
  for (int o = 0; o < k; ++o) {
    float best = -CUDART_INF_F;
    int best_pos = -1;
    for (int pos = 0; pos < total; ++pos) {
      const int j = pos / kb, e = pos - j * kb;
      const int64_t at = ((int64_t)j * S + q) * kb + e;
      const int32_t id = idx[at];
      const float sc = score[at] + (id >= shard_counts[j] ? bad : 0.f);
      const bool eligible = sc < last_sc || (sc == last_sc && pos > last_pos);
      if (eligible && sc > best) { best = sc; best_pos = pos; }
    }
    if (best_pos < 0) {  // fewer than k candidates in total
      out_score[(int64_t)q * k + o] = -CUDART_INF_F;
      out_id[(int64_t)q * k + o] = -1;
      continue;
    }
    const int j = best_pos / kb, e = best_pos - j * kb;
    const int32_t id = idx[((int64_t)j * S + q) * kb + e];
    out_score[(int64_t)q * k + o] = best;
    out_id[(int64_t)q * k + o] = shard_idx_to_entity[(int64_t)j * Es + min(id, Es - 1)];
    last_sc = best;
    last_pos = best_pos;
  }
}

// ---------------------------------------------------------------------------
// AllScores post-processing (pipeline.py:227-298): the reference re-orders the
// [query, n_shard * n_step * window] block scores into global-entity order on the
// host (np.unique over the scored ids), drops padding queries, writes -inf on
// non-candidate / filtered completions and reads / restores the ground-truth
// scores with fancy indexing.  Here these are three small device kernels.
//   select: out[i, e] = col_idx[e] >= 0 ? src[row_idx[i], col_idx[e]] : fill
//   pairs_get / pairs_set: v[t] = mat[rows[t], cols[t]] (rows == NULL -> t)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) select_scores_kernel(const float* __restrict__ src, int64_t ld_src,
                                                            const int32_t* __restrict__ row_idx,
                                                            const int32_t* __restrict__ col_idx,
                                                            int n_cols, float fill,
                                                            float* __restrict__ out, int64_t ld_out) {
  const int i = blockIdx.y;
  const float* srow = src + (int64_t)(row_idx != nullptr ? row_idx[i] : i) * ld_src;
  float* orow = out + (int64_t)i * ld_out;
  for (int e = blockIdx.x * 256 + threadIdx.x; e < n_cols; e += gridDim.x * 256) {
    const int32_t c = __ldg(col_idx + e);
    orow[e] = c >= 0 ? __ldg(srow + c) : fill;
  }
}
__global__ void pairs_get_kernel(const float* __restrict__ mat, int64_t ld,
                                 const int32_t* __restrict__ rows, const int32_t* __restrict__ cols,
                                 int n, float* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  out[t] = mat[(int64_t)(rows != nullptr ? rows[t] : t) * ld + cols[t]];
}
__global__ void pairs_set_kernel(float* __restrict__ mat, int64_t ld, const int32_t* __restrict__ rows,
                                 const int32_t* __restrict__ cols, int n,
                                 const float* __restrict__ values, float value) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  mat[(int64_t)(rows != nullptr ? rows[t] : t) * ld + cols[t]] = values != nullptr ? values[t] : value;
}

}  // namespace bess

using namespace bess;

extern "C" int bess_topk_finalize(const float* score, const int32_t* idx, int n_shard, int n_query,
                                  int kb, const int32_t* shard_counts,
                                  const int32_t* shard_idx_to_entity, int max_entity_per_shard, int k,
                                  float bad_score, float* out_score, int32_t* out_id, void* stream) {
  if (n_query == 0) return BESS_OK;
  BESS_CHECK_ARG(k >= 1 && k <= n_shard * kb, "k=%d exceeds the %d merged entries", k, n_shard * kb);
  topk_finalize_kernel<<<ceil_div(n_query, 128), 128, 0, (cudaStream_t)stream>>>(
      score, idx, n_shard, n_query, kb, shard_counts, shard_idx_to_entity, max_entity_per_shard, k,
      bad_score, out_score, out_id);
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

extern "C" int bess_mask_add(float* score, int n_row, int n_col, int64_t ld, const uint8_t* mask,
                             int64_t ld_mask, int mask_rows, int flag, float value, void* stream) {
  if (n_row == 0 || n_col == 0) return BESS_OK;
  BESS_CHECK_ARG(mask_rows == 1 || mask_rows == n_row, "mask rows %d vs %d", mask_rows, n_row);
  mask_add_kernel<<<n_row, L_THREADS, 0, (cudaStream_t)stream>>>(score, n_row, n_col, ld, mask,
                                                                 ld_mask, mask_rows, flag, value);
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

extern "C" int bess_mask_diag(float* score, int n_row, int64_t ld, int step, int half_group,
                              int group, float value, void* stream) {
  if (n_row == 0) return BESS_OK;
  mask_diag_kernel<<<ceil_div(n_row, 256), 256, 0, (cudaStream_t)stream>>>(
      score, n_row, ld, step, half_group, group, value, (int)ld);
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

template <int KIND, int OUT, int TPB>
static void launch_loss_tpb(const LossArgs& a0, cudaStream_t st) {
  // enough CTAs to keep ~48 KB of row loads in flight per SM, but persistent
  const int per_sm = TPB == 64 ? 16 : (TPB == 128 ? 8 : 4);
  LossArgs a = a0;
  const int out_elem = (OUT == GRAD_BF16 || OUT == GRAD_F16 || OUT == GRAD_F16X3) ? 2 : 4;
  const bool lo_used = OUT == GRAD_TF32 || OUT == GRAD_F16X3;
  a.full_rows = a.n_neg == L_CACHE * TPB && (a.ld & 3) == 0 && (a.ld_g & 3) == 0 &&
                (reinterpret_cast<uintptr_t>(a.neg) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(a.g_hi) & (4 * out_elem - 1)) == 0 &&
                (!lo_used || (reinterpret_cast<uintptr_t>(a.g_lo) & (4 * out_elem - 1)) == 0) &&
                getenv("BESS_LOSS_GENERIC") == nullptr;
  const int grid = a.n < kNumSM * per_sm ? a.n : kNumSM * per_sm;
  loss_kernel<KIND, OUT, TPB><<<grid, TPB, 0, st>>>(a);
}

template <int KIND, int OUT>
static void launch_loss_kind(const LossArgs& a, cudaStream_t st) {
  if (a.n_neg <= 64 * L_CACHE) launch_loss_tpb<KIND, OUT, 64>(a, st);
  else if (a.n_neg <= 128 * L_CACHE) launch_loss_tpb<KIND, OUT, 128>(a, st);
  else launch_loss_tpb<KIND, OUT, 256>(a, st);
}

template <int OUT>
static int launch_loss(int kind, float margin, int adversarial, float adv_scale, float loss_scale,
                       int64_t n_entity, const float* pos, float* neg, int n, int n_neg, int64_t ld,
                       const float* weight, int weight_n, float* row_loss, float* d_pos, void* g_hi,
                       void* g_lo, int64_t ld_g, cudaStream_t st, const float* g_scale = nullptr) {
  LossArgs a;
  a.g_scale = g_scale;
  a.margin = margin; a.adversarial = adversarial; a.adv_scale = adv_scale; a.loss_scale = loss_scale;
  a.ce_shift = 0.f; a.pos = pos; a.neg = neg; a.n = n; a.n_neg = n_neg; a.ld = ld; a.weight = weight;
  a.weight_n = weight_n; a.row_loss = row_loss; a.d_pos = d_pos; a.g_hi = g_hi; a.g_lo = g_lo;
  a.ld_g = ld_g;
  switch (kind) {
    case BESS_LOSS_LOGSIGMOID: launch_loss_kind<BESS_LOSS_LOGSIGMOID, OUT>(a, st); break;
    case BESS_LOSS_MARGIN_RANKING: launch_loss_kind<BESS_LOSS_MARGIN_RANKING, OUT>(a, st); break;
    case BESS_LOSS_SOFTMAX_CE:
      a.adversarial = 0;
      a.ce_shift = (float)(log((double)(n_entity - 1)) - log((double)n_neg));
      launch_loss_kind<BESS_LOSS_SOFTMAX_CE, OUT>(a, st);
      break;
    default:
      bess_set_error("unknown loss kind %d", kind);
      return BESS_ERR_INVALID_ARG;
  }
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

extern "C" int bess_loss_fwd_bwd(int kind, float margin, int adversarial, float adv_scale,
                                 float loss_scale, int64_t n_entity, const float* pos, float* neg,
                                 int n, int n_neg, int64_t ld, const float* weight, int weight_n,
                                 float* row_loss, float* d_pos, float* d_neg, void* stream) {
  if (n == 0) return BESS_OK;
  BESS_CHECK_ARG(n_neg > 0, "loss needs at least one negative");
  BESS_CHECK_ARG(weight_n == 1 || weight_n == n, "triple_weight has %d entries, need 1 or %d", weight_n, n);
  return launch_loss<GRAD_F32>(kind, margin, adversarial, adv_scale, loss_scale, n_entity, pos, neg, n,
                               n_neg, ld, weight, weight_n, row_loss, d_pos, d_neg, nullptr, ld,
                               (cudaStream_t)stream);
}

extern "C" int bess_loss_fwd_bwd_operand(int kind, float margin, int adversarial, float adv_scale,
                                         float loss_scale, int64_t n_entity, const float* pos,
                                         float* neg, int n, int n_neg, int64_t ld,
                                         const float* weight, int weight_n, float* row_loss,
                                         float* d_pos, int grad_dtype, void* d_neg_hi, void* d_neg_lo,
                                         int64_t ld_grad, const float* grad_scale, void* stream) {
  if (n == 0) return BESS_OK;
  BESS_CHECK_ARG(n_neg > 0, "loss needs at least one negative");
  BESS_CHECK_ARG(weight_n == 1 || weight_n == n, "triple_weight has %d entries, need 1 or %d", weight_n, n);
  BESS_CHECK_ARG(d_neg_hi != nullptr &&
                     ((grad_dtype != BESS_F32 && grad_dtype != BESS_F16X3) || d_neg_lo != nullptr),
                 "bess_loss_fwd_bwd_operand: missing gradient output");
  BESS_CHECK_ARG(grad_dtype != BESS_F16X3 || grad_scale != nullptr,
                 "bess_loss_fwd_bwd_operand: BESS_F16X3 needs the gradient operand's scale");
  cudaStream_t st = (cudaStream_t)stream;
  switch (grad_dtype) {
    case BESS_F16X3:
      return launch_loss<GRAD_F16X3>(kind, margin, adversarial, adv_scale, loss_scale, n_entity, pos,
                                     neg, n, n_neg, ld, weight, weight_n, row_loss, d_pos, d_neg_hi,
                                     d_neg_lo, ld_grad, st, grad_scale);
    case BESS_F32:
      return launch_loss<GRAD_TF32>(kind, margin, adversarial, adv_scale, loss_scale, n_entity, pos, neg,
                                    n, n_neg, ld, weight, weight_n, row_loss, d_pos, d_neg_hi, d_neg_lo,
                                    ld_grad, st);
    case BESS_BF16:
      return launch_loss<GRAD_BF16>(kind, margin, adversarial, adv_scale, loss_scale, n_entity, pos, neg,
                                    n, n_neg, ld, weight, weight_n, row_loss, d_pos, d_neg_hi, nullptr,
                                    ld_grad, st);
    case BESS_F16:
      return launch_loss<GRAD_F16>(kind, margin, adversarial, adv_scale, loss_scale, n_entity, pos, neg,
                                   n, n_neg, ld, weight, weight_n, row_loss, d_pos, d_neg_hi, nullptr,
                                   ld_grad, st);
    default:
      bess_set_error("bess_loss_fwd_bwd_operand: unknown gradient dtype %d", grad_dtype);
      return BESS_ERR_INVALID_ARG;
  }
}

extern "C" int bess_sum_f32(const float* x, int n, float* out, void* stream) {
  sum_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(x, n, out);
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

extern "C" int bess_rank_from_scores(const float* pos, const float* neg, int n, int n_neg,
                                     int64_t ld, int mode, int worst_rank_infty, float* rank,
                                     void* stream) {
  if (n == 0) return BESS_OK;
  BESS_CHECK_ARG(mode >= 0 && mode <= 2, "rank mode %d", mode);
  rank_kernel<<<n, L_THREADS, 0, (cudaStream_t)stream>>>(pos, neg, n, n_neg, ld, mode,
                                                          worst_rank_infty, rank);
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

extern "C" int bess_fill_f32(float* p, int64_t n, float v, void* stream) {
  if (n == 0) return BESS_OK;
  fill_f32_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(p, n, v);
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}
extern "C" int bess_fill_i32(int32_t* p, int64_t n, int32_t v, void* stream) {
  if (n == 0) return BESS_OK;
  fill_i32_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(p, n, v);
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}
extern "C" int bess_cast_from_f32(const float* src, void* dst, int dtype, int64_t n, void* stream) {
  if (n == 0) return BESS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == BESS_F16) cast_kernel<__half><<<ceil_div(n, 256), 256, 0, st>>>(src, (__half*)dst, n);
  else if (dtype == BESS_BF16) cast_kernel<__nv_bfloat16><<<ceil_div(n, 256), 256, 0, st>>>(src, (__nv_bfloat16*)dst, n);
  else cast_kernel<float><<<ceil_div(n, 256), 256, 0, st>>>(src, (float*)dst, n);
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

extern "C" int bess_topk_merge(const float* win_score, int64_t ld, int n_query, int n_win,
                               const int32_t* win_ids, int64_t ld_ids, int win_id0,
                               float* best_score, int32_t* best_id, int k, void* stream) {
  if (n_query == 0 || n_win == 0) return BESS_OK;
  BESS_CHECK_ARG(k >= 1 && k <= TK_MAXK, "k=%d out of range (max %d)", k, TK_MAXK);
  // BESS_TOPK_MERGE_V1=1 forces the shared-memory list kernel (A/B switch)
  static const bool force_v1 = [] { const char* e = getenv("BESS_TOPK_MERGE_V1"); return e && e[0] == '1'; }();
  if (k <= 32 && !force_v1)
    topk_merge_reg_kernel<<<ceil_div(n_query, TKR_WARPS), TKR_WARPS * 32, 0, (cudaStream_t)stream>>>(
        win_score, ld, n_query, n_win, win_ids, ld_ids, win_id0, best_score, best_id, k);
  else
    topk_merge_kernel<<<ceil_div(n_query, 4), 128, 0, (cudaStream_t)stream>>>(
        win_score, ld, n_query, n_win, win_ids, ld_ids, win_id0, best_score, best_id, k);
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

extern "C" int bess_select_scores(const float* src, int64_t ld_src, const int32_t* row_idx, int n_rows,
                                  const int32_t* col_idx, int n_cols, float fill, float* out,
                                  int64_t ld_out, void* stream) {
  if (n_rows == 0 || n_cols == 0) return BESS_OK;
  BESS_CHECK_ARG(col_idx != nullptr, "bess_select_scores: column map required");
  BESS_CHECK_ARG(n_rows <= 65535, "bess_select_scores: at most 65535 rows per call (got %d)", n_rows);
  const int gx = ceil_div(n_cols, 256) < 64 ? ceil_div(n_cols, 256) : 64;
  select_scores_kernel<<<dim3(gx, n_rows), 256, 0, (cudaStream_t)stream>>>(src, ld_src, row_idx, col_idx,
                                                                           n_cols, fill, out, ld_out);
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}
extern "C" int bess_pairs_get(const float* mat, int64_t ld, const int32_t* rows, const int32_t* cols,
                              int n, float* out, void* stream) {
  if (n == 0) return BESS_OK;
  pairs_get_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(mat, ld, rows, cols, n, out);
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}
extern "C" int bess_pairs_set(float* mat, int64_t ld, const int32_t* rows, const int32_t* cols, int n,
                              const float* values, float value, void* stream) {
  if (n == 0) return BESS_OK;
  pairs_set_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(mat, ld, rows, cols, n, values,
                                                                        value);
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}
