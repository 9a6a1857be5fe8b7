// Exact re-rank of the top-k candidate lists.
//
// "MRR / Hits@k must match exactly" cannot be promised by comparing two different fp32
// summation orders: wherever two candidates are closer than the rounding noise of a
// D-term sum (~1e-6 relative) the order is decided by the order of the additions, and the
// reference's own order (a BLAS matmul on the host, or the IPU's) is not a property of the
// algorithm.  What CAN be made exact is the ranking under ONE fixed, documented fp32
// arithmetic: the tensor-core / tile scorers only SELECT the top-(k + 1 + margin) local
// candidates of every (query, scoring shard); this kernel then re-scores just those in a
// fixed summation order — coordinate 0, 1, ..., W-1, every product and sum rounded to fp32
// separately (no FMA contraction), IEEE sqrt — and re-sorts the list by (score descending,
// local id ascending).  The oracle restates the same arithmetic with one rounding per
// torch op (oracle/besskge_oracle.py: exact_scores), so ids, scores, ranks and MRR agree
// BIT FOR BIT, with no near-tie tolerance.  The selection stage can only lose a true top-k
// member if more than `margin` candidates lie within its ~1e-6 relative error of the k-th
// score.
//
// Families: TransE (L1 / L2), DistMult, ComplEx — arithmetic made of + - x sqrt |.| only
// (RotatE needs sin / cos, BoxE exp / tanh: not bit-reproducible across libms).
//   score_tails / score_heads formulas: scoring.py:335-356 (TransE), :815-840 (DistMult),
//   :918-946 (ComplEx, with utils.complex_multiplication :72-89).
#include <math_constants.h>

#include "common.cuh"

namespace bess {

template <typename T>
__global__ void __launch_bounds__(256) topk_exact_rescore_kernel(
    int family, int norm_p, int mode, const T* __restrict__ fixed, int64_t fixed_pitch,
    const T* __restrict__ rel_table, int64_t rel_pitch, const int32_t* __restrict__ rel_id,
    const T* __restrict__ table, int64_t table_pitch, int Es, int W,
    const int32_t* __restrict__ ids_in, const float* __restrict__ sc_in, int n_query, int kbm,
    int kb, float* __restrict__ sc_out, int32_t* __restrict__ id_out) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= n_query) return;  // warp-uniform
  int32_t id = lane < kbm ? ids_in[(int64_t)q * kbm + lane] : 0x7fffffff;
  float score = lane < kbm ? sc_in[(int64_t)q * kbm + lane] : -CUDART_INF_F;
  const bool live = lane < kbm && id >= 0 && id < Es;  // id == Es: the list's empty slots
  const T* f = fixed + (int64_t)q * fixed_pitch;
  const T* r = rel_table + (int64_t)__ldg(rel_id + q) * rel_pitch;
  const T* c = table + (int64_t)(live ? id : 0) * table_pitch;
  const bool tails = mode == BESS_MODE_TAILS;
  float acc = 0.f;
  if (family == BESS_DISTMULT) {
    // tails: sum_k (h_k r_k) c_k ; heads: sum_k (r_k t_k) c_k — the same expression
    for (int k = 0; k < W; ++k) {
      const float qk = __fmul_rn(ldf(f + k), ldf(r + k));
      acc = __fadd_rn(acc, __fmul_rn(qk, ldf(c + k)));
    }
    if (live) score = acc;
  } else if (family == BESS_COMPLEX) {
    const int e = W >> 1;
    for (int k = 0; k < W; ++k) {
      const int j = k < e ? k : k - e;
      const float f_re = ldf(f + j), f_im = ldf(f + e + j);
      const float r_re = ldf(r + j), r_im = ldf(r + e + j);
      float qk;
      if (tails) {  // q = h (x) r
        qk = k < e ? __fsub_rn(__fmul_rn(f_re, r_re), __fmul_rn(f_im, r_im))
                   : __fadd_rn(__fmul_rn(f_re, r_im), __fmul_rn(f_im, r_re));
      } else {  // q = conj(r) (x) t, conj(r) = (r_re, -r_im)
        const float n_im = -r_im;
        qk = k < e ? __fsub_rn(__fmul_rn(r_re, f_re), __fmul_rn(n_im, f_im))
                   : __fadd_rn(__fmul_rn(r_re, f_im), __fmul_rn(n_im, f_re));
      }
      acc = __fadd_rn(acc, __fmul_rn(qk, ldf(c + k)));
    }
    if (live) score = acc;
  } else {  // TransE: tails -||(h + r) - c||_p, heads -||(t - r) - c||_p
    for (int k = 0; k < W; ++k) {
      const float qk = tails ? __fadd_rn(ldf(f + k), ldf(r + k)) : __fsub_rn(ldf(f + k), ldf(r + k));
      const float d = __fsub_rn(qk, ldf(c + k));
      acc = __fadd_rn(acc, norm_p == 1 ? fabsf(d) : __fmul_rn(d, d));
    }
    if (live) score = -(norm_p == 1 ? acc : __fsqrt_rn(acc));
  }
  // position in (score descending, id ascending, lane ascending) order
  int pos = 0;
#pragma unroll 1
  for (int i = 0; i < 32; ++i) {
    const float s_i = __shfl_sync(FULL, score, i);
    const int32_t id_i = __shfl_sync(FULL, id, i);
    if (s_i > score || (s_i == score && (id_i < id || (id_i == id && i < lane)))) ++pos;
  }
  if (lane < kbm && pos < kb) {
    sc_out[(int64_t)q * kb + pos] = score;
    id_out[(int64_t)q * kb + pos] = id;
  }
}

}  // namespace bess

using namespace bess;

extern "C" int bess_topk_exact_supported(int family) {
  return family == BESS_TRANSE || family == BESS_DISTMULT || family == BESS_COMPLEX;
}

extern "C" int bess_topk_exact_rescore(const bess_score_cfg_t* cfg, int dtype, int mode,
                                       const void* fixed, int64_t fixed_pitch,
                                       const void* rel_table, int64_t rel_pitch,
                                       const int32_t* rel_id, const void* table,
                                       int64_t table_pitch, int n_table_rows, int row_elems,
                                       const int32_t* ids_in, const float* score_in, int n_query,
                                       int k_in, int k_out, float* score_out, int32_t* ids_out,
                                       void* stream) {
  BESS_CHECK_ARG(bess_topk_exact_supported(cfg->family),
                 "bess_topk_exact_rescore: family %d has no bit-reproducible arithmetic",
                 cfg->family);
  BESS_CHECK_ARG(k_in >= 1 && k_in <= 32 && k_out >= 1 && k_out <= k_in,
                 "bess_topk_exact_rescore: need 1 <= k_out <= k_in <= 32 (got %d, %d)", k_out, k_in);
  if (n_query == 0) return BESS_OK;
  const int blocks = ceil_div((int64_t)n_query * 32, 256);
  cudaStream_t st = (cudaStream_t)stream;
#define BESS_EXACT_LAUNCH(T)                                                                      \
  topk_exact_rescore_kernel<T><<<blocks, 256, 0, st>>>(                                           \
      cfg->family, cfg->norm_p, mode, (const T*)fixed, fixed_pitch, (const T*)rel_table, rel_pitch, \
      rel_id, (const T*)table, table_pitch, n_table_rows, row_elems, ids_in, score_in, n_query,   \
      k_in, k_out, score_out, ids_out)
  switch (dtype) {
    case BESS_F32: BESS_EXACT_LAUNCH(float); break;
    case BESS_F16: BESS_EXACT_LAUNCH(__half); break;
    case BESS_BF16: BESS_EXACT_LAUNCH(__nv_bfloat16); break;
    default: bess_set_error("unknown dtype %d", dtype); return BESS_ERR_INVALID_ARG;
  }
#undef BESS_EXACT_LAUNCH
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}
