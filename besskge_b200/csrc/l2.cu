// Norm-expanded L2 distance on the tensor cores (BASELINE north_star (3); the reference's
// `pea.distance_matrix(v1, v2, p=2)` = torch.cdist, scoring.py:195-197, whose own large-input path
// is the same expansion):
//     ||q - c||^2 = ||q||^2 + ||c||^2 - 2 q.c
// The q.c block comes from the tcgen05 GEMM (gemm_tc.cu); this file holds the small kernels
// around it.  Forward:  score = -sqrt(max(qn + cn - 2 dot, 0)).  Backward, with
// b = (dL/dscore) / dist (0 where dist == 0):
//     dL/dq_s = B C - rb_s q_s        rb_s = sum_c b_sc
//     dL/dc_c = B^T Q - cb_c c_c      cb_c = sum_s b_sc
// i.e. the two contractions of the DOT families plus a row scaling.  Every reduction here
// runs in a fixed order (deterministic).
#include "common.cuh"

namespace bess {

#define L2_DISPATCH_DTYPE(dtype, ...)                                  \
  switch (dtype) {                                                     \
    case BESS_F32: { using T = float; __VA_ARGS__; break; }            \
    case BESS_F16: { using T = __half; __VA_ARGS__; break; }           \
    case BESS_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }   \
    default: bess_set_error("unknown dtype %d", dtype); return BESS_ERR_INVALID_ARG; \
  }

// out[i] = sum_k row_i[k]^2 ; one warp per row, lanes stride the row, shuffle tree
template <typename T>
__global__ void __launch_bounds__(256) row_sqnorm_kernel(bess_rows_t rows, int n, int width,
                                                         float* __restrict__ out) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n) return;
  const T* r = static_cast<const T*>(rows.base) + src_row(rows, w) * rows.pitch;
  float a = 0.f;
  for (int k = threadIdx.x & 31; k < width; k += 32) {
    const float v = ldf(r + k);
    a = fmaf(v, v, a);
  }
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) out[w] = a;
}

// score[map(q), col0 + c] <- -sqrt(max(qn[q] + cn[c] - 2 * score, 0))   (in place over the dots)
__global__ void __launch_bounds__(256) l2_from_dot_kernel(float* score, bess_rowmap_t map, int64_t ld,
                                                          int col0, int n_q, int n_c,
                                                          const float* __restrict__ qn,
                                                          const float* __restrict__ cn) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= n_c) return;
  const float cc = cn[c];
  for (int q = blockIdx.y; q < n_q; q += gridDim.y) {
    float* p = score + (int64_t)map_row(map, q) * ld + col0 + c;
    const float d2 = fmaf(-2.f, *p, qn[q] + cc);
    *p = -sqrtf(fmaxf(d2, 0.f));
  }
}

// coef[q, c] = g / dist = -g / score (0 where score == 0) and row_sum[q] = sum_c coef[q, c].
// One CTA per query row.
__global__ void __launch_bounds__(256) l2_coef_kernel(const float* __restrict__ d_score,
                                                      const float* __restrict__ score,
                                                      bess_rowmap_t map, int64_t ld, int col0, int n_c,
                                                      float* __restrict__ coef, int64_t ld_coef,
                                                      float* __restrict__ row_sum) {
  __shared__ float red[8];
  const int q = blockIdx.x;
  const int64_t off = (int64_t)map_row(map, q) * ld + col0;
  float a = 0.f;
  for (int c = threadIdx.x; c < n_c; c += 256) {
    const float s = score[off + c];
    const float b = s != 0.f ? -d_score[off + c] / s : 0.f;
    coef[(int64_t)q * ld_coef + c] = b;
    a += b;
  }
  a = block_sum<256>(a, red);
  if (threadIdx.x == 0) row_sum[q] = a;
}

// column sums of coef [n_q, n_c]: stage 1 = per chunk of rows (thread = column, rows in order),
// stage 2 = chunks in order
constexpr int kColChunk = 128;
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ coef,
                                                             int64_t ld, int n_q, int n_c,
                                                             float* __restrict__ partial) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= n_c) return;
  const int q0 = blockIdx.y * kColChunk;
  const int q1 = min(n_q, q0 + kColChunk);
  float a = 0.f;
  for (int q = q0; q < q1; ++q) a += coef[(int64_t)q * ld + c];
  partial[(int64_t)blockIdx.y * n_c + c] = a;
}
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ partial,
                                                           int n_chunk, int n_c,
                                                           float* __restrict__ out) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= n_c) return;
  float a = 0.f;
  for (int i = 0; i < n_chunk; ++i) a += partial[(int64_t)i * n_c + c];
  out[c] = a;
}

// out_i[k] += scale * alpha[i] * src_i[k]; one warp per row; out rows fp32
template <typename T>
__global__ void __launch_bounds__(256) rows_axpy_kernel(const float* __restrict__ alpha, float scale,
                                                        bess_rows_t src, bess_rows_t out, int n,
                                                        int width) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n) return;
  const T* s = static_cast<const T*>(src.base) + src_row(src, w) * src.pitch;
  float* o = static_cast<float*>(const_cast<void*>(out.base)) + src_row(out, w) * out.pitch;
  const float a = scale * alpha[w];
  for (int k = threadIdx.x & 31; k < width; k += 32) o[k] = fmaf(a, ldf(s + k), o[k]);
}

}  // namespace bess

using namespace bess;

static inline dim3 warp_grid_(int n_rows) { return dim3(ceil_div((int64_t)n_rows * 32, 256)); }

extern "C" int bess_row_sqnorm(int dtype, bess_rows_t rows, int n, int width, float* out,
                               void* stream) {
  if (n == 0) return BESS_OK;
  L2_DISPATCH_DTYPE(dtype, row_sqnorm_kernel<T><<<warp_grid_(n), 256, 0, (cudaStream_t)stream>>>(
                               rows, n, width, out));
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

extern "C" int bess_l2_from_dot(float* score, bess_rowmap_t score_map, int64_t ld, int col0, int n_query,
                                int n_cand, const float* q_sqnorm, const float* c_sqnorm,
                                void* stream) {
  if (n_query == 0 || n_cand == 0) return BESS_OK;
  const int gy = n_query < 4 * kNumSM ? n_query : 4 * kNumSM;
  l2_from_dot_kernel<<<dim3(ceil_div(n_cand, 256), gy), 256, 0, (cudaStream_t)stream>>>(
      score, score_map, ld, col0, n_query, n_cand, q_sqnorm, c_sqnorm);
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

extern "C" int64_t bess_l2_coef_workspace(int n_query, int n_cand) {
  return (int64_t)ceil_div(n_query, kColChunk) * n_cand * sizeof(float);
}

extern "C" int bess_l2_coef(const float* d_score, const float* score, bess_rowmap_t score_map,
                            int64_t ld, int col0, int n_query, int n_cand, float* coef,
                            int64_t ld_coef, float* row_sum, float* col_sum, void* workspace,
                            void* stream) {
  if (n_query == 0 || n_cand == 0) return BESS_OK;
  BESS_CHECK_ARG(workspace != nullptr, "bess_l2_coef: workspace required");
  cudaStream_t st = (cudaStream_t)stream;
  l2_coef_kernel<<<n_query, 256, 0, st>>>(d_score, score, score_map, ld, col0, n_cand, coef, ld_coef,
                                          row_sum);
  BESS_CHECK_LAUNCH();
  const int n_chunk = ceil_div(n_query, kColChunk);
  float* partial = static_cast<float*>(workspace);
  colsum_partial_kernel<<<dim3(ceil_div(n_cand, 256), n_chunk), 256, 0, st>>>(coef, ld_coef, n_query,
                                                                            n_cand, partial);
  BESS_CHECK_LAUNCH();
  colsum_final_kernel<<<ceil_div(n_cand, 256), 256, 0, st>>>(partial, n_chunk, n_cand, col_sum);
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

extern "C" int bess_rows_axpy(int dtype, const float* alpha, float scale, bess_rows_t src,
                              bess_rows_t out, int n, int width, void* stream) {
  if (n == 0) return BESS_OK;
  L2_DISPATCH_DTYPE(dtype, rows_axpy_kernel<T><<<warp_grid_(n), 256, 0, (cudaStream_t)stream>>>(
                               alpha, scale, src, out, n, width));
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}
