// Negative scoring for the score functions whose residual reads TWO candidate elements per
// coordinate: InterHT (scoring.py:1418-1572) and TranS (scoring.py:1575-1750).
//
//   score(q, c) = -|| e ||_p ,  e_k = qv0[k] * c^m_k + qv0[d + k] * c^a_k + qv1[k] ,  k < d
//
// with c^m / c^a the (optionally L2-normalised) main / auxiliary halves of the candidate row
// [main d | aux d] and (qv0, qv1) the query vectors of the prologue (families.cuh).  These are
// SURVEY 8(f) rank-4 families: plain CUDA-core kernels, deterministic (fixed summation orders,
// no atomics), correctness first — the register-tiled, pipe-tuned kernels of pair.cu serve
// the families BASELINE.json names.
//   shared negatives : 64 x 64 (query, candidate) tiles through shared memory, 4 x 4 per thread
//   their backward   : thread = coordinate; dQ streams all candidates per 8-query tile, dC
//                      streams a slice of the queries per 8-candidate tile (+ fixed-order reduce)
//   per-triple       : block per query, warp per candidate (fused gather + score)
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

struct P2Args {
  int p, d, W;
  const float* qv;      // [n_query, 2, W]
  int n_query;
  bess_rows_t cand;     // rows of W = 2d elements
  const float* scale;   // [2, n_cand] inverse norms of the halves (shared kernels) or null
  int n_cand;
  bess_rowmap_t score_map;
  int64_t ld;
  int col0;
  const float* score;
  const float* d_score;
  float* out;
  int normalize;        // per-triple kernels normalise in line
};

// dL/de_k = coef * (p == 1 ? sign(e_k) : e_k), coef from the score (= -norm) and dL/dscore
BESS_D float p2_coef(int p, float score, float g) {
  if (p == 1) return -g;
  return score != 0.f ? g / score : 0.f;
}
BESS_D float p2_de(int p, float coef, float e) { return p == 1 ? sign_mul(coef, e) : coef * e; }

// ------------------------------------------------------------ shared: forward --
constexpr int P2_T = 64, P2_K = 16;
template <typename CT>
__global__ void __launch_bounds__(256) pair2_fwd_kernel(P2Args a) {
  __shared__ float Qm[P2_T][P2_K + 1], Qa[P2_T][P2_K + 1], Qz[P2_T][P2_K + 1];
  __shared__ float Cm[P2_T][P2_K + 1], Ca[P2_T][P2_K + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int q0 = blockIdx.y * P2_T, c0 = blockIdx.x * P2_T;
  const CT* cbase = static_cast<const CT*>(a.cand.base);
  float acc[4][4] = {};
  for (int k0 = 0; k0 < a.d; k0 += P2_K) {
    for (int t = threadIdx.x; t < P2_T * P2_K; t += 256) {
      const int r = t / P2_K, kk = t - r * P2_K, k = k0 + kk;
      const int q = q0 + r, c = c0 + r;
      float m = 0.f, x = 0.f, z = 0.f, cm = 0.f, ca = 0.f;
      if (k < a.d) {
        if (q < a.n_query) {
          const float* qr = a.qv + (int64_t)q * 2 * a.W;
          m = qr[k]; x = qr[a.d + k]; z = qr[a.W + k];
        }
        if (c < a.n_cand) {
          const CT* row = cbase + src_row(a.cand, c) * a.cand.pitch;
          cm = ldf(row + k); ca = ldf(row + a.d + k);
          if (a.scale != nullptr) { cm *= a.scale[c]; ca *= a.scale[a.n_cand + c]; }
        }
      }
      Qm[r][kk] = m; Qa[r][kk] = x; Qz[r][kk] = z; Cm[r][kk] = cm; Ca[r][kk] = ca;
    }
    __syncthreads();
    const int kmax = min(P2_K, a.d - k0);
    for (int kk = 0; kk < kmax; ++kk) {
      float m[4], x[4], z[4], cm[4], ca[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        m[i] = Qm[ty * 4 + i][kk]; x[i] = Qa[ty * 4 + i][kk]; z[i] = Qz[ty * 4 + i][kk];
        cm[i] = Cm[tx * 4 + i][kk]; ca[i] = Ca[tx * 4 + i][kk];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += nacc(a.p, fmaf(m[i], cm[j], fmaf(x[i], ca[j], z[i])));
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + ty * 4 + i;
    if (q >= a.n_query) continue;
    float* orow = a.out + (int64_t)map_row(a.score_map, q) * a.ld + a.col0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + tx * 4 + j;
      if (c < a.n_cand) orow[c] = -nfin(a.p, acc[i][j]);
    }
  }
}

// ---------------------------------------------------- shared: dL/d(query vectors) --
// grid (ceil(n_query / 8), ceil(d / 128)); thread = coordinate k; every block streams all
// candidates in order -> deterministic sums.
constexpr int P2_BQ = 8, P2_CH = 64;
template <typename CT>
__global__ void __launch_bounds__(128) pair2_bwd_q_kernel(P2Args a, float* d_qv) {
  __shared__ float coef[P2_BQ][P2_CH];
  const int k = blockIdx.y * 128 + threadIdx.x;
  const int q0 = blockIdx.x * P2_BQ;
  const bool live = k < a.d;
  const CT* cbase = static_cast<const CT*>(a.cand.base);
  float m[P2_BQ], x[P2_BQ], z[P2_BQ], dm[P2_BQ] = {}, dx[P2_BQ] = {}, dz[P2_BQ] = {};
#pragma unroll
  for (int i = 0; i < P2_BQ; ++i) {
    const int q = q0 + i;
    m[i] = x[i] = z[i] = 0.f;
    if (live && q < a.n_query) {
      const float* qr = a.qv + (int64_t)q * 2 * a.W;
      m[i] = qr[k]; x[i] = qr[a.d + k]; z[i] = qr[a.W + k];
    }
  }
  for (int c0 = 0; c0 < a.n_cand; c0 += P2_CH) {
    __syncthreads();
    for (int t = threadIdx.x; t < P2_BQ * P2_CH; t += 128) {
      const int i = t / P2_CH, cc = t - i * P2_CH, q = q0 + i, c = c0 + cc;
      float v = 0.f;
      if (q < a.n_query && c < a.n_cand) {
        const int64_t at = (int64_t)map_row(a.score_map, q) * a.ld + a.col0 + c;
        v = p2_coef(a.p, a.score[at], a.d_score[at]);
      }
      coef[i][cc] = v;
    }
    __syncthreads();
    const int cmax = min(P2_CH, a.n_cand - c0);
    if (live) {
      for (int cc = 0; cc < cmax; ++cc) {
        const int c = c0 + cc;
        const CT* row = cbase + src_row(a.cand, c) * a.cand.pitch;
        float cm = ldf(row + k), ca = ldf(row + a.d + k);
        if (a.scale != nullptr) { cm *= a.scale[c]; ca *= a.scale[a.n_cand + c]; }
#pragma unroll
        for (int i = 0; i < P2_BQ; ++i) {
          const float de = p2_de(a.p, coef[i][cc], fmaf(m[i], cm, fmaf(x[i], ca, z[i])));
          dm[i] = fmaf(de, cm, dm[i]); dx[i] = fmaf(de, ca, dx[i]); dz[i] += de;
        }
      }
    }
  }
  if (!live) return;
#pragma unroll
  for (int i = 0; i < P2_BQ; ++i) {
    const int q = q0 + i;
    if (q >= a.n_query) continue;
    float* o = d_qv + (int64_t)q * 2 * a.W;
    o[k] = dm[i]; o[a.d + k] = dx[i]; o[a.W + k] = dz[i]; o[a.W + a.d + k] = 0.f;
  }
}

// ---------------------------------------------- shared: dL/d(normalised candidates) --
// grid (ceil(n_cand / 8), ceil(d / 128), split); block streams the queries of its split.
template <typename CT>
__global__ void __launch_bounds__(128) pair2_bwd_c_kernel(P2Args a, float* partial, int q_per_split) {
  __shared__ float coef[P2_CH][P2_BQ];
  const int k = blockIdx.y * 128 + threadIdx.x;
  const int c0 = blockIdx.x * P2_BQ;
  const bool live = k < a.d;
  const CT* cbase = static_cast<const CT*>(a.cand.base);
  float cm[P2_BQ], ca[P2_BQ], dcm[P2_BQ] = {}, dca[P2_BQ] = {};
#pragma unroll
  for (int j = 0; j < P2_BQ; ++j) {
    const int c = c0 + j;
    cm[j] = ca[j] = 0.f;
    if (live && c < a.n_cand) {
      const CT* row = cbase + src_row(a.cand, c) * a.cand.pitch;
      cm[j] = ldf(row + k); ca[j] = ldf(row + a.d + k);
      if (a.scale != nullptr) { cm[j] *= a.scale[c]; ca[j] *= a.scale[a.n_cand + c]; }
    }
  }
  const int qb = blockIdx.z * q_per_split, qe = min(a.n_query, qb + q_per_split);
  for (int q0 = qb; q0 < qe; q0 += P2_CH) {
    __syncthreads();
    for (int t = threadIdx.x; t < P2_CH * P2_BQ; t += 128) {
      const int qq = t / P2_BQ, j = t - qq * P2_BQ, q = q0 + qq, c = c0 + j;
      float v = 0.f;
      if (q < qe && c < a.n_cand) {
        const int64_t at = (int64_t)map_row(a.score_map, q) * a.ld + a.col0 + c;
        v = p2_coef(a.p, a.score[at], a.d_score[at]);
      }
      coef[qq][j] = v;
    }
    __syncthreads();
    const int qmax = min(P2_CH, qe - q0);
    if (live) {
      for (int qq = 0; qq < qmax; ++qq) {
        const float* qr = a.qv + (int64_t)(q0 + qq) * 2 * a.W;
        const float m = qr[k], x = qr[a.d + k], z = qr[a.W + k];
#pragma unroll
        for (int j = 0; j < P2_BQ; ++j) {
          const float de = p2_de(a.p, coef[qq][j], fmaf(m, cm[j], fmaf(x, ca[j], z)));
          dcm[j] = fmaf(de, m, dcm[j]); dca[j] = fmaf(de, x, dca[j]);
        }
      }
    }
  }
  if (!live) return;
#pragma unroll
  for (int j = 0; j < P2_BQ; ++j) {
    const int c = c0 + j;
    if (c >= a.n_cand) continue;
    float* o = partial + ((int64_t)blockIdx.z * a.n_cand + c) * a.W;
    o[k] = dcm[j]; o[a.d + k] = dca[j];
  }
}

__global__ void pair2_bwd_c_reduce_kernel(const float* partial, int split, int n_cand, int W,
                                          bess_rows_t d_cand, int add) {
  const int64_t total = (int64_t)n_cand * W;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(t / W), k = (int)(t - (int64_t)c * W);
    float s = 0.f;
    for (int z = 0; z < split; ++z) s += partial[((int64_t)z * n_cand + c) * W + k];
    float* o = reinterpret_cast<float*>(const_cast<void*>(d_cand.base)) + src_row(d_cand, c) * d_cand.pitch + k;
    *o = add ? *o + s : s;
  }
}

// ------------------------------------------------------------------ per-triple --
// Block per query, warp per candidate (round robin): candidate c of the query at position
// qpos is logical row cand.map(c) + qpos * q_stride (then cand.idx); normalisation in line.
constexpr int P2_WARPS = 8;
template <typename CT>
BESS_D void p2_row_norms(const P2Args& a, const CT* row, int lane, float& s1, float& s2, float& n1,
                         float& n2) {
  s1 = s2 = n1 = n2 = 1.f;
  if (!a.normalize) return;
  float u = 0.f, v = 0.f;
  for (int k = lane; k < a.d; k += 32) {
    const float x = ldf(row + k), y = ldf(row + a.d + k);
    u += x * x; v += y * y;
  }
  n1 = sqrtf(warp_sum(u)); n2 = sqrtf(warp_sum(v));
  s1 = 1.f / fmaxf(n1, 1e-12f); s2 = 1.f / fmaxf(n2, 1e-12f);
}

template <typename CT>
__global__ void __launch_bounds__(P2_WARPS * 32) pair2_pt_fwd_kernel(P2Args a, int64_t q_stride,
                                                                      int n_per) {
  extern __shared__ float sq[];  // [3 d]: qv0 main, qv0 aux, qv1
  const int q = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* qr = a.qv + (int64_t)q * 2 * a.W;
  for (int k = threadIdx.x; k < a.d; k += blockDim.x) {
    sq[k] = qr[k]; sq[a.d + k] = qr[a.d + k]; sq[2 * a.d + k] = qr[a.W + k];
  }
  __syncthreads();
  const int qpos = map_row(a.score_map, q);
  const CT* cbase = static_cast<const CT*>(a.cand.base);
  for (int c = warp; c < n_per; c += P2_WARPS) {
    int r = map_row(a.cand.map, c) + (int)(qpos * q_stride);
    if (a.cand.idx != nullptr) r = __ldg(a.cand.idx + r);
    const CT* row = cbase + (int64_t)r * a.cand.pitch;
    float s1, s2, n1, n2;
    p2_row_norms(a, row, lane, s1, s2, n1, n2);
    float acc = 0.f;
    for (int k = lane; k < a.d; k += 32)
      acc += nacc(a.p, fmaf(sq[k], ldf(row + k) * s1, fmaf(sq[a.d + k], ldf(row + a.d + k) * s2,
                                                           sq[2 * a.d + k])));
    acc = warp_sum(acc);
    if (lane == 0) a.out[(int64_t)qpos * a.ld + a.col0 + c] = -nfin(a.p, acc);
  }
}

// backward: d_qv accumulated per warp in shared memory (lane-owned coordinates), reduced over the
// warps in a fixed order at the end; d_cand row written per candidate incl. the normalisation
// chain rule.
template <typename CT>
__global__ void __launch_bounds__(P2_WARPS * 32) pair2_pt_bwd_kernel(P2Args a, int64_t q_stride,
                                                                      int n_per, float* d_qv,
                                                                      bess_rows_t d_cand) {
  extern __shared__ float sm[];  // [3 d] query | [P2_WARPS][3 d] per-warp gradient accumulators
  const int q = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d = a.d;
  float* sq = sm;
  float* acc = sm + 3 * d + warp * 3 * d;
  const float* qr = a.qv + (int64_t)q * 2 * a.W;
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    sq[k] = qr[k]; sq[d + k] = qr[d + k]; sq[2 * d + k] = qr[a.W + k];
  }
  for (int k = lane; k < 3 * d; k += 32) acc[k] = 0.f;
  __syncthreads();
  const int qpos = map_row(a.score_map, q);
  const CT* cbase = static_cast<const CT*>(a.cand.base);
  for (int c = warp; c < n_per; c += P2_WARPS) {
    const int lr = map_row(a.cand.map, c) + (int)(qpos * q_stride);
    const int r = a.cand.idx != nullptr ? __ldg(a.cand.idx + lr) : lr;
    const CT* row = cbase + (int64_t)r * a.cand.pitch;
    float s1, s2, n1, n2;
    p2_row_norms(a, row, lane, s1, s2, n1, n2);
    const int64_t at = (int64_t)qpos * a.ld + a.col0 + c;
    const float coef = p2_coef(a.p, a.score[at], a.d_score[at]);
    // pass B: query gradients and the projections c^ . dc^ of the two halves
    float pm = 0.f, pa = 0.f;
    for (int k = lane; k < d; k += 32) {
      const float cm = ldf(row + k) * s1, ca = ldf(row + d + k) * s2;
      const float de = p2_de(a.p, coef, fmaf(sq[k], cm, fmaf(sq[d + k], ca, sq[2 * d + k])));
      acc[k] = fmaf(de, cm, acc[k]); acc[d + k] = fmaf(de, ca, acc[d + k]); acc[2 * d + k] += de;
      pm += cm * (de * sq[k]); pa += ca * (de * sq[d + k]);
    }
    if (a.normalize) { pm = warp_sum(pm); pa = warp_sum(pa); }
    // pass C: the candidate's gradient row (addressed like the candidate, without its index list)
    float* o = reinterpret_cast<float*>(const_cast<void*>(d_cand.base)) + (int64_t)(map_row(d_cand.map, c) + (int)(qpos * q_stride)) * d_cand.pitch;
    for (int k = lane; k < d; k += 32) {
      const float cm = ldf(row + k) * s1, ca = ldf(row + d + k) * s2;
      const float de = p2_de(a.p, coef, fmaf(sq[k], cm, fmaf(sq[d + k], ca, sq[2 * d + k])));
      o[k] = unnorm_grad(a.normalize, de * sq[k], cm, pm, n1, s1);
      o[d + k] = unnorm_grad(a.normalize, de * sq[d + k], ca, pa, n2, s2);
    }
  }
  __syncthreads();
  float* o = d_qv + (int64_t)q * 2 * a.W;
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
    for (int w = 0; w < P2_WARPS; ++w) {
      const float* aw = sm + 3 * d + w * 3 * d;
      g0 += aw[k]; g1 += aw[d + k]; g2 += aw[2 * d + k];
    }
    o[k] = g0; o[d + k] = g1; o[a.W + k] = g2; o[a.W + d + k] = 0.f;
  }
}

static P2Args p2_args(const bess_score_cfg_t* cfg, const float* qv, int n_query, bess_rows_t cand,
                      const float* scale, int n_cand, bess_rowmap_t score_map, int64_t ld, int col0) {
  P2Args a;
  a.p = cfg->norm_p; a.d = cfg->d; a.W = 2 * cfg->d; a.qv = qv; a.n_query = n_query; a.cand = cand;
  a.scale = scale; a.n_cand = n_cand; a.score_map = score_map; a.ld = ld; a.col0 = col0;
  a.score = nullptr; a.d_score = nullptr; a.out = nullptr; a.normalize = cfg->normalize;
  return a;
}
static int p2_split(int n_query, int n_cand, int d) {
  const int tiles = ceil_div(n_cand, P2_BQ) * ceil_div(d, 128);
  int split = ceil_div(4 * kNumSM, tiles);
  const int max_split = ceil_div(n_query, 4 * P2_CH);
  if (split > max_split) split = max_split;
  return split < 1 ? 1 : split;
}

#define P2_DISPATCH(dtype, ...)                                      \
  switch (dtype) {                                                   \
    case BESS_F32: { using CT = float; __VA_ARGS__; break; }         \
    case BESS_F16: { using CT = __half; __VA_ARGS__; break; }        \
    case BESS_BF16: { using CT = __nv_bfloat16; __VA_ARGS__; break; } \
    default: bess_set_error("unknown dtype %d", dtype); return BESS_ERR_INVALID_ARG; \
  }

// ---- entry points called from the pair.cu C-ABI functions for OP_PAIR2 families -------------
int pair2_shared_fwd(const bess_score_cfg_t* cfg, int dtype, const float* qv, int n_query,
                     bess_rows_t cand, const float* scale, int n_cand, float* out,
                     bess_rowmap_t score_map, int64_t ld, int col0, cudaStream_t st) {
  if (n_query == 0 || n_cand == 0) return BESS_OK;
  P2Args a = p2_args(cfg, qv, n_query, cand, scale, n_cand, score_map, ld, col0);
  a.out = out;
  const dim3 grid(ceil_div(n_cand, P2_T), ceil_div(n_query, P2_T));
  P2_DISPATCH(dtype, pair2_fwd_kernel<CT><<<grid, 256, 0, st>>>(a));
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

int pair2_shared_bwd_query(const bess_score_cfg_t* cfg, int dtype, const float* qv, int n_query,
                           bess_rows_t cand, const float* scale, int n_cand, const float* score,
                           const float* d_score, bess_rowmap_t score_map, int64_t ld, int col0,
                           float* d_qv, cudaStream_t st) {
  if (n_query == 0) return BESS_OK;
  P2Args a = p2_args(cfg, qv, n_query, cand, scale, n_cand, score_map, ld, col0);
  a.score = score; a.d_score = d_score;
  const dim3 grid(ceil_div(n_query, P2_BQ), ceil_div(a.d, 128));
  P2_DISPATCH(dtype, pair2_bwd_q_kernel<CT><<<grid, 128, 0, st>>>(a, d_qv));
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

int64_t pair2_bwd_cand_workspace(const bess_score_cfg_t* cfg, int n_query, int n_cand) {
  return (int64_t)p2_split(n_query, n_cand, cfg->d) * n_cand * 2 * cfg->d * sizeof(float);
}

int pair2_shared_bwd_cand(const bess_score_cfg_t* cfg, int dtype, const float* qv, int n_query,
                          bess_rows_t cand, const float* scale, int n_cand, const float* score,
                          const float* d_score, bess_rowmap_t score_map, int64_t ld, int col0,
                          bess_rows_t d_cand, int add_cand, void* workspace, cudaStream_t st) {
  if (n_cand == 0) return BESS_OK;
  BESS_CHECK_ARG(workspace != nullptr, "workspace required");
  P2Args a = p2_args(cfg, qv, n_query, cand, scale, n_cand, score_map, ld, col0);
  a.score = score; a.d_score = d_score;
  const int split = p2_split(n_query, n_cand, a.d);
  int q_per_split = ceil_div(ceil_div(n_query, split), P2_CH) * P2_CH;
  if (q_per_split < P2_CH) q_per_split = P2_CH;
  float* partial = static_cast<float*>(workspace);
  const dim3 grid(ceil_div(n_cand, P2_BQ), ceil_div(a.d, 128), split);
  P2_DISPATCH(dtype, pair2_bwd_c_kernel<CT><<<grid, 128, 0, st>>>(a, partial, q_per_split));
  BESS_CHECK_LAUNCH();
  const int64_t total = (int64_t)n_cand * a.W;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 8 * kNumSM) blocks = 8 * kNumSM;
  pair2_bwd_c_reduce_kernel<<<blocks, 256, 0, st>>>(partial, split, n_cand, a.W, d_cand, add_cand);
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

int pair2_pertriple_fwd(const bess_score_cfg_t* cfg, int dtype, const float* qv, int n_query,
                        bess_rows_t cand, int64_t q_stride, int n_per, float* out,
                        bess_rowmap_t score_map, int64_t ld, int col0, cudaStream_t st) {
  if (n_query == 0 || n_per == 0) return BESS_OK;
  P2Args a = p2_args(cfg, qv, n_query, cand, nullptr, 0, score_map, ld, col0);
  a.out = out;
  const size_t smem = (size_t)3 * a.d * sizeof(float);
  BESS_CHECK_ARG(smem <= 48 * 1024, "embedding_size too large for the per-triple kernel");
  P2_DISPATCH(dtype, pair2_pt_fwd_kernel<CT><<<n_query, P2_WARPS * 32, smem, st>>>(a, q_stride, n_per));
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

int pair2_pertriple_bwd(const bess_score_cfg_t* cfg, int dtype, const float* qv, int n_query,
                        bess_rows_t cand, int64_t q_stride, int n_per, const float* score,
                        const float* d_score, bess_rowmap_t score_map, int64_t ld, int col0,
                        float* d_qv, bess_rows_t d_cand, cudaStream_t st) {
  if (n_query == 0) return BESS_OK;
  P2Args a = p2_args(cfg, qv, n_query, cand, nullptr, 0, score_map, ld, col0);
  a.score = score; a.d_score = d_score;
  const size_t smem = (size_t)3 * a.d * sizeof(float) * (1 + P2_WARPS);
  BESS_CHECK_ARG(smem <= 200 * 1024, "embedding_size too large for the per-triple backward kernel");
  P2_DISPATCH(dtype, {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(pair2_pt_bwd_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    pair2_pt_bwd_kernel<CT><<<n_query, P2_WARPS * 32, smem, st>>>(a, q_stride, n_per, d_qv, d_cand);
  });
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

}  // namespace bess
