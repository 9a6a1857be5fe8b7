// Negative-sample scoring kernels (CUDA-core path).
//
//  * shared negatives  (negative_sample_sharing=True, scoring.py:176-200,
//    231-255, 569-573, 1378-1389): every query against ONE candidate list.
//    Register-tiled [128 x 128] CTA tiles over (query, candidate) with the
//    embedding dimension streamed through shared memory in chunks, like a SIMT
//    GEMM whose inner product is replaced by the family's pair function
//    (|q-c|, (q-c)^2, q*c, PairRE, BoxE).  Backward = two contractions of the
//    same shape (over candidates for dQ, over queries for dC).
//  * per-triple negatives (scoring.py:199, 254): pure streaming; one CTA per
//    query, one warp per candidate row, rows are read in place through an
//    index list (fused gather + score).
//
// The DOT / L2 cases also have a tcgen05 tensor-core implementation
// (gemm_tc.cu); this file is the exact-fp32 CUDA-core path and the only path
// for L1 / PairRE / BoxE.
#include <stdlib.h>

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

static inline FamCfg to_cfg(const bess_score_cfg_t* c) {
  FamCfg f;
  f.family = c->family; f.norm_p = c->norm_p; f.d = c->d; f.normalize = c->normalize;
  f.apply_tanh = c->apply_tanh; f.per_dim = c->per_dim; f.eps = c->eps; f.rel_u = c->rel_u;
  return f;
}

template <int OP>
struct OpTraits {
  static constexpr int NV = OP == OP_PAIRRE ? 2 : (OP == OP_BOXE ? 3 : 1);
  static constexpr int NSEG = OP == OP_BOXE ? 2 : 1;
};

// Load kVec consecutive coordinates [k, k+kVec) of a candidate row as floats;
// coordinate kk reads element (kk + rot) % W; coordinates >= W read as 0.
template <typename CT>
BESS_D void load_cand_vec(const CT* row, int k, int W, int rot, bool vec_ok,
                          float (&out)[Elem<CT>::kVec]) {
  constexpr int V = Elem<CT>::kVec;
  if (vec_ok && k + V <= W) {
    int e = k + rot;
    if (e >= W) e -= W;
    Elem<CT>::load_vec(row + e, out);
  } else {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int kk = k + i;
      float v = 0.f;
      if (kk < W) {
        int e = kk + rot;
        if (e >= W) e -= W;
        v = ldf(row + e);
      }
      out[i] = v;
    }
  }
}

struct PairArgs {
  const float* qv;   // [n_query, NV, W]
  int n_query;
  bess_rows_t cand;
  const float* cand_scale;
  int n_cand;
  int W;
  int rot;           // BoxE candidate rotation
  int apply_tanh;
  int vec_ok;
  bess_rowmap_t score_map;
  int64_t ld;
  int col0;
  float* out;        // fwd: scores
  float* aux;        // BoxE p=2: first-box norm
  const float* score;    // bwd
  const float* d_score;  // bwd
};

// ---------------------------------------------------------------------------
// Forward tile kernel.  512 threads = 16 warps per CTA, one CTA per SM: warp ty
// owns MR query rows, lane tx owns 4 candidate columns, so a CTA covers
// [16 * MR queries x 128 candidates] with an MR x 4 register tile per thread.
//   * 16 resident warps (4 per scheduler) hide the shared-memory and FADD
//     latencies that limited the earlier 256-thread / 8 x 8 version to 66 % issue
//     (ncu, profiles/r01f_*: 2 warps per scheduler, stalls wait + short_scoreboard).
//   * MR in {6, 7, 8} is chosen per launch so that the tile count is as close as
//     possible to a multiple of the 148 SMs (S = 8192, N = 256: 74 x 2 = 148 tiles
//     of 112 queries instead of 128 tiles of 128).
//   * software pipeline: the next K chunk travels global -> registers while the
//     current one is reduced from shared memory.
//   * a warp reads its query values as a broadcast (all lanes share ty) and its
//     candidate values as one conflict-free 512-byte row segment.
// ---------------------------------------------------------------------------
constexpr int F_TC = 128, F_KC = 16, F_NT = 512;

template <int OP, int P, typename CT, int MR>
__global__ void __launch_bounds__(F_NT) pair_fwd_kernel(PairArgs a) {
  constexpr int NV = OpTraits<OP>::NV;
  constexpr int NSEG = OpTraits<OP>::NSEG;
  constexpr int V = Elem<CT>::kVec;
  constexpr int TQ = 16 * MR;
  // two stages: the next chunk is stored while other warps still read the current one, so a
  // chunk costs ONE barrier (ncu on the single-stage version: barrier 0.88 stalls per issue)
  // (BoxE's three query vectors would exceed the 48 KB static limit: single stage, two barriers)
  constexpr int NST = NV <= 2 ? 2 : 1;
  __shared__ __align__(16) float Qs[NST][NV][F_KC][16][8];  // [stage][v][k][warp][row slot]; slots >= MR unused
  __shared__ __align__(16) float Cs[NST][F_KC][F_TC];

  const int tid = threadIdx.x;
  const int tx = tid & 31, ty = tid >> 5;
  const int q0 = blockIdx.y * TQ, c0 = blockIdx.x * F_TC;
  const int W = a.W;
  const int qv_row = NV * W;

  float acc[MR][4];
  constexpr int SZI = NSEG == 2 ? MR : 1, SZJ = NSEG == 2 ? 4 : 1;  // BoxE keeps the first-box norm
  float tot[SZI][SZJ];
#pragma unroll
  for (int i = 0; i < MR; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
  for (int i = 0; i < SZI; ++i)
#pragma unroll
    for (int j = 0; j < SZJ; ++j) tot[i][j] = 0.f;

  constexpr int Q_LOADS = NV * TQ * (F_KC / 4);            // float4 loads of a query tile
  constexpr int QL = (Q_LOADS + F_NT - 1) / F_NT;
  constexpr int C_LOADS = F_TC * (F_KC / V);               // 128-bit loads of a candidate tile
  static_assert(C_LOADS <= F_NT, "candidate tile: at most one load per thread");
  float4 qreg[QL];
  float creg[V];
  const bool q_vec = (W & 3) == 0;

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int l = 0; l < QL; ++l) {
      const int s = tid + l * F_NT;
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (s < Q_LOADS) {
        const int q = s % TQ;
        const int kv = (s / TQ) % (F_KC / 4);
        const int v = s / (TQ * (F_KC / 4));
        const int k = k0 + kv * 4;
        if (q0 + q < a.n_query) {
          const float* src = a.qv + (int64_t)(q0 + q) * qv_row + v * W + k;
          if (k + 4 <= W && q_vec) {
            val = *reinterpret_cast<const float4*>(src);
          } else {
            if (k + 0 < W) val.x = src[0];
            if (k + 1 < W) val.y = src[1];
            if (k + 2 < W) val.z = src[2];
            if (k + 3 < W) val.w = src[3];
          }
        }
      }
      qreg[l] = val;
    }
#pragma unroll
    for (int i = 0; i < V; ++i) creg[i] = 0.f;
    if (tid < C_LOADS) {
      const int c = tid % F_TC;
      const int kv = tid / F_TC;
      if (c0 + c < a.n_cand) {
        const CT* row = static_cast<const CT*>(a.cand.base) + src_row(a.cand, c0 + c) * a.cand.pitch;
        load_cand_vec<CT>(row, k0 + kv * V, W, a.rot, a.vec_ok, creg);
        if (OP == OP_PAIRRE && a.cand_scale != nullptr) {
          const float sc = a.cand_scale[c0 + c];
#pragma unroll
          for (int i = 0; i < V; ++i) creg[i] *= sc;
        }
      }
    }
  };
  auto store_tiles = [&](int st) {
#pragma unroll
    for (int l = 0; l < QL; ++l) {
      const int s = tid + l * F_NT;
      if (s < Q_LOADS) {
        const int q = s % TQ;
        const int kv = (s / TQ) % (F_KC / 4);
        const int v = s / (TQ * (F_KC / 4));
        const int wq = q / MR, slot = q - wq * MR;
        Qs[st][v][kv * 4 + 0][wq][slot] = qreg[l].x; Qs[st][v][kv * 4 + 1][wq][slot] = qreg[l].y;
        Qs[st][v][kv * 4 + 2][wq][slot] = qreg[l].z; Qs[st][v][kv * 4 + 3][wq][slot] = qreg[l].w;
      }
    }
    if (tid < C_LOADS) {
      const int c = tid % F_TC;
      const int kv = tid / F_TC;
#pragma unroll
      for (int i = 0; i < V; ++i) Cs[st][kv * V + i][c] = creg[i];
    }
  };

  load_tiles(0);
  store_tiles(0);
  __syncthreads();

  int st = 0;
  for (int k0 = 0; k0 < W; k0 += F_KC, st = (st ^ 1) & (NST - 1)) {
    const bool has_next = k0 + F_KC < W;
    if (has_next) load_tiles(k0 + F_KC);

    if (NSEG == 2 && k0 == W / 2) {  // BoxE: first box finished, start the second
#pragma unroll
      for (int i = 0; i < MR; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          tot[i % SZI][j % SZJ] = nfin(P, acc[i][j]);
          acc[i][j] = 0.f;
        }
    }

#pragma unroll 8
    for (int k = 0; k < F_KC; ++k) {
      float qf[NV][8];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const float4 lo = *reinterpret_cast<const float4*>(&Qs[st][v][k][ty][0]);
        qf[v][0] = lo.x; qf[v][1] = lo.y; qf[v][2] = lo.z; qf[v][3] = lo.w;
        if (MR > 4) {
          const float4 hi = *reinterpret_cast<const float4*>(&Qs[st][v][k][ty][4]);
          qf[v][4] = hi.x; qf[v][5] = hi.y; qf[v][6] = hi.z; qf[v][7] = hi.w;
        }
      }
      const float4 c4 = *reinterpret_cast<const float4*>(&Cs[st][k][tx * 4]);
      const float cf[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
      for (int i = 0; i < MR; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          acc[i][j] += pair_elem<OP>(P, a.apply_tanh, qf[0][i], NV > 1 ? qf[NV > 1 ? 1 : 0][i] : 0.f,
                                     NV > 2 ? qf[NV > 2 ? 2 : 0][i] : 0.f, cf[j]);
    }
    // stage st^1 was last read two chunks ago; every warp has passed the barrier that followed
    // those reads, so it can be overwritten now; one barrier publishes it for the next chunk
    if (NST == 2) {
      if (has_next) store_tiles(st ^ 1);
      __syncthreads();
    } else {
      __syncthreads();
      if (has_next) {
        store_tiles(0);
        __syncthreads();
      }
    }
  }

  // ---- epilogue: a lane owns 4 consecutive columns -> one 128-bit store per row when aligned
  const int c = c0 + tx * 4;
  const bool st_vec = (a.ld & 3) == 0 && ((a.col0 + c) & 3) == 0 && c + 4 <= a.n_cand &&
                      (reinterpret_cast<uintptr_t>(a.out) & 15) == 0 &&
                      (a.aux == nullptr || (reinterpret_cast<uintptr_t>(a.aux) & 15) == 0);
#pragma unroll
  for (int i = 0; i < MR; ++i) {
    const int q = q0 + ty * MR + i;
    if (q >= a.n_query) continue;
    const int64_t orow = (int64_t)map_row(a.score_map, q) * a.ld + a.col0;
    float s[4], t0[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      t0[j] = NSEG == 2 ? tot[i % SZI][j % SZJ] : 0.f;
      if (OP == OP_DOT) s[j] = acc[i][j];
      else if (NSEG == 2) s[j] = -(t0[j] + nfin(P, acc[i][j]));
      else s[j] = -nfin(P, acc[i][j]);
    }
    if (st_vec) {
      *reinterpret_cast<float4*>(a.out + orow + c) = make_float4(s[0], s[1], s[2], s[3]);
      if (NSEG == 2 && a.aux != nullptr)
        *reinterpret_cast<float4*>(a.aux + orow + c) = make_float4(t0[0], t0[1], t0[2], t0[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c + j < a.n_cand) {
          a.out[orow + c + j] = s[j];
          if (NSEG == 2 && a.aux != nullptr) a.aux[orow + c + j] = t0[j];
        }
    }
  }
}

// L1 sub-gradient fast path (half-precision tables only, see pair_bwd_q_kernel)
template <int OP, int P, typename CT>
struct kFastL1 { static constexpr bool value = false; };
template <>
struct kFastL1<OP_DIST, 1, __half> { static constexpr bool value = true; };
template <>
struct kFastL1<OP_DIST, 1, __nv_bfloat16> { static constexpr bool value = true; };
// 1 / 0.5 / 0 for e > 0 / e == 0 / e < 0: one FFMA.SAT (every normal e saturates)
BESS_D float sat_half_sign(float e) { return __saturatef(fmaf(e, 8.507059173023462e37f, 0.5f)); }

// Per-pair gradient coefficient (see pair_elem_bwd): returns coefficient for
// norm segment `seg`.
template <int OP, int P>
BESS_D float pair_coef(float g, float score, float aux, int seg) {
  if (OP == OP_DOT) return g;
  if (P == 1) return -g;
  float nv;
  if (OP == OP_BOXE) {
    nv = seg == 0 ? aux : (-score - aux);
  } else {
    nv = -score;
  }
  return nv > 0.f ? -g / nv : 0.f;
}

// ---------------------------------------------------------------------------
// Backward w.r.t. query vectors: CTA owns [64 queries x 64 coordinates] and
// streams all candidates.
// ---------------------------------------------------------------------------
constexpr int B_T = 64, B_TK = 64, B_CH = 32, B_GS = B_T + 4;

// NT threads own [NT / 4 queries x 64 coordinates]; NT = 128 halves the tile when that fills
// the 148 SMs more evenly (bwd_q_threads()).
template <int OP, int P, typename CT, int NT>
__global__ void __launch_bounds__(NT, OpTraits<OP>::NV == 1 ? (NT == 128 ? 7 : 3) : 1)
pair_bwd_q_kernel(PairArgs a, float* d_qv) {
  constexpr int BT = NT / 4;       // queries per CTA
  constexpr int BGS = BT + 4;
  constexpr int NV = OpTraits<OP>::NV;
  constexpr int NSEG = OpTraits<OP>::NSEG;
  constexpr int V = Elem<CT>::kVec;
  __shared__ __align__(16) float Gs[NSEG][B_CH][BGS];  // coef[c][q]
  __shared__ __align__(16) float Cs[B_CH][B_TK];        // cand[c][k]

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int q0 = blockIdx.x * BT, k0 = blockIdx.y * B_TK;
  const int W = a.W, qv_row = NV * W;

  float qr[NV][4][4], acc[NV][4][4];
  // FAST (L1 distance, half tables): acc += (2 coef) * u with u = sat(e * 2^126 + 0.5) in
  // {0, 0.5, 1}, and the constant part sum_c coef is subtracted once at the end — 3 FMA-pipe
  // instructions per element instead of 4.  The two sums cancel, which costs ~N/2 ulp of
  // relative accuracy: fine at the 1e-2 bar of bf16 / fp16 tables, not used for fp32 ones.
  constexpr bool FAST = kFastL1<OP, P, CT>::value;
  float csum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = q0 + ty * 4 + i, k = k0 + tx * 4 + j;
        qr[v][i][j] = (q < a.n_query && k < W) ? a.qv[(int64_t)q * qv_row + v * W + k] : 0.f;
        acc[v][i][j] = 0.f;
      }
  // segment of each of this thread's coordinates (BoxE)
  int segj[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) segj[j] = (NSEG == 2 && (k0 + tx * 4 + j) >= W / 2) ? 1 : 0;

  for (int c0 = 0; c0 < a.n_cand; c0 += B_CH) {
    // coef tile: warp w loads query rows w, w+8, ...; lane = candidate (coalesced)
    {
      const int lane = tid & 31, w = tid >> 5;
      for (int q = w; q < BT; q += NT / 32) {
        const int c = c0 + lane;
        float cf0 = 0.f, cf1 = 0.f;
        if (q0 + q < a.n_query && c < a.n_cand) {
          const int64_t off = (int64_t)map_row(a.score_map, q0 + q) * a.ld + a.col0 + c;
          const float g = a.d_score[off];
          const float sc = (P == 2 && OP != OP_DOT) ? a.score[off] : 0.f;
          const float ax = (P == 2 && OP == OP_BOXE) ? a.aux[off] : 0.f;
          cf0 = pair_coef<OP, P>(g, sc, ax, 0);
          if (NSEG == 2) cf1 = pair_coef<OP, P>(g, sc, ax, 1);
        }
        Gs[0][lane][q] = cf0;
        if (NSEG == 2) Gs[NSEG - 1][lane][q] = cf1;
      }
    }
    // candidate tile
    for (int s = tid; s < B_CH * (B_TK / V); s += NT) {
      const int kv = s % (B_TK / V);
      const int c = s / (B_TK / V);
      float vals[V];
#pragma unroll
      for (int i = 0; i < V; ++i) vals[i] = 0.f;
      if (c0 + c < a.n_cand) {
        const CT* row = static_cast<const CT*>(a.cand.base) + src_row(a.cand, c0 + c) * a.cand.pitch;
        load_cand_vec<CT>(row, k0 + kv * V, W, a.rot, a.vec_ok, vals);
        if (OP == OP_PAIRRE && a.cand_scale != nullptr) {
          const float sc = a.cand_scale[c0 + c];
#pragma unroll
          for (int i = 0; i < V; ++i) vals[i] *= sc;
        }
      }
#pragma unroll
      for (int i = 0; i < V; ++i) Cs[c][kv * V + i] = vals[i];
    }
    __syncthreads();

#pragma unroll 4
    for (int c = 0; c < B_CH; ++c) {
      const float4 g0 = *reinterpret_cast<const float4*>(&Gs[0][c][ty * 4]);
      float4 g1 = g0;
      if (NSEG == 2) g1 = *reinterpret_cast<const float4*>(&Gs[NSEG - 1][c][ty * 4]);
      const float4 cv4 = *reinterpret_cast<const float4*>(&Cs[c][tx * 4]);
      const float gq0[4] = {g0.x, g0.y, g0.z, g0.w};
      const float gq1[4] = {g1.x, g1.y, g1.z, g1.w};
      const float cv[4] = {cv4.x, cv4.y, cv4.z, cv4.w};
      if (FAST) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float g2 = gq0[i] + gq0[i];
          csum[i] += g2;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            acc[0][i][j] = fmaf(g2, sat_half_sign(qr[0][i][j] - cv[j]), acc[0][i][j]);
        }
        continue;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float coef = (NSEG == 2 && segj[j]) ? gq1[i] : gq0[i];
          float d0, d1, d2, dc;
          pair_elem_bwd<OP>(P, a.apply_tanh, qr[0][i][j], NV > 1 ? qr[NV > 1 ? 1 : 0][i][j] : 0.f,
                            NV > 2 ? qr[NV > 2 ? 2 : 0][i][j] : 0.f, cv[j], coef, d0, d1, d2, dc);
          acc[0][i][j] += d0;
          if (NV > 1) acc[NV > 1 ? 1 : 0][i][j] += d1;
          if (NV > 2) acc[NV > 2 ? 2 : 0][i][j] += d2;
        }
    }
    __syncthreads();
  }

#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = q0 + ty * 4 + i;
      if (q >= a.n_query) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + tx * 4 + j;
        if (k < W) d_qv[(int64_t)q * qv_row + v * W + k] = FAST ? fmaf(-0.5f, csum[i], acc[v][i][j]) : acc[v][i][j];
      }
    }
}

// ---------------------------------------------------------------------------
// Backward w.r.t. candidate rows: CTA owns [64 candidates x 64 coordinates]
// and streams a slice of the queries (split-S; partials are reduced in a fixed
// order afterwards -> deterministic).
// ---------------------------------------------------------------------------
template <int OP, int P, typename CT>
__global__ void __launch_bounds__(256, OpTraits<OP>::NV == 1 ? 4 : 1)
pair_bwd_c_kernel(PairArgs a, float* partial, int q_per_split) {
  constexpr int NV = OpTraits<OP>::NV;
  constexpr int NSEG = OpTraits<OP>::NSEG;
  __shared__ __align__(16) float Gs[NSEG][B_CH][B_T];      // coef[q][c]
  __shared__ __align__(16) float Qs[NV][B_CH][B_TK];       // qv[q][k]

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int c0 = blockIdx.x * B_T, k0 = blockIdx.y * B_TK;
  const int W = a.W, qv_row = NV * W;
  const int qs = blockIdx.z * q_per_split;
  const int qe = min(a.n_query, qs + q_per_split);

  float cr[4][4], acc[4][4];
  constexpr bool FAST = kFastL1<OP, P, CT>::value;  // see pair_bwd_q_kernel
  float csum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty * 4 + i;
    const CT* row = nullptr;
    float sc = 1.f;
    if (c < a.n_cand) {
      row = static_cast<const CT*>(a.cand.base) + src_row(a.cand, c) * a.cand.pitch;
      if (OP == OP_PAIRRE && a.cand_scale != nullptr) sc = a.cand_scale[c];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      float v = 0.f;
      if (row != nullptr && k < W) {
        int e = k + a.rot;
        if (e >= W) e -= W;
        v = ldf(row + e) * sc;
      }
      cr[i][j] = v;
      acc[i][j] = 0.f;
    }
  }
  int segj[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) segj[j] = (NSEG == 2 && (k0 + tx * 4 + j) >= W / 2) ? 1 : 0;

  for (int qb = qs; qb < qe; qb += B_CH) {
    // coef tile Gs[q][c]: warp w loads query rows w, w+8, ...; lanes cover 64 candidates
    {
      const int lane = tid & 31, w = tid >> 5;
      for (int q = w; q < B_CH; q += 8) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int cl = lane + 32 * h;
          const int c = c0 + cl;
          float cf0 = 0.f, cf1 = 0.f;
          if (qb + q < qe && c < a.n_cand) {
            const int64_t off = (int64_t)map_row(a.score_map, qb + q) * a.ld + a.col0 + c;
            const float g = a.d_score[off];
            const float sc = (P == 2 && OP != OP_DOT) ? a.score[off] : 0.f;
            const float ax = (P == 2 && OP == OP_BOXE) ? a.aux[off] : 0.f;
            cf0 = pair_coef<OP, P>(g, sc, ax, 0);
            if (NSEG == 2) cf1 = pair_coef<OP, P>(g, sc, ax, 1);
          }
          Gs[0][q][cl] = cf0;
          if (NSEG == 2) Gs[NSEG - 1][q][cl] = cf1;
        }
      }
    }
    // query tile Qs[v][q][k]
    for (int s = tid; s < NV * B_CH * (B_TK / 4); s += 256) {
      const int kv = s % (B_TK / 4);
      const int q = (s / (B_TK / 4)) % B_CH;
      const int v = s / ((B_TK / 4) * B_CH);
      const int k = k0 + kv * 4;
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (qb + q < qe) {
        const float* src = a.qv + (int64_t)(qb + q) * qv_row + v * W + k;
        if (k + 4 <= W && (W & 3) == 0) {
          val = *reinterpret_cast<const float4*>(src);
        } else {
          if (k + 0 < W) val.x = src[0];
          if (k + 1 < W) val.y = src[1];
          if (k + 2 < W) val.z = src[2];
          if (k + 3 < W) val.w = src[3];
        }
      }
      *reinterpret_cast<float4*>(&Qs[v][q][kv * 4]) = val;
    }
    __syncthreads();

#pragma unroll 4
    for (int q = 0; q < B_CH; ++q) {
      const float4 g0 = *reinterpret_cast<const float4*>(&Gs[0][q][ty * 4]);
      float4 g1 = g0;
      if (NSEG == 2) g1 = *reinterpret_cast<const float4*>(&Gs[NSEG - 1][q][ty * 4]);
      const float gc0[4] = {g0.x, g0.y, g0.z, g0.w};
      const float gc1[4] = {g1.x, g1.y, g1.z, g1.w};
      float qf[NV][4];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const float4 t = *reinterpret_cast<const float4*>(&Qs[v][q][tx * 4]);
        qf[v][0] = t.x; qf[v][1] = t.y; qf[v][2] = t.z; qf[v][3] = t.w;
      }
      if (FAST) {  // d/dc of -|q - c| coefficient form: coef * sign(c - q)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float g2 = gc0[i] + gc0[i];
          csum[i] += g2;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            acc[i][j] = fmaf(g2, sat_half_sign(cr[i][j] - qf[0][j]), acc[i][j]);
        }
        continue;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float coef = (NSEG == 2 && segj[j]) ? gc1[i] : gc0[i];
          float d0, d1, d2, dc;
          pair_elem_bwd<OP>(P, a.apply_tanh, qf[0][j], NV > 1 ? qf[NV > 1 ? 1 : 0][j] : 0.f,
                            NV > 2 ? qf[NV > 2 ? 2 : 0][j] : 0.f, cr[i][j], coef, d0, d1, d2, dc);
          acc[i][j] += dc;
        }
    }
    __syncthreads();
  }

  // partial[split][c][k]  (k = coordinate; element rotation applied by the reducer)
  float* dst = partial + (int64_t)blockIdx.z * a.n_cand * W;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty * 4 + i;
    if (c >= a.n_cand) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < W) dst[(int64_t)c * W + k] = FAST ? fmaf(-0.5f, csum[i], acc[i][j]) : acc[i][j];
    }
  }
}

// Sum the split partials in order and write fp32 candidate-gradient rows.
__global__ void __launch_bounds__(256) pair_bwd_c_reduce_kernel(const float* partial, int n_split,
                                                                 int n_cand, int W, int rot,
                                                                 bess_rows_t d_cand, int add) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)n_cand * W) return;
  const int c = (int)(t / W), k = (int)(t % W);
  float s = 0.f;
  for (int i = 0; i < n_split; ++i) s += partial[(int64_t)i * n_cand * W + t];
  int e = k + rot;
  if (e >= W) e -= W;
  float* row = static_cast<float*>(const_cast<void*>(d_cand.base)) + src_row(d_cand, c) * d_cand.pitch;
  row[e] = add ? row[e] + s : s;
}

// ---------------------------------------------------------------------------
// Per-triple negatives: streaming kernels.  CTA = one query, warp = one
// candidate row at a time.
// ---------------------------------------------------------------------------
struct PerArgs {
  const float* qv;
  int n_query;
  bess_rows_t cand;
  int n_per;
  int W;
  int rot;
  int apply_tanh;
  int normalize;  // PairRE
  int vec_ok;
  bess_rowmap_t score_map;
  int64_t ld;
  int col0;
  float* out;
  float* aux;
  const float* score;
  const float* d_score;
  float* d_qv;
  bess_rows_t d_cand;
};

constexpr int PT_WARPS = 4;

template <int OP, int P, typename CT>
__global__ void __launch_bounds__(PT_WARPS * 32) pertriple_fwd_kernel(PerArgs a, int64_t q_stride) {
  constexpr int NV = OpTraits<OP>::NV;
  constexpr int NSEG = OpTraits<OP>::NSEG;
  constexpr int V = Elem<CT>::kVec;
  extern __shared__ __align__(16) float sq[];  // [NV][W]
  const int q = blockIdx.x;
  const int W = a.W;
  for (int i = threadIdx.x; i < NV * W; i += blockDim.x) sq[i] = a.qv[(int64_t)q * NV * W + i];
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int qpos = map_row(a.score_map, q);
  const int64_t orow = (int64_t)qpos * a.ld + a.col0;
  const int inv_rot = a.rot == 0 ? 0 : W - a.rot;  // coordinate of element e is (e + inv_rot) % W

  for (int c = w; c < a.n_per; c += PT_WARPS) {
    int lr = map_row(a.cand.map, c) + (int)(qpos * q_stride);
    if (a.cand.idx != nullptr) lr = __ldg(a.cand.idx + lr);
    const CT* row = static_cast<const CT*>(a.cand.base) + (int64_t)lr * a.cand.pitch;
    float scale = 1.f;
    if (OP == OP_PAIRRE && a.normalize) {
      float n2 = 0.f;
      for (int e = lane; e < W; e += 32) { const float v = ldf(row + e); n2 += v * v; }
      scale = 1.f / fmaxf(sqrtf(warp_sum(n2)), 1e-12f);
    }
    float acc0 = 0.f, acc1 = 0.f;
    for (int e0 = lane * V; e0 < W; e0 += 32 * V) {
      float vals[V];
      load_cand_vec<CT>(row, e0, W, 0, a.vec_ok, vals);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int e = e0 + i;
        if (e >= W) break;
        int k = e + inv_rot;
        if (k >= W) k -= W;
        const float t = pair_elem<OP>(P, a.apply_tanh, sq[k], NV > 1 ? sq[(NV > 1 ? W : 0) + k] : 0.f,
                                      NV > 2 ? sq[(NV > 2 ? 2 * W : 0) + k] : 0.f, vals[i] * scale);
        if (NSEG == 2 && k >= W / 2) acc1 += t; else acc0 += t;
      }
    }
    acc0 = warp_sum(acc0);
    if (NSEG == 2) acc1 = warp_sum(acc1);
    if (lane == 0) {
      float s;
      if (OP == OP_DOT) s = acc0;
      else if (NSEG == 2) {
        const float n0 = nfin(P, acc0);
        s = -(n0 + nfin(P, acc1));
        if (a.aux != nullptr) a.aux[orow + c] = n0;
      } else s = -nfin(P, acc0);
      a.out[orow + c] = s;
    }
  }
}

// Same computation for rows of up to 1024 elements that allow 128-bit loads
// (every BASELINE shape: RotatE d=512 -> 1024, PairRE d=512 -> 512): a warp
// pulls its WHOLE candidate row into registers with PT_U independent streaming
// 128-bit loads per lane, and the loads of the warp's NEXT row (and the index
// of the one after) are issued before the current row is reduced, so two rows
// per warp are in flight: the kernel is a pure HBM stream of randomly placed
// rows (measured before the prefetch: 4 KB rows 6.9 TB/s, 2 KB rows only
// 3.2 TB/s — too few bytes in flight per SM).  PT_U = ceil(W / (32 * V)) rounded
// to {1/4, 1/2, 1} of the 1024-element maximum (template parameter F) keeps the
// register count proportional to the row.  PairRE's row norm comes from the
// registers instead of a second pass over the row, and the query vectors are
// read from shared memory as 128-bit words (conflict-free) when the BoxE
// rotation keeps a lane's elements contiguous.
template <int OP, int P, typename CT, int F>
__global__ void __launch_bounds__(PT_WARPS * 32) pertriple_fwd_row_kernel(PerArgs a, int64_t q_stride) {
  constexpr int NV = OpTraits<OP>::NV;
  constexpr int NSEG = OpTraits<OP>::NSEG;
  constexpr int V = Elem<CT>::kVec;
  constexpr int PT_U = 1024 / (32 * V) / F;  // row chunks per lane
  static_assert(PT_U >= 1, "row chunk factor");
  extern __shared__ __align__(16) float sq[];  // [NV][W]
  const int q = blockIdx.x;
  const int W = a.W;
  for (int i = threadIdx.x; i < NV * W; i += blockDim.x) sq[i] = a.qv[(int64_t)q * NV * W + i];
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int qpos = map_row(a.score_map, q);
  const int64_t orow = (int64_t)qpos * a.ld + a.col0;
  const int inv_rot = a.rot == 0 ? 0 : W - a.rot;  // coordinate of element e is (e + inv_rot) % W
  const bool sq_vec = (inv_rot % V) == 0 && (W & 3) == 0;
  const int qbase = (int)(qpos * q_stride);

  auto row_index = [&](int c) -> int {
    int lr = map_row(a.cand.map, c) + qbase;
    if (a.cand.idx != nullptr) lr = __ldg(a.cand.idx + lr);
    return lr;
  };
  auto load_row = [&](int lr, uint4 (&raw)[PT_U]) {
    const CT* row = static_cast<const CT*>(a.cand.base) + (int64_t)lr * a.cand.pitch;
#pragma unroll
    for (int u = 0; u < PT_U; ++u) {
      const int e = (u * 32 + lane) * V;
      raw[u] = e < W ? ld_stream(reinterpret_cast<const uint4*>(row + e)) : make_uint4(0u, 0u, 0u, 0u);
    }
  };

  uint4 raw[PT_U], raw_next[PT_U];
  int lr_next = 0;  // storage row of candidate c + PT_WARPS (its loads go out one iteration ahead)
  if (w < a.n_per) load_row(row_index(w), raw);
  if (w + PT_WARPS < a.n_per) lr_next = row_index(w + PT_WARPS);

  // Prefetching pays for rows of <= 2 KB (3.2 -> 4.0 TB/s measured); 4 KB rows already keep
  // enough bytes in flight and lose occupancy to the second register set (6.9 -> 5.3 TB/s).
  constexpr bool PREFETCH = F >= 2;
  for (int c = w; c < a.n_per; c += PT_WARPS) {
    const bool has_next = c + PT_WARPS < a.n_per;
    if (PREFETCH) {
      if (has_next) load_row(lr_next, raw_next);
      if (c + 2 * PT_WARPS < a.n_per) lr_next = row_index(c + 2 * PT_WARPS);
    }

    float vals[PT_U][V];
#pragma unroll
    for (int u = 0; u < PT_U; ++u) Elem<CT>::unpack(raw[u], vals[u]);
    float scale = 1.f;
    if (OP == OP_PAIRRE && a.normalize) {
      float n2 = 0.f;
#pragma unroll
      for (int u = 0; u < PT_U; ++u)
#pragma unroll
        for (int i = 0; i < V; ++i) n2 += vals[u][i] * vals[u][i];  // chunks beyond W are zero
      scale = 1.f / fmaxf(sqrtf(warp_sum(n2)), 1e-12f);
    }
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
    for (int u = 0; u < PT_U; ++u) {
      const int e0 = (u * 32 + lane) * V;
      if (e0 < W) {
        int k0 = e0 + inv_rot;
        if (k0 >= W) k0 -= W;
        float qa[NV][V];
        if (sq_vec) {
#pragma unroll
          for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int i = 0; i < V; i += 4) {
              const float4 t = *reinterpret_cast<const float4*>(&sq[v * W + k0 + i]);
              qa[v][i] = t.x; qa[v][i + 1] = t.y; qa[v][i + 2] = t.z; qa[v][i + 3] = t.w;
            }
        } else {
#pragma unroll
          for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int i = 0; i < V; ++i) {
              int k = k0 + i;
              if (k >= W) k -= W;
              qa[v][i] = sq[v * W + k];
            }
        }
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float t = pair_elem<OP>(P, a.apply_tanh, qa[0][i], NV > 1 ? qa[NV > 1 ? 1 : 0][i] : 0.f,
                                        NV > 2 ? qa[NV > 2 ? 2 : 0][i] : 0.f, vals[u][i] * scale);
          if (NSEG == 2) {
            int k = k0 + i;
            if (k >= W) k -= W;
            if (k >= W / 2) acc1 += t; else acc0 += t;
          } else {
            acc0 += t;
          }
        }
      }
    }
    acc0 = warp_sum(acc0);
    if (NSEG == 2) acc1 = warp_sum(acc1);
    if (lane == 0) {
      float s;
      if (OP == OP_DOT) s = acc0;
      else if (NSEG == 2) {
        const float n0 = nfin(P, acc0);
        s = -(n0 + nfin(P, acc1));
        if (a.aux != nullptr) a.aux[orow + c] = n0;
      } else s = -nfin(P, acc0);
      a.out[orow + c] = s;
    }
    if (has_next) {
      if (PREFETCH) {
#pragma unroll
        for (int u = 0; u < PT_U; ++u) raw[u] = raw_next[u];
      } else {
        load_row(row_index(c + PT_WARPS), raw);
      }
    }
  }
}

template <int OP, int P, typename CT>
__global__ void __launch_bounds__(PT_WARPS * 32) pertriple_bwd_kernel(PerArgs a, int64_t q_stride) {
  constexpr int NV = OpTraits<OP>::NV;
  constexpr int NSEG = OpTraits<OP>::NSEG;
  extern __shared__ __align__(16) float sm[];  // sq [NV*W] | dq [PT_WARPS][NV*W]
  const int q = blockIdx.x;
  const int W = a.W, NW = NV * W;
  float* sq = sm;
  float* dq = sm + NW;
  for (int i = threadIdx.x; i < NW; i += blockDim.x) sq[i] = a.qv[(int64_t)q * NW + i];
  for (int i = threadIdx.x; i < PT_WARPS * NW; i += blockDim.x) dq[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* mydq = dq + w * NW;
  const int qpos = map_row(a.score_map, q);
  const int64_t orow = (int64_t)qpos * a.ld + a.col0;
  const int inv_rot = a.rot == 0 ? 0 : W - a.rot;

  for (int c = w; c < a.n_per; c += PT_WARPS) {
    const int lrow = map_row(a.cand.map, c) + (int)(qpos * q_stride);
    int lr = lrow;
    if (a.cand.idx != nullptr) lr = __ldg(a.cand.idx + lr);
    const CT* row = static_cast<const CT*>(a.cand.base) + (int64_t)lr * a.cand.pitch;
    int gr = map_row(a.d_cand.map, c) + (int)(qpos * q_stride);
    if (a.d_cand.idx != nullptr) gr = __ldg(a.d_cand.idx + gr);
    float* grow = static_cast<float*>(const_cast<void*>(a.d_cand.base)) + (int64_t)gr * a.d_cand.pitch;

    float scale = 1.f, nrm = 1.f;
    if (OP == OP_PAIRRE && a.normalize) {
      float n2 = 0.f;
      for (int e = lane; e < W; e += 32) { const float v = ldf(row + e); n2 += v * v; }
      nrm = sqrtf(warp_sum(n2));
      scale = 1.f / fmaxf(nrm, 1e-12f);
    }
    const float g = a.d_score[orow + c];
    const float sc = (P == 2 && OP != OP_DOT) ? a.score[orow + c] : 0.f;
    const float ax = (P == 2 && OP == OP_BOXE) ? a.aux[orow + c] : 0.f;
    const float cf0 = pair_coef<OP, P>(g, sc, ax, 0);
    const float cf1 = NSEG == 2 ? pair_coef<OP, P>(g, sc, ax, 1) : cf0;
    float proj = 0.f;
    for (int e = lane; e < W; e += 32) {
      int k = e + inv_rot;
      if (k >= W) k -= W;
      const float cv = ldf(row + e) * scale;
      float d0, d1, d2, dc;
      pair_elem_bwd<OP>(P, a.apply_tanh, sq[k], NV > 1 ? sq[(NV > 1 ? W : 0) + k] : 0.f,
                        NV > 2 ? sq[(NV > 2 ? 2 * W : 0) + k] : 0.f, cv,
                        (NSEG == 2 && k >= W / 2) ? cf1 : cf0, d0, d1, d2, dc);
      mydq[k] += d0;
      if (NV > 1) mydq[(NV > 1 ? W : 0) + k] += d1;
      if (NV > 2) mydq[(NV > 2 ? 2 * W : 0) + k] += d2;
      grow[e] = dc;
      proj += cv * dc;
    }
    if (OP == OP_PAIRRE && a.normalize) {
      proj = warp_sum(proj);
      for (int e = lane; e < W; e += 32) {
        const float ch = ldf(row + e) * scale;
        grow[e] = nrm > 1e-12f ? (grow[e] - ch * proj) * scale : grow[e] * scale;
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NW; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int ww = 0; ww < PT_WARPS; ++ww) s += dq[ww * NW + i];
    a.d_qv[(int64_t)q * NW + i] = s;
  }
}

}  // namespace bess

using namespace bess;

static inline int elem_size(int dtype) { return dtype == BESS_F32 ? 4 : 2; }

static bool cand_vec_ok(const bess_rows_t& cand, int dtype, int W, int rot) {
  const int es = elem_size(dtype);
  const int V = 16 / es;
  return ((uintptr_t)cand.base % 16 == 0) && ((cand.pitch * es) % 16 == 0) && (W % V == 0) &&
         (rot % V == 0);
}

static int check_pair(const bess_score_cfg_t* cfg, int mode, FamCfg& f, int& op, int& rot) {
  BESS_CHECK_ARG(cfg != nullptr, "null score config");
  BESS_CHECK_ARG(mode == BESS_MODE_TAILS || mode == BESS_MODE_HEADS, "bad mode %d", mode);
  f = to_cfg(cfg);
  op = pair_op(f);
  if (op != OP_DOT)
    BESS_CHECK_ARG(f.norm_p == 1 || f.norm_p == 2, "scoring_norm %d not supported", f.norm_p);
  if (op == OP_BOXE)
    BESS_CHECK_ARG(f.per_dim, "BoxE dist_func_per_dim=False is not supported for negative scoring");
  // BoxE coordinate k pairs with candidate element (k + rot) % W (families.cuh prologue_fwd)
  rot = (op == OP_BOXE && mode == BESS_MODE_TAILS) ? f.d : 0;
  return BESS_OK;
}

// dispatch over (op, p, dtype); the statement sees OP, P and CT
#define PAIR_DISPATCH_T(OPV, PV, dtype, ...)                                                   \
  switch (dtype) {                                                                             \
    case BESS_F32: { using CT = float; constexpr int OP = OPV; constexpr int P = PV; __VA_ARGS__; break; } \
    case BESS_F16: { using CT = __half; constexpr int OP = OPV; constexpr int P = PV; __VA_ARGS__; break; } \
    case BESS_BF16: { using CT = __nv_bfloat16; constexpr int OP = OPV; constexpr int P = PV; __VA_ARGS__; break; } \
    default: bess_set_error("unknown dtype %d", dtype); return BESS_ERR_INVALID_ARG;          \
  }
#define PAIR_DISPATCH(op, p, dtype, ...)                                                       \
  do {                                                                                         \
    if (op == OP_DOT) { PAIR_DISPATCH_T(OP_DOT, 1, dtype, __VA_ARGS__) }                       \
    else if (op == OP_DIST && p == 1) { PAIR_DISPATCH_T(OP_DIST, 1, dtype, __VA_ARGS__) }      \
    else if (op == OP_DIST) { PAIR_DISPATCH_T(OP_DIST, 2, dtype, __VA_ARGS__) }                \
    else if (op == OP_PAIRRE && p == 1) { PAIR_DISPATCH_T(OP_PAIRRE, 1, dtype, __VA_ARGS__) }  \
    else if (op == OP_PAIRRE) { PAIR_DISPATCH_T(OP_PAIRRE, 2, dtype, __VA_ARGS__) }            \
    else if (p == 1) { PAIR_DISPATCH_T(OP_BOXE, 1, dtype, __VA_ARGS__) }                       \
    else { PAIR_DISPATCH_T(OP_BOXE, 2, dtype, __VA_ARGS__) }                                   \
  } while (0)

static PairArgs make_pair_args(const FamCfg& f, int dtype, int rot, const float* qv, int n_query,
                               bess_rows_t cand, const float* cand_scale, int n_cand,
                               bess_rowmap_t score_map, int64_t ld, int col0) {
  PairArgs a;
  a.qv = qv; a.n_query = n_query; a.cand = cand; a.cand_scale = cand_scale; a.n_cand = n_cand;
  a.W = ent_width(f); a.rot = rot; a.apply_tanh = f.apply_tanh;
  a.vec_ok = cand_vec_ok(cand, dtype, a.W, rot) ? 1 : 0;
  a.score_map = score_map; a.ld = ld; a.col0 = col0;
  a.out = nullptr; a.aux = nullptr; a.score = nullptr; a.d_score = nullptr;
  return a;
}

// InterHT / TranS (two candidate elements per coordinate): kernels in pair2.cu
namespace bess {
int pair2_shared_fwd(const bess_score_cfg_t*, int, const float*, int, bess_rows_t, const float*, int, float*,
                     bess_rowmap_t, int64_t, int, cudaStream_t);
int pair2_shared_bwd_query(const bess_score_cfg_t*, int, const float*, int, bess_rows_t, const float*, int,
                           const float*, const float*, bess_rowmap_t, int64_t, int, float*, cudaStream_t);
int64_t pair2_bwd_cand_workspace(const bess_score_cfg_t*, int, int);
int pair2_shared_bwd_cand(const bess_score_cfg_t*, int, const float*, int, bess_rows_t, const float*, int,
                          const float*, const float*, bess_rowmap_t, int64_t, int, bess_rows_t, int, void*,
                          cudaStream_t);
int pair2_pertriple_fwd(const bess_score_cfg_t*, int, const float*, int, bess_rows_t, int64_t, int, float*,
                        bess_rowmap_t, int64_t, int, cudaStream_t);
int pair2_pertriple_bwd(const bess_score_cfg_t*, int, const float*, int, bess_rows_t, int64_t, int,
                        const float*, const float*, bess_rowmap_t, int64_t, int, float*, bess_rows_t,
                        cudaStream_t);
}  // namespace bess
static int check_pair2(const bess_score_cfg_t* cfg) {
  BESS_CHECK_ARG(cfg->d > 0 && (cfg->norm_p == 1 || cfg->norm_p == 2),
                 "InterHT / TranS need embedding_size > 0 and scoring_norm 1 or 2");
  return BESS_OK;
}

extern "C" int bess_score_shared_fwd(const bess_score_cfg_t* cfg, int dtype, int mode,
                                     const float* qv, int n_query, bess_rows_t cand,
                                     const float* cand_scale, int n_cand, float* out,
                                     bess_rowmap_t score_map, int64_t ld_out, int col0, float* aux,
                                     void* stream) {
  if (is_pair2(cfg->family)) {
    if (int e = check_pair2(cfg)) return e;
    return pair2_shared_fwd(cfg, dtype, qv, n_query, cand, cand_scale, n_cand, out, score_map, ld_out, col0,
                            (cudaStream_t)stream);
  }
  FamCfg f; int op, rot;
  if (int e = check_pair(cfg, mode, f, op, rot)) return e;
  if (n_query == 0 || n_cand == 0) return BESS_OK;
  if (op == OP_BOXE)
    BESS_CHECK_ARG(f.d % F_KC == 0, "BoxE negative scoring needs embedding_size %% %d == 0", F_KC);
  if (op == OP_BOXE && f.norm_p == 2) BESS_CHECK_ARG(aux != nullptr, "BoxE p=2 needs the aux buffer");
  PairArgs a = make_pair_args(f, dtype, rot, qv, n_query, cand, cand_scale, n_cand, score_map, ld_out, col0);
  a.out = out; a.aux = aux;
  // rows per warp (MR): the choice that leaves the busiest SM with the least work
  const int nct = ceil_div(n_cand, F_TC);
  int mr = 8;
  int64_t best = -1;
  for (int m = 8; m >= 6; --m) {
    const int64_t tiles = (int64_t)ceil_div(n_query, 16 * m) * nct;
    const int64_t cost = ceil_div(tiles, (int64_t)kNumSM) * m;
    if (best < 0 || cost < best) { best = cost; mr = m; }
  }
  dim3 grid(nct, ceil_div(n_query, 16 * mr));
  cudaStream_t st = (cudaStream_t)stream;
  if (mr == 8) { PAIR_DISPATCH(op, f.norm_p, dtype, pair_fwd_kernel<OP, P, CT, 8><<<grid, F_NT, 0, st>>>(a)); }
  else if (mr == 7) { PAIR_DISPATCH(op, f.norm_p, dtype, pair_fwd_kernel<OP, P, CT, 7><<<grid, F_NT, 0, st>>>(a)); }
  else { PAIR_DISPATCH(op, f.norm_p, dtype, pair_fwd_kernel<OP, P, CT, 6><<<grid, F_NT, 0, st>>>(a)); }
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

extern "C" int bess_score_shared_bwd_query(const bess_score_cfg_t* cfg, int dtype, int mode,
                                           const float* qv, int n_query, bess_rows_t cand,
                                           const float* cand_scale, int n_cand, const float* score,
                                           const float* d_score, bess_rowmap_t score_map,
                                           int64_t ld, int col0, const float* aux, float* d_qv,
                                           void* stream) {
  if (is_pair2(cfg->family)) {
    if (int e = check_pair2(cfg)) return e;
    return pair2_shared_bwd_query(cfg, dtype, qv, n_query, cand, cand_scale, n_cand, score, d_score,
                                  score_map, ld, col0, d_qv, (cudaStream_t)stream);
  }
  FamCfg f; int op, rot;
  if (int e = check_pair(cfg, mode, f, op, rot)) return e;
  if (n_query == 0) return BESS_OK;
  PairArgs a = make_pair_args(f, dtype, rot, qv, n_query, cand, cand_scale, n_cand, score_map, ld, col0);
  a.score = score; a.d_score = d_score; a.aux = const_cast<float*>(aux);
  // 64- or 32-query tiles: whichever leaves the busiest SM with less work (every CTA streams all
  // candidates, so the grid is fixed by the shape and quantises against the 148 SMs)
  const int kt = ceil_div(a.W, B_TK);
  const int64_t t64 = (int64_t)ceil_div(n_query, 64) * kt, t32 = (int64_t)ceil_div(n_query, 32) * kt;
  const int64_t cost64 = 2 * ceil_div(t64, (int64_t)kNumSM), cost32 = ceil_div(t32, (int64_t)kNumSM);
  if (cost32 < cost64) {
    dim3 grid(ceil_div(n_query, 32), kt);
    PAIR_DISPATCH(op, f.norm_p, dtype,
                  pair_bwd_q_kernel<OP, P, CT, 128><<<grid, 128, 0, (cudaStream_t)stream>>>(a, d_qv));
  } else {
    dim3 grid(ceil_div(n_query, 64), kt);
    PAIR_DISPATCH(op, f.norm_p, dtype,
                  pair_bwd_q_kernel<OP, P, CT, 256><<<grid, 256, 0, (cudaStream_t)stream>>>(a, d_qv));
  }
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

static int choose_split(int n_query, int n_cand, int W) {
  const int tiles = ceil_div(n_cand, B_T) * ceil_div(W, B_TK);
  int split = ceil_div(4 * kNumSM, tiles);
  const int max_split = ceil_div(n_query, 4 * B_CH);  // >= 128 queries per split
  if (split > max_split) split = max_split;
  if (split < 1) split = 1;
  if (split > 64) split = 64;
  return split;
}

extern "C" int64_t bess_shared_bwd_cand_workspace(const bess_score_cfg_t* cfg, int n_query,
                                                  int n_cand) {
  if (is_pair2(cfg->family)) return pair2_bwd_cand_workspace(cfg, n_query, n_cand);
  const FamCfg f = to_cfg(cfg);
  const int W = ent_width(f);
  return (int64_t)choose_split(n_query, n_cand, W) * n_cand * W * sizeof(float);
}

extern "C" int bess_score_shared_bwd_cand(const bess_score_cfg_t* cfg, int dtype, int mode,
                                          const float* qv, int n_query, bess_rows_t cand,
                                          const float* cand_scale, int n_cand, const float* score,
                                          const float* d_score, bess_rowmap_t score_map,
                                          int64_t ld, int col0, const float* aux,
                                          bess_rows_t d_cand, int add_cand, void* workspace,
                                          void* stream) {
  if (is_pair2(cfg->family)) {
    if (int e = check_pair2(cfg)) return e;
    return pair2_shared_bwd_cand(cfg, dtype, qv, n_query, cand, cand_scale, n_cand, score, d_score, score_map,
                                 ld, col0, d_cand, add_cand, workspace, (cudaStream_t)stream);
  }
  FamCfg f; int op, rot;
  if (int e = check_pair(cfg, mode, f, op, rot)) return e;
  if (n_cand == 0) return BESS_OK;
  BESS_CHECK_ARG(workspace != nullptr, "workspace required");
  PairArgs a = make_pair_args(f, dtype, rot, qv, n_query, cand, cand_scale, n_cand, score_map, ld, col0);
  a.score = score; a.d_score = d_score; a.aux = const_cast<float*>(aux);
  const int split = choose_split(n_query, n_cand, a.W);
  int q_per_split = ceil_div(n_query, split);
  q_per_split = ceil_div(q_per_split, B_CH) * B_CH;
  if (q_per_split < B_CH) q_per_split = B_CH;
  float* partial = static_cast<float*>(workspace);
  dim3 grid(ceil_div(n_cand, B_T), ceil_div(a.W, B_TK), split);
  PAIR_DISPATCH(op, f.norm_p, dtype,
                pair_bwd_c_kernel<OP, P, CT><<<grid, 256, 0, (cudaStream_t)stream>>>(a, partial, q_per_split));
  BESS_CHECK_LAUNCH();
  const int64_t total = (int64_t)n_cand * a.W;
  pair_bwd_c_reduce_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
      partial, split, n_cand, a.W, rot, d_cand, add_cand);
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

static PerArgs make_per_args(const FamCfg& f, int dtype, int rot, const float* qv, int n_query,
                             bess_rows_t cand, int n_per, bess_rowmap_t score_map, int64_t ld,
                             int col0) {
  PerArgs a;
  a.qv = qv; a.n_query = n_query; a.cand = cand; a.n_per = n_per; a.W = ent_width(f); a.rot = rot;
  a.apply_tanh = f.apply_tanh; a.normalize = (f.family == FAM_PAIRRE || f.family == FAM_TRIPLERE) ? f.normalize : 0;
  a.vec_ok = cand_vec_ok(cand, dtype, a.W, 0) ? 1 : 0;
  a.score_map = score_map; a.ld = ld; a.col0 = col0;
  a.out = nullptr; a.aux = nullptr; a.score = nullptr; a.d_score = nullptr; a.d_qv = nullptr;
  a.d_cand = cand;
  return a;
}

extern "C" int bess_score_pertriple_fwd(const bess_score_cfg_t* cfg, int dtype, int mode,
                                        const float* qv, int n_query, bess_rows_t cand,
                                        int64_t cand_q_stride, int n_per, float* out,
                                        bess_rowmap_t score_map, int64_t ld_out, int col0,
                                        float* aux, void* stream) {
  if (is_pair2(cfg->family)) {
    if (int e = check_pair2(cfg)) return e;
    return pair2_pertriple_fwd(cfg, dtype, qv, n_query, cand, cand_q_stride, n_per, out, score_map, ld_out,
                               col0, (cudaStream_t)stream);
  }
  FamCfg f; int op, rot;
  if (int e = check_pair(cfg, mode, f, op, rot)) return e;
  if (n_query == 0 || n_per == 0) return BESS_OK;
  if (op == OP_BOXE && f.norm_p == 2) BESS_CHECK_ARG(aux != nullptr, "BoxE p=2 needs the aux buffer");
  PerArgs a = make_per_args(f, dtype, rot, qv, n_query, cand, n_per, score_map, ld_out, col0);
  a.out = out; a.aux = aux;
  const size_t smem = (size_t)pair_nvec(f) * a.W * sizeof(float);
  BESS_CHECK_ARG(smem <= 48 * 1024, "row too wide for the per-triple kernel");
  // whole-row-in-registers variant: 128-bit loads possible and the row has <= 1024 elements
  static const bool force_v1 = [] { const char* e = getenv("BESS_PERTRIPLE_V1"); return e && e[0] == '1'; }();
  const int vec_elems = dtype == BESS_F32 ? 4 : 8;
  const bool row_regs = !force_v1 && a.vec_ok && a.W % vec_elems == 0 && a.W <= 1024;
  // row chunk factor: registers for a quarter / half / all of the 1024-element maximum
  const int fct = a.W <= 256 ? 4 : (a.W <= 512 ? 2 : 1);
  PAIR_DISPATCH(op, f.norm_p, dtype, {
    if (row_regs && fct == 4)
      pertriple_fwd_row_kernel<OP, P, CT, 4><<<n_query, PT_WARPS * 32, smem, (cudaStream_t)stream>>>(a, cand_q_stride);
    else if (row_regs && fct == 2)
      pertriple_fwd_row_kernel<OP, P, CT, 2><<<n_query, PT_WARPS * 32, smem, (cudaStream_t)stream>>>(a, cand_q_stride);
    else if (row_regs)
      pertriple_fwd_row_kernel<OP, P, CT, 1><<<n_query, PT_WARPS * 32, smem, (cudaStream_t)stream>>>(a, cand_q_stride);
    else
      pertriple_fwd_kernel<OP, P, CT><<<n_query, PT_WARPS * 32, smem, (cudaStream_t)stream>>>(a, cand_q_stride);
  });
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

extern "C" int bess_score_pertriple_bwd(const bess_score_cfg_t* cfg, int dtype, int mode,
                                        const float* qv, int n_query, bess_rows_t cand,
                                        int64_t cand_q_stride, int n_per, const float* score,
                                        const float* d_score, bess_rowmap_t score_map, int64_t ld,
                                        int col0, const float* aux, float* d_qv,
                                        bess_rows_t d_cand, void* stream) {
  if (is_pair2(cfg->family)) {
    if (int e = check_pair2(cfg)) return e;
    return pair2_pertriple_bwd(cfg, dtype, qv, n_query, cand, cand_q_stride, n_per, score, d_score, score_map,
                               ld, col0, d_qv, d_cand, (cudaStream_t)stream);
  }
  FamCfg f; int op, rot;
  if (int e = check_pair(cfg, mode, f, op, rot)) return e;
  if (n_query == 0) return BESS_OK;
  PerArgs a = make_per_args(f, dtype, rot, qv, n_query, cand, n_per, score_map, ld, col0);
  a.score = score; a.d_score = d_score; a.aux = const_cast<float*>(aux); a.d_qv = d_qv;
  a.d_cand = d_cand;
  const size_t smem = (size_t)pair_nvec(f) * a.W * sizeof(float) * (1 + PT_WARPS);
  BESS_CHECK_ARG(smem <= 200 * 1024, "row too wide for the per-triple backward kernel");
  PAIR_DISPATCH(op, f.norm_p, dtype, {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(pertriple_bwd_kernel<OP, P, CT>,
                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    pertriple_bwd_kernel<OP, P, CT><<<n_query, PT_WARPS * 32, smem, (cudaStream_t)stream>>>(a, cand_q_stride);
  });
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}
