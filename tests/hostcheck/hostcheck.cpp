// TEST INFRASTRUCTURE: compiles besskge_b200/csrc/families.cuh for the HOST so the
// score-function math and the hand-derived gradients used by the CUDA kernels
// can be checked against the oracle on a machine without a GPU.  Never linked
// into the product library.
#include <cstdint>
#include <vector>
#include "../../besskge_b200/csrc/families.cuh"

using namespace bess;

// TripleRE's v2 offset u travels in a global set by hc_set_rel_u (keeps every signature unchanged)
static float g_rel_u = 0.f;
extern "C" void hc_set_rel_u(float u) { g_rel_u = u; }
static FamCfg mk(int family, int p, int d, int normalize, int apply_tanh, int per_dim, float eps) {
  FamCfg c; c.family = family; c.norm_p = p; c.d = d; c.normalize = normalize;
  c.apply_tanh = apply_tanh; c.per_dim = per_dim; c.eps = eps; c.rel_u = g_rel_u; return c;
}

extern "C" {

void hc_triple_fwd(int family, int p, int d, int normalize, int apply_tanh, int per_dim, float eps,
                   int n, const float* h, const float* rel_table, const int* r, const float* t,
                   float* out) {
  FamCfg c = mk(family, p, d, normalize, apply_tanh, per_dim, eps);
  const int W = ent_width(c), Wr = rel_width(c);
  for (int i = 0; i < n; ++i)
    out[i] = triple_fwd<HostCtx, float>(c, h + (int64_t)i * W, rel_table + (int64_t)r[i] * Wr, t + (int64_t)i * W);
}

void hc_triple_bwd(int family, int p, int d, int normalize, int apply_tanh, int per_dim, float eps,
                   int n, const float* h, const float* rel_table, const int* r, const float* t,
                   const float* g, float* dh, float* dr_rows, float* dt) {
  FamCfg c = mk(family, p, d, normalize, apply_tanh, per_dim, eps);
  const int W = ent_width(c), Wr = rel_width(c);
  for (int i = 0; i < n; ++i) {
    const float* hh = h + (int64_t)i * W; const float* tt = t + (int64_t)i * W;
    const float* rr = rel_table + (int64_t)r[i] * Wr;
    const float s = triple_fwd<HostCtx, float>(c, hh, rr, tt);
    triple_bwd<HostCtx, float>(c, hh, rr, tt, s, g[i], dh + (int64_t)i * W, dr_rows + (int64_t)i * Wr,
                               dt + (int64_t)i * W, 0, 0, 0);
    if (family == FAM_BOXE) boxe_rel_finalize<HostCtx, float>(c, rr, dr_rows + (int64_t)i * Wr);
  }
}

// scores of n queries against candidates: shared (cand [nc, W]) or per query (cand [n, nc, W])
static float pair_score(const FamCfg& c, int op, int rot, const float* qv, const float* cand, float scale,
                        float* seg0_out) {
  const int W = ent_width(c), p = c.norm_p;
  const int nseg = pair_nseg(op);
  float acc[2] = {0.f, 0.f};
  for (int k = 0; k < W; ++k) {
    int e = k + rot; if (e >= W) e -= W;
    const float cv = cand[e] * scale;
    float t;
    switch (op) {
      case OP_DOT: t = pair_elem<OP_DOT>(p, c.apply_tanh, qv[k], 0.f, 0.f, cv); break;
      case OP_DIST: t = pair_elem<OP_DIST>(p, c.apply_tanh, qv[k], 0.f, 0.f, cv); break;
      case OP_PAIRRE: t = pair_elem<OP_PAIRRE>(p, c.apply_tanh, qv[k], qv[W + k], 0.f, cv); break;
      default: t = pair_elem<OP_BOXE>(p, c.apply_tanh, qv[k], qv[W + k], qv[2 * W + k], cv); break;
    }
    acc[(nseg == 2 && k >= W / 2) ? 1 : 0] += t;
  }
  if (op == OP_DOT) return acc[0];
  if (nseg == 2) { const float n0 = nfin(p, acc[0]); if (seg0_out) *seg0_out = n0; return -(n0 + nfin(p, acc[1])); }
  return -nfin(p, acc[0]);
}

void hc_candidates_fwd(int family, int p, int d, int normalize, int apply_tanh, int per_dim, float eps,
                       int mode, int n, const float* fixed, const float* rel_table, const int* r,
                       const float* cand, int nc, int shared, float* out) {
  FamCfg c = mk(family, p, d, normalize, apply_tanh, per_dim, eps);
  const int W = ent_width(c), Wr = rel_width(c), NV = pair_nvec(c), op = pair_op(c);
  const int rot = (op == OP_BOXE && mode == MODE_TAILS) ? d : 0;
  std::vector<float> qv((size_t)NV * W);
  for (int i = 0; i < n; ++i) {
    prologue_fwd<HostCtx, float>(c, mode, fixed + (int64_t)i * W, rel_table + (int64_t)r[i] * Wr, qv.data());
    for (int j = 0; j < nc; ++j) {
      const float* cr = shared ? cand + (int64_t)j * W : cand + ((int64_t)i * nc + j) * W;
      if (op == OP_PAIR2) {  // host restatement of csrc/pair2.cu
        const Pair2Norms nn = pair2_norms<HostCtx, float>(c, cr);
        float acc = 0.f;
        for (int k = 0; k < d; ++k)
          acc += nacc(p, qv[k] * cr[k] * nn.im + qv[d + k] * cr[d + k] * nn.ia + qv[W + k]);
        out[(int64_t)i * nc + j] = -nfin(p, acc);
        continue;
      }
      float scale = 1.f;
      if (op == OP_PAIRRE && normalize) {
        float a = 0.f; for (int k = 0; k < W; ++k) a += cr[k] * cr[k];
        scale = 1.f / fmaxf(sqrtf(a), 1e-12f);
      }
      out[(int64_t)i * nc + j] = pair_score(c, op, rot, qv.data(), cr, scale, nullptr);
    }
  }
}

// full backward of the candidate scoring given g [n, nc]: d_fixed [n, W], d_rel rows [n, Wr],
// d_cand (shared: [nc, W] accumulated over queries; else [n, nc, W])
void hc_candidates_bwd(int family, int p, int d, int normalize, int apply_tanh, int per_dim, float eps,
                       int mode, int n, const float* fixed, const float* rel_table, const int* r,
                       const float* cand, int nc, int shared, const float* g, float* d_fixed,
                       float* d_rel_rows, float* d_cand) {
  FamCfg c = mk(family, p, d, normalize, apply_tanh, per_dim, eps);
  const int W = ent_width(c), Wr = rel_width(c), NV = pair_nvec(c), op = pair_op(c);
  const int rot = (op == OP_BOXE && mode == MODE_TAILS) ? d : 0;
  std::vector<float> qv((size_t)NV * W), dqv((size_t)NV * W), dch(W);
  const int64_t ncand_rows = shared ? nc : (int64_t)n * nc;
  for (int64_t i = 0; i < ncand_rows * W; ++i) d_cand[i] = 0.f;
  for (int i = 0; i < n; ++i) {
    const float* x = fixed + (int64_t)i * W; const float* rr = rel_table + (int64_t)r[i] * Wr;
    prologue_fwd<HostCtx, float>(c, mode, x, rr, qv.data());
    for (auto& v : dqv) v = 0.f;
    for (int j = 0; j < nc; ++j) {
      const float* cr = shared ? cand + (int64_t)j * W : cand + ((int64_t)i * nc + j) * W;
      float* dc = shared ? d_cand + (int64_t)j * W : d_cand + ((int64_t)i * nc + j) * W;
      if (op == OP_PAIR2) {
        const Pair2Norms nn = pair2_norms<HostCtx, float>(c, cr);
        float acc = 0.f;
        for (int k = 0; k < d; ++k)
          acc += nacc(p, qv[k] * cr[k] * nn.im + qv[d + k] * cr[d + k] * nn.ia + qv[W + k]);
        const float sc2 = -nfin(p, acc), gg2 = g[(int64_t)i * nc + j];
        const float cf = p == 1 ? -gg2 : (sc2 != 0.f ? gg2 / sc2 : 0.f);
        float pm = 0.f, pa = 0.f;
        for (int k = 0; k < d; ++k) {
          const float cm = cr[k] * nn.im, ca = cr[d + k] * nn.ia;
          const float e = qv[k] * cm + qv[d + k] * ca + qv[W + k];
          const float de = p == 1 ? cf * fsign(e) : cf * e;
          dqv[k] += de * cm; dqv[d + k] += de * ca; dqv[W + k] += de;
          pm += cm * de * qv[k]; pa += ca * de * qv[d + k];
        }
        for (int k = 0; k < d; ++k) {
          const float cm = cr[k] * nn.im, ca = cr[d + k] * nn.ia;
          const float e = qv[k] * cm + qv[d + k] * ca + qv[W + k];
          const float de = p == 1 ? cf * fsign(e) : cf * e;
          dc[k] += unnorm_grad(c.normalize, de * qv[k], cm, pm, nn.nm, nn.im);
          dc[d + k] += unnorm_grad(c.normalize, de * qv[d + k], ca, pa, nn.na, nn.ia);
        }
        continue;
      }
      float scale = 1.f, nrm = 1.f;
      if (op == OP_PAIRRE && normalize) {
        float a = 0.f; for (int k = 0; k < W; ++k) a += cr[k] * cr[k];
        nrm = sqrtf(a); scale = 1.f / fmaxf(nrm, 1e-12f);
      }
      float seg0 = 0.f;
      const float sc = pair_score(c, op, rot, qv.data(), cr, scale, &seg0);
      const float gg = g[(int64_t)i * nc + j];
      float coef[2];
      for (int s = 0; s < 2; ++s) {
        if (op == OP_DOT) coef[s] = gg;
        else if (p == 1) coef[s] = -gg;
        else {
          float nv = op == OP_BOXE ? (s == 0 ? seg0 : (-sc - seg0)) : -sc;
          coef[s] = nv > 0.f ? -gg / nv : 0.f;
        }
      }
      float proj = 0.f;
      for (int k = 0; k < W; ++k) {
        int e = k + rot; if (e >= W) e -= W;
        const float cv = cr[e] * scale;
        const float cf = coef[(pair_nseg(op) == 2 && k >= W / 2) ? 1 : 0];
        float d0 = 0, d1 = 0, d2 = 0, dcv = 0;
        switch (op) {
          case OP_DOT: pair_elem_bwd<OP_DOT>(p, c.apply_tanh, qv[k], 0.f, 0.f, cv, cf, d0, d1, d2, dcv); break;
          case OP_DIST: pair_elem_bwd<OP_DIST>(p, c.apply_tanh, qv[k], 0.f, 0.f, cv, cf, d0, d1, d2, dcv); break;
          case OP_PAIRRE: pair_elem_bwd<OP_PAIRRE>(p, c.apply_tanh, qv[k], qv[W + k], 0.f, cv, cf, d0, d1, d2, dcv); break;
          default: pair_elem_bwd<OP_BOXE>(p, c.apply_tanh, qv[k], qv[W + k], qv[2 * W + k], cv, cf, d0, d1, d2, dcv); break;
        }
        dqv[k] += d0;
        if (NV > 1) dqv[W + k] += d1;
        if (NV > 2) dqv[2 * W + k] += d2;
        dch[e] = dcv;
        proj += cv * dcv;
      }
      for (int e = 0; e < W; ++e) {
        float v = dch[e];
        if (op == OP_PAIRRE && normalize) v = nrm > 1e-12f ? (v - cr[e] * scale * proj) * scale : v * scale;
        dc[e] += v;
      }
    }
    prologue_bwd<HostCtx, float>(c, mode, x, rr, dqv.data(), d_fixed + (int64_t)i * W,
                                 d_rel_rows + (int64_t)i * Wr, 0, 0);
    if (family == FAM_BOXE) boxe_rel_finalize<HostCtx, float>(c, rr, d_rel_rows + (int64_t)i * Wr);
  }
}

}  // extern "C"
