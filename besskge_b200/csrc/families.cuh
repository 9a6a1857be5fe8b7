// Score-function family math (TransE, RotatE, DistMult, ComplEx, PairRE, BoxE).
//
// Written as host+device row functions parameterised by a reduction context so
// the SAME arithmetic is (a) executed by one warp per row inside the kernels
// and (b) compiled for the host by tests/hostcheck to verify the formulas and
// the hand-derived gradients against the oracle without a GPU.
//
// Reference semantics (file:line are /root/reference/besskge/):
//   TransE    scoring.py:321-354   -||h + r - t||_p
//   RotatE    scoring.py:423-462   -||h o e^{i r} - t||_p over the 2d real vector
//                                  (utils.py:92-112: cos/sin of r, NO pi factor)
//   DistMult  scoring.py:804-837   sum h*r*t
//   ComplEx   scoring.py:905-946   sum (h (x) r) * t, halves = re | im
//   PairRE    scoring.py:540-593   -||h^ o r_h - t^ o r_t||_p, h^ = h/max(||h||,1e-12)
//   BoxE      scoring.py:1250-1415 box distance, see boxe_* below
//   InterHT   scoring.py:1418-1572 -||h^ (t~^ + o) + r - t^ (h~^ + o)||_p, entity rows [main | aux],
//   TranS     scoring.py:1575-1750 -||h^ (t~^ + o + r_bar) - t^ (h~^ + o - r_hat) + r||_p,
//                                  relation rows [r | r_bar | r_hat]; ^ = L2-normalised half
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef BESS_HD
#ifdef __CUDACC__
#define BESS_HD __host__ __device__ __forceinline__
#else
#define BESS_HD inline
#endif
#endif

namespace bess {

enum Family { FAM_TRANSE = 0, FAM_ROTATE = 1, FAM_DISTMULT = 2, FAM_COMPLEX = 3, FAM_PAIRRE = 4, FAM_BOXE = 5,
              FAM_TRIPLERE = 6, FAM_INTERHT = 7, FAM_TRANS = 8 };
// OP_PAIR2: two candidate elements per coordinate (InterHT / TranS), own kernels in pair2.cu
enum PairOp { OP_DIST = 0, OP_DOT = 1, OP_PAIRRE = 2, OP_BOXE = 3, OP_PAIR2 = 4 };
enum Mode { MODE_TAILS = 0, MODE_HEADS = 1 };  // which side the candidates replace

struct FamCfg {
  int family;
  int norm_p;      // 1 or 2 (distance families)
  int d;           // constructor embedding_size
  int normalize;   // PairRE normalize_entities
  int apply_tanh;  // BoxE
  int per_dim;     // BoxE dist_func_per_dim
  float eps;       // BoxE
  float rel_u;     // TripleRE v2 offset added to both relation projections (0 = v1);
                   // InterHT / TranS: the `offset` added to the auxiliary entity embeddings
};
BESS_HD bool is_pair2(int family) { return family == FAM_INTERHT || family == FAM_TRANS; }

// PairRE and TripleRE share one code path: -|| h^ (r_h + u) - t^ (r_t + u) (+ r_m) ||_p with the
// relation row laid out [r_h | r_t] (PairRE, scoring.py:465-593) or [r_h | r_m | r_t] (TripleRE,
// scoring.py:596-743).  Offsets of the three parts inside a relation row; mid < 0 = no r_m.
struct ProjLayout { int oh, ot, om; float u; };
BESS_HD ProjLayout proj_layout(const FamCfg& c) {
  ProjLayout l;
  if (c.family == FAM_TRIPLERE) { l.oh = 0; l.om = c.d; l.ot = 2 * c.d; l.u = c.rel_u; }
  else { l.oh = 0; l.ot = c.d; l.om = -1; l.u = 0.f; }
  return l;
}

BESS_HD int ent_width(const FamCfg& c) {
  return (c.family == FAM_ROTATE || c.family == FAM_COMPLEX || c.family == FAM_BOXE ||
          is_pair2(c.family)) ? 2 * c.d : c.d;
}
BESS_HD int rel_width(const FamCfg& c) {
  switch (c.family) {
    case FAM_COMPLEX: case FAM_PAIRRE: return 2 * c.d;
    case FAM_TRIPLERE: case FAM_TRANS: return 3 * c.d;
    case FAM_BOXE: return 4 * c.d + 2;
    default: return c.d;
  }
}
BESS_HD int pair_op(const FamCfg& c) {
  switch (c.family) {
    case FAM_TRANSE: case FAM_ROTATE: return OP_DIST;
    case FAM_DISTMULT: case FAM_COMPLEX: return OP_DOT;
    case FAM_PAIRRE: case FAM_TRIPLERE: return OP_PAIRRE;
    case FAM_INTERHT: case FAM_TRANS: return OP_PAIR2;
    default: return OP_BOXE;
  }
}
// number of query-side vectors of width ent_width() the pair kernels consume
BESS_HD int pair_nvec(const FamCfg& c) {
  if (is_pair2(c.family)) return 2;
  return (c.family == FAM_PAIRRE || c.family == FAM_TRIPLERE) ? 2 : (c.family == FAM_BOXE ? 3 : 1);
}

// ---- p-norm pieces ---------------------------------------------------------
BESS_HD float nacc(int p, float e) { return p == 1 ? fabsf(e) : e * e; }
BESS_HD float nfin(int p, float a) { return p == 1 ? a : sqrtf(a); }
BESS_HD float fsign(float e) { return (float)((e > 0.f) - (e < 0.f)); }
// coef * sign(e) with sign(0) = 0 (the L1 sub-gradient torch uses).  Device: FMA-pipe
// instructions only — u = sat(e * 2^126 + 0.5) is 1 / 0.5 / 0 for e > 0 / e == 0 / e < 0
// (one FFMA.SAT; every normal e saturates), so (2 coef) * (u - 0.5) is exactly +coef / 0 /
// -coef.  The earlier LOP3 + FSETP + FSEL form ran on the half-rate ALU pipe, which ncu showed
// to be the limiter of the L1 backward tile kernels (alu pipe 66-75 %, fma pipe 23-25 %).
BESS_HD float sign_mul(float coef, float e) {
#ifdef __CUDA_ARCH__
  const float u = __saturatef(fmaf(e, 8.507059173023462e37f /* 2^126 */, 0.5f));
  return (coef + coef) * (u - 0.5f);
#else
  return coef * fsign(e);
#endif
}
// d(norm)/de given the finished norm value nv
BESS_HD float ndiff(int p, float e, float nv) {
  if (p == 1) return fsign(e);
  return nv > 0.f ? e / nv : 0.f;
}
BESS_HD float eluf(float x) { return x > 0.f ? x : expm1f(x); }
BESS_HD float elu_grad(float x) { return x > 0.f ? 1.f : expf(x); }

// ---- typed element access (kernels: fp32/fp16/bf16; host check: float) ------
template <typename T>
struct Ld {
  static BESS_HD float f(const T* p, int i);
};
template <>
struct Ld<float> {
  static BESS_HD float f(const float* p, int i) { return p[i]; }
};

// Host context: a "warp" of one lane.
struct HostCtx {
  static BESS_HD int lane() { return 0; }
  static BESS_HD int lanes() { return 1; }
  static BESS_HD float sum(float v) { return v; }
  static BESS_HD int all(int v) { return v; }
};

// ---------------------------------------------------------------------------
// BoxE helpers.  Relation row: [cen_h d | cen_t d | wid_h d | wid_t d | size_h | size_t]
// Box b uses cen = r[b*d ..], wid = r[2d + b*d ..], size = r[4d + b].
// bumped[b=0] = h_base + t_bump, bumped[b=1] = h_bump + t_base (scoring.py:1354-1357).
// ---------------------------------------------------------------------------
struct BoxNorm {
  float den, gm, scale;
};
template <typename Ctx, typename T>
BESS_HD BoxNorm boxe_norm(const FamCfg& c, const T* r, int b) {
  const int d = c.d;
  float lg = 0.f;
  for (int k = Ctx::lane(); k < d; k += Ctx::lanes())
    lg += logf(fmaxf(fabsf(Ld<T>::f(r, 2 * d + b * d + k)), c.eps));
  lg = Ctx::sum(lg);
  BoxNorm n;
  n.gm = expf(lg / (float)d);
  n.den = fmaxf(n.gm, c.eps);
  n.scale = 1.f + eluf(Ld<T>::f(r, 4 * d + b));
  return n;
}
// effective (post-tanh) centre and width of one box dimension
struct BoxDim {
  float wn, low, up, cen, w;
};
BESS_HD BoxDim boxe_dim(const FamCfg& c, const BoxNorm& n, float cen_raw, float wid_raw) {
  BoxDim o;
  o.wn = fabsf(wid_raw) / n.den * n.scale;
  if (c.apply_tanh) {
    o.low = tanhf(cen_raw - 0.5f * o.wn);
    o.up = tanhf(o.low + o.wn);
    o.cen = 0.5f * (o.low + o.up);
    o.w = o.up - o.low;
  } else {
    o.low = o.up = 0.f;
    o.cen = cen_raw;
    o.w = o.wn;
  }
  return o;
}
// distance of one bumped coordinate to its box (scoring.py:1299-1338)
struct BoxDist {
  float x, cd, wp1, dist;
  int inside;
};
BESS_HD BoxDist boxe_dist(int apply_tanh, float bumped, float cen, float w, int force_mode /*-1 per-dim, 0 out, 1 in*/) {
  BoxDist o;
  o.x = apply_tanh ? tanhf(bumped) : bumped;
  o.cd = fabsf(o.x - cen);
  o.wp1 = 1.f + w;
  o.inside = force_mode < 0 ? (o.cd <= 0.5f * w) : force_mode;
  o.dist = o.inside ? o.cd / o.wp1 : o.cd * o.wp1 - 0.5f * w * (o.wp1 - 1.f / o.wp1);
  return o;
}
// gradient of boxe_dist w.r.t. (bumped, cen, w) given dL/d(dist)
BESS_HD void boxe_dist_bwd(int apply_tanh, const BoxDist& f, float cen, float w, float dd, float& dbumped, float& dcen, float& dw) {
  float dcd, dwp1;
  dw = 0.f;
  if (f.inside) {
    dcd = dd / f.wp1;
    dwp1 = -dd * f.cd / (f.wp1 * f.wp1);
  } else {
    dcd = dd * f.wp1;
    dwp1 = dd * f.cd;
    const float inv = 1.f / f.wp1;
    dw += -dd * 0.5f * (f.wp1 - inv);           // via k, dk/dw (explicit)
    dwp1 += -dd * 0.5f * w * (1.f + inv * inv);  // via k, dk/dwp1
  }
  dw += dwp1;
  const float dx = dcd * fsign(f.x - cen);
  dcen = -dx;
  dbumped = apply_tanh ? dx * (1.f - f.x * f.x) : dx;
}

// Turn accumulated (dL/dcen', dL/dw') stored in a relation-gradient row into
// (dL/dcen, dL/dwid, dL/dsize) in place.  The map is linear in the incoming
// gradients, so the triple path and the negative path may both accumulate into
// the row before this runs once.  g: fp32 row [4d+2]; r: relation row.
template <typename Ctx, typename T>
BESS_HD void boxe_rel_finalize(const FamCfg& c, const T* r, float* g) {
  const int d = c.d;
  for (int b = 0; b < 2; ++b) {
    const BoxNorm n = boxe_norm<Ctx, T>(c, r, b);
    float s1 = 0.f;
    for (int k = Ctx::lane(); k < d; k += Ctx::lanes()) {
      const float cen_raw = Ld<T>::f(r, b * d + k), wid_raw = Ld<T>::f(r, 2 * d + b * d + k);
      const BoxDim bd = boxe_dim(c, n, cen_raw, wid_raw);
      const float dcp = g[b * d + k], dwp = g[2 * d + b * d + k];
      float dcen, dwn;
      if (c.apply_tanh) {
        float dlow = 0.5f * dcp - dwp;
        const float dup = 0.5f * dcp + dwp;
        const float db = dup * (1.f - bd.up * bd.up);
        dlow += db;
        const float da = dlow * (1.f - bd.low * bd.low);
        dcen = da;
        dwn = db - 0.5f * da;
      } else {
        dcen = dcp;
        dwn = dwp;
      }
      g[b * d + k] = dcen;
      g[2 * d + b * d + k] = dwn;  // finished in the second pass
      s1 += dwn * bd.wn;
    }
    s1 = Ctx::sum(s1);
    const float dden = -s1 / n.den;
    const float dscale = s1 / n.scale;
    const float dgm = (n.gm >= c.eps) ? dden : 0.f;
    const float dl = dgm * n.gm / (float)d;
    for (int k = Ctx::lane(); k < d; k += Ctx::lanes()) {
      const float wid_raw = Ld<T>::f(r, 2 * d + b * d + k);
      const float aw = fabsf(wid_raw);
      float daw = g[2 * d + b * d + k] * n.scale / n.den;
      if (aw >= c.eps) daw += dl / aw;
      g[2 * d + b * d + k] = daw * fsign(wid_raw);
    }
    if (Ctx::lane() == 0) g[4 * d + b] = dscale * elu_grad(Ld<T>::f(r, 4 * d + b));
  }
}

// ---------------------------------------------------------------------------
// InterHT / TranS helpers: entity row [main d | aux d], each half L2-normalised
// (torch.nn.functional.normalize: x / max(||x||, 1e-12)) when cfg.normalize.
// ---------------------------------------------------------------------------
struct Pair2Norms {
  float nm, na;  // norms of the two halves (1 when not normalising)
  float im, ia;  // 1 / max(norm, 1e-12)
};
template <typename Ctx, typename T>
BESS_HD Pair2Norms pair2_norms(const FamCfg& c, const T* x) {
  Pair2Norms n;
  n.nm = n.na = n.im = n.ia = 1.f;
  if (c.normalize) {
    float a = 0.f, b = 0.f;
    for (int k = Ctx::lane(); k < c.d; k += Ctx::lanes()) {
      const float u = Ld<T>::f(x, k), v = Ld<T>::f(x, c.d + k);
      a += u * u; b += v * v;
    }
    n.nm = sqrtf(Ctx::sum(a)); n.na = sqrtf(Ctx::sum(b));
    n.im = 1.f / fmaxf(n.nm, 1e-12f); n.ia = 1.f / fmaxf(n.na, 1e-12f);
  }
  return n;
}
// gradient w.r.t. the raw half from the gradient w.r.t. the normalised half:
// (g - x^ (x^ . g)) / ||x|| above the clamp, g / 1e-12 below it
BESS_HD float unnorm_grad(int normalize, float g, float xhat, float proj, float norm, float inv) {
  if (!normalize) return g;
  return norm > 1e-12f ? (g - xhat * proj) * inv : g * inv;
}

// ---------------------------------------------------------------------------
// score_triple forward: one row (h, r, t) -> score.
// ---------------------------------------------------------------------------
template <typename Ctx, typename T>
BESS_HD float triple_fwd(const FamCfg& c, const T* h, const T* r, const T* t) {
  const int d = c.d, p = c.norm_p;
  const int l0 = Ctx::lane(), ls = Ctx::lanes();
  float acc = 0.f;
  switch (c.family) {
    case FAM_TRANSE: {
      for (int k = l0; k < d; k += ls) acc += nacc(p, Ld<T>::f(h, k) + Ld<T>::f(r, k) - Ld<T>::f(t, k));
      return -nfin(p, Ctx::sum(acc));
    }
    case FAM_ROTATE: {
      for (int k = l0; k < d; k += ls) {
        float s, co;
        sincosf(Ld<T>::f(r, k), &s, &co);
        const float hr = Ld<T>::f(h, k), hi = Ld<T>::f(h, d + k);
        acc += nacc(p, hr * co - hi * s - Ld<T>::f(t, k)) + nacc(p, hr * s + hi * co - Ld<T>::f(t, d + k));
      }
      return -nfin(p, Ctx::sum(acc));
    }
    case FAM_DISTMULT: {
      for (int k = l0; k < d; k += ls) acc += Ld<T>::f(h, k) * Ld<T>::f(r, k) * Ld<T>::f(t, k);
      return Ctx::sum(acc);
    }
    case FAM_COMPLEX: {
      for (int k = l0; k < d; k += ls) {
        const float hr = Ld<T>::f(h, k), hi = Ld<T>::f(h, d + k), rr = Ld<T>::f(r, k), ri = Ld<T>::f(r, d + k);
        acc += (hr * rr - hi * ri) * Ld<T>::f(t, k) + (hr * ri + hi * rr) * Ld<T>::f(t, d + k);
      }
      return Ctx::sum(acc);
    }
    case FAM_PAIRRE: case FAM_TRIPLERE: {
      float ih = 1.f, it = 1.f;
      if (c.normalize) {
        float nh = 0.f, nt = 0.f;
        for (int k = l0; k < d; k += ls) {
          const float a = Ld<T>::f(h, k), b = Ld<T>::f(t, k);
          nh += a * a; nt += b * b;
        }
        ih = 1.f / fmaxf(sqrtf(Ctx::sum(nh)), 1e-12f);
        it = 1.f / fmaxf(sqrtf(Ctx::sum(nt)), 1e-12f);
      }
      const ProjLayout pl = proj_layout(c);
      for (int k = l0; k < d; k += ls) {
        float e = Ld<T>::f(h, k) * ih * (Ld<T>::f(r, pl.oh + k) + pl.u) -
                  Ld<T>::f(t, k) * it * (Ld<T>::f(r, pl.ot + k) + pl.u);
        if (pl.om >= 0) e += Ld<T>::f(r, pl.om + k);
        acc += nacc(p, e);
      }
      return -nfin(p, Ctx::sum(acc));
    }
    case FAM_INTERHT: case FAM_TRANS: {
      const Pair2Norms nh = pair2_norms<Ctx, T>(c, h), nt = pair2_norms<Ctx, T>(c, t);
      const bool ts = c.family == FAM_TRANS;
      const float o = c.rel_u;
      for (int k = l0; k < d; k += ls) {
        const float hm = Ld<T>::f(h, k) * nh.im, ha = Ld<T>::f(h, d + k) * nh.ia;
        const float tm = Ld<T>::f(t, k) * nt.im, ta = Ld<T>::f(t, d + k) * nt.ia;
        const float rb = ts ? Ld<T>::f(r, d + k) : 0.f, rh = ts ? Ld<T>::f(r, 2 * d + k) : 0.f;
        acc += nacc(p, hm * (ta + o + rb) - tm * (ha + o - rh) + Ld<T>::f(r, k));
      }
      return -nfin(p, Ctx::sum(acc));
    }
    default: {  // FAM_BOXE
      float total = 0.f;
      for (int b = 0; b < 2; ++b) {
        const BoxNorm n = boxe_norm<Ctx, T>(c, r, b);
        int mode = -1;
        if (!c.per_dim) {
          int all_in = 1;
          for (int k = l0; k < d; k += ls) {
            const BoxDim bd = boxe_dim(c, n, Ld<T>::f(r, b * d + k), Ld<T>::f(r, 2 * d + b * d + k));
            const float bumped = Ld<T>::f(h, b * d + k) + Ld<T>::f(t, (1 - b) * d + k);
            all_in &= boxe_dist(c.apply_tanh, bumped, bd.cen, bd.w, -1).inside;
          }
          mode = Ctx::all(all_in);
        }
        float a = 0.f;
        for (int k = l0; k < d; k += ls) {
          const BoxDim bd = boxe_dim(c, n, Ld<T>::f(r, b * d + k), Ld<T>::f(r, 2 * d + b * d + k));
          const float bumped = Ld<T>::f(h, b * d + k) + Ld<T>::f(t, (1 - b) * d + k);
          a += nacc(p, boxe_dist(c.apply_tanh, bumped, bd.cen, bd.w, mode).dist);
        }
        total += nfin(p, Ctx::sum(a));
      }
      return -total;
    }
  }
}

// ---------------------------------------------------------------------------
// score_triple backward.  g = dL/dscore.  Writes (or accumulates, add != 0)
// fp32 gradient rows dh, dt [W] and dr [Wr].  `score` is the forward value.
// For BoxE, dr receives PRE-CHAIN gradients (see boxe_rel_finalize).
// ---------------------------------------------------------------------------
BESS_HD void put(float* p, int i, float v, int add) { p[i] = add ? p[i] + v : v; }

template <typename Ctx, typename T>
BESS_HD void triple_bwd(const FamCfg& c, const T* h, const T* r, const T* t, float score, float g,
                        float* dh, float* dr, float* dt, int add_h, int add_r, int add_t) {
  const int d = c.d, p = c.norm_p;
  const int l0 = Ctx::lane(), ls = Ctx::lanes();
  switch (c.family) {
    case FAM_TRANSE: {
      const float nv = -score;
      for (int k = l0; k < d; k += ls) {
        const float e = Ld<T>::f(h, k) + Ld<T>::f(r, k) - Ld<T>::f(t, k);
        const float de = -g * ndiff(p, e, nv);
        put(dh, k, de, add_h); put(dr, k, de, add_r); put(dt, k, -de, add_t);
      }
      return;
    }
    case FAM_ROTATE: {
      const float nv = -score;
      for (int k = l0; k < d; k += ls) {
        float s, co;
        sincosf(Ld<T>::f(r, k), &s, &co);
        const float hr = Ld<T>::f(h, k), hi = Ld<T>::f(h, d + k);
        const float qr = hr * co - hi * s, qi = hr * s + hi * co;
        const float dre = -g * ndiff(p, qr - Ld<T>::f(t, k), nv);
        const float dim = -g * ndiff(p, qi - Ld<T>::f(t, d + k), nv);
        put(dh, k, dre * co + dim * s, add_h);
        put(dh, d + k, -dre * s + dim * co, add_h);
        put(dt, k, -dre, add_t);
        put(dt, d + k, -dim, add_t);
        put(dr, k, -dre * qi + dim * qr, add_r);
      }
      return;
    }
    case FAM_DISTMULT: {
      for (int k = l0; k < d; k += ls) {
        const float a = Ld<T>::f(h, k), b = Ld<T>::f(r, k), e = Ld<T>::f(t, k);
        put(dh, k, g * b * e, add_h); put(dr, k, g * a * e, add_r); put(dt, k, g * a * b, add_t);
      }
      return;
    }
    case FAM_COMPLEX: {
      for (int k = l0; k < d; k += ls) {
        const float hr = Ld<T>::f(h, k), hi = Ld<T>::f(h, d + k), rr = Ld<T>::f(r, k), ri = Ld<T>::f(r, d + k);
        const float tr = Ld<T>::f(t, k), ti = Ld<T>::f(t, d + k);
        const float mr = hr * rr - hi * ri, mi = hr * ri + hi * rr;
        const float dmr = g * tr, dmi = g * ti;
        put(dt, k, g * mr, add_t); put(dt, d + k, g * mi, add_t);
        put(dh, k, dmr * rr + dmi * ri, add_h); put(dh, d + k, -dmr * ri + dmi * rr, add_h);
        put(dr, k, dmr * hr + dmi * hi, add_r); put(dr, d + k, -dmr * hi + dmi * hr, add_r);
      }
      return;
    }
    case FAM_PAIRRE: case FAM_TRIPLERE: {
      const float nv = -score;
      float nh = 1.f, nt = 1.f, ih = 1.f, it = 1.f;
      if (c.normalize) {
        float a2 = 0.f, b2 = 0.f;
        for (int k = l0; k < d; k += ls) {
          const float a = Ld<T>::f(h, k), b = Ld<T>::f(t, k);
          a2 += a * a; b2 += b * b;
        }
        nh = sqrtf(Ctx::sum(a2)); nt = sqrtf(Ctx::sum(b2));
        ih = 1.f / fmaxf(nh, 1e-12f); it = 1.f / fmaxf(nt, 1e-12f);
      }
      // first pass: relation grads and the projections hh.dhh, tt.dtt
      const ProjLayout pl = proj_layout(c);
      float ph = 0.f, pt = 0.f;
      for (int k = l0; k < d; k += ls) {
        const float hh = Ld<T>::f(h, k) * ih, tt = Ld<T>::f(t, k) * it;
        const float rh = Ld<T>::f(r, pl.oh + k) + pl.u, rt = Ld<T>::f(r, pl.ot + k) + pl.u;
        const float rm = pl.om >= 0 ? Ld<T>::f(r, pl.om + k) : 0.f;
        const float de = -g * ndiff(p, hh * rh - tt * rt + rm, nv);
        put(dr, pl.oh + k, de * hh, add_r); put(dr, pl.ot + k, -de * tt, add_r);
        if (pl.om >= 0) put(dr, pl.om + k, de, add_r);
        ph += hh * (de * rh); pt += tt * (-de * rt);
      }
      if (c.normalize) { ph = Ctx::sum(ph); pt = Ctx::sum(pt); }
      for (int k = l0; k < d; k += ls) {
        const float hh = Ld<T>::f(h, k) * ih, tt = Ld<T>::f(t, k) * it;
        const float rh = Ld<T>::f(r, pl.oh + k) + pl.u, rt = Ld<T>::f(r, pl.ot + k) + pl.u;
        const float rm = pl.om >= 0 ? Ld<T>::f(r, pl.om + k) : 0.f;
        const float de = -g * ndiff(p, hh * rh - tt * rt + rm, nv);
        float dhh = de * rh, dtt = -de * rt;
        if (c.normalize) {
          // d/dx of x/max(||x||,eps): (I - x^ x^T)/||x|| above eps, I/eps below
          dhh = nh > 1e-12f ? (dhh - hh * ph) * ih : dhh * ih;
          dtt = nt > 1e-12f ? (dtt - tt * pt) * it : dtt * it;
        }
        put(dh, k, dhh, add_h); put(dt, k, dtt, add_t);
      }
      return;
    }
    case FAM_INTERHT: case FAM_TRANS: {
      const float nv = -score, o = c.rel_u;
      const bool ts = c.family == FAM_TRANS;
      const Pair2Norms nh = pair2_norms<Ctx, T>(c, h), nt = pair2_norms<Ctx, T>(c, t);
      // pass 1: relation gradients and the projections x^ . dx^ of the four halves
      float phm = 0.f, pha = 0.f, ptm = 0.f, pta = 0.f;
      for (int k = l0; k < d; k += ls) {
        const float hm = Ld<T>::f(h, k) * nh.im, ha = Ld<T>::f(h, d + k) * nh.ia;
        const float tm = Ld<T>::f(t, k) * nt.im, ta = Ld<T>::f(t, d + k) * nt.ia;
        const float rb = ts ? Ld<T>::f(r, d + k) : 0.f, rh = ts ? Ld<T>::f(r, 2 * d + k) : 0.f;
        const float A = ta + o + rb, Bq = ha + o - rh;
        const float de = -g * ndiff(p, hm * A - tm * Bq + Ld<T>::f(r, k), nv);
        put(dr, k, de, add_r);
        if (ts) { put(dr, d + k, de * hm, add_r); put(dr, 2 * d + k, de * tm, add_r); }
        phm += hm * (de * A); pha += ha * (-de * tm); ptm += tm * (-de * Bq); pta += ta * (de * hm);
      }
      if (c.normalize) { phm = Ctx::sum(phm); pha = Ctx::sum(pha); ptm = Ctx::sum(ptm); pta = Ctx::sum(pta); }
      for (int k = l0; k < d; k += ls) {
        const float hm = Ld<T>::f(h, k) * nh.im, ha = Ld<T>::f(h, d + k) * nh.ia;
        const float tm = Ld<T>::f(t, k) * nt.im, ta = Ld<T>::f(t, d + k) * nt.ia;
        const float rb = ts ? Ld<T>::f(r, d + k) : 0.f, rh = ts ? Ld<T>::f(r, 2 * d + k) : 0.f;
        const float A = ta + o + rb, Bq = ha + o - rh;
        const float de = -g * ndiff(p, hm * A - tm * Bq + Ld<T>::f(r, k), nv);
        put(dh, k, unnorm_grad(c.normalize, de * A, hm, phm, nh.nm, nh.im), add_h);
        put(dh, d + k, unnorm_grad(c.normalize, -de * tm, ha, pha, nh.na, nh.ia), add_h);
        put(dt, k, unnorm_grad(c.normalize, -de * Bq, tm, ptm, nt.nm, nt.im), add_t);
        put(dt, d + k, unnorm_grad(c.normalize, de * hm, ta, pta, nt.na, nt.ia), add_t);
      }
      return;
    }
    default: {  // FAM_BOXE
      for (int b = 0; b < 2; ++b) {
        const BoxNorm n = boxe_norm<Ctx, T>(c, r, b);
        int mode = -1;
        if (!c.per_dim) {
          int all_in = 1;
          for (int k = l0; k < d; k += ls) {
            const BoxDim bd = boxe_dim(c, n, Ld<T>::f(r, b * d + k), Ld<T>::f(r, 2 * d + b * d + k));
            all_in &= boxe_dist(c.apply_tanh, Ld<T>::f(h, b * d + k) + Ld<T>::f(t, (1 - b) * d + k), bd.cen, bd.w, -1).inside;
          }
          mode = Ctx::all(all_in);
        }
        float a = 0.f;
        for (int k = l0; k < d; k += ls) {
          const BoxDim bd = boxe_dim(c, n, Ld<T>::f(r, b * d + k), Ld<T>::f(r, 2 * d + b * d + k));
          a += nacc(p, boxe_dist(c.apply_tanh, Ld<T>::f(h, b * d + k) + Ld<T>::f(t, (1 - b) * d + k), bd.cen, bd.w, mode).dist);
        }
        const float nv = nfin(p, Ctx::sum(a));
        for (int k = l0; k < d; k += ls) {
          const BoxDim bd = boxe_dim(c, n, Ld<T>::f(r, b * d + k), Ld<T>::f(r, 2 * d + b * d + k));
          const BoxDist f = boxe_dist(c.apply_tanh, Ld<T>::f(h, b * d + k) + Ld<T>::f(t, (1 - b) * d + k), bd.cen, bd.w, mode);
          const float dd = -g * ndiff(p, f.dist, nv);
          float dbump, dcen, dw;
          boxe_dist_bwd(c.apply_tanh, f, bd.cen, bd.w, dd, dbump, dcen, dw);
          put(dh, b * d + k, dbump, add_h);
          put(dt, (1 - b) * d + k, dbump, add_t);
          put(dr, b * d + k, dcen, add_r);
          put(dr, 2 * d + b * d + k, dw, add_r);
        }
        if (l0 == 0 && !add_r) dr[4 * d + b] = 0.f;
      }
      return;
    }
  }
}

// ---------------------------------------------------------------------------
// Query prologue for score_heads / score_tails: (fixed entity row x, relation
// row r) -> NV query vectors qv[v*W + k] consumed by the pair kernels.
//   OP_DIST   score = -|| qv - c ||_p              (TransE, RotatE)
//   OP_DOT    score = sum qv * c                    (DistMult, ComplEx)
//   OP_PAIRRE score = -|| c^ * qv1 - qv0 ||_p       (c^ = normalised candidate)
//   OP_BOXE   qv0 = partner coordinate, qv1 = centre', qv2 = width';
//             coordinate k pairs with candidate element (k + rot) mod W where
//             rot = d for MODE_TAILS and 0 for MODE_HEADS.
// ---------------------------------------------------------------------------
template <typename Ctx, typename T>
BESS_HD void prologue_fwd(const FamCfg& c, int mode, const T* x, const T* r, float* qv) {
  const int d = c.d, W = ent_width(c);
  const int l0 = Ctx::lane(), ls = Ctx::lanes();
  switch (c.family) {
    case FAM_TRANSE:
      for (int k = l0; k < d; k += ls)
        qv[k] = mode == MODE_TAILS ? Ld<T>::f(x, k) + Ld<T>::f(r, k) : Ld<T>::f(x, k) - Ld<T>::f(r, k);
      return;
    case FAM_ROTATE:
      for (int k = l0; k < d; k += ls) {
        float s, co;
        const float ang = Ld<T>::f(r, k);
        sincosf(mode == MODE_TAILS ? ang : -ang, &s, &co);
        const float xr = Ld<T>::f(x, k), xi = Ld<T>::f(x, d + k);
        qv[k] = xr * co - xi * s;
        qv[d + k] = xr * s + xi * co;
      }
      return;
    case FAM_DISTMULT:
      for (int k = l0; k < d; k += ls) qv[k] = Ld<T>::f(x, k) * Ld<T>::f(r, k);
      return;
    case FAM_COMPLEX:
      for (int k = l0; k < d; k += ls) {
        const float xr = Ld<T>::f(x, k), xi = Ld<T>::f(x, d + k), rr = Ld<T>::f(r, k), ri = Ld<T>::f(r, d + k);
        if (mode == MODE_TAILS) {  // h (x) r
          qv[k] = xr * rr - xi * ri;
          qv[d + k] = xr * ri + xi * rr;
        } else {  // conj(r) (x) t  (scoring.py:928-932)
          qv[k] = rr * xr + ri * xi;
          qv[d + k] = rr * xi - ri * xr;
        }
      }
      return;
    case FAM_PAIRRE: case FAM_TRIPLERE: {
      float inv = 1.f;
      if (c.normalize) {
        float a2 = 0.f;
        for (int k = l0; k < d; k += ls) { const float a = Ld<T>::f(x, k); a2 += a * a; }
        inv = 1.f / fmaxf(sqrtf(Ctx::sum(a2)), 1e-12f);
      }
      // tails: fixed = head -> qv0 = h^ r_h (+ r_m), qv1 = r_t ; heads: qv0 = t^ r_t (- r_m), qv1 = r_h
      // (score = -|| c^ qv1 - qv0 ||; TripleRE: scoring.py:699-743)
      const ProjLayout pl = proj_layout(c);
      const int own = mode == MODE_TAILS ? pl.oh : pl.ot, other = mode == MODE_TAILS ? pl.ot : pl.oh;
      const float sm = mode == MODE_TAILS ? 1.f : -1.f;
      for (int k = l0; k < d; k += ls) {
        float q0 = Ld<T>::f(x, k) * inv * (Ld<T>::f(r, own + k) + pl.u);
        if (pl.om >= 0) q0 += sm * Ld<T>::f(r, pl.om + k);
        qv[k] = q0;
        qv[W + k] = Ld<T>::f(r, other + k) + pl.u;
      }
      return;
    }
    case FAM_INTERHT: case FAM_TRANS: {
      // residual e_k = qv0[k] * c^m_k + qv0[d + k] * c^a_k + qv1[k] over the candidate's
      // normalised halves (scoring.py:1530-1572, 1700-1750):
      //   tails (x = head):  e = -(h~ + o - r_hat) c^m + h c^a + h (o + r_bar) + r
      //   heads (x = tail):  e =  (t~ + o + r_bar) c^m - t c^a + r - t (o - r_hat)
      const Pair2Norms nx = pair2_norms<Ctx, T>(c, x);
      const bool ts = c.family == FAM_TRANS;
      const float o = c.rel_u;
      for (int k = l0; k < d; k += ls) {
        const float xm = Ld<T>::f(x, k) * nx.im, xa = Ld<T>::f(x, d + k) * nx.ia;
        const float rb = ts ? Ld<T>::f(r, d + k) : 0.f, rh = ts ? Ld<T>::f(r, 2 * d + k) : 0.f;
        const float rr = Ld<T>::f(r, k);
        if (mode == MODE_TAILS) {
          qv[k] = -(xa + o - rh); qv[d + k] = xm; qv[W + k] = xm * (o + rb) + rr;
        } else {
          qv[k] = xa + o + rb; qv[d + k] = -xm; qv[W + k] = rr - xm * (o - rh);
        }
        qv[W + d + k] = 0.f;
      }
      return;
    }
    default: {  // FAM_BOXE
      for (int b = 0; b < 2; ++b) {
        const BoxNorm n = boxe_norm<Ctx, T>(c, r, b);
        for (int k = l0; k < d; k += ls) {
          const BoxDim bd = boxe_dim(c, n, Ld<T>::f(r, b * d + k), Ld<T>::f(r, 2 * d + b * d + k));
          // box b coordinate k: tails: h[b*d+k] + cand[(1-b)*d+k]; heads: cand[b*d+k] + t[(1-b)*d+k]
          qv[b * d + k] = mode == MODE_TAILS ? Ld<T>::f(x, b * d + k) : Ld<T>::f(x, (1 - b) * d + k);
          qv[W + b * d + k] = bd.cen;
          qv[2 * W + b * d + k] = bd.w;
        }
      }
      return;
    }
  }
}

// Backward of the prologue: dqv [NV*W] -> dx [W] (accumulated if add_x) and
// dr [Wr] (accumulated if add_r; BoxE: pre-chain slots).
template <typename Ctx, typename T>
BESS_HD void prologue_bwd(const FamCfg& c, int mode, const T* x, const T* r, const float* dqv,
                          float* dx, float* dr, int add_x, int add_r) {
  const int d = c.d, W = ent_width(c);
  const int l0 = Ctx::lane(), ls = Ctx::lanes();
  switch (c.family) {
    case FAM_TRANSE:
      for (int k = l0; k < d; k += ls) {
        put(dx, k, dqv[k], add_x);
        put(dr, k, mode == MODE_TAILS ? dqv[k] : -dqv[k], add_r);
      }
      return;
    case FAM_ROTATE:
      for (int k = l0; k < d; k += ls) {
        float s, co;
        const float ang = Ld<T>::f(r, k);
        sincosf(mode == MODE_TAILS ? ang : -ang, &s, &co);
        const float xr = Ld<T>::f(x, k), xi = Ld<T>::f(x, d + k);
        const float qr = xr * co - xi * s, qi = xr * s + xi * co;
        const float a = dqv[k], b = dqv[d + k];
        put(dx, k, a * co + b * s, add_x);
        put(dx, d + k, -a * s + b * co, add_x);
        const float dang = -a * qi + b * qr;
        put(dr, k, mode == MODE_TAILS ? dang : -dang, add_r);
      }
      return;
    case FAM_DISTMULT:
      for (int k = l0; k < d; k += ls) {
        put(dx, k, dqv[k] * Ld<T>::f(r, k), add_x);
        put(dr, k, dqv[k] * Ld<T>::f(x, k), add_r);
      }
      return;
    case FAM_COMPLEX:
      for (int k = l0; k < d; k += ls) {
        const float xr = Ld<T>::f(x, k), xi = Ld<T>::f(x, d + k), rr = Ld<T>::f(r, k), ri = Ld<T>::f(r, d + k);
        const float a = dqv[k], b = dqv[d + k];
        if (mode == MODE_TAILS) {
          put(dx, k, a * rr + b * ri, add_x); put(dx, d + k, -a * ri + b * rr, add_x);
          put(dr, k, a * xr + b * xi, add_r); put(dr, d + k, -a * xi + b * xr, add_r);
        } else {
          put(dx, k, a * rr - b * ri, add_x); put(dx, d + k, a * ri + b * rr, add_x);
          put(dr, k, a * xr + b * xi, add_r); put(dr, d + k, a * xi - b * xr, add_r);
        }
      }
      return;
    case FAM_PAIRRE: case FAM_TRIPLERE: {
      float nx = 1.f, inv = 1.f;
      if (c.normalize) {
        float a2 = 0.f;
        for (int k = l0; k < d; k += ls) { const float a = Ld<T>::f(x, k); a2 += a * a; }
        nx = sqrtf(Ctx::sum(a2));
        inv = 1.f / fmaxf(nx, 1e-12f);
      }
      const ProjLayout pl = proj_layout(c);
      const int own = mode == MODE_TAILS ? pl.oh : pl.ot, other = mode == MODE_TAILS ? pl.ot : pl.oh;
      const float sm = mode == MODE_TAILS ? 1.f : -1.f;
      float proj = 0.f;
      for (int k = l0; k < d; k += ls) {
        const float xh = Ld<T>::f(x, k) * inv;
        put(dr, own + k, dqv[k] * xh, add_r);
        put(dr, other + k, dqv[W + k], add_r);
        if (pl.om >= 0) put(dr, pl.om + k, sm * dqv[k], add_r);
        proj += xh * dqv[k] * (Ld<T>::f(r, own + k) + pl.u);
      }
      if (c.normalize) proj = Ctx::sum(proj);
      for (int k = l0; k < d; k += ls) {
        const float xh = Ld<T>::f(x, k) * inv;
        float dxh = dqv[k] * (Ld<T>::f(r, own + k) + pl.u);
        if (c.normalize) dxh = nx > 1e-12f ? (dxh - xh * proj) * inv : dxh * inv;
        put(dx, k, dxh, add_x);
      }
      return;
    }
    case FAM_INTERHT: case FAM_TRANS: {
      const Pair2Norms nx = pair2_norms<Ctx, T>(c, x);
      const bool ts = c.family == FAM_TRANS;
      const float o = c.rel_u;
      float pm = 0.f, pa = 0.f;
      for (int k = l0; k < d; k += ls) {
        const float xm = Ld<T>::f(x, k) * nx.im, xa = Ld<T>::f(x, d + k) * nx.ia;
        const float rb = ts ? Ld<T>::f(r, d + k) : 0.f, rh = ts ? Ld<T>::f(r, 2 * d + k) : 0.f;
        const float g0 = dqv[k], g1 = dqv[d + k], gz = dqv[W + k];
        float dxm, dxa;
        if (mode == MODE_TAILS) {
          dxa = -g0; dxm = g1 + gz * (o + rb);
          if (ts) { put(dr, 2 * d + k, g0, add_r); put(dr, d + k, gz * xm, add_r); }
        } else {
          dxa = g0; dxm = -g1 - gz * (o - rh);
          if (ts) { put(dr, d + k, g0, add_r); put(dr, 2 * d + k, gz * xm, add_r); }
        }
        put(dr, k, gz, add_r);
        pm += xm * dxm; pa += xa * dxa;
      }
      if (c.normalize) { pm = Ctx::sum(pm); pa = Ctx::sum(pa); }
      for (int k = l0; k < d; k += ls) {
        const float xm = Ld<T>::f(x, k) * nx.im, xa = Ld<T>::f(x, d + k) * nx.ia;
        const float rb = ts ? Ld<T>::f(r, d + k) : 0.f, rh = ts ? Ld<T>::f(r, 2 * d + k) : 0.f;
        const float g0 = dqv[k], g1 = dqv[d + k], gz = dqv[W + k];
        const float dxa = mode == MODE_TAILS ? -g0 : g0;
        const float dxm = mode == MODE_TAILS ? g1 + gz * (o + rb) : -g1 - gz * (o - rh);
        put(dx, k, unnorm_grad(c.normalize, dxm, xm, pm, nx.nm, nx.im), add_x);
        put(dx, d + k, unnorm_grad(c.normalize, dxa, xa, pa, nx.na, nx.ia), add_x);
      }
      return;
    }
    default: {  // FAM_BOXE
      for (int b = 0; b < 2; ++b) {
        for (int k = l0; k < d; k += ls) {
          const int xi = mode == MODE_TAILS ? b * d + k : (1 - b) * d + k;
          put(dx, xi, dqv[b * d + k], add_x);
          put(dr, b * d + k, dqv[W + b * d + k], add_r);
          put(dr, 2 * d + b * d + k, dqv[2 * W + b * d + k], add_r);
        }
        if (l0 == 0 && !add_r) dr[4 * d + b] = 0.f;
      }
      return;
    }
  }
}

// ---------------------------------------------------------------------------
// Pair-op element functions used by the shared-negative tile kernels and the
// per-triple-negative streaming kernels.  `cv` is the (already normalised for
// PAIRRE, already rotated for BOXE) candidate element.
// contribution to the reduction:
// ---------------------------------------------------------------------------
template <int OP>
BESS_HD float pair_elem(int p, int apply_tanh, float q0, float q1, float q2, float cv) {
  if (OP == OP_DOT) return q0 * cv;
  if (OP == OP_DIST) return nacc(p, q0 - cv);
  if (OP == OP_PAIRRE) return nacc(p, cv * q1 - q0);
  return nacc(p, boxe_dist(apply_tanh, q0 + cv, q1, q2, -1).dist);
}
// Derivatives.  `coef` is dL/d(sum) folded with the norm: for OP_DOT coef = g;
// for p=1 coef = -g; for p=2 coef = -g / norm (0 when norm == 0) so that the
// element derivative is coef * d(nacc)/2... (see pair kernels): we define
//   dL/d(e) = coef * (p==1 ? sign(e) : e)   for the inner residual e.
template <int OP>
BESS_HD void pair_elem_bwd(int p, int apply_tanh, float q0, float q1, float q2, float cv, float coef,
                           float& dq0, float& dq1, float& dq2, float& dcv) {
  if (OP == OP_DOT) {
    dq0 = coef * cv; dcv = coef * q0; dq1 = dq2 = 0.f;
  } else if (OP == OP_DIST) {
    const float e = q0 - cv;
    const float de = p == 1 ? sign_mul(coef, e) : coef * e;
    dq0 = de; dcv = -de; dq1 = dq2 = 0.f;
  } else if (OP == OP_PAIRRE) {
    const float e = cv * q1 - q0;
    const float de = p == 1 ? sign_mul(coef, e) : coef * e;
    dq0 = -de; dq1 = de * cv; dcv = de * q1; dq2 = 0.f;
  } else {
    const BoxDist f = boxe_dist(apply_tanh, q0 + cv, q1, q2, -1);
    const float dd = coef * (p == 1 ? fsign(f.dist) : f.dist);
    float dbump, dcen, dw;
    boxe_dist_bwd(apply_tanh, f, q1, q2, dd, dbump, dcen, dw);
    dq0 = dbump; dcv = dbump; dq1 = dcen; dq2 = dw;
  }
}
// number of independent norm segments along the row (BoxE sums two norms)
BESS_HD int pair_nseg(int op) { return op == OP_BOXE ? 2 : 1; }

}  // namespace bess
