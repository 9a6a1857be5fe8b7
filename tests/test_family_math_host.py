"""The score-function math compiled into the CUDA kernels (csrc/families.cuh:
forward formulas, query prologues, pair functions and the hand-derived
gradients) is also compilable for the host.  This test builds it with g++ and
checks it against the oracle (forward) and torch autograd of the oracle
(gradients) — catching formula errors without a GPU.  It exercises the same
source lines the kernels run, but none of the kernel scaffolding."""
import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
from torch.testing import assert_close

from oracle import besskge_oracle as O

ROOT = Path(__file__).resolve().parents[1]
FAM_ID = {"TransE": 0, "RotatE": 1, "DistMult": 2, "ComplEx": 3, "PairRE": 4, "BoxE": 5,
          "TripleRE": 6, "InterHT": 7, "TranS": 8}
EW = {"TransE": 1, "RotatE": 2, "DistMult": 1, "ComplEx": 2, "PairRE": 1, "BoxE": 2, "TripleRE": 1,
      "InterHT": 2, "TranS": 2}
RW = {"TransE": lambda d: d, "RotatE": lambda d: d, "DistMult": lambda d: d,
      "ComplEx": lambda d: 2 * d, "PairRE": lambda d: 2 * d, "BoxE": lambda d: 4 * d + 2,
      "TripleRE": lambda d: 3 * d, "InterHT": lambda d: d, "TranS": lambda d: 3 * d}


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("hostcheck") / "libhostcheck.so"
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", str(out),
                    str(ROOT / "tests" / "hostcheck" / "hostcheck.cpp")], check=True)
    return C.CDLL(str(out))


def fp(t):
    return t.contiguous().data_ptr()


VARIANTS = [
    ("TransE", dict(p=1)), ("TransE", dict(p=2)), ("RotatE", dict(p=1)), ("RotatE", dict(p=2)),
    ("DistMult", dict(p=2)), ("ComplEx", dict(p=2)),
    ("PairRE", dict(p=1)), ("PairRE", dict(p=2)), ("PairRE", dict(p=2, normalize=False)),
    ("BoxE", dict(p=1)), ("BoxE", dict(p=2)), ("BoxE", dict(p=2, apply_tanh=False)),
    ("TripleRE", dict(p=1)), ("TripleRE", dict(p=2, rel_u=0.5)),
    ("TripleRE", dict(p=1, normalize=False, rel_u=1.25)),
    ("InterHT", dict(p=1, rel_u=1.0)), ("InterHT", dict(p=2, rel_u=0.5)),
    ("InterHT", dict(p=2, normalize=False, rel_u=1.0)),
    ("TranS", dict(p=1, rel_u=1.0)), ("TranS", dict(p=2, rel_u=0.25)),
    ("TranS", dict(p=1, normalize=False, rel_u=1.0)),
]


def cfgs(fam, v, d):
    o = dict(family=fam, d=d, norm_p=v["p"], normalize=v.get("normalize", True),
             apply_tanh=v.get("apply_tanh", True), per_dim=v.get("per_dim", True), eps=1e-6,
             rel_u=v.get("rel_u", 0.0))
    c_args = (FAM_ID[fam], v["p"], d, int(o["normalize"]), int(o["apply_tanh"]), int(o["per_dim"]),
              C.c_float(1e-6))
    return o, c_args


def data(fam, d, n, nc, seed=0):
    g = torch.Generator().manual_seed(seed)
    W, Wr = EW[fam] * d, RW[fam](d)
    h = torch.randn(n, W, generator=g)
    t = torch.randn(n, W, generator=g)
    rel = torch.randn(5, Wr, generator=g)
    r = torch.randint(5, (n,), generator=g, dtype=torch.int32)
    cs = torch.randn(nc, W, generator=g)
    cp = torch.randn(n, nc, W, generator=g)
    return h, t, rel, r, cs, cp


@pytest.mark.parametrize("fam,v", VARIANTS + [("BoxE", dict(p=1, per_dim=False))])
def test_triple_fwd_bwd(lib, fam, v):
    d, n = 16, 11
    o, ca = cfgs(fam, v, d)
    lib.hc_set_rel_u(C.c_float(o["rel_u"]))
    h, t, rel, r, _, _ = data(fam, d, n, 3)
    out = torch.empty(n)
    lib.hc_triple_fwd(*ca, n, C.c_void_p(fp(h)), C.c_void_p(fp(rel)), C.c_void_p(fp(r)),
                      C.c_void_p(fp(t)), C.c_void_p(fp(out)))
    hh, tt, rr = (x.clone().requires_grad_(True) for x in (h, t, rel))
    ref = O.score_triple(o, hh, rr, r, tt)
    assert_close(out, ref.detach(), rtol=2e-5, atol=2e-5)
    g = torch.randn(n, generator=torch.Generator().manual_seed(1))
    ref.backward(g)
    dh, dt = torch.empty_like(h), torch.empty_like(t)
    dr_rows = torch.zeros(n, rel.shape[1])
    lib.hc_triple_bwd(*ca, n, C.c_void_p(fp(h)), C.c_void_p(fp(rel)), C.c_void_p(fp(r)),
                      C.c_void_p(fp(t)), C.c_void_p(fp(g)), C.c_void_p(fp(dh)),
                      C.c_void_p(fp(dr_rows)), C.c_void_p(fp(dt)))
    dr = torch.zeros_like(rel).index_add_(0, r.long(), dr_rows)
    assert_close(dh, hh.grad, rtol=2e-4, atol=2e-5)
    assert_close(dt, tt.grad, rtol=2e-4, atol=2e-5)
    assert_close(dr, rr.grad, rtol=2e-4, atol=5e-5)


@pytest.mark.parametrize("fam,v", VARIANTS)
@pytest.mark.parametrize("mode", ["t", "h"])
@pytest.mark.parametrize("shared", [True, False])
def test_candidates_fwd_bwd(lib, fam, v, mode, shared):
    d, n, nc = 16, 7, 5
    o, ca = cfgs(fam, v, d)
    lib.hc_set_rel_u(C.c_float(o["rel_u"]))
    h, t, rel, r, cs, cp = data(fam, d, n, nc, seed=3)
    fixed = h if mode == "t" else t
    cand = cs if shared else cp
    out = torch.empty(n, nc)
    m = 0 if mode == "t" else 1
    lib.hc_candidates_fwd(*ca, m, n, C.c_void_p(fp(fixed)), C.c_void_p(fp(rel)), C.c_void_p(fp(r)),
                          C.c_void_p(fp(cand)), nc, int(shared), C.c_void_p(fp(out)))
    ff, cc, rr = (x.clone().requires_grad_(True) for x in (fixed, cand, rel))
    ref = O.score_candidates(o, mode, ff, rr, r, cc.unsqueeze(0) if shared else cc, shared)
    assert_close(out, ref.detach(), rtol=2e-5, atol=2e-5)
    g = torch.randn(n, nc, generator=torch.Generator().manual_seed(2))
    ref.backward(g)
    d_fixed = torch.empty_like(fixed)
    dr_rows = torch.zeros(n, rel.shape[1])
    d_cand = torch.empty_like(cand)
    lib.hc_candidates_bwd(*ca, m, n, C.c_void_p(fp(fixed)), C.c_void_p(fp(rel)), C.c_void_p(fp(r)),
                          C.c_void_p(fp(cand)), nc, int(shared), C.c_void_p(fp(g)),
                          C.c_void_p(fp(d_fixed)), C.c_void_p(fp(dr_rows)), C.c_void_p(fp(d_cand)))
    dr = torch.zeros_like(rel).index_add_(0, r.long(), dr_rows)
    assert_close(d_fixed, ff.grad, rtol=2e-4, atol=2e-5)
    assert_close(d_cand, cc.grad, rtol=2e-4, atol=2e-5)
    assert_close(dr, rr.grad, rtol=2e-4, atol=5e-5)
