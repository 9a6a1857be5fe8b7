"""TEST INFRASTRUCTURE — not product code.

Loads the UNMODIFIED reference (`/root/reference/besskge/*.py`) as a CPU oracle
inside this container.  Only `tests/golden/make_golden.py` and the CPU tests
that are skipped when `/root/reference` is absent may call this; nothing in
`besskge_b200/`, `bench.py` or the `-m gpu` tests does (the reference does not
exist on the GPU box).

Why stubs are needed: `besskge/__init__.py:37` dlopens a PopART-linked `.so`,
and `bess.py:12-19`, `scoring.py:11`, `batch_sampler.py:14`, `dataset.py:17`
import `poptorch`, `poptorch_experimental_addons` and `ogb`, none of which are
installed.  We register a synthetic `besskge` package whose `__path__` points
at the reference directory (so its `__init__` is skipped) and provide three
tiny stand-ins for the missing third-party modules.  Every reference file is
imported byte-for-byte.

Third-party semantics restated here (source not under /root/reference;
pinned `poptorch-experimental-addons @ 899aec4a`, requirements.txt:5):
  * `pea.distance_matrix(a, b, p)`   -> pairwise p-distance  [S,D]x[N,D]->[S,N]
  * `all_to_all_single_cross_replica(x[n,...], n)` on replica r:
        out[j] = x_j[r]            (x_j = the tensor held by replica j)
  * `all_gather_cross_replica(x, n)` -> stack_j x_j  (new leading axis)
  * `poptorch.identity_loss(x, reduction="none")` -> x
  * `poptorch.for_loop(n, body, xs)` -> plain iteration
"""
from __future__ import annotations

import copy
import sys
import threading
import types
from pathlib import Path
from typing import Any, Callable, Dict, List, Optional

import torch

REFERENCE_ROOT = Path("/root/reference")


def reference_available() -> bool:
    return (REFERENCE_ROOT / "besskge" / "bess.py").exists()


class LockstepExchange:
    """Cross-replica collectives for one-python-thread-per-replica emulation.

    Results are built with torch.stack so autograd flows across replicas.
    """

    def __init__(self, n: int) -> None:
        self.n = n
        self.slots: List[Optional[torch.Tensor]] = [None] * n
        self.barrier = threading.Barrier(n) if n > 1 else None
        self.tl = threading.local()

    def _rank(self) -> int:
        return getattr(self.tl, "rank", 0)

    def all_to_all(self, x: torch.Tensor, n: int) -> torch.Tensor:
        if self.n == 1:
            return x
        r = self._rank()
        self.slots[r] = x
        self.barrier.wait()
        out = torch.stack([self.slots[j][r] for j in range(self.n)])
        self.barrier.wait()
        return out

    def all_gather(self, x: torch.Tensor, n: int) -> torch.Tensor:
        if self.n == 1:
            return x.unsqueeze(0)
        r = self._rank()
        self.slots[r] = x
        self.barrier.wait()
        out = torch.stack(list(self.slots))
        self.barrier.wait()
        return out


_EXCHANGE: Dict[str, LockstepExchange] = {"ex": LockstepExchange(1)}


def set_n_replica(n: int) -> LockstepExchange:
    _EXCHANGE["ex"] = LockstepExchange(n)
    return _EXCHANGE["ex"]


def _install_stubs() -> None:
    pt = types.ModuleType("poptorch")
    pt.identity_loss = lambda x, reduction="none": x

    def _for_loop(n: int, body: Callable[..., Any], xs: List[Any]) -> List[Any]:
        for _ in range(n):
            xs = list(body(*xs))
        return xs

    pt.for_loop = _for_loop
    pt.ipuHardwareIsAvailable = lambda num_ipus=1: False

    # --- PopTorch runtime used by pipeline.py:130-144 (inferenceModel over n replicas,
    # deviceIterations, OutputMode.All, async DataLoader): restated as "run the module once
    # per (step, replica) on the (1, ...) slice of every input and concatenate the outputs"
    class _Options:
        def __init__(self) -> None:
            self.replication_factor = 1
            self.device_iterations = 1

        def deviceIterations(self, n: int) -> "_Options":
            self.device_iterations = int(n)
            return self

        def outputMode(self, *_a: Any) -> "_Options":
            return self

        def useIpuModel(self, *_a: Any) -> "_Options":
            return self

    class _Var:
        def replicaGrouping(self, *_a: Any) -> None:
            return None

    class _InferenceModel:
        def __init__(self, module: torch.nn.Module, options: "_Options") -> None:
            self.module = module
            self.options = options
            self.entity_embedding = _Var()

        def __call__(self, **inp: torch.Tensor) -> Any:
            n = int(self.options.replication_factor)
            bps = int(self.options.device_iterations)
            batch = {k: v.reshape(bps, n, *v.shape[1:]) for k, v in inp.items()}
            res = run_replicated(self.module, batch, n, bps)
            return res["__tensor__"] if "__tensor__" in res else res

    def _dataloader(options: Any = None, dataset: Any = None, batch_size: Any = None,
                    sampler: Any = None, **_kw: Any) -> Any:
        return torch.utils.data.DataLoader(dataset, batch_size=None, sampler=sampler)

    pt.Options = _Options
    pt.DataLoader = _dataloader
    pt.inferenceModel = lambda module, options=None: _InferenceModel(module, options)
    pt.OutputMode = types.SimpleNamespace(All=0)
    pt.CommGroupType = types.SimpleNamespace(NoGrouping=0)
    pt.VariableRetrievalMode = types.SimpleNamespace(OnePerGroup=0)
    pt.DataLoaderMode = types.SimpleNamespace(Async=0)
    pt.SharingStrategy = types.SimpleNamespace(SharedMemory=0)

    pea = types.ModuleType("poptorch_experimental_addons")
    pea.distance_matrix = lambda a, b, p: _pairwise_distance(a, b, p)
    col = types.ModuleType("poptorch_experimental_addons.collectives")
    col.all_gather_cross_replica = lambda x, n: _EXCHANGE["ex"].all_gather(x, n)
    col.all_to_all_single_cross_replica = lambda x, n: _EXCHANGE["ex"].all_to_all(x, n)
    pea.collectives = col

    ogb = types.ModuleType("ogb")
    ogb_lp = types.ModuleType("ogb.linkproppred")
    ogb.linkproppred = ogb_lp

    sys.modules.update(
        {
            "poptorch": pt,
            "poptorch_experimental_addons": pea,
            "poptorch_experimental_addons.collectives": col,
            "ogb": ogb,
            "ogb.linkproppred": ogb_lp,
        }
    )
    pkg = types.ModuleType("besskge")
    pkg.__path__ = [str(REFERENCE_ROOT / "besskge")]
    sys.modules["besskge"] = pkg

    # CPU-torch compatibility shim: loss.py:239-247 passes an int32 class target
    # to cross_entropy, which PopTorch accepts but CPU torch rejects ("expected
    # scalar type Long").  Cast the target; the arithmetic is untouched.
    import torch.nn.functional as F

    if not getattr(F.cross_entropy, "_bess_shim", False):
        _orig_ce = F.cross_entropy

        def _ce(input, target, *a, **k):
            if target.dtype == torch.int32:
                target = target.long()
            return _orig_ce(input, target, *a, **k)

        _ce._bess_shim = True
        F.cross_entropy = _ce


def _pairwise_distance(a: torch.Tensor, b: torch.Tensor, p: int) -> torch.Tensor:
    # Direct (non norm-expanded) evaluation so that the fixtures are free of the
    # cancellation error of the mm-based cdist path.
    return torch.norm(a.unsqueeze(1) - b.unsqueeze(0), p=p, dim=-1)


_LOADED: Dict[str, Any] = {}


def load_reference() -> types.SimpleNamespace:
    """Import the reference modules; returns a namespace of modules."""
    if "ns" in _LOADED:
        return _LOADED["ns"]
    if not reference_available():
        raise RuntimeError("/root/reference is not present on this machine")
    _install_stubs()
    import importlib

    names = [
        "utils",
        "dataset",
        "sharding",
        "embedding",
        "negative_sampler",
        "batch_sampler",
        "loss",
        "metric",
        "scoring",
        "bess",
        "pipeline",
    ]
    ns = types.SimpleNamespace()
    for nm in names:
        setattr(ns, nm, importlib.import_module(f"besskge.{nm}"))
    _LOADED["ns"] = ns
    return ns


def run_replicated(
    model: torch.nn.Module,
    batch: Dict[str, torch.Tensor],
    n_shard: int,
    batches_per_step: int,
    grad: bool = False,
) -> Dict[str, Any]:
    """What PopTorch does implicitly: run `model` once per (step, replica).

    `batch[k]` carries the host layout (bps, n_shard, ...); replica r at step s
    sees the (1, ...) slice `flatten(0,1)[s*n + r]` and `entity_embedding[r]`.
    Outputs are concatenated in (step, shard) order (OutputMode.All).
    With grad=True the summed loss is returned un-detached so the caller can
    call backward() once and read `model.entity_embedding.grad` ([n,Es,D]).
    """
    ex = set_n_replica(n_shard)
    table = model.entity_embedding  # [n, Es, D] Parameter
    flat = {k: v.flatten(end_dim=1) for k, v in batch.items()}
    results: List[List[Optional[Dict[str, Any]]]] = [
        [None] * n_shard for _ in range(batches_per_step)
    ]
    errors: List[BaseException] = []

    def make_replica(r: int) -> torch.nn.Module:
        rep = copy.copy(model)
        rep._parameters = dict(model._parameters)
        rep._modules = dict(model._modules)
        rep._parameters.pop("entity_embedding", None)
        rep.__dict__["entity_embedding"] = table[r]
        sf = copy.copy(model.score_fn)
        sf._parameters = dict(model.score_fn._parameters)
        sf._parameters.pop("entity_embedding", None)
        sf.__dict__["entity_embedding"] = table[r]
        rep._modules["score_fn"] = sf
        return rep

    def worker(r: int) -> None:
        try:
            ex.tl.rank = r
            rep = make_replica(r)
            for s in range(batches_per_step):
                i = s * n_shard + r
                kwargs = {k: v[i : i + 1] for k, v in flat.items()}
                if "triple_weight" not in kwargs and hasattr(rep, "loss_fn"):
                    kwargs["triple_weight"] = torch.tensor([1.0])
                with torch.set_grad_enabled(grad):
                    res = rep(**kwargs)
                    # AllScoresBESS.forward returns a bare tensor (bess.py:1062)
                    results[s][r] = res if isinstance(res, dict) else {"__tensor__": res}
        except BaseException as e:  # pragma: no cover
            errors.append(e)
            if ex.barrier is not None:
                ex.barrier.abort()

    if n_shard == 1:
        worker(0)
    else:
        ts = [threading.Thread(target=worker, args=(r,)) for r in range(n_shard)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
    if errors:
        raise errors[0]

    out: Dict[str, Any] = {}
    keys = results[0][0].keys()
    for k in keys:
        vals = [results[s][r][k] for s in range(batches_per_step) for r in range(n_shard)]
        if k == "loss":
            out[k] = torch.stack([v.reshape(()) for v in vals])
        else:
            out[k] = torch.cat(vals, dim=0)
    return out
