"""besskge_b200 — B200-native implementation of the BESS sharded KGE step
(graphcore-research/bess-kge hot path) behind the reference's Python API.

Host side (numpy, bit-exact with the reference): `sharding`, `negative_sampler`,
`batch_sampler`, `embedding`, `dataset`.
Device side (hand-written sm_100a CUDA through the C-ABI in
`include/besskge_b200.h`): `scoring`, `loss`, `metric`, `bess`.
There is no CPU fallback for the device side: importing works anywhere, but
calling a device entry point without the built library or on CPU tensors
raises `BessLibraryError`.
"""
from . import (  # noqa: F401
    batch_sampler,
    bess,
    dataset,
    embedding,
    loss,
    metric,
    negative_sampler,
    optim,
    scoring,
    sharding,
)
from ._lib import BessLibraryError, library_path, load  # noqa: F401

__version__ = "0.1.0"


def load_custom_ops_so() -> None:
    """Counterpart of reference `besskge.load_custom_ops_so`
    (besskge/__init__.py:10-37): loads the compiled kernel library and fails
    loudly when it is missing."""
    load()
