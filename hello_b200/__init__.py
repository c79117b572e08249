"""hello_b200 -- B200 (sm_100a) forward of HELLO's mixture-of-experts variant-calling DNN.

One hot path of anands-repo/hello, rebuilt from scratch: the batched MoE forward over ragged per-site,
per-allele read feature tensors (reference: python/MixtureOfExpertsAdvanced.py:71-252, 487-589), behind the
reference's own Python call surface.  See DESIGN.md.
"""
from . import arch, weights, synth  # noqa: F401

__all__ = ["arch", "weights", "synth", "load_wrapper"]


def load_wrapper(path, device="cuda:0", precision="bf16x3", reference_python=None, **engine_kwargs):
    """``network = hello_b200.load_wrapper(path)`` replaces ``network = torch.load(path, map_location='cpu')``
    (python/caller_calling.py:863): see hello_b200.model.load_wrapper."""
    from .model import load_wrapper as _load
    return _load(path, device=device, precision=precision, reference_python=reference_python, **engine_kwargs)
