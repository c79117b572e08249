"""hello_b200 -- B200 (sm_100a) forward of HELLO's mixture-of-experts variant-calling DNN.

One hot path of anands-repo/hello, rebuilt from scratch: the batched MoE forward over ragged per-site,
per-allele read feature tensors (reference: python/MixtureOfExpertsAdvanced.py:71-252, 487-589), behind the
reference's own Python call surface.  See DESIGN.md.
"""
from . import arch, weights, synth  # noqa: F401

__all__ = ["arch", "weights", "synth"]
