"""The reference's own model modules, unmodified, as a CPU baseline.  TEST / MEASUREMENT INFRASTRUCTURE: only tests/,
bench.py's CPU legs and oracle/ scripts import this; nothing under hello_b200/ does.

`oracle/Makefile` copies MixtureOfExpertsAdvanced.py, NNTools.py, Attention.py, architectures/ and the
moe_attention_config_* modules from /root/reference/python into the git-ignored oracle/_ref/python/ (it travels to the GPU
box with the snapshot like oracle/_ref/libref_encoder.so).  This module puts that directory (or /root/reference/python
itself where it exists) on sys.path and builds the reference's `MoEMergedWrapperAdvanced` exactly as the product does:

    create_moe_attention_model(configDict)            python/MixtureOfExpertsAdvanced.py:657-703
    createMoEFullMergedAdvancedModelWrapper(moe)      python/MixtureOfExpertsAdvanced.py:704-707 (what create_model_wrapper.py pickles)
    network.eval(); network.providePredictions = True python/caller_calling.py:863-867
    network(featureDict, ref_segment) under no_grad   python/caller_calling.py:651-652   (one site per call)

One configuration per process: the reference's architecture modules are mutable singletons.
"""
from __future__ import annotations

import importlib
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = (os.path.join(HERE, "_ref", "python"), "/root/reference/python")


def reference_dir():
    for d in CANDIDATES:
        if os.path.exists(os.path.join(d, "MixtureOfExpertsAdvanced.py")):
            return d
    return None


def available() -> bool:
    return reference_dir() is not None


def import_reference():
    """-> the reference's MixtureOfExpertsAdvanced module (NNTools is imported by it and patches torch.nn)."""
    d = reference_dir()
    if d is None:
        raise ImportError("the reference's python modules are neither under oracle/_ref/python (make -C oracle) nor "
                          "under /root/reference/python")
    if d not in sys.path:
        sys.path.insert(0, d)
    sys.dont_write_bytecode = True
    warnings.filterwarnings("ignore")                    # torch.nn.utils.weight_norm deprecation
    return importlib.import_module("MixtureOfExpertsAdvanced")


def build_moe(cfg_name: str, params=None):
    """The reference's MoEAttention for one of hello_b200.arch.CONFIGS, optionally with `params` loaded."""
    import types
    from hello_b200 import arch
    M = import_reference()
    if cfg_name in arch.REFERENCE_ADDENDUM_MODULE:
        # transfer-learning model: build_on_top stacks the addendum networks on a base model
        # (python/MixtureOfExpertsAdvancedXferLearning.py:94-183)
        X = importlib.import_module("MixtureOfExpertsAdvancedXferLearning")
        base_name, add_module = arch.REFERENCE_ADDENDUM_MODULE[cfg_name]
        base = M.create_moe_attention_model(importlib.import_module(arch.REFERENCE_CONFIG_MODULE[base_name]).configDict)
        add = importlib.import_module(add_module).configDict
        holder = types.SimpleNamespace(module=types.SimpleNamespace(dnn=base))
        moe, _ = X.build_on_top(holder, **{k: X.make_network(add, k) for k in add})
    else:
        moe = M.create_moe_attention_model(importlib.import_module(arch.REFERENCE_CONFIG_MODULE[cfg_name]).configDict)
    moe = moe.eval()
    if params is not None:
        moe.load_state_dict(params)
    return moe


def build_wrapper(cfg_name: str, params=None, provide_predictions: bool = True):
    """The object caller_calling.py gets from torch.load(...): MoEMergedWrapperAdvanced in eval mode."""
    M = import_reference()
    net = M.createMoEFullMergedAdvancedModelWrapper(build_moe(cfg_name, params)).eval()
    net.providePredictions = provide_predictions
    return net
