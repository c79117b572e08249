"""`.wrapper.dnn` drop-in (north_star: "keep ... the .wrapper.dnn model load"): the reference pickles the whole
MoEMergedWrapperAdvanced module (python/create_model_wrapper.py:7-10) and python/caller_calling.py:863 loads it with
torch.load.  Here the reference's own modules build such a file for every supported wiring, hello_b200 reads it back on the
CPU (read_wrapper: no GPU needed) and the packed weight blob must be byte-identical to the one packed from the same
parameters directly.  Needs the reference's python modules (oracle/_ref/python or /root/reference/python); skipped
otherwise.  One configuration per subprocess: the reference's architecture modules are mutable singletons."""
import hashlib
import os
import subprocess
import sys

import pytest
import torch

from helpers import params_for
from hello_b200 import arch, weights
from oracle import ref_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ["single_tech", "single_tech_hp", "hybrid_no_ensemble", "hybrid_ensemble2", "hybrid_full", "single_tech_addendum"]

CHILD = r"""
import hashlib, sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, %(root)r)
import torch
from hello_b200 import arch, weights
from oracle import ref_model
name, path = sys.argv[1], sys.argv[2]
cfg = arch.CONFIGS[name]
net = ref_model.build_wrapper(name, weights.init_params(cfg, seed=13), provide_predictions=True)
torch.save(net, path)                                   # python/create_model_wrapper.py:10
del net
from hello_b200 import model
got_cfg, sd, provide = model.read_wrapper(path)         # NNTools is already importable: ref_model put it on sys.path
blob = weights.pack_blob(got_cfg, sd)
print("RESULT", got_cfg.name, provide, hashlib.sha256(blob).hexdigest(), len(sd))
"""


@pytest.mark.skipif(not ref_model.available(), reason="the reference's python modules are not available")
@pytest.mark.parametrize("name", CASES)
def test_wrapper_pickle_round_trip(name, tmp_path):
    path = str(tmp_path / (name + ".wrapper.dnn"))
    out = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}, name, path], capture_output=True, text=True,
                         timeout=600, env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1"))
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][-1].split()
    cfg = arch.CONFIGS[name]
    want = hashlib.sha256(weights.pack_blob(cfg, params_for(cfg))).hexdigest()
    assert line[1] == name and line[2] == "True"
    assert line[3] == want, "blob packed from the unpickled wrapper differs from the blob packed from the parameters"
    assert int(line[4]) == len(weights.param_shapes(cfg))
    assert os.path.getsize(path) > 1_000_000


def test_state_dict_without_weight_norm_is_refused():
    """A model built without weight-norm (plain Conv1d weights, BatchNorm statistics) must be refused, not silently
    stripped of its weights (python/NNTools.py:27-45,84-104; moe_attention_config_single_tech_old_equivalent_layer_norm.py)."""
    cfg = arch.CONFIGS["single_tech"]
    sd = dict(params_for(cfg))
    k = next(iter(sd)).rsplit(".", 1)[0]
    # a materialised weight next to weight_g / weight_v is redundant and dropped
    with_copy = dict(sd)
    with_copy[k + ".weight"] = weights.fold_weight_norm(sd[k + ".weight_v"], sd[k + ".weight_g"])
    assert set(weights.weight_norm_state(with_copy)) == set(sd)
    assert weights.cfg_from_state_dict(weights.weight_norm_state(with_copy)).name == "single_tech"
    plain = {kk: v for kk, v in sd.items() if not kk.startswith(k + ".weight_")}
    plain[k + ".weight"] = with_copy[k + ".weight"]
    with pytest.raises(ValueError, match="without weight-norm"):
        weights.weight_norm_state(plain)
    bn = dict(sd)
    bn["read_convolver0.network.1.running_mean"] = torch.zeros(16)
    with pytest.raises(ValueError, match="without weight-norm"):
        weights.weight_norm_state(bn)


def test_load_wrapper_reports_missing_reference_modules(tmp_path):
    """Without NNTools importable the pickle cannot be read: a clear error, not an AttributeError from the unpickler."""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from hello_b200 import model, _lib\n"
            "try:\n    model.read_wrapper('/nonexistent.wrapper.dnn')\n"
            "except _lib.HelloMoEError as e:\n    print('OK', 'NNTools' in str(e))\n") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
    assert out.stdout.strip() == "OK True", out.stdout + out.stderr[-500:]


def test_legacy_wiring_state_dict_maps_onto_single_tech():
    """MoEMergedAdvanced (legacy wiring, python/MixtureOfExpertsAdvanced.py:255-484) in its single-technology form has the
    parameters of MoEAttention in the same order under other names; tests/golden/legacy_single_tech.npz was produced by the
    reference's createMoEFullMergedAdvancedModel with these very parameters, and the oracle / CUDA forward of `single_tech`
    reproduce it (test_oracle_golden, test_forward_matches_reference_golden).  (The hybrid legacy wirings have reference-made
    golden files of their own: legacy_hybrid_additive, legacy_hybrid_combiners.)"""
    cfg = arch.CONFIGS["single_tech"]
    params = params_for(cfg)
    legacy = {}
    for k in weights.param_shapes(cfg):                       # the reference's registration order: bias, weight_g, weight_v
        v = params[k]
        net, rest = k.split(".", 1)
        legacy[{"read_convolver0": "readConv0", "compressor0": "alleleConv0", "xattn0": "expert0"}[net] + "." + rest] = v
    # the live xattn has two parameter-free front-end modules: the legacy expert's Sequential slots are shifted by two
    legacy = {(k.replace("expert0.network.%d." % s, "expert0.network.%d." % (s - 2)) if k.startswith("expert0.") else k): v
              for k, v in legacy.items() for s in [int(k.split(".")[2])]}
    back = weights.legacy_state_to_attention(legacy)
    assert list(back.keys()) == list(weights.param_shapes(cfg).keys())
    assert all(torch.equal(back[k], params[k]) for k in params)
    assert weights.cfg_from_state_dict(back).name == "single_tech"
    assert weights.legacy_state_to_attention(params) == dict(params)               # non-legacy dicts pass through
    combined = dict(legacy)
    combined["alleleConvCombiner.network.0.conv1d.bias"] = torch.zeros(16)      # one ConvCombiner without the other is refused
    with pytest.raises(ValueError, match="ConvCombiner"):
        weights.legacy_state_to_attention(combined)
    broken = dict(legacy)
    broken["readConv1.network.0.conv1d.bias"] = torch.zeros(16)                # an incomplete second technology
    with pytest.raises(ValueError, match="does not match"):
        weights.legacy_state_to_attention(broken)


CHILD_BN = r"""
import hashlib, importlib, sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, %(root)r)
import torch
from hello_b200 import arch, weights, _lib
from oracle import ref_model
M = ref_model.import_reference()
kind, path = sys.argv[1], sys.argv[2]
if kind == "batchnorm":
    import architectures.read_convolver as rc, architectures.compressor_conv_small as cc, architectures.xattn_subtract as xa
    for m in (rc, cc, xa):
        m.weight_norm = False
        m.gen_config()
    moe = M.create_moe_attention_model({"read_conv0": rc.config, "compressor0": cc.config, "xattn0": xa.config}).eval()
    state = weights.init_batchnorm_state([(k, tuple(v.shape)) for k, v in moe.state_dict().items()], seed=13)
    moe.load_state_dict(state)
else:                                                   # the Softplus / no-normalisation configuration
    cfgd = importlib.import_module("moe_attention_config_single_tech_old_equivalent_layer_norm").configDict
    moe = M.create_moe_attention_model(cfgd).eval()
    state = weights.init_batchnorm_state([(k, tuple(v.shape)) for k, v in moe.state_dict().items()], seed=13)
    moe.load_state_dict(state)
torch.save(M.createMoEFullMergedAdvancedModelWrapper(moe.eval()).eval(), path)
from hello_b200 import model
try:
    cfg, sd, provide = model.read_wrapper(path)
    print("RESULT", cfg.name, hashlib.sha256(weights.pack_blob(cfg, sd)).hexdigest())
except _lib.HelloMoEError as e:
    print("REFUSED", str(e)[:160].replace("\n", " "))
"""


@pytest.mark.skipif(not ref_model.available(), reason="the reference's python modules are not available")
def test_batchnorm_wrapper_is_folded_and_softplus_model_is_recognised(tmp_path):
    """A .wrapper.dnn of a model built without weight-norm loads (BatchNorm1d folded, same blob as folding the state dict
    directly); the reference's Softplus configuration (moe_attention_config_single_tech_old_equivalent_layer_norm.py) is
    recognised by its modules -- the state dict alone would read as the ReLU model -- and packs Softplus layer records."""
    from helpers import batchnorm_params
    run = lambda kind: subprocess.run([sys.executable, "-c", CHILD_BN % {"root": ROOT}, kind, str(tmp_path / (kind + ".dnn"))],
                                      capture_output=True, text=True, timeout=600, env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1"))
    out = run("batchnorm")
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][-1].split()
    _, params = batchnorm_params()
    assert line[1] == "single_tech"
    assert line[2] == hashlib.sha256(weights.pack_blob(arch.CONFIGS["single_tech"], params)).hexdigest()
    out = run("softplus")
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][-1].split()
    assert line[1] == "single_tech_softplus"
    cfg_sp = arch.CONFIGS["single_tech_softplus"]
    _, params_sp = batchnorm_params("single_tech_softplus")          # same deterministic state as the child built
    assert line[2] == hashlib.sha256(weights.pack_blob(cfg_sp, params_sp)).hexdigest()
    assert line[2] != hashlib.sha256(weights.pack_blob(arch.CONFIGS["single_tech"], params_sp)).hexdigest()
