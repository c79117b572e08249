"""The C ABI bound from plain C (examples/c_host.c): include/hello_moe.h is valid C, the program links against
libhello_moe.so with nothing but the CUDA runtime, and -- on a GPU -- produces exactly what the Python host gets."""
import os
import shutil
import struct
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def build_c_host(out_path):
    import __graft_entry__ as g
    g.build()
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = [shutil.which("gcc") or "gcc", "-O2", "-Wall", "-Werror", "-std=c11", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(cuda, "include"), os.path.join(ROOT, "examples", "c_host.c"), "-o", out_path,
           "-L", os.path.join(ROOT, "hello_b200"), "-lhello_moe", "-L", os.path.join(cuda, "lib64"), "-lcudart",
           "-Wl,-rpath," + os.path.join(ROOT, "hello_b200")]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr[-2000:]
    return out_path


def test_header_is_plain_c_and_the_host_links(tmp_path):
    exe = build_c_host(str(tmp_path / "c_host"))
    assert os.path.getsize(exe) > 0
    # no GPU needed to see that it refuses a bad case file before touching CUDA
    bad = tmp_path / "bad.bin"
    bad.write_bytes(b"NOTACASE" + b"\0" * 64)
    proc = subprocess.run([exe, str(bad), str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert proc.returncode == 1 and "bad case file" in proc.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("config", ["single_tech", "hybrid_ensemble2"])
def test_c_host_matches_python_host(tmp_path, config):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import export_case
    from hello_b200 import model
    exe = build_c_host(str(tmp_path / "c_host"))
    case, out = str(tmp_path / "case.bin"), str(tmp_path / "out.bin")
    cfg, params, pl = export_case.export(case, config, sites=40, coverage=12, precision="bf16x3", seed=7)
    proc = subprocess.run([exe, case, out], capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    assert "sites" in proc.stdout
    raw = open(out, "rb").read()
    S, A, P = struct.unpack_from("<3q", raw, 0)
    off = 24
    def take(n, dt):
        nonlocal off
        a = np.frombuffer(raw, dtype=dt, count=n, offset=off)
        off += a.nbytes
        return a
    logits, meta, pp = take(3 * A, np.float32).reshape(3, A), take(3 * S, np.float32).reshape(S, 3), take(4 * P, np.float32).reshape(4, P)
    best_pair, best_prob, qual = take(2 * S, np.int32).reshape(S, 2), take(S, np.float32), take(5 * S, np.float64).reshape(S, 5)
    assert off == len(raw)
    eng = model.MoEEngine(cfg, params, device="cuda:0", precision="bf16x3")
    r = eng.run(model.DeviceBatch.from_pileups(pl, "cuda:0"))
    torch.cuda.synchronize()
    assert (S, A, P) == (pl.n_sites, pl.n_alleles, r.pair_prob.shape[1])
    assert np.array_equal(logits, r.logits.cpu().numpy()) and np.array_equal(meta, r.meta.cpu().numpy())
    assert np.array_equal(pp, r.pair_prob.cpu().numpy()) and np.array_equal(best_pair, r.best_pair.cpu().numpy())
    assert np.array_equal(best_prob, r.best_prob.cpu().numpy()) and np.array_equal(qual, r.call_qual.cpu().numpy())
