"""bench.py's own arm on a small workload: one JSON line carrying every key of the measurement contract."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_line_contract():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--sites", "4096", "--steps", "2", "--warmup", "3",
                          "--no-cpu-baseline", "--other-sites", "2048"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-800:]
    lines = out.stdout.splitlines()
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "gpu_launches", "e2e", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["metric"] == "candidate_sites_per_sec" and d["unit"] == "sites/s" and d["scaling"] == "weak"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["vs_baseline"] is None
    assert d["value"] > 0 and d["gpu_launches"] >= 2 * 6                       # six kernels per chunk in the fused mode
    assert d["config"]["workload"].startswith("illumina_30x") and "model" not in d["config"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 4096 * 20 * 900 and e["d2h_bytes_per_step"] > 0
    assert e["value"] != d["value"]
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and 0 < r["frac"] < 1 / 3 + 1e-6
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    # BASELINE configs 3-5 ride along without touching the headline fields
    ow = d["other_workloads"]
    assert set(ow) == {"pacbio_hp_30x", "hybrid_no_ensemble_30x", "hybrid_ensemble2_30x", "wgs_ragged_15_60x",
                       "illumina_30x_softplus", "hybrid_no_ensemble_wide_30x"}
    for name, w in ow.items():
        assert "error" not in w, (name, w)
        assert w["sites"] == (256 if name == "hybrid_no_ensemble_wide_30x" else 2048)       # the wide model: 1/8 of the sites
        assert w["sites_per_sec"] > 0 and 0 < w["read_convolver_stage_share"] < 1
        assert w["max_abs_dlogit_vs_oracle"] < 1e-3 and w["oracle_sample_sites"] == 32


def test_bench_balanced_partition_single_rank():
    """--partition balanced on one rank: the whole deterministic dataset is this rank's shard; the line carries the
    partition record (ranges, cost shares, per-rank forward time)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "wgs_ragged_15_60x", "--partition",
                          "balanced", "--total-sites", "20000", "--steps", "2", "--warmup", "3", "--no-cpu-baseline",
                          "--no-other-workloads", "--no-e2e"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-800:]
    d = json.loads(out.stdout.splitlines()[-1])
    p = d["partition"]
    assert p["mode"] == "balanced" and p["total_sites"] == 20000 and p["site_ranges"] == [[0, 20000]]
    assert abs(sum(p["cost_share"]) - 1.0) < 1e-9 and len(p["rank_forward_ms"]) == 1
    assert d["config"]["sites_per_gpu"] == 20000 and d["scaling"] == "strong"
