"""CPU-side checks: the C-ABI library loads and exports every symbol include/hello_moe.h declares, the ctypes
structs match the header, the product path refuses to run without a GPU, and the host-side packing logic."""
import ctypes
import os
import re
import struct

import numpy as np
import pytest
import torch

from helpers import params_for
from hello_b200 import _lib, arch, synth, weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hello_moe.h")


def header_functions():
    names = set()
    for header in sorted(os.listdir(os.path.join(ROOT, "include"))):           # every include/*.h
        text = open(os.path.join(ROOT, "include", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(hello_(?:moe|encode)_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = header_functions()
    assert len(names) >= 11
    for name in names:
        assert hasattr(lib, name), name
    assert set(names) == set(_lib.EXPORTS), "binding and header disagree on the exported functions"
    lib.hello_moe_abi_version.restype = ctypes.c_int
    assert lib.hello_moe_abi_version() == _lib.ABI_VERSION


def test_ctypes_structs_match_header_layout():
    # hello_cfg: 12 int32; hello_batch: 4 int64 + 2 int32 + 11 pointers; hello_result: 9 pointers
    assert ctypes.sizeof(_lib.HelloCfg) == 12 * 4
    assert ctypes.sizeof(_lib.HelloBatch) == 4 * 8 + 2 * 4 + 11 * 8
    assert ctypes.sizeof(_lib.HelloResult) == 9 * 8
    from hello_b200 import encoder
    assert ctypes.sizeof(encoder.HelloEncodeBatch) == 8 + 2 * 4 + 16 * 8   # include/hello_encode.h
    text = open(HEADER).read()
    for field in ("struct_size", "n_tech", "read_channels", "xattn_present", "has_combiners", "meta_kind",
                  "feature_length", "precision", "max_chunk_sites"):
        assert field in text
    assert [f[0] for f in _lib.HelloCfg._fields_] == ["struct_size", "n_tech", "read_channels", "xattn_present",
                                                      "has_combiners", "meta_kind", "feature_length", "precision",
                                                      "max_chunk_sites"]


def test_create_rejects_bad_arguments_without_touching_a_gpu():
    lib = _lib.load()
    handle = ctypes.c_void_p()
    cfg = _lib.HelloCfg()
    cfg.struct_size = 4                       # wrong size -> HELLO_ERR_ARG before any CUDA call
    buf = (ctypes.c_char * 256)()
    assert lib.hello_moe_create(buf, 256, ctypes.byref(cfg), 0, ctypes.byref(handle)) == -1
    assert b"struct_size" in lib.hello_moe_last_error(None)
    assert lib.hello_moe_launch_count(None) == 0
    assert lib.hello_moe_workspace_bytes(None, 1, 0, 1, 1) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_a_gpu():
    from hello_b200 import model
    cfg = arch.CONFIGS["single_tech"]
    with pytest.raises(_lib.HelloMoEError):
        model.MoEEngine(cfg, params_for(cfg))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "hello_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
    # outside tests/ the oracle is imported only by smoke() and by the CPU-baseline / checker legs of the benchmarks
    allowed = {"__graft_entry__.py", "bench.py", os.path.join("tools", "bench_encoder.py"),
               os.path.join("tools", "bench_serving.py")}        # its reference-latency leg times oracle/_ref/python
    for rel in ["__graft_entry__.py", "bench.py"] + [os.path.join("tools", f) for f in os.listdir(os.path.join(ROOT, "tools"))]:
        if rel.endswith(".py") and rel not in allowed:
            text = open(os.path.join(ROOT, rel)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), rel


@pytest.mark.parametrize("name", sorted(arch.CONFIGS))
def test_blob_roundtrip(name):
    """pack_blob: header, layer records and folded weights are where hello_moe_create expects them."""
    cfg = arch.CONFIGS[name]
    params = params_for(cfg)
    blob = weights.pack_blob(cfg, params)
    magic, version, n_slots, rec_off, n_rec, data_off, n_floats = struct.unpack_from("<8sIIQQQQ", blob, 0)
    assert magic == weights.BLOB_MAGIC and version == 1 and n_slots == 10
    assert rec_off == 128 and data_off == rec_off + n_rec * 128 and data_off % 16 == 0
    assert len(blob) == data_off + 4 * n_floats
    first = struct.unpack_from("<10I", blob, 48)
    count = struct.unpack_from("<10I", blob, 88)
    nets = cfg.networks()
    for net, nid in weights.NET_IDS.items():
        n_layers = len([l for l in nets.get(net, []) if not isinstance(l, arch.Front)])
        assert count[nid] == n_layers
    # first conv of read_convolver0: folded weight g*v/|v| stored as [k*cin, cout]
    rec = np.frombuffer(blob, np.int32, 32, rec_off + first[0] * 128)
    cin, cout, k = rec[2], rec[3], rec[4]
    data = np.frombuffer(blob, np.float32, n_floats, data_off)
    w = data[rec[8]:rec[8] + k * cin * cout].reshape(k * cin, cout)
    key = cfg.keyed("read_convolver0")[0][0] + ".conv1d"          # "read_convolver0.0.network.0..." with an addendum
    v = params[key + ".weight_v"].numpy().astype(np.float64)
    g = params[key + ".weight_g"].numpy().astype(np.float64)
    ref = g * v / np.sqrt((v ** 2).sum(axis=(1, 2), keepdims=True))
    np.testing.assert_allclose(w, ref.transpose(2, 1, 0).reshape(k * cin, cout), rtol=2e-6, atol=1e-7)


def test_cfg_from_state_dict_recognises_every_wiring():
    for name, cfg in arch.CONFIGS.items():
        assert weights.cfg_from_state_dict(params_for(cfg), softplus_nets=cfg.softplus_nets).name == name


def test_synthetic_pileups_follow_the_codebook():
    pl = synth.make_pileups(50, coverage=12, channels=(7,), seed=5)
    r = pl.reads[0]
    assert r.dtype == torch.uint8 and r.shape[1:] == (150, 7)
    assert set(np.unique(r[..., 0].numpy())) <= {0, 30, 100, 180, 250}
    assert set(np.unique(r[..., 4].numpy())) <= {0, 70, 240}
    assert set(np.unique(r[..., 6].numpy())) <= {0, 120, 240}
    sao, aro = pl.site_allele_off.numpy(), pl.allele_read_off[0].numpy()
    assert sao[0] == 0 and aro[0] == 0 and aro[-1] == r.shape[0] and (np.diff(sao) >= 1).all() and (np.diff(aro) >= 1).all()


def test_encoder_host_packing():
    """pack_sites / row_plan: flat arrays, BAM CIGAR encoding, row order site -> allele -> supporting reads of the
    technology, -1 row for an allele without support (c++/src/AlleleSearcherLiteFiltered.cpp:958-968, 1037-1043)."""
    from hello_b200 import encoder
    s0 = encoder.SitePileup(["ACGT", "TTGCA"], [[30] * 4, [20] * 5], [[(0, 4)], [(4, 1), (0, 2), (1, 1), (0, 1)]], [100, 101],
                            [60, 255], [1, -1], [False, True], [0, 2], "ACGTACGTAC", 95, 101, 102, {"A": [1, 0], "C": [0]})
    s1 = encoder.SitePileup(["GG"], [[7, 8]], [[(0, 2)]], [50], [3], [1], [False], [1], "GGGG", 49, 50, 51, {"G": [0]})
    p = encoder.pack_sites([s0, s1])
    assert p.read_off.tolist() == [0, 4, 9, 11] and p.cigar_off.tolist() == [0, 1, 5, 6]
    assert p.cigars.tolist() == [4 << 4, 1 << 4 | 4, 2 << 4, 1 << 4 | 1, 1 << 4, 2 << 4]
    assert bytes(p.bases) == b"ACGTTTGCAGG" and p.read_base.tolist() == [0, 2, 3] and p.ref_off.tolist() == [0, 10, 14]
    rr, rs, counts = encoder.row_plan([s0, s1], [["A", "C", "T"], ["G"]], False, p.read_base)
    assert rr.tolist() == [0, 0, -1, 2] and rs.tolist() == [0, 0, 0, 1] and counts == [1, 1, 1, 1]
    rr, rs, counts = encoder.row_plan([s0, s1], [["A", "C", "T"], ["G"]], True, p.read_base)
    assert rr.tolist() == [1, -1, -1, -1] and counts == [1, 1, 1, 1]
    with pytest.raises(ValueError):
        encoder.pack_sites([encoder.SitePileup(["AC"], [[1]], [[(0, 2)]], [0], [0], [1], [False], [0], "AC", 0, 0, 1, {})])
    with pytest.raises(ValueError):
        encoder.pack_sites([encoder.SitePileup(["AC"], [[1, 2]], [[(0, 3)]], [0], [0], [1], [False], [0], "AC", 0, 0, 1, {})])


def test_feature_records_match_the_callers_pickle_format():
    """model.feature_records builds the `.features` list of caller_calling.py:746-754 (what prepareVcf.vcfRecords reads;
    tests/golden/final_calls.npz was produced by feeding exactly these records to the reference)."""
    import pickle
    from hello_b200 import model
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "final_calls.npz"))
    names_all = [str(x) for x in g["names"]]
    n_alleles = g["n_alleles"]
    alleles = [[names_all[i] for i in g["allele_name_idx"][s, :n]] for s, n in enumerate(n_alleles)]
    pair_off = np.concatenate([[0], np.cumsum(n_alleles * (n_alleles + 1) // 2)])
    loci = [("chr1", 1000 + 10 * s, len(a[0])) for s, a in enumerate(alleles)]
    recs = pickle.loads(pickle.dumps(model.feature_records(g["experts"], g["meta"], pair_off, alleles, loci)))
    assert len(recs) == len(alleles)
    for s, r in enumerate(recs):
        assert set(r) == {"chromosome", "position", "length", "meta", "expertPredictions"}
        assert (r["chromosome"], r["position"], r["length"]) == loci[s]
        assert isinstance(r["meta"], np.ndarray) and r["meta"].dtype == np.float32 and r["meta"].shape == (3,)
        np.testing.assert_array_equal(r["meta"], g["meta"][s])
        assert int(np.argmax(r["meta"])) == int(g["best_expert"][s])
        n = len(alleles[s])
        keys = [(alleles[s][i], alleles[s][j]) for i in range(n) for j in range(i, n)]
        assert len(r["expertPredictions"]) == 3
        for e, d in enumerate(r["expertPredictions"]):
            assert list(d.keys()) == keys                                    # itertools.product order, i <= j
            for q, k in enumerate(keys):
                assert torch.is_tensor(d[k]) and d[k].dim() == 0 and d[k].dtype == torch.float32
                assert float(d[k]) == float(g["experts"][e, pair_off[s] + q])
        # the caller's rule on these records gives the golden calls (sorted((v, k)) picks max value, then greatest key)
        for e in range(3):
            top = sorted([(float(v), k) for k, v in r["expertPredictions"][e].items()], reverse=True)[0][1]
            assert [alleles[s].index(top[0]), alleles[s].index(top[1])] == list(g["call_pair"][s, e])
    with pytest.raises(ValueError):
        model.feature_records(g["experts"], g["meta"], pair_off, alleles[:-1], loci)
    with pytest.raises(ValueError):
        model.feature_records(g["experts"], g["meta"], pair_off, [a[:1] for a in alleles], loci)


def test_gpu_numa_binding_helpers(tmp_path, monkeypatch):
    """shard.bind_rank_to_gpu_numa: cpulist parsing, sysfs lookup by PCI address, and the no-op fallbacks."""
    import types
    from hello_b200 import shard
    assert shard.parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert shard.parse_cpulist("") == [] and shard.parse_cpulist("5") == [5]
    dev = tmp_path / "0000:1b:00.0"
    dev.mkdir()
    mine = sorted(os.sched_getaffinity(0))
    (dev / "local_cpulist").write_text("%d\n" % mine[0])
    props = types.SimpleNamespace(pci_domain_id=0, pci_bus_id=0x1b, pci_device_id=0)
    monkeypatch.setattr(torch.cuda, "get_device_properties", lambda i: props)
    assert shard.gpu_local_cpus(0, sysfs=str(tmp_path)) == [mine[0]]
    assert shard.gpu_local_cpus(0, sysfs=str(tmp_path / "missing")) == []
    real = shard.gpu_local_cpus
    try:
        monkeypatch.setattr(shard, "gpu_local_cpus", lambda i: real(i, sysfs=str(tmp_path)))
        assert shard.bind_rank_to_gpu_numa(0) == ([mine[0]] if len(mine) > 1 else mine)
        monkeypatch.setattr(shard, "gpu_local_cpus", lambda i: [10 ** 6])          # no overlap with the allowed set
        before = sorted(os.sched_getaffinity(0))
        assert shard.bind_rank_to_gpu_numa(0) == before
    finally:
        os.sched_setaffinity(0, mine)


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the driver's reference arm, CPU only): stdout is exactly one JSON line with the contract's
    keys; anything else a library writes to file descriptor 1 goes to stderr."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sites", "32", "--cpu-workers", "2"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    lines = out.stdout.splitlines()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "candidate_sites_per_sec" and d["unit"] == "sites/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["value"] > 0
    from oracle import ref_model
    # the reference's own modules wherever oracle/_ref/python exists (make -C oracle; it travels to the GPU box),
    # the labelled port otherwise
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_model.available() else "port")
    assert d["cpu_baseline"]["cores"] == 2
    assert d["e2e"] == {"value": d["value"], "unit": "sites/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("illumina_30x")
    port = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                           "--cpu-sites", "32", "--cpu-workers", "2", "--cpu-port"], capture_output=True, text=True, timeout=300)
    assert port.returncode == 0 and json.loads(port.stdout.splitlines()[-1])["cpu_baseline"]["kind"] == "port"


def test_streamed_ranges_cover_the_batch_and_start_small():
    """forward_host / forward_host_packed stream the batch in site ranges: contiguous, complete, never larger than the
    requested size, and the first one an eighth of it (its copy is the one transfer no computation hides)."""
    from hello_b200.model import MoEEngine
    for n_sites, chunk in ((1_000_000, 65536), (5000, 65536), (70_000, 8192), (1, 64), (8192, 8192)):
        r = list(MoEEngine._ranges(n_sites, chunk))
        assert r[0][0] == 0 and r[-1][1] == n_sites
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert all(0 < s1 - s0 <= max(chunk, 1024) for s0, s1 in r)
        assert r[0][1] - r[0][0] == min(n_sites, max(1024, chunk // 8))
    assert list(MoEEngine._ranges(0, 4096)) == []
