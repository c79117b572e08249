#!/usr/bin/env python
"""Benchmark of the GPU read-feature encoder (include/hello_encode.h), the step before the network.

    python tools/bench_encoder.py [--sites N] [--steps K] [--warmup W]

One JSON line, same conventions as bench.py: `value` = read rows encoded per second with the packed reads resident in
HBM; `e2e` = the same with the packed reads coming from pinned host memory every step (H2D inside the timed region,
an 8-byte checksum read back); `roofline` = achieved algorithmic bytes/s of encode_reads_kernel against the measured
HBM copy bandwidth; `cpu_baseline` = the reference's own compiled C++ encoder (oracle/_ref/libref_encoder.so, kind
"reference") when it was built, else the oracle port (oracle/encoder_oracle.py, Python loops restating the C++), on a
bounded sample.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _cpu_rows(args):
    seed, n, use_ref = args
    import numpy as np
    from oracle import encoder_oracle as E, ref_encoder as R
    fn = R.compute_features_colored_simple if use_ref else E.compute_features_colored_simple
    rng = np.random.default_rng(seed)
    sites = [E.random_site(rng, n_reads=30, border_cases=False) for _ in range(20)]      # made outside the timed region
    done, k = 0, 0
    t0 = time.perf_counter()
    while done < n:
        site = sites[k % len(sites)]
        k += 1
        for allele in site.supports:
            done += fn(site, allele, 150, False, False).shape[0]
    return done, time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=int, default=200_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    import numpy as np
    import torch
    from hello_b200 import encoder, synth
    dev = torch.device("cuda:0")
    packed, row_read, row_site = synth.make_packed_reads(args.sites, 30, seed=13)
    R = int(row_read.size)
    dp = encoder.DevicePackedReads(packed, dev)
    out = torch.empty((R, 150, 6), dtype=torch.uint8, device=dev)
    rr_d, rs_d = torch.from_numpy(row_read).to(dev), torch.from_numpy(row_site).to(dev)   # resident, like the reads
    for _ in range(args.warmup):
        dp.encode(rr_d, rs_d, 6, out=out)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(args.steps):
        dp.encode(rr_d, rs_d, 6, out=out)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / args.steps
    in_bytes = packed.nbytes() + row_read.nbytes + row_site.nbytes
    out_bytes = R * 900
    # e2e: pinned host -> device every step
    pinned = encoder.PackedReads(*[torch.from_numpy(getattr(packed, f)).pin_memory().numpy() if f != "read_base"
                                   else getattr(packed, f) for f in packed.__dataclass_fields__])
    chk = torch.empty(1, dtype=torch.int64).pin_memory()
    def e2e_step():
        d = encoder.DevicePackedReads(pinned, dev)
        d.encode(row_read, row_site, 6, out=out)
        chk.copy_(out[:: max(R // 4096, 1)].sum(dtype=torch.int64).reshape(1), non_blocking=True)
    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.steps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs") or 6500.0
    achieved = (in_bytes + out_bytes) / (ms / 1e3) / 1e9
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import ref_encoder
        use_ref = ref_encoder.available()
        workers = len(os.sched_getaffinity(0))
        with mp.get_context("fork").Pool(workers) as pool:
            res = pool.map(_cpu_rows, [(100 + w, 20000 if use_ref else 600, use_ref) for w in range(workers)])
        rows = sum(r[0] for r in res)
        what = ("the reference's compiled computeFeaturesColoredSimple, one AlleleSearcherLiteFiltered built per call as "
                "the caller does per site" if use_ref else "pure-Python restatement of the C++ loop")
        cpu = {"value": rows / max(r[1] for r in res), "unit": "rows/s", "cores": workers, "kind": "reference" if use_ref else "port",
               "sample": "%d rows of random 30-read sites, %d worker processes (%s)" % (rows, workers, what)}
    print(json.dumps({
        "metric": "encoded_read_rows_per_sec", "value": R / (ms / 1e3), "unit": "rows/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "illumina_30x_%dk_sites_encoder" % (args.sites // 1000), "rows": R, "feature_length": 150,
                   "channels": 6, "l2": "input %.2f GB + output %.2f GB per step, far larger than L2" % (in_bytes / 1e9, out_bytes / 1e9)},
        "gpu_launches": args.steps,
        "e2e": {"value": R / e2e_s, "unit": "rows/s", "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": 8},
        "roofline": {"kernel": "encode_reads_kernel", "bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                     "frac": achieved / hbm, "traffic": None, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback",
                     "algorithmic_bytes_per_row": (in_bytes + out_bytes) / R},
        "cpu_baseline": cpu}), flush=True)


if __name__ == "__main__":
    main()
