#!/usr/bin/env python
"""What the product actually calls: one site per `network(featureDict, ref_segment)` (python/caller_calling.py:651-652).

    python tools/bench_serving.py [--sites 400] [--producers 16 32] > profiles/r02_serving_latency_throughput.json

1. latency of the strict drop-in call (MoEMergedWrapperB200, one site per call, results read back on the host): p50 / p90 / p99
   in milliseconds, next to the reference's own wrapper on one host core when oracle/_ref/python is there;
2. throughput of the cross-process batching shim (hello_b200/serving.py): N forked producer processes -- the stand-ins for
   call.py's worker pool (python/call.py:111,215-221) -- each scoring its share of the sites through a RemoteNetwork against
   one ScoringServer that owns the GPU.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hello_b200 import arch, model, serving, synth, weights          # noqa: E402


def _producer(net, fds, barrier, out_q):
    net.providePredictions = True
    barrier.wait()
    t0 = time.perf_counter()
    lat = []
    for fd, seg in fds:
        a = time.perf_counter()
        r = net(fd, seg)
        lat.append(time.perf_counter() - a)
    out_q.put((len(fds), time.perf_counter() - t0, lat))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=int, default=400)
    ap.add_argument("--producers", type=int, nargs="*", default=[16, 32])
    ap.add_argument("--sites-per-producer", type=int, default=200)
    ap.add_argument("--config", default="single_tech")
    args = ap.parse_args()
    cfg = arch.CONFIGS[args.config]
    params = weights.init_params(cfg, seed=13)
    pl = synth.make_pileups(max(args.sites, args.sites_per_producer), coverage=30, channels=cfg.read_cin, seed=13)
    fds = [pl.site_feature_dict(s) for s in range(pl.n_sites)]
    out = {"config": args.config, "coverage": 30, "gpu": torch.cuda.get_device_name(0), "host_cores": len(os.sched_getaffinity(0))}

    # ---- 1. strict per-site call
    net = model.MoEMergedWrapperB200(model.MoEAttentionB200(cfg, params, device="cuda:0", precision="bf16x3")).eval()
    net.providePredictions = True
    for fd, seg in fds[:20]:
        net(fd, seg)
    lat = []
    for fd, seg in fds[:args.sites]:
        a = time.perf_counter()
        r = net(fd, seg)
        float(next(iter(r[0].values())))
        lat.append(time.perf_counter() - a)
    lat = np.array(lat) * 1e3
    out["per_site_call"] = {"api": "MoEMergedWrapperB200(featureDict, segment), providePredictions=True, one site per call",
                            "sites": len(lat), "p50_ms": float(np.percentile(lat, 50)), "p90_ms": float(np.percentile(lat, 90)),
                            "p99_ms": float(np.percentile(lat, 99)), "mean_ms": float(lat.mean()),
                            "sites_per_sec_one_caller": float(1e3 / lat.mean())}
    del net
    try:
        from oracle import ref_model
        if ref_model.available():
            torch.set_num_threads(1)
            ref = ref_model.build_wrapper(args.config, params)
            rl = []
            for fd, seg in fds[:min(60, args.sites)]:
                a = time.perf_counter()
                with torch.no_grad():
                    ref(fd, seg)
                rl.append(time.perf_counter() - a)
            rl = np.array(rl[5:]) * 1e3
            out["per_site_call"]["reference_cpu_one_core_p50_ms"] = float(np.percentile(rl, 50))
            out["per_site_call"]["reference_cpu_one_core_p99_ms"] = float(np.percentile(rl, 99))
    except Exception as exc:
        out["per_site_call"]["reference_error"] = repr(exc)[:200]

    # ---- 2. the batching shim under N producers
    out["scoring_server"] = []
    ctx = mp.get_context("fork")
    for n_prod in args.producers:
        with serving.ScoringServer(args.config, params, n_clients=n_prod, device="cuda:0", precision="bf16x3", max_sites=4096) as server:
            barrier, out_q = ctx.Barrier(n_prod), ctx.Queue()
            procs = []
            for c in range(n_prod):
                share = [fds[(c * 7 + k) % len(fds)] for k in range(args.sites_per_producer)]
                p = ctx.Process(target=_producer, args=(server.client(c), share, barrier, out_q))
                p.start()
                procs.append(p)
            res = [out_q.get(timeout=900) for _ in procs]
            for p in procs:
                p.join(timeout=60)
        n = sum(r[0] for r in res)
        wall = max(r[1] for r in res)
        lats = np.concatenate([np.array(r[2]) for r in res]) * 1e3
        out["scoring_server"].append({"producers": n_prod, "sites": n, "sites_per_sec": n / wall, "wall_s": wall,
                                      "request_latency_p50_ms": float(np.percentile(lats, 50)),
                                      "request_latency_p99_ms": float(np.percentile(lats, 99)),
                                      "note": "producers are single-threaded processes doing the reference's per-site packing "
                                              "(featureDict -> uint8 arrays -> queue); the server batches whatever is pending"})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
