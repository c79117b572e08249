#!/usr/bin/env python
"""Benchmark of the HELLO MoE forward on B200: candidate sites/sec through the batched forward (read convolver ->
allele/site heads -> genotype posteriors + argmax), BASELINE.json config 2 ("Illumina 30x model on 1M synthetic
sites, 1 B200") by default.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference's own modules on the host CPU cores

One JSON line on stdout (rank 0).  `value` = sites/s with inputs resident in HBM; `e2e` = the same job through
MoEEngine.forward_host with pinned HOST buffers (H2D of the pileups and D2H of the per-site results inside the
timed region); `roofline` = the read-convolver stage (dominant kernel) timed with CUDA events by the library;
`cpu_baseline` = the reference's own per-site call (oracle/_ref/python) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "candidate_sites_per_sec"
WORKLOADS = {
    # name: (model config, coverage, description)
    "illumina_30x": ("single_tech", 30, "Illumina 30x single-technology model"),
    "pacbio_hp_30x": ("single_tech_hp", 30, "PacBio 30x model with haplotag channel"),
    "hybrid_no_ensemble_30x": ("hybrid_no_ensemble", 30, "hybrid Illumina+PacBio no_ensemble model"),
    "hybrid_ensemble2_30x": ("hybrid_ensemble2", 30, "hybrid 2-expert gated model"),
    "wgs_ragged_15_60x": ("single_tech", (15, 60), "ragged coverage 15x-60x sweep"),
    "hybrid_no_ensemble_wide_30x": ("hybrid_no_ensemble_wide", 30, "hybrid no_ensemble model with 2x channels (layer-wise tensor-core kernels)"),
    "hybrid_no_ensemble_addendum_30x": ("hybrid_no_ensemble_addendum", 30, "hybrid transfer-learning model"),
    "illumina_30x_addendum": ("single_tech_addendum", 30, "Illumina 30x transfer-learning model (two more residual blocks per sub-network)"),
    "illumina_30x_softplus": ("single_tech_softplus", 30, "Illumina 30x model of the Softplus configuration (fused kernels with a Softplus epilogue)"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="illumina_30x", choices=sorted(WORKLOADS))
    ap.add_argument("--sites", type=int, default=1_000_000, help="sites per GPU (weak scaling)")
    ap.add_argument("--precision", default=os.environ.get("HELLO_PRECISION", "bf16x3"),
                    choices=["bf16x3", "bf16", "fp32"],
                    help="bf16x3 (default): read convolver on tcgen05 with hi+lo bf16 operands, fp32 accumulate, "
                         "posteriors within 1e-3 of the reference; bf16: single-MMA fast mode; fp32: CUDA cores only")
    ap.add_argument("--no-gather", action="store_true", help="N>1: skip the NCCL gather of per-site results")
    ap.add_argument("--partition", default="replicate-shape", choices=["replicate-shape", "balanced"],
                    help="replicate-shape (default): every rank generates --sites sites of its own (weak scaling); "
                         "balanced: ONE deterministic dataset of --total-sites sites (default --sites x ranks) is cut into "
                         "contiguous site ranges of equal algorithmic cost (shard.site_costs / balanced_ranges), each rank "
                         "generates and scores only its own range")
    ap.add_argument("--total-sites", type=int, default=0, help="--partition balanced: sites of the whole dataset")
    ap.add_argument("--no-other-workloads", action="store_true",
                    help="skip the other_workloads legs (BASELINE configs 3-5 at --other-sites sites, 1 GPU only)")
    ap.add_argument("--other-sites", type=int, default=131072)
    ap.add_argument("--workspace-gb", type=float, default=6.0)
    ap.add_argument("--chunk-sites", type=int, default=0, help="cap on sites per internal chunk (0 = auto)")
    ap.add_argument("--e2e-chunk-sites", type=int, default=0, help="sites per streamed range (0 = min(65536, sites/8))")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e-packed", action="store_true",
                    help="skip the second end-to-end figure (aligned reads on the host, rows encoded on the GPU)")
    ap.add_argument("--e2e-packed-sites", type=int, default=262144, help="sites per GPU of the e2e_packed leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sites", type=int, default=0, help="sites per CPU step (0 = 128 per worker)")
    ap.add_argument("--cpu-workers", type=int, default=0)
    ap.add_argument("--cpu-port", action="store_true", help="CPU arm: time the oracle port instead of the reference modules")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own per-site call on the host cores
#   kind "reference": oracle/_ref/python (the reference's unmodified MixtureOfExpertsAdvanced / NNTools / architectures,
#                     copied there by oracle/Makefile) -- MoEMergedWrapperAdvanced(featureDict, segment), one site per call,
#                     under no_grad, one torch thread per worker process: python/caller_calling.py:39,651-652, call.py:26,30
#   kind "port":      labelled fallback when oracle/_ref/python is missing -- oracle/hello_oracle.py in 16-site calls
_W = {}


def _ref_worker_init(cfg_name):
    import torch
    torch.set_num_threads(1)                      # python/caller_calling.py:39
    from hello_b200 import arch, weights
    from oracle import ref_model
    cfg = arch.CONFIGS[cfg_name]
    _W["net"] = ref_model.build_wrapper(cfg_name, weights.init_params(cfg, seed=13), provide_predictions=True)


def _ref_worker_run(task):
    """scoreSite (python/caller_calling.py:612-654): uint8 arrays -> torch.Tensor floats -> network(featureDict, segment)."""
    import torch
    net = _W["net"]
    done = 0
    for alleles, seg in task:
        fd = {name: (torch.Tensor(t0), torch.Tensor(t1) if t1 is not None else None) for name, t0, t1 in alleles}
        with torch.no_grad():
            out = net(fd, seg)
        best = sorted([(v, k) for k, v in out[0].items()], reverse=True)[0]        # caller_calling.py:702-705
        done += best is not None
    return done


def _ref_tasks(pl, cfg, sites_per_task):
    tasks, cur = [], []
    sao = pl.site_allele_off
    for s in range(pl.n_sites):
        a0, a1 = int(sao[s]), int(sao[s + 1])
        alleles = []
        for k, a in enumerate(range(a0, a1)):
            parts = []
            for t in range(len(cfg.read_cin)):
                r0, r1 = int(pl.allele_read_off[t][a]), int(pl.allele_read_off[t][a + 1])
                parts.append(pl.reads[t][r0:r1].numpy())
            alleles.append(("ACGT"[k % 4] * (1 + k // 4), parts[0], parts[1] if len(parts) > 1 else None))
        cur.append((alleles, pl.ref_onehot[s:s + 1].clone()))
        if len(cur) == sites_per_task:
            tasks.append(cur)
            cur = []
    if cur:
        tasks.append(cur)
    return tasks


def _cpu_worker_init(cfg_name):
    import torch
    torch.set_num_threads(1)                      # the reference runs 1 torch thread per worker process
    from hello_b200 import arch, weights          # (python/caller_calling.py:39, python/call.py:26,30)
    from oracle import hello_oracle as O
    cfg = arch.CONFIGS[cfg_name]
    _W["model"] = O.OracleModel(cfg, weights.init_params(cfg, seed=13))
    _W["cfg"] = cfg
    _W["O"] = O


def _cpu_worker_run(task):
    import torch
    tensors, naps, nrpa, ref = task
    model, cfg, O = _W["model"], _W["cfg"], _W["O"]
    res = model.forward(tensors, naps, nrpa, ref)
    post = O.batched_posteriors(cfg, res, naps)
    return len(post)


def _cpu_tasks(pl, cfg, batch_sites):
    import torch
    tasks = []
    sao = pl.site_allele_off
    for s0 in range(0, pl.n_sites, batch_sites):
        s1 = min(pl.n_sites, s0 + batch_sites)
        a0, a1 = int(sao[s0]), int(sao[s1])
        tensors, nrpa = [], []
        for t in range(len(cfg.read_cin)):
            aro = pl.allele_read_off[t]
            r0, r1 = int(aro[a0]), int(aro[a1])
            tensors.append(pl.reads[t][r0:r1].transpose(1, 2).contiguous())
            nrpa.append(torch.diff(aro[a0:a1 + 1]).tolist())
        if len(tensors) == 1:
            tensors.append(None)
            nrpa.append(None)
        tasks.append((tuple(tensors), torch.diff(sao[s0:s1 + 1]).tolist(), tuple(nrpa), pl.ref_onehot[s0:s1]))
    return tasks


def run_reference(args):
    """Times the reference's implementation of the path on the host cores: its own modules, its own per-site call and its
    own threading model (a pool of single-threaded worker processes, python/call.py:111,215-221), on `cpu_sites` sites per
    step.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from hello_b200 import arch, synth
    from oracle import ref_model
    cfg_name, cov, desc = WORKLOADS[args.workload]
    cfg = arch.CONFIGS[cfg_name]
    workers = args.cpu_workers or len(os.sched_getaffinity(0))
    kind = "reference" if (ref_model.available() and not args.cpu_port) else "port"
    # a bounded sample: ~60 sites/s/core through the stock per-site call, so 64 sites per worker and step is ~1 s of work
    n_sites = args.cpu_sites or (64 if kind == "reference" else 128) * workers
    pl = synth.make_pileups(n_sites, coverage=cov, channels=cfg.read_cin, seed=13)
    if kind == "reference":
        tasks, init, run = _ref_tasks(pl, cfg, 4), _ref_worker_init, _ref_worker_run
        how = "the reference's own modules (oracle/_ref/python, unmodified): MoEMergedWrapperAdvanced(featureDict, segment) " \
              "once per site under no_grad + the caller's argmax, as python/caller_calling.py:651-652,702-705"
    else:
        tasks, init, run = _cpu_tasks(pl, cfg, 16), _cpu_worker_init, _cpu_worker_run
        how = "oracle/hello_oracle.py (a port of the reference forward: oracle/_ref/python is missing), 16-site calls"
    ctx = mp.get_context("fork")
    with ctx.Pool(workers, initializer=init, initargs=(cfg_name,)) as pool:
        def step():
            return sum(pool.map(run, tasks, chunksize=1))
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        done = 0
        for _ in range(args.steps):
            done += step()
        dt = time.perf_counter() - t0
    value = done / dt
    sample = "%d synthetic %s sites per step (seed 13, CPU generator); %s; %d worker processes x 1 torch thread" % (
        n_sites, args.workload, how, workers)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "sites/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "wiring": cfg_name, "sites_per_step": n_sites, "coverage": cov},
        "cpu_baseline": {"value": value, "unit": "sites/s", "cores": workers, "kind": kind, "sample": sample,
                         "per_core": value / max(workers, 1)},
        "e2e": {"value": value, "unit": "sites/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return line


# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, "/tmp/hello_clocks_%d_%d.csv" % (os.getpid(), index)

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons),
                       power_w_max=max(power), samples=len(sm))
        return out


def generate_on_device(cfg, cov, n_sites, device, seed, gen_chunk=16384):
    """Synthetic pileups generated on the GPU in slices, concatenated into one resident batch."""
    import torch
    from hello_b200 import synth
    reads = [[] for _ in cfg.read_cin]
    rpa = [[] for _ in cfg.read_cin]
    apS, refs = [], []
    for i, s0 in enumerate(range(0, n_sites, gen_chunk)):
        n = min(gen_chunk, n_sites - s0)
        pl = synth.make_pileups(n, coverage=cov, channels=cfg.read_cin, seed=seed + 7919 * i, device=device)
        for t in range(len(cfg.read_cin)):
            reads[t].append(pl.reads[t])
            rpa[t].append(torch.diff(pl.allele_read_off[t]))
        apS.append(torch.diff(pl.site_allele_off))
        if cfg.meta == "meta_convolver_ref":
            refs.append(pl.ref_onehot)
    def csr(parts):
        c = torch.cat(parts).to(torch.int64)
        off = torch.zeros(c.numel() + 1, dtype=torch.int64)
        off[1:] = torch.cumsum(c, 0)
        return off.to(torch.int32)
    reads = tuple(torch.cat(r) for r in reads)
    return reads, tuple(csr(r) for r in rpa), csr(apS), (torch.cat(refs) if refs else None)


def generate_balanced_shard(cfg, cov, total_sites, world, rank, device, seed, block=16384):
    """--partition balanced: the dataset is `total_sites` sites in blocks of `block` (block b = synth.make_pileups with seed
    + 7919 b, so it does not depend on the number of ranks).  Every rank replays the cheap head of each block's random
    stream (synth.site_counts) to get alleles and reads per site, prices the sites with shard.site_costs, cuts the prefix
    sum with shard.balanced_ranges and generates only the blocks that overlap its own range (shard.take_shard trims the
    first and last one).  Mirrors the reference's shardHotspots.py:78-137 (contiguous hotspot shards, python/call.py:171-221)."""
    import numpy as np
    import torch
    from hello_b200 import shard, synth
    n_blocks = (total_sites + block - 1) // block
    na, nr = [], []
    for b in range(n_blocks):
        n = min(block, total_sites - b * block)
        a, r = synth.site_counts(n, cov, cfg.read_cin, seed + 7919 * b, device)
        na.append(a.cpu())
        nr.append(r.cpu())
    na, nr = torch.cat(na).numpy(), torch.cat(nr).numpy()
    sao_all = np.concatenate(([0], np.cumsum(na)))
    cost = shard.site_costs(cfg, sao_all, reads_per_site=[nr])
    ranges = shard.balanced_ranges(cost, world)
    s0, s1 = ranges[rank]
    reads, rpa, apS = [], [], []
    for b in range(s0 // block, (max(s1, s0 + 1) - 1) // block + 1):
        if s1 <= s0:
            break
        n = min(block, total_sites - b * block)
        pl = synth.make_pileups(n, coverage=cov, channels=cfg.read_cin, seed=seed + 7919 * b, device=device)
        lo, hi = max(s0 - b * block, 0), min(s1 - b * block, n)
        sh = shard.take_shard(pl.site_allele_off, pl.allele_read_off, lo, hi)
        r0, r1 = sh.read_range[0]
        reads.append(pl.reads[0][r0:r1].clone() if (lo, hi) != (0, n) else pl.reads[0])
        rpa.append(torch.diff(sh.allele_read_off[0]))
        apS.append(torch.diff(sh.site_allele_off))
        del pl

    def csr(parts):
        c = torch.cat(parts).to(torch.int64) if parts else torch.zeros(0, dtype=torch.int64)
        off = torch.zeros(c.numel() + 1, dtype=torch.int64)
        off[1:] = torch.cumsum(c, 0)
        return off.to(torch.int32)
    reads_t = torch.cat(reads) if reads else torch.zeros((0, 150, cfg.read_cin[0]), dtype=torch.uint8, device=device)
    total_cost = float(cost.sum())
    info = {"mode": "balanced", "total_sites": int(total_sites), "block_sites": block,
            "site_ranges": [[int(a), int(b)] for a, b in ranges],
            "cost_share": [float(cost[a:b].sum() / total_cost) for a, b in ranges],
            "reads_per_rank": [int(nr[a:b].sum()) for a, b in ranges]}
    return (reads_t,), (csr(rpa),), csr(apS), None, info


def _oracle_logit_check(cfg, params, reads, aro, sao, ref, logits, n_check=32):
    """Checker only (never timed): max |dlogit| of the first `n_check` sites against the CPU oracle."""
    import torch
    from oracle import hello_oracle as O
    n = min(n_check, sao.numel() - 1)
    a1 = int(sao[n])
    tensors, nrpa = [], []
    for t in range(len(cfg.read_cin)):
        r1 = int(aro[t][a1])
        tensors.append(reads[t][:r1].cpu().transpose(1, 2).contiguous())
        nrpa.append(torch.diff(aro[t][:a1 + 1]).tolist())
    if len(tensors) == 1:
        tensors.append(None)
        nrpa.append(None)
    torch.set_num_threads(max(1, min(16, len(os.sched_getaffinity(0)))))
    res = O.OracleModel(cfg, params).forward(tuple(tensors), torch.diff(sao[:n + 1]).tolist(), tuple(nrpa),
                                             ref[:n].cpu() if ref is not None else None)
    got = logits[:, :a1].cpu()
    if cfg.returns_meta:
        want = torch.stack([e.reshape(-1) for e in res[0]])
        present = [e for e in range(3) if cfg.xattn_present[e]]
        return float((got[present] - want[present]).abs().max()), n
    head = 0 if cfg.xattn_present[0] else 2
    return float((got[head] - res.reshape(-1)).abs().max()), n


def run_other_workloads(args, dev, skip):
    """BASELINE.json configs 3-5 next to the headline (config 2): each at --other-sites sites on this GPU, resident inputs,
    2 timed steps after 1 warm-up; sites/s, share of the step spent in the read-convolver stage and max |dlogit| against the
    CPU oracle on a sample.  Does not touch the headline fields."""
    import torch
    from hello_b200 import _lib, arch, model, weights
    out = {}
    # after BASELINE's configs: the two reference configurations outside them that differ in kernels --
    # Softplus epilogues on the fused kernels, and the 2x-wide model on the layer-wise kernel (1/8 of the sites: 16x slower)
    for name in ("pacbio_hp_30x", "hybrid_no_ensemble_30x", "hybrid_ensemble2_30x", "wgs_ragged_15_60x",
                 "illumina_30x_softplus", "hybrid_no_ensemble_wide_30x"):
        if name == skip:
            continue
        n_sites = max(1, args.other_sites // 8) if name == "hybrid_no_ensemble_wide_30x" else args.other_sites
        cfg_name, cov, desc = WORKLOADS[name]
        cfg = arch.CONFIGS[cfg_name]
        try:
            params = weights.init_params(cfg, seed=13)
            eng = model.MoEEngine(cfg, params, device=dev, precision=args.precision,
                                  workspace_bytes=int(args.workspace_gb * (1 << 30)))
            reads, aro, sao, ref = generate_on_device(cfg, cov, n_sites, dev, seed=4242)
            batch = model.DeviceBatch.from_host(reads, _lib.LAYOUT_RLC, aro, sao, ref, dev)
            res = eng.alloc_result(batch)
            eng.run(batch, res)
            torch.cuda.synchronize(dev)
            eng.profile_enable(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                eng.run(batch, res)
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1)
            rc_ms, _ = eng.profile_collect()
            eng.profile_enable(False)
            f_read, f_allele, f_site = arch.flops_model(cfg)
            R = [int(r.shape[0]) for r in reads]
            flops = sum(R[t] * f_read[t] for t in range(len(R))) + batch.n_alleles * f_allele + batch.n_sites * f_site
            err, n_checked = _oracle_logit_check(cfg, params, reads, aro, sao, ref, res.logits)
            out[name] = {"description": desc, "wiring": cfg_name, "sites": batch.n_sites, "reads": R,
                         "sites_per_sec": 2 * batch.n_sites / (ms / 1e3), "ms_per_step": ms / 2,
                         "read_convolver_stage_share": rc_ms / ms if ms > 0 else None,
                         "algorithmic_tflops": 2 * flops / (ms / 1e3) / 1e12,
                         "max_abs_dlogit_vs_oracle": err, "oracle_sample_sites": n_checked}
            eng.close()
            del eng, batch, res, reads
            torch.cuda.empty_cache()
        except Exception as exc:                   # report, never hide
            out[name] = {"error": repr(exc)[:300]}
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from hello_b200 import _lib, arch, model, shard, weights

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this repo has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus = None
    if world > 1:
        # one process per GPU: stay on the GPU's socket so the pinned buffers of the e2e leg are allocated next to it
        numa_cpus = shard.bind_rank_to_gpu_numa(local)
        dist.init_process_group("nccl", device_id=dev)

    cfg_name, cov, desc = WORKLOADS[args.workload]
    cfg = arch.CONFIGS[cfg_name]
    params = weights.init_params(cfg, seed=13)
    engine = model.MoEEngine(cfg, params, device=dev, precision=args.precision,
                             workspace_bytes=int(args.workspace_gb * (1 << 30)), max_chunk_sites=args.chunk_sites)

    partition_info = None
    if args.partition == "balanced":
        total_sites = args.total_sites or args.sites * world
        reads, aro, sao, ref, partition_info = generate_balanced_shard(cfg, cov, total_sites, world, rank, dev, seed=13)
    else:
        reads, aro, sao, ref = generate_on_device(cfg, cov, args.sites, dev, seed=13 + 1000 * rank)
    batch = model.DeviceBatch.from_host(reads, _lib.LAYOUT_RLC, aro, sao, ref, dev)
    S, A = batch.n_sites, batch.n_alleles
    gather = world > 1 and not args.no_gather
    # N>1: the per-site results live in one packed buffer per rank (two slots), so the gather is ONE all_gather per step, queued
    # on a side stream where it overlaps the next step's forward (shard.SiteGatherer)
    gatherer = shard.SiteGatherer(dev).plan(S, A, batch.n_pairs) if gather else None
    results = [engine.result_from_views(batch, gatherer.result_views(k)) for k in range(2)] if gather \
        else [engine.alloc_result(batch)]
    result = results[0]
    R = [int(r.shape[0]) for r in reads]
    input_gb = sum(r.numel() for r in reads) / 1e9
    f_read, f_allele, f_site = arch.flops_model(cfg)
    flops_read = sum(R[t] * f_read[t] for t in range(len(R)))
    flops_step = flops_read + A * f_allele + S * f_site

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    step_no = [0]
    fwd_events = []

    def step(timed=False):
        """One pass of the hot path over this rank's shard; with N>1 the per-site results are then gathered over
        NCCL (the only collective of the path -- nothing is exchanged inside the forward)."""
        slot = step_no[0] % len(results)
        step_no[0] += 1
        if gather:
            gatherer.before_overwrite(slot)          # the gather of two steps ago has finished reading this slot
        if timed:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        engine.run(batch, results[slot])
        if timed:
            b.record()
            fwd_events.append((a, b))
        if gather:
            gatherer.gather_async(slot)

    # ---- value: inputs resident in HBM ---------------------------------------------------------------------------
    for _ in range(args.warmup):
        step()
    barrier()
    engine.profile_enable(True)
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = engine.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step(timed=True)
    if gather:
        gatherer.wait()                              # every queued gather is inside the timed region
    ev1.record()
    barrier()
    clocks = sampler.stop()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    launches = engine.launch_count() - launches0
    rc_ms, rc_regions = engine.profile_collect()
    engine.profile_enable(False)
    sites_all = S
    if world > 1:
        t = torch.tensor([S], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        sites_all = int(t.item())
    value = (sites_all if args.partition == "balanced" else world * S) * args.steps / (ms_total / 1e3)
    # per-rank forward time (load balance of the partition): mean over the timed steps, gathered to every rank
    my_fwd_ms = sum(a.elapsed_time(b) for a, b in fwd_events) / max(len(fwd_events), 1)
    rank_ms = [my_fwd_ms]
    if world > 1:
        t = torch.zeros(world, dtype=torch.float64, device=dev)
        t[rank] = my_fwd_ms
        dist.all_reduce(t)
        rank_ms = [float(x) for x in t.tolist()]
    if partition_info is not None:
        partition_info.update(rank_forward_ms=rank_ms, imbalance_max_over_mean=max(rank_ms) / (sum(rank_ms) / len(rank_ms)))

    # ---- e2e: host buffers, copies inside the timed region -------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        # The host copy of the pileups is pinned; keep it inside half of this rank's share of the free host memory (all
        # ranks of a node pin at once) by timing the e2e leg on a prefix of the sites if need be.
        try:
            avail = next(int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable"))
        except Exception:
            avail = 64 << 30
        if args.e2e_chunk_sites <= 0:
            args.e2e_chunk_sites = max(8192, min(65536, S // 8))
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
        in_bytes = sum(r.numel() for r in reads) + (ref.numel() * 4 if ref is not None else 0)
        budget = 0.5 * avail / max(local_world, 1)
        S_e = S if in_bytes <= budget else max(args.e2e_chunk_sites, int(S * budget / in_bytes))
        S_e = min(S, S_e)
        a_e = int(sao[S_e])

        def pinned_copy(t, n):                                # device -> pinned host, no pageable intermediate
            h = torch.empty((n,) + tuple(t.shape[1:]), dtype=t.dtype, pin_memory=True)
            h.copy_(t[:n])
            return h
        hb = model.HostBatch([pinned_copy(r, int(aro[t][a_e])) for t, r in enumerate(reads)], _lib.LAYOUT_RLC,
                             [o[:a_e + 1].clone() for o in aro], sao[:S_e + 1].clone(),
                             pinned_copy(ref, S_e) if ref is not None else None, pin=True)
        del batch, reads, results, result
        gatherer = None
        torch.cuda.empty_cache()
        engine.forward_host(hb, args.e2e_chunk_sites)                       # warm-up (allocations, pinned outputs)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out = engine.forward_host(hb, args.e2e_chunk_sites)
            torch.cuda.synchronize(dev)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        S_e_all = S_e
        if world > 1 and args.partition == "balanced":
            t = torch.tensor([S_e], dtype=torch.int64, device=dev)
            dist.all_reduce(t)
            S_e_all = int(t.item())
        e2e = {"value": (S_e_all if args.partition == "balanced" else world * S_e) * args.steps / dt, "unit": "sites/s", "h2d_bytes_per_step": hb.input_nbytes(),
               "d2h_bytes_per_step": out.nbytes(), "sites_per_gpu": S_e,
               "api": "MoEEngine.forward_host (pinned host buffers, read rows streamed in ranges of up to %d sites -- the "
                      "first ones an eighth, a quarter, half of that -- on a copy stream while the previous range "
                      "computes)" % args.e2e_chunk_sites}
        if numa_cpus is not None:
            e2e["host_binding"] = "rank 0 runs on %d cores local to its GPU (pinned buffers first-touched there)" % len(numa_cpus)

    # ---- e2e_packed: the host holds ALIGNED READS (what exists right after read sampling); the GPU encodes the rows --------
    e2e_packed = None
    if not args.no_e2e and not args.no_e2e_packed and len(cfg.read_cin) == 1 and not isinstance(cov, tuple):
        import numpy as np
        from hello_b200 import synth as _synth
        batch = reads = results = result = hb = out = None
        torch.cuda.empty_cache()
        # an extra figure must never cost the headline line: set-up failures are reported in the record, and the ranks agree
        # on whether to enter the timed region (a rank waiting alone in a barrier would hang the job)
        prep_err, hpb = None, None
        try:
            gen_sites = 32768
            times = max(1, min(args.e2e_packed_sites, S) // gen_sites)
            packed, rr, rs = _synth.make_packed_reads(gen_sites, coverage=cov, seed=13 + rank, hp=cfg.read_cin[0] == 7)
            aro_p, sao_p = _synth.packed_allele_csr(np.diff(packed.read_base), seed=13 + rank)
            packed, rr, rs = _synth.tile_packed_reads(packed, rr, rs, times)
            rep = lambda off: torch.cat([off[:-1].long() + k * int(off[-1]) for k in range(times)] +
                                        [torch.tensor([times * int(off[-1])])]).to(torch.int32)
            hpb = model.HostPackedBatch(packed, (rr,), (rs,), (rep(aro_p),), rep(sao_p), pin=True)
            S_p = hpb.n_sites
            chunk_p = max(8192, min(65536, S_p // 8))
            engine.forward_host_packed(hpb, chunk_p)                       # warm-up (allocations, pinned outputs)
            torch.cuda.synchronize(dev)
        except Exception as exc:
            prep_err = repr(exc)[:300]
        all_ok = prep_err is None
        if world > 1:
            t = torch.tensor([1 if all_ok else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            all_ok = bool(t.item())
        if not all_ok:
            e2e_packed = {"error": prep_err or "set-up failed on another rank"}
        else:
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                out_p = engine.forward_host_packed(hpb, chunk_p)
                torch.cuda.synchronize(dev)
            barrier()
            dt = max_over_ranks(time.perf_counter() - t0)
            e2e_packed = {"value": world * S_p * args.steps / dt, "unit": "sites/s", "h2d_bytes_per_step": hpb.input_nbytes(),
                          "d2h_bytes_per_step": out_p.nbytes(), "sites_per_gpu": S_p, "rows_per_gpu": int(rr.size),
                          "h2d_bytes_per_row": hpb.input_nbytes() / max(int(rr.size), 1),
                          "api": "MoEEngine.forward_host_packed (pinned host buffers of aligned reads -- bases, qualities, CIGARs, "
                                 "reference windows -- streamed in ranges of up to %d sites; hello_encode_reads builds the [R,150,C] rows "
                                 "on the GPU; then hello_moe_forward_range)" % chunk_p,
                          "data": "%d generated sites (synth.make_packed_reads) tiled %d times" % (gen_sites, times)}
        hpb = out_p = None

    if rank != 0:
        print("bench.py: rank %d of %d done" % (rank, world), file=sys.stderr, flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tensor_peak = peaks.get("bf16_tflops_sustained") or 1400.0
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback (B200_PROFILING.md, sustained)"
    achieved = flops_read * args.steps / (rc_ms / 1e3) / 1e12 if rc_ms > 0 else None
    traffic, traffic_src = None, None
    try:   # dram bytes per read of the dominant kernel, from the committed `ncu --set full` capture
        prof = json.load(open(os.path.join(ROOT, "profiles", "readconv_tc_summary.json")))
        if args.precision in prof and rc_regions > 0:
            traffic = prof[args.precision]["dram_bytes_per_read"] * sum(R) * args.steps / rc_regions
            traffic_src = prof[args.precision]["source"]
    except Exception:
        pass
    roofline = {
        "kernel": "read convolver stage (%s)" % ("readconv_tc" if args.precision != "fp32" else
                                                "conv1d_fp32_kernel x19 + maxpool per chunk"),
        "bound": "tensor", "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s",
        "frac": (achieved / tensor_peak) if achieved else None, "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": peak_src,
        "algorithmic_flops_per_launch": flops_read * args.steps / rc_regions if rc_regions else None,
        "ms_per_launch": rc_ms / rc_regions if rc_regions else None,
        "algorithmic_flops_per_read": list(f_read), "reads_per_step": R,
        "stage_ms_per_step": rc_ms / max(args.steps, 1), "regions_per_step": rc_regions / max(args.steps, 1),
        "stage_share_of_step": rc_ms / ms_total if ms_total > 0 else None,
        "note": {"fp32": "fp32 mode runs the convolutions on CUDA cores (FFMA); the fraction is quoted against the bf16 "
                         "tensor-core peak the north star targets",
                 "bf16x3": "one fused tcgen05 kernel per chunk; bf16x3 computes 3 products per algorithmic MAC (hi*hi + hi*lo "
                           "+ lo*hi as two instructions, fp32 accumulate), so the ceiling of `frac` is 1/3 by construction",
                 "bf16": "one fused tcgen05 kernel per chunk; single bf16 MMA per MAC (fast mode, ~1e-1 logit error)"
                 }[args.precision],
    }
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1",
               "--workload", args.workload]
        try:
            out_ = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, RANK="0"))
            cpu_baseline = json.loads(out_.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as exc:  # report, never hide
            cpu_baseline = {"error": repr(exc)[:200]}
    other = None
    if world == 1 and not args.no_other_workloads:
        batch = reads = results = result = hb = out = None          # release the headline workload (HBM and pinned host memory)
        engine.close()
        torch.cuda.empty_cache()
        other = run_other_workloads(args, dev, skip=args.workload)
    scaling = "weak" if (args.partition != "balanced" or not args.total_sites) else "strong"
    line = {
        "metric": METRIC, "value": value, "unit": "sites/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / max(args.steps, 1), "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None,
        "dtype": {"fp32": "f32", "bf16x3": "bf16x3 (fp32 accumulate)", "bf16": "bf16 (fp32 accumulate)"}[args.precision],
        "data": "synthetic",
        "config": {"workload": "%s_%dk_sites_per_gpu" % (args.workload, args.sites // 1000), "wiring": cfg_name,
                   "weights": "random-init (seed 13), shipped blobs are git-lfs pointers", "sites_per_gpu": S,
                   "alleles_per_gpu": A, "reads_per_gpu": R, "coverage": cov, "precision": args.precision,
                   "partition": ("one dataset cut into contiguous site ranges of equal algorithmic cost, " if partition_info
                                 else "every rank scores its own sites, ") + "no collective inside the forward" +
                                (", ONE packed NCCL all_gather of the per-site results per step on a side stream "
                                 "(overlaps the next step's forward)" if gather else ""),
                   "l2": "inputs (%.1f GB per step) are far larger than L2; no flush needed" % input_gb,
                   "flops_per_step_per_gpu": flops_step},
        "clocks": clocks, "gpu_launches": launches, "e2e": e2e, "e2e_packed": e2e_packed, "roofline": roofline,
        "cpu_baseline": cpu_baseline, "rank_forward_ms": rank_ms,
    }
    if partition_info is not None:
        line["partition"] = partition_info
    if other is not None:
        line["other_workloads"] = other
    text = json.dumps(line)
    print(text, flush=True)
    print("bench.py: rank 0 wrote the result line (%d bytes)" % len(text), file=sys.stderr, flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    # stdout carries exactly ONE line, the JSON result: libraries that write to file descriptor 1 behind Python's back
    # (NCCL prints "NCCL version ..." there when the process group comes up) are sent to stderr instead.
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(result_fd, "w")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
