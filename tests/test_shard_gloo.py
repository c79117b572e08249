"""Multi-GPU host logic on the CPU: site sharding + the result gather, world_size 2 and 3 over gloo.

The forward itself is stood in for by the CPU oracle (allowed in tests/): every rank runs its shard of sites, the
per-site records are all-gathered, and rank 0's gathered result must equal the unsharded run: same genotype calls,
same shapes and order, floats within the oracle's own batch-size noise (its reduceSlots is a cumulative sum over the
whole batch, python/MixtureOfExpertsAdvanced.py:29-34, so its rounding depends on what shares the batch).
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from hello_b200 import arch, shard, synth, weights   # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _forward_shard(cfg, params, pl, sh):
    """Oracle forward + wrapper tail on one shard; returns the tensors hello_moe_forward would produce."""
    from oracle import hello_oracle as O
    S = sh.s1 - sh.s0
    if S == 0:
        z = torch.zeros
        return (z((0, 2), dtype=torch.int32), z(0), z((0, 3)), z((4, 0)), z(0, dtype=torch.float64), z((3, 0)))
    tensors, nrpa = [], []
    for t in range(len(cfg.read_cin)):
        r0, r1 = sh.read_range[t]
        tensors.append(pl.reads[t][r0:r1].transpose(1, 2).contiguous())
        nrpa.append(torch.diff(sh.allele_read_off[t]).tolist())
    if len(tensors) == 1:
        tensors.append(None)
        nrpa.append(None)
    naps = torch.diff(sh.site_allele_off).tolist()
    res = O.OracleModel(cfg, params).forward(tuple(tensors), naps, tuple(nrpa), pl.ref_onehot[sh.s0:sh.s1])
    post = O.batched_posteriors(cfg, res, naps)
    if cfg.returns_meta:
        logits = torch.stack([e.reshape(-1) for e in res[0]])
    else:
        logits = torch.zeros((3, sum(naps)))
        logits[0 if cfg.xattn_present[0] else 2] = res.reshape(-1)
    best_pair = torch.tensor([p[3] for p in post], dtype=torch.int32).reshape(-1, 2)
    best_prob = torch.tensor([p[4] for p in post], dtype=torch.float32)
    meta = torch.stack([p[2] for p in post]).float()
    pair = torch.stack([torch.cat([p[0] for p in post])] + [torch.cat([p[1][e] for p in post]) for e in range(3)])
    mix64 = torch.tensor([v for p in post for v in O.remix_float64(p[1], p[2])], dtype=torch.float64)
    return best_pair, best_prob, meta, pair, mix64, logits


def _worker(rank, world, port, cfg_name, n_sites, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = arch.CONFIGS[cfg_name]
        params = weights.init_params(cfg, seed=13)
        pl = synth.make_pileups(n_sites, coverage=7, channels=cfg.read_cin, seed=91)     # same batch on every rank
        cost = shard.site_costs(cfg, pl.site_allele_off.numpy(), [o.numpy() for o in pl.allele_read_off])
        ranges = shard.balanced_ranges(cost, world)
        s0, s1 = ranges[rank]
        sh = shard.take_shard(pl.site_allele_off, pl.allele_read_off, s0, s1)
        got = shard.gather_site_results(*_forward_shard(cfg, params, pl, sh))
        if rank == 0:
            whole = shard.take_shard(pl.site_allele_off, pl.allele_read_off, 0, pl.n_sites)
            bp, bq, meta, pair, mix64, logits = _forward_shard(cfg, params, pl, whole)
            close = lambda a, b: a.shape == b.shape and torch.allclose(a, b, rtol=0, atol=2e-4)
            ok = (torch.equal(got.best_pair, bp) and close(got.best_prob, bq) and close(got.meta, meta)
                  and close(got.pair_prob, pair) and close(got.pair_mix64, mix64)
                  and close(got.logits, logits))
            with open(out_path, "w") as f:
                f.write("ok" if ok else "mismatch")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,cfg_name,n_sites", [(2, "single_tech", 9), (3, "hybrid_ensemble2", 7), (2, "single_tech", 1)])
def test_sharded_forward_equals_unsharded(tmp_path, world, cfg_name, n_sites):
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(world, _free_port(), cfg_name, n_sites, out), nprocs=world, join=True)
    assert open(out).read() == "ok"


def test_balanced_ranges_cover_and_balance():
    rng = np.random.default_rng(5)
    cost = rng.integers(1, 100, size=1000).astype(np.float64)
    for world in (1, 2, 3, 8):
        r = shard.balanced_ranges(cost, world)
        assert r[0][0] == 0 and r[-1][1] == 1000
        assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
        loads = np.array([cost[a:b].sum() for a, b in r])
        assert loads.max() - loads.min() <= 2 * cost.max()
    # fewer sites than ranks: nothing lost, nothing duplicated
    r = shard.balanced_ranges(np.ones(3), 8)
    assert sum(b - a for a, b in r) == 3 and all(b >= a for a, b in r)
    assert shard.balanced_ranges(np.zeros(0), 4) == [(0, 0)] * 4


def test_site_costs_follow_the_flop_model():
    cfg = arch.CONFIGS["single_tech"]
    f_read, f_allele, f_site = arch.flops_model(cfg)
    sao = np.array([0, 2, 3])
    aro = np.array([0, 5, 9, 10])
    cost = shard.site_costs(cfg, sao, [aro])
    assert cost.tolist() == [9 * f_read[0] + 2 * f_allele + f_site, 1 * f_read[0] + 1 * f_allele + f_site]


def _gatherer_worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ragged shard sizes, one rank without sites
        shapes = [(5, 8, 13), (0, 0, 0), (3, 4, 5)][:world]
        S, A, P = shapes[rank]
        g = shard.SiteGatherer("cpu", slots=2).plan(S, A, P)
        ok = g.collectives == 1
        for step in range(3):                       # steady state: one collective per step, alternating slots
            slot = step % 2
            v = g.result_views(slot)
            for k, (name, dt, _) in enumerate(shard.RESULT_FIELDS):
                v[name].copy_((torch.arange(v[name].numel()).reshape(v[name].shape) + 1000 * rank + 10 * k + step).to(dt))
            before = g.collectives
            g.gather_async(slot)
            g.wait()
            ok = ok and g.collectives == before + 1
            got = g.concat(slot)
            for r in range(world):
                rv = g.rank_views(slot, r)
                Sr, Ar, Pr = shapes[r]
                ok = ok and rv["logits"].shape == (3, Ar) and rv["pair_prob"].shape == (4, Pr) and rv["call_pair"].shape == (Sr, 5, 2)
                for k, (name, dt, _) in enumerate(shard.RESULT_FIELDS):
                    want = (torch.arange(rv[name].numel()).reshape(rv[name].shape) + 1000 * r + 10 * k + step).to(dt)
                    ok = ok and torch.equal(rv[name], want)
            ok = ok and got.best_pair.shape == (sum(s[0] for s in shapes), 2) and got.logits.shape == (3, sum(s[1] for s in shapes))
            ok = ok and got.pair_prob.shape == (4, sum(s[2] for s in shapes)) and got.call_qual.dtype == torch.float64
        if rank == 0:
            with open(out_path, "w") as f:
                f.write("ok" if ok else "mismatch")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_packed_gather_is_one_collective_per_step(tmp_path, world):
    """SiteGatherer: the nine per-site result tensors are views into one buffer per rank, the counts are exchanged once,
    every step is a single all_gather, ragged (and empty) shards come back in rank == site order."""
    out = str(tmp_path / "result.txt")
    mp.spawn(_gatherer_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert open(out).read() == "ok"


def test_packed_layout_is_aligned_and_disjoint():
    layout, total = shard.packed_layout(7, 11, 19)
    spans = sorted((off, off + int(np.prod(shp)) * torch.empty(0, dtype=dt).element_size()) for off, shp, dt in layout.values())
    assert all(a[1] <= b[0] for a, b in zip(spans[:-1], spans[1:])) and spans[-1][1] <= total
    assert all(off % 256 == 0 for off, _, _ in layout.values())
    v = shard.packed_views(torch.zeros(total, dtype=torch.uint8), 7, 11, 19)
    assert v["pair_mix64"].dtype == torch.float64 and v["pair_mix64"].shape == (19,) and v["call_pair"].shape == (7, 5, 2)
