"""Cross-process batching shim (hello_b200/serving.py): host logic on the CPU with an injected scorer (the oracle), and
the real GPU server against direct per-site calls."""
import multiprocessing as mp
import threading

import numpy as np
import pytest
import torch

from helpers import params_for
from hello_b200 import arch, serving, synth


def oracle_run_batch(cfg):
    from oracle import hello_oracle as O
    orc = O.OracleModel(cfg, params_for(cfg))

    def run(reads, offs, sao, rank, ref):
        naps = torch.diff(sao).tolist()
        tensors = tuple(r.transpose(1, 2) for r in reads) + ((None,) if len(reads) == 1 else ())
        nrpa = tuple(torch.diff(o).tolist() for o in offs) + ((None,) if len(reads) == 1 else ())
        res = orc.forward(tensors, naps, nrpa, ref)
        post = O.batched_posteriors(cfg, res, naps, allele_rank=rank)
        pair = torch.cat([torch.stack([p[0]] + list(p[1])) for p in post], dim=1).numpy()
        S = len(naps)
        return {"pair_prob": pair, "meta": torch.stack([torch.as_tensor(p[2]).float() for p in post]).numpy(),
                "best_pair": np.array([p[3] for p in post], np.int32), "call_pair": np.zeros((S, 5, 2), np.int32),
                "call_qual": np.zeros((S, 5)), "best_expert": np.zeros(S, np.int32)}
    return run


def _client(net, pl, sites, out_q):
    net.providePredictions = True
    res = []
    for s in sites:
        fd, seg = pl.site_feature_dict(s, allele_names=["T", "AC", "A", "G"][:len(pl.site_feature_dict(s)[0])])
        r = net(fd, seg)
        res.append((s, [{k: float(v) for k, v in d.items()} for d in r[:4]], r[4].tolist()))
    out_q.put(res)


def test_batch_planner_roundtrip():
    reqs = [serving.SiteRequest(0, 1, ["C", "A"], [[np.ones((2, 150, 6), np.uint8), np.zeros((1, 150, 6), np.uint8)]], None),
            serving.SiteRequest(1, 7, ["G"], [[np.full((3, 150, 6), 7, np.uint8)]], None)]
    plan = serving.BatchPlanner(1)
    for r in reqs:
        plan.add(r)
    reads, offs, sao, rank, ref = plan.build()
    assert reads[0].shape == (6, 150, 6) and offs[0].tolist() == [0, 2, 3, 6] and sao.tolist() == [0, 2, 3]
    assert rank.tolist() == [1, 0, 0] and ref is None
    pp = np.arange(16, dtype=np.float32).reshape(4, 4)
    parts = plan.split(pp, np.eye(3, dtype=np.float32)[:2], np.zeros((2, 2), np.int32), np.zeros((2, 5, 2), np.int32),
                       np.zeros((2, 5)), np.array([0, 2], np.int32))
    assert parts[0][1]["pair_prob"].shape == (4, 3) and parts[1][1]["pair_prob"].tolist() == [[3.0], [7.0], [11.0], [15.0]]
    assert parts[1][1]["best_expert"] == 2 and parts[1][0].seq == 7


@pytest.mark.parametrize("name", ["single_tech", "hybrid_ensemble2"])
def test_serve_loop_batches_requests_from_several_processes(name):
    """Three forked client processes score sites through RemoteNetwork; the server thread batches whatever is pending
    and every client gets exactly the per-site results of the reference wrapper call."""
    from oracle import hello_oracle as O
    cfg = arch.CONFIGS[name]
    pl = synth.make_pileups(18, coverage=6, channels=cfg.read_cin, seed=17)
    ctx = mp.get_context("fork")
    requests, responses = ctx.Queue(), [ctx.Queue() for _ in range(3)]
    stats = {}
    th = threading.Thread(target=serving.serve_loop, args=(oracle_run_batch(cfg), len(cfg.read_cin), requests, responses, 8, 0.05,
                                                            stats))
    th.start()
    out_q = ctx.Queue()
    procs = []
    for c in range(3):
        net = serving.RemoteNetwork(requests, responses[c], c, len(cfg.read_cin), cfg.meta == "meta_convolver_ref")
        p = ctx.Process(target=_client, args=(net, pl, list(range(c, 18, 3)), out_q))
        p.start()
        procs.append(p)
    got = {}
    for _ in procs:
        for s, dicts, meta in out_q.get(timeout=120):
            got[s] = (dicts, meta)
    for p in procs:
        p.join(timeout=30)
    requests.put(serving._STOP)
    th.join(timeout=30)
    assert len(got) == 18 and stats["sites"] == 18 and stats["batches"] < 18, stats
    orc = O.OracleModel(cfg, params_for(cfg))
    for s in range(18):
        fd, seg = pl.site_feature_dict(s, allele_names=["T", "AC", "A", "G"][:len(pl.site_feature_dict(s)[0])])
        ref = O.wrapper_forward(orc, fd, seg, provide_predictions=True)
        for dg, dr in zip(got[s][0], ref[:4]):
            assert list(dg.keys()) == list(dr.keys())
            for k in dg:
                assert abs(dg[k] - float(dr[k])) < 5e-5
        assert np.allclose(got[s][1], ref[4].numpy(), atol=5e-5)


@pytest.mark.gpu
def test_gpu_scoring_server_equals_direct_calls():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from hello_b200 import model
    cfg = arch.CONFIGS["single_tech"]
    params = params_for(cfg)
    pl = synth.make_pileups(24, coverage=8, channels=cfg.read_cin, seed=23)
    direct = model.MoEMergedWrapperB200(model.MoEAttentionB200(cfg, params, device="cuda:0", precision="bf16x3"))
    direct.providePredictions = True
    want = {}
    for s in range(24):
        fd, seg = pl.site_feature_dict(s, allele_names=["T", "AC", "A", "G"][:len(pl.site_feature_dict(s)[0])])
        r = direct(fd, seg)
        want[s] = ([{k: float(v) for k, v in d.items()} for d in r[:4]], r[4].tolist())
    ctx = mp.get_context("fork")
    with serving.ScoringServer("single_tech", params, n_clients=2, device="cuda:0", precision="bf16x3", max_sites=16,
                               max_wait_s=0.02) as server:
        out_q = ctx.Queue()
        procs = [ctx.Process(target=_client, args=(server.client(c), pl, list(range(c, 24, 2)), out_q)) for c in range(2)]
        for p in procs:
            p.start()
        got = {}
        for _ in procs:
            for s, dicts, meta in out_q.get(timeout=300):
                got[s] = (dicts, meta)
        for p in procs:
            p.join(timeout=30)
    assert len(got) == 24
    for s in range(24):
        assert got[s][0] == want[s][0], "batched-on-the-server results must equal the direct per-site call bit for bit"
        assert got[s][1] == want[s][1]


def test_one_bad_site_does_not_poison_the_batch():
    """A malformed request (an allele without rows) is answered with its exception on its own; a site that makes the
    scorer fail is isolated by re-scoring the batch site by site: every other client still gets its result."""
    cfg = arch.CONFIGS["single_tech"]
    pl = synth.make_pileups(6, coverage=5, channels=cfg.read_cin, seed=5)
    good = oracle_run_batch(cfg)

    def run(reads, offs, sao, rank, ref):
        if int((reads[0][:, 0, 0] == 255).sum()) > 0:          # poisoned site: first byte of a row is 255
            raise RuntimeError("scorer rejected the batch")
        return good(reads, offs, sao, rank, ref)

    import queue as pyqueue
    requests, responses = pyqueue.Queue(), [pyqueue.Queue() for _ in range(3)]
    reqs = []
    for s in range(6):
        fd, seg = pl.site_feature_dict(s)
        reqs.append(serving.request_from_feature_dict(s % 3, s + 1, fd, None, 1))
    reqs[1].reads[0][0] = np.zeros((0, 150, 6), np.uint8)       # malformed: allele without rows
    reqs[4].reads[0][0] = reqs[4].reads[0][0].copy()
    reqs[4].reads[0][0][0, 0, 0] = 255                          # makes the scorer fail
    for r in reqs:
        requests.put(r)
    requests.put(serving._STOP)
    stats = {}
    serving.serve_loop(run, 1, requests, responses, max_sites=16, max_wait_s=0.05, stats=stats)
    answers = {}
    for c in range(3):
        while not responses[c].empty():
            seq, site = responses[c].get()
            answers[seq] = site
    assert sorted(answers) == [1, 2, 3, 4, 5, 6]
    assert isinstance(answers[2], ValueError) and "at least one" in str(answers[2])
    assert isinstance(answers[5], RuntimeError)
    for seq in (1, 3, 4, 6):
        assert isinstance(answers[seq], dict) and answers[seq]["pair_prob"].shape[0] == 4


def _die_quietly(requests):
    requests.get()                                              # take the request, never answer, exit


def test_remote_network_notices_a_dead_server():
    """The client polls with a timeout and checks the server's process state: a server that died mid-request raises a
    clear error in the worker instead of hanging it forever."""
    cfg = arch.CONFIGS["single_tech"]
    pl = synth.make_pileups(1, coverage=4, channels=cfg.read_cin, seed=2)
    ctx = mp.get_context("fork")
    requests, response = ctx.Queue(), ctx.Queue()
    server = ctx.Process(target=_die_quietly, args=(requests,))
    server.start()
    net = serving.RemoteNetwork(requests, response, 0, 1, False, server_pid=server.pid)
    fd, seg = pl.site_feature_dict(0)
    with pytest.raises(RuntimeError, match="scoring server process"):
        net(fd, seg)
    server.join(timeout=10)
    # a hung (alive but silent) server is caught by the heartbeat
    hb = ctx.Value("d", 1.0, lock=False)                        # last heartbeat: 1970
    net2 = serving.RemoteNetwork(requests, response, 0, 1, False, server_pid=None, heartbeat=hb, dead_after_s=5.0)
    with pytest.raises(RuntimeError, match="has not answered"):
        net2(fd, seg)
