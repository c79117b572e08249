import cProfile, pstats, sys, io, time
sys.path.insert(0, '.')
import torch
from hello_b200 import arch, model, synth, weights
cfg = arch.CONFIGS["single_tech"]
params = weights.init_params(cfg, seed=13)
pl = synth.make_pileups(300, coverage=30, channels=cfg.read_cin, seed=13)
fds = [pl.site_feature_dict(s) for s in range(pl.n_sites)]
net = model.MoEMergedWrapperB200(model.MoEAttentionB200(cfg, params, device="cuda:0", precision="bf16x3")).eval()
net.providePredictions = True
for fd, seg in fds[:30]: net(fd, seg)
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for fd, seg in fds: 
    r = net(fd, seg); float(next(iter(r[0].values())))
pr.disable()
print("per call ms (under profile)", (time.perf_counter()-t0)/len(fds)*1e3)
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35); print(s.getvalue()[:6000])
