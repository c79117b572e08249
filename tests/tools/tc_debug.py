#!/usr/bin/env python
"""Layer-by-layer check of the fused tcgen05 read convolver against the CPU oracle (run on a GPU box).

    python tests/tools/tc_debug.py [bf16x3|bf16] [n_reads]

Prints, for each of the 17 layer phases, the max abs error of the kernel's post-activation values against the
oracle's fp32 values of the same layer, plus the error of the final [R,36,64] features.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from hello_b200 import arch, model, synth, weights          # noqa: E402
from helpers import readconv_phase_from_dump, readconv_phase_reference   # noqa: E402

if __name__ == "__main__":
    prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
    n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    cfg = arch.CONFIGS["single_tech"]
    params = weights.init_params(cfg, seed=13)
    eng = model.MoEEngine(cfg, params, device="cuda:0", precision=prec)
    pl = synth.make_pileups(max(2, n_reads // 8), coverage=12, channels=cfg.read_cin, seed=21)
    reads = pl.reads[0][:n_reads]
    print("precision", prec, "reads", tuple(reads.shape))
    ref = readconv_phase_reference(cfg, params, reads)
    for ph in range(17):
        out, dbg = eng.readconv_debug(reads, ph)
        torch.cuda.synchronize()
        got = readconv_phase_from_dump(dbg.cpu(), ph, reads.shape[0], ref[ph].shape[2])
        err, scale = (got - ref[ph]).abs().max().item(), ref[ph].abs().max().item()
        print("phase %2d  shape %s  max|err| %.3e  max|ref| %.3e  rel %.2e" % (ph, tuple(ref[ph].shape[1:]), err,
                                                                            scale, err / max(scale, 1e-30)))
    err = (out.cpu().transpose(1, 2) - ref[-1]).abs().max().item()
    print("final features max|err| %.3e (max|ref| %.3e)" % (err, ref[-1].abs().max().item()))
