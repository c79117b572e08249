#!/usr/bin/env python
"""Timeline of the fused tcgen05 read convolver on one SM (run on a GPU box).

    python tools/tc_trace.py [bf16x3|bf16] [items_per_cta]

CTA 0 stamps clock64() per (work item, read group, layer phase): MMA issue start / end, accumulators seen by the
epilogue, epilogue end (readconv_tc.cuh: trace_slot).  Prints per-phase averages in SM cycles.
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hello_b200 import arch, model, weights          # noqa: E402

NG, NPH, ITEMS = 3, 17, 16
STRIDE = 21                     # tc::MAX_PHASES: phase slots per (item, group) in the trace record

if __name__ == "__main__":
    prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
    per_cta = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    cfg = arch.CONFIGS["single_tech"]
    eng = model.MoEEngine(cfg, weights.init_params(cfg, seed=13), device="cuda:0", precision=prec)
    n = 148 * 3 * NG * per_cta
    g = torch.Generator().manual_seed(1)
    reads = torch.randint(0, 256, (n, 150, 6), generator=g, dtype=torch.uint8).cuda()
    out = torch.empty((n, 36, 64), dtype=torch.float32, device="cuda")
    tr = torch.zeros((ITEMS, NG, STRIDE, 8), dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        rc = eng.lib.hello_moe_readconv_debug(eng.handle, 0, reads.data_ptr(), n, 1, -2, out.data_ptr(), tr.data_ptr(),
                                              C.c_void_p(st))
        assert rc == 0, eng.lib.hello_moe_last_error(eng.handle)
    torch.cuda.synchronize()
    t = tr.cpu()
    items = min(ITEMS, per_cta)
    t0 = int(t[0, :, 0, 0].min())
    print("item span (cycles):", [int(t[i, :, NPH - 1, 3].max() - t[i, :, 0, 0].min()) for i in range(items)])
    print("steady-state cycles per item: %.0f" % ((int(t[items - 1, :, NPH - 1, 3].max()) - int(t[1, :, NPH - 1, 3].max())) / (items - 2)))
    print("phase |  issue   mma_wait(issue_end->acc_seen)   epilogue   act->next issue start | per group")
    sel = t[2:items]
    for ph in range(NPH):
        issue = (sel[:, :, ph, 1] - sel[:, :, ph, 0]).float().mean().item()
        wait = (sel[:, :, ph, 2] - sel[:, :, ph, 1]).float().mean().item()
        total = (sel[:, :, ph, 2] - sel[:, :, ph, 0]).float().mean().item()
        epi = (sel[:, :, ph, 3] - sel[:, :, ph, 2]).float().mean().item()
        nxt = (sel[:, :, ph + 1, 0] - sel[:, :, ph, 3]).float().mean().item() if ph + 1 < NPH else float("nan")
        f = lambda a, b: (sel[:, :, ph, a] - sel[:, :, ph, b]).float().mean().item()
        inner = "" if ph == 2 else "  [epilogue: first ld %5.0f | block0 %5.0f | blocks1-3 %5.0f | tail %5.0f | fence+arrive %4.0f]" % (
            f(4, 2), f(5, 4), f(6, 5), f(3, 6), f(7, 3))
        print("%5d | %7.0f %7.0f (issue->acc %7.0f) %9.0f %9.0f%s" % (ph, issue, wait, total, epi, nxt, inner))
    # one item in full: when is each group in which state
    i = 3
    print("item %d timeline (cycles since first issue of the item), group: [issue_start, acc_seen, epi_end] per phase" % i)
    base = int(t[i, :, 0, 0].min())
    for gi in range(NG):
        print(" g%d " % gi + " ".join("%d/%d/%d" % (int(t[i, gi, ph, 0]) - base, int(t[i, gi, ph, 2]) - base,
                                                     int(t[i, gi, ph, 3]) - base) for ph in range(NPH)))
    # item boundary: last epilogue of item i (incl. the allele sum) -> first issue of item i+1, per group
    gap = (t[3:items, :, 0, 0] - t[2:items - 1, :, NPH - 1, 3]).float()
    print("item boundary, epilogue end of the last phase -> first MMA issue of the next item (cycles): mean %.0f, per group %s"
          % (gap.mean().item(), [round(x) for x in gap.mean(0).tolist()]))
    idle = (t[3:items, :, 0, 0].min(1).values - t[2:items - 1, :, NPH - 1, 1].max(1).values).float()
    print("tensor pipe: last MMA issue of item i -> first MMA issue of item i+1 (cycles): mean %.0f" % idle.mean().item())
    # where the boundary goes (spare phase slot NPH): times relative to the previous item's last epilogue end
    prev_end = t[2:items - 1, :, NPH - 1, 3]
    names = ["operand store start", "operand stored", "operand arrive", "issuer: weights landed",
             "issuer: operand seen", "issuer: token received", "first MMA issue"]
    vals = [t[3:items, :, NPH, k] for k in (0, 1, 2, 3, 4, 5)] + [t[3:items, :, 0, 0]]
    for nm, v in zip(names, vals):
        print("  %-32s +%6.0f" % (nm, (v - prev_end).float().mean().item()))

