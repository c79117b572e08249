#!/usr/bin/env python
"""Timeline of the fused tcgen05 head kernel on one SM (run on a GPU box):  python tools/head_trace.py [xattn0|compressor0]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hello_b200 import arch, model, weights          # noqa: E402

if __name__ == "__main__":
    net = sys.argv[1] if len(sys.argv) > 1 else "xattn0"
    cfg = arch.CONFIGS["single_tech"]
    eng = model.MoEEngine(cfg, weights.init_params(cfg, seed=13), device="cuda:0", precision="bf16x3")
    shape, ngrp, per = ((36, 64), 2, 12) if net.startswith("compressor") else ((18, 128), 1, 12)
    n = 148 * per * 12
    x = (torch.randn((n,) + shape, generator=torch.Generator().manual_seed(1)) * 20).cuda()
    co, lo = arch.net_out_shape(cfg.networks()[net], shape[0])
    out = torch.empty((n, lo, co), dtype=torch.float32, device="cuda")
    tr = torch.zeros((16, ngrp, 12, 4), dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        rc = eng.lib.hello_moe_headconv_debug(eng.handle, weights.NET_IDS[net], x.data_ptr(), n, -2, out.data_ptr(), tr.data_ptr(),
                                              C.c_void_p(st))
        assert rc == 0, eng.lib.hello_moe_last_error(eng.handle)
    torch.cuda.synchronize()
    t = tr.cpu()
    print("cycles per item (steady):", float(t[11, 0, 6, 3] - t[1, 0, 6, 3]) / 10)
    sel = t[2:12]
    print("operand load: %.0f" % (sel[:, :, 11, 1] - sel[:, :, 11, 0]).float().mean().item())
    print("phase | issue (start->end)  end->acc seen  epilogue  epi end->next issue start")
    for ph in range(7):
        f = lambda a, b, p=ph: (sel[:, :, p, a] - sel[:, :, p, b]).float().mean().item()
        nxt = (sel[:, :, ph + 1, 0] - sel[:, :, ph, 3]).float().mean().item() if ph < 6 else float("nan")
        print("%5d | %8.0f %8.0f %8.0f %8.0f" % (ph, f(1, 0), f(2, 1), f(3, 2), nxt))
