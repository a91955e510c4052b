#!/usr/bin/env python
"""K3 A/B: ds_simmat with the row-major operand (blocked = 0) against the k-blocked copy (blocked = 1), self (symmetric) and
rows x other (blocked = 2: copy and GEMM back to back; 1: staged, the copy of stage s + 1 under the GEMM of stage s), at the
Sref diffeats shape and a ragged one; the paths must agree bitwise (same MMA order)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from diffsim_b200 import _native as N
from diffsim_b200 import ops

lib = N.load()
shapes = ((2032, 655360), (2032, 163840), (777, 100007 // 8 * 8 + 8), (4096, 65536))
for n, L in shapes:
    f = torch.empty(n, L, dtype=torch.float16, device="cuda")
    for i in range(0, n, 127):
        f[i:i + 127] = torch.randn(min(127, n - i), L, device="cuda").half()
    g = f.clone()
    res = {}
    for blocked in (0, 2, 1, 0, 2, 1):
        lib.ds_debug_set_simmat_blocked(blocked)
        out = []
        for name, fn in (("self", lambda: ops.simmat(f, None, "cosine")), ("full", lambda: ops.simmat(f, g, "cosine"))):
            for _ in range(3):
                c = fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            tf = 2.0 * n * n * L / (ms * 1e-3) / 1e12
            out.append(f"{name} {ms:.3f} ms ({tf:.0f} TFLOP/s algorithmic)")
            res.setdefault(name, {})[blocked] = c
        print(f"N={n} L={L} blocked={blocked}: " + ", ".join(out), flush=True)
    for name in res:
        a, b = res[name][0], res[name][1]
        print(f"   {name}: bitwise equal {bool(torch.equal(a, b))} / {bool(torch.equal(a, res[name][2]))}, max |diff| {float((a - b).abs().max()):.2e}, "
              f"diag err {float((b.diagonal() - 1).abs().max()):.1e}", flush=True)
    del f, g, res
lib.ds_debug_set_simmat_blocked(-1)
