#!/usr/bin/env python
"""K3 workload for an ncu capture: ds_simmat on the Sref diffeats shape (2032 x 655360 fp16).
usage: ncu --set full --clock-control none --import-source on -k regex:gemm_tn -s 1 -c 1 -o gpurun_out/x python tools/ncu_simmat.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffsim_b200 import ops

n, L = int(os.environ.get("SIMMAT_N", 2032)), int(os.environ.get("SIMMAT_L", 655360))
f = torch.empty(n, L, dtype=torch.float16, device="cuda")
for i in range(0, n, 127):
    f[i:i + 127] = torch.randn(min(127, n - i), L, device="cuda").half()
for _ in range(3):
    c = ops.simmat(f, None, "cosine")
torch.cuda.synchronize()
print("ok", float(c.diagonal().mean()))
