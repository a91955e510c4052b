#!/usr/bin/env python
"""Small attention workload for an `ncu --set full` capture: 3 launches of ds_aas_pairs on 256 SD-1.5 up0 pairs.
usage: ncu --set full --clock-control none --import-source on -k regex:aas_attn -s 2 -c 1 -o gpurun_out/x python tools/ncu_attn.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffsim_b200 import ops, synth

B, H, S, D = 2, 8, 256, 160
n_img = 768
q, k, v = synth.device_cache(B, H, S, D, n_img, torch.float16, "cuda")
T = n_img // 3
pairs = torch.tensor([(3 * t, 3 * t + 1) for t in range(T)], dtype=torch.int32, device="cuda")
for _ in range(3):
    out = ops.aas_pairs(q, k, v, pairs, "cosine")
torch.cuda.synchronize()
print("ok", float(out.mean()))
