#!/usr/bin/env python
"""Pair-scoring throughput of the non-headline shapes (DiT-XL/2, SDXL, CLIP/DINO-like, SD-1.5 other layers), one line per
shape; the library is taken from DIFFSIM_B200_LIB (A/B of builds: run once per library)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffsim_b200 import ops, synth

def case(name, B, H, S, D, n_img, dtype=torch.float16):
    q, k, v = synth.device_cache(B, H, S, D, n_img, dtype, "cuda")
    T = n_img // 3
    pairs = torch.tensor([(3 * t, 3 * t + 1) for t in range(T)] + [(3 * t, 3 * t + 2) for t in range(T)], dtype=torch.int32, device="cuda")
    for _ in range(2): ops.aas_pairs(q, k, v, pairs, "cosine")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ops.aas_pairs(q, k, v, pairs, "cosine")
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    P = pairs.shape[0]
    print(f"[{name}] ({B},{H},{S},{D}): {P} pairs {ms:.3f} ms {P / ms * 1e3:.0f} pairs/s {4 * P * 4 * B * H * S * S * D / ms / 1e9:.0f} TFLOP/s", flush=True)

print(os.path.basename(os.environ.get("DIFFSIM_B200_LIB", "default")))
case("DiT-XL/2", 2, 16, 256, 72, 768)
case("SDXL up0", 2, 20, 1024, 64, 96)
case("SDXL up1", 2, 10, 4096, 64, 24)
case("SD-1.5 up1", 2, 8, 1024, 80, 192)
case("SD-1.5 up2", 2, 8, 4096, 40, 24)
case("SD-1.5 up0", 2, 8, 256, 160, 768)
