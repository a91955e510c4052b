#!/usr/bin/env python
"""Wait accounting of CTA 0 of the attention kernel (library built with DS_EXTRA_NVCC_FLAGS=-DDS_ACCT): per role, the SM clocks
spent in each kind of wait / section, per 128-row half-step."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffsim_b200 import ops, synth, _native as N

CAP = 64
B, H, S, D = 2, 8, 256, 160
n_img = 768
q, k, v = synth.device_cache(B, H, S, D, n_img, torch.float16, "cuda")
pairs = torch.tensor([(3 * t, 3 * t + 1) for t in range(n_img // 3)] + [(3 * t, 3 * t + 2) for t in range(n_img // 3)],
                     dtype=torch.int32, device="cuda")
buf = torch.zeros(8 * CAP + 1, dtype=torch.int64, device="cuda")
lib = N.load()
ops.aas_pairs(q, k, v, pairs, "cosine")
torch.cuda.synchronize()
lib.ds_debug_set_trace(buf.data_ptr(), CAP)
ops.aas_pairs(q, k, v, pairs, "cosine")
torch.cuda.synchronize()
lib.ds_debug_set_trace(None, 0)
t = buf.cpu()
total = int(t[8 * CAP])
attn = 4 * pairs.shape[0]
halfsteps = attn * B * H * (S // 128) * (S // 128) / 148
names = {0: ("producer", ["wait kv_empty", "wait q_empty"]),
         1: ("mma", ["wait q_full", "wait s_free", "wait kv_full (QK)", "issue QK mmas", "wait p_full", "wait o_empty", "wait kv_full (PV)", "issue PV mmas"]),
         2: ("softmax warp 0", ["wait s_full", "S load + max", "quadrant barrier", "p_free wait.. exp + P store + arrive", "loop overhead", "wait p_free"]),
         3: ("softmax warp 15", ["wait s_full", "S load + max", "quadrant barrier", "exp + P store + arrive", "loop overhead", "wait p_free"]),
         4: ("epilogue warp 0", ["wait o_full", "O -> registers", "arithmetic + reduction"])}
print(f"CTA 0: {total} clk total, {halfsteps:.0f} half-steps -> {total / halfsteps:.0f} clk per half-step")
for slot, (role, labels) in names.items():
    acc = 0
    for i, lab in enumerate(labels):
        w = int(t[slot * CAP + i])
        n, c = w >> 40, w & ((1 << 40) - 1)
        acc += c
        print(f"  {role:16s} {lab:40s} {c / halfsteps:8.0f} clk/half-step  ({n} times, {c / max(n, 1):7.0f} each)")
    print(f"  {role:16s} {'accounted':40s} {acc / halfsteps:8.0f} clk/half-step")
