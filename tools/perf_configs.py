#!/usr/bin/env python
"""Throughput of the other BASELINE.json configs: DiT-XL/2 pairs, SDXL shapes, N x N AAS matrix, K3 feature GEMM, K2."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffsim_b200 import ops, synth, scoring


def timeit(fn, iters=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def pairs_case(name, B, H, S, D, n_img, dtype=torch.float16):
    q, k, v = synth.device_cache(B, H, S, D, n_img, dtype, "cuda")
    T = n_img // 3
    pairs = torch.tensor([(3 * t, 3 * t + 1) for t in range(T)] + [(3 * t, 3 * t + 2) for t in range(T)], dtype=torch.int32, device="cuda")
    ms = timeit(lambda: ops.aas_pairs(q, k, v, pairs, "cosine"))
    P = pairs.shape[0]
    fl = 4 * P * 4 * B * H * S * S * D
    print(f"[{name}] ({B},{H},{S},{D}) {str(dtype)[6:]}: {P} pairs {ms:.3f} ms  {P / ms * 1e3:.0f} pairs/s  {fl / ms / 1e9:.0f} TFLOP/s", flush=True)


pairs_case("cfg5 DiT-XL/2", 2, 16, 256, 72, 768)
pairs_case("cfg4 SDXL up0", 2, 20, 1024, 64, 96)
pairs_case("cfg4 SDXL up1", 2, 10, 4096, 64, 24)
pairs_case("cfg4 literal", 2, 20, 4096, 64, 12)
pairs_case("sd15 up1", 2, 8, 1024, 80, 192)
pairs_case("sd15 mid", 2, 8, 64, 160, 1536)
pairs_case("sd15 up0 bf16", 2, 8, 256, 160, 768, torch.bfloat16)

# cfg3-shaped: N x N AAS matrix (N reduced to keep the run short; per-attention cost is what matters)
B, H, S, D = 2, 8, 256, 160
N = 256
q, k, v = synth.device_cache(B, H, S, D, N, torch.float16, "cuda")
ms = timeit(lambda: ops.aas_matrix(q, k, v, k, v, "cosine"), iters=2, warmup=1)
attn = N * N + N  # cross (incl. diagonal) + self per row chunk (approx)
print(f"[cfg3 AAS matrix] N={N}: {ms:.2f} ms  {N * N / ms * 1e3:.0f} directional scores/s  ~{attn * 4 * B * H * S * S * D / ms / 1e9:.0f} TFLOP/s "
      f"(Sref N=2032 would take ~{ms * (2032 / N) ** 2 / 1e3:.1f} s on one GPU)", flush=True)

# K3: feature GEMM, diffeats-shaped L = 655360
for n, L in ((512, 655360), (2032, 65536), (2032, 655360)):
    try:
        f = torch.randn(n, L, device="cuda", dtype=torch.float16)
        ms = timeit(lambda: ops.simmat(f, f, "cosine"), iters=2, warmup=1)
        print(f"[K3 simmat] N={n} L={L}: {ms:.2f} ms  {2 * n * n * L / ms / 1e9:.0f} TFLOP/s", flush=True)
        del f
    except Exception as e:
        print("[K3 simmat]", n, L, "failed:", repr(e)[:200])
f = torch.randn(2032, 65536, device="cuda", dtype=torch.float16)
ms = timeit(lambda: (f @ f.T), iters=2, warmup=1)
print(f"[torch matmul fp16 2032x2032x65536] {ms:.2f} ms {2 * 2032 * 2032 * 65536 / ms / 1e9:.0f} TFLOP/s")
del f

# K4: QKV projection of the hooked layer, SD-1.5 up0: rows = images x 512, C = 1280 -> 3 x 1280
for n_img in (96, 768):
    hid, w = synth.device_hidden(2, 8, 256, 160, n_img, torch.float16, "cuda")
    outs = [torch.empty(n_img, 2, 256, 1280, dtype=torch.float16, device="cuda") for _ in range(3)]
    ms = timeit(lambda: ops.qkv_project(hid, w, None, 3, out=outs), iters=5, warmup=2)
    fl = 2 * n_img * 512 * 1280 * 3840
    print(f"[K4 qkv_project] {n_img} images (rows {n_img * 512} x 1280 -> 3840): {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s  "
          f"{n_img / ms * 1e3:.0f} images/s", flush=True)
    h2 = hid.view(-1, 1280)
    ms = timeit(lambda: h2 @ w.t(), iters=5, warmup=2)
    print(f"[torch matmul same shape] {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s", flush=True)
    del hid, outs, h2
