#!/usr/bin/env python
"""Bring-up checks for the CUDA kernels, one case per process (a trapped kernel poisons the CUDA context).

    python tools/gpu_check.py --case all          # runs every case in its own subprocess
    python tools/gpu_check.py --case attn_d160    # one case, in-process

Each case prints a compact error report (max abs / rel error, and WHERE the error sits: by row block and
column block) so that a wrong descriptor or swizzle can be localised from the log alone.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = ["reduce", "simmat", "attn_pv", "attn_qk", "attn_d160", "attn_d64", "attn_d72", "attn_d40", "attn_d80",
         "attn_d128", "attn_tails", "attn_long", "aas_pairs", "aas_groups", "aas_matrix", "perf"]


def report(name, got, ref, tol, row_block=32, col_block=32):
    import torch

    got = got.double().cpu()
    ref = ref.double().cpu()
    err = (got - ref).abs()
    scale = ref.abs().max().item() + 1e-30
    mx = err.max().item()
    ok = bool(torch.isfinite(got).all()) and mx <= tol * max(1.0, scale)
    print(f"[{name}] max_abs_err={mx:.3e} ref_absmax={scale:.3e} finite={bool(torch.isfinite(got).all())} -> {'OK' if ok else 'FAIL'}")
    if not ok and got.dim() >= 2:
        e2 = err.reshape(-1, got.shape[-2], got.shape[-1]).amax(0)
        R, Cc = e2.shape
        rb = [e2[r:r + row_block].max().item() for r in range(0, R, row_block)]
        cb = [e2[:, c:c + col_block].max().item() for c in range(0, Cc, col_block)]
        print(f"   err by row block({row_block}): " + " ".join(f"{x:.2e}" for x in rb))
        print(f"   err by col block({col_block}): " + " ".join(f"{x:.2e}" for x in cb))
        flat = err.reshape(-1)
        idx = int(flat.argmax())
        print(f"   worst flat index {idx}: got {got.reshape(-1)[idx].item():.6f} ref {ref.reshape(-1)[idx].item():.6f}")
        print("   got[0,0,:8]:", [round(x, 4) for x in got.reshape(-1, got.shape[-1])[0, :8].tolist()])
        print("   ref[0,0,:8]:", [round(x, 4) for x in ref.reshape(-1, ref.shape[-1])[0, :8].tolist()])
    return ok


def timeit(fn, iters=20, warmup=3):
    import torch

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


def case_reduce():
    import torch
    from diffsim_b200 import ops
    from oracle import aas_oracle as O

    ok = True
    g = torch.Generator().manual_seed(1)
    for dtype in (torch.float16, torch.bfloat16, torch.float32):
        for P, E in ((1, 655360), (5, 1000), (3, 8191), (16, 589824), (2, 7)):
            x = torch.randn(P, E, generator=g) * 0.7 + 0.2
            y = 0.6 * x + 0.4 * torch.randn(P, E, generator=g)
            xd, yd = x.to(dtype).cuda(), y.to(dtype).cuda()
            for mode in ("cosine", "mse", "minmax_cosine"):
                got = ops.pair_reduce(xd, yd, mode).cpu()
                ref = torch.tensor([O.similarity(xd[p].cpu(), yd[p].cpu(), mode) for p in range(P)])
                err = ((got.double() - ref).abs() / ref.abs().clamp_min(1e-6)).max().item()
                good = err < 2e-5
                ok &= good
                print(f"[reduce {dtype} P={P} E={E} {mode}] rel_err={err:.2e} {'OK' if good else 'FAIL'}")
    return ok


def case_simmat():
    import torch
    from diffsim_b200 import ops

    ok = True
    g = torch.Generator().manual_seed(2)
    for dtype in (torch.float16, torch.bfloat16):
        for nr, nc, L in ((128, 128, 64), (128, 128, 256), (160, 96, 200), (300, 300, 4096), (40, 520, 10240)):
            a = (torch.randn(nr, L, generator=g) * 0.5 + 0.1).to(dtype).cuda()
            b = (torch.randn(nc, L, generator=g) * 0.5 - 0.2).to(dtype).cuda()
            for mode in ("cosine", "minmax_cosine"):
                got = ops.simmat(a, b, mode)
                ad, bd = a.double(), b.double()
                if mode == "minmax_cosine":
                    ad = (ad - ad.amin(1, keepdim=True)) / (ad.amax(1, keepdim=True) - ad.amin(1, keepdim=True))
                    bd = (bd - bd.amin(1, keepdim=True)) / (bd.amax(1, keepdim=True) - bd.amin(1, keepdim=True))
                ref = (ad @ bd.T) / (ad.norm(dim=1, keepdim=True).clamp_min(1e-8) * bd.norm(dim=1).clamp_min(1e-8))
                ok &= report(f"simmat {dtype} {nr}x{nc}x{L} {mode}", got, ref, 2e-5, 32, 32)
    return ok


def _attn_inputs(B, H, Sq, Skv, D, dtype, seed=0, kind="random"):
    import torch

    g = torch.Generator().manual_seed(seed)

    def mk(S, std):
        mem = torch.randn(B, S, H * D, generator=g) * std
        return mem.to(dtype).cuda().view(B, S, H, D).transpose(1, 2)

    q, k, v = mk(Sq, 1.5), mk(Skv, 1.5), mk(Skv, 1.0)
    if kind == "ramp":     # later kv rows carry much larger logits: forces the lazy running maximum to move (O rescale path)
        ramp = (1.0 + 9.0 * torch.arange(Skv).float() / Skv).view(1, 1, Skv, 1).cuda()
        k = (k.float() * ramp).to(dtype)
    if kind == "pv":       # K = 0: uniform softmax, O = mean(V): isolates the P-write + PV (V descriptor) path
        k = torch.zeros_like(k)
    elif kind == "qk":     # V = one-hot over d: O[:, d] = sum_{j = d mod D} P[:, j]: exposes P (hence S = QK^T)
        vm = torch.zeros(B, Skv, H, D)
        for j in range(Skv):
            vm[:, j, :, j % D] = 1.0
        v = vm.to(dtype).cuda().view(B, Skv, H, D).transpose(1, 2)
    return q, k, v


def _attn_case(B, H, Sq, Skv, D, dtype, kind="random", tol=None, seed=0):
    import torch
    from diffsim_b200 import ops
    from oracle import aas_oracle as O

    q, k, v = _attn_inputs(B, H, Sq, Skv, D, dtype, seed, kind)
    out = ops.attn_fwd(q, k, v)
    torch.cuda.synchronize()
    ref = O.attention(q.cpu(), k.cpu(), v.cpu())
    tol = tol if tol is not None else (2e-2 if dtype == torch.bfloat16 else 4e-3)
    return report(f"attn {kind} B{B} H{H} Sq{Sq} Skv{Skv} D{D} {str(dtype)[6:]}", out, ref, tol, 32, 16)


def case_attn_pv():
    import torch

    ok = _attn_case(1, 1, 128, 64, 64, torch.float16, "pv")
    ok &= _attn_case(1, 1, 128, 256, 64, torch.float16, "pv")
    ok &= _attn_case(1, 2, 128, 256, 160, torch.float16, "pv")
    ok &= _attn_case(2, 8, 256, 256, 160, torch.bfloat16, "pv")
    return ok


def case_attn_qk():
    import torch

    ok = _attn_case(1, 1, 128, 64, 64, torch.float16, "qk")
    ok &= _attn_case(1, 1, 128, 256, 64, torch.float16, "qk")
    ok &= _attn_case(1, 2, 128, 256, 160, torch.float16, "qk")
    return ok


def case_attn_d160():
    import torch

    ok = _attn_case(2, 8, 256, 256, 160, torch.float16)
    ok &= _attn_case(2, 8, 256, 256, 160, torch.bfloat16)
    ok &= _attn_case(2, 8, 64, 64, 160, torch.float16)   # SD-1.5 mid block
    return ok


def case_attn_d64():
    import torch

    ok = _attn_case(1, 4, 256, 256, 64, torch.float16)
    ok &= _attn_case(2, 20, 256, 256, 64, torch.bfloat16)
    return ok


def case_attn_d72():
    import torch
    from diffsim_b200 import ops, synth
    from oracle import aas_oracle as O

    ok = _attn_case(2, 16, 256, 256, 72, torch.float16)
    ok &= _attn_case(2, 16, 256, 256, 72, torch.bfloat16)
    # DiT packed-qkv strides (diffsim/diffsim_dit.py:22-23)
    m = synth.SynthModel(2, 16, 256, 72, seed=5)
    q, k, v = m.image(m.new_base(), 0.9, torch.float16, layout="dit")
    dev_packed = torch.empty(2, 256, 3, 16, 72, dtype=torch.float16, device="cuda")
    dev_packed[:, :, 0] = q.permute(0, 2, 1, 3).cuda()
    dev_packed[:, :, 1] = k.permute(0, 2, 1, 3).cuda()
    dev_packed[:, :, 2] = v.permute(0, 2, 1, 3).cuda()
    qkv = dev_packed.permute(2, 0, 3, 1, 4)
    out = ops.attn_fwd(qkv[0], qkv[1], qkv[2])
    ok &= report("attn dit-packed D72", out, O.attention(q, k, v), 4e-3, 32, 16)
    return ok


def case_attn_d40():
    import torch

    return _attn_case(2, 8, 256, 256, 40, torch.float16)


def case_attn_d80():
    import torch

    return _attn_case(2, 8, 256, 256, 80, torch.float16)


def case_attn_d128():
    import torch

    return _attn_case(1, 4, 256, 256, 128, torch.bfloat16)


def case_attn_tails():
    import torch

    ok = _attn_case(1, 12, 50, 50, 64, torch.float16)      # CLIP ViT-B/32: 50 tokens
    ok &= _attn_case(1, 6, 257, 200, 64, torch.float16)    # ragged q and kv
    ok &= _attn_case(2, 3, 100, 130, 160, torch.bfloat16)
    ok &= _attn_case(1, 2, 1, 1, 64, torch.float16)
    return ok


def case_attn_long():
    import torch

    # kv lengths beyond one 256-row group: online softmax across groups (SDXL / SD-1.5 high-resolution layers)
    ok = _attn_case(1, 2, 128, 320, 64, torch.float16)       # 1 full group + a 64-row tail (half A only)
    ok &= _attn_case(1, 2, 200, 448, 64, torch.float16)      # tail group with a partial half B
    ok &= _attn_case(1, 2, 256, 512, 64, torch.bfloat16)
    ok &= _attn_case(2, 4, 1024, 1024, 64, torch.float16)    # SDXL up_blocks[0]-like
    ok &= _attn_case(1, 2, 1024, 1024, 80, torch.float16)    # SD-1.5 up_blocks[1]-like
    ok &= _attn_case(1, 2, 512, 4096, 40, torch.float16)     # SD-1.5 up_blocks[2]-like
    ok &= _attn_case(1, 2, 4096, 4096, 64, torch.bfloat16)   # SDXL up_blocks[1]-like
    ok &= _attn_case(1, 1, 300, 700, 160, torch.float16)
    ok &= _attn_case(1, 2, 256, 256, 64, torch.float16, "ramp")
    ok &= _attn_case(1, 2, 256, 1024, 64, torch.float16, "ramp")
    ok &= _attn_case(1, 2, 256, 256, 160, torch.float16, "ramp")
    ok &= _attn_case(1, 2, 128, 2048, 64, torch.bfloat16, "ramp")
    return ok


def _pairs_setup(shape, n_pairs, dtype, seed=0, layout="sd"):
    from diffsim_b200 import synth

    B, H, S, D = shape
    m = synth.SynthModel(B, H, S, D, seed=2334)
    images, pairs = synth.make_pairs(m, n_pairs, dtype, seed=seed, layout=layout)
    return images, pairs


def case_aas_pairs():
    import torch
    from diffsim_b200 import ops, synth
    from oracle import aas_oracle as O

    ok = True
    for shape, dtype, npairs in (((2, 8, 256, 160), torch.float16, 4), ((2, 8, 256, 160), torch.bfloat16, 3),
                                 ((2, 16, 256, 72), torch.float16, 3), ((2, 4, 64, 64), torch.float16, 3),
                                 ((2, 4, 640, 64), torch.float16, 2), ((1, 2, 1024, 64), torch.bfloat16, 2)):
        images, pairs = _pairs_setup(shape, npairs, dtype)
        q, k, v = synth.stack_cache(images, "cuda")
        for mode in ("cosine", "mse"):
            got = ops.aas_pairs(q, k, v, pairs, mode).cpu()
            ref = torch.tensor([O.aas_pair_score(*images[a], *images[b], mode=mode) for a, b in pairs])
            t0 = torch.tensor([O.aas_pair_score(*images[a], *images[b], mode=mode, tier="T0") for a, b in pairs])
            rel = ((got.double() - ref).abs() / ref.abs().clamp_min(1e-9)).max().item()
            rel0 = ((got.double() - t0).abs() / t0.abs().clamp_min(1e-9)).max().item()
            good = rel < 1e-3
            ok &= good
            print(f"[aas_pairs {shape} {str(dtype)[6:]} {mode}] rel_err_vs_T1={rel:.2e} vs_T0={rel0:.2e} {'OK' if good else 'FAIL'}")
            print("    got", [round(x, 6) for x in got.tolist()])
            print("    ref", [round(x, 6) for x in ref.tolist()])
    return ok


def case_aas_groups():
    import torch
    from diffsim_b200 import ops, synth
    from oracle import aas_oracle as O

    shape, dtype = (2, 8, 256, 160), torch.float16
    m = synth.SynthModel(*shape, seed=2334)
    images, trips = synth.make_triplets(m, 3, dtype, seed=3)
    q, k, v = synth.stack_cache(images, "cuda")
    # groups by unique query image: ref -> [left, right], left -> [ref], right -> [ref]
    gq, go, kv = [], [0], []
    for (r, l, rt) in trips:
        for qi, lst in ((r, [l, rt]), (l, [r]), (rt, [r])):
            gq.append(qi)
            kv.extend(lst)
            go.append(len(kv))
    got = ops.aas_groups(q, k, v, k, v, gq, go, kv, "cosine").cpu()
    ref = []
    for gi, qi in enumerate(gq):
        for t in range(go[gi], go[gi + 1]):
            j = kv[t]
            ref.append(O.aas_directional(images[qi][0], images[qi][1], images[qi][2], images[j][1], images[j][2]))
    ref = torch.tensor(ref)
    rel = ((got.double() - ref).abs() / ref.abs().clamp_min(1e-9)).max().item()
    print(f"[aas_groups triplets] rel_err={rel:.2e} {'OK' if rel < 1e-3 else 'FAIL'}")
    print("    got", [round(x, 6) for x in got.tolist()])
    print("    ref", [round(x, 6) for x in ref.tolist()])
    return rel < 1e-3


def case_aas_matrix():
    import torch
    from diffsim_b200 import ops, synth
    from oracle import aas_oracle as O

    shape, dtype = (2, 4, 128, 64), torch.float16
    m = synth.SynthModel(*shape, seed=2334)
    images, labels = synth.make_styles(m, 5, 4, dtype, seed=4)
    q, k, v = synth.stack_cache(images, "cuda")
    got = ops.aas_matrix(q, k, v, k, v, "cosine").cpu()
    ref = O.aas_matrix([im[0] for im in images], [im[1] for im in images], [im[2] for im in images])
    ok = report("aas_matrix 20x20", got, ref, 1e-3, 4, 4)
    # row-block sharding must be bit-identical to the full matrix
    half = ops.aas_matrix(q[10:], k[10:], v[10:], k, v, "cosine").cpu()
    same = bool((half == got[10:]).all())
    print(f"[aas_matrix row-block == full rows, bitwise] {'OK' if same else 'FAIL'}")
    return ok and same


def case_perf():
    import torch
    from diffsim_b200 import ops, synth

    dev = "cuda"
    res = {}
    # K2
    for dtype in (torch.float16,):
        P, E = 256, 655360
        x = torch.randn(P, E, device=dev).to(dtype)
        y = torch.randn(P, E, device=dev).to(dtype)
        for mode in ("cosine", "mse", "minmax_cosine"):
            ms = timeit(lambda: ops.pair_reduce(x, y, mode))
            gbs = 2 * P * E * x.element_size() / ms / 1e6
            res[f"reduce_{mode}_GBs"] = gbs
            print(f"[perf reduce {mode}] {ms:.3f} ms  {gbs:.0f} GB/s")
    # K1 pairs, SD-1.5 up0
    B, H, S, D = 2, 8, 256, 160
    n_img = 768
    q, k, v = synth.device_cache(B, H, S, D, n_img, torch.float16, dev)
    pairs = torch.tensor([(3 * t, 3 * t + 1) for t in range(n_img // 3)] + [(3 * t, 3 * t + 2) for t in range(n_img // 3)],
                         dtype=torch.int32, device=dev)
    ms = timeit(lambda: ops.aas_pairs(q, k, v, pairs, "cosine"), iters=10)
    P = pairs.shape[0]
    from diffsim_b200 import _native as N
    cyc = torch.zeros(8 * 4 + 1, dtype=torch.int64, device=dev)
    N.load().ds_debug_set_trace(cyc.data_ptr(), 4)
    ops.aas_pairs(q, k, v, pairs, "cosine")
    torch.cuda.synchronize()
    N.load().ds_debug_set_trace(None, 0)
    items_per_sm = 4 * P * B * H * (S // 128) / 148
    print(f"[perf aas_pairs cycles] CTA0 {int(cyc[-1])} SM clocks, {int(cyc[-1]) / items_per_sm:.0f} clk per item (tensor floor 2560)")
    fl = 4 * P * 4 * B * H * S * S * D
    print(f"[perf aas_pairs sd15_up0 fp16] {P} pairs {ms:.3f} ms  {P / ms * 1e3:.0f} pairs/s  {fl / ms / 1e9:.1f} TFLOP/s")
    res["aas_pairs_tflops"] = fl / ms / 1e9
    # torch SDPA reference on the same device, same tensors (per pair: 4 SDPA + 2 cosine)
    import torch.nn.functional as F

    def torch_ref(np_=64):
        out = []
        for i in range(np_):
            a, b = int(pairs[i, 0]), int(pairs[i, 1])
            qa, ka, va, qb, kb, vb = q[a], k[a], v[a], q[b], k[b], v[b]
            ab = F.scaled_dot_product_attention(qa, kb, vb)
            ba = F.scaled_dot_product_attention(qb, ka, va)
            sa = F.scaled_dot_product_attention(qa, ka, va)
            sb = F.scaled_dot_product_attention(qb, kb, vb)
            out.append((F.cosine_similarity(ab.reshape(1, -1), sa.reshape(1, -1)) +
                        F.cosine_similarity(ba.reshape(1, -1), sb.reshape(1, -1))) / 2)
        return out

    pl = pairs.cpu()
    ms_t = timeit(lambda: torch_ref(64), iters=3, warmup=1)
    print(f"[perf torch-CUDA reference lines] 64 pairs {ms_t:.3f} ms  {64 / ms_t * 1e3:.0f} pairs/s")
    res["torch_cuda_pairs_per_s"] = 64 / ms_t * 1e3
    # batched torch SDPA only (upper bound for the library path)
    qa = q[pairs[:128, 0].long()].reshape(-1, H, S, D)
    kb = k[pairs[:128, 1].long()].reshape(-1, H, S, D)
    vb = v[pairs[:128, 1].long()].reshape(-1, H, S, D)
    ms_b = timeit(lambda: F.scaled_dot_product_attention(qa, kb, vb), iters=10)
    print(f"[perf torch SDPA batched 128 attentions] {ms_b:.3f} ms {128 * 4 * B * H * S * S * D / ms_b / 1e9:.1f} TFLOP/s")
    print(json.dumps(res))
    return True


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="all")
    ap.add_argument("--timeout", type=int, default=240)
    args = ap.parse_args()
    if args.case == "all" or "," in args.case:
        cases = CASES if args.case == "all" else args.case.split(",")
        summary = {}
        for c in cases:
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", c], timeout=args.timeout)
                summary[c] = "OK" if r.returncode == 0 else f"FAIL(rc={r.returncode})"
            except subprocess.TimeoutExpired:
                summary[c] = "TIMEOUT"
            print(f"=== case {c}: {summary[c]} ({time.time() - t0:.1f}s)", flush=True)
        print("SUMMARY " + json.dumps(summary))
        sys.exit(0 if all(v == "OK" for v in summary.values()) else 1)
    fn = globals()["case_" + args.case]
    ok = fn()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
