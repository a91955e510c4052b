import torch, sys
sys.path.insert(0, ".")
from diffsim_b200 import ops, _native as N
lib = N.load()
for n, L in ((2032, 655360), (2032, 163840)):
    f = torch.empty(n, L, dtype=torch.float16, device="cuda")
    for i in range(0, n, 127):
        f[i:i + 127] = torch.randn(min(127, n - i), L, device="cuda").half()
    g = f.clone()
    for pf in (256, 128, 64, 32, 256):
        lib.ds_debug_set_simmat_max_kb(pf)
        out = []
        for name, fn in (("self", lambda: ops.simmat(f, None, "cosine")), ("full", lambda: ops.simmat(f, g, "cosine"))):
            for _ in range(3): fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): fn()
            e1.record(); torch.cuda.synchronize()
            out.append(f"{name} {e0.elapsed_time(e1) / 10:.3f} ms")
        d = float((ops.simmat(f, None, "cosine").diagonal() - 1).abs().max())
        print(f"N={n} L={L} max_kb_per_partial={pf}: " + ", ".join(out) + f", diag err {d:.1e}", flush=True)
    del f, g
