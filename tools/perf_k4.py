import os, sys, torch
sys.path.insert(0, ".")
from diffsim_b200 import ops, synth, _native as N
lib = N.load()
def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
print(os.path.basename(os.environ.get("DIFFSIM_B200_LIB", "default")))
for n_img in (96, 768):
    hid, w = synth.device_hidden(2, 8, 256, 160, n_img, torch.float16, "cuda")
    outs = [torch.empty(n_img, 2, 256, 1280, dtype=torch.float16, device="cuda") for _ in range(3)]
    fl = 2 * n_img * 512 * 1280 * 3840
    line = f"[K4 {n_img} images]"
    for var in (0, 2, 0, 2):
        lib.ds_debug_set_gemm_variant(var)
        ms = timeit(lambda: ops.qkv_project(hid, w, None, 3, out=outs))
        line += f" variant {var}: {ms:.3f} ms {fl / ms / 1e9:.0f} TFLOP/s |"
    print(line, flush=True)
