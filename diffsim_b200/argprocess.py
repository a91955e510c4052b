"""Command-line flags of the benchmark drivers -- same names, choices and defaults as the reference's
argprocess.py:3-19 (including `--similarity` defaulting to mse although every launcher script passes cosine, and
the parsed-but-unused --use_mask / --use_text_attn / --out_path)."""
import argparse

METRICS = ["diffsim", "diffsim_xl", "clip_i", "clip_cross", "dino", "dinov1", "dino_cross", "cute", "lpips", "gram",
           "diffeats", "clipfeats", "dinofeats", "ensemble", "dit"]

# per-benchmark settings of the reference's launcher scripts (cute_main.sh:3, night_main.sh:3, style_main.sh:4,7,
# ipref_main.sh:4, tid_main.sh:3, dreambench_main.sh:3); all use --similarity cosine --seed 2334
BENCHMARK_PRESETS = {
    "cute": dict(target_block="up_blocks", target_layer=[0], target_step=600),
    "nights": dict(target_block="up_blocks", target_layer=[0], target_step=500),
    "sref": dict(target_block="up_blocks", target_layer=[0], target_step=900),
    "ipref": dict(target_block="up_blocks", target_layer=[5], target_step=750),
    "tid": dict(target_block="up_blocks", target_layer=[0], target_step=900),
    "dreambench": dict(target_block="up_blocks", target_layer=[0], target_step=750),
}


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="DiffSim scoring flags (reference-compatible).")
    p.add_argument("--image_path", type=str, help="Path to image folder")
    p.add_argument("--original_path", type=str, default=None, help="Path to original images for ipref")
    p.add_argument("--out_path", type=str, help="Output folder (ckpt folder / retrieval result folder)")
    p.add_argument("--image_size", type=int, default=512, help="(Resized) resolution of the compared images")
    p.add_argument("--target_block", type=str, choices=["down_blocks", "mid_blocks", "up_blocks"], default="up_blocks")
    p.add_argument("--target_layer", type=int, default=2, nargs="+",
                   help="Target layer; SDXL takes 3 numbers: block, attention and transformer-block index")
    p.add_argument("--target_step", type=int, default=100, help="Index of the denoising step")
    p.add_argument("--metric", type=str, choices=METRICS, default="diffsim")
    p.add_argument("--similarity", type=str, choices=["cosine", "mse"], default="mse")
    p.add_argument("--prompt", type=str, default="High quality image")
    p.add_argument("--ip_adapter", action="store_true")
    p.add_argument("--use_mask", action="store_true")
    p.add_argument("--use_text_attn", action="store_true")
    p.add_argument("--seed", type=int, default=2333)
    return p


def arg_parse(argv=None):
    return build_parser().parse_args(argv)
