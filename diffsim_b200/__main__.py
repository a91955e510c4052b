"""`python -m diffsim_b200 <benchmark> [reference flags]` -- the driver layer end to end on synthetic data.

The reference's `*_main.py` scripts walk a dataset on disk and a diffusion pipeline with downloaded weights; neither
exists offline, so this entry point keeps their command line (argprocess.py flags, per-benchmark presets of the launcher
scripts) and their printed report, and replaces the dataset by a synthetic one of the same SHAPE served by SyntheticTrunk:

    python -m diffsim_b200 cute   --similarity cosine --n 450        # CUTE-shaped 2AFC triplets, cute_main.py rule
    python -m diffsim_b200 nights --similarity cosine --n 256        # NIGHTS: predicted vs the annotators' vote
    python -m diffsim_b200 sref   --n 32 --out_path /tmp/sref        # style retrieval: N x N matrix -> result files

With a real pipeline, construct `DiffSim(trunk=DiffusersTrunk(pipe))` and call `drivers.run_2afc / run_nights` or
`scoring.aas_matrix_*` + `retrieval.write_retrieval_results` with the dataset's image paths instead.
"""
from __future__ import annotations

import argparse
import sys

from . import argprocess


def build_parser() -> argparse.ArgumentParser:
    p = argprocess.build_parser()
    p.prog = "python -m diffsim_b200"
    p.add_argument("benchmark", choices=sorted(argprocess.BENCHMARK_PRESETS), help="which driver to run")
    p.add_argument("--n", type=int, default=128, help="synthetic triplets (2AFC benchmarks) or styles (sref)")
    p.add_argument("--device", default="cuda")
    return p


def resolve(args) -> dict:
    """Flags -> scoring settings: explicit flags win, else the launcher script's preset for the benchmark."""
    preset = argprocess.BENCHMARK_PRESETS[args.benchmark]
    defaults = argprocess.build_parser()
    out = {}
    for key in ("target_block", "target_layer", "target_step"):
        val = getattr(args, key)
        out[key] = preset[key] if val == defaults.get_default(key) else val
    if isinstance(out["target_layer"], int):
        out["target_layer"] = [out["target_layer"]]
    return out


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    cfg = resolve(args)
    import torch

    from . import drivers, retrieval, scoring
    from .diffsim import DiffSim, SyntheticTrunk

    dtype = torch.float16
    ds = DiffSim(dtype, args.device, trunk=SyntheticTrunk((2, 8, 256, 160), dtype, args.device, seed=args.seed))
    common = dict(img_size=args.image_size, prompt=args.prompt, target_block=cfg["target_block"],
                  target_layer=cfg["target_layer"], target_step=cfg["target_step"], similarity=args.similarity,
                  seed=str(args.seed), device=args.device)
    g = torch.Generator().manual_seed(args.seed)
    if args.benchmark == "sref":
        names, images = [], []
        for s in range(args.n):
            for i in range(1, 5):
                names.append(f"{s:03d}_{i}")
                images.append(f"style{s}@0.8")       # four noisy renderings of one style concept
        # SyntheticTrunk keys images by (concept, alpha, step): give each rendering its own step offset so they differ
        qkvs = [ds.trunk.extract(im, target_step=cfg["target_step"] + j % 4) for j, im in enumerate(images)]
        cache = scoring.QKVCache.from_images(qkvs, args.device)
        score = scoring.symmetrize(scoring.aas_matrix_local(cache, cache, args.similarity))
        acc = retrieval.retrieval_accuracy(score, [n.split("_")[0] for n in names], topk=3,
                                           larger_is_closer=args.similarity == "cosine")
        print(f"Current total samples: {len(names)}")
        print(f"Sref retrieval precision@3: {100 * acc['precision@k']:.2f}%  hit@3: {100 * acc['hit@k']:.2f}%")
        if args.out_path:
            paths = retrieval.write_retrieval_results(score, names, args.out_path, topk=10,
                                                      larger_is_closer=args.similarity == "cosine")
            print(f"wrote {len(paths)} retrieval result files under {args.out_path}")
        return 0
    rows = []
    for t in range(args.n):
        a_pos = 0.55 + 0.4 * float(torch.rand((), generator=g))
        a_neg = 0.15 + 0.3 * float(torch.rand((), generator=g))
        ref, pos, neg = f"c{t}@1.0", f"c{t}@{a_pos:.3f}", f"c{t}@{a_neg:.3f}"
        if args.benchmark == "nights":
            left_is_pos = bool(torch.rand((), generator=g) < 0.5)
            rows.append((ref, pos, neg, 1) if left_is_pos else (ref, neg, pos, 0))
        else:
            rows.append((ref, pos, neg))
    if args.benchmark == "nights":
        r = drivers.run_nights(ds, rows, **common)
        print(f"Total samples now: {r.total}")
        print(f"Final validation accuracy: {r.accuracy:.2f}%")
    else:
        r = drivers.run_2afc(ds, rows, **common)
        for line in drivers.format_report(args.benchmark.upper(), r):
            print(line)
    return 0


if __name__ == "__main__":
    sys.exit(main())
