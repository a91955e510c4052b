"""Batched equivalents of the reference's benchmark drivers.

The reference walks a dataset, calls `DiffSim.diffsim` twice per triplet (extracting the reference image twice,
cute_main.py:111-132) and synchronises with the device on every comparison (cute_main.py:196-205,
night_main.py:157-163).  Here a driver
  1. extracts every DISTINCT image once through the scorer's trunk (the per-image capture of diffsim/diffsim.py:122-169),
  2. stacks the (q,k,v) into a cache,
  3. scores all triplets with ONE library call (ds_aas_triplets: 7 attentions per triplet instead of 8, decisions and
     counts on the device) and reads the counts back once.
Dataset walking (folder layouts, csv columns) stays with the caller: the functions take triplets of image
identifiers -- paths for a real trunk, ids for the synthetic one.

    run_2afc     cute_main.py / style_main.py / ipref_main.py / tid_main.py / dreambench_main.py rule:
                 cosine: correct if diff_ab > diff_ac (2x: diff_ab > 2 diff_ac); mse: `<` (cute_main.py:196-205)
    run_nights   night_main.py:157-163: predicted = 1 if ab > ac (cosine) / ab < ac (mse); compared with the
                 annotators' `left_vote` column as the reference does
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Hashable, List, Optional, Sequence, Tuple

import torch

from . import ops
from .scoring import QKVCache


@dataclass
class TwoAFCResult:
    total: int
    correct: int
    correct_2x: int
    diff_ab: torch.Tensor   # float32 [T] on the device
    diff_ac: torch.Tensor
    flags: torch.Tensor     # uint8 [T]: the per-triplet decision

    @property
    def accuracy(self) -> float:
        return 100.0 * self.correct / self.total if self.total else 0.0   # the reference prints percentages

    @property
    def accuracy_2x(self) -> float:
        return 100.0 * self.correct_2x / self.total if self.total else 0.0


def build_cache(scorer, images: Sequence[Hashable], img_size, prompt, target_block, target_layer, target_step,
                seed="2333", device="cuda") -> Tuple[QKVCache, Dict[Hashable, int]]:
    """Extract each distinct image once (order of first appearance) and stack the results."""
    index: Dict[Hashable, int] = {}
    qkvs = []
    for im in images:
        if im in index:
            continue
        index[im] = len(qkvs)
        # extract() = the layer diffsim() scores (the reference's diffsim_value() indexes the blocks differently)
        fn = getattr(scorer, "extract", None) or scorer.diffsim_value
        qkvs.append(fn(im, img_size, prompt, target_block, target_layer, target_step, seed=seed, device=device))
    return QKVCache.from_images(qkvs, device), index


def _score(cache: QKVCache, index, triplets, similarity, round_scores):
    trip = torch.tensor([[index[a], index[b], index[c]] for a, b, c in triplets], dtype=torch.int32)
    return ops.aas_triplets(cache.q, cache.k, cache.v, trip, similarity, None, round_scores)


def run_2afc(scorer, triplets: Sequence[Tuple[Hashable, Hashable, Hashable]], img_size=512, prompt="",
             target_block="up_blocks", target_layer=(0,), target_step=600, similarity="cosine", seed="2333",
             device="cuda", round_scores: bool = True, cache: Optional[Tuple[QKVCache, Dict]] = None) -> TwoAFCResult:
    """triplets: (image_A, image_B, image_C) with B the positive -- the loop body of cute_main.py:108-205 for all
    triplets at once.  round_scores=True compares scores rounded to the input dtype, as the reference's fp16 score
    tensors are (diffsim/diffsim.py:197)."""
    if cache is None:
        cache = build_cache(scorer, [im for t in triplets for im in t], img_size, prompt, target_block, target_layer,
                            target_step, seed, device)
    c, index = cache
    ab, ac, counts, flags = _score(c, index, triplets, similarity, round_scores)
    n = counts.cpu()          # the one host sync of the run
    return TwoAFCResult(len(triplets), int(n[0]), int(n[1]), ab, ac, flags)


def run_nights(scorer, rows: Sequence[Tuple[Hashable, Hashable, Hashable, int]], img_size=512, prompt="",
               target_block="up_blocks", target_layer=(0,), target_step=600, similarity="cosine", seed="2333",
               device="cuda", round_scores: bool = True) -> TwoAFCResult:
    """rows: (ref, left, right, vote) with `vote` the csv column the reference compares `predicted` with
    (night_main.py:66,157-163).  `correct` counts predicted == vote; correct_2x is not defined for NIGHTS (0)."""
    trips = [(r, l, rt) for r, l, rt, _ in rows]
    cache, index = build_cache(scorer, [im for t in trips for im in t], img_size, prompt, target_block, target_layer,
                               target_step, seed, device)
    ab, ac, _, flags = _score(cache, index, trips, similarity, round_scores)
    votes = torch.tensor([int(v) for *_, v in rows], dtype=torch.uint8, device=flags.device)
    correct = int((flags == votes).sum().cpu())
    return TwoAFCResult(len(rows), correct, 0, ab, ac, flags)


def format_report(name: str, r: TwoAFCResult) -> List[str]:
    """The lines the reference drivers print at the end of a run."""
    return [f"Current total samples: {r.total}", f"{name} accuracy: {r.accuracy:.2f}%", f"{name} 2x accuracy: {r.accuracy_2x:.2f}%"]


def ensemble_votes(metric_scores: Sequence[Tuple[torch.Tensor, torch.Tensor]], votes: Optional[torch.Tensor] = None) -> int:
    """The reference's 'ensemble' metric: per triplet each member metric votes `0 if ab < ac else 1` -- ties count FOR the
    positive, unlike the strict single-metric rule -- and the triplet is correct when at least two of the three agree
    (cute_main.py:187-195); on NIGHTS the majority is compared with the annotators' vote: correct when (vote == 1 and
    sum >= 2) or (vote == 0 and sum <= 1) (night_main.py:148-152).  metric_scores: [(ab, ac)] per member metric, each a
    [T] tensor.  Returns the number of correct triplets (one device sync)."""
    if len(metric_scores) != 3:
        raise ValueError("the reference's ensemble has exactly three members (diffsim, clip_i, dino)")
    total = None
    for ab, ac in metric_scores:
        corr = (~(ab < ac)).to(torch.int32)
        total = corr if total is None else total + corr.to(total.device)
    if votes is None:
        return int((total >= 2).sum())
    votes = votes.to(total.device)
    return int((((votes == 1) & (total >= 2)) | ((votes == 0) & (total <= 1))).sum())
