"""Per-image Q/K/V store and retrieval result files.

* `save_qkv` / `load_qkv`: what `DiffSim.diffsim_value` returns per image (diffsim/diffsim.py:201-258) kept as three
  (N,B,S,H*D) tensors in one safetensors file, keyed by (model, block, layer, step, seed) in the metadata, so that an
  all-pairs run extracts every image once.
* `write_retrieval_results`: one text file per query image, one retrieved image per line, best first, first token
  "<cls>_<imgid>:" -- the format retrieval_vis.py parses (:57-68 Sref/class-folder layout `<out>/<cls>/<imgid>.txt`,
  :121-132 flat COCO layout `<out>/<imgid>.txt` with first token "<imgid>:").  The reference ships only the parser
  (and the plots made from it); the N x N scoring that produces such files is ds_aas_matrix / ds_simmat.
* `read_retrieval_result`: the parser's logic (first tokens of the first `limit` lines), for tests and consumers;
  `skip_first_id` reproduces the IP variant that drops "<cls>_1:" (:197).
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from .scoring import QKVCache


# --------------------------------------------------------------------------------------------------------
# Q/K/V store
# --------------------------------------------------------------------------------------------------------
def store_key(model: str, target_block: str, target_layer, target_step: int, seed, img_size: int) -> str:
    layer = ",".join(str(x) for x in target_layer) if isinstance(target_layer, (list, tuple)) else str(target_layer)
    return f"{model}|{target_block}|{layer}|t{int(target_step)}|seed{seed}|{int(img_size)}"


def save_qkv(path: str, cache: QKVCache, names: Sequence[str], key: str = "") -> None:
    """Write the cache (memory order (N,B,S,H*D), the hook's layout) and the image names."""
    from safetensors.torch import save_file

    if len(names) != cache.n_images:
        raise ValueError(f"{len(names)} names for {cache.n_images} images")
    qm, km, vm = cache.memory()
    B, H, S, D = cache.shape
    meta = {"names": json.dumps(list(names)), "shape_BHSD": json.dumps([B, H, S, D]), "key": key, "format": "diffsim_b200.qkv.v1"}
    save_file({"q": qm.cpu().contiguous(), "k": km.cpu().contiguous(), "v": vm.cpu().contiguous()}, path, metadata=meta)


def load_qkv(path: str, device="cpu", expect_key: Optional[str] = None) -> Tuple[QKVCache, List[str]]:
    from safetensors import safe_open

    with safe_open(path, framework="pt", device="cpu") as f:
        meta = f.metadata() or {}
        if meta.get("format") != "diffsim_b200.qkv.v1":
            raise ValueError(f"{path}: not a diffsim_b200 Q/K/V store")
        if expect_key is not None and meta.get("key") != expect_key:
            raise ValueError(f"{path}: stored for '{meta.get('key')}', wanted '{expect_key}' (different model / layer / step)")
        B, H, S, D = json.loads(meta["shape_BHSD"])
        names = json.loads(meta["names"])
        mems = [f.get_tensor(n) for n in ("q", "k", "v")]
    view = lambda m: m.to(device).view(m.shape[0], B, S, H, D).permute(0, 1, 3, 2, 4)  # noqa: E731
    return QKVCache(view(mems[0]), view(mems[1]), view(mems[2])), names


# --------------------------------------------------------------------------------------------------------
# retrieval result files
# --------------------------------------------------------------------------------------------------------
def ranked_indices(score: torch.Tensor, topk: Optional[int] = None, larger_is_closer: bool = True,
                   skip_self: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per query row: the indices of the other images ordered best first, and their scores.  Ties break towards
    the lower index (stable sort), so files are reproducible."""
    s = score.detach().float().cpu().clone()
    n_rows, n_cols = s.shape
    if skip_self:
        if n_rows != n_cols:
            raise ValueError("skip_self needs a square matrix")
        s.fill_diagonal_(float("-inf") if larger_is_closer else float("inf"))
    order = torch.sort(s, dim=1, descending=larger_is_closer, stable=True).indices
    k = n_cols - (1 if skip_self else 0)
    k = k if topk is None else min(k, int(topk))
    order = order[:, :k]
    return order, torch.gather(s, 1, order)


def write_retrieval_results(score: torch.Tensor, names: Sequence[str], out_dir: str, topk: Optional[int] = 10,
                            larger_is_closer: bool = True, layout: str = "class_folders") -> List[str]:
    """names: "<cls>_<imgid>" per image (layout 'class_folders', retrieval_vis.py:57-68) or "<imgid>" (layout 'flat',
    :121-132).  Writes `<out_dir>/<cls>/<imgid>.txt` or `<out_dir>/<imgid>.txt`; returns the paths."""
    if layout not in ("class_folders", "flat"):
        raise ValueError(layout)
    n = len(names)
    if score.shape != (n, n):
        raise ValueError(f"score matrix {tuple(score.shape)} does not match {n} names")
    order, vals = ranked_indices(score, topk, larger_is_closer)
    paths = []
    for i, name in enumerate(names):
        if layout == "class_folders":
            cls, _, img_id = name.partition("_")
            if not img_id or "_" in img_id:
                raise ValueError(f"'{name}': class_folders layout needs names of the form <cls>_<imgid> (one underscore; "
                                 "the reference parser splits on it)")
            d = os.path.join(out_dir, cls)
            path = os.path.join(d, f"{img_id}.txt")
        else:
            d, path = out_dir, os.path.join(out_dir, f"{name}.txt")
        os.makedirs(d, exist_ok=True)
        with open(path, "w") as f:
            for j, v in zip(order[i].tolist(), vals[i].tolist()):
                f.write(f"{names[j]}: {v:.6f}\n")
        paths.append(path)
    return paths


def read_retrieval_result(path: str, limit: int = 4, layout: str = "class_folders", skip_first_id: bool = False) -> List[str]:
    """What retrieval_vis.py's read_image_path extracts from a result file: the relative image paths of the first
    `limit` lines ('<cls>/<imgid>.png' or '<imgid>.jpg')."""
    out = []
    with open(path) as f:
        for line in f:
            parts = line.strip().split()
            if len(parts) < 1:
                continue
            if layout == "class_folders":
                cls, img_id = parts[0].split("_")
                if skip_first_id and img_id == "1:":
                    continue
                out.append(os.path.join(cls, f"{img_id[:-1]}.png"))
            else:
                out.append(f"{parts[0][:-1]}.jpg")
            if len(out) >= limit:
                break
    return out


def retrieval_accuracy(score: torch.Tensor, labels: Sequence, topk: int = 1, larger_is_closer: bool = True) -> Dict[str, float]:
    """Top-k hit rate and mean precision@k of same-label retrieval (Sref: 508 styles x 4 images)."""
    order, _ = ranked_indices(score, topk, larger_is_closer)
    lab = {l: i for i, l in enumerate(dict.fromkeys(labels))}
    y = torch.tensor([lab[l] for l in labels])
    hit = (y[order] == y[:, None])
    return {"hit@k": hit.any(dim=1).float().mean().item(), "precision@k": hit.float().mean().item(), "k": int(order.shape[1])}
