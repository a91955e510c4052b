"""Synthetic Q/K/V at the hook boundary (no weights, datasets or network exist offline).

Emulates "same model, different images": the projections Wq, Wk, Wv are shared by all images; an image's
hidden state is h = alpha * base + sqrt(1 - alpha^2) * noise, so images built from the same base with a large
alpha are near-duplicates and alpha -> 0 gives unrelated images.  gain_qk makes the logits peaky like real
attention.  iid-Gaussian Q/K/V are useless for parity: every score would be ~0.

Layout: q, k, v come out exactly as the reference's hook sees them -- (B,H,S,D) views over (B,S,H*D) memory,
strides (S*H*D, D, H*D, 1) (diffsim/hacked_attn.py:61-77), or DiT's packed-qkv strides
(diffsim/diffsim_dit.py:22-23).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

# (B, H, S, D) at the hooked layer -- SURVEY.md section 8
SHAPES = {
    "sd15_up0": (2, 8, 256, 160),      # SD-1.5 512^2, up_blocks layer 0 (configs 1-3)
    "sd15_mid": (2, 8, 64, 160),
    "sd15_up1": (2, 8, 1024, 80),
    "sd15_up2": (2, 8, 4096, 40),
    "sdxl_up0": (2, 20, 1024, 64),     # SDXL 1024^2 (config 4, real layers)
    "sdxl_up1": (2, 10, 4096, 64),
    "sdxl_literal": (2, 20, 4096, 64),  # config 4 as literally worded
    "dit_xl2": (2, 16, 256, 72),       # DiT-XL/2 256^2 (config 5)
}


@dataclass
class SynthModel:
    B: int
    H: int
    S: int
    D: int
    seed: int = 2334
    gain_qk: float = 2.2
    device: str = "cpu"

    def __post_init__(self):
        g = torch.Generator().manual_seed(self.seed)
        C = self.H * self.D
        std = 1.0 / math.sqrt(C)
        self.Wq = (torch.randn(C, C, generator=g) * std * self.gain_qk).to(self.device)
        self.Wk = (torch.randn(C, C, generator=g) * std * self.gain_qk).to(self.device)
        self.Wv = (torch.randn(C, C, generator=g) * std).to(self.device)
        self._g = g

    # ---- hidden states -------------------------------------------------------------------------
    def new_base(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """(B,S,C) hidden state of a "concept"; the second batch row (cond vs uncond half of CFG) is a
        small perturbation of the first."""
        g = generator or self._g
        C = self.H * self.D
        base = torch.randn(self.B, self.S, C, generator=g)
        for b in range(1, self.B):
            base[b] = base[0] + 0.1 * torch.randn(self.S, C, generator=g)
        return base

    def hidden(self, base: torch.Tensor, alpha: float, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        g = generator or self._g
        noise = torch.randn(base.shape, generator=g)
        return alpha * base + math.sqrt(max(0.0, 1.0 - alpha * alpha)) * noise

    # ---- projections ---------------------------------------------------------------------------
    def qkv(self, hidden: torch.Tensor, dtype=torch.float16, layout: str = "sd") -> Tuple[torch.Tensor, ...]:
        """hidden (B,S,C) -> q,k,v (B,H,S,D) views in the reference's memory layout."""
        h = hidden.to(self.Wq.device, torch.float32)
        B, S, C = h.shape
        q, k, v = h @ self.Wq, h @ self.Wk, h @ self.Wv
        if layout == "sd":
            # diffsim/hacked_attn.py:74-77: .view(B,-1,H,D).transpose(1,2)
            f = lambda t: t.to(dtype).contiguous().view(B, S, self.H, self.D).transpose(1, 2)  # noqa: E731
            return f(q), f(k), f(v)
        if layout == "dit":
            # diffsim/diffsim_dit.py:22-23: qkv(x).reshape(B,N,3,H,D).permute(2,0,3,1,4).unbind(0)
            packed = torch.stack([q, k, v], dim=2).to(dtype).contiguous()  # (B,S,3,C)
            qkv = packed.view(B, S, 3, self.H, self.D).permute(2, 0, 3, 1, 4)
            return qkv[0], qkv[1], qkv[2]
        raise ValueError(layout)

    def linear_weights(self, dtype=torch.float16) -> torch.Tensor:
        """[W_q; W_k; W_v] stacked along dim 0 in nn.Linear layout ([out, in], 3C x C): the to_q / to_k / to_v weights
        of the hooked layer (diffsim/hacked_attn.py:61-69) as the projection kernel takes them."""
        return torch.cat([self.Wq.t(), self.Wk.t(), self.Wv.t()], dim=0).to(dtype).contiguous()

    def image(self, base: torch.Tensor, alpha: float, dtype=torch.float16, layout: str = "sd",
              generator: Optional[torch.Generator] = None):
        return self.qkv(self.hidden(base, alpha, generator), dtype, layout)


def stack_cache(images: Sequence[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]], device=None):
    """[(q,k,v) per image, each a (B,H,S,D) view over (B,S,H*D)] -> three (N,B,H,S,D) cache views over
    (N,B,S,H*D) memory (the layout the library's Q/K/V caches use)."""
    outs = []
    for idx in range(3):
        ts = [im[idx] for im in images]
        B, H, S, D = ts[0].shape
        mem = torch.stack([t.permute(0, 2, 1, 3).reshape(B, S, H * D) for t in ts], dim=0)  # (N,B,S,H*D)
        if device is not None:
            mem = mem.to(device)
        outs.append(mem.view(len(ts), B, S, H, D).permute(0, 1, 3, 2, 4))
    return tuple(outs)


def make_pairs(model: SynthModel, n_pairs: int, dtype=torch.float16, alpha_lo: float = 0.3, alpha_hi: float = 0.99,
               seed: int = 0, layout: str = "sd"):
    """CUTE-shaped batch (config 1): n_pairs (A,B) pairs, B = noisy version of A's concept with alpha ~ U(lo,hi).
    Returns (images, pair_idx) with images a list of (q,k,v)."""
    g = torch.Generator().manual_seed(seed)
    images, pairs = [], []
    for p in range(n_pairs):
        base = model.new_base(g)
        alpha = alpha_lo + (alpha_hi - alpha_lo) * float(torch.rand((), generator=g))
        images.append(model.image(base, 1.0, dtype, layout, g))
        images.append(model.image(base, alpha, dtype, layout, g))
        pairs.append((2 * p, 2 * p + 1))
    return images, pairs


def make_triplets(model: SynthModel, n_triplets: int, dtype=torch.float16, seed: int = 0, layout: str = "sd",
                  near_tie_fraction: float = 0.1):
    """NIGHTS-shaped 2AFC batch (config 2): (ref, left, right) with alpha_left, alpha_right drawn so that most
    triplets have a clear margin and a tail of near-ties.  Returns (images, triplets) with images a list of
    (q,k,v) (3 per triplet) and triplets a list of (ref, left, right) image indices."""
    g = torch.Generator().manual_seed(seed)
    images, trips = [], []
    for t in range(n_triplets):
        base = model.new_base(g)
        a_left = 0.35 + 0.6 * float(torch.rand((), generator=g))
        if float(torch.rand((), generator=g)) < near_tie_fraction:
            a_right = min(0.99, max(0.05, a_left + 0.01 * (float(torch.rand((), generator=g)) - 0.5)))
        else:
            a_right = 0.35 + 0.6 * float(torch.rand((), generator=g))
        images.append(model.image(base, 1.0, dtype, layout, g))
        images.append(model.image(base, a_left, dtype, layout, g))
        images.append(model.image(base, a_right, dtype, layout, g))
        trips.append((3 * t, 3 * t + 1, 3 * t + 2))
    return images, trips


def make_styles(model: SynthModel, n_styles: int, per_style: int = 4, dtype=torch.float16, alpha: float = 0.8,
                seed: int = 0, layout: str = "sd"):
    """Sref-shaped retrieval set (config 3): n_styles x per_style images, images of one style share a base
    (alpha within style), so ground-truth neighbours exist.  Returns (images, style_of_image)."""
    g = torch.Generator().manual_seed(seed)
    images, labels = [], []
    for s in range(n_styles):
        base = model.new_base(g)
        for _ in range(per_style):
            images.append(model.image(base, alpha, dtype, layout, g))
            labels.append(s)
    return images, labels


def device_cache(B: int, H: int, S: int, D: int, n_images: int, dtype=torch.float16, device="cuda", seed: int = 0,
                 n_bases: int = 64, alpha_lo: float = 0.3, alpha_hi: float = 0.99):
    """Large synthetic Q/K/V cache generated ON THE DEVICE for throughput runs (the CPU generator above is too
    slow for tens of thousands of images).  Same recipe: shared projections, image = alpha * base + noise, with
    image i built from base (i // 3) % n_bases so that triplets (3t, 3t+1, 3t+2) share a concept.
    Returns q, k, v as (N,B,H,S,D) views over (N,B,S,H*D) memory."""
    g = torch.Generator(device=device).manual_seed(seed)
    C = H * D
    std = 1.0 / math.sqrt(C)
    Wq = torch.randn(C, C, generator=g, device=device) * (std * 2.2)
    Wk = torch.randn(C, C, generator=g, device=device) * (std * 2.2)
    Wv = torch.randn(C, C, generator=g, device=device) * std
    bases = torch.randn(n_bases, 1, S, C, generator=g, device=device).repeat(1, B, 1, 1)
    if B > 1:
        bases[:, 1:] += 0.1 * torch.randn(n_bases, B - 1, S, C, generator=g, device=device)
    mems = [torch.empty(n_images, B, S, C, dtype=dtype, device=device) for _ in range(3)]
    chunk = 64
    for i0 in range(0, n_images, chunk):
        i1 = min(n_images, i0 + chunk)
        idx = (torch.arange(i0, i1, device=device) // 3) % n_bases
        alpha = alpha_lo + (alpha_hi - alpha_lo) * torch.rand(i1 - i0, 1, 1, 1, generator=g, device=device)
        alpha[(torch.arange(i0, i1, device=device) % 3) == 0] = 1.0  # image 3t is the clean reference
        h = alpha * bases[idx] + torch.sqrt(1 - alpha * alpha) * torch.randn(i1 - i0, B, S, C, generator=g,
                                                                               device=device)
        mems[0][i0:i1] = (h @ Wq).to(dtype)
        mems[1][i0:i1] = (h @ Wk).to(dtype)
        mems[2][i0:i1] = (h @ Wv).to(dtype)
    return tuple(m.view(n_images, B, S, H, D).permute(0, 1, 3, 2, 4) for m in mems)


def device_hidden(B: int, H: int, S: int, D: int, n_images: int, dtype=torch.float16, device="cuda", seed: int = 0,
                  n_bases: int = 64, alpha_lo: float = 0.3, alpha_hi: float = 0.99, pin_host: bool = False):
    """Hook-INPUT form of device_cache: hidden states (N,B,S,C) of n_images synthetic images (same recipe) plus the
    stacked projection weight [W_q; W_k; W_v] (3C, C) in nn.Linear layout, both in `dtype`.  With pin_host the hidden
    states are returned in pinned host memory (generated on the device in chunks, copied back)."""
    g = torch.Generator(device=device).manual_seed(seed)
    C = H * D
    std = 1.0 / math.sqrt(C)
    Wq = torch.randn(C, C, generator=g, device=device) * (std * 2.2)
    Wk = torch.randn(C, C, generator=g, device=device) * (std * 2.2)
    Wv = torch.randn(C, C, generator=g, device=device) * std
    weight = torch.cat([Wq.t(), Wk.t(), Wv.t()], dim=0).to(dtype).contiguous()
    bases = torch.randn(n_bases, 1, S, C, generator=g, device=device).repeat(1, B, 1, 1)
    if B > 1:
        bases[:, 1:] += 0.1 * torch.randn(n_bases, B - 1, S, C, generator=g, device=device)
    out = torch.empty(n_images, B, S, C, dtype=dtype, device="cpu" if pin_host else device, pin_memory=pin_host)
    chunk = 64
    for i0 in range(0, n_images, chunk):
        i1 = min(n_images, i0 + chunk)
        idx = (torch.arange(i0, i1, device=device) // 3) % n_bases
        alpha = alpha_lo + (alpha_hi - alpha_lo) * torch.rand(i1 - i0, 1, 1, 1, generator=g, device=device)
        alpha[(torch.arange(i0, i1, device=device) % 3) == 0] = 1.0  # image 3t is the clean reference
        h = alpha * bases[idx] + torch.sqrt(1 - alpha * alpha) * torch.randn(i1 - i0, B, S, C, generator=g, device=device)
        out[i0:i1].copy_(h.to(dtype))
    return out, weight


def device_style_cache(B: int, H: int, S: int, D: int, i0: int, i1: int, per_style: int = 4, dtype=torch.float16, device="cuda",
                       alpha: float = 0.8, seed: int = 2334):
    """Images [i0, i1) of an Sref-shaped style set (image i belongs to style i // per_style; the images of a style share a
    base, alpha = 0.8), generated on the device.  Deterministic per image -- the generator is re-seeded per style and per
    image -- so any rank can build any image of the set (row-block sharded retrieval, and its single-GPU check).
    Returns q, k, v as (n,B,H,S,D) views over (n,B,S,H*D) memory."""
    C = H * D
    g = torch.Generator(device=device).manual_seed(seed)
    std = 1.0 / math.sqrt(C)
    Wq = torch.randn(C, C, generator=g, device=device) * (std * 2.2)
    Wk = torch.randn(C, C, generator=g, device=device) * (std * 2.2)
    Wv = torch.randn(C, C, generator=g, device=device) * std
    mems = [torch.empty(i1 - i0, B, S, C, dtype=dtype, device=device) for _ in range(3)]
    for i in range(i0, i1):
        gs = torch.Generator(device=device).manual_seed(1000003 + i // per_style)
        base = torch.randn(1, S, C, generator=gs, device=device).repeat(B, 1, 1)
        if B > 1:
            base[1:] += 0.1 * torch.randn(B - 1, S, C, generator=gs, device=device)
        gi = torch.Generator(device=device).manual_seed(7000001 + i)
        h = alpha * base + math.sqrt(1 - alpha * alpha) * torch.randn(B, S, C, generator=gi, device=device)
        mems[0][i - i0] = (h @ Wq).to(dtype)
        mems[1][i - i0] = (h @ Wk).to(dtype)
        mems[2][i - i0] = (h @ Wv).to(dtype)
    return tuple(m.view(i1 - i0, B, S, H, D).permute(0, 1, 3, 2, 4) for m in mems)
