"""CPU oracle of the DiffSim scoring hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module, and only as the checker or the reported CPU baseline -- never as part of the product path.

The reference is pure Python; the arithmetic of its hot path lives in a third-party dependency that is not
under /root/reference: torch (pinned torch==2.3.0 in the reference's requirements.txt:19; this image has
torch 2.11.0).  The functions below RESTATE that arithmetic explicitly (matmul / softmax / sums written out,
float64 by default) instead of calling F.scaled_dot_product_attention / F.cosine_similarity / F.mse_loss, so
that the oracle is independent of the library routines the reference happens to call.

Pinning: the reference has no tests, golden vectors or fixtures for this path.  The
oracle is pinned instead against outputs of the reference's own code run in the build container:
tests/golden/make_golden.py imports /root/reference/diffsim/diffsim.py (third-party imports that are absent --
diffusers, seaborn, matplotlib -- are stubbed; the trunk is replaced by a fake pipeline that deposits synthetic
Q/K/V in `module.stores`) and executes DiffSim.diffsim, i.e. lines 98-197 including the verbatim score
arithmetic 177-197, and the attention_calc / min_max_normalize helpers of metrics/dino.py and
metrics/diffeats.py.  tests/test_oracle.py checks this module against those vectors.

Tiers:
  T0  float64 arithmetic on the (16-bit) inputs, nothing rounded            -- "truth"
  T1  float64 attention, attention outputs rounded to the input dtype (as the reference's SDPA outputs are
      and as the CUDA kernel does), float64 reductions                       -- what the kernel is compared to
  T2  the reference's own torch calls in the native dtype (reference_pair_score below)
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch


# ----------------------------------------------------------------------------------------------------
# attention  (diffsim/diffsim.py:177-180, diffsim/hacked_attn.py:81-83; explicit form metrics/dino.py:120-131)
# ----------------------------------------------------------------------------------------------------
def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: Optional[float] = None,
              round_to: Optional[torch.dtype] = None, work_dtype: torch.dtype = torch.float64) -> torch.Tensor:
    """softmax(q k^T * scale) v over the last two dims; non-causal, no mask, no dropout.
    scale defaults to 1/sqrt(D) like F.scaled_dot_product_attention.  If round_to is given the result is
    rounded to that dtype and returned in work_dtype (models an SDPA output stored in fp16/bf16)."""
    D = q.shape[-1]
    sc = (1.0 / math.sqrt(D)) if scale is None else float(scale)
    qf, kf, vf = q.to(work_dtype), k.to(work_dtype), v.to(work_dtype)
    s = torch.matmul(qf, kf.transpose(-1, -2)) * sc
    s = s - s.amax(dim=-1, keepdim=True)
    p = torch.exp(s)
    p = p / p.sum(dim=-1, keepdim=True)
    o = torch.matmul(p, vf)
    if round_to is not None:
        o = o.to(round_to).to(work_dtype)
    return o


# ----------------------------------------------------------------------------------------------------
# reductions
# ----------------------------------------------------------------------------------------------------
def flat_cosine(x: torch.Tensor, y: torch.Tensor, eps: float = 1e-8) -> float:
    """F.cosine_similarity(x.reshape(-1).unsqueeze(0), y.reshape(-1).unsqueeze(0)) (diffsim/diffsim.py:187-188):
    torch divides each vector by max(|v|, eps) and sums the products."""
    xf, yf = x.reshape(-1).to(torch.float64), y.reshape(-1).to(torch.float64)
    nx = max(float(torch.sqrt((xf * xf).sum())), eps)
    ny = max(float(torch.sqrt((yf * yf).sum())), eps)
    return float((xf * yf).sum()) / (nx * ny)


def mse(x: torch.Tensor, y: torch.Tensor) -> float:
    """F.mse_loss(x, y) (diffsim/diffsim.py:194-195): mean of squared differences."""
    d = x.reshape(-1).to(torch.float64) - y.reshape(-1).to(torch.float64)
    return float((d * d).mean())


def min_max_normalize(t: torch.Tensor) -> torch.Tensor:
    """metrics/diffeats.py:136-140."""
    tf = t.to(torch.float64)
    mn, mx = tf.min(), tf.max()
    return (tf - mn) / (mx - mn)


def minmax_cosine(x: torch.Tensor, y: torch.Tensor) -> float:
    """metrics/diffeats.py:202-205: min-max normalise both feature maps, then the flat cosine."""
    return flat_cosine(min_max_normalize(x), min_max_normalize(y))


def similarity(x: torch.Tensor, y: torch.Tensor, mode: str) -> float:
    if mode == "cosine":
        return flat_cosine(x, y)
    if mode == "minmax_cosine":
        return minmax_cosine(x, y)
    return mse(x, y)  # the reference treats every non-'cosine' value as MSE (diffsim/diffsim.py:182,189)


# ----------------------------------------------------------------------------------------------------
# Aligned Attention Score
# ----------------------------------------------------------------------------------------------------
def aas_directional(q_i, k_i, v_i, k_j, v_j, mode: str = "cosine", scale: Optional[float] = None,
                    tier: str = "T1") -> float:
    """sim( Attn(Q_i,K_j,V_j), Attn(Q_i,K_i,V_i) ): one of the two terms of diffsim/diffsim.py:187-188 / 194-195."""
    rnd = q_i.dtype if (tier == "T1" and q_i.dtype in (torch.float16, torch.bfloat16)) else None
    cross = attention(q_i, k_j, v_j, scale, round_to=rnd)
    self_ = attention(q_i, k_i, v_i, scale, round_to=rnd)
    return similarity(cross, self_, mode)


def aas_pair_score(qa, ka, va, qb, kb, vb, mode: str = "cosine", scale: Optional[float] = None,
                   tier: str = "T1") -> float:
    """(diffsim_a_on_b + diffsim_b_on_a) / 2 -- diffsim/diffsim.py:177-197 (copies diffsim_xl.py:135-155,
    diffsim_dit.py:130-142)."""
    a_on_b = aas_directional(qa, ka, va, kb, vb, mode, scale, tier)
    b_on_a = aas_directional(qb, kb, vb, ka, va, mode, scale, tier)
    return (a_on_b + b_on_a) / 2.0


def aas_ip_adapter_score(qa, ka_list, va_list, qb, kb_list, vb_list, scale: Optional[float] = None,
                         tier: str = "T1") -> float:
    """IP-Adapter ("DiffSim-C") cosine form, diffsim/diffsim.py:172-175,184-185: one attention per ip key/value
    tensor, mean of the per-tensor cosines."""
    ab = [aas_directional(qa, ka, va, kb, vb, "cosine", scale, tier)
          for ka, va, kb, vb in zip(ka_list, va_list, kb_list, vb_list)]
    ba = [aas_directional(qb, kb, vb, ka, va, "cosine", scale, tier)
          for ka, va, kb, vb in zip(ka_list, va_list, kb_list, vb_list)]
    return (sum(ab) / len(ab) + sum(ba) / len(ba)) / 2.0


def aas_matrix(q: Sequence[torch.Tensor], k: Sequence[torch.Tensor], v: Sequence[torch.Tensor], mode: str = "cosine",
               scale: Optional[float] = None, tier: str = "T1", rows: Optional[Sequence[int]] = None) -> torch.Tensor:
    """Directional matrix Dm[i,j] = dir(i -> j): the pair formula applied to every (i,j) (the reference ships no
    N x N code, only the consumer of its output -- retrieval_vis.py:57-68)."""
    n = len(q)
    rows = list(range(n)) if rows is None else list(rows)
    out = torch.zeros((len(rows), n), dtype=torch.float64)
    rnd = q[0].dtype if (tier == "T1" and q[0].dtype in (torch.float16, torch.bfloat16)) else None
    for ri, i in enumerate(rows):
        self_ = attention(q[i], k[i], v[i], scale, round_to=rnd)
        for j in range(n):
            cross = attention(q[i], k[j], v[j], scale, round_to=rnd)
            out[ri, j] = similarity(cross, self_, mode)
    return out


def symmetrize(dm: torch.Tensor) -> torch.Tensor:
    """S = (Dm + Dm^T) / 2: the pair score of diffsim/diffsim.py:197 for every (i,j)."""
    return (dm + dm.transpose(0, 1)) / 2.0


def simmat(rows: torch.Tensor, cols: torch.Tensor, mode: str = "cosine") -> torch.Tensor:
    """All-pairs flat cosine / min-max cosine of feature vectors (metrics/diffeats.py:202-205, clip_i.py:183)."""
    out = torch.zeros((rows.shape[0], cols.shape[0]), dtype=torch.float64)
    for i in range(rows.shape[0]):
        for j in range(cols.shape[0]):
            out[i, j] = similarity(rows[i], cols[j], mode)
    return out


# ----------------------------------------------------------------------------------------------------
# decisions (cute_main.py:196-205, night_main.py:157-163)
# ----------------------------------------------------------------------------------------------------
def twoafc(ab: Sequence[float], ac: Sequence[float], mode: str = "cosine"):
    """Returns (correct, correct_2x, flags) with the drivers' strict inequalities."""
    correct = correct2 = 0
    flags = []
    for a, c in zip(ab, ac):
        if mode == "cosine":
            ok, ok2 = a > c, a > 2 * c
        else:
            ok, ok2 = a < c, a * 2 < c
        flags.append(bool(ok))
        correct += int(ok)
        correct2 += int(ok2)
    return correct, correct2, flags


# ----------------------------------------------------------------------------------------------------
# T2: the reference's own torch calls, native dtype.  Used as the CPU baseline ("what the reference's CPU
# inference path executes per pair") and reported beside the oracle; the formula is diffsim/diffsim.py:177-197.
# ----------------------------------------------------------------------------------------------------
def reference_pair_score(qa, ka, va, qb, kb, vb, mode: str = "cosine") -> torch.Tensor:
    import torch.nn.functional as F

    sdpa = lambda q, k, v: F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False)  # noqa: E731
    a_on_b, b_on_a = sdpa(qa, kb, vb), sdpa(qb, ka, va)
    self_a, self_b = sdpa(qa, ka, va), sdpa(qb, kb, vb)
    if mode == "cosine":
        flat = lambda t: t.reshape(-1).unsqueeze(0)  # noqa: E731
        s_ab = F.cosine_similarity(flat(a_on_b), flat(self_a))
        s_ba = F.cosine_similarity(flat(b_on_a), flat(self_b))
    else:
        s_ab, s_ba = F.mse_loss(a_on_b, self_a), F.mse_loss(b_on_a, self_b)
    return (s_ab + s_ba) / 2


# ----------------------------------------------------------------------------------------------------
# QKV projection of the hooked layer (diffsim/hacked_attn.py:61-69,74-77; diffsim/diffsim_dit.py:21-23)
# ----------------------------------------------------------------------------------------------------
def project_qkv(hidden: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, n_outputs: int = 3,
                round_to: Optional[torch.dtype] = None):
    """y = hidden @ weight^T (+ bias) in float64 (weight in nn.Linear layout [n_out, C_in]; for separate
    to_q / to_k / to_v modules the three weights stacked along dim 0), split into n_outputs column blocks.
    round_to: round the result to that dtype (what an fp16 / bf16 nn.Linear returns) and hand it back as such."""
    y = torch.matmul(hidden.to(torch.float64), weight.to(torch.float64).t())
    if bias is not None:
        y = y + bias.to(torch.float64)
    if round_to is not None:
        y = y.to(round_to)
    return tuple(y.chunk(n_outputs, dim=-1))


def reference_capture(hidden: torch.Tensor, wq: torch.Tensor, wk: torch.Tensor, wv: torch.Tensor, heads: int):
    """T2: the reference's capture lines in the native dtype -- attn.to_q / to_k / to_v on the hook input and the
    head-split views (diffsim/hacked_attn.py:61-69,74-77); bias-free as SD's attn1 projections are."""
    import torch.nn.functional as F

    B, S, _ = hidden.shape
    out = []
    for w in (wq, wk, wv):
        y = F.linear(hidden, w)
        out.append(y.view(B, S, heads, y.shape[-1] // heads).transpose(1, 2))
    return tuple(out)


# ----------------------------------------------------------------------------------------------------
# baseline-metric AAS variants
# ----------------------------------------------------------------------------------------------------
def clip_attention_calc(q, k, v, scale: float, hidden_size_shape, out_w: torch.Tensor, out_b: Optional[torch.Tensor],
                        round_to: Optional[torch.dtype] = None) -> torch.Tensor:
    """metrics/clip_i.py:113-127: attention with an explicit scale, heads merged to (bsz, tgt_len, embed_dim), out_proj.
    round_to models the 16-bit storage of the attention output and of the projection result."""
    bsz, tgt_len, embed_dim = hidden_size_shape
    o = attention(q, k, v, scale, round_to=round_to)
    o = o.transpose(1, 2).reshape(bsz, tgt_len, embed_dim)
    y = torch.matmul(o, out_w.to(torch.float64).t())
    if out_b is not None:
        y = y + out_b.to(torch.float64)
    if round_to is not None:
        y = y.to(round_to).to(torch.float64)
    return y


def clip_cross_score(qa, ka, va, qb, kb, vb, scale: float, hidden_size_shape, out_w, out_b,
                     round_to: Optional[torch.dtype] = None) -> float:
    """metrics/clip_i.py:130-159."""
    f = lambda q, k, v: clip_attention_calc(q, k, v, scale, hidden_size_shape, out_w, out_b, round_to)  # noqa: E731
    a_on_b, b_on_a, self_a, self_b = f(qa, kb, vb), f(qb, ka, va), f(qa, ka, va), f(qb, kb, vb)
    return (flat_cosine(a_on_b, self_a) + flat_cosine(b_on_a, self_b)) / 2.0


def gram_matrix(features: torch.Tensor, round_to: Optional[torch.dtype] = None) -> torch.Tensor:
    """metrics/vgg_gram.py:57-69: (b,d,h,w) -> (b*d, h*w) -> F F^T."""
    b, d, h, w = features.shape
    f = features.reshape(b * d, h * w).to(torch.float64)
    g = f @ f.t()
    return g.to(round_to).to(torch.float64) if round_to is not None else g


def gram_similarity(fa: torch.Tensor, fb: torch.Tensor, round_to: Optional[torch.dtype] = None) -> float:
    """metrics/vgg_gram.py:81: cosine of the last ROW of each Gram matrix (as written in the reference)."""
    return flat_cosine(gram_matrix(fa, round_to)[-1], gram_matrix(fb, round_to)[-1])
