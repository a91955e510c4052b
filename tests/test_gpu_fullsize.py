"""Parity at the REAL sizes of BASELINE.json's configs (the other GPU tests use shapes the CPU oracle finishes in seconds).

The checker is still oracle/aas_oracle.py -- the same float64 restatement, the same code -- but handed CUDA tensors, so that
its matmuls run in float64 on the GPU (the oracle is plain device-agnostic torch; a 64-image matrix at the SD-1.5 shape is
2.7 TFLOP of float64 work).  Every test first pins "oracle on the GPU == oracle on the CPU" on a sample of its own entries.
Decisions at full size (config 2) are compared with the reference's own torch lines run on the GPU in the native dtype
(tier T2: oracle.reference_pair_score = 4 x F.scaled_dot_product_attention + 2 x F.cosine_similarity,
diffsim/diffsim.py:177-197), identical outside the tolerance band, near-ties counted.
"""
import pytest
import torch

from conftest import checksum, regenerate_case
from oracle import aas_oracle as O

pytestmark = pytest.mark.gpu

REL_16BIT = 1e-3      # north_star: scores within 1e-3 relative for 16-bit inputs
SD15 = (2, 8, 256, 160)


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from diffsim_b200 import _native as N

    N.check(N.load().ds_device_ok())
    return "cuda"


def _lists(cache):
    q, k, v = cache
    n = q.shape[0]
    return [q[i] for i in range(n)], [k[i] for i in range(n)], [v[i] for i in range(n)]


def _topk_equal_outside_ties(s_got, s_ref, topk, band):
    """Per row: the top-k sets agree unless the reference's k-th and (k+1)-th scores are within the band (a tie the
    tolerance cannot resolve).  Returns (rows compared, rows skipped as ties)."""
    n = s_ref.shape[0]
    skipped = 0
    for i in range(n):
        ref_sorted, ref_idx = torch.sort(s_ref[i], descending=True)
        if (ref_sorted[topk - 1] - ref_sorted[topk]).item() <= band * abs(ref_sorted[topk - 1].item()):
            skipped += 1
            continue
        got_idx = torch.topk(s_got[i], topk).indices
        assert set(got_idx.tolist()) == set(ref_idx[:topk].tolist()), f"row {i}"
        if (ref_sorted[0] - ref_sorted[1]).item() > band * abs(ref_sorted[0].item()):
            assert int(s_got[i].argmax()) == int(ref_idx[0]), f"row {i} argmax"
    return n - skipped, skipped


# ------------------------------------------------------------------------------------------------------
# config 3: Sref-shaped retrieval at (2,8,256,160)
# ------------------------------------------------------------------------------------------------------
def test_cfg3_matrix_64_images_at_the_sd15_shape_matches_the_oracle():
    dev = _cuda()
    from diffsim_b200 import ops, scoring, synth

    B, H, S, D = SD15
    n = 64
    cache = synth.device_style_cache(B, H, S, D, 0, n, 4, torch.float16, dev)
    q, k, v = cache
    dm = ops.aas_matrix(q, k, v, k, v, "cosine")
    ql, kl, vl = _lists(cache)
    ref = O.aas_matrix(ql, kl, vl, "cosine")                      # the oracle, float64 on the GPU
    # the oracle gives the same numbers on the CPU (sampled entries)
    for i, j in ((0, 1), (5, 40), (63, 62), (17, 17)):
        cpu = O.aas_directional(q[i].cpu(), k[i].cpu(), v[i].cpu(), k[j].cpu(), v[j].cpu(), "cosine")
        assert cpu == pytest.approx(ref[i, j].item(), rel=1e-9, abs=1e-12)
    rel = ((dm.double().cpu() - ref).abs() / ref.abs().clamp_min(1e-9)).max().item()
    assert rel < REL_16BIT, rel
    assert dm.diagonal().sub(1).abs().max().item() < 1e-4
    s_got, s_ref = scoring.symmetrize(dm).double().cpu(), O.symmetrize(ref)
    s_got.fill_diagonal_(-1)
    s_ref.fill_diagonal_(-1)
    compared, skipped = _topk_equal_outside_ties(s_got, s_ref, 3, 2 * REL_16BIT)
    assert compared >= n - 4, (compared, skipped)
    # ground truth of the synthetic set: the three nearest neighbours are the other images of the style
    top3 = torch.topk(s_got, 3).indices
    assert all(set(top3[i].tolist()) == {j for j in range(4 * (i // 4), 4 * (i // 4) + 4) if j != i} for i in range(n))
    # MSE form on a row block
    dm_mse = ops.aas_matrix(q[:4], k[:4], v[:4], k, v, "mse")
    ref_mse = O.aas_matrix(ql, kl, vl, "mse", rows=range(4))
    off = ref_mse > 1e-6                                          # the diagonal is exactly 0 in both
    assert ((dm_mse.double().cpu() - ref_mse).abs()[off] / ref_mse[off]).max().item() < REL_16BIT
    assert dm_mse.diagonal().abs().max().item() == 0.0


def test_cfg3_sampled_rows_of_the_2032_image_matrix_match_the_oracle():
    dev = _cuda()
    from diffsim_b200 import ops, scoring, synth

    B, H, S, D = SD15
    n = 2032
    cache = synth.device_style_cache(B, H, S, D, 0, n, 4, torch.float16, dev)
    q, k, v = cache
    dm = ops.aas_matrix(q, k, v, k, v, "cosine")                  # 4.13 M directional scores, one call
    assert torch.isfinite(dm).all()
    rows = [0, 777, 1290, 2031]
    ql, kl, vl = _lists(cache)
    ref = O.aas_matrix(ql, kl, vl, "cosine", rows=rows)           # 4 x 2032 entries, float64 on the GPU
    got = dm[rows].double().cpu()
    assert ((got - ref).abs() / ref.abs().clamp_min(1e-9)).max().item() < REL_16BIT
    for r, i in enumerate(rows):                                  # directional top-3 of the sampled rows
        g, rf = got[r].clone(), ref[r].clone()
        g[i] = rf[i] = -1
        rs, ri = torch.sort(rf, descending=True)
        if (rs[2] - rs[3]).item() > 2 * REL_16BIT * abs(rs[2].item()):
            assert set(torch.topk(g, 3).indices.tolist()) == set(ri[:3].tolist())
    # a row block recomputed on its own (what a rank of the sharded run computes) is bit-identical
    blk = ops.aas_matrix(q[1000:1016], k[1000:1016], v[1000:1016], k, v, "cosine")
    assert torch.equal(blk, dm[1000:1016])
    # retrieval on the full symmetrised matrix recovers the styles
    from diffsim_b200 import retrieval

    acc = retrieval.retrieval_accuracy(scoring.symmetrize(dm), [i // 4 for i in range(n)], topk=3)
    assert acc["hit@k"] == 1.0 and acc["precision@k"] > 0.99, acc


# ------------------------------------------------------------------------------------------------------
# config 4: SDXL, 4096-token layers (real up_blocks[1] shape and the literal "4096 tokens x 1280 channels")
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(2, 10, 4096, 64), (2, 20, 4096, 64)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_cfg4_sdxl_4096_token_pair_matches_the_oracle(shape, dtype):
    dev = _cuda()
    from diffsim_b200 import ops, synth

    B, H, S, D = shape
    q, k, v = synth.device_cache(B, H, S, D, 2, dtype, dev, seed=11, alpha_lo=0.6, alpha_hi=0.9)
    for sim in ("cosine", "mse"):
        got = float(ops.aas_pairs(q, k, v, [(0, 1)], sim)[0])
        ref = O.aas_pair_score(q[0], k[0], v[0], q[1], k[1], v[1], mode=sim)      # T1, float64 on the GPU
        assert got == pytest.approx(ref, rel=REL_16BIT), (sim, got, ref)
    # the oracle on the CPU agrees with the oracle on the GPU on one (b, h) slice of the same inputs
    sl = lambda t, i: t[i][:1, :1].cpu()  # noqa: E731
    cpu = O.aas_directional(sl(q, 0), sl(k, 0), sl(v, 0), sl(k, 1), sl(v, 1), "cosine")
    gpu = O.aas_directional(q[0][:1, :1], k[0][:1, :1], v[0][:1, :1], k[1][:1, :1], v[1][:1, :1], "cosine")
    assert cpu == pytest.approx(gpu, rel=1e-9)


# ------------------------------------------------------------------------------------------------------
# config 2: 2048 NIGHTS-shaped triplets at the SD-1.5 shape, decisions against the reference's own lines on the GPU
# ------------------------------------------------------------------------------------------------------
def test_cfg2_2048_triplets_decide_like_the_reference_lines_on_the_gpu():
    dev = _cuda()
    from diffsim_b200 import ops, synth

    B, H, S, D = SD15
    T = 2048
    q, k, v = synth.device_cache(B, H, S, D, 3 * T, torch.float16, dev, seed=21)
    trips = torch.arange(3 * T, dtype=torch.int32, device=dev).view(T, 3)
    ab, ac, counts, flags = ops.aas_triplets(q, k, v, trips, "cosine")
    # tier T2: the reference's torch calls in fp16 on the same tensors (4 SDPA + 2 cosine per pair, fp16 scores)
    r_ab = torch.stack([O.reference_pair_score(q[3 * t], k[3 * t], v[3 * t], q[3 * t + 1], k[3 * t + 1], v[3 * t + 1]).reshape(())
                        for t in range(T)]).float()
    r_ac = torch.stack([O.reference_pair_score(q[3 * t], k[3 * t], v[3 * t], q[3 * t + 2], k[3 * t + 2], v[3 * t + 2]).reshape(())
                        for t in range(T)]).float()
    # scores: ours (fp32) vs the reference's fp16 scores -- equal up to the fp16 quantum of the score plus the tolerance
    quantum = 2.0 ** -11
    for got, ref in ((ab, r_ab), (ac, r_ac)):
        err = (got - ref).abs() / ref.abs().clamp_min(1e-6)
        assert err.max().item() < REL_16BIT + 2 * quantum, err.max().item()
    # decisions: identical wherever the reference's margin is outside the band; near-ties are counted and must be rare
    margin = (r_ab - r_ac).abs() / torch.maximum(r_ab.abs(), r_ac.abs()).clamp_min(1e-6)
    clear = margin > 2 * (REL_16BIT + quantum)
    ref_flags = (r_ab > r_ac)
    assert torch.equal(flags.bool()[clear], ref_flags[clear])
    near = int((~clear).sum())
    assert near < 0.05 * T, near
    assert int(counts[0]) == int(flags.sum())
    # and a sample of the scores against the float64 oracle (T1)
    for t in (0, 517, 2047):
        o = O.aas_pair_score(q[3 * t], k[3 * t], v[3 * t], q[3 * t + 1], k[3 * t + 1], v[3 * t + 1])
        assert float(ab[t]) == pytest.approx(o, rel=REL_16BIT)


# ------------------------------------------------------------------------------------------------------
# goldens: the SDXL-like case on the GPU, and the inputs really are the ones the reference saw
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["sdxl_like_f16", "sd15_up0_f16_cute16", "dit_xl2_f16_packed"])
def test_golden_inputs_are_the_reference_inputs_and_scores_match(golden, name):
    dev = _cuda()
    from diffsim_b200 import ops, synth

    case = {c["name"]: c for c in golden["cases"]}[name]
    images = regenerate_case(case)
    for im, cs in zip(images, case["checksums"]):
        assert [checksum(t) for t in im] == cs                   # bit-identical to what make_golden.py fed the reference
    q, k, v = synth.stack_cache(images, dev)
    for sim in ("cosine", "mse"):
        got = ops.aas_pairs(q, k, v, case["pairs"], sim).cpu().double()
        ref32 = torch.tensor(case["scores"][sim]["reference_fp32_math"], dtype=torch.float64)
        assert ((got - ref32).abs() / ref32.abs().clamp_min(1e-6)).max().item() < REL_16BIT


# ------------------------------------------------------------------------------------------------------
# K3 at the bench shape: 2032 x 655 360 features, sampled rows against the oracle
# ------------------------------------------------------------------------------------------------------
def test_k3_simmat_at_the_sref_shape_matches_the_oracle_on_sampled_rows():
    dev = _cuda()
    from diffsim_b200 import ops

    n, L = 2032, 655360
    g = torch.Generator(device=dev).manual_seed(5)
    feats = torch.empty(n, L, dtype=torch.float16, device=dev)
    base = torch.randn(64, L, generator=g, device=dev)
    for i0 in range(0, n, 127):
        i1 = min(n, i0 + 127)
        idx = torch.arange(i0, i1, device=dev) % 64
        feats[i0:i1] = (0.6 * base[idx] + 0.8 * torch.randn(i1 - i0, L, generator=g, device=dev) + 0.25).to(torch.float16)
    del base
    rows = torch.tensor([0, 1, 64, 65, 500, 501, 999, 1000, 1023, 1024, 1500, 1777, 2000, 2029, 2030, 2031], device=dev)
    for mode in ("cosine", "minmax_cosine"):
        c = ops.simmat(feats, None, mode)
        ref = O.simmat(feats[rows], feats, mode)                  # float64 on the GPU
        assert (c[rows].double() - ref.to(dev)).abs().max().item() < 2e-4, mode
        assert torch.equal(c, c.t())
        assert (c.diagonal() - 1).abs().max().item() < 2e-4
