"""Edge cases of the C ABI on the GPU: empty work lists, degenerate sizes, misuse that must fail loudly, and size-independent
properties at the full benchmark shape (symmetry of the pair score, triplets == pairs, matrix == pairs, self-score == 1)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def test_empty_work_lists_are_no_ops():
    dev = _cuda()
    from diffsim_b200 import ops, synth

    q, k, v = synth.device_cache(1, 2, 128, 64, 3, torch.float16, dev)
    assert ops.aas_pairs(q, k, v, torch.zeros((0, 2), dtype=torch.int32), "cosine").shape == (0,)
    ab, ac, counts, flags = ops.aas_triplets(q, k, v, torch.zeros((0, 3), dtype=torch.int32), "cosine")
    assert ab.shape == (0,) and ac.shape == (0,) and flags.shape == (0,)
    assert counts.tolist() == [0, 0]
    x = torch.zeros((0, 16), dtype=torch.float16, device=dev)
    assert ops.pair_reduce(x, x, "cosine").shape == (0,)
    c, f = ops.twoafc(torch.zeros(0, device=dev), torch.zeros(0, device=dev))
    assert c.tolist() == [0, 0] and f.shape == (0,)


def test_degenerate_values_follow_torch_semantics():
    """All-zero vectors: F.cosine_similarity clamps each norm at 1e-8 -> 0; identical vectors -> 1; mse of equal -> 0."""
    dev = _cuda()
    from diffsim_b200 import ops

    z = torch.zeros((2, 4096), dtype=torch.float16, device=dev)
    o = torch.ones((2, 4096), dtype=torch.float16, device=dev)
    assert ops.pair_reduce(z, o, "cosine").tolist() == [0.0, 0.0]
    assert ops.pair_reduce(z, z, "cosine").tolist() == [0.0, 0.0]
    assert ops.pair_reduce(o, o, "cosine").tolist() == pytest.approx([1.0, 1.0], abs=1e-6)
    assert ops.pair_reduce(o, o, "mse").tolist() == [0.0, 0.0]
    # constant feature maps: min-max normalisation divides by zero in the reference (nan); the kernel must not hide that
    r = ops.pair_reduce(o, o, "minmax_cosine")
    assert torch.isnan(r).all() or torch.isinf(r).all() or (r == 0).all()


def test_misuse_fails_loudly():
    dev = _cuda()
    from diffsim_b200 import ops, synth
    from diffsim_b200._native import DiffSimError

    q, k, v = synth.device_cache(1, 2, 128, 64, 4, torch.float16, dev)
    with pytest.raises(DiffSimError):       # mismatched head counts between q and k
        ops.aas_pairs(q, k[:, :, :1], v[:, :, :1], [(0, 1)], "cosine")
    with pytest.raises(DiffSimError):       # innermost stride must be 1
        ops.attn_fwd(q[0].transpose(-1, -2).contiguous().transpose(-1, -2), k[0], v[0])
    with pytest.raises((DiffSimError, RuntimeError)):
        ops.pair_reduce(q[0].reshape(1, -1), k[0].reshape(1, -1)[:, :-8], "cosine")
    # the reference treats every similarity string other than 'cosine' as MSE (diffsim/diffsim.py:182,189): mirrored
    assert torch.equal(ops.aas_pairs(q, k, v, [(0, 1)], "l1"), ops.aas_pairs(q, k, v, [(0, 1)], "mse"))
    with pytest.raises(DiffSimError):       # ... but an out-of-range mode code at the ABI is an error
        ops.aas_pairs(q, k, v, [(0, 1)], 7)
    with pytest.raises((DiffSimError, RuntimeError)):
        ops.simmat(q[0].reshape(2, -1).float())            # fp32 features are not accepted by the GEMM


def test_full_size_properties_without_an_oracle():
    """SD-1.5 up0 shape, 64 images: properties that hold at any size."""
    dev = _cuda()
    from diffsim_b200 import ops, synth

    n = 64
    q, k, v = synth.device_cache(2, 8, 256, 160, n, torch.float16, dev, seed=11)
    idx = torch.arange(n, dtype=torch.int32)
    pairs = torch.stack([idx, (idx * 5 + 1) % n], 1)
    s_ab = ops.aas_pairs(q, k, v, pairs, "cosine")
    s_ba = ops.aas_pairs(q, k, v, pairs.flip(1), "cosine")
    assert torch.equal(s_ab, s_ba)                                      # diffsim(A,B) == diffsim(B,A), bit for bit
    self_pairs = torch.stack([idx, idx], 1)
    assert (ops.aas_pairs(q, k, v, self_pairs, "cosine") - 1).abs().max().item() < 2e-6
    assert ops.aas_pairs(q, k, v, self_pairs, "mse").abs().max().item() == 0.0
    # triplets are two pairs sharing the reference image
    trips = torch.stack([idx[:21] * 3 % n, (idx[:21] * 3 + 1) % n, (idx[:21] * 3 + 2) % n], 1)
    ab, ac, counts, flags = ops.aas_triplets(q, k, v, trips, "cosine")
    assert torch.equal(ab, ops.aas_pairs(q, k, v, trips[:, [0, 1]], "cosine"))
    assert torch.equal(ac, ops.aas_pairs(q, k, v, trips[:, [0, 2]], "cosine"))
    assert int(counts[0]) == int((ab > ac).sum()) == int(flags.sum())
    # the all-pairs matrix holds the same directional values: S = (D + D^T) / 2 reproduces the pair scores
    dm = ops.aas_matrix(q[:16], k[:16], v[:16], k[:16], v[:16], "cosine")
    s = (dm + dm.t()) * 0.5
    p16 = torch.tensor([(i, j) for i in range(16) for j in range(16)], dtype=torch.int32)
    assert torch.equal(s.reshape(-1), ops.aas_pairs(q[:16], k[:16], v[:16], p16, "cosine"))
    # launching twice gives the same bits (fixed reduction order, no atomics on floats)
    assert torch.equal(dm, ops.aas_matrix(q[:16], k[:16], v[:16], k[:16], v[:16], "cosine"))


def test_out_of_range_image_indices_give_nan_not_a_plausible_score():
    """An index past the cache would become a TMA coordinate outside the tensor (zero-filled: a finite, wrong score): the
    work list is validated on the device and every score of the call comes back NaN."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from diffsim_b200 import ops, synth

    q, k, v = synth.device_cache(2, 4, 128, 64, 6, torch.float16, "cuda", seed=1)
    good = ops.aas_pairs(q, k, v, [(0, 1), (2, 3)], "cosine")
    assert torch.isfinite(good).all()
    for bad_pairs in ([(0, 1), (2, 6)], [(-1, 1)], [(0, 99999)]):
        assert torch.isnan(ops.aas_pairs(q, k, v, bad_pairs, "cosine")).all()
    ab, ac, counts, flags = ops.aas_triplets(q, k, v, [(0, 1, 2), (3, 4, 7)], "cosine")
    assert torch.isnan(ab).all() and torch.isnan(ac).all() and int(counts[0]) == 0
    assert torch.isfinite(ops.aas_pairs(q, k, v, [(4, 5)], "mse")).all()        # the flag does not stick to the next call
    # two streams do not share a workspace
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    with torch.cuda.stream(s1):
        a = ops.aas_pairs(q, k, v, [(0, 1)] * 64, "cosine")
    with torch.cuda.stream(s2):
        b = ops.aas_pairs(q, k, v, [(2, 3)] * 64, "cosine")
    torch.cuda.synchronize()
    assert torch.equal(a, good[:1].expand(64)) and torch.equal(b, good[1:].expand(64))
