"""Host-side logic that needs no GPU: layouts, caches, flags, sharding arithmetic, retrieval output format."""
import pytest
import torch

from diffsim_b200 import argprocess, scoring, synth
from diffsim_b200.diffsim import SyntheticTrunk, resolve_sd15_layer


def test_synth_layouts_match_the_reference_views():
    m = synth.SynthModel(2, 8, 256, 160)
    q, k, v = m.image(m.new_base(), 0.9)
    assert q.shape == (2, 8, 256, 160) and q.stride() == (327680, 160, 1280, 1)  # hacked_attn.py:74-77
    m = synth.SynthModel(2, 16, 256, 72)
    q, k, v = m.image(m.new_base(), 0.9, layout="dit")
    assert q.stride() == (256 * 3 * 16 * 72, 72, 3 * 16 * 72, 1)                  # diffsim_dit.py:22-23
    assert k.data_ptr() - q.data_ptr() == 16 * 72 * 2 and v.data_ptr() - k.data_ptr() == 16 * 72 * 2


def test_cache_views_are_zero_copy():
    m = synth.SynthModel(2, 2, 64, 40)
    base = m.new_base()
    images = [m.image(base, a) for a in (1.0, 0.8, 0.5)]
    cache = scoring.QKVCache.from_images(images)
    assert cache.n_images == 3 and cache.shape == (2, 2, 64, 40)
    assert cache.q.stride() == (2 * 64 * 80, 64 * 80, 40, 80, 1)
    for i, im in enumerate(images):
        assert torch.equal(cache.q[i], im[0]) and torch.equal(cache.v[i], im[2])
    qm, km, vm = cache.memory()
    assert qm.data_ptr() == cache.q.data_ptr() and qm.is_contiguous() and qm.shape == (3, 2, 64, 80)
    assert cache.bytes_per_image == 3 * 2 * 2 * 64 * 40 * 2
    sub = cache.slice(1, 3)
    assert sub.n_images == 2 and sub.k.data_ptr() == cache.k[1].data_ptr()


def test_row_blocks_cover_everything_once():
    for n, w in ((2032, 8), (2032, 3), (5, 8), (0, 2), (17, 4)):
        blocks = [scoring.row_block(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1


def test_ranked_lists_summary_lines():
    names = [f"{c}_{i}" for c in ("cat", "dog") for i in range(3)]
    S = torch.eye(6) * 0 + torch.tensor([[1 if a // 3 == b // 3 else 0 for b in range(6)] for a in range(6)]).float()
    S = S + 0.01 * torch.arange(6).float()[None]
    lines = scoring.ranked_lists(S, names, topk=2)
    assert len(lines) == 6
    for ln, name in zip(lines, names):
        head, rest = ln.split(":")            # '<query>: <best> <2nd>' (the parser-format files: tests/test_metrics_cpu.py)
        assert head == name and len(rest.split()) == 2
        assert all(r.split("_")[0] == name.split("_")[0] for r in rest.split())
        assert name not in rest.split()


def test_flags_keep_reference_names_and_defaults():
    a = argprocess.arg_parse([])
    assert (a.image_size, a.target_block, a.target_layer, a.target_step) == (512, "up_blocks", 2, 100)
    assert (a.metric, a.similarity, a.prompt, a.seed) == ("diffsim", "mse", "High quality image", 2333)
    a = argprocess.arg_parse("--target_block up_blocks --target_layer 0 --target_step 600 --similarity cosine --seed 2334".split())
    assert a.target_layer == [0] and a.similarity == "cosine"           # cute_main.sh:3
    a = argprocess.arg_parse("--metric diffsim_xl --target_layer 0 1 2".split())
    assert a.target_layer == [0, 1, 2]
    with pytest.raises(SystemExit):
        argprocess.arg_parse(["--similarity", "l1"])
    assert argprocess.BENCHMARK_PRESETS["nights"]["target_step"] == 500


def test_layer_collapse_quirk():
    assert resolve_sd15_layer([5]) == 0                 # diffsim/diffsim.py:99-100 (ipref_main.sh passes 5)
    assert resolve_sd15_layer([5], compat_layer_collapse=False) == 5
    with pytest.raises(ValueError):
        resolve_sd15_layer([0, 1])


def test_synthetic_trunk_is_deterministic_and_similarity_ordered():
    from oracle import aas_oracle as O

    tr = SyntheticTrunk((2, 2, 64, 40), torch.float32, "cpu")
    a1 = tr.extract("cat@1.0")
    a2 = tr.extract("cat@1.0")
    assert all(torch.equal(x, y) for x, y in zip(a1, a2))
    near, far, other = tr.extract("cat@0.95"), tr.extract("cat@0.5"), tr.extract("dog@1.0")
    s = [O.aas_pair_score(*a1, *b) for b in (near, far, other)]
    assert s[0] > s[1] > s[2]
    assert a1[0].stride() == (64 * 80, 40, 80, 1)


def test_triplet_generator_has_near_ties_and_margins():
    m = synth.SynthModel(2, 2, 64, 40)
    images, trips = synth.make_triplets(m, 6, torch.float16, seed=1)
    assert len(images) == 18 and trips[2] == (6, 7, 8)
    images2, _ = synth.make_triplets(m, 6, torch.float16, seed=1)
    assert all(torch.equal(a[0], b[0]) for a, b in zip(images, images2))


def test_hostbind_parses_sysfs_lists_and_never_raises():
    from diffsim_b200 import hostbind

    assert hostbind._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert hostbind._parse_cpulist("") == []
    r = hostbind.bind_to_gpu_node(0)          # no GPU here: the affinity is left alone, with the reason stated
    assert r["bound"] is False and isinstance(r["why"], str)


def test_cli_keeps_reference_flags_and_launcher_presets():
    from diffsim_b200 import __main__ as cli

    a = cli.build_parser().parse_args(["nights", "--similarity", "cosine", "--seed", "2334"])
    assert cli.resolve(a) == {"target_block": "up_blocks", "target_layer": [0], "target_step": 500}   # night_main.sh
    a = cli.build_parser().parse_args(["cute", "--target_step", "750", "--target_layer", "3", "--target_block", "down_blocks"])
    assert cli.resolve(a) == {"target_block": "down_blocks", "target_layer": [3], "target_step": 750}   # explicit flags win
    assert a.similarity == "mse" and a.metric == "diffsim"                                            # reference defaults


def test_torch_custom_ops_are_registered_and_have_no_cpu_fallback():
    import torch as _t

    from diffsim_b200 import torch_ops

    for name in torch_ops.OPERATORS:
        assert hasattr(_t.ops.diffsim_b200, name)
    x = _t.randn(2, 64).half()
    with pytest.raises(NotImplementedError):          # the dispatcher has no CPU kernel to fall back to
        _t.ops.diffsim_b200.pair_reduce(x, x, "cosine")
    q = _t.randn(1, 2, 64, 64).half()
    with pytest.raises(NotImplementedError):
        _t.ops.diffsim_b200.attn_fwd(q, q, q)
