"""The reference-facing Python surface on the GPU: scorer classes, capture hooks / processor, host-buffer scorer."""
import pytest
import torch

from oracle import aas_oracle as O

pytestmark = pytest.mark.gpu


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return "cuda"


def test_diffsim_class_returns_reference_dtype_and_shape():
    dev = _cuda()
    from diffsim_b200.diffsim import DiffSim, SyntheticTrunk, diffsim_DiT, diffsim_xl

    ds = DiffSim(torch.float16, dev, trunk=SyntheticTrunk((2, 8, 256, 160), torch.float16, dev))
    kw = dict(img_size=512, prompt="p", target_block="up_blocks", target_layer=[0], target_step=600, seed=2334, device=dev)
    s = ds.diffsim("cat@1.0", "cat@0.8", similarity="cosine", **kw)
    assert s.dtype == torch.float16 and s.shape == (1,)             # diffsim/diffsim.py:197 (cosine keeps a dim)
    m = ds.diffsim("cat@1.0", "cat@0.8", similarity="mse", **kw)
    assert m.dtype == torch.float16 and m.shape == ()
    A = ds.diffsim_value("cat@1.0", **kw)
    Bm = ds.diffsim_value("cat@0.8", **kw)
    assert A[0].shape == (2, 8, 256, 160) and A[0].stride() == (327680, 160, 1280, 1)
    ref = O.aas_pair_score(*[t.cpu() for t in A], *[t.cpu() for t in Bm])
    assert float(s) == pytest.approx(ref, rel=2e-3)                  # fp16 quantisation of the returned score
    assert float(ds.diffsim("cat@1.0", "cat@1.0", similarity="cosine", **kw)) == 1.0
    # the drivers' comparison works on the returned tensors (cute_main.py:196-205)
    far = ds.diffsim("cat@1.0", "dog@1.0", similarity="cosine", **kw)
    assert bool(s > far)
    # fp32 scores on request
    ds32 = DiffSim(torch.float16, dev, trunk=ds.trunk, match_reference_dtype=False)
    s32 = ds32.diffsim("cat@1.0", "cat@0.8", similarity="cosine", **kw)
    assert s32.dtype == torch.float32 and float(s32) == pytest.approx(ref, rel=1e-3)
    # DiT: packed-qkv views are consumed in place
    with pytest.raises(ValueError, match="needs a trunk"):
        DiffSim(torch.float16, dev)                                   # no silent synthetic default
    dit = diffsim_DiT(256, 600, dev, trunk=SyntheticTrunk((2, 16, 256, 72), torch.float16, dev, layout="dit"))
    d = dit.diffsim_score("cat@1.0", "cat@0.9", 256, "p", "up_blocks", [14], 600, "cosine", 2334)
    A, Bm = dit.trunk.extract("cat@1.0", target_step=600), dit.trunk.extract("cat@0.9", target_step=600)
    assert A[0].stride()[2] == 3 * 16 * 72
    assert float(d) == pytest.approx(O.aas_pair_score(*[t.cpu() for t in A], *[t.cpu() for t in Bm]), rel=2e-3)
    xl = diffsim_xl(torch.float16, dev, trunk=SyntheticTrunk((2, 20, 256, 64), torch.float16, dev))
    x = xl.diffsim_score("cat@1.0", "cat@0.9", 1024, "p", "up_blocks", [0, 1, 2], 600, "cosine", 2334)
    assert x.shape == (1,) and 0 < float(x) <= 1


class _FakeAttention(torch.nn.Module):
    """Shape of diffusers.models.attention_processor.Attention as far as the processors use it."""

    def __init__(self, dim, heads, dtype, dev):
        super().__init__()
        self.heads = heads
        self.to_q = torch.nn.Linear(dim, dim, bias=False, dtype=dtype, device=dev)
        self.to_k = torch.nn.Linear(dim, dim, bias=False, dtype=dtype, device=dev)
        self.to_v = torch.nn.Linear(dim, dim, bias=False, dtype=dtype, device=dev)
        self.to_out = torch.nn.ModuleList([torch.nn.Linear(dim, dim, dtype=dtype, device=dev), torch.nn.Dropout(0.0)])
        self.spatial_norm = self.group_norm = None
        self.norm_cross = False
        self.residual_connection = False
        self.rescale_output_factor = 1.0
        self.processor = None

    def forward(self, hidden_states):
        return self.processor(self, hidden_states)[0] if self.processor else hidden_states


def test_hooks_and_processor_follow_the_reference_contract():
    dev = _cuda()
    import torch.nn.functional as F

    from diffsim_b200 import hooks

    torch.manual_seed(0)
    attn = _FakeAttention(1280, 8, torch.float16, dev)
    x = torch.randn(2, 256, 1280, device=dev, dtype=torch.float16)
    # protocol 2: forward-pre-hook leaves module.stores = [q, k, v]
    with hooks.capture(attn, hooks.make_sd_pre_hook()):
        attn(x)
    q, k, v = attn.stores
    assert q.shape == (2, 8, 256, 160) and q.stride() == (327680, 160, 1280, 1)
    # the three projections ran as ONE ds_qkv_project call (K4) on the stacked weight: against the oracle's restatement of
    # hacked_attn.py:61-77 (float64) and, to an ulp of fp16, against the module's own nn.Linear (cuBLAS)
    ref_q, ref_k, ref_v = O.reference_capture(x.cpu(), attn.to_q.weight.cpu(), attn.to_k.weight.cpu(), attn.to_v.weight.cpu(), 8)
    for got, ref in ((q, ref_q), (k, ref_k), (v, ref_v)):
        assert got.shape == ref.shape
        assert (got.double().cpu() - ref.double()).abs().max().item() < 4e-3 * max(1.0, ref.abs().max().item())
    assert (q.float() - attn.to_q(x).view(2, 256, 8, 160).transpose(1, 2).float()).abs().max().item() < 4e-3
    assert hasattr(attn, "_ds_qkv_stack") and attn._ds_qkv_stack[1].shape == (3 * 1280, 1280)
    n0 = __import__("diffsim_b200").ops.LAUNCHES
    with hooks.capture(attn, hooks.make_sd_pre_hook()):
        attn(x)
    assert __import__("diffsim_b200").ops.LAUNCHES == n0 + 1      # one library launch per capture, stacked weight cached
    # fused_qkv=False: the module's own nn.Linear layers, bit for bit
    with hooks.capture(attn, hooks.make_sd_pre_hook(fused_qkv=False)):
        attn(x)
    assert torch.equal(attn.stores[0], attn.to_q(x).view(2, 256, 8, 160).transpose(1, 2))
    # capture straight into a cache slot
    from diffsim_b200 import scoring

    cache = scoring.QKVCache.empty(3, 2, 8, 256, 160, torch.float16, dev)
    slot = [m[1] for m in cache.memory()]
    with hooks.capture(attn, hooks.make_sd_pre_hook(out=slot)):
        attn(x)
    assert attn.stores[0].data_ptr() == cache.q[1].data_ptr() and torch.equal(cache.q[1], q) and torch.equal(cache.v[1], v)
    # weights replaced in place -> the stacked copy is rebuilt
    with torch.no_grad():
        attn.to_k.weight.mul_(0.5)
    with hooks.capture(attn, hooks.make_sd_pre_hook()):
        attn(x)
    assert (attn.stores[1].float() - 0.5 * k.float()).abs().max().item() < 4e-3
    with torch.no_grad():
        attn.to_k.weight.mul_(2.0)
    # a layer with spatial_norm: the pre-hook cannot serve it (no temb) and says so; the processor applies the norm itself
    attn.spatial_norm = lambda h, temb: h * 2.0
    with hooks.capture(attn, hooks.make_sd_pre_hook()):
        with pytest.raises(NotImplementedError, match="spatial_norm"):
            attn(x)
    _, q_sn, _, _, _ = hooks.B200AttnProcessor()(attn, x, temb=torch.zeros(1, device=dev))
    assert (q_sn.float() - 2.0 * q.float()).abs().max().item() < 8e-3
    attn.spatial_norm = None
    assert len(attn._forward_pre_hooks) == 0                      # removed (the reference accumulates them)
    with hooks.capture(attn, hooks.make_sd_pre_hook(early_exit=True)):
        with pytest.raises(hooks.StopForward):
            attn(x)
    # protocol 1: processor returning (hidden_states, q, k, v, residual) -- hacked_attn.py:101
    out, q2, k2, v2, res = hooks.B200AttnProcessor()(attn, x)
    assert torch.equal(q2, q) and res is x
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(2, 256, 1280)
    ref = attn.to_out[0](ref)
    assert (out.float() - ref.float()).abs().max().item() < 2e-2

    class _FakeTimmAttention(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.num_heads, self.head_dim = 16, 72
            self.qkv = torch.nn.Linear(1152, 3 * 1152, dtype=torch.float16, device=dev)
            self.q_norm = self.k_norm = torch.nn.Identity()

        def forward(self, x):
            return x

    t = _FakeTimmAttention()
    xt = torch.randn(2, 256, 1152, device=dev, dtype=torch.float16)
    with hooks.capture(t, hooks.make_dit_pre_hook()):
        t(xt)
    q, k, v = t.stores
    assert q.shape == (2, 16, 256, 72) and q.stride() == (256 * 3 * 1152, 72, 3 * 1152, 1)
    # module.qkv (with bias) ran on ds_qkv_project: equal to the module's own Linear up to an ulp of fp16
    packed = t.qkv(xt).reshape(2, 256, 3, 16, 72).permute(2, 0, 3, 1, 4)
    assert (q.float() - packed[0].float()).abs().max().item() < 4e-3 and (v.float() - packed[2].float()).abs().max().item() < 4e-3
    from diffsim_b200 import ops

    o = ops.attn_fwd(q, k, v)
    assert (o.float() - F.scaled_dot_product_attention(q, k, v).float()).abs().max().item() < 4e-3


def test_host_scorer_equals_device_scorer():
    dev = _cuda()
    from diffsim_b200 import scoring, synth

    shape = (2, 4, 128, 64)
    T = 50
    q, k, v = synth.device_cache(*shape, 3 * T, torch.float16, dev, seed=3)
    cache = scoring.QKVCache(q, k, v)
    trips = torch.arange(3 * T, dtype=torch.int32, device=dev).view(T, 3)
    _, _, counts, _ = scoring.score_triplets(cache, trips, "cosine")
    host = scoring.QKVCache.empty(3 * T, *shape, torch.float16, "cpu", pin=True)
    for hm, dm in zip(host.memory(), cache.memory()):
        hm.copy_(dm)
    scorer = scoring.HostTripletScorer(shape, torch.float16, dev, chunk_triplets=16)   # 4 chunks, ragged tail
    got = scorer.score(host, T)
    assert got == (int(counts[0]), int(counts[1]))
    assert scorer.h2d_bytes == 3 * T * cache.bytes_per_image and scorer.d2h_bytes == 8


def test_ip_adapter_processor_follows_the_reference_contract():
    """B200IPAdapterAttnProcessor vs a torch restatement of hacked_IPAdapterAttnProcessor2_0.__call__ (diffsim/hacked_attn.py:
    146-335, unmasked path): same 5-tuple, ip keys / values as head-split views, hidden states within fp16 tolerance; the
    returned (query, ip_keys, ip_values) then score through aas_score_ip_adapter."""
    dev = _cuda()
    import torch.nn.functional as F

    from diffsim_b200 import hooks
    from diffsim_b200.diffsim import aas_score_ip_adapter

    torch.manual_seed(1)
    C, heads, T_text, T_ip = 1280, 8, 77, 16
    attn = _FakeAttention(C, heads, torch.float16, dev)
    proc = hooks.B200IPAdapterAttnProcessor(C, C, torch.float16, num_tokens=(T_ip,), scale=0.7).to(dev)
    assert [tuple(m.weight.shape) for m in proc.to_k_ip] == [(C, C)]

    def ref_call(x, text, ip):
        q = attn.to_q(x).view(2, -1, heads, C // heads).transpose(1, 2)
        k = attn.to_k(text).view(2, -1, heads, C // heads).transpose(1, 2)
        v = attn.to_v(text).view(2, -1, heads, C // heads).transpose(1, 2)
        h = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(2, -1, C)
        ik = proc.to_k_ip[0](ip).view(2, -1, heads, C // heads).transpose(1, 2)
        iv = proc.to_v_ip[0](ip).view(2, -1, heads, C // heads).transpose(1, 2)
        h = h + 0.7 * F.scaled_dot_product_attention(q, ik, iv).transpose(1, 2).reshape(2, -1, C)
        return attn.to_out[0](h), q, ik, iv

    outs = []
    for seed in (0, 1):
        g = torch.Generator(device=dev).manual_seed(seed)
        x = torch.randn(2, 256, C, device=dev, dtype=torch.float16, generator=g)
        text = torch.randn(2, T_text, C, device=dev, dtype=torch.float16, generator=g)
        ip = torch.randn(2, T_ip, C, device=dev, dtype=torch.float16, generator=g)
        hs, q, ipk, ipv, res = proc(attn, x, encoder_hidden_states=(text, [ip]))
        ref_h, ref_q, ref_k, ref_v = ref_call(x, text, ip)
        assert res is x and len(ipk) == len(ipv) == 1
        assert ipk[0].shape == (2, heads, T_ip, C // heads) and ipk[0].stride(-1) == 1
        assert torch.equal(q, ref_q) and torch.equal(ipk[0], ref_k) and torch.equal(ipv[0], ref_v)
        assert (hs.float() - ref_h.float()).abs().max().item() < 3e-2
        # deprecated single-tensor form: the last num_tokens rows are the image tokens (hacked_attn.py:163-173)
        hs2, *_ = proc(attn, x, encoder_hidden_states=torch.cat([text, ip], dim=1))
        assert torch.equal(hs2, hs)
        outs.append((q, ipk, ipv))
    s = aas_score_ip_adapter(outs[0], outs[1], "cosine", match_reference_dtype=False)
    same = aas_score_ip_adapter(outs[0], outs[0], "cosine", match_reference_dtype=False)
    assert float(same) == pytest.approx(1.0, abs=1e-5) and -1.0 <= float(s) < 1.0
    with pytest.raises(NotImplementedError):
        proc(attn, x, encoder_hidden_states=(text, [ip]), ip_adapter_masks=[torch.ones(1, 1, 16, 16, device=dev)])


def test_cli_drivers_run_end_to_end(tmp_path, capsys):
    _cuda()
    from diffsim_b200 import __main__ as cli

    assert cli.main(["cute", "--similarity", "cosine", "--n", "6"]) == 0
    out = capsys.readouterr().out
    assert "Current total samples: 6" in out and "CUTE accuracy: 100.00%" in out     # positives are far closer than negatives
    assert cli.main(["nights", "--similarity", "mse", "--n", "5"]) == 0
    assert "Final validation accuracy: 100.00%" in capsys.readouterr().out
    assert cli.main(["sref", "--similarity", "cosine", "--n", "3", "--out_path", str(tmp_path)]) == 0
    out = capsys.readouterr().out
    assert "precision@3: 100.00%" in out and (tmp_path / "001" / "2.txt").exists()


def test_torch_custom_ops_equal_the_direct_bindings():
    dev = _cuda()
    from diffsim_b200 import ops, synth, torch_ops  # noqa: F401  (registers torch.ops.diffsim_b200.*)

    T = torch.ops.diffsim_b200
    q, k, v = synth.device_cache(2, 4, 128, 64, 6, torch.float16, dev, seed=5)
    pairs = torch.tensor([(0, 1), (2, 3), (4, 5)], dtype=torch.int32, device=dev)
    assert torch.equal(T.aas_pairs(q, k, v, pairs, "cosine"), ops.aas_pairs(q, k, v, pairs, "cosine"))
    assert torch.equal(T.aas_pairs(q, k, v, pairs, "mse", 0.1), ops.aas_pairs(q, k, v, pairs, "mse", 0.1))
    assert torch.equal(T.attn_fwd(q[0], k[1], v[1]), ops.attn_fwd(q[0], k[1], v[1]))
    s = T.aas_score(q[0], k[0], v[0], q[1], k[1], v[1], "cosine")
    assert torch.equal(s, ops.aas_pairs(q, k, v, pairs[:1], "cosine"))
    ab, ac, counts, flags = T.aas_triplets(q, k, v, torch.tensor([[0, 1, 2], [3, 4, 5]], dtype=torch.int32, device=dev))
    assert torch.equal(ab, ops.aas_pairs(q, k, v, torch.tensor([(0, 1), (3, 4)]), "cosine")) and counts.shape == (2,)
    assert torch.equal(T.aas_matrix(q, k, v, k, v), ops.aas_matrix(q, k, v, k, v))
    x = q[0].reshape(1, -1)
    y = q[1].reshape(1, -1)
    assert torch.equal(T.pair_reduce(x, y, "minmax_cosine"), ops.pair_reduce(x, y, "minmax_cosine"))
    f = torch.randn(40, 256, device=dev).half()
    assert torch.equal(T.simmat(f), ops.simmat(f)) and torch.equal(T.simmat(f, f[:8]), ops.simmat(f, f[:8]))
    h = torch.randn(64, 128, device=dev).half()
    w = torch.randn(3 * 128, 128, device=dev).half()
    outs = T.qkv_project(h, w, None, 3)
    assert len(outs) == 3 and all(torch.equal(a, b) for a, b in zip(outs, ops.qkv_project(h, w, None, 3)))
