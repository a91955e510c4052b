"""DiffusersTrunk on a fake diffusers pipeline (CPU): which attn1 module each (--target_block, --target_layer) pair selects --
checked against the reference's own indexing expressions (diffsim/diffsim.py:122-145 for diffsim(), :221-244 for
diffsim_value() whose slices are swapped, diffsim/diffsim_xl.py:88-107) --, the hook contract, the early exit, and that
--target_step is an INDEX into the 1000-step timestep array (diffsim_pipeline.py:153-157)."""
import types

import pytest
import torch
from PIL import Image

from diffsim_b200 import hooks
from diffsim_b200.diffsim import DiffusersTrunk


class _Attn(torch.nn.Module):
    def __init__(self, name, dim=32, heads=4):
        super().__init__()
        self.name, self.heads, self.calls = name, heads, 0
        self.to_q = torch.nn.Linear(dim, dim, bias=False)
        self.to_k = torch.nn.Linear(dim, dim, bias=False)
        self.to_v = torch.nn.Linear(dim, dim, bias=False)
        self.spatial_norm = self.group_norm = None
        self.norm_cross = False

    def forward(self, x):
        self.calls += 1
        return x


def _block(name, n_att=2, n_tb=2):
    atts = []
    for a in range(n_att):
        tbs = [types.SimpleNamespace(attn1=_Attn(f"{name}.a{a}.t{t}")) for t in range(n_tb)]
        atts.append(types.SimpleNamespace(transformer_blocks=tbs))
    return types.SimpleNamespace(attentions=atts)


class _UNet:
    def __init__(self):
        self.down_blocks = [_block(f"down{i}") for i in range(4)]
        self.mid_block = _block("mid")
        self.up_blocks = [_block(f"up{i}") for i in range(4)]
        self.seen_t = None

    def all_attn(self):
        for blk in self.down_blocks + [self.mid_block] + self.up_blocks:
            for a in blk.attentions:
                for tb in a.transformer_blocks:
                    yield tb.attn1

    def __call__(self, sample, t, encoder_hidden_states=None):
        self.seen_t = int(t)
        x = torch.randn(sample.shape[0], 16, 32, generator=torch.Generator().manual_seed(0))
        for m in self.all_attn():          # execution order of a UNet: down, mid, up
            x = m(x)
        return (x,)


class _Pipe:
    def __init__(self):
        self.unet = _UNet()
        lat = types.SimpleNamespace(sample=lambda generator=None: torch.zeros(1, 4, 8, 8))
        self.vae = types.SimpleNamespace(encode=lambda x: types.SimpleNamespace(latent_dist=lat),
                                         config=types.SimpleNamespace(scaling_factor=0.18215))
        ts = torch.arange(999, -1, -1)
        self.scheduler = types.SimpleNamespace(timesteps=ts, set_timesteps=lambda n, device=None: None,
                                               add_noise=lambda lat, noise, t: lat + noise,
                                               scale_model_input=lambda x, t: x)
        self.encode_calls = 0

    def encode_prompt(self, prompt, device, n, cfg, neg):
        self.encode_calls += 1
        return torch.zeros(1, 77, 32), torch.zeros(1, 77, 32)


def _ref_sd15(unet, block, layer, value_mode):
    """The reference's indexing, as written (diffsim/diffsim.py:125-145 / :224-244)."""
    if block == "down_blocks":
        blocks = unet.down_blocks[1:] if value_mode else unet.down_blocks[:-1]
    elif block == "mid_blocks":
        return unet.mid_block.attentions[-1].transformer_blocks[-1].attn1
    else:
        blocks = unet.up_blocks[:-1] if value_mode else unet.up_blocks[1:]
    return blocks[layer].attentions[-1].transformer_blocks[-1].attn1


@pytest.mark.parametrize("value_mode", [False, True])
@pytest.mark.parametrize("block,layer", [("down_blocks", 0), ("down_blocks", 2), ("mid_blocks", 0), ("up_blocks", 0), ("up_blocks", 2)])
def test_sd15_target_module_matches_the_reference_indexing(block, layer, value_mode):
    pipe = _Pipe()
    trunk = DiffusersTrunk(pipe, "cpu", torch.float32, value_mode=value_mode)
    assert trunk.target_module(block, layer) is _ref_sd15(pipe.unet, block, layer, value_mode)


def test_sdxl_target_module_takes_three_indices():
    pipe = _Pipe()
    trunk = DiffusersTrunk(pipe, "cpu", torch.float32, kind="sdxl")
    u = pipe.unet
    assert trunk.target_module("up_blocks", [1, 0, 1]) is u.up_blocks[:-1][1].attentions[0].transformer_blocks[1].attn1
    assert trunk.target_module("down_blocks", [0, 1, 0]) is u.down_blocks[1:][0].attentions[1].transformer_blocks[0].attn1
    assert trunk.target_module("mid_blocks", [1, 0]) is u.mid_block.attentions[1].transformer_blocks[0].attn1


def test_extract_captures_qkv_stops_early_and_leaves_no_hook():
    pipe = _Pipe()
    trunk = DiffusersTrunk(pipe, "cpu", torch.float32)
    img = Image.new("RGB", (20, 12), (200, 30, 90))
    q, k, v = trunk.extract(img, 16, "a photo", "up_blocks", 0, 600, torch.Generator().manual_seed(1))
    target = pipe.unet.up_blocks[1].attentions[-1].transformer_blocks[-1].attn1
    assert q.shape == (2, 4, 16, 8) and q.stride() == (16 * 32, 8, 32, 1)       # (B,H,S,D) view over (B,S,H*D): hacked_attn.py:74-77
    assert target.stores[0] is q and len(target._forward_pre_hooks) == 0           # hook removed (the reference leaks one per call)
    order = list(pipe.unet.all_attn())
    idx = order.index(target)
    assert all(m.calls == 1 for m in order[:idx]) and all(m.calls == 0 for m in order[idx:])   # nothing ran past the hooked layer
    assert pipe.unet.seen_t == 999 - 600                                           # target_step indexes timesteps[...]
    trunk.extract(img, 16, "a photo", "mid_blocks", 0, 0, torch.Generator().manual_seed(1))
    assert pipe.encode_calls == 1                                                  # prompt embeddings cached per prompt
    with pytest.raises(IndexError):
        trunk.extract(img, 16, "a photo", "up_blocks", 3, 0, torch.Generator().manual_seed(1))   # up_blocks[1:] has 3 entries
    assert isinstance(hooks.StopForward(), Exception)


def test_diffsim_value_reproduces_the_reference_slice_quirk_and_extract_does_not():
    from diffsim_b200.diffsim import DiffSim

    pipe = _Pipe()
    ds = DiffSim(torch.float32, "cpu", trunk=DiffusersTrunk(pipe, "cpu", torch.float32))
    img = Image.new("RGB", (8, 8), (1, 2, 3))
    args = (img, 16, "p", "up_blocks", [0], 600)
    u = pipe.unet
    scored = u.up_blocks[1:][0].attentions[-1].transformer_blocks[-1].attn1       # diffsim/diffsim.py:143-145
    quirky = u.up_blocks[:-1][0].attentions[-1].transformer_blocks[-1].attn1      # diffsim/diffsim.py:242-244
    q, _, _ = ds.extract(*args, seed="2333", device="cpu")
    assert scored.stores[0] is q and getattr(quirky, "stores", None) is None
    qv, _, _ = ds.diffsim_value(*args, seed="2333", device="cpu")
    assert quirky.stores[0] is qv and ds.trunk.value_mode is False                  # restored afterwards
    ds2 = DiffSim(torch.float32, "cpu", trunk=DiffusersTrunk(_Pipe(), "cpu", torch.float32), compat_value_slices=False)
    q2, _, _ = ds2.diffsim_value(*args, seed="2333", device="cpu")
    assert ds2.trunk.pipe.unet.up_blocks[1].attentions[-1].transformer_blocks[-1].attn1.stores[0] is q2


class _SDXLPipe(_Pipe):
    """SDXL-shaped fake: encode_prompt by keyword returning four tensors, _get_add_time_ids, a UNet that insists on
    added_cond_kwargs, a VAE that records the dtype it is asked to encode in (diffsim/diffsim_xl.py:58-63 upcasts it)."""

    def __init__(self):
        super().__init__()
        self.vae_dtypes = []
        self.vae_float_calls = 0
        lat = types.SimpleNamespace(sample=lambda generator=None: torch.zeros(1, 4, 8, 8))

        def enc(x):
            self.vae_dtypes.append(x.dtype)
            return types.SimpleNamespace(latent_dist=lat)

        def vae_float():
            self.vae_float_calls += 1

        self.vae = types.SimpleNamespace(encode=enc, float=vae_float, config=types.SimpleNamespace(scaling_factor=0.13025))
        self.text_encoder_2 = types.SimpleNamespace(config=types.SimpleNamespace(projection_dim=1280))
        self.unet_kwargs = None
        unet = self.unet

        def call(sample, t, encoder_hidden_states=None, added_cond_kwargs=None):
            assert added_cond_kwargs is not None, "the SDXL UNet needs added_cond_kwargs"
            self.unet_kwargs = added_cond_kwargs
            return _UNet.__call__(unet, sample, t, encoder_hidden_states)

        self.unet_call = call

    def encode_prompt(self, prompt=None, prompt_2=None, device=None, num_images_per_prompt=1, do_classifier_free_guidance=True,
                      negative_prompt=None, negative_prompt_2=None):
        assert isinstance(prompt, str) and prompt_2 is None and not isinstance(device, str) or device == "cpu"
        self.encode_calls += 1
        return torch.ones(1, 77, 32), torch.zeros(1, 77, 32), torch.full((1, 1280), 2.0), torch.full((1, 1280), 3.0)

    def _get_add_time_ids(self, original_size, crops, target_size, dtype=None, text_encoder_projection_dim=None):
        assert text_encoder_projection_dim == 1280
        return torch.tensor([list(original_size + crops + target_size)], dtype=dtype)


def test_sdxl_trunk_passes_added_cond_kwargs_and_encodes_in_fp32():
    pipe = _SDXLPipe()
    pipe.unet.__class__ = type("_U", (_UNet,), {"__call__": lambda self, *a, **k: pipe.unet_call(*a, **k)})
    trunk = DiffusersTrunk(pipe, "cpu", torch.bfloat16, kind="sdxl")
    img = Image.new("RGB", (20, 12), (200, 30, 90))
    lat = trunk.encode(img, 32, torch.Generator().manual_seed(1))
    assert lat.dtype == torch.bfloat16                                                   # latents cast back to the pipeline dtype
    pipe.vae_dtypes.clear()
    pipe.vae_float_calls = 0
    trunk = DiffusersTrunk(pipe, "cpu", torch.float32, kind="sdxl")                      # (the fake UNet computes in fp32)
    q, k, v = trunk.extract(img, 32, "a photo", "up_blocks", [1, 0, 1], 600, torch.Generator().manual_seed(1))
    assert pipe.vae_dtypes == [torch.float32] and pipe.vae_float_calls == 1          # VAE upcast, latents cast back
    added = pipe.unet_kwargs
    assert set(added) == {"text_embeds", "time_ids"}
    assert added["text_embeds"].shape == (2, 1280) and added["text_embeds"][0, 0] == 3.0 and added["text_embeds"][1, 0] == 2.0
    assert added["time_ids"].tolist() == [[32, 32, 0, 0, 32, 32]] * 2                   # (original, crop, target) x [neg, pos]
    target = pipe.unet.up_blocks[:-1][1].attentions[0].transformer_blocks[1].attn1
    assert target.stores[0] is q


def test_pair_scoring_consumes_the_generator_in_the_reference_order():
    """diffsim(A, B): VAE sample A, VAE sample B, noise A, noise B from ONE generator (diffsim/diffsim.py:109-113,
    diffsim_pipeline.py:174-176); the cached path seeds each image afresh (VAE sample, noise), like diffsim_value."""
    from diffsim_b200 import diffsim as D

    calls = []

    class T(D.Trunk):
        def encode(self, image, img_size, generator):
            calls.append(("vae", image, float(torch.rand(1, generator=generator))))
            return image

        def forward(self, latents, prompt, target_block, target_layer, target_step, generator):
            calls.append(("noise", latents, float(torch.rand(1, generator=generator))))
            return None

    t = T()
    g = torch.Generator().manual_seed(2334)
    D._extract_pair(t, "A", "B", 512, "p", "up_blocks", 0, 600, g)
    assert [c[:2] for c in calls] == [("vae", "A"), ("vae", "B"), ("noise", "A"), ("noise", "B")]
    g2 = torch.Generator().manual_seed(2334)
    expect = [float(torch.rand(1, generator=g2)) for _ in range(4)]
    assert [c[2] for c in calls] == expect
    calls.clear()
    t.extract("B", 512, "p", "up_blocks", 0, 600, torch.Generator().manual_seed(2334))
    assert [c[:2] for c in calls] == [("vae", "B"), ("noise", "B")] and calls[0][2] == expect[0]   # B's VAE draw differs from pair mode


def test_scorers_refuse_to_run_without_a_trunk():
    from diffsim_b200.diffsim import DiffSim, diffsim_DiT, diffsim_xl

    for make in (lambda: DiffSim(torch.float32, "cpu"), lambda: diffsim_xl(torch.float32, "cpu"), lambda: diffsim_DiT(256, 600, "cpu")):
        with pytest.raises(ValueError, match="needs a trunk"):
            make()
    with pytest.raises(ValueError, match="ckpt"):
        diffsim_DiT(256, 600, "cpu", ckpt="DiT-XL-2-256x256.pt", trunk=DiffusersTrunk(_Pipe(), "cpu", torch.float32))
