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
