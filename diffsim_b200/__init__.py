"""diffsim_b200 -- B200-native (sm_100a) implementation of DiffSim's Aligned Attention Score hot path.

The compute lives in libdiffsim_b200.so (hand-written CUDA: tcgen05 / TMEM / TMA attention, vectorised
reductions, tensor-core similarity GEMM) behind the C ABI of include/diffsim_b200.h; this package is the
Python host that mirrors the reference's call surface (DiffSim.diffsim / diffsim_value, diffsim_xl.diffsim_score,
diffsim_DiT.diffsim_score, the hook / processor contracts and the --target_* / --similarity flags); the trunk
(DiffSimPipeline.step: one noised UNet forward) stays on PyTorch behind diffsim.Trunk.
"""
__version__ = "0.1.0"

from . import _native  # noqa: F401  (does not load the shared library until first use)
