"""ldt_b200 -- B200 (sm_100a) implementation of the LDT sampling hot path.

Drop-in surface (same names and call signatures as the reference, Negai-98/LDT):

    Score(cfg.score)                      reference model/scorenet/score.py:47
    Compressor(cfg.compressor)            reference model/Compressor/Network.py:105  (sampling decoder)
    DiffusionVPSDE(cfg.sde)               reference diffusion/diffusion_continuous.py:626
    metrics.compute_CD_metrics(...)       reference evaluation/evaluation_metrics.py:299
    shims/StructuralLosses, shims/StructuralLossesBackend   reference evaluation/pytorch_structural_losses

All arithmetic runs in ``csrc/libldt_b200.so`` (hand-written CUDA behind the C ABI of ``include/ldt_b200.h``);
importing this package does not need a GPU, calling it does.
"""
from . import _lib  # noqa: F401
from .compressor import Compressor
from .score import Score
from .sde import DiffusionVPSDE

__all__ = ["Score", "Compressor", "DiffusionVPSDE"]
__version__ = "0.1.0"
