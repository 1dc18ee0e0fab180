"""NeRF module behind the reference's interface (noisy_src/model.py:20-221).

Same class names, constructor arguments, attribute / state_dict keys and shapes, so reference
checkpoints load here and ours load there.  `forward` runs the sm_100a path:
encode (PE) -> 10 tcgen05 GEMM launches -> fp32 heads; backward is the matching kernel chain.
Only the reference's default architecture is implemented (hidden 256, 8 layers, skip (4,),
L = 10 / 4, view directions on); anything else raises NotImplementedError -- there is no fallback.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn as nn

from . import ops
from .config import ModelConfig


class PositionalEncoding(nn.Module):
    """[x, sin(2^k x), cos(2^k x)]_k -- no pi factor (noisy_src/model.py:43,76-78)."""

    def __init__(self, num_freqs: int, include_input: bool = True, log_sampling: bool = True) -> None:
        super().__init__()
        if not include_input or not log_sampling:
            raise NotImplementedError("only include_input=True, log_sampling=True (the reference's only use)")
        self.num_freqs = num_freqs
        self.include_input = include_input
        self.register_buffer("freq_bands", 2.0 ** torch.linspace(0.0, num_freqs - 1, num_freqs))

    @property
    def output_dim(self) -> int:
        return 2 * self.num_freqs + (1 if self.include_input else 0)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.PosEnc.apply(x, self.num_freqs)


def _check_config(config) -> None:
    ok = (config.pos_freqs == 10 and config.dir_freqs == 4 and config.hidden_dim == 256
          and config.num_hidden_layers == 8 and tuple(config.skips) == (4,) and config.use_view_dirs)
    if not ok:
        raise NotImplementedError(
            "the sm_100a kernels are specialised for the reference's default ModelConfig "
            "(pos_freqs=10, dir_freqs=4, hidden_dim=256, num_hidden_layers=8, skips=(4,), use_view_dirs=True)")


class NeRF(nn.Module):
    def __init__(self, config: ModelConfig | None = None) -> None:
        super().__init__()
        if config is None:
            config = ModelConfig()
        _check_config(config)
        self.config = config
        self.pos_encoder = PositionalEncoding(config.pos_freqs)
        self.dir_encoder = PositionalEncoding(config.dir_freqs)
        pos_dim = 3 * self.pos_encoder.output_dim
        dir_dim = 3 * self.dir_encoder.output_dim
        # identical construction order to the reference => identical nn.Linear default init under a seed
        self.pts_linears = nn.ModuleList()
        in_dim = pos_dim
        for i in range(config.num_hidden_layers):
            self.pts_linears.append(nn.Linear(in_dim, config.hidden_dim))
            in_dim = config.hidden_dim
            if i in config.skips:
                in_dim += pos_dim
        self.sigma_linear = nn.Linear(config.hidden_dim, 1)
        self.feature_linear = nn.Linear(config.hidden_dim, config.hidden_dim)
        self.dir_linear = nn.Linear(config.hidden_dim + dir_dim, config.hidden_dim // 2)
        self.rgb_linear = nn.Linear(config.hidden_dim // 2, 3)
        self._packed = ops.PackedWeights()

    def kernel_params(self):
        """The 24 parameter tensors in state_dict order (the order the C ABI expects)."""
        ps = []
        for lin in self.pts_linears:
            ps += [lin.weight, lin.bias]
        for lin in (self.sigma_linear, self.feature_linear, self.dir_linear, self.rgb_linear):
            ps += [lin.weight, lin.bias]
        return ps

    def forward_raw(self, x: torch.Tensor, d: torch.Tensor, group: int = 1) -> torch.Tensor:
        """(rgb pre-sigmoid, sigma pre-ReLU) [M,4]; d has one row per `group` consecutive points."""
        return ops.NeRFMLP.apply(x, d, group, self._packed, torch.is_grad_enabled(), *self.kernel_params())

    def forward(self, x: torch.Tensor, d: torch.Tensor | None = None) -> Tuple[torch.Tensor, torch.Tensor]:
        if d is None:
            # the reference crashes here too (283-wide dir_linear fed 256 features, model.py:187-193)
            raise RuntimeError("NeRF.forward needs view directions d when use_view_dirs=True")
        shp = x.shape[:-1]
        raw = self.forward_raw(x.reshape(-1, 3), d.reshape(-1, 3), 1)
        rgb, sigma = ops.HeadAct.apply(raw)
        return rgb.reshape(*shp, 3), sigma.reshape(*shp, 1)


def create_nerf(config: ModelConfig | None = None) -> Tuple[NeRF, NeRF | None]:
    if config is None:
        config = ModelConfig()
    return NeRF(config), NeRF(config)
