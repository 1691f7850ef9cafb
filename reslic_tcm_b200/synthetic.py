"""Synthetic latents of the BASELINE.json config shapes.

Per-image generators (seed = 100_000*cfg + image index) make every image's tensors
identical no matter how the batch is sharded across ranks: rank r materialises only its
images and a 1-GPU run sees the same data as an N-GPU run.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch

M_LATENT = 320      # y channels (tcm.py:311,331)
Z_CHANNELS = 192    # z channels (tcm.py:416)
NUM_SLICES = 5      # tcm.py num_slices -> 64 channels per slice


@dataclass(frozen=True)
class Config:
    cfg: int
    name: str
    batch: int
    height: int          # padded image height
    width: int
    training: bool
    with_indexes: bool

    @property
    def y_hw(self) -> Tuple[int, int]:
        return self.height // 16, self.width // 16

    @property
    def z_hw(self) -> Tuple[int, int]:
        return self.height // 64, self.width // 64

    @property
    def y_elems_per_image(self) -> int:
        return M_LATENT * self.y_hw[0] * self.y_hw[1]

    @property
    def z_elems_per_image(self) -> int:
        return Z_CHANNELS * self.z_hw[0] * self.z_hw[1]

    @property
    def num_pixels_per_image(self) -> int:
        return self.height * self.width


# BASELINE.json "configs", in order (cfg numbers are 1-based as in SURVEY.md §8)
CONFIGS: Dict[int, Config] = {
    1: Config(1, "tcm64_256x256_b1_forward", 1, 256, 256, False, False),
    2: Config(2, "tcm64_kodak768x512_b24_eval_round_indexes", 24, 768, 512, False, True),
    3: Config(3, "tcm128_kodak768x512_b64_likelihood_bpp", 64, 768, 512, False, False),
    4: Config(4, "tcm128_clic2048x1408_b16_compress_symbols_indexes", 16, 2048, 1408, False, True),
    5: Config(5, "tcm64_train256x256_b256_noise", 256, 256, 256, True, False),
}


def scale_table(device=None) -> torch.Tensor:
    """exp(linspace(ln 0.11, ln 256, 64)) — tcm.py:26-34."""
    t = torch.exp(torch.linspace(math.log(0.11), math.log(256), 64))
    return t.to(device) if device is not None else t


def make_image(cfg: int, image_index: int, y_hw: Tuple[int, int], z_hw: Tuple[int, int],
               with_noise: bool = False) -> Dict[str, torch.Tensor]:
    """One image's (y, mu, sigma, z[, noise]) on CPU, fp32."""
    g = torch.Generator().manual_seed(100_000 * cfg + image_index)
    h, w = y_hw
    shape = (M_LATENT, h, w)
    mu = torch.randn(shape, generator=g)
    # log-uniform scales: exercises the 0.11 bound and ~58 of the 64 table bins, + 1 % tail
    sigma = torch.exp(torch.empty(shape).uniform_(math.log(0.05), math.log(64.0), generator=g))
    tail = torch.rand(shape, generator=g) < 0.01
    sigma = torch.where(tail, torch.empty(shape).uniform_(64.0, 300.0, generator=g), sigma)
    y = mu + sigma * torch.randn(shape, generator=g)
    z = 2.0 * torch.randn((Z_CHANNELS, z_hw[0], z_hw[1]), generator=g)
    out = {"y": y, "mu": mu, "sigma": sigma, "z": z}
    if with_noise:
        out["noise_y"] = torch.empty(shape).uniform_(-0.5, 0.5, generator=g)
        out["noise_z"] = torch.empty(z.shape).uniform_(-0.5, 0.5, generator=g)
    return out


def make_batch(cfg: int, images: range, y_hw: Optional[Tuple[int, int]] = None,
               z_hw: Optional[Tuple[int, int]] = None, with_noise: bool = False,
               pin: bool = False) -> Dict[str, torch.Tensor]:
    """Stack images [images.start, images.stop) of config `cfg` into NCHW host tensors."""
    c = CONFIGS[cfg]
    y_hw = y_hw or c.y_hw
    z_hw = z_hw or c.z_hw
    items = [make_image(cfg, i, y_hw, z_hw, with_noise) for i in images]
    out = {k: torch.stack([it[k] for it in items]) for k in items[0]}
    if pin:
        out = {k: v.pin_memory() for k, v in out.items()}
    return out


def eb_parameters(channels: int = Z_CHANNELS, trained_like: bool = True, seed: int = 1234
                  ) -> Dict[str, torch.Tensor]:
    """EntropyBottleneck parameters: CompressAI init
    (adaptive_entropy_bottleneck.py:341-362), optionally perturbed so that the tanh terms
    and the medians are live ("trained-like")."""
    g = torch.Generator().manual_seed(seed)
    filters = (1, 3, 3, 3, 3, 1)
    scale = 10.0 ** (1 / 5)
    p: Dict[str, torch.Tensor] = {}
    for i in range(5):
        init = math.log(math.expm1(1 / scale / filters[i + 1]))
        p[f"_matrix{i}"] = torch.full((channels, filters[i + 1], filters[i]), init)
        p[f"_bias{i}"] = torch.empty(channels, filters[i + 1], 1).uniform_(-0.5, 0.5, generator=g)
        if i < 4:
            p[f"_factor{i}"] = torch.zeros(channels, filters[i + 1], 1)
    q = torch.tensor([-10.0, 0.0, 10.0]).repeat(channels, 1, 1)
    if trained_like:
        for i in range(5):
            p[f"_matrix{i}"] = p[f"_matrix{i}"] + 0.3 * torch.randn(p[f"_matrix{i}"].shape, generator=g)
        for i in range(4):
            p[f"_factor{i}"] = 0.5 * torch.randn(p[f"_factor{i}"].shape, generator=g)
        q[:, 0, 1] = torch.randn(channels, generator=g)
    p["quantiles"] = q
    return p


def load_eb_parameters(module, params: Dict[str, torch.Tensor]) -> None:
    with torch.no_grad():
        for k, v in params.items():
            getattr(module, k).copy_(v.to(getattr(module, k).device))
