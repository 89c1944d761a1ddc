"""Deterministic synthetic weights for HiT-SIR state_dicts -- TEST INFRASTRUCTURE ONLY.

Parity tests must not depend on torch's RNG stream (it can change between torch
builds), so every tensor is filled from a numpy PCG64 stream seeded by
(seed, crc32(key)).  Two regimes:

* ``"init"``    - the statistics of the reference's own initialisation
  (hit_sir_pro.py:1267-1274: Linear ~ trunc_normal(std .02), bias 0, LayerNorm 1/0;
  Conv2d keeps torch's default kaiming-uniform(a=sqrt(5)) -> U(-1/sqrt(fan_in), +)).
* ``"stress"``  - every parameter matters: Linear/Conv ~ U(-1,1)*g/sqrt(fan_in), non-zero
  biases, LayerNorm gamma ~ 1 +- 0.3 and beta ~ +-0.2.  At default init zeroing the whole
  spatial self-correlation moves the output by only 5.7e-3, so a broken
  attention kernel would hide inside the tolerance; under "stress" it cannot.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict

import numpy as np
import torch


def _rng(seed: int, key: str) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([seed, zlib.crc32(key.encode())]))


def fill_state_dict(sd: Dict[str, torch.Tensor], seed: int = 0, mode: str = "stress") -> Dict[str, torch.Tensor]:
    """Return a new state_dict with the same keys/shapes as ``sd`` and deterministic fp32 values."""
    out = {}
    for key, ref in sd.items():
        shape = tuple(ref.shape)
        g = _rng(seed, key)
        is_bias = key.endswith(".bias")
        is_norm = (".norm" in key or key.startswith("norm.") or
                   ".pos1.0." in key or ".pos2.0." in key or ".pos3.0." in key)
        if is_norm:
            if mode == "init":
                a = np.zeros(shape) if is_bias else np.ones(shape)
            else:
                a = g.uniform(-0.2, 0.2, shape) if is_bias else 1.0 + g.uniform(-0.3, 0.3, shape)
        elif is_bias:
            if mode == "init":
                # Linear bias 0 (hit_sir_pro.py:1270-1271); Conv2d bias U(+-1/sqrt(fan_in)) -> fan_in unknown
                # from the bias alone, use a small fixed range
                a = np.zeros(shape) if len(shape) == 1 and _is_linear_bias(key) else g.uniform(-0.05, 0.05, shape)
            else:
                a = g.uniform(-0.1, 0.1, shape)
        else:
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
            if mode == "init":
                if len(shape) == 2:
                    a = np.clip(g.normal(0.0, 0.02, shape), -2.0, 2.0)
                else:
                    bound = 1.0 / math.sqrt(fan_in)
                    a = g.uniform(-bound, bound, shape)
            else:
                gain = 1.5 if len(shape) == 2 else 1.2
                a = g.uniform(-1.0, 1.0, shape) * gain / math.sqrt(fan_in)
        out[key] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).reshape(shape)
    return out


_LINEAR_TAILS = ("proj.bias", "fc1.bias", "fc2.bias", "spatial_linear.bias", "k_generate1.bias",
                 "k_generate2.bias", "pos_proj.bias", ".2.bias", "_first.bias", "_second.bias")


def _is_linear_bias(key: str) -> bool:
    return key.endswith(_LINEAR_TAILS)


def synthetic_image(batch: int, height: int, width: int, seed: int = 1234, chans: int = 3) -> torch.Tensor:
    """Deterministic LR batch in [0,1): low-frequency structure + uniform noise (numpy PCG64)."""
    g = np.random.Generator(np.random.PCG64([seed, batch, height, width]))
    yy, xx = np.meshgrid(np.linspace(0, 1, height), np.linspace(0, 1, width), indexing="ij")
    img = np.empty((batch, chans, height, width), dtype=np.float64)
    for b in range(batch):
        for c in range(chans):
            f1, f2, ph = g.uniform(1, 6), g.uniform(1, 6), g.uniform(0, 6.28)
            img[b, c] = 0.5 + 0.25 * np.sin(6.28 * (f1 * yy + f2 * xx) + ph)
    img += g.uniform(-0.25, 0.25, img.shape)
    return torch.from_numpy(np.clip(img, 0.0, 0.999).astype(np.float32))
