"""Generate golden vectors from the UNMODIFIED reference module.

Run in a container where /root/reference exists (it does not on the GPU box):

    python tests/golden/make_golden.py

For each case: build the reference `HiT_SIR` (/root/reference/models/hit_sir_pro.py:1065)
through the three-symbol timm shim (oracle/ref_shim), load deterministic weights from
`oracle.weights.fill_state_dict` (numpy PCG64, independent of torch's RNG), run
`model(x)` under eval/no_grad in fp32 on CPU and store the output plus strided samples of
intermediate activations captured with forward hooks.  The committed .npz files are what
`tests/test_oracle_golden.py` pins the oracle to and what the `-m gpu` tests compare the
CUDA path against.
"""
import contextlib
import io
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle.weights import fill_state_dict, synthetic_image  # noqa: E402

with contextlib.redirect_stdout(io.StringIO()):
    from models.hit_sir_pro import HiT_SIR  # noqa: E402

PRO = dict(embed_dim=180, base_win_size=[8, 8], depths=[6] * 6, num_heads=[6] * 6, mlp_ratio=2,
           hier_win_ratios=[0.5, 1, 2, 4, 6, 8, 10, 12])

# name, (mulsizeconv, casa, fusion), upsampler, upscale, weight mode, weight seed, (B,H,W), input seed
CASES = [
    ("cfg1_pro_x4_64", (True, True, True), "nearest+conv", 4, "stress", 0, (1, 64, 64), 1234),
    ("pro_x4_init_b2_40x52", (True, True, True), "nearest+conv", 4, "init", 1, (2, 40, 52), 7),
    ("cfg5_ablation_x4_33x47", (False, False, False), "nearest+conv", 4, "stress", 2, (1, 33, 47), 8),
    ("pro_x2_pixelshuffle_48x36", (True, True, True), "pixelshuffle", 2, "stress", 3, (1, 48, 36), 9),
    ("mixed_x3_direct_35x40", (False, True, False), "pixelshuffledirect", 3, "stress", 4, (1, 35, 40), 10),
]
TAP_STRIDE = 211


def tap_modules(model):
    taps = {"shallow": model.conv_first, "norm": model.norm, "conv_after_body": model.conv_after_body}
    if isinstance(model.fusion, torch.nn.Module):
        taps["fused"] = model.fusion
    for i in (0, len(model.layers) - 1):
        for j, blk in enumerate(model.layers[i].residual_group.blocks):
            taps[f"block{i}.{j}"] = blk
    for i, layer in enumerate(model.layers):
        taps[f"layer{i}"] = layer
    return taps


def to_nhwc_flat(name, t, hw):
    if t.dim() == 4:                      # (B,C,H,W) conv-style outputs
        t = t.permute(0, 2, 3, 1)
    return t.reshape(-1)                  # (B,HW,C) token outputs are already NHWC order


def main():
    for name, flags, up, scale, mode, wseed, shape, xseed in CASES:
        with contextlib.redirect_stdout(io.StringIO()):
            m = HiT_SIR(*flags, upsampler=up, upscale=scale, **PRO).eval()
        n_params = sum(p.numel() for p in m.parameters())
        sd = fill_state_dict(m.state_dict(), wseed, mode)
        m.load_state_dict(sd, strict=True)
        x = synthetic_image(*shape, seed=xseed)
        got = {}
        hooks = []
        for tname, mod in tap_modules(m).items():
            hooks.append(mod.register_forward_hook(
                lambda _m, _i, out, tname=tname: got.__setitem__(tname, out.detach())))
        with torch.no_grad():
            y = m(x)
        for h in hooks:
            h.remove()
        arrays = {"y": y.numpy(), "n_params": np.int64(n_params), "n_keys": np.int64(len(sd)),
                  "tap_stride": np.int64(TAP_STRIDE),
                  # reference state_dict keys + shapes: pins the drop-in's parameter tree on boxes without /root/reference
                  "keys": np.array("\n".join(f"{k} {tuple(v.shape)}" for k, v in m.state_dict().items()))}
        for tname, t in got.items():
            arrays["tap_" + tname] = to_nhwc_flat(tname, t, shape[1:])[::TAP_STRIDE].numpy().copy()
        meta = dict(flags=list(flags), upsampler=up, upscale=scale, mode=mode, wseed=wseed,
                    shape=list(shape), xseed=xseed)
        arrays["meta"] = np.array(repr(meta))
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(f"{name}: params {n_params} keys {len(sd)} y {tuple(y.shape)} mean {y.mean():.6f} std {y.std():.6f}"
              f" -> {os.path.getsize(path) / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
