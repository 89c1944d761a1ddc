"""Measured parity errors of the CUDA path (GPU box only): the numbers tests/helpers.py turns into per-tap bounds.

    python tools/measure_taps.py > gpurun_out/taps.json

* the taps of tests/test_gpu_parity.py::test_cuda_taps_match_oracle (stress weights, 1x56x72): rel-L2 per tap vs the CPU oracle;
* S-SC / C-SC halves of every window size against the reference golden (scc_parts_56x72);
* end-to-end max-abs / PSNR of every golden case and of the BASELINE-size fixtures.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.weights import synthetic_image  # noqa: E402
from tests.helpers import GOLDEN_CASES, TAP_NAMES, build_pair, load_golden, psnr, rel_l2  # noqa: E402
from tests.test_oracle_golden import BIG_CASES, check_big  # noqa: E402

DEV = "cuda:0"


def main():
    out = {"taps": {}, "scc_parts": {}, "golden": {}, "big": {}}
    model, oracle = build_pair((1, 1, 1), "nearest+conv", 4, "stress", 21)
    x = synthetic_image(1, 56, 72, seed=4)
    taps = {}
    with torch.no_grad():
        oracle.forward(x, taps)
    model = model.to(DEV)
    xd = x.to(DEV)
    for n in TAP_NAMES:
        ref = taps[n].contiguous()
        dst = torch.full((ref.numel(),), float("nan"), device=DEV)
        model.set_tap(DEV, n, dst, stop=True)
        with torch.no_grad():
            model(xd)
        torch.cuda.synchronize()
        out["taps"][n] = rel_l2(dst.cpu().view(ref.shape), ref)
    g, meta = load_golden("scc_parts_56x72")
    for j in range(6):
        dst = torch.full((56 * 72 * 180,), float("nan"), device=DEV)
        model.set_tap(DEV, f"block0.{j}.scc", dst, stop=True)
        with torch.no_grad():
            model(xd)
        torch.cuda.synchronize()
        got = dst.cpu().view(1, 56, 72, 180)
        for kind, part in (("ssc", got[..., :90]), ("csc", got[..., 90:])):
            out["scc_parts"][f"{kind}{j}"] = rel_l2(part.reshape(-1)[::meta["s_stride"]], torch.from_numpy(g[f"{kind}{j}"]))
    model.set_tap(DEV, None)
    del model
    for name in GOLDEN_CASES:
        g, meta = load_golden(name)
        model, _ = build_pair(meta["flags"], meta["upsampler"], meta["upscale"], meta["mode"], meta["wseed"])
        x = synthetic_image(*meta["shape"], seed=meta["xseed"])
        with torch.no_grad():
            y = model.to(DEV)(x.to(DEV)).cpu()
        ref = torch.from_numpy(g["y"])
        scale = max(1.0, ref.abs().max().item())
        out["golden"][name] = {"mode": meta["mode"], "max_abs": (y - ref).abs().max().item() / scale, "psnr": psnr(y / scale, ref / scale)}
        del model
    for name in BIG_CASES:
        g, meta = load_golden(name)
        model, _ = build_pair(meta["flags"], meta["upsampler"], meta["upscale"], meta["mode"], meta["wseed"])
        x = synthetic_image(*meta["shape"], seed=meta["xseed"])
        with torch.no_grad():
            y = model.to(DEV)(x.to(DEV)).cpu()
        err, p = check_big(y, g, 1.0)
        out["big"][name] = {"mode": meta["mode"], "max_abs": err, "psnr": p}
        del model
        torch.cuda.empty_cache()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
