"""CPU tests: pin the oracle to the reference's golden vectors, and the drop-in's parameter tree to the
reference's state_dict.  (Parity of the CUDA path itself is in test_gpu_parity.py, -m gpu.)"""
import os

import numpy as np
import pytest
import torch

import hitsir_b200
from oracle.weights import synthetic_image
from tests.helpers import GOLDEN_CASES, build_pair, load_golden


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_reference_golden(name):
    g, meta = load_golden(name)
    model, oracle = build_pair(meta["flags"], meta["upsampler"], meta["upscale"], meta["mode"], meta["wseed"])
    # parameter tree of the drop-in == reference state_dict (names, order, shapes, count)
    mine = "\n".join(f"{k} {tuple(v.shape)}" for k, v in model.state_dict().items())
    assert mine == str(g["keys"])
    assert sum(p.numel() for p in model.parameters()) == int(g["n_params"])
    assert len(list(model.buffers())) == 0
    x = synthetic_image(*meta["shape"], seed=meta["xseed"])
    taps = {}
    with torch.no_grad():
        y = oracle.forward(x, taps)
    ref = torch.from_numpy(g["y"])
    assert y.shape == ref.shape
    # fp32 restatement vs fp32 reference: only summation-order noise
    assert (y - ref).abs().max().item() < 2e-5
    stride = int(g["tap_stride"])
    checked = 0
    for key in g.files:
        if not key.startswith("tap_") or key == "tap_stride":
            continue
        t = taps[key[4:]].reshape(-1)[::stride]
        r = torch.from_numpy(g[key])
        assert t.shape == r.shape, key
        scale = r.abs().max().item() + 1e-6
        assert (t - r).abs().max().item() / scale < 1e-4, key
        checked += 1
    assert checked >= 15


def test_known_answer_param_count():
    # logs/.../模型参数量.txt:1 of the reference: the one known-answer value that pins the architecture
    m = hitsir_b200.HiT_SIR(True, True, True, **hitsir_b200.PRO_KWARGS)
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 10220014
    assert len(m.state_dict()) == 1650


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference tree not present on this box")
def test_oracle_matches_live_reference():
    import contextlib, io, sys, warnings
    warnings.filterwarnings("ignore")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "oracle", "ref_shim"))
    sys.path.insert(0, "/root/reference")
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            from models.hit_sir_pro import HiT_SIR as RefHiT
            ref = RefHiT(True, True, True, **hitsir_b200.PRO_KWARGS).eval()
        model, oracle = build_pair((1, 1, 1), "nearest+conv", 4, "stress", 11)
        ref.load_state_dict(model.state_dict(), strict=True)
        x = synthetic_image(1, 37, 45, seed=5)          # needs reflect padding for every window size
        with torch.no_grad():
            assert (ref(x) - oracle(x)).abs().max().item() < 2e-5
    finally:
        sys.path.remove("/root/reference")
        sys.modules.pop("models.hit_sir_pro", None)
        sys.modules.pop("models", None)


def test_oracle_edge_cases():
    model, oracle = build_pair((0, 0, 0), "pixelshuffledirect", 2, "init", 3)
    with pytest.raises(RuntimeError):               # 32x32 cannot be reflect-padded to window 64 (hit_sir_pro.py:672)
        oracle(torch.rand(1, 3, 32, 32))
    y = oracle(torch.rand(1, 3, 33, 33))
    assert y.shape == (1, 3, 66, 66)
    with pytest.raises(AssertionError):
        from oracle.hitsir_oracle import HiTSIROracle, OracleConfig
        HiTSIROracle(model.state_dict(), OracleConfig(False, False, False, upscale=2, upsampler="nearest+conv"))
