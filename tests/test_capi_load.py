"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/hitsir_b200.h declares
(no compute calls without a GPU), and the product has no silent fallback."""
import ctypes
import os
import re

import pytest
import torch

import hitsir_b200
from hitsir_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    header = open(os.path.join(ROOT, "include", "hitsir_b200.h")).read()
    declared = set(re.findall(r"HITSIR_API\s+[\w\s\*]+?\b(hitsir_\w+)\s*\(", header))
    assert declared == set(_capi.SYMBOLS), declared ^ set(_capi.SYMBOLS)
    lib = _capi.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.hitsir_version()


def test_config_struct_layout_matches_header():
    # 6 int32 + 16 + 16 int32 + 2 int32 + float + int32 + float + int32 + int32 + 16 float
    assert ctypes.sizeof(_capi.HitsirConfig) == 4 * (6 + 16 + 16 + 2 + 1 + 1 + 1 + 1 + 1 + 16 + 2)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    m = hitsir_b200.HiT_SIR(False, False, False, **hitsir_b200.PRO_KWARGS)
    cfg = m._config()
    out = ctypes.c_void_p()
    status = _capi.load().hitsir_create(ctypes.byref(cfg), ctypes.byref(out))
    assert status != 0 and not out
    assert "CPU" in _capi.last_error() or "device" in _capi.last_error()


def test_module_refuses_cpu_and_unsupported_configs():
    m = hitsir_b200.HiT_SIR(True, True, True, **hitsir_b200.PRO_KWARGS).eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.rand(1, 3, 64, 64))
    with pytest.raises(NotImplementedError):
        hitsir_b200.HiT_SIR(True, True, True)                      # reference default embed_dim=60 is not the pro width
    kw = dict(hitsir_b200.PRO_KWARGS)
    with pytest.raises(AssertionError, match="only support x4 now"):   # hit_sir_pro.py:1248
        hitsir_b200.HiT_SIR(True, True, True, **{**kw, "upscale": 2})
    with pytest.raises(ValueError, match="not supported"):             # hit_sir_pro.py:1042
        hitsir_b200.HiT_SIR(True, True, True, **{**kw, "upsampler": "pixelshuffle", "upscale": 5})
    with pytest.raises(NotImplementedError):
        hitsir_b200.HiT_SIR(True, True, True, **{**kw, "patch_norm": False})


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "single-image-super-resolution-application_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.replace("oracle tap", ""), os.path.join(dirpath, f)


def test_load_checkpoint_in_reference_format(tmp_path):
    """experiments/experiment.py:257-263 writes {'start_epoch', 'model', 'optimizer'}; :222-223 / test_experiment.py:43-44 read it."""
    kw = dict(hitsir_b200.PRO_KWARGS)
    kw.update(depths=[2], num_heads=[6])
    src = hitsir_b200.HiT_SIR(True, True, True, **kw)
    sd = {k: torch.randn_like(v) for k, v in src.state_dict().items()}
    path = tmp_path / "new_epoch_model.pth"
    torch.save({"start_epoch": 17, "model": sd, "optimizer": {"state": {}, "param_groups": []}}, path)
    dst = hitsir_b200.HiT_SIR(True, True, True, **kw)
    assert dst.load_checkpoint(str(path)) == 17
    for k, v in dst.state_dict().items():
        assert torch.equal(v, sd[k]), k
    assert dst.load_checkpoint(sd) is None                       # a bare state_dict works too
    bad = dict(sd); bad.pop(next(iter(bad)))
    with pytest.raises(RuntimeError):                            # strict, like the reference's load_state_dict
        dst.load_checkpoint({"model": bad})
