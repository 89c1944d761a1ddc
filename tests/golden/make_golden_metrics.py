"""Golden vectors for the evaluation-metric oracle, from the UNMODIFIED reference `convert_image`.

Run where /root/reference exists (it does not on the GPU box):

    python tests/golden/make_golden_metrics.py

Seeded RGB batches (values slightly outside [0,1] so that the clip matters for the SR side) go through
`utils.utils.convert_image(img, source='[0,1]', target='y-channel', ...)` (/root/reference/utils/utils.py:170-186) exactly as
`experiments/experiment.py:442-455` calls it; inputs and outputs are stored in tests/golden/metrics_y_channel.npz.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")

from utils.utils import convert_image  # noqa: E402


def main():
    rng = np.random.Generator(np.random.PCG64(20261018))
    hr = rng.random((3, 3, 37, 53), dtype=np.float32)
    sr = (hr + rng.normal(0.0, 0.05, hr.shape).astype(np.float32)).astype(np.float32)      # leaves [0,1] in places
    out = {}
    for name, img in (("hr", hr), ("sr_clipped", np.clip(sr, 0, 1).astype(np.float32))):
        y = convert_image(torch.from_numpy(img), source="[0,1]", target="y-channel", is_lr=False, is_lr_amplify=False, scaling_factor=4)
        out["y_" + name] = y.numpy()
    np.savez_compressed(os.path.join(HERE, "metrics_y_channel.npz"), hr=hr, sr=sr, **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
