"""CPU oracle for the evaluation metric around the path -- TEST INFRASTRUCTURE ONLY.

Restates, in numpy fp32, what `Experiment` computes per test image
(/root/reference/experiments/experiment.py:436-463 after the clip of :746-748):

* `y_channel`: `convert_image(img, source='[0,1]', target='y-channel')`
  (/root/reference/utils/utils.py:170-186): `16/255 + (65.738 R + 129.057 G + 25.064 B) / 256`, element-wise in float32 in that order.
* `psnr`: `skimage.metrics.peak_signal_noise_ratio(hr_y, sr_y, data_range=1)`.  scikit-image is a third-party dependency that is
  not vendored in /root/reference and not installed in this image (requirements.txt pins scikit-image==0.24.0); its published
  algorithm is `10 * log10(data_range**2 / mean((a - b)**2, dtype=float64))` with the difference and square taken in the
  input precision (float32 here).

Pinning: `tests/golden/metrics_y_channel.npz` holds the output of the UNMODIFIED reference `convert_image` on seeded inputs
(generator: `tests/golden/make_golden_metrics.py`, run where /root/reference exists); `tests/test_metrics.py` checks `y_channel`
against it bit for bit.  The PSNR formula has no reference-side fixture (skimage absent): "parity unpinned" for that one line.

Only `tests/` may import this module; the product (`hitsir_b200.metrics`) never does.
"""
import numpy as np


def y_channel(img: np.ndarray) -> np.ndarray:
    """[N,3,H,W] float32 RGB in [0,1] -> [N,H,W] float32 (utils/utils.py:170-180)."""
    img = np.asarray(img, dtype=np.float32)
    f = np.float32
    t = (f(65.738) * img[:, 0] + f(129.057) * img[:, 1]) + f(25.064) * img[:, 2]
    return (f(16.0 / 255.0) + t / f(256.0)).astype(np.float32)


def mse_y(sr: np.ndarray, hr: np.ndarray, clip: bool = True) -> np.ndarray:
    """Per-image mean squared Y difference, float64 [N] (experiment.py:442-463 with the clip of :748)."""
    sr = np.asarray(sr, dtype=np.float32)
    if clip:
        sr = np.clip(sr, np.float32(0), np.float32(1))
    d = y_channel(hr) - y_channel(sr)
    return np.mean((d * d).reshape(d.shape[0], -1), axis=1, dtype=np.float64)


def psnr_y(sr: np.ndarray, hr: np.ndarray, clip: bool = True) -> np.ndarray:
    with np.errstate(divide="ignore"):
        return 10.0 * np.log10(1.0 / mse_y(sr, hr, clip))
