"""Bit-exact checks of the INDEX work of the path (-m gpu): reflect padding (hit_sir_pro.py:664-674), window partition / reverse with
the crop (:236-271, :696), nearest upsampling (:1331-1332) and PixelShuffle (:1024-1062) are integer re-indexing; north_star asks
them to be bit-exact, a relative-L2 bound on smooth images would not notice an off-by-one in a reflected strip or a swapped
sub-pixel.  Each test isolates one stage with the C-ABI test hooks (hitsir_set_inject feeds the stage white noise, hitsir_set_tap
reads it back) and plants weights under which the stage's arithmetic is exact (identity taps, counting sums of small integers), so
`torch.equal` against a torch restatement of the reference's indexing is a legitimate assertion."""
import pytest
import torch
import torch.nn.functional as F

from oracle.hitsir_oracle import from_windows, reflect_index, to_windows
from tests.helpers import build_pair

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
H, W = 56, 72                      # not a multiple of 16, 32, 48 or 64: every large window needs its reflected strip
WINDOWS = [4, 8, 16, 32, 48, 64]
MEAN = torch.tensor([0.485, 0.456, 0.4060]).view(1, 3, 1, 1)


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def noise(*shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(*shape, generator=g)


def run_tap(model, x, name, numel):
    dst = torch.full((numel,), float("nan"), device=DEV)
    model.set_tap(DEV, name, dst, stop=True)
    with torch.no_grad():
        model(x)
    torch.cuda.synchronize()
    model.set_tap(DEV, None)
    out = dst.cpu()
    assert not torch.isnan(out).any(), name
    return out


def padded(n, w):
    return -(-n // w) * w


@pytest.mark.parametrize("j", range(6))
def test_reflect_pad_and_window_tokens_are_bit_exact(j):
    """casa off: the window tokens are the reflect-padded stream rounded to bf16, position for position (:664-674, :689)."""
    model, _ = build_pair((0, 0, 0), "nearest+conv", 4, "init", 5)
    model = model.to(DEV)
    x = noise(1, 3, H, W, seed=1).to(DEV)
    src = (noise(1, H, W, 180, seed=100 + j) * 8.0 - 4.0)                 # white noise, several binades: bf16 rounding is not trivial
    model.set_inject(DEV, f"block0.{j}.in", src.reshape(-1, 180))
    w = WINDOWS[j]
    Hp, Wp = padded(H, w), padded(W, w)
    got = run_tap(model, x, f"block0.{j}.qkv", Hp * Wp * 180).view(1, Hp, Wp, 180)
    model.set_inject(DEV, None)
    want = bf16_round(src[:, reflect_index(H, Hp)][:, :, reflect_index(W, Wp)])
    assert torch.equal(got, want)


def counting_weights(sd):
    """Under these weights SCC.forward counts: k = 0 (k_generate* = 0), the pooled bias is exactly 1 (pos3.2: weight 0, bias 1; the
    mean of a constant), pooling sums (spatial_linear weight 1, bias 0).  With q = 0 and v in {0, 1} the S-SC output of a token is
    sum over ITS window of v (hit_sir_pro.py:475-511) -- small integers, exact in bf16 operands / fp32 accumulation -- and C-SC is 0."""
    for k in sd:
        if ".correlation.k_generate" in k or k.endswith("correlation.pos.pos3.2.weight") or k.endswith("correlation.spatial_linear.bias"):
            sd[k].zero_()
        elif k.endswith("correlation.pos.pos3.2.bias") or k.endswith("correlation.spatial_linear.weight"):
            sd[k].fill_(1.0)


@pytest.mark.parametrize("j", range(6))
def test_window_partition_reverse_and_crop_are_bit_exact(j):
    """Window membership, window_reverse and the crop of :696, through the real self-correlation kernels: every output token must
    carry the count of ones of ITS window of the reflect-padded map -- a token assigned to a neighbouring window, a mirrored strip
    that starts one pixel off or a tile stored at the wrong place changes integers."""
    model, _ = build_pair((0, 0, 0), "nearest+conv", 4, "init", 6, edit=counting_weights)
    model = model.to(DEV)
    x = noise(1, 3, H, W, seed=2).to(DEV)
    g = torch.Generator().manual_seed(200 + j)
    src = torch.zeros(1, H, W, 180)
    src[..., 90:] = (torch.rand(1, H, W, 90, generator=g) < 1.0 / 40.0).float()          # q = 0, v sparse in {0, 1}
    model.set_inject(DEV, f"block0.{j}.in", src.reshape(-1, 180))
    got = run_tap(model, x, f"block0.{j}.scc", H * W * 180).view(1, H, W, 180)
    model.set_inject(DEV, None)
    w = WINDOWS[j]
    Hp, Wp = padded(H, w), padded(W, w)
    vp = src[:, reflect_index(H, Hp)][:, :, reflect_index(W, Wp)][..., 90:]              # (1,Hp,Wp,90)
    win = to_windows(vp, (w, w))                                                        # window_partition (:236-248)
    cnt = win.sum(dim=1, keepdim=True).expand(-1, w * w, -1)
    want_s = from_windows(cnt.contiguous(), (w, w), 1, Hp, Wp)[:, :H, :W, :]            # window_reverse + crop
    assert want_s.max().item() <= 256.0                                                 # stays exact in the bf16 result
    assert torch.equal(got[..., :90], want_s)
    assert torch.equal(got[..., 90:], torch.zeros_like(want_s))


def identity_conv(w, rule):
    """3x3 conv weight (Co, Ci, 3, 3): out channel o = centre tap of input channel rule(o)."""
    w.zero_()
    for o in range(w.shape[0]):
        w[o, rule(o), 1, 1] = 1.0


def test_nearest_upsampling_is_bit_exact():
    """'nearest+conv' head with identity convolutions: up1 / up2 / hr must be the x2 / x4 nearest replication of the bf16 feature map
    and the image its first three channels + mean, pixel for pixel (hit_sir_pro.py:1326-1334, 1342)."""
    def edit(sd):
        identity_conv(sd["conv_before_upsample.0.weight"], lambda o: o)
        for k in ("conv_up1", "conv_up2", "conv_hr"):
            identity_conv(sd[k + ".weight"], lambda o: o)
        identity_conv(sd["conv_last.weight"], lambda o: o + 5)
        for k in ("conv_before_upsample.0", "conv_up1", "conv_up2", "conv_hr", "conv_last"):
            sd[k + ".bias"].zero_()
    model, _ = build_pair((0, 0, 0), "nearest+conv", 4, "init", 7, edit=edit)
    model = model.to(DEV)
    x = noise(1, 3, 40, 52, seed=3).to(DEV)
    h, w = 40, 52
    src = noise(1, h, w, 180, seed=300)                                                 # >= 0: LeakyReLU is the identity
    model.set_inject(DEV, "fused", src.reshape(-1, 180))
    f64 = bf16_round(src)[..., :64]
    up1 = run_tap(model, x, "up1", 4 * h * w * 64).view(1, 2 * h, 2 * w, 64)
    assert torch.equal(up1, f64.repeat_interleave(2, 1).repeat_interleave(2, 2))
    up2 = run_tap(model, x, "up2", 16 * h * w * 64).view(1, 4 * h, 4 * w, 64)
    want = f64.repeat_interleave(4, 1).repeat_interleave(4, 2)
    assert torch.equal(up2, want)
    hr = run_tap(model, x, "hr", 16 * h * w * 64).view(1, 4 * h, 4 * w, 64)
    assert torch.equal(hr, want)
    with torch.no_grad():
        y = model(x).cpu()
    model.set_inject(DEV, None)
    want_y = F.interpolate(f64[..., 5:8].permute(0, 3, 1, 2), scale_factor=4, mode="nearest") + MEAN
    assert torch.equal(y, want_y)


@pytest.mark.parametrize("scale", [2, 3, 4])
def test_pixelshuffle_is_bit_exact(scale):
    """Upsample (conv 64 -> 64 s^2 + PixelShuffle(s), hit_sir_pro.py:1024-1043): with one-tap weights that send feature channel
    (co % 64) to conv output co, every output sub-pixel (i, j) of every channel shows a DIFFERENT input channel, so a swapped or
    transposed sub-pixel cannot hide."""
    stages = 1 if scale == 3 else {2: 1, 4: 2}[scale]

    def edit(sd):
        identity_conv(sd["conv_before_upsample.0.weight"], lambda o: o)
        sd["conv_before_upsample.0.bias"].zero_()
        for s in range(stages):
            identity_conv(sd[f"upsample.{2 * s}.weight"], lambda o: (o * 5 + 3) % 64)
            sd[f"upsample.{2 * s}.bias"].zero_()
        identity_conv(sd["conv_last.weight"], lambda o: 7 * o + 2)
        sd["conv_last.bias"].zero_()
    model, _ = build_pair((0, 0, 0), "pixelshuffle", scale, "init", 8, edit=edit)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    model = model.to(DEV)
    h, w = 36, 44
    x = noise(1, 3, h, w, seed=4).to(DEV)
    src = noise(1, h, w, 180, seed=400 + scale)
    model.set_inject(DEV, "fused", src.reshape(-1, 180))
    with torch.no_grad():
        y = model(x).cpu()
    model.set_inject(DEV, None)
    t = bf16_round(src)[..., :64].permute(0, 3, 1, 2)                                    # conv_before_upsample = identity on >= 0
    for s in range(stages):
        t = F.pixel_shuffle(F.conv2d(t, sd[f"upsample.{2 * s}.weight"], None, 1, 1), 3 if scale == 3 else 2)
    want = F.conv2d(t, sd["conv_last.weight"], None, 1, 1) + MEAN
    assert y.shape == (1, 3, scale * h, scale * w)
    assert torch.equal(y, want)


@pytest.mark.parametrize("scale", [2, 3, 4])
def test_pixelshuffledirect_is_bit_exact(scale):
    """UpsampleOneStep (conv C -> 3 s^2 + PixelShuffle(s), hit_sir_pro.py:1046-1062, 1320-1325)."""
    def edit(sd):
        identity_conv(sd["upsample.0.weight"], lambda o: (o * 11 + 4) % 180)
        sd["upsample.0.bias"].zero_()
    model, _ = build_pair((0, 0, 0), "pixelshuffledirect", scale, "init", 9, edit=edit)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    model = model.to(DEV)
    h, w = 35, 41
    x = noise(1, 3, h, w, seed=5).to(DEV)
    src = noise(1, h, w, 180, seed=500 + scale) * 2.0 - 1.0
    model.set_inject(DEV, "fused", src.reshape(-1, 180))
    with torch.no_grad():
        y = model(x).cpu()
    model.set_inject(DEV, None)
    t = bf16_round(src).permute(0, 3, 1, 2)
    want = F.pixel_shuffle(F.conv2d(t, sd["upsample.0.weight"], None, 1, 1), scale) + MEAN
    assert torch.equal(y, want)
