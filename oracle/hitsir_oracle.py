"""CPU oracle for the HiT-SIR-pro forward pass -- TEST INFRASTRUCTURE ONLY.

A plain fp32 PyTorch *restatement* of `HiT_SIR.forward`
(/root/reference/models/hit_sir_pro.py:1304-1344) driven by a reference
``state_dict``.  It is written from the maths of the reference (every function
cites the reference lines it follows) in the token-major formulation the CUDA
path uses (NHWC tokens, pooled relative-position bias precomputed from the
weights, reflect padding by index map), so each CUDA stage has an oracle tap
with the same name.

Pinning: `tests/test_oracle_golden.py` checks this file against golden vectors
produced by the UNMODIFIED reference module (generator:
`tests/golden/make_golden.py`, run where /root/reference exists) and, when
/root/reference is present, against the live reference module itself.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
``--impl reference`` legs may import this module.  The product package
(`hitsir_b200`) never does and has no CPU fallback.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

RGB_MEAN = (0.485, 0.456, 0.4060)  # hit_sir_pro.py:1128


@dataclass
class OracleConfig:
    """Mirror of the constructor arguments that change the forward
    (hit_sir_pro.py:1091-1120)."""
    is_mult_size_conv_feat_extract: bool = True
    is_channel_spatial_attn: bool = True
    is_fusion: bool = True
    in_chans: int = 3
    embed_dim: int = 180
    depths: Sequence[int] = (6, 6, 6, 6, 6, 6)
    num_heads: Sequence[int] = (6, 6, 6, 6, 6, 6)
    base_win_size: Sequence[int] = (8, 8)
    mlp_ratio: float = 2.0
    upscale: int = 4
    img_range: float = 1.0
    upsampler: str = "nearest+conv"
    hier_win_ratios: Sequence[float] = (0.5, 1, 2, 4, 6, 8, 10, 12)

    def windows(self, depth: int) -> List[Tuple[int, int]]:
        # BasicLayer.__init__, hit_sir_pro.py:802-817: only the first `depth` ratios are used
        return [(int(self.base_win_size[0] * r), int(self.base_win_size[1] * r))
                for r in list(self.hier_win_ratios)[:depth]]


# --------------------------------------------------------------------------
# weight-only precomputation
# --------------------------------------------------------------------------
def dynamic_pos_bias_table(sd: Dict[str, Tensor], prefix: str, win: Tuple[int, int]) -> Tensor:
    """DynamicPosBias MLP evaluated on every relative offset
    (hit_sir_pro.py:274-313 with residual=False, called at :479-483).
    Returns ((2wh-1)*(2ww-1), heads)."""
    wh, ww = win
    oh = torch.arange(1 - wh, wh)
    ow = torch.arange(1 - ww, ww)
    grid = torch.stack(torch.meshgrid(oh, ow, indexing="ij"))       # (2, 2wh-1, 2ww-1)
    off = grid.flatten(1).transpose(0, 1).contiguous().float()     # (n, 2)
    p = prefix + "pos."
    x = F.linear(off.to(sd[p + "pos_proj.weight"].device), sd[p + "pos_proj.weight"], sd[p + "pos_proj.bias"])
    for name in ("pos1", "pos2", "pos3"):
        d = x.shape[-1]
        x = F.layer_norm(x, (d,), sd[p + name + ".0.weight"], sd[p + name + ".0.bias"], 1e-5)
        x = F.relu(x)
        x = F.linear(x, sd[p + name + ".2.weight"], sd[p + name + ".2.bias"])
    return x


def pooled_rpe_bias(tbl: Tensor, win: Tuple[int, int], base: Tuple[int, int]) -> Tensor:
    """Relative-position bias gathered per (token l, token m) and averaged over
    each r x r pooling cell (hit_sir_pro.py:486-501).  Returns (heads, L, Lb).

    bias[h, l, cell] = mean_{m in cell} tbl[(yl-ym+wh-1)*(2ww-1) + (xl-xm+ww-1), h]
    Computed without materialising the L*L gather (separable index arithmetic)."""
    wh, ww = win
    bh, bw = base
    rh, rw = wh // bh, ww // bw
    heads = tbl.shape[1]
    t = tbl.view(2 * wh - 1, 2 * ww - 1, heads)
    yl = torch.arange(wh).view(wh, 1, 1)
    cy = torch.arange(bh).view(1, bh, 1)
    iy = torch.arange(rh).view(1, 1, rh)
    dy = yl - (cy * rh + iy) + wh - 1                                # (wh, bh, rh)
    xl = torch.arange(ww).view(ww, 1, 1)
    cx = torch.arange(bw).view(1, bw, 1)
    ix = torch.arange(rw).view(1, 1, rw)
    dx = xl - (cx * rw + ix) + ww - 1                                # (ww, bw, rw)
    # sum over (iy, ix) in float64 then divide -> matches the reference's fp32 mean to ~1e-7
    g = t.double()[dy.view(wh, 1, bh, 1, rh, 1), dx.view(1, ww, 1, bw, 1, rw)]  # (wh,ww,bh,bw,rh,rw,heads)
    g = g.mean(dim=(4, 5))                                           # (wh, ww, bh, bw, heads)
    return g.permute(4, 0, 1, 2, 3).reshape(heads, wh * ww, bh * bw).float().contiguous()


# --------------------------------------------------------------------------
# building blocks (token-major: tensors are (B, H, W, C) unless noted)
# --------------------------------------------------------------------------
def conv_nhwc(x: Tensor, w: Tensor, b: Optional[Tensor], pad: int) -> Tensor:
    """nn.Conv2d(stride 1, zero padding) on an NHWC tensor."""
    return F.conv2d(x.permute(0, 3, 1, 2), w, b, 1, pad).permute(0, 2, 3, 1)


def multi_size_conv_extract(x_nchw: Tensor, sd, p="conv_first.") -> Tensor:
    """MultipleSizeConvExtract.forward, hit_sir_pro.py:67-100 (self.norm is never applied).
    Returns NHWC."""
    x1 = F.conv2d(x_nchw, sd[p + "conv_x.weight"], sd[p + "conv_x.bias"])
    outs = []
    for k in (3, 5, 7, 9):
        xk = F.conv2d(x_nchw, sd[p + f"conv{k}.weight"], sd[p + f"conv{k}.bias"], 1, k // 2)
        outs.append(xk * torch.sigmoid(x1 * xk) + xk)
    y = F.conv2d(torch.cat(outs, 1), sd[p + "conv_last.weight"], sd[p + "conv_last.bias"])
    return y.permute(0, 2, 3, 1).contiguous()


def reflect_index(n: int, n_pad: int) -> Tensor:
    """Source index of F.pad(..., 'reflect') on the bottom/right side
    (hit_sir_pro.py:664-674): padded i >= n reads 2(n-1)-i."""
    i = torch.arange(n_pad)
    return torch.where(i < n, i, 2 * (n - 1) - i)


def reflect_pad_nhwc(x: Tensor, win: Tuple[int, int]) -> Tensor:
    B, H, W, C = x.shape
    ph = (win[0] - H % win[0]) % win[0]
    pw = (win[1] - W % win[1]) % win[1]
    if ph >= H or pw >= W:
        # same condition under which F.pad(..., 'reflect') raises (hit_sir_pro.py:672)
        raise RuntimeError(
            f"Padding size should be less than the corresponding input dimension, but got: "
            f"padding ({pw}, {ph}) at dimension of input {list(x.shape)}")
    return x[:, reflect_index(H, H + ph)][:, :, reflect_index(W, W + pw)]


def spatial_channel_attention(x: Tensor, sd, p: str) -> Tensor:
    """SpatialChannelAttention.forward ("casa"), hit_sir_pro.py:338-359, on the padded NHWC map."""
    B, Hp, Wp, C = x.shape
    cavg = x.mean(dim=3, keepdim=True)
    cmax = x.max(dim=3, keepdim=True)[0]
    a1 = F.leaky_relu(conv_nhwc(cavg, sd[p + "linear1.weight"], sd[p + "linear1.bias"], 1), 0.2)
    a2 = F.leaky_relu(conv_nhwc(cmax, sd[p + "linear2.weight"], sd[p + "linear2.bias"], 1), 0.2)
    savg = x.mean(dim=(1, 2))                                        # (B, C)
    smax = x.amax(dim=(1, 2))
    s1 = F.linear(F.linear(savg, sd[p + "linear1_first.weight"], sd[p + "linear1_first.bias"]),
                  sd[p + "linear1_second.weight"], sd[p + "linear1_second.bias"])
    s2 = F.linear(F.linear(smax, sd[p + "linear2_first.weight"], sd[p + "linear2_first.bias"]),
                  sd[p + "linear2_second.weight"], sd[p + "linear2_second.bias"])
    attn = (a1 * s1.view(B, 1, 1, C) + a2 * s2.view(B, 1, 1, C)) / 2.0
    return attn + x


def to_windows(x: Tensor, win: Tuple[int, int]) -> Tensor:
    """window_partition, hit_sir_pro.py:236-248 -> (B*nW, L, C), windows row-major, tokens row-major."""
    B, H, W, C = x.shape
    x = x.view(B, H // win[0], win[0], W // win[1], win[1], C)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(-1, win[0] * win[1], C)


def from_windows(xw: Tensor, win: Tuple[int, int], B: int, H: int, W: int) -> Tensor:
    """window_reverse, hit_sir_pro.py:251-271."""
    C = xw.shape[-1]
    x = xw.view(B, H // win[0], W // win[1], win[0], win[1], C)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, C)


def spatial_linear_pool(t: Tensor, win, base, w_sl: Tensor, b_sl: Tensor) -> Tensor:
    """SCC.spatial_linear_projection, hit_sir_pro.py:435-456, on (nW, L, C') -> (nW, Lb, C').
    cell (cy,cx) = sum_{i,j<r} W[0, i*rw+j] * t[(cy*rh+i, cx*rw+j)] + bias."""
    nW, L, Cc = t.shape
    rh, rw = win[0] // base[0], win[1] // base[1]
    tt = t.view(nW, base[0], rh, base[1], rw, Cc)
    wsl = w_sl.view(rh, rw)
    return torch.einsum("nairjc,ij->narc", tt, wsl).reshape(nW, base[0] * base[1], Cc) + b_sl


def scc_forward(x_pad: Tensor, sd, p: str, win, base_win, heads: int, casa: bool,
                bias: Tensor, taps: Optional[dict] = None) -> Tensor:
    """SCC.forward, hit_sir_pro.py:542-599 (with :458-540).  x_pad (B,Hp,Wp,C) -> same shape."""
    B, Hp, Wp, C = x_pad.shape
    base = (min(win[0], base_win[0]), min(win[1], base_win[1]))
    t = spatial_channel_attention(x_pad, sd, p + "qkv.") if casa else x_pad
    if taps is not None:
        taps["qkv"] = t
    tw = to_windows(t, win)                                          # (nW, L, C)
    nW, L, _ = tw.shape
    half = C // 2
    d = half // heads
    q = tw[..., :half].reshape(nW, L, heads, d)                      # channel = t*half + h*d + j (:569)
    v = tw[..., half:].reshape(nW, L, heads, d)
    k = (F.linear(q, sd[p + "k_generate1.weight"], sd[p + "k_generate1.bias"]) +
         F.linear(v, sd[p + "k_generate2.weight"], sd[p + "k_generate2.bias"])) / 2.0   # :572
    # S-SC (:458-513)
    w_sl, b_sl = sd[p + "spatial_linear.weight"], sd[p + "spatial_linear.bias"]
    k_p = spatial_linear_pool(k.reshape(nW, L, half), win, base, w_sl, b_sl).view(nW, -1, heads, d)
    v_p = spatial_linear_pool(v.reshape(nW, L, half), win, base, w_sl, b_sl).view(nW, -1, heads, d)
    corr = torch.einsum("nlhd,nmhd->nhlm", q, k_p) / d + bias.unsqueeze(0)   # scale = head_dim (:421,475)
    x_sp = torch.einsum("nhlm,nmhd->nlhd", corr, v_p).reshape(nW, L, half)
    # C-SC (:515-540): heads merged, corr = q^T k / L
    qm, km, vm = q.reshape(nW, L, half), k.reshape(nW, L, half), v.reshape(nW, L, half)
    cc = torch.einsum("nlc,nle->nce", qm, km) / L
    x_ch = torch.einsum("nce,nle->nlc", cc, vm)
    out = torch.cat([x_sp, x_ch], dim=-1)                            # :596
    if taps is not None:
        taps["scc_pre_proj"] = from_windows(out, win, B, Hp, Wp)
    out = F.linear(out, sd[p + "proj.weight"], sd[p + "proj.bias"])
    return from_windows(out, win, B, Hp, Wp)


def conv_ffn(x: Tensor, sd, p: str) -> Tensor:
    """ConvFFN.forward + dwconv.forward, hit_sir_pro.py:39-46, 20-24 (NHWC in/out)."""
    h = F.gelu(F.linear(x, sd[p + "fc1.weight"], sd[p + "fc1.bias"]))
    wd, bd = sd[p + "dwconv.depthwise_conv.0.weight"], sd[p + "dwconv.depthwise_conv.0.bias"]
    dw = F.conv2d(h.permute(0, 3, 1, 2), wd, bd, 1, 2, 1, wd.shape[0]).permute(0, 2, 3, 1)
    h = h + F.gelu(dw)
    return F.linear(h, sd[p + "fc2.weight"], sd[p + "fc2.bias"])


def resi_conv(t: Tensor, sd, p: str) -> Tensor:
    """The conv before a residual connection: '1conv' = Conv2d(C, C, 3) (:911-912, :1221-1222); '3conv' = Conv3x3(C -> C/4), LeakyReLU(0.2),
    Conv1x1, LeakyReLU(0.2), Conv3x3(C/4 -> C) (:913-918, :1224-1231).  Told apart by the state_dict keys.  NHWC in/out."""
    if p + "weight" in sd:
        return conv_nhwc(t, sd[p + "weight"], sd[p + "bias"], 1)
    t = F.leaky_relu(conv_nhwc(t, sd[p + "0.weight"], sd[p + "0.bias"], 1), 0.2)
    t = F.leaky_relu(conv_nhwc(t, sd[p + "2.weight"], sd[p + "2.bias"], 0), 0.2)
    return conv_nhwc(t, sd[p + "4.weight"], sd[p + "4.bias"], 1)


def ln(x: Tensor, sd, p: str) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[p + "weight"], sd[p + "bias"], 1e-5)


def union_attention(x: Tensor, sd, p: str) -> Tensor:
    """UnionAttention.forward, hit_sir_pro.py:113-133.  x NHWC -> NHWC."""
    B, H, W, C = x.shape
    xn = x.permute(0, 3, 1, 2)                                        # (B,C,H,W)
    c_att = F.conv2d(torch.cat((xn.mean(1, keepdim=True), xn.max(1, keepdim=True)[0]), 1),
                     sd[p + "conv1.weight"], sd[p + "conv1.bias"], 1, 1)         # (B,1,H,W)
    h_in = torch.stack((xn.mean(2), xn.max(2)[0]), 1)                # (B,2,C,W): plane (channel,width)
    h_att = F.conv2d(h_in, sd[p + "conv2.weight"], sd[p + "conv2.bias"], 1, 1)   # (B,1,C,W)
    w_in = torch.stack((xn.mean(3), xn.max(3)[0]), 1)                # (B,2,C,H): plane (channel,height)
    w_att = F.conv2d(w_in, sd[p + "conv3.weight"], sd[p + "conv3.bias"], 1, 1)   # (B,1,C,H)
    s = c_att + w_att.view(B, C, H, 1) + h_att.view(B, C, 1, W)
    y = F.conv2d(s, sd[p + "conv_last.weight"], sd[p + "conv_last.bias"], 1, 1)
    return y.permute(0, 2, 3, 1)


def fusion(first: Tensor, second: Tensor, sd, p="fusion.") -> Tensor:
    """Fusion.forward(shallow=first, deep=second), hit_sir_pro.py:145-162.
    NOTE the call site passes (conv_after_body(deep), shallow) positionally (:1330)."""
    a1 = union_attention(first, sd, p + "union_attention1.")
    att = torch.sigmoid(union_attention(first + second, sd, p + "union_attention2."))
    a3 = union_attention(second, sd, p + "union_attention3.")
    return first * torch.sigmoid(a1 * att) + second * torch.sigmoid(a3 * (1 - att))


# --------------------------------------------------------------------------
# the forward pass
# --------------------------------------------------------------------------
class HiTSIROracle:
    def __init__(self, state_dict: Dict[str, Tensor], cfg: OracleConfig):
        self.cfg = cfg
        self.sd = {k: v.detach().float().cpu() for k, v in state_dict.items()}
        if cfg.upsampler == "nearest+conv" and cfg.upscale != 4:
            raise AssertionError("only support x4 now.")                # hit_sir_pro.py:1248
        self._bias: Dict[Tuple[int, int], Tensor] = {}
        self.mean = (torch.tensor(RGB_MEAN).view(1, 3, 1, 1) if cfg.in_chans == 3
                     else torch.zeros(1, 1, 1, 1))

    def pooled_bias(self, i: int, j: int) -> Tensor:
        """Weight-only; cached (the reference rebuilds it every forward, :477-503)."""
        if (i, j) not in self._bias:
            cfg = self.cfg
            win = cfg.windows(cfg.depths[i])[j]
            base = (min(win[0], cfg.base_win_size[0]), min(win[1], cfg.base_win_size[1]))
            p = f"layers.{i}.residual_group.blocks.{j}.correlation."
            tbl = dynamic_pos_bias_table(self.sd, p, win)
            self._bias[(i, j)] = pooled_rpe_bias(tbl, win, base)
        return self._bias[(i, j)]

    def block(self, x: Tensor, i: int, j: int, taps: Optional[dict] = None) -> Tensor:
        """HierarchicalTransformerBlock.forward, hit_sir_pro.py:676-706.  x NHWC."""
        cfg, sd = self.cfg, self.sd
        B, H, W, C = x.shape
        win = cfg.windows(cfg.depths[i])[j]
        p = f"layers.{i}.residual_group.blocks.{j}."
        sub = {} if taps is not None else None
        xp = reflect_pad_nhwc(x, win)
        a = scc_forward(xp, sd, p + "correlation.", win, tuple(cfg.base_win_size), cfg.num_heads[i],
                        cfg.is_channel_spatial_attn, self.pooled_bias(i, j), sub)
        a = a[:, :H, :W, :]
        x = x + ln(a, sd, p + "norm1.")
        if taps is not None:
            taps[f"block{i}.{j}.qkv"] = sub["qkv"]
            taps[f"block{i}.{j}.scc"] = sub["scc_pre_proj"][:, :H, :W, :]
            taps[f"block{i}.{j}.attn"] = x
        x = x + ln(conv_ffn(x, sd, p + "mlp."), sd, p + "norm2.")
        if taps is not None:
            taps[f"block{i}.{j}"] = x
        return x

    def forward(self, x: Tensor, taps: Optional[dict] = None) -> Tensor:
        """HiT_SIR.forward, hit_sir_pro.py:1304-1344.  x (B,3,H,W) fp32 -> (B,3,sH,sW)."""
        cfg, sd = self.cfg, self.sd
        x = x.float()
        B, _, H, W = x.shape
        mean = self.mean
        x = (x - mean) * cfg.img_range
        if cfg.is_mult_size_conv_feat_extract:
            shallow = multi_size_conv_extract(x, sd)
        else:
            shallow = F.conv2d(x, sd["conv_first.weight"], sd["conv_first.bias"], 1, 1).permute(0, 2, 3, 1)
        if taps is not None:
            taps["shallow"] = shallow
        # forward_features (:1284-1302); patch_embed = flatten + LayerNorm (:975-983)
        t = ln(shallow, sd, "patch_embed.norm.")
        if "absolute_pos_embed" in sd:                 # ape=True (:1293-1294): (1, num_patches, C) broadcast over the batch; H*W must equal num_patches
            t = t + sd["absolute_pos_embed"].view(1, H, W, -1)
        if taps is not None:
            taps["embed"] = t
        for i in range(len(cfg.depths)):
            t_in = t
            for j in range(cfg.depths[i]):
                t = self.block(t, i, j, taps)
            # RHTB.forward (:928-936)
            t = resi_conv(t, sd, f"layers.{i}.conv.") + t_in
            if taps is not None:
                taps[f"layer{i}"] = t
        t = ln(t, sd, "norm.")
        if taps is not None:
            taps["norm"] = t
        cab = resi_conv(t, sd, "conv_after_body.")
        if taps is not None:
            taps["conv_after_body"] = cab
        if cfg.is_fusion:
            f = fusion(cab, shallow, sd)           # positional binding, :1330
        else:
            f = cab + shallow                      # :1153
        if taps is not None:
            taps["fused"] = f
        fn = f.permute(0, 3, 1, 2)
        if cfg.upsampler == "pixelshuffle":        # :1313-1319
            y = F.leaky_relu(F.conv2d(fn, sd["conv_before_upsample.0.weight"], sd["conv_before_upsample.0.bias"], 1, 1), 0.01)
            if (cfg.upscale & (cfg.upscale - 1)) == 0:
                for s in range(int(math.log(cfg.upscale, 2))):
                    y = F.pixel_shuffle(F.conv2d(y, sd[f"upsample.{2 * s}.weight"], sd[f"upsample.{2 * s}.bias"], 1, 1), 2)
            elif cfg.upscale == 3:
                y = F.pixel_shuffle(F.conv2d(y, sd["upsample.0.weight"], sd["upsample.0.bias"], 1, 1), 3)
            else:
                raise ValueError(f"scale {cfg.upscale} is not supported. Supported scales: 2^n and 3.")
            y = F.conv2d(y, sd["conv_last.weight"], sd["conv_last.bias"], 1, 1)
        elif cfg.upsampler == "pixelshuffledirect":  # :1320-1325
            y = F.pixel_shuffle(F.conv2d(fn, sd["upsample.0.weight"], sd["upsample.0.bias"], 1, 1), cfg.upscale)
        elif cfg.upsampler == "nearest+conv":      # :1326-1334
            y = F.leaky_relu(F.conv2d(fn, sd["conv_before_upsample.0.weight"], sd["conv_before_upsample.0.bias"], 1, 1), 0.01)
            if taps is not None:
                taps["conv_before_upsample"] = y.permute(0, 2, 3, 1)
            y = F.leaky_relu(F.conv2d(F.interpolate(y, scale_factor=2, mode="nearest"), sd["conv_up1.weight"], sd["conv_up1.bias"], 1, 1), 0.2)
            if taps is not None:
                taps["up1"] = y.permute(0, 2, 3, 1)
            y = F.leaky_relu(F.conv2d(F.interpolate(y, scale_factor=2, mode="nearest"), sd["conv_up2.weight"], sd["conv_up2.bias"], 1, 1), 0.2)
            if taps is not None:
                taps["up2"] = y.permute(0, 2, 3, 1)
            y = F.leaky_relu(F.conv2d(y, sd["conv_hr.weight"], sd["conv_hr.bias"], 1, 1), 0.2)
            if taps is not None:
                taps["hr"] = y.permute(0, 2, 3, 1)
            y = F.conv2d(y, sd["conv_last.weight"], sd["conv_last.bias"], 1, 1)
        else:                                      # :1335-1340 (denoise mode)
            y = x + F.conv2d(fn, sd["conv_last.weight"], sd["conv_last.bias"], 1, 1)
        y = y / cfg.img_range + mean
        return y[:, :, :H * cfg.upscale, :W * cfg.upscale]     # upsampler=None: y is x-sized and the slice is a no-op (:1344)

    __call__ = forward
