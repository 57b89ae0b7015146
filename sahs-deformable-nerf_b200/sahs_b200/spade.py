"""Stage-II SPADE generators: `Generator` and `Generator_audio` of the reference
(nerf/_init_spade.py:318-328, :359-373) for inference -- the consumer of the Stage-I frames
(eval_get_texture_photo_audio.py:170-206: `G(identity_photo, stage1_frame, audio) -> refined frame`).

Same constructor, attribute tree and state_dict keys as the reference modules (checkpoints load with strict=True; the
two names `conv1` / `conv1_sn` of a spectral-normed conv are one module here as there).  The forward pass is this
library's: activations are NHWC fp16 in HBM, every 3x3 convolution is `sahs_spade_conv` (implicit-GEMM tcgen05 kernel,
csrc/spade_conv.cu) with bias / ReLU / residual add / the whole SPADE modulation + LeakyReLU fused into its epilogue;
nearest resizes of the conditioning maps and nn.Upsample are folded into the conv's gather and never stored; eval-mode
BatchNorm and spectral norm are folded into the packed weights.  What stays in PyTorch is the tiny audio encoder (four
Conv1d over a 16-step window, like Stage I's AudioNet) and the image layout conversion at both ends.

Training of Stage II (discriminator, VGG loss; train_get_texture_photo*.py) is out of scope."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import custom_ops  # noqa: F401  (registers torch.ops.sahs_b200.spade_conv / instnorm_stats / avgpool2)

MODE_S1, MODE_S2, MODE_T2, MODE_FIRST, MODE_T2C = 0, 1, 2, 3, 4
EPI_RELU, EPI_ADD, EPI_SPADE, EPI_F32 = 1, 2, 4, 8
T2_CLASS_MIN_PIXELS = 128 * 128      # transposed convs on smaller inputs run as one nine-tap launch (measured: 64^2 input 0.043 vs 0.067 ms)


# ---- parameter holders with the reference's attribute names -------------------------------------------------------------
class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder: the generator's forward runs the fused CUDA path")


class _Conv(_Holder):
    def __init__(self, cin, cout):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, 3, 3).normal_(0, (2.0 / (9 * cin)) ** 0.5))
        self.bias = nn.Parameter(torch.zeros(cout))


class _ConvT(_Holder):
    def __init__(self, cin, cout):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cin, cout, 3, 3).normal_(0, (2.0 / (9 * cin)) ** 0.5))
        self.bias = nn.Parameter(torch.zeros(cout))


class _SNConv(_Holder):
    """spectral_norm(nn.Conv2d): bias, weight_orig, buffers weight_u / weight_v (torch.nn.utils.spectral_norm)"""

    def __init__(self, cin, cout):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(cout))
        self.weight_orig = nn.Parameter(torch.empty(cout, cin, 3, 3).normal_(0, (2.0 / (9 * cin)) ** 0.5))
        self.register_buffer("weight_u", F.normalize(torch.randn(cout), dim=0))
        self.register_buffer("weight_v", F.normalize(torch.randn(cin * 9), dim=0))


class _BN(_Holder):
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class ResBlock2d(_Holder):
    """ref: nerf/_init_spade.py:7-38"""

    def __init__(self, cin, cout, downsample=False):
        super().__init__()
        self.downsample = downsample
        self.initial = nn.Sequential(_Conv(cin, cout), _BN(cout), nn.ReLU())
        if downsample:
            self.downsample_layer = _Conv(cin, cout)
            self.residual_downsample = _Conv(cout, cout)
        self.residual = nn.Sequential(_Conv(cout, cout), _BN(cout), nn.ReLU())


class IdEncoder(_Holder):
    """ref: :183-199"""

    def __init__(self):
        super().__init__()
        self.layer1 = nn.Sequential(_Conv(3, 64), nn.AvgPool2d(2, stride=2))
        self.layer2 = ResBlock2d(64, 64)
        self.layer3 = ResBlock2d(64, 128, downsample=True)
        self.layer4 = ResBlock2d(128, 256, downsample=True)


class SPADELayer(_Holder):
    """ref: :114-139"""

    def __init__(self, norm_nc, label_nc):
        super().__init__()
        self.mlp_shared = nn.Sequential(_Conv(label_nc, 128), nn.ReLU())
        self.conv_gamma = _Conv(128, norm_nc)
        self.conv_beta = _Conv(128, norm_nc)


class SPADEBlock(_Holder):
    """ref: :235-282"""

    def __init__(self, cin, cout, fid_channels, downsample=False, upsample=False):
        super().__init__()
        self.spade1 = SPADELayer(cin, fid_channels)
        self.conv1 = _SNConv(cin, cout)
        self.conv1_sn = self.conv1
        self.spade2 = SPADELayer(cout, fid_channels)
        self.conv2 = _SNConv(cout, cout)
        self.conv2_sn = self.conv2
        self.downsample, self.upsample = downsample, upsample
        if downsample:
            self.residual_downsample = _Conv(cin, cin)
        if upsample:
            self.residual_upsample = _ConvT(cin, cin)
        self.spade_s = SPADELayer(cin, fid_channels)
        self.conv_s = _SNConv(cin, cout)


class RefineNetwork(_Holder):
    """ref: :286-315"""

    def __init__(self, fid1, fid2, fid3):
        super().__init__()
        self.layer1 = nn.Sequential(_Conv(3, 64), nn.AvgPool2d(2, stride=2))
        self.layer2 = SPADEBlock(64, 64, fid1, downsample=True)
        self.layer3 = SPADEBlock(64, 128, fid2, downsample=True)
        self.layer4 = SPADEBlock(128, 256, fid3)
        self.layer5 = SPADEBlock(256, 256, fid3, upsample=True)
        self.layer6 = SPADEBlock(256, 128, fid2, upsample=True)
        self.layer7 = SPADEBlock(128, 64, fid1, upsample=True)
        self.layer8 = _Conv(64, 3)


class AudioNet(nn.Module):
    """The audio encoder of _init_spade.py:330-357 (LeakyReLU 0.02; forward = the conv stack, encoder_fc1 is unused)."""

    def __init__(self, dim_aud=76, win_size=16):
        super().__init__()
        self.win_size, self.dim_aud = win_size, dim_aud
        convs = []
        for cin, cout in ((29, 32), (32, 32), (32, 64), (64, 64)):
            convs += [nn.Conv1d(cin, cout, kernel_size=3, stride=2, padding=1, bias=True), nn.LeakyReLU(0.02, True)]
        self.encoder_conv = nn.Sequential(*convs)
        self.encoder_fc1 = nn.Sequential(nn.Linear(64, 64), nn.LeakyReLU(0.02, True), nn.Linear(64, dim_aud))

    def forward(self, x):
        half_w = int(self.win_size / 2)
        x = x[:, 8 - half_w:8 + half_w, :].permute(0, 2, 1)
        return self.encoder_conv(x).squeeze(-1)


# ---- weight packing ------------------------------------------------------------------------------------------------------
def _swizzle_rows(t: torch.Tensor) -> torch.Tensor:
    """[..., rows, 64] fp16 -> the same bytes as K-major SWIZZLE_128B operand tiles (16-byte unit u of row r at unit
    u ^ (r & 7)); rows % 8 == 0."""
    lead, rows = t.shape[:-2], t.shape[-2]
    v = t.reshape(*lead, rows // 8, 8, 8, 8)
    unit = torch.arange(8, device=t.device)
    src = (unit[None, :] ^ unit[:, None])                       # [r, physical unit] -> logical unit
    idx = src.reshape(*([1] * (len(lead) + 1)), 8, 8, 1).expand(*lead, rows // 8, 8, 8, 8)
    return torch.gather(v, len(lead) + 2, idx).reshape(*lead, rows, 64).contiguous()


class _Packed:
    """One convolution ready for sahs_spade_conv."""

    def __init__(self, rows: torch.Tensor, bias: torch.Tensor, ntile: int, cin: int, cout: int, first: bool, taps=(3, 3)):
        # rows: [ntiles * ntile, nky * nkx, cin_padded] fp32 (first: [.., 64] with the nine taps folded into one chunk)
        n = rows.shape[0]
        self.ntile, self.ntiles, self.cin, self.cout, self.first = ntile, n // ntile, cin, cout, first
        self.taps = taps
        if first:
            blocks = rows.reshape(self.ntiles, ntile, 1, 64).permute(0, 2, 1, 3)
        else:
            nky, nkx = taps
            kcs = rows.shape[2] // 64                   # K chunk order: [ky][64-channel chunk][kx] (csrc/spade_conv.cu)
            blocks = rows.reshape(self.ntiles, ntile, nky, nkx, kcs, 64).permute(0, 2, 4, 3, 1, 5).reshape(
                self.ntiles, nky * nkx * kcs, ntile, 64)
        self.packed = _swizzle_rows(blocks.to(torch.float16).contiguous())
        self.bias = bias.float().contiguous()


def _pad_rows(t: torch.Tensor, n: int) -> torch.Tensor:
    if t.shape[0] == n:
        return t
    out = torch.zeros((n,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    out[:t.shape[0]] = t
    return out


class _PackedT2:
    """nn.ConvTranspose2d(3x3, stride 2, padding 1, output_padding 1) as four output-parity classes (py, px): output
    (2i + py, 2j + px) only receives the taps with (o + 1 - k) even -- k = 1 for parity 0, k = 0 and 2 for parity 1 -- so
    each class is a plain conv over the input grid with 1, 2, 2 or 4 taps (csrc/spade_conv.cu mode 4)."""

    def __init__(self, rows: torch.Tensor, b: torch.Tensor, cin: int, cout: int):
        ntile = 128 if cout >= 128 else (64 if cout >= 64 else 16)
        n = (cout + ntile - 1) // ntile * ntile
        self.cin, self.cout, self.first = cin, cout, False
        # small inputs (few tiles) run better as ONE launch over all nine taps (mode 2, zero taps skipped per tile)
        self.full = _Packed(_pad_rows(rows.reshape(cout, 9, cin), n), _pad_rows(b, n), ntile, cin, cout, False)
        self.classes = []
        for py in (0, 1):
            for px in (0, 1):
                kys, kxs = ([0, 2] if py else [1]), ([0, 2] if px else [1])
                sel = rows[:, kys][:, :, kxs].reshape(cout, len(kys) * len(kxs), cin)
                self.classes.append(_Packed(_pad_rows(sel, n), _pad_rows(b, n), ntile, cin, cout, False, (len(kys), len(kxs))))


def _pack_conv(w: torch.Tensor, b: torch.Tensor, transposed=False):
    """w: effective fp32 weight, [cout, cin, 3, 3] (transposed: [cin, cout, 3, 3])"""
    rows = (w.permute(1, 2, 3, 0) if transposed else w.permute(0, 2, 3, 1))          # [cout, ky, kx, cin]
    cout, cin = rows.shape[0], rows.shape[3]
    if transposed:
        return _PackedT2(rows, b, cin, cout)
    if cin == 3:
        r = torch.zeros(cout, 9, 4, dtype=w.dtype, device=w.device)
        r[:, :, :3] = rows.reshape(cout, 9, 3)
        r = F.pad(r.reshape(cout, 36), (0, 28))
        ntile = 64 if cout >= 64 else 16
        n = (cout + ntile - 1) // ntile * ntile
        return _Packed(_pad_rows(r, n), _pad_rows(b, n), ntile, 3, cout, True)
    ntile = 128 if cout >= 128 else (64 if cout >= 64 else 16)
    n = (cout + ntile - 1) // ntile * ntile
    return _Packed(_pad_rows(rows.reshape(cout, 9, cin), n), _pad_rows(b, n), ntile, cin, cout, False)


def _pack_gamma_beta(wg, bg, wb, bb) -> _Packed:
    """conv_gamma and conv_beta as one conv whose N tiles are [gamma rows 64t..64t+63 | beta rows 64t..64t+63]"""
    c, cin = wg.shape[0], wg.shape[1]
    g = wg.permute(0, 2, 3, 1).reshape(c // 64, 64, 9, cin)
    bt = wb.permute(0, 2, 3, 1).reshape(c // 64, 64, 9, cin)
    rows = torch.cat((g, bt), 1).reshape(2 * c, 9, cin)
    bias = torch.cat((bg.reshape(c // 64, 64), bb.reshape(c // 64, 64)), 1).reshape(2 * c)
    p = _Packed(rows, bias, 128, cin, c, False)
    return p


class Identity:
    """Everything a generator derives from the identity photo alone: the IdEncoder feature maps and -- filled on the first
    frame that uses them -- the `mlp_shared` activations of every SPADE layer conditioned on one of those maps.  A clip is
    refined with ONE identity photo (ref: eval_get_texture_photo_audio.py:163-173, `frame = indenty_photo` inside the frame
    loop), so all of this is frame independent: `G.refine(identity, frame)` skips the IdEncoder and 18 (Generator) / 12
    (Generator_audio) of the 70 convolutions of `G(I_src, frame)` and returns the same bits."""

    def __init__(self, fids, shape):
        self.fids, self.shape = fids, shape
        self.actv: Dict[Tuple[str, int, int], torch.Tensor] = {}


# ---- the generators ------------------------------------------------------------------------------------------------------
class Generator(nn.Module):
    """ref: nerf/_init_spade.py:318-328.  forward(I_src, I_raw) -> refined frame, all [1, 3, H, W] fp32 (H, W multiples of 8)."""

    def __init__(self):
        super().__init__()
        self.idencoder = IdEncoder()
        self.refine_network = RefineNetwork(64, 128, 256)
        self._packed: Optional[Dict[str, _Packed]] = None
        self.tally: Optional[Dict[str, float]] = None     # set to {} to count launches and algorithmic FLOPs of a forward

    # -- packing (redone after load_state_dict / device moves; call invalidate_packed() after editing parameters in place)
    def invalidate_packed(self):
        self._packed = None

    def load_state_dict(self, *a, **k):
        self._packed = None
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    @torch.no_grad()
    def _prepare(self) -> Dict[str, _Packed]:
        if self._packed is not None:
            return self._packed
        P: Dict[str, _Packed] = {}
        mods = dict(self.named_modules())

        def plain(name, transposed=False):
            m = mods[name]
            P[name] = _pack_conv(m.weight.float(), m.bias.float(), transposed)

        def with_bn(name):     # Sequential(conv, BatchNorm2d (eval), ReLU): y = (conv(x) + b - rm) * s + beta, s = g / sqrt(rv + eps)
            conv, bn = mods[name + ".0"], mods[name + ".1"]
            s = bn.weight.float() / torch.sqrt(bn.running_var.float() + 1e-5)
            P[name] = _pack_conv(conv.weight.float() * s[:, None, None, None],
                                 (conv.bias.float() - bn.running_mean.float()) * s + bn.bias.float())

        def sn(name):          # spectral_norm in eval mode: weight_orig / (u . (W v)), no power iteration
            m = mods[name]
            w = m.weight_orig.float()
            sigma = torch.dot(m.weight_u.float(), torch.mv(w.reshape(w.shape[0], -1), m.weight_v.float()))
            P[name] = _pack_conv(w / sigma, m.bias.float())

        def spade(name):
            plain(name + ".mlp_shared.0")
            g, b = mods[name + ".conv_gamma"], mods[name + ".conv_beta"]
            P[name + ".gb"] = _pack_gamma_beta(g.weight.float(), g.bias.float(), b.weight.float(), b.bias.float())

        plain("idencoder.layer1.0")
        for lay, down in (("layer2", False), ("layer3", True), ("layer4", True)):
            with_bn(f"idencoder.{lay}.initial")
            if down:
                plain(f"idencoder.{lay}.downsample_layer")
                plain(f"idencoder.{lay}.residual_downsample")
            else:
                with_bn(f"idencoder.{lay}.residual")
        plain("refine_network.layer1.0")
        for i in range(2, 8):
            pfx = f"refine_network.layer{i}"
            blk = mods[pfx]
            for s_ in ("spade1", "spade2", "spade_s"):
                spade(f"{pfx}.{s_}")
            for c_ in ("conv1", "conv2", "conv_s"):
                sn(f"{pfx}.{c_}")
            if blk.downsample:
                plain(f"{pfx}.residual_downsample")
            if blk.upsample:
                plain(f"{pfx}.residual_upsample", transposed=True)
        plain("refine_network.layer8")
        self._packed = P
        return P

    # -- kernels (torch.ops.sahs_b200.* custom operators over the C ABI, sahs_b200/custom_ops.py)
    def _conv(self, p: _Packed, x: torch.Tensor, out_h: int, out_w: int, mode: int, up=0, down=0, relu=False, add=None,
              spade=None, aux_shift=0, f32=False) -> torch.Tensor:
        if isinstance(p, _PackedT2) and x.shape[0] * x.shape[1] < T2_CLASS_MIN_PIXELS:
            p = p.full                                   # one launch, mode 2
        if isinstance(p, _PackedT2):                 # transposed conv: four parity-class launches into one output
            if relu or add is not None or spade is not None or f32:
                raise RuntimeError("the parity-class transposed conv has the plain bias epilogue only")
            out = torch.ops.sahs_b200.spade_conv_t2(x, [c.packed for c in p.classes], [c.bias for c in p.classes], p.cin, p.cout)
            if self.tally is not None:
                self.tally["flop"] = self.tally.get("flop", 0.0) + 2.0 * out_h * out_w * p.cout * 2.25 * p.cin
                self.tally["conv_launches"] = self.tally.get("conv_launches", 0) + 4
            return out
        if p.first:
            mode = MODE_FIRST
        epi = (EPI_RELU if relu else 0) | (EPI_ADD if add is not None else 0) | (EPI_SPADE if spade is not None else 0) | \
              (EPI_F32 if f32 else 0)
        aux = spade[0] if spade is not None else add
        mean, rstd = (spade[1], spade[2]) if spade is not None else (None, None)
        out = torch.ops.sahs_b200.spade_conv(x, p.packed, p.bias, p.cin, p.cout, out_h, out_w, mode, up, down, epi, aux,
                                             aux_shift, mean, rstd)
        if self.tally is not None:
            n_out = 2 * p.cout if spade is not None else p.cout
            taps = 9.0 / 4.0 if mode == MODE_T2 else 9.0          # a stride-2 transposed conv touches 9/4 taps per output
            self.tally["flop"] = self.tally.get("flop", 0.0) + 2.0 * out_h * out_w * n_out * taps * p.cin
            self.tally["conv_launches"] = self.tally.get("conv_launches", 0) + 1
        return out

    def _stats(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.tally is not None:
            self.tally["other_launches"] = self.tally.get("other_launches", 0) + 2
        return torch.ops.sahs_b200.instnorm_stats(x, 1e-5)

    def _avgpool(self, x: torch.Tensor) -> torch.Tensor:
        if self.tally is not None:
            self.tally["other_launches"] = self.tally.get("other_launches", 0) + 1
        return torch.ops.sahs_b200.avgpool2(x)

    # -- network
    @staticmethod
    def _image(img: torch.Tensor) -> torch.Tensor:
        """[1, 3, H, W] fp32 -> [H, W, 4] fp16 (channel 3 = 0)"""
        if img.dim() != 4 or img.shape[0] != 1 or img.shape[1] != 3:
            raise RuntimeError("expected a [1, 3, H, W] image")
        if not img.is_cuda:
            raise RuntimeError("sahs_b200 ops need CUDA tensors (there is no CPU path)")
        out = torch.zeros(img.shape[2], img.shape[3], 4, dtype=torch.float16, device=img.device)
        out[..., :3] = img[0].permute(1, 2, 0)
        return out

    def _id_encoder(self, P, src):
        H, W = src.shape[0], src.shape[1]
        t = self._avgpool(self._conv(P["idencoder.layer1.0"], src, H, W, MODE_FIRST))
        h, w = H // 2, W // 2
        o = self._conv(P["idencoder.layer2.initial"], t, h, w, MODE_S1, relu=True)
        x1 = self._conv(P["idencoder.layer2.residual"], o, h, w, MODE_S1, relu=True, add=t)
        feats = [x1]
        for lay in ("layer3", "layer4"):
            x = feats[-1]
            o = self._conv(P[f"idencoder.{lay}.initial"], x, h, w, MODE_S1, relu=True)
            idn = self._conv(P[f"idencoder.{lay}.downsample_layer"], x, h // 2, w // 2, MODE_S2)
            feats.append(self._conv(P[f"idencoder.{lay}.residual_downsample"], o, h // 2, w // 2, MODE_S2, add=idn))
            h, w = h // 2, w // 2
        return feats

    def _fid_at(self, fid, h, w):
        """conditioning map + the power-of-two nearest resize that brings it to (h, w)"""
        if callable(fid):
            return fid(h, w), 0, 0
        fh, fw = fid.shape[0], fid.shape[1]
        if (fh, fw) == (h, w):
            return fid, 0, 0
        if (fh * 2, fw * 2) == (h, w):
            return fid, 1, 0
        if (fh, fw) == (h * 2, w * 2):
            return fid, 0, 1
        raise RuntimeError(f"conditioning map {fh}x{fw} cannot be resized to {h}x{w} by a factor of two")

    def _spade_act(self, P, pfx, x, x_shift, fid, h, w, ident=None):
        """leaky_relu(SPADELayer(x, fid), 0.2) at (h, w); x is stored at (h >> x_shift, w >> x_shift)"""
        mean, rstd = self._stats(x)          # statistics of a nearest-upsampled tensor are those of the stored one
        key = (pfx, h, w)
        actv = ident.actv.get(key) if (ident is not None and not callable(fid)) else None
        if actv is None:
            f, up, down = self._fid_at(fid, h, w)
            actv = self._conv(P[pfx + ".mlp_shared.0"], f, h, w, MODE_S1, up=up, down=down, relu=True)
            if ident is not None and not callable(fid):      # a map of the identity photo: the same for every frame
                ident.actv[key] = actv
        return self._conv(P[pfx + ".gb"], actv, h, w, MODE_S1, spade=(x, mean, rstd), aux_shift=x_shift)

    def _block(self, P, pfx, x, fid, down=False, up=False, ident=None):
        h, w = x.shape[0], x.shape[1]
        x1 = self._conv(P[pfx + ".conv1"], self._spade_act(P, pfx + ".spade1", x, 0, fid, h, w, ident), h, w, MODE_S1)
        idn, oh, ow, s1 = x, h, w, 0
        if down:
            x1 = self._avgpool(x1)
            oh, ow = h // 2, w // 2
            idn = self._conv(P[pfx + ".residual_downsample"], x, oh, ow, MODE_S2)
        if up:
            oh, ow, s1 = 2 * h, 2 * w, 1                    # x1 stays at (h, w): nn.Upsample is folded into its consumers
            idn = self._conv(P[pfx + ".residual_upsample"], x, oh, ow, MODE_T2)
        x2 = self._conv(P[pfx + ".conv2"], self._spade_act(P, pfx + ".spade2", x1, s1, fid, oh, ow, ident), oh, ow, MODE_S1)
        hs = self._spade_act(P, pfx + ".spade_s", idn, 0, fid, oh, ow, ident)
        return self._conv(P[pfx + ".conv_s"], hs, oh, ow, MODE_S1, add=x2)

    def _refine(self, P, raw, fid1, fid2, fid3, taps=None, ident=None):
        H, W = raw.shape[0], raw.shape[1]
        x = self._avgpool(self._conv(P["refine_network.layer1.0"], raw, H, W, MODE_FIRST))
        for name, fid, kw in (("layer2", fid1, dict(down=True)), ("layer3", fid2, dict(down=True)), ("layer4", fid3, {}),
                              ("layer5", fid3, dict(up=True)), ("layer6", fid2, dict(up=True)), ("layer7", fid1, dict(up=True))):
            x = self._block(P, "refine_network." + name, x, fid, ident=ident, **kw)
            if taps is not None:
                taps[name] = x
        return self._conv(P["refine_network.layer8"], x, H, W, MODE_S1, f32=True)

    def _check(self, *imgs):
        h, w = imgs[0].shape[2], imgs[0].shape[3]
        if h % 8 or w % 8 or h < 8 or w < 8:
            raise RuntimeError("image height and width must be multiples of 8")
        for im in imgs:
            if tuple(im.shape) != (1, 3, h, w):
                raise RuntimeError("I_src and I_raw must both be [1, 3, H, W]")

    @torch.no_grad()
    def encode_identity(self, I_src) -> Identity:
        """The frame-independent part of a clip: IdEncoder(I_src) (+ the conditioning activations, filled by the first
        `refine`).  Invalid after the weights change."""
        self._check(I_src, I_src)
        P = self._prepare()
        return Identity(tuple(self._id_encoder(P, self._image(I_src))), tuple(I_src.shape))

    @torch.no_grad()
    def refine(self, identity: Identity, I_raw, taps=None):
        """G(I_src, I_raw) with everything that depends on I_src alone taken from `identity`: same result, bit for bit."""
        self._check(I_raw, I_raw)
        if tuple(I_raw.shape) != identity.shape:
            raise RuntimeError("the frame and the identity photo must have the same size")
        P = self._prepare()
        fid1, fid2, fid3 = identity.fids
        out = self._refine(P, self._image(I_raw), fid1, fid2, fid3, taps, identity)
        return out.permute(2, 0, 1).unsqueeze(0)

    @torch.no_grad()
    def forward(self, I_src, I_raw, taps=None):
        self._check(I_src, I_raw)
        P = self._prepare()
        fid1, fid2, fid3 = self._id_encoder(P, self._image(I_src))
        out = self._refine(P, self._image(I_raw), fid1, fid2, fid3, taps)
        return out.permute(2, 0, 1).unsqueeze(0)


class Generator_audio(Generator):
    """ref: nerf/_init_spade.py:359-373.  forward(I_src, I_raw, driving_data [16, 29]): the third conditioning map is the
    audio feature, `a.unsqueeze(1).repeat(1, 256, 64, 64)` = [1, 256, 64, 64 * 64], which every SPADE layer resizes with
    nearest interpolation: channel-independent, constant over rows, a[floor(x * 4096 / w) % 64] along a row of width w."""

    def __init__(self):
        super().__init__()
        self.AudioNet = AudioNet(76, 16)

    def _audio_fid(self, driving_data):
        a = self.AudioNet(driving_data.unsqueeze(0).float())[0]                    # [64]
        maps: Dict[Tuple[int, int], torch.Tensor] = {}

        def fid3(h, w):
            if (h, w) not in maps:
                cols = torch.arange(4096, dtype=torch.float32, device=a.device).remainder(64).view(1, 1, 1, 4096)
                idx = F.interpolate(cols, size=(1, w), mode="nearest").view(w).long()
                maps[(h, w)] = a[idx].to(torch.float16).view(1, w, 1).expand(h, w, 256).contiguous()
            return maps[(h, w)]

        return fid3

    @torch.no_grad()
    def refine(self, identity: Identity, I_raw, driving_data, taps=None):
        """G(I_src, I_raw, driving_data) with the identity photo's part taken from `identity` (the audio-conditioned SPADE
        layers are per frame and are not cached)."""
        self._check(I_raw, I_raw)
        if tuple(I_raw.shape) != identity.shape:
            raise RuntimeError("the frame and the identity photo must have the same size")
        P = self._prepare()
        fid1, fid2, _ = identity.fids
        out = self._refine(P, self._image(I_raw), fid1, fid2, self._audio_fid(driving_data), taps, identity)
        return out.permute(2, 0, 1).unsqueeze(0)

    @torch.no_grad()
    def forward(self, I_src, I_raw, driving_data, taps=None):
        self._check(I_src, I_raw)
        P = self._prepare()
        fid1, fid2, _ = self._id_encoder(P, self._image(I_src))
        fid3 = self._audio_fid(driving_data)
        out = self._refine(P, self._image(I_raw), fid1, fid2, fid3, taps)
        return out.permute(2, 0, 1).unsqueeze(0)


class GraphedGenerator:
    """One CUDA graph per frame: the ~110 launches of a forward (convs, statistics, pooling, layout conversion) are
    recorded once for fixed input shapes and replayed; inputs are copied into the graph's static buffers.  Use for clip
    refinement (one identity photo, many Stage-I frames)."""

    def __init__(self, gen: Generator, *example_inputs: torch.Tensor, warmup: int = 2, identity: Optional[Identity] = None):
        """example_inputs: (I_src, I_raw[, audio]) -- or, with `identity=gen.encode_identity(I_src)`, (I_raw[, audio]): the
        graph then holds only the per-frame work (`gen.refine`), the identity photo's part having been computed once."""
        self.gen = gen
        self.inputs = [t.detach().clone() for t in example_inputs]
        run = (lambda: gen.refine(identity, *self.inputs)) if identity is not None else (lambda: gen(*self.inputs))
        side = torch.cuda.Stream(device=self.inputs[0].device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):              # (fills the identity's conditioning activations)
                run()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.output = run()

    def __call__(self, *inputs: torch.Tensor) -> torch.Tensor:
        for dst, src in zip(self.inputs, inputs):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.output
