"""CPU interpretation of what csrc/spade_conv.cu computes from a packed convolution (test infrastructure, like
tests/test_host_plan.py's numpy interpretation of the field plan): the packed, swizzled fp16 weight blocks are
unswizzled and multiplied with operand tiles gathered by the kernel's coordinate rules (src_pixel).  Subclassing the
generators with these three methods runs the product's whole host-side orchestration -- BatchNorm / spectral-norm
folding, gamma|beta tiling, resize shifts, upsample folding, residual wiring -- on the CPU against the oracle."""
import torch

from sahs_b200 import spade as SP


def unswizzle(packed: torch.Tensor) -> torch.Tensor:
    """[ntiles, nchunks, ntile, 64] swizzled fp16 -> logical fp32"""
    nt, nq, rows, _ = packed.shape
    v = packed.reshape(nt, nq, rows // 8, 8, 8, 8)
    unit = torch.arange(8)
    src = (unit[None, :] ^ unit[:, None]).reshape(1, 1, 1, 8, 8, 1).expand(nt, nq, rows // 8, 8, 8, 8)
    return torch.gather(v, 4, src).reshape(nt, nq, rows, 64).float()          # the XOR permutation is its own inverse


def src_pixel(mode, oy, ox, ky, kx, in_h, in_w, out_h, out_w, up, down):
    if mode == SP.MODE_S2:
        iy, ix = 2 * oy + ky - 1, 2 * ox + kx - 1
        return (iy >= 0) & (iy < in_h) & (ix >= 0) & (ix < in_w), iy, ix
    if mode == SP.MODE_T2:
        ty, tx = oy + 1 - ky, ox + 1 - kx
        iy, ix = ty >> 1, tx >> 1
        return (ty >= 0) & (tx >= 0) & (ty % 2 == 0) & (tx % 2 == 0) & (iy < in_h) & (ix < in_w), iy, ix
    ly, lx = oy + ky - 1, ox + kx - 1
    return (ly >= 0) & (ly < out_h) & (lx >= 0) & (lx < out_w), (ly << down) >> up, (lx << down) >> up


def conv_t2(p, x):
    """_PackedT2: four parity classes (py, px), each a plain conv over the input grid with the class's taps (mode 4 of the
    kernel: tap index ky -> source row i + dy[ky], dy = [1, 0] for parity 1 and [0] for parity 0), scattered to
    (2i + py, 2j + px)."""
    in_h, in_w = x.shape[0], x.shape[1]
    out = torch.zeros(2 * in_h, 2 * in_w, p.cout)
    xf = x.float()
    pix = torch.arange(in_h * in_w)
    oy, ox = pix // in_w, pix % in_w
    kcs = p.cin // 64
    for cls, c in enumerate(p.classes):
        py, px = cls >> 1, cls & 1
        dys, dxs = ([1, 0] if py else [0]), ([1, 0] if px else [0])
        nky, nkx = c.taps
        assert (nky, nkx) == (len(dys), len(dxs))
        W = unswizzle(c.packed.cpu())
        acc = torch.zeros(in_h * in_w, c.ntiles * c.ntile)
        for q in range(W.shape[1]):
            kx, g = q % nkx, q // nkx
            ky, kc = g // kcs, g % kcs
            iy, ix = oy + dys[ky], ox + dxs[kx]
            ok = (iy < in_h) & (ix < in_w)
            vals = xf[iy.clamp(0, in_h - 1), ix.clamp(0, in_w - 1), kc * 64:(kc + 1) * 64]
            A = torch.where(ok[:, None], vals, torch.zeros_like(vals))
            for t in range(c.ntiles):
                acc[:, t * c.ntile:(t + 1) * c.ntile] += A @ W[t, q].T
        acc = (acc + c.bias.cpu()[None, :])[:, :p.cout]
        out[2 * oy + py, 2 * ox + px] = acc
    return out


def conv(p, x, out_h, out_w, mode, up=0, down=0, relu=False, add=None, spade=None, aux_shift=0, f32=False):
    """same signature as Generator._conv; tensors are fp32 [H, W, C] on the CPU (fp16 rounding of activations is the
    GPU tests' subject)."""
    if isinstance(p, SP._PackedT2):                  # same choice as Generator._conv
        if x.shape[0] * x.shape[1] >= SP.T2_CLASS_MIN_PIXELS:
            return conv_t2(p, x)
        p = p.full
    if p.first:
        mode = SP.MODE_FIRST
    W = unswizzle(p.packed.cpu())
    in_h, in_w = x.shape[0], x.shape[1]
    P = out_h * out_w
    pix = torch.arange(P)
    oy, ox = pix // out_w, pix % out_w
    acc = torch.zeros(P, p.ntiles * p.ntile)
    xf = x.float()
    nq = W.shape[1]
    for q in range(nq):
        A = torch.zeros(P, 64)
        if p.first:
            for k in range(9):
                ok, iy, ix = src_pixel(mode, oy, ox, k // 3, k % 3, in_h, in_w, out_h, out_w, up, down)
                vals = xf[iy.clamp(0, in_h - 1), ix.clamp(0, in_w - 1), :4]
                A[:, 4 * k:4 * k + 4] = torch.where(ok[:, None], vals, torch.zeros_like(vals))
        else:
            kcs = p.cin // 64
            kx, g = q % 3, q // 3                       # chunk order [ky][64-channel chunk][kx]
            ky, kc = g // kcs, g % kcs
            tap = 3 * ky + kx
            ok, iy, ix = src_pixel(mode, oy, ox, tap // 3, tap % 3, in_h, in_w, out_h, out_w, up, down)
            vals = xf[iy.clamp(0, in_h - 1), ix.clamp(0, in_w - 1), kc * 64:(kc + 1) * 64]
            A = torch.where(ok[:, None], vals, torch.zeros_like(vals))
        for t in range(p.ntiles):
            acc[:, t * p.ntile:(t + 1) * p.ntile] += A @ W[t, q].T
    acc = acc + p.bias.cpu()[None, :]
    if spade is not None:
        aux, mean, rstd = spade
        ay, ax = oy >> aux_shift, ox >> aux_shift
        xa = aux.float()[ay, ax, :p.cout]
        g = acc.reshape(P, p.ntiles, 2, 64)[:, :, 0, :].reshape(P, p.cout)
        b = acc.reshape(P, p.ntiles, 2, 64)[:, :, 1, :].reshape(P, p.cout)
        y = (xa - mean[None, :]) * rstd[None, :] * (1 + g) + b
        return torch.where(y > 0, y, 0.2 * y).reshape(out_h, out_w, p.cout)
    out = acc[:, :p.cout]
    if relu:
        out = out.clamp_min(0)
    if add is not None:
        out = out + add.float().reshape(P, -1)[:, :p.cout]
    return out.reshape(out_h, out_w, p.cout)


class EmulatedMixin:
    def _conv(self, p, x, out_h, out_w, mode, **kw):
        return conv(p, x, out_h, out_w, mode, **kw)

    def _stats(self, x):
        f = x.float().reshape(-1, x.shape[2]).double()
        mean = f.mean(0)
        var = (f * f).mean(0) - mean * mean
        return mean.float(), (1.0 / torch.sqrt(var.clamp_min(0) + 1e-5)).float()

    def _avgpool(self, x):
        h, w, c = x.shape
        return x.float().reshape(h // 2, 2, w // 2, 2, c).mean((1, 3))

    @staticmethod
    def _image(img):
        out = torch.zeros(img.shape[2], img.shape[3], 4)
        out[..., :3] = img[0].permute(1, 2, 0)
        return out


class EmulatedGenerator(EmulatedMixin, SP.Generator):
    pass


class EmulatedGeneratorAudio(EmulatedMixin, SP.Generator_audio):
    pass
