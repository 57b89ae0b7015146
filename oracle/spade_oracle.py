"""TEST INFRASTRUCTURE -- CPU restatement (plain torch fp32 ops, functional, driven by a state_dict) of the reference's
Stage-II SPADE generators, SURVEY.md section 8(f) row 3.  Only tests/, __graft_entry__.smoke() and bench.py's baseline
legs may import this file; the product (sahs_b200/spade.py + csrc/spade.cu) never does.

Restates nerf/_init_spade.py of the reference:
  ResBlock2d :7-38, SPADELayer :114-139, IdEncoder :183-199, SPADEBlock :235-282, RefineNetwork :286-315,
  Generator :318-328, AudioNet (conv part) :330-357, Generator_audio :359-373.
Inference only (module.eval()): BatchNorm uses its running statistics, spectral_norm divides weight_orig by
sigma = u . (W v) without a power iteration (torch.nn.utils.spectral_norm, eval branch).

Parity pinning: oracle/make_golden_spade.py loads the same synthetic state_dict (tests/spade_fixtures.py) into the
UNMODIFIED reference modules, imported from /root/reference, and checks this restatement against them (bit-exact on CPU:
same ATen ops in the same order) before writing tests/golden/spade_*.npz."""
import torch
import torch.nn.functional as F


def conv(sd, p, x, stride=1):
    """nn.Conv2d(kernel_size=3, padding=1)"""
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=1)


def sn_weight(sd, p):
    """effective weight of spectral_norm(conv) in eval mode (no power iteration)"""
    w = sd[p + ".weight_orig"]
    sigma = torch.dot(sd[p + ".weight_u"], torch.mv(w.reshape(w.shape[0], -1), sd[p + ".weight_v"]))
    return w / sigma


def sn_conv(sd, p, x):
    return F.conv2d(x, sn_weight(sd, p), sd[p + ".bias"], padding=1)


def bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False,
                        0.1, 1e-5)


def res_block(sd, p, x, downsample):
    """ref: _init_spade.py:7-38"""
    identity = x
    out = F.relu(bn(sd, p + ".initial.1", conv(sd, p + ".initial.0", x)))
    if downsample:
        identity = conv(sd, p + ".downsample_layer", identity, stride=2)
        out = conv(sd, p + ".residual_downsample", out, stride=2)
    else:
        out = F.relu(bn(sd, p + ".residual.1", conv(sd, p + ".residual.0", out)))
    return out + identity


def id_encoder(sd, p, x):
    """ref: :183-199.  Returns (fid1 [64, H/2], fid2 [128, H/4], fid3 [256, H/8])."""
    x = F.avg_pool2d(conv(sd, p + ".layer1.0", x), 2, stride=2)
    x1 = res_block(sd, p + ".layer2", x, False)
    x2 = res_block(sd, p + ".layer3", x1, True)
    x3 = res_block(sd, p + ".layer4", x2, True)
    return x1, x2, x3


def spade_layer(sd, p, x, fid):
    """ref: :114-139"""
    normalized = F.instance_norm(x, eps=1e-5)
    fid = F.interpolate(fid, size=x.shape[2:], mode="nearest")
    actv = F.relu(conv(sd, p + ".mlp_shared.0", fid))
    gamma = conv(sd, p + ".conv_gamma", actv)
    beta = conv(sd, p + ".conv_beta", actv)
    return normalized * (1 + gamma) + beta


def spade_block(sd, p, x, fid, downsample=False, upsample=False):
    """ref: :235-282"""
    identity = x
    x1 = sn_conv(sd, p + ".conv1", F.leaky_relu(spade_layer(sd, p + ".spade1", x, fid), 0.2))
    if downsample:
        x1 = F.avg_pool2d(x1, 2, stride=2)
        identity = conv(sd, p + ".residual_downsample", identity, stride=2)
    if upsample:
        x1 = F.interpolate(x1, scale_factor=2, mode="nearest")
        identity = F.conv_transpose2d(identity, sd[p + ".residual_upsample.weight"], sd[p + ".residual_upsample.bias"],
                                      stride=2, padding=1, output_padding=1)
    x2 = sn_conv(sd, p + ".conv2", F.leaky_relu(spade_layer(sd, p + ".spade2", x1, fid), 0.2))
    xs = sn_conv(sd, p + ".conv_s", F.leaky_relu(spade_layer(sd, p + ".spade_s", identity, fid), 0.2))
    return xs + x2


def refine_network(sd, p, x, fid1, fid2, fid3, return_intermediates=False):
    """ref: :286-315"""
    inter = {}
    x = F.avg_pool2d(conv(sd, p + ".layer1.0", x), 2, stride=2)
    inter["layer1"] = x
    for name, fid, kw in (("layer2", fid1, dict(downsample=True)), ("layer3", fid2, dict(downsample=True)),
                          ("layer4", fid3, {}), ("layer5", fid3, dict(upsample=True)),
                          ("layer6", fid2, dict(upsample=True)), ("layer7", fid1, dict(upsample=True))):
        x = spade_block(sd, p + "." + name, x, fid, **kw)
        inter[name] = x
    x = conv(sd, p + ".layer8", x)
    return (x, inter) if return_intermediates else x


def generator(sd, i_src, i_raw, return_intermediates=False):
    """Generator.forward, ref: :318-328.  i_src, i_raw: [1, 3, H, W] (H, W multiples of 8)."""
    fid1, fid2, fid3 = id_encoder(sd, "idencoder", i_src)
    return refine_network(sd, "refine_network", i_raw, fid1, fid2, fid3, return_intermediates)


def audio_feature(sd, p, audio):
    """AudioNet.forward of _init_spade.py (:330-357: the conv stack only, LeakyReLU 0.02; encoder_fc1 is unused).
    audio [16, 29] -> [1, 64]"""
    x = audio.unsqueeze(0)[:, 0:16, :].permute(0, 2, 1)
    for i in range(4):
        x = F.leaky_relu(F.conv1d(x, sd[f"{p}.encoder_conv.{2 * i}.weight"], sd[f"{p}.encoder_conv.{2 * i}.bias"], stride=2,
                                  padding=1), 0.02)
    return x.squeeze(-1)


def generator_audio(sd, i_src, i_raw, audio, return_intermediates=False):
    """Generator_audio.forward, ref: :359-373.  The third identity feature is REPLACED by the audio feature: the
    [1, 64] vector is repeated to [1, 256, 64, 64 * 64] (`.unsqueeze(1).repeat(1, 256, 64, 64)` on a 3-d tensor) and every
    SPADE layer then resizes it with nearest interpolation, i.e. channel c of the conditioning map is a horizontal
    pattern of the 64 audio features that depends on the target width (kept as is: it is what the reference computes)."""
    fid1, fid2, _ = id_encoder(sd, "idencoder", i_src)
    a = audio_feature(sd, "AudioNet", audio)
    fid3 = a.unsqueeze(1).repeat(1, 256, 64, 64)
    return refine_network(sd, "refine_network", i_raw, fid1, fid2, fid3, return_intermediates)
