"""Synthetic Stage-II (SPADE generator) fixtures: a deterministic state_dict with the reference's keys and shapes
(tests/golden/spade_keys.json, written by oracle/make_golden_spade.py from the live reference modules) and seeded
inputs.  The weights are 17.3 M parameters, so they are regenerated from the seed instead of being committed."""
import json
import os

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = os.path.join(REPO, "tests", "golden", "spade_keys.json")


def load_keys(kind: str) -> dict:
    """kind: "generator" | "generator_audio" -> {key: shape list} in the reference's state_dict order."""
    with open(KEYS) as f:
        return json.load(f)[kind]


def make_state_dict(kind: str, seed: int = 0) -> dict:
    """He-scaled conv weights, small biases, non-trivial BatchNorm statistics and spectral-norm vectors: every folding
    step (BN into conv, weight_orig / sigma) changes the result.  The two names of a spectral-normed conv
    (`conv1` / `conv1_sn`, the same module in the reference) get the same tensors."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape in load_keys(kind).items():
        alias = key.replace("_sn.", ".")
        if alias != key and alias in sd:
            sd[key] = sd[alias]
            continue
        leaf = key.rsplit(".", 1)[1]
        if leaf == "num_batches_tracked":
            t = torch.tensor(100, dtype=torch.int64)
        elif leaf == "running_var":
            t = 0.5 + torch.rand(shape, generator=gen)
        elif leaf == "running_mean":
            t = 0.1 * torch.randn(shape, generator=gen)
        elif leaf in ("weight_u", "weight_v"):
            # what power iteration converges to in training: the leading singular pair of weight_orig (a random pair would
            # make sigma = u.Wv tiny and the "normalised" weights huge), tilted a little so that sigma is not exactly s_max
            w = sd[key.rsplit(".", 1)[0] + ".weight_orig"]
            U, _, Vh = torch.linalg.svd(w.reshape(w.shape[0], -1), full_matrices=False)
            t = (U[:, 0] if leaf == "weight_u" else Vh[0]) + 0.05 * torch.randn(shape, generator=gen) / shape[0] ** 0.5
            t = t / t.norm()
        elif leaf in ("weight", "weight_orig") and len(shape) >= 3:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = torch.randn(shape, generator=gen) * (1.4 / fan_in ** 0.5)
            if ".conv_gamma." in key or ".conv_beta." in key:
                t = t * 0.5
        elif leaf == "weight" and len(shape) == 2:
            t = torch.randn(shape, generator=gen) * (1.0 / shape[1] ** 0.5)
        elif leaf == "weight":                       # BatchNorm scale
            t = 0.5 + torch.rand(shape, generator=gen)
        else:                                        # biases
            t = 0.1 * torch.randn(shape, generator=gen)
        sd[key] = t
    return sd


def make_inputs(H: int, W: int, seed: int = 0) -> dict:
    """identity photo, Stage-I frame (both [1,3,H,W] in [0,1], smooth + noise) and a DeepSpeech window [16,29]."""
    gen = torch.Generator().manual_seed(1000 + seed)
    ys = torch.linspace(0, 1, H)[:, None].expand(H, W)
    xs = torch.linspace(0, 1, W)[None, :].expand(H, W)
    base = torch.stack((ys, xs, 0.5 * (ys + xs)))
    i_src = (0.6 * base + 0.4 * torch.rand(3, H, W, generator=gen)).unsqueeze(0).contiguous()
    i_raw = (0.5 * base.flip(0) + 0.5 * torch.rand(3, H, W, generator=gen)).unsqueeze(0).contiguous()
    audio = torch.randn(16, 29, generator=gen)
    return {"i_src": i_src, "i_raw": i_raw, "audio": audio}
