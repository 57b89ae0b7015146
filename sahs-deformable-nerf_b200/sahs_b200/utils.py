"""Label / colour helpers of the reference's `nerf.utils` (ref: nerf/utils.py:5-26 shrink, :29-70 color2label_np,
:112-140 label2color), written as table look-ups.  Host-side data preparation: the renderer's own palette output is
`sahs_b200.frame_postprocess` (CUDA)."""
from __future__ import annotations

import numpy as np
import torch

NUM_CLASSES = 12
# RGB palette of the 12 semantic classes (background, face, nose, glasses, eye, brow, ear, inner mouth, lips, hair,
# neck, torso) -- ref: nerf/utils.py:31-46
PALETTE = np.array([[0, 0, 0], [204, 0, 0], [76, 153, 0], [204, 204, 0], [51, 51, 255], [0, 255, 255], [102, 51, 0],
                    [102, 204, 0], [255, 255, 0], [0, 0, 204], [255, 153, 51], [0, 204, 0]], dtype=np.int32)


def shrink(mask: np.ndarray) -> np.ndarray:
    """[H,W,C] scores -> one-hot int32 [H,W,12] of the arg-max class (ref: nerf/utils.py:5-26)."""
    label = np.argmax(mask, axis=-1)
    return np.eye(NUM_CLASSES, dtype=np.int32)[np.clip(label, 0, NUM_CLASSES - 1)] * (label < NUM_CLASSES)[..., None]


def color2label_np(target: np.ndarray) -> np.ndarray:
    """[H,W,3] palette-coloured parsing map -> one-hot int32 [H,W,12]; pixels of no palette colour stay all-zero
    (ref: nerf/utils.py:29-70)."""
    hit = np.all(np.asarray(target)[:, :, None, :] == PALETTE[None, None, :, :], axis=-1)      # [H,W,12]
    return hit.astype(np.int32)


def label2color(mask: torch.Tensor) -> torch.Tensor:
    """[H,W,12] class scores -> float32 [H,W,3] in [0,1], palette entries written in reversed channel order exactly as
    the reference does (ref: nerf/utils.py:112-140); always a CPU tensor, as there."""
    label = torch.argmax(mask, dim=-1).cpu()
    table = torch.from_numpy(PALETTE[:, ::-1].copy()).to(torch.float32)
    color = torch.zeros(label.shape + (3,), dtype=torch.float32)
    ok = label < NUM_CLASSES
    color[ok] = table[label[ok]]
    return color / 255.
