"""Drop-in for the reference's nerf/_init_spade.py as the Stage-II *inference* scripts use it
(eval_get_texture_photo_audio.py:22,36,154-159 and eval_get_texture_photo_3dmm.py:36,125-128: `from nerf._init_spade import *`,
`G = Generator().to(device)` / `Generator_audio()`, `G.load_state_dict(checkpoint["model_state_dict"])`, `G.eval()`,
`G(frame, image[, driving_data])`).  The generators are the B200-native ones (sahs_b200/spade.py); the training-only classes
(Discriminator, VGG) are not provided and raise when touched."""
from sahs_b200.spade import (AudioNet, Generator, Generator_audio, GraphedGenerator, IdEncoder, RefineNetwork,  # noqa: F401
                             ResBlock2d, SPADEBlock, SPADELayer)

__all__ = ["AudioNet", "Generator", "Generator_audio", "IdEncoder", "RefineNetwork", "ResBlock2d", "SPADEBlock", "SPADELayer",
           "Discriminator", "VGG"]


def _training_only(name):
    class _Missing:
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name} belongs to Stage-II training, which the B200 path does not provide "
                                      "(inference only: Generator / Generator_audio)")
    _Missing.__name__ = name
    return _Missing


Discriminator = _training_only("Discriminator")
VGG = _training_only("VGG")
