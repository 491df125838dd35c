"""viscy_b200: B200-native (sm_100a) implementation of the VisCy convolutional hot path.

Drop-in modules with the reference's constructor / forward / state_dict surface:
`UNeXt2` (VM/unet/unext2.py), `ContrastiveEncoder` (VM/contrastive/encoder.py), `Unet3d` / `UNet3DBase`
(VM/unet/unet3d*.py), `Unet25d` (VM/unet/unet25d.py).  `patch_viscy()` swaps them into the reference's
architecture registries.
"""

from .unext2 import UNeXt2  # noqa: F401

__all__ = ["UNeXt2"]
