"""viscy_b200: B200-native (sm_100a) implementation of the VisCy convolutional hot path.

Drop-in modules with the reference's constructor / forward / state_dict surface:
`UNeXt2` (VM/unet/unext2.py), `FullyConvolutionalMAE` (VM/unet/fcmae.py), `ContrastiveEncoder` (VM/contrastive/encoder.py), `Unet3d` / `UNet3DBase`
(VM/unet/unet3d*.py), `Unet25d` (VM/unet/unet25d.py), `Unet2d` (VM/unet/unet2d.py); `losses.MixedLoss`
(VU/losses/mixed_loss.py) and `predict.AugmentedPredictionVSUNet` (CY/engine.py).  `patch_viscy()` swaps them into the reference's
architecture registries.
"""

from .contrastive import ContrastiveEncoder  # noqa: F401
from .fcmae import FullyConvolutionalMAE  # noqa: F401
from .unet2d import ConvBlock2D, Unet2d  # noqa: F401
from .unet25d import ConvBlock3D, Unet25d  # noqa: F401
from .unet3d import UNet3DBase, Unet3d  # noqa: F401
from .unext2 import UNeXt2  # noqa: F401

__all__ = ["UNeXt2", "FullyConvolutionalMAE", "ContrastiveEncoder", "Unet3d", "UNet3DBase", "Unet25d", "ConvBlock3D", "Unet2d", "ConvBlock2D"]


def patch_viscy() -> list[str]:
    """Swap the reference's architecture registry entries for the B200-native classes.

    Touches only modules that are importable; returns the list of patched registries."""
    import importlib

    patched = []
    for mod_name, attr in (("cytoland.engine", "_UNET_ARCHITECTURE"), ("dynacell.engine", "_ARCHITECTURE")):
        try:
            mod = importlib.import_module(mod_name)
        except Exception:
            continue
        reg = getattr(mod, attr, None)
        if isinstance(reg, dict) and "UNeXt2" in reg:
            reg["UNeXt2"] = UNeXt2
            for key, cls in (("FNet3D", Unet3d), ("2.5D", Unet25d), ("2D", Unet2d), ("fcmae", FullyConvolutionalMAE),
                             ("UNeXt2_2D", FullyConvolutionalMAE)):
                if key in reg:
                    reg[key] = cls
            patched.append(f"{mod_name}.{attr}")
    return patched


__all__.append("patch_viscy")
