"""Execute the reference's own hot-path files from /root/reference (AUTHORING CONTAINER ONLY; TEST INFRASTRUCTURE).

`load()` returns a namespace with the reference classes.  The package `viscy_models.__init__` cannot be imported
(importlib.metadata + timm/monai/...), so lightweight package shells are registered and each file is
loaded by path, unmodified.  Real timm/monai are used when importable; otherwise oracle/ref_timm.py / ref_monai.py.
"""
from __future__ import annotations

import importlib
import importlib.util
import sys
import types
from pathlib import Path

REF_SRC = Path("/root/reference/packages/viscy-models/src/viscy_models")
_FILES = [
    ("viscy_models.schedule", "schedule.py"),
    ("viscy_models.components.conv_block_3d", "components/conv_block_3d.py"),
    ("viscy_models.components.conv_block_2d", "components/conv_block_2d.py"),
    ("viscy_models.unet.unet2d", "unet/unet2d.py"),
    ("viscy_models.components.stems", "components/stems.py"),
    ("viscy_models.components.blocks", "components/blocks.py"),
    ("viscy_models.components.heads", "components/heads.py"),
    ("viscy_models.unet.blocks", "unet/blocks.py"),
    ("viscy_models.unet.unet3d_base", "unet/unet3d_base.py"),
    ("viscy_models.unet.unet3d", "unet/unet3d.py"),
    ("viscy_models.unet.unet25d", "unet/unet25d.py"),
    ("viscy_models.unet.unext2", "unet/unext2.py"),
    ("viscy_models.unet.fcmae", "unet/fcmae.py"),
    ("viscy_models.contrastive.encoder", "contrastive/encoder.py"),
]
_ns = None


def available() -> bool:
    return REF_SRC.exists()


def _ensure_third_party() -> dict:
    info = {}
    try:
        import timm  # noqa: F401
        info["timm"] = "real"
    except Exception:
        from . import ref_timm
        t = ref_timm.as_module()
        sys.modules["timm"] = t
        sys.modules["timm.layers"] = t.layers
        sys.modules["timm.models"] = t.models
        sys.modules["timm.models.convnext"] = t.models.convnext
        info["timm"] = "restated"
    try:
        import monai  # noqa: F401
        info["monai"] = "real"
    except Exception:
        from . import ref_monai
        sys.modules.update(ref_monai.as_modules())
        info["monai"] = "restated"
    return info


def load() -> types.SimpleNamespace:
    global _ns
    if _ns is not None:
        return _ns
    if not available():
        raise RuntimeError("/root/reference is not present (GPU box): use tests/golden fixtures + oracle.models instead")
    info = _ensure_third_party()
    for pkg in ("viscy_models", "viscy_models.components", "viscy_models.unet", "viscy_models.contrastive"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = [str(REF_SRC / "/".join(pkg.split(".")[1:]))]
            sys.modules[pkg] = m
    mods = {}
    for name, rel in _FILES:
        spec = importlib.util.spec_from_file_location(name, REF_SRC / rel)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    _ns = types.SimpleNamespace(
        third_party=info,
        UNeXt2=mods["viscy_models.unet.unext2"].UNeXt2,
        Unet25d=mods["viscy_models.unet.unet25d"].Unet25d,
        Unet3d=mods["viscy_models.unet.unet3d"].Unet3d,
        UNet3DBase=mods["viscy_models.unet.unet3d_base"].UNet3DBase,
        ConvBlock3D=mods["viscy_models.components.conv_block_3d"].ConvBlock3D,
        UNeXt2Stem=mods["viscy_models.components.stems"].UNeXt2Stem,
        StemDepthtoChannels=mods["viscy_models.components.stems"].StemDepthtoChannels,
        UNeXt2Decoder=mods["viscy_models.components.blocks"].UNeXt2Decoder,
        UNeXt2UpStage=mods["viscy_models.components.blocks"].UNeXt2UpStage,
        PixelToVoxelHead=mods["viscy_models.components.heads"].PixelToVoxelHead,
        ContrastiveEncoder=mods["viscy_models.contrastive.encoder"].ContrastiveEncoder,
        ResnetBlock=mods["viscy_models.unet.blocks"].ResnetBlock,
        Unet2d=mods["viscy_models.unet.unet2d"].Unet2d,
        ConvBlock2D=mods["viscy_models.components.conv_block_2d"].ConvBlock2D,
        FullyConvolutionalMAE=mods["viscy_models.unet.fcmae"].FullyConvolutionalMAE,
        fcmae=mods["viscy_models.unet.fcmae"],
        PixelToVoxelShuffleHead=mods["viscy_models.components.heads"].PixelToVoxelShuffleHead,
    )
    return _ns


_loss_ns = None


def load_losses() -> types.SimpleNamespace:
    """The reference's own `MixedLoss` / `ms_ssim_25d` / `ssim_25d` (viscy-utils), executed unmodified.  metrics.py imports
    scipy (present) and skimage / torchmetrics / torchvision symbols that the SSIM functions never touch: absent packages
    are replaced by empty shells carrying just those names."""
    global _loss_ns
    if _loss_ns is not None:
        return _loss_ns
    if not available():
        raise RuntimeError("/root/reference is not present (GPU box): use tests/golden fixtures instead")
    src = Path("/root/reference/packages/viscy-utils/src/viscy_utils")

    def shell(name, **attrs):
        try:
            return importlib.import_module(name)
        except Exception:
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            sys.modules[name] = m
            return m

    unused = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("not part of the SSIM path"))  # noqa: E731
    shell("skimage")
    shell("skimage.measure", label=unused, regionprops=unused)
    shell("torchmetrics")
    shell("torchmetrics.detection")
    shell("torchmetrics.detection.mean_ap", MeanAveragePrecision=unused)
    shell("torchvision")
    shell("torchvision.ops", masks_to_boxes=unused)
    for pkg, sub in (("viscy_utils", ""), ("viscy_utils.evaluation", "evaluation"), ("viscy_utils.losses", "losses")):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = [str(src / sub)]
            sys.modules[pkg] = m
    mods = {}
    for name, rel in (("viscy_utils.evaluation.metrics", "evaluation/metrics.py"),
                      ("viscy_utils.losses.mixed_loss", "losses/mixed_loss.py")):
        spec = importlib.util.spec_from_file_location(name, src / rel)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    met = mods["viscy_utils.evaluation.metrics"]
    _loss_ns = types.SimpleNamespace(MixedLoss=mods["viscy_utils.losses.mixed_loss"].MixedLoss,
                                     ms_ssim_25d=met.ms_ssim_25d, ssim_25d=met.ssim_25d)
    return _loss_ns
