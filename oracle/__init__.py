"""ORACLE -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's algorithm for the VisCy convolutional hot path, used as the
checker in `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / --impl reference leg.
Nothing under `viscy_b200/` imports this package; the product path has no CPU or oracle fallback.

Layers
  * `ref_timm.py`, `ref_monai.py`  -- restatements of the third-party pieces the reference composes and
    that are NOT present under /root/reference: timm 1.0.27 ConvNeXt / ConvNeXt-V2 (uv.lock:6269-6270),
    monai 1.5.2 UpSample(pixelshuffle) / Convolution+ADN (uv.lock:3358-3359).  Semantics: SURVEY.md App. B.
  * `reference_loader.py`          -- (authoring container only) executes the reference's own, unmodified
    composition code from /root/reference (unext2.py, blocks.py, heads.py, stems.py, encoder.py, unet25d.py,
    unet3d*.py, conv_block_3d.py) against real timm/monai when importable, else against the restatements.
  * `models.py`                    -- self-contained functional restatement that travels to the GPU box
    (no /root/reference at run time): unext2 / contrastive encoder / unet25d / unet3d forward from a
    reference-layout state_dict.  Pinned here against `reference_loader` outputs (tests/golden/).

Parity status: Unet25d / Unet3d are pinned against the reference's own code run here.  UNeXt2 /
ContrastiveEncoder are pinned against the reference's composition code running over the *restated*
timm/monai blocks: the reference's tests hold no numeric goldens for them, so at the
timm/monai boundary the parity is UNPINNED by the reference's own tests; key counts (213/273/194) and
sentinel keys from packages/viscy-models/tests/test_state_dict_compat.py are checked instead.
"""
