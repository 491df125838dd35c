"""The C-ABI library loads and exports every symbol include/viscy_b200.h declares (no compute calls: no GPU here)."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    from viscy_b200 import build
    lib_path = build.build()
    lib = ctypes.CDLL(str(lib_path))
    header = (ROOT / "include" / "viscy_b200.h").read_text()
    names = sorted(set(re.findall(r"\b(?:int|int64_t)\s+(vb200_\w+)\s*\(", header)))
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.vb200_abi_version() >= 1


def test_error_reporting_without_gpu():
    from viscy_b200 import _lib
    lib = _lib.lib()
    assert lib.vb200_gemm(None, None) == _lib.ERR_INVALID
    assert "null descriptor" in _lib.last_error()


def test_gemm_desc_layout_matches_header():
    """ctypes mirror of vb200_gemm_desc: field order as in the header."""
    from viscy_b200 import _lib
    header = (ROOT / "include" / "viscy_b200.h").read_text()
    body = header[header.index("typedef struct vb200_gemm_desc {"):header.index("} vb200_gemm_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in re.findall(r"(?:int32_t|int64_t|const void\*|void\*|const float\*|float\*)\s+([^;]+);", body):
        fields += [f.strip() for f in decl.split(",")]
    assert fields == [f[0] for f in _lib.GemmDesc._fields_]


def test_conv3d_desc_layout_matches_header():
    """ctypes mirror of vb200_conv3d_desc: field order as in the header."""
    from viscy_b200 import _lib
    header = (ROOT / "include" / "viscy_b200.h").read_text()
    body = header[header.index("typedef struct vb200_conv3d_desc {"):header.index("} vb200_conv3d_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in re.findall(r"(?:const int32_t\*|int32_t|int64_t|const void\*|void\*|const float\*|float\*)\s+([^;]+);", body):
        fields += [re.sub(r"\[\d+\]", "", f).strip() for f in decl.split(",")]
    assert fields == [f[0] for f in _lib.Conv3dDesc._fields_]
    import ctypes as C
    # natural C layout: 24 int32 (one of them alignment padding), 6 int64, 8 pointers
    assert C.sizeof(_lib.Conv3dDesc) == 96 + 48 + 64


def test_conv3d_igemm_geometry_query_is_host_only():
    """vb200_conv3d_igemm_supported: 128 (forward) / 64 (wgrad) consecutive output voxels must form a TMA box."""
    from viscy_b200 import ops
    k3, p1 = (3, 3, 3), (1, 1, 1)
    for shape in [(1, 128, 128, 128, 32), (1, 64, 64, 64, 64), (1, 8, 8, 8, 512), (2, 4, 8, 8, 64), (1, 4, 4, 256, 64)]:
        assert ops.conv3d_igemm_supported(shape, 64, k3, p1), shape
        assert ops.conv3d_igemm_supported(shape, 64, k3, p1, wgrad=True), shape
    assert not ops.conv3d_igemm_supported((1, 7, 12, 20, 64), 64, k3, p1)          # 20-wide rows do not tile 128
    assert not ops.conv3d_igemm_supported((1, 8, 8, 8, 16), 64, k3, p1)            # forward needs 32-channel granules
    assert ops.conv3d_igemm_supported((1, 8, 8, 8, 16), 64, k3, p1, wgrad=True)    # the weight gradient does not
    assert ops.conv3d_igemm_supported((1, 23, 128, 128, 32), 32, k3, (0, 1, 1))    # valid in Z (UNeXt2 head geometry)
    assert not ops.conv3d_igemm_supported((1, 8, 8, 8, 64), 20, k3, p1)            # cout % 8
