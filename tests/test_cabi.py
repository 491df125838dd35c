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
    for decl in re.findall(r"(?:int32_t|int64_t|const void\*|void\*|const float\*)\s+([^;]+);", body):
        fields += [f.strip() for f in decl.split(",")]
    assert fields == [f[0] for f in _lib.GemmDesc._fields_]
