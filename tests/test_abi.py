"""The C-ABI library loads without a GPU and exports every symbol include/xspect_b200.h declares."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols() -> set[str]:
    text = (ROOT / "include/xspect_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(xs_[a-z0-9_]+)\s*\(", text))


def test_header_symbols_are_exported_and_bound():
    from xspect2_b200 import _abi
    L = _abi.lib()
    decl = declared_symbols()
    assert decl, "no declarations parsed"
    for name in decl:
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert decl == set(_abi.SYMBOLS), decl ^ set(_abi.SYMBOLS)
    assert L.xs_version() >= 100


def test_argument_errors_without_gpu():
    from xspect2_b200 import _abi
    L = _abi.lib()
    h = C.c_void_p()
    assert L.xs_cobs_open(b"/nonexistent/index.cobs_classic", 0, 0, 0, C.byref(h)) == _abi.XS_ERR_IO
    assert b"nonexistent" in L.xs_last_error()
    with pytest.raises(FileNotFoundError):
        _abi.check(_abi.XS_ERR_IO)
    assert L.xs_bloom_open(b"/nonexistent/filter.bloom", 21, 0, C.byref(h)) == _abi.XS_ERR_IO
    assert L.xs_cobs_query(None, None, 0, None, None, 0, 1, 4, None) == _abi.XS_ERR_ARG


def test_format_errors_without_gpu(tmp_path):
    from xspect2_b200 import _abi
    L = _abi.lib()
    p = tmp_path / "bad.cobs_classic"
    p.write_bytes(b"NOTCOBS" * 10)
    h = C.c_void_p()
    assert L.xs_cobs_open(str(p).encode(), 0, 0, 0, C.byref(h)) == _abi.XS_ERR_FORMAT
    with pytest.raises(ValueError):
        _abi.check(_abi.XS_ERR_FORMAT)


def test_result_order_is_host_only():
    import numpy as np
    from xspect2_b200.engine import CobsIndex
    s = np.array([3, 9, 9, 0, 3, 1], np.uint32)
    o = CobsIndex.result_order(s)
    assert sorted(o.tolist()) == list(range(6))
    assert all(s[o[i]] >= s[o[i + 1]] for i in range(5))


def test_no_oracle_import_in_product():
    for f in (ROOT / "xspect2_b200").rglob("*.py"):
        t = f.read_text()
        assert "import oracle" not in t and "from oracle" not in t, f
    for f in (ROOT / "xspect2_b200").rglob("*.cu*"):
        assert "oracle/" not in f.read_text().replace("the oracle", ""), f
