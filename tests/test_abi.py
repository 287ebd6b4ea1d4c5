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


def _probe(path):
    from xspect2_b200 import _abi
    hd = _abi.CobsHeader()
    rc = _abi.lib().xs_cobs_probe_header(str(path).encode(), C.byref(hd))
    return rc, hd


def test_header_candidate_layouts_and_compact_padding(tmp_path):
    """SURVEY.md A.1.3 / ADVICE r1: the header layout of cobs-reloaded cannot be verified offline, so the reader tries
    the documented field order first and neighbouring ones after it, and accepts a reading only when the end magic and
    the size identity hold; a compact file whose data is already page-aligned may carry 0 or page_size padding bytes."""
    import struct
    from oracle import oracle
    from xspect2_b200 import _abi
    names = [f"doc{i}" for i in range(11)]
    sig, h, k = 37, 7, 21
    data = bytes(range(256)) * 10
    data = data[: sig * 2]
    doc = oracle.classic_header(k, 1, names, sig, h) + data
    (tmp_path / "a.cobs_classic").write_bytes(doc)
    rc, hd = _probe(tmp_path / "a.cobs_classic")
    assert rc == 0 and hd.layout == b"cobs v1 (documented)" and (hd.n_docs, hd.num_hashes, hd.sig_size_max, hd.term_size) == (11, 7, sig, k)
    assert hd.file_size - hd.data_offset == sig * 2
    # num_hashes and signature_size swapped
    alt = b"COBS:CLASSIC_INDEX" + struct.pack("<IIB", 1, k, 1) + struct.pack("<I", 11) + struct.pack("<QQ", h, sig)
    alt += b"".join(n.encode() + b"\n" for n in names) + b"CLASSIC_INDEX" + data
    (tmp_path / "b.cobs_classic").write_bytes(alt)
    rc, hd = _probe(tmp_path / "b.cobs_classic")
    assert rc == 0 and hd.layout == b"num_hashes before signature_size" and (hd.num_hashes, hd.sig_size_max) == (7, sig)
    # 32-bit num_hashes
    alt = b"COBS:CLASSIC_INDEX" + struct.pack("<IIB", 1, k, 1) + struct.pack("<I", 11) + struct.pack("<QI", sig, h)
    alt += b"".join(n.encode() + b"\n" for n in names) + b"CLASSIC_INDEX" + data
    (tmp_path / "c.cobs_classic").write_bytes(alt)
    rc, hd = _probe(tmp_path / "c.cobs_classic")
    assert rc == 0 and hd.layout == b"32-bit num_hashes" and (hd.num_hashes, hd.sig_size_max) == (7, sig)
    # a truncated data section fails under every reading, and the message is the documented layout's
    (tmp_path / "d.cobs_classic").write_bytes(doc[:-3])
    rc, _ = _probe(tmp_path / "d.cobs_classic")
    assert rc == _abi.XS_ERR_FORMAT and b"size identity" in _abi.lib().xs_last_error()

    # compact: page_size 1 (every locus with < 32 alleles) and page sizes where the remainder is 0
    for page_size in (1, 2, 3, 8):
        for n_names in range(3, 12):
            nm = [f"a{i}" for i in range(n_names)]
            n_pages = (n_names + 8 * page_size - 1) // (8 * page_size)
            sigs = [19 + p for p in range(n_pages)]
            body = b"\x5a" * (sum(sigs) * page_size)
            for full in (False, True):
                head = oracle.compact_header(31, 1, nm, page_size, sigs, 1, pad_full_page=full)
                f = tmp_path / f"c_{page_size}_{n_names}_{int(full)}.cobs_compact"
                f.write_bytes(head + body)
                rc, hd = _probe(f)
                assert rc == 0, (_abi.lib().xs_last_error(), page_size, n_names, full)
                assert hd.data_offset == len(head) and hd.n_pages == n_pages and hd.page_bytes == page_size and hd.n_docs == n_names
                assert oracle.parse_cobs(f)["data_off"] == len(head)
