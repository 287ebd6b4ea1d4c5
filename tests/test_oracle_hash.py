"""Pins the oracle's XXH64 / XXH3-64 and the device-side (__host__ __device__) hash, canonical-form, LCG and
Barrett functions against python-xxhash known answers (tests/golden/xxhash_kat.json) — no GPU needed."""
import ctypes as C
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
KAT = json.loads((ROOT / "tests/golden/xxhash_kat.json").read_text())["vectors"]


@pytest.fixture(scope="module")
def hc():
    so = ROOT / "tests/native/libxs_hostcheck.so"
    src = ROOT / "tests/native/xs_hostcheck.cu"
    dev = ROOT / "xspect2_b200/csrc/xs_device.cuh"
    if not so.exists() or so.stat().st_mtime < max(src.stat().st_mtime, dev.stat().st_mtime):
        subprocess.run(["nvcc", "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-o", str(so), str(src)], check=True)
    L = C.CDLL(str(so))
    L.hc_xxh64.restype = C.c_uint64
    L.hc_xxh64.argtypes = [C.c_char_p, C.c_uint32, C.c_uint64]
    L.hc_xxh3.restype = C.c_uint64
    L.hc_xxh3.argtypes = [C.c_char_p, C.c_uint32]
    L.hc_canonical.restype = C.c_uint64
    L.hc_canonical.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_char_p, C.POINTER(C.c_int)]
    L.hc_literal.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_int, C.c_char_p]
    L.hc_lcg.argtypes = [C.c_uint64, C.c_uint32, C.c_void_p]
    L.hc_mod.restype = C.c_uint64
    L.hc_mod.argtypes = [C.c_uint64, C.c_uint64]
    L.hc_mod_small.restype = C.c_uint32
    L.hc_mod_small.argtypes = [C.c_uint64, C.c_uint32]
    L.hc_biocomp.restype = C.c_uint8
    L.hc_biocomp.argtypes = [C.c_uint8]
    L.hc_cobscomp.restype = C.c_uint8
    L.hc_cobscomp.argtypes = [C.c_uint8]
    return L


def test_survey_known_answers(oracle):
    # SURVEY.md A.5 (verified there with python-xxhash 3.7.0)
    assert oracle.xxh64(b"", 0) == 0xEF46DB3751D8E999
    assert oracle.xxh3_64(b"") == 0x2D06800538D394C2
    kmer = b"AGAGATTACGTCTGGTTGCAA"
    exp = [0x17B5F0EF3E7A1BE1, 0x2D30DBFB1F27A784, 0xC771E587E0BB0B24, 0xC3254F0823C2F3C0,
           0x6FA7F711E5572DE0, 0x0973C2601C8CE8F1, 0xB7175C98A250C276]
    assert [oracle.xxh64(kmer, j) for j in range(7)] == exp
    assert oracle.xxh3_64(kmer) == 11849584422377248794
    assert oracle.xxh3_64(b"TAAATAAATTTATATAGCTAA") == 0x8CE889D5DA5A0CC9
    assert oracle.xxh3_64(b"AAATAAATTTATATAGCTAAA") == 0xB659ED5B07C7CEE3


def test_oracle_hashes_match_golden(oracle):
    for v in KAT:
        d = v["data"].encode()
        for s, h in v["xxh64"].items():
            assert oracle.xxh64(d, int(s)) == h, (d, s)
        assert oracle.xxh3_64(d) == v["xxh3_64"], d


def test_oracle_hashes_match_live_xxhash(oracle):
    xxhash = pytest.importorskip("xxhash")
    rng = np.random.default_rng(5)
    for _ in range(300):
        n = int(rng.integers(0, 129))
        d = bytes(rng.integers(0, 256, size=n, dtype=np.uint8))
        s = int(rng.integers(0, 2**63))
        assert oracle.xxh64(d, s) == xxhash.xxh64_intdigest(d, seed=s)
        assert oracle.xxh3_64(d) == xxhash.xxh3_64_intdigest(d)


def test_device_hash_functions_on_host(hc):
    for v in KAT:
        d = v["data"].encode()
        if not 1 <= len(d) <= 32:
            continue
        for s, h in v["xxh64"].items():
            assert hc.hc_xxh64(d, len(d), int(s)) == h, (d, s)
        assert hc.hc_xxh3(d, len(d)) == v["xxh3_64"], d


def test_device_canonical_and_expand_on_host(hc, oracle):
    rng = np.random.default_rng(7)
    for k in (1, 2, 7, 8, 15, 16, 17, 21, 24, 31, 32):
        seq = bytes(np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=200)])
        for g in range(0, 200 - k + 1, 3):
            out = C.create_string_buffer(32)
            inv = C.c_int()
            msb = hc.hc_canonical(seq, len(seq), g, k, out, C.byref(inv))
            assert inv.value == 0
            exp = oracle.cobs_term(seq[g:g + k])
            assert out.raw[:k] == exp
            code = 0
            for ch in exp:
                code = (code << 2) | b"ACGT".index(ch)
            assert msb == code
    # invalid bitmap
    seq = b"ACGTNACGTACGTACGTACGTACGTAaGTACGTACGTACGTACGTACGTACGTACGTACGT"
    for g in range(len(seq) - 21 + 1):
        out = C.create_string_buffer(32)
        inv = C.c_int()
        hc.hc_canonical(seq, len(seq), g, 21, out, C.byref(inv))
        assert (inv.value & 1) == int(any(c not in b"ACGT" for c in seq[g:g + 21]))


def test_device_literal_terms_on_host(hc, oracle):
    rng = np.random.default_rng(11)
    alphabet = np.frombuffer(b"ACGTNacgtnRYKMSWBDHVUu-", np.uint8)
    table = oracle.bio_complement_table()
    for c in range(256):
        assert hc.hc_biocomp(c) == table[c]
        assert hc.hc_cobscomp(c) == {65: 84, 67: 71, 71: 67, 84: 65}.get(c, 0)
    for k in (5, 21, 31, 32):
        seq = bytes(alphabet[rng.integers(0, alphabet.size, size=120)])
        for g in range(0, 120 - k + 1):
            out = C.create_string_buffer(32)
            hc.hc_literal(seq, g, k, None, 1, out)
            assert out.raw[:k] == oracle.bloom_term(seq[g:g + k])
            cobs = oracle.cobs_term(seq[g:g + k], 1, oracle.POLICY_LITERAL)
            tbl = bytes({65: 84, 67: 71, 71: 67, 84: 65}.get(c, 0) for c in range(256))
            hc.hc_literal(seq, g, k, tbl, 1, out)
            assert out.raw[:k] == cobs


def test_device_lcg_and_barrett_on_host(hc):
    M = 47026247687942121848144207491837418733
    rng = np.random.default_rng(3)
    for _ in range(50):
        h0 = int(rng.integers(0, 2**63)) * 2 + int(rng.integers(0, 2))
        out = np.zeros(8, np.uint64)
        hc.hc_lcg(h0, 8, out.ctypes.data)
        st = h0
        for i in range(8):
            st = (st * M + 1) % (1 << 128)
            assert int(out[i]) == (st >> 32) % (1 << 64)
    for _ in range(2000):
        x = int(rng.integers(0, 2**63)) * 2 + int(rng.integers(0, 2))
        m = int(rng.integers(1, 2**62)) if rng.random() < 0.5 else int(rng.integers(1, 2**20))
        assert hc.hc_mod(x, m) == x % m
    for m in (1, 2, 3, 2**32, 2**32 + 1, 150000001, 13800000008):
        for x in (0, 1, m - 1, m, m + 1, 2**64 - 1):
            assert hc.hc_mod(x, m) == x % m
    # the 32-bit remainder variant of the bucketed kernels (signature_size < 2^31)
    for _ in range(4000):
        x = int(rng.integers(0, 2**63)) * 2 + int(rng.integers(0, 2))
        m = int(rng.integers(1, 2**31)) if rng.random() < 0.7 else int(rng.integers(1, 2**12))
        assert hc.hc_mod_small(x, m) == x % m
    for m in (1, 2, 3, 2**31 - 1, 2**30, 150000001, 20000003):
        for x in (0, 1, m - 1, m, m + 1, 2**32 - 1, 2**32, 2**64 - 1, 2**64 - m):
            assert hc.hc_mod_small(x, m) == x % m
