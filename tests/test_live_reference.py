"""Opt-in live differential tests against the real third-party libraries the reference calls
(`cobs_index` from cobs-reloaded, `rbloom`).  They are not installable offline, so these tests skip here; on a
box that has them they pin every [UNVERIFIED-3P] assumption of SURVEY.md Appendix A in one go:
the oracle must read files written by the real libraries and return the real libraries' answers."""
import numpy as np
import pytest

from tests import synth


def test_cobs_classic_against_real_cobs(tmp_path, oracle):
    cobs = pytest.importorskip("cobs_index")
    rng = np.random.default_rng(1)
    docs = synth.make_genomes(rng, 5, 3000)
    d = tmp_path / "docs"
    d.mkdir()
    for name, (g,) in docs.items():
        (d / f"{name}.fasta").write_text(f">{name}\n{g.tobytes().decode()}\n")
    doclist = cobs.DocumentList()
    for f in sorted(d.iterdir()):
        doclist.add(str(f))
    params = cobs.ClassicIndexParameters()
    params.term_size, params.num_hashes, params.false_positive_rate, params.clobber = 21, 7, 0.01, True
    path = tmp_path / "index.cobs_classic"
    cobs.classic_construct_list(doclist, str(path), params)       # probabilistic_filter_model.py:186-192
    real = cobs.Search(str(path), True)
    orc = oracle.CobsOracle(path)                                  # header layout, bit order
    bases, b, e = synth.sample_reads(rng, [g for (g,) in docs.values()], 200, (22, 400), sub=0.01, n_rate=0.003)
    for x, y in zip(b, e):
        q = bases[int(x):int(y)].tobytes().decode()
        for step in (1, 3):
            got = [(r.doc_name, r.score) for r in orc.search(q, step)]
            exp = [(r.doc_name, r.score) for r in real.search(q, step=step)]   # hash seeds, N handling, result order
            assert got == exp, q


def test_bloom_against_real_rbloom(tmp_path, oracle):
    rbloom = pytest.importorskip("rbloom")
    xxhash = pytest.importorskip("xxhash")
    rng = np.random.default_rng(2)
    g = synth.random_dna(rng, 5000).tobytes().decode()
    bf = rbloom.Bloom(len(g) - 21 + 1, 0.01, hash_func=xxhash.xxh3_64_intdigest)   # probabilistic_single_filter_model.py:88
    for i in range(len(g) - 20):
        bf.add(oracle.bloom_term(g[i:i + 21].encode()).decode())
    p = tmp_path / "filter.bloom"
    bf.save(str(p))
    orc = oracle.BloomOracle(p, 21)                                # file layout, k, LCG constants
    for i in range(0, len(g) - 20, 7):
        assert g[i:i + 21] in orc or oracle.bloom_term(g[i:i + 21].encode()).decode() in orc
    for _ in range(2000):
        km = synth.random_dna(rng, 21).tobytes().decode()
        assert (km in orc) == (km in bf)


def test_reference_flows_with_the_real_wheels(tmp_path):
    """The committed model / workflow / CLI golden vectors were produced by the reference's own Python code over
    oracle-backed stand-ins for its two native wheels (tests/golden/make_reference_flows.py).  Where the real wheels and
    the reference's source are available, the same generator runs on them — and must write the same file: one command
    that pins header layout, bit order, hash seeds, non-ACGT handling, result order and the rbloom constants at once."""
    import json
    import os
    import subprocess
    import sys
    from pathlib import Path
    pytest.importorskip("cobs_index")
    pytest.importorskip("rbloom")
    pytest.importorskip("Bio")
    src = Path(os.environ.get("XS_REFERENCE_SRC", "/root/reference/src"))
    if not (src / "xspect" / "models").is_dir():
        pytest.skip("the reference's source tree is not available (XS_REFERENCE_SRC)")
    root = Path(__file__).resolve().parent.parent
    out = tmp_path / "flows_real.json"
    subprocess.run([sys.executable, str(root / "tests" / "golden" / "make_reference_flows.py"), "--real", "--out", str(out),
                    "--reference-src", str(src)], check=True, cwd=root)
    got, gold = json.loads(out.read_text()), json.loads((root / "tests" / "golden" / "reference_flows.json").read_text())
    for section in gold:
        assert got[section] == gold[section], section
