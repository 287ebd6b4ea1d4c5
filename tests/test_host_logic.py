"""Host-side logic that needs no GPU: reader, result API, slugs/paths, MLST splitter — mirrored on the
reference's own unit tests (tests/test_model_result.py, test_model_management.py,
test_probabilistic_filter_mlst_model.py:127-177, test_probabilistic_filter_model.py:139-146)."""
from pathlib import Path

import numpy as np
import pytest

from xspect2_b200 import model_management as mm
from xspect2_b200 import seqio
from xspect2_b200.file_io import get_record_iterator, prepare_input_output_paths, filter_sequences
from xspect2_b200.models.mlst_result import MlstResult
from xspect2_b200.models.probabilistic_filter_mlst_model import ProbabilisticFilterMlstSchemeModel
from xspect2_b200.models.probabilistic_filter_model import ProbabilisticFilterModel
from xspect2_b200.models.result import ModelResult
from xspect2_b200.seqio import Seq, SeqRecord, SequenceBatch


# ---------------------------------------------------------------- ModelResult (G10)
def test_get_scores_rounding_table():
    hits = {"subsequence1": {"label1": 10, "label2": 5}, "subsequence2": {"label1": 8, "label2": 3}}
    res = ModelResult("test_slug", hits, {"subsequence1": 100, "subsequence2": 50})
    assert res.get_scores() == {
        "subsequence1": {"label1": 0.1, "label2": 0.05},
        "subsequence2": {"label1": 0.16, "label2": 0.06},
        "total": {"label1": 0.12, "label2": 0.05},
    }
    assert res.get_total_hits() == {"label1": 18, "label2": 8}


def test_model_result_contract(tmp_path):
    with pytest.raises(ValueError):
        ModelResult("s", {"total": {"a": 1}}, {"total": 1})
    hits = {"r1": {"a": 9, "b": 1}, "r2": {"a": 2, "b": 8}, "misclassified": {"x": 1}}
    res = ModelResult("s", hits, {"r1": 10, "r2": 10}, sparse_sampling_step=2, prediction="a")
    assert res.misclassified == {"x": 1} and "misclassified" not in res.hits
    assert res.get_filter_mask("a", 0.7) == {"r1": True, "r2": False}
    assert res.get_filter_mask("b", -1) == {"r1": False, "r2": True}
    assert res.get_filtered_subsequence_labels("a") == ["r1"]
    with pytest.raises(ValueError):
        res.get_filter_mask("a", 1.5)
    d = res.to_dict()
    assert list(d) == ["model_slug", "sparse_sampling_step", "hits", "scores", "num_kmers", "misclassified", "input_source", "prediction"]
    res.save(tmp_path / "sub" / "out.json")
    assert (tmp_path / "sub" / "out.json").read_text().startswith("{\n    \"model_slug\"")
    m = MlstResult("Oxford", 1, {"rec": [{"Strain type": {}}, {"All results": {}}]}, "x.fna")
    assert list(m.to_dict()) == ["Scheme", "Steps", "Results", "Input_source"]


# ---------------------------------------------------------------- slugs and model paths
def test_slug_and_model_paths(tmp_path, monkeypatch):
    monkeypatch.setattr(mm, "get_xspect_model_path", lambda: tmp_path)
    assert "acinetobacter-genus.json" in str(mm.get_genus_model_path("Acinetobacter"))
    assert "salmonella-species.json" in str(mm.get_species_model_path("Salmonella"))
    assert mm.get_mlst_model_path("abaumannii", "MLST (Oxford)").name == "abaumannii-mlst-oxford-mlst.json"
    assert mm.slugify("Acinetobacter-Species") == "acinetobacter-species"
    assert mm.slugify("Test Filter-Species") == "test-filter-species"
    (tmp_path / "acinetobacter-species.json").write_text(
        '{"model_type": "Species", "model_display_name": "Acinetobacter", "model_class": "ProbabilisticFilterSVMModel", "display_names": {"470": "Acinetobacter baumannii"}}')
    assert mm.is_svm_model("acinetobacter-species")
    assert mm.get_models() == {"Species": ["Acinetobacter"]}
    assert mm.get_model_display_names("acinetobacter-species") == ["Acinetobacter baumannii"]
    with pytest.raises(ValueError):
        mm.get_model_metadata("nope")


# ---------------------------------------------------------------- model metadata (no index needed)
def _model(tmp_path, k=21):
    return ProbabilisticFilterModel(k, "Test Filter", "John Doe", "john.doe@example.com", "Species", Path(tmp_path))


def test_model_metadata_and_validation(tmp_path):
    m = _model(tmp_path)
    assert m.slug() == "test-filter-species"
    assert m.get_cobs_index_path().endswith("test-filter-species/index.cobs_classic")
    d = m.to_dict()
    assert d["model_class"] == "ProbabilisticFilterModel" and d["k"] == 21 and d["num_hashes"] == 7 and d["fpr"] == 0.01
    for bad in (dict(k=0), dict(model_display_name=""), dict(model_type=""), dict(base_path="str")):
        args = dict(k=21, model_display_name="x", author=None, author_email=None, model_type="Species", base_path=Path(tmp_path))
        args.update(bad)
        with pytest.raises(ValueError):
            ProbabilisticFilterModel(**args)
    m.save()
    assert (Path(tmp_path) / "test-filter-species.json").is_file()
    with pytest.raises(FileNotFoundError):
        ProbabilisticFilterModel.load(Path(tmp_path) / "test-filter-species.json")


def test_count_kmers_g4(tmp_path):
    m = _model(tmp_path)
    seq = Seq("AGAGATTACGTCTGGTTGCAAGAGATCATGACAGGGGGAATTGGTTGAAAATAAATATATCGCCAGCAGCACATGAACAA")
    assert m._count_kmers(seq) == 60                      # tests/test_probabilistic_filter_model.py:139-146
    assert m._count_kmers(SeqRecord(seq, "x")) == 60
    assert [m._count_kmers(seq, step=s) for s in (1, 2, 3, 4)] == [60, 30, 20, 15]
    assert m._count_kmers([seq, seq], step=7) == 18
    with pytest.raises(ValueError):
        m._count_kmers("not a seq")


# ---------------------------------------------------------------- MLST host pieces (G9)
def _mlst(tmp_path, k=4):
    return ProbabilisticFilterMlstSchemeModel(k=k, model_display_name="Test Filter", organism="Test Organism",
                                              base_path=Path(tmp_path), scheme_url="")


def test_sequence_splitter_reference_case(tmp_path):
    m = _mlst(tmp_path)
    seq = "AGCTATTTCGCTGATGTCGACTGATCAAAAAGCCGGCGCGCTTTCGTATAGGCTAGCTACGACATACGATCGATCACTGA"
    res = m.sequence_splitter(seq, 20)
    assert len(res) == 5 and len(res[-1]) == 12 and all(len(c) == 20 for c in res[:4])
    assert m.slug() == "test-organism-test-filter-mlst"


def test_sequence_splitter_matches_oracle_and_segments(tmp_path, oracle):
    rng = np.random.default_rng(5)
    for k, allele, n in [(21, 450, 10000), (21, 450, 10011), (31, 398, 12345), (4, 20, 80), (4, 20, 82), (21, 60, 1000005), (31, 100, 99999)]:
        m = _mlst(tmp_path, k)
        a = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)]
        s = a.tobytes().decode()
        chunks = m.sequence_splitter(s, allele)
        assert chunks == oracle.sequence_splitter(s, allele, k)
        bases, b, e = m._chunk_segments(a, allele)
        assert [bases[int(x):int(y)].tobytes().decode() for x, y in zip(b, e)] == chunks


def test_has_sufficient_score(tmp_path):
    m = _mlst(tmp_path)
    sizes = [405, 400, 484, 305, 457, 371, 513]
    good = {"Scores": {"Oxf_cpn60": 265, "Oxf_gdhB": 381, "Oxf_gltA": 241, "Oxf_gpi": 541, "Oxf_gyrB": 352, "Oxf_recA": 254, "Oxf_rpoD": 286}}
    bad = {"Scores": {"Oxf_cpn60": 100, "Oxf_gdhB": 20, "Oxf_gltA": 6, "Oxf_gpi": 55, "Oxf_gyrB": 21, "Oxf_recA": 0, "Oxf_rpoD": 7}}
    assert m.has_sufficient_score(good, sizes)
    assert not m.has_sufficient_score(bad, sizes)
    assert m.has_sufficient_score({"a": {"x": 10}, "b": {}, "c": {"y": 300}}, [100, 100, 500]) is True


# ---------------------------------------------------------------- reader
FASTA = ">seq1 first record\nACGTAC GT\nNNacgt\r\n\n>seq2\nTTTT\n>seq3 empty\n>seq1 duplicate id\nGG\n"
FASTQ = "@r1 desc\nACGTN\n+\nIIIII\n@r2\nGGCC\n+r2\nJJJJ\n"


def test_fasta_iterator_and_batch_agree(tmp_path):
    p = tmp_path / "x.fna"
    p.write_text(FASTA)
    recs = list(get_record_iterator(p))
    assert [(r.id, str(r.seq)) for r in recs] == [("seq1", "ACGTACGTNNacgt"), ("seq2", "TTTT"), ("seq3", ""), ("seq1", "GG")]
    assert recs[0].description == "seq1 first record"
    b = SequenceBatch.from_file(p)
    assert b.ids == [r.id for r in recs]
    assert [b.sequence(i) for i in range(len(b))] == [str(r.seq) for r in recs]
    b2 = SequenceBatch.from_records(recs)
    assert np.array_equal(b2.lengths, b.lengths)


def test_fastq_iterator_and_batch_agree(tmp_path):
    p = tmp_path / "x.fq"
    p.write_text(FASTQ)
    recs = list(get_record_iterator(p))
    assert [(r.id, str(r.seq)) for r in recs] == [("r1", "ACGTN"), ("r2", "GGCC")]
    b = SequenceBatch.from_file(p)
    assert b.ids == ["r1", "r2"] and [b.sequence(i) for i in range(2)] == ["ACGTN", "GGCC"]
    # wrapped FASTQ falls back to the general parser
    p2 = tmp_path / "w.fastq"
    p2.write_text("@w1\nACGT\nAC\n+\nIIII\nII\n@w2\nTT\n+\nII\n")
    b = SequenceBatch.from_file(p2)
    assert [b.sequence(i) for i in range(2)] == ["ACGTAC", "TT"]
    bad = tmp_path / "bad.fq"
    bad.write_text("@w1\nACGT\n+\nIII\n")
    with pytest.raises(ValueError):
        SequenceBatch.from_file(bad)


def test_large_random_fasta_roundtrip(tmp_path):
    rng = np.random.default_rng(3)
    p = tmp_path / "big.fasta"
    want = []
    with open(p, "w") as f:
        for i in range(200):
            s = "".join(rng.choice(list("ACGTNacgt"), size=int(rng.integers(0, 500))))
            want.append((f"c{i}", s))
            f.write(f">c{i} len={len(s)}\n")
            for j in range(0, len(s), 70):
                f.write(s[j:j + 70] + ("\r\n" if i % 2 else "\n"))
    b = SequenceBatch.from_file(p)
    assert [(b.ids[i], b.sequence(i)) for i in range(len(b))] == want


def test_record_iterator_errors_and_path_fanout(tmp_path):
    with pytest.raises(ValueError):
        get_record_iterator("str")
    with pytest.raises(ValueError):
        get_record_iterator(tmp_path / "missing.fna")
    (tmp_path / "a.txt").write_text("x")
    with pytest.raises(ValueError):
        get_record_iterator(tmp_path / "a.txt")
    with pytest.raises(ValueError):
        get_record_iterator(tmp_path)
    (tmp_path / "d").mkdir()
    for n in ("a.fasta", "b.fq", "c.txt"):
        (tmp_path / "d" / n).write_text(">x\nACGT\n" if n.endswith("a") else "@x\nAC\n+\nII\n")
    paths, out = prepare_input_output_paths(tmp_path / "d")
    assert sorted(p.name for p in paths) == ["a.fasta", "b.fq"]
    assert out(0, tmp_path / "res.json").name == "res_1.json"
    paths, out = prepare_input_output_paths(tmp_path / "d" / "a.fasta")
    assert out(0, tmp_path / "res.json").name == "res.json"
    with pytest.raises(ValueError):
        prepare_input_output_paths(tmp_path / "nope")


def test_filter_sequences_writes_fasta(tmp_path):
    p = tmp_path / "x.fna"
    p.write_text(FASTA)
    out = tmp_path / "f.fasta"
    filter_sequences(p, out, ["seq2", "seq1"])
    got = list(get_record_iterator(out))
    assert [(r.id, str(r.seq)) for r in got] == [("seq1", "ACGTACGTNNacgt"), ("seq2", "TTTT"), ("seq1", "GG")]
    filter_sequences(p, tmp_path / "none.fasta", [])
    assert not (tmp_path / "none.fasta").exists()


def test_seq_semantics_used_by_the_bloom_path(oracle):
    """min(kmer, str(kmer.reverse_complement())) as in probabilistic_single_filter_model.py:175-180."""
    rng = np.random.default_rng(9)
    alphabet = list("ACGTNacgtRYKMSWBDHVUu-")
    for _ in range(300):
        s = "".join(rng.choice(alphabet, size=21))
        kmer = Seq(s)
        assert str(min(kmer, str(kmer.reverse_complement()))).encode() == oracle.bloom_term(s.encode())
    assert seqio.is_seq(Seq("A")) and not seqio.is_seq("A") and seqio.is_record(SeqRecord("ACGT", "i"))


def test_native_reader_many_records_matches_iterators(tmp_path):
    """> 2 checkpoint segments (parallel fill pass), ragged reads, '@' and '+' opening quality lines, CRLF."""
    rng = np.random.default_rng(17)
    n = 70_001
    lens = rng.integers(0, 90, size=n)
    lens[::501] = 0
    allb = np.frombuffer(b"ACGTNacgt", np.uint8)[rng.integers(0, 9, size=int(lens.sum()))]
    ends = np.cumsum(lens)
    starts = ends - lens
    fq = tmp_path / "many.fq"
    with open(fq, "wb") as f:
        for i in range(n):
            s = allb[starts[i]:ends[i]].tobytes()
            q = (b"@" if i % 3 == 0 else b"+") * len(s)
            nl = b"\r\n" if i % 5 == 0 else b"\n"
            f.write(b"@r%d\tpair=%d" % (i, i % 2) + nl + s + nl + b"+" + nl + q + nl)
    b = SequenceBatch.from_file(fq)
    assert len(b) == n and np.array_equal(b.bases, allb)
    assert np.array_equal(b.begin, starts.astype(np.uint64)) and np.array_equal(b.end, ends.astype(np.uint64))
    assert b.ids == [f"r{i}" for i in range(n)]
    it = list(get_record_iterator(fq))
    assert [r.id for r in it[:2000]] == b.ids[:2000]
    assert [str(r.seq) for r in it[40000:41000]] == [b.sequence(i) for i in range(40000, 41000)]
    fa = tmp_path / "many.fna"
    with open(fa, "wb") as f:
        f.write(b"; comment before the first record\n")
        for i in range(n):
            s = allb[starts[i]:ends[i]].tobytes()
            f.write(b">c%d some description\n" % i)
            for j in range(0, len(s), 40):
                f.write(s[j:j + 40][:20] + b" " + s[j:j + 40][20:] + (b"\r\n" if i % 2 else b"\n"))
    b2 = SequenceBatch.from_file(fa)
    assert np.array_equal(b2.bases, allb) and np.array_equal(b2.end, ends.astype(np.uint64))
    assert b2.ids[0] == "c0" and b2.ids[-1] == f"c{n - 1}"
    empty = tmp_path / "empty.fasta"
    empty.write_text("")
    assert len(SequenceBatch.from_file(empty)) == 0
    with pytest.raises(ValueError):
        SequenceBatch.from_file(tmp_path / "x.txt")
    bad = tmp_path / "bad2.fastq"
    bad.write_text("@a\nAC GT\n+\nIIIII\n")
    with pytest.raises(ValueError, match="Whitespace"):
        SequenceBatch.from_file(bad)
    bad.write_text("ACGT\n")
    with pytest.raises(ValueError, match="should start with '@'"):
        SequenceBatch.from_file(bad)


def test_min_hits_table_is_exact():
    """The integer threshold table reproduces `round(hits / num_kmers, 2) >= threshold` (result.py:59,92-123)."""
    from xspect2_b200.pipeline import min_hits_table
    for thr in (0.0, 0.005, 0.01, 0.5, 0.7, 0.85, 0.995, 1.0):
        t = min_hits_table(400, thr)
        for n in list(range(1, 140)) + [150, 199, 200, 399, 400]:
            for h in range(n + 1):
                assert (h >= t[n]) == (round(h / n, 2) >= thr), (thr, n, h)


def test_columnar_result_json_is_byte_identical(tmp_path):
    """ColumnarModelResult.save (native writer) == json.dumps(ModelResult.to_dict(), indent=4), incl. duplicate ids,
    excluded documents, rewritten keys, ids that need escaping, score rounding for many (hits, num_kmers) pairs."""
    import json
    from xspect2_b200.engine import CobsIndex
    from xspect2_b200.models.result import ColumnarModelResult
    rng = np.random.default_rng(23)
    for case in range(6):
        n, d = (1, 3) if case == 0 else (int(rng.integers(2, 400)), int(rng.integers(1, 130)))
        nk = rng.integers(1, 5000 if case % 2 else 131, size=n).astype(np.int64)
        counts = np.minimum(rng.integers(0, 9000, size=(n, d)) * (rng.random((n, d)) < 0.4), nk[:, None]).astype(np.uint32)
        ids = [f"read{i}/1" for i in range(n)]
        if n > 5:
            ids[3] = ids[1]                       # duplicate id: first position, last values
            ids[4] = 'we"ird\\\\id\tx'
        names = [str(470 + 3 * j) for j in range(d)]
        keys = [f"{nm} - species {nm}" for nm in names] if case % 3 == 0 else list(names)
        include = np.ones(d, bool)
        if d > 2 and case % 2:
            include[[0, d - 1]] = False
        pred = None if case < 2 else "471"
        col = ColumnarModelResult("test-slug", ids, names, keys, include, counts, nk, sparse_sampling_step=1 + case % 3, prediction=pred,
                                  input_source=None if case == 1 else "in.fq")
        out = tmp_path / f"col{case}.json"
        col.save(out)
        # the reference's dict-based construction of the same result
        hits, num_kmers = {}, {}
        for i, rid in enumerate(ids):
            order = CobsIndex.result_order(counts[i])
            hits[rid] = {keys[j]: int(counts[i, j]) for j in order.tolist() if include[j]}
            num_kmers[rid] = int(nk[i])
        ref = ModelResult("test-slug", hits, num_kmers, 1 + case % 3, pred, None if case == 1 else "in.fq")
        assert out.read_text() == json.dumps(ref.to_dict(), indent=4)
        assert col.get_total_hits() == ref.get_total_hits() and list(col.get_total_hits()) == list(ref.get_total_hits())
        assert col.get_total_scores() == ref.get_scores()["total"]
        for lab in [key for key, keep in zip(keys, include) if keep][:3]:
            for thr in (0.0, 0.01, 0.3, 0.7, 1.0, -1):
                assert col.get_filter_mask(lab, thr) == ref.get_filter_mask(lab, thr), (case, lab, thr)
        assert col._hits is None                  # all of the above without building the nested dictionaries
        with pytest.raises(ValueError):
            col.get_filter_mask(keys[0], 1.5)
        assert col.hits == ref.hits and col.num_kmers == ref.num_kmers and col.to_dict() == ref.to_dict()
        assert col.get_filtered_subsequence_labels(keys[1 if d > 2 else 0] if include[1 if d > 2 else 0] else keys[1], 0.3) == \
            ref.get_filtered_subsequence_labels(keys[1 if d > 2 else 0] if include[1 if d > 2 else 0] else keys[1], 0.3)
    with pytest.raises(ValueError):
        ColumnarModelResult("s", ["total"], ["a"], ["a"], np.ones(1, bool), np.zeros((1, 1), np.uint32), np.ones(1, np.int64))


def test_native_reader_fuzz_against_python_iterators(tmp_path):
    """Randomised FASTA / FASTQ texts (blank lines, CRLF, wrapped records, odd characters, missing final newline):
    the native reader and the pure-Python iterators agree on ids and sequences, and fail on the same inputs."""
    from hypothesis import given, settings, strategies as st
    from xspect2_b200.seqio import FastaIterator, FastqPhredIterator

    seq_chars = "ACGTNacgtnRYKM*-"
    word = st.text(alphabet="abcXYZ019_./|:=", min_size=0, max_size=8)
    title = st.builds(lambda w, rest: (w + (" " + rest if rest else "")), word, st.text(alphabet="abc =;,", max_size=10))
    seq = st.text(alphabet=seq_chars, min_size=0, max_size=70)
    nl = st.sampled_from(["\n", "\r\n"])

    fasta_rec = st.builds(lambda t, s, w, e: ">" + t + e + "".join(s[i:i + w] + e for i in range(0, len(s), w)) + (e if len(s) % 7 == 0 else ""),
                          title, seq, st.integers(5, 40), nl)
    fasta_text = st.builds(lambda recs, cut: ("".join(recs))[: None if not cut else -1] if recs else "", st.lists(fasta_rec, max_size=8), st.booleans())

    @settings(max_examples=150, deadline=None)
    @given(fasta_text)
    def check_fasta(text):
        p = tmp_path / "h.fasta"
        p.write_bytes(text.encode())
        exp = [(r.id, str(r.seq)) for r in FastaIterator(p)]
        b = SequenceBatch.from_file(p)
        assert [(b.ids[i], b.sequence(i)) for i in range(len(b))] == exp

    fastq_rec = st.builds(lambda t, s, w, e, qc: "@" + t + e + ("".join(s[i:i + w] + e for i in range(0, len(s), w)) if s else e) + "+" + e +
                          ("".join((qc * len(s))[i:i + w] + e for i in range(0, len(s), w)) if s else e),
                          title, st.text(alphabet="ACGTNacgt", min_size=0, max_size=60), st.integers(4, 80), nl, st.sampled_from("I@+#5"))
    fastq_text = st.builds(lambda recs, junk: "".join(recs) + junk, st.lists(fastq_rec, max_size=8), st.sampled_from(["", "\n", "@x\nAC\n+\nI\n", "ACGT\n"]))

    @settings(max_examples=150, deadline=None)
    @given(fastq_text)
    def check_fastq(text):
        p = tmp_path / "h.fastq"
        p.write_bytes(text.encode())
        try:
            exp = [(r.id, str(r.seq)) for r in FastqPhredIterator(p)]
        except ValueError:
            with pytest.raises(ValueError):
                SequenceBatch.from_file(p)
            return
        b = SequenceBatch.from_file(p)
        assert [(b.ids[i], b.sequence(i)) for i in range(len(b))] == exp

    check_fasta()
    check_fastq()


def test_native_filtered_fasta_equals_python_writer(tmp_path):
    """filter_sequences through the native writer == SeqIO.write-style output of the kept parsed records."""
    import io
    rng = np.random.default_rng(41)
    recs = []
    for i in range(300):
        s = "".join(rng.choice(list("ACGTNacgt"), size=int(rng.integers(0, 250))))
        recs.append((f"r{i} sample={i % 7} len={len(s)}" if i % 4 else f"r{i}", s))
    fq = tmp_path / "in.fastq"
    fq.write_text("".join(f"@{t}\n{s}\n+\n{'I' * len(s)}\n" for t, s in recs))
    fa = tmp_path / "in.fasta"
    fa.write_text("".join(f">{t}\r\n" + "".join(s[j:j + 70] + "\r\n" for j in range(0, len(s), 70)) for t, s in recs))
    wanted = [f"r{i}" for i in range(0, 300, 3)]
    for src in (fq, fa):
        out = tmp_path / f"kept_{src.suffix[1:]}.fasta"
        filter_sequences(src, out, wanted)
        buf = io.StringIO()
        seqio.write_fasta((r for r in get_record_iterator(src) if r.id in set(wanted)), buf)
        assert out.read_text() == buf.getvalue()
        assert [r.id for r in get_record_iterator(out)] == wanted


def test_columnar_result_json_multi_threaded_writer(tmp_path, monkeypatch):
    """More records than one formatting block: the writer formats blocks on several host threads and must still
    produce the reference's bytes (and the same bytes as with one thread)."""
    import json
    import numpy as np
    from xspect2_b200.engine import CobsIndex
    from xspect2_b200.models.result import ColumnarModelResult, ModelResult
    rng = np.random.default_rng(77)
    n, d = 9001, 7
    nk = rng.integers(1, 300, size=n).astype(np.int64)
    counts = np.minimum(rng.integers(0, 400, size=(n, d)) * (rng.random((n, d)) < 0.5), nk[:, None]).astype(np.uint32)
    ids = [f"r{i}" for i in range(n)]
    names = [str(100 + j) for j in range(d)]
    include = np.ones(d, bool)
    include[2] = False
    col = ColumnarModelResult("slug", ids, names, list(names), include, counts, nk, sparse_sampling_step=1, prediction="101", input_source="x.fq")
    outs = []
    for threads in ("1", "5", "32"):
        monkeypatch.setenv("XS_RESULT_THREADS", threads)
        out = tmp_path / f"t{threads}.json"
        col.save(out)
        outs.append(out.read_bytes())
    assert outs[0] == outs[1] == outs[2]
    hits = {rid: {names[j]: int(counts[i, j]) for j in CobsIndex.result_order(counts[i]).tolist() if include[j]} for i, rid in enumerate(ids)}
    ref = ModelResult("slug", hits, {rid: int(v) for rid, v in zip(ids, nk)}, 1, "101", "x.fq")
    assert outs[0].decode() == json.dumps(ref.to_dict(), indent=4)


def test_pipeline_block_plan():
    """pipeline._plan_blocks: record ranges that cover the batch in order, bounded in records and bytes, with the
    byte span and the length extremes of each range; records laid out of order fall back to one block."""
    from xspect2_b200.pipeline import _plan_blocks, min_hits_table
    rng = np.random.default_rng(3)
    lens = rng.integers(22, 400, size=10_000).astype(np.uint64)
    end = np.cumsum(lens, dtype=np.uint64)
    begin = end - lens
    blocks = _plan_blocks(begin, end, block_records=1500, block_bytes=100_000)
    assert blocks[0][0] == 0 and blocks[-1][1] == lens.size
    assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
    for i0, i1, lo, hi, mn, mx in blocks:
        assert 0 < i1 - i0 <= 1500 and lo == begin[i0] and hi == end[i1 - 1]
        assert hi - lo <= 4 * 100_000 + (1 << 20)
        assert mn == lens[i0:i1].min() and mx == lens[i0:i1].max()
    assert len(blocks) > 10
    perm = rng.permutation(lens.size)                     # same records, arbitrary order: spans cover the whole buffer
    one = _plan_blocks(begin[perm], end[perm], block_records=1500, block_bytes=1000)
    assert one == [(0, lens.size, 0, int(end[-1]), int(lens.min()), int(lens.max()))]
    assert _plan_blocks(begin[:0], end[:0], 10, 10) == []
    # the exact threshold table: smallest h with round(h / n, 2) >= t
    t = min_hits_table(300, 0.7)
    for n in (1, 2, 3, 7, 130, 299, 300):
        h = int(t[n])
        assert round(h / n, 2) >= 0.7 and (h == 0 or round((h - 1) / n, 2) < 0.7)


def test_svm_is_refit_only_when_its_inputs_change(tmp_path):
    """ProbabilisticFilterSVMModel._get_svm keeps the fitted classifier while scores.csv, kernel, C and the excluded
    ids stay the same (the reference refits per predict call: same deterministic fit, same predictions)."""
    import csv
    import time

    from xspect2_b200.models.probabilistic_filter_svm_model import ProbabilisticFilterSVMModel as M
    m = M.__new__(M)
    m.base_path, m.kernel, m.c = tmp_path, "rbf", 1.5
    m.display_names = {str(1000 + d): f"s{d}" for d in range(5)}
    m.slug = lambda: "g-species"
    (tmp_path / "g-species").mkdir()

    def write(seed):
        rng = np.random.default_rng(seed)
        with open(tmp_path / "g-species" / "scores.csv", "w", encoding="utf-8") as f:
            w = csv.writer(f)
            w.writerow(["file"] + sorted(m.display_names) + ["label"])
            for d in range(5):
                for r in range(4):
                    x = rng.random(5) * 0.3
                    x[d] = 0.95
                    w.writerow([f"a{d}{r}"] + [f"{v:.2f}" for v in x] + [str(1000 + d)])

    write(1)
    a = m._get_svm(None)
    assert m._get_svm(None) is a
    c = m._get_svm(["1001"])                       # other excluded ids: a different classifier (4 classes, 4 features)
    assert c is not a and len(c.classes_) == 4 and c.n_features_in_ == 4
    d = m._get_svm(None)
    assert d is not a and list(d.classes_) == list(a.classes_)
    assert str(d.predict([[0.95, 0.1, 0.1, 0.1, 0.1]])[0]) == "1000"
    time.sleep(0.01)
    write(2)                                       # retrained model directory: refit
    assert m._get_svm(None) is not d
