"""Model / workflow / CLI level parity on the GPU: the reference's API driven on synthetic model directories,
compared with the reference's per-record loops restated over the CPU oracle (oracle.reference_predict*).
Mirrors tests/test_probabilistic_filter_model.py, test_probabilistic_single_filter_model.py,
test_probabilistic_filter_svm_model.py, test_probabilistic_filter_mlst_model.py and test_cli.py of the reference."""
import importlib
import json
from pathlib import Path

import numpy as np
import pytest

from tests import model_fixtures as mf
from tests import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world(tmp_path_factory, oracle, gpu):
    """One synthetic data root (models dir) shared by the module."""
    root = tmp_path_factory.mktemp("xspect-data")
    models = root / "models"
    models.mkdir()
    rng = np.random.default_rng(2026)
    sp_json, genomes, svm_genomes = mf.species_model(oracle, models, rng)
    ge_json = mf.genus_model(oracle, models, list(genomes.values()))
    ml_json, alleles = mf.mlst_model(oracle, models, rng)
    return dict(root=root, models=models, sp_json=sp_json, ge_json=ge_json, ml_json=ml_json, genomes=genomes,
                svm_genomes=svm_genomes, alleles=alleles, rng=rng)


def _records(world, n_reads=60):
    rng = np.random.default_rng(77)
    gl = list(world["genomes"].values())
    bases, b, e = synth.sample_reads(rng, gl, n_reads, (40, 400), sub=0.01, n_rate=0.002)
    recs = [(f"read{i}", bases[int(x):int(y)].tobytes().decode()) for i, (x, y) in enumerate(zip(b, e))]
    recs.append(("contig_long", np.concatenate([gl[0], gl[2][:3000]]).tobytes().decode()))
    recs.append(("read3", gl[1][:500].tobytes().decode()))        # duplicate id: overwrites the earlier read3
    return recs


# ------------------------------------------------------------------------------ species model
def test_species_predict_matches_reference_loop(world, oracle, tmp_path):
    from xspect2_b200.models.probabilistic_filter_model import ProbabilisticFilterModel
    from xspect2_b200.seqio import Seq, SeqRecord
    model = ProbabilisticFilterModel.load(world["sp_json"])
    orc = oracle.CobsOracle(model.get_cobs_index_path())
    recs = _records(world)
    fasta = tmp_path / "in.fna"
    mf.write_fasta(fasta, recs)
    for step in (1, 3):
        res = model.predict(fasta, step=step)
        hits, nk = oracle.reference_predict(orc, recs, model.k, None, step)
        assert res.hits == hits
        assert [list(v.items()) for v in res.hits.values()] == [list(v.items()) for v in hits.values()]  # dict order too
        assert res.num_kmers == nk and res.sparse_sampling_step == step
    # other input kinds: record, list of records, iterator, fastq path
    rl = [SeqRecord(Seq(s), rid) for rid, s in recs]
    h1, _ = oracle.reference_predict(orc, recs, model.k)
    assert model.predict(rl).hits == h1
    assert model.predict(rl[5]).hits == {"read5": h1["read5"]}
    fq = tmp_path / "in.fastq"
    mf.write_fastq(fq, recs)
    assert model.predict(fq).hits == h1
    from xspect2_b200.file_io import get_record_iterator
    assert model.predict(get_record_iterator(fasta)).hits == h1
    # calculate_hits on one Seq, with exclusion
    ex = [orc.names[1], orc.names[3]]
    got = model.calculate_hits(Seq(recs[0][1]), exclude_ids=ex, step=2)
    exp, _ = oracle.reference_predict(orc, recs[:1], model.k, ex, 2)
    assert list(got.items()) == list(exp[recs[0][0]].items())
    assert model.predict(fasta, exclude_ids=ex).hits == oracle.reference_predict(orc, recs, model.k, ex)[0]


def test_species_predict_display_names_and_errors(world, oracle, tmp_path):
    from xspect2_b200.models.probabilistic_filter_model import ProbabilisticFilterModel
    from xspect2_b200.seqio import Seq, SeqRecord
    model = ProbabilisticFilterModel.load(world["sp_json"])
    g = next(iter(world["genomes"].values()))
    rec = SeqRecord(Seq(g[:300].tobytes().decode()), "r")
    res = model.predict(rec, display_name=True)
    keys = list(res.hits["r"])
    assert all(" - species" in key for key in keys)          # "<id> - <name minus genus>" (tests/test_cli.py:88-128)
    assert {key.split(" -")[0] for key in keys} == set(model.display_names)
    with pytest.raises(ValueError, match="longer than k"):
        model.predict([rec, SeqRecord(Seq("ACGT" * 5 + "A"), "short")])       # len == k aborts the whole call
    with pytest.raises(ValueError):
        model.predict("not a record")
    with pytest.raises(ValueError):
        model.calculate_hits("ACGT" * 30)
    with pytest.raises(ValueError):
        model.predict(tmp_path / "missing.fna")
    with pytest.raises(NotImplementedError):                                  # never a silently unfiltered result
        model.predict(rec, validation=True)
    # 'misclassified' is a reserved record id of ModelResult (result.py:29)
    res_m = model.predict([SeqRecord(rec.seq, "misclassified"), rec])
    assert "misclassified" not in res_m.hits and res_m.misclassified == res_m.hits[rec.id]
    # whole training genome scores 1.0 on its own species (G3 shape)
    tid = next(iter(world["genomes"]))
    total = model.predict(SeqRecord(Seq(g.tobytes().decode()), "g")).get_scores()["total"]
    assert total[tid] == 1.0 and all(v < 1.0 for key, v in total.items() if key != tid)


def test_species_predict_arrays(world, oracle, tmp_path):
    from xspect2_b200.models.probabilistic_filter_model import ProbabilisticFilterModel
    model = ProbabilisticFilterModel.load(world["sp_json"])
    recs = _records(world)
    fasta = tmp_path / "in.fna"
    mf.write_fasta(fasta, recs)
    arr = model.predict_arrays(fasta, step=1)
    res = model.predict(fasta)
    # columnar totals count duplicate ids twice; compare per record instead
    for i, rid in enumerate(arr.ids):
        if rid != "read3":
            assert dict(zip(arr.names, arr.counts[i].tolist())) == res.hits[rid]
    best, tie = arr.argmax()
    assert best.shape == tie.shape == (len(recs),)


def test_species_predict_summary_matches_predict(world, oracle, tmp_path):
    from xspect2_b200.models.probabilistic_filter_model import ProbabilisticFilterModel
    model = ProbabilisticFilterModel.load(world["sp_json"])
    recs = [r for r in _records(world, 300) if r[0] != "read3"]          # unique ids: totals comparable
    fq = tmp_path / "in.fq"
    mf.write_fastq(fq, recs)
    for step in (1, 2):
        summ = model.predict_summary(fq, step=step)
        res = model.predict(fq, step=step)
        assert summ["batch"].ids == [r[0] for r in recs] and summ["labels"] == model.index.index.names
        assert summ["total_hits"] == {lab: res.get_total_hits()[lab] for lab in summ["labels"]}
        assert summ["total_scores"] == {lab: res.get_scores()["total"][lab] for lab in summ["labels"]}
        for i, (rid, _) in enumerate(recs):
            h = res.hits[rid]
            mx = max(h.values())
            winners = [lab for lab in summ["labels"] if h[lab] == mx]
            assert summ["labels"][int(summ["best"][i])] == winners[0]
            assert int(summ["best_hits"][i]) == mx and bool(summ["ambiguous"][i]) == (len(winners) > 1)
            assert int(summ["num_kmers"][i]) == res.num_kmers[rid]


def test_streamed_file_calls_equal_two_pass_reader(world, oracle, tmp_path):
    """xs_cobs_classify_file (blocks of the file parsed on all host threads while earlier blocks are scored) returns
    what the two-pass reader + xs_cobs_classify return — FASTA with wrapped lines, 4-line FASTQ, many tiny blocks —
    and hands wrapped FASTQ and short records over to the paths that own those cases."""
    from xspect2_b200.models.probabilistic_filter_model import ProbabilisticFilterModel
    from xspect2_b200.seqio import SequenceBatch
    model = ProbabilisticFilterModel.load(world["sp_json"])
    ix = model.index.index
    recs = _records(world, 900)
    fq, fa = tmp_path / "in.fastq", tmp_path / "in.fna"
    mf.write_fastq(fq, recs)
    mf.write_fasta(fa, recs, wrap=60)
    for path, fmt in ((fq, 2), (fa, 1)):
        batch = SequenceBatch.from_file(path)
        best, cnt, nb, totals = ix.classify(batch.bases, batch.begin, batch.end, 1)
        for block in (0, 3000, 40_000):
            r = ix.classify_file(path, fmt, 1, block_bytes=block)
            assert np.array_equal(r["best"], best) and np.array_equal(r["best_hits"], cnt) and np.array_equal(r["n_best"], nb), (path, block)
            assert np.array_equal(r["totals"], totals) and np.array_equal(r["seq_len"], batch.end - batch.begin)
            assert r["id_buf"].tobytes() == batch._id_buf.tobytes() and np.array_equal(r["id_end"], batch._id_end)
            assert r["n_short"] == 0 and r["n_bases"] == batch.bases.size
    summ = model.predict_summary(fq)
    assert "streamed" in summ and summ["batch"].ids == [r[0] for r in recs]
    # shapes of real files: CRLF line ends, no newline at the end, blank lines, one record, nothing at all
    def same(path, fmt):
        batch = SequenceBatch.from_file(path)
        exp = ix.classify(batch.bases, batch.begin, batch.end, 1) if len(batch) else None
        for block in (0, 777):
            r = ix.classify_file(path, fmt, 1, block_bytes=block)
            assert r["best"].size == len(batch) and np.array_equal(r["seq_len"], batch.end - batch.begin)
            assert r["id_buf"].tobytes() == batch._id_buf.tobytes()
            if exp is not None:
                assert np.array_equal(r["best"], exp[0]) and np.array_equal(r["best_hits"], exp[1]) and np.array_equal(r["n_best"], exp[2])
                assert np.array_equal(r["totals"], exp[3])
    text = lambda x: x if isinstance(x, str) else x.tobytes().decode()
    crlf = tmp_path / "crlf.fastq"
    crlf.write_bytes("".join(f"@{rid} extra words\r\n{text(sq)}\r\n+\r\n{'I' * len(text(sq))}\r\n" for rid, sq in recs[:40]).encode())
    same(crlf, 2)
    noeol = tmp_path / "noeol.fastq"
    noeol.write_bytes(("\n\n" + "".join(f"@{rid}\n{text(sq)}\n+{rid}\n{'#' * len(text(sq))}\n\n" for rid, sq in recs[:33])).rstrip("\n").encode())
    same(noeol, 2)
    fa2 = tmp_path / "odd.fasta"
    fa2.write_bytes(("; comment line\n" + "".join(f">{rid} desc\r\n{text(sq)[:70]}\r\n\r\n{text(sq)[70:].lower()}\n" for rid, sq in recs[:25]) + ">last\nACGTNNNNACGTACGTACGTACGTACGTAAAC").encode())
    same(fa2, 1)
    one = tmp_path / "one.fna"
    mf.write_fasta(one, recs[:1])
    same(one, 1)
    empty = tmp_path / "empty.fastq"
    empty.write_bytes(b"")
    same(empty, 2)
    # wrapped FASTQ: the streaming reader declines, predict_summary falls back to the two-pass reader
    wrapped = tmp_path / "wrapped.fastq"
    with open(wrapped, "w") as f:
        for rid, sq in recs[:50]:
            sq = sq if isinstance(sq, str) else sq.tobytes().decode()
            h = len(sq) // 2
            f.write(f"@{rid}\n{sq[:h]}\n{sq[h:]}\n+\n{'I' * h}\n{'I' * (len(sq) - h)}\n")
    with pytest.raises(ValueError):
        ix.classify_file(wrapped, 2, 1)
    s2 = model.predict_summary(wrapped)
    assert "streamed" not in s2 and np.array_equal(s2["best"], summ["best"][:50])
    # one record not longer than k aborts, as in the reference's loop
    short = tmp_path / "short.fastq"
    mf.write_fastq(short, recs[:10] + [("tiny", "ACGTACGTAC")])
    with pytest.raises(ValueError, match="longer than k"):
        model.predict_summary(short)
    # malformed FASTQ keeps Biopython's message through the fallback
    bad = tmp_path / "bad.fastq"
    bad.write_text("@r1\nACGTACGTACGTACGTACGTACGTACGT\n+\nIIII\n")
    with pytest.raises(ValueError, match="Lengths of sequence and quality"):
        model.predict_summary(bad)


# ------------------------------------------------------------------------------ SVM model
def test_svm_prediction_matches_reference_flow(world, oracle, tmp_path):
    from sklearn.svm import SVC
    from xspect2_b200.models.probabilistic_filter_svm_model import ProbabilisticFilterSVMModel
    from xspect2_b200.models.result import ModelResult
    model = ProbabilisticFilterSVMModel.load(world["sp_json"])
    assert model.to_dict()["model_class"] == "ProbabilisticFilterSVMModel" and model.kernel == "rbf"
    orc = oracle.CobsOracle(model.get_cobs_index_path())
    import csv
    rows = list(csv.reader(open(model.base_path / model.slug() / "scores.csv")))[1:]
    for acc, (tid, genome) in list(world["svm_genomes"].items())[::4]:
        contigs = [(f"{acc}_c{i}", genome[i * 1500:(i + 1) * 1500].tobytes().decode()) for i in range(4)]
        fasta = tmp_path / f"{acc}.fna"
        mf.write_fasta(fasta, contigs)
        for ex in (None, [sorted(model.display_names)[2]]):
            res = model.predict(fasta, exclude_ids=ex, step=1)
            hits, nk = oracle.reference_predict(orc, contigs, model.k, ex, 1)
            ref = ModelResult(model.slug(), hits, nk)
            x = [list(dict(sorted(ref.get_scores()["total"].items())).values())]
            assert model.svm_input(res) == x                                  # bit-identical SVM input
            keys = list(model.display_names)
            drop = {i for i, key in enumerate(keys) if ex and key in ex}
            xt = [[float(v) for i, v in enumerate(r[1:-1]) if i not in drop] for r in rows if not (ex and r[-1] in ex)]
            yt = [r[-1] for r in rows if not (ex and r[-1] in ex)]
            exp = str(SVC(kernel="rbf", C=1.0).fit(xt, yt).predict(x)[0])
            assert res.prediction == exp
            if ex is None:
                assert res.prediction == tid                                  # G7 shape: the right species
            assert res.hits == hits and "prediction" in res.to_dict()


# ------------------------------------------------------------------------------ genus model
def test_genus_predict_matches_reference_loop(world, oracle, tmp_path):
    from xspect2_b200.models.probabilistic_single_filter_model import ProbabilisticSingleFilterModel
    from xspect2_b200.seqio import Seq
    model = ProbabilisticSingleFilterModel.load(world["ge_json"])
    bf = oracle.BloomOracle(model.base_path / model.slug() / "filter.bloom", model.k)
    recs = _records(world)
    rng = np.random.default_rng(5)
    recs += [(f"junk{i}", synth.mutate(rng, synth.random_dna(rng, 200), lower=0.1, iupac=0.05).tobytes().decode()) for i in range(10)]
    fasta = tmp_path / "in.fa"
    mf.write_fasta(fasta, recs)
    for step in (1, 4):
        res = model.predict(fasta, step=step)
        hits, nk = oracle.reference_predict_bloom(bf, "Testgenus", recs, model.k, step)
        assert res.hits == hits and res.num_kmers == nk
    assert model.calculate_hits(Seq(recs[0][1]), step=4) == hits[recs[0][0]]
    kmers = list(model._generate_kmers(Seq(recs[1][1]), step=4))
    assert sum(1 for km in kmers if km in model.bf) == hits[recs[1][0]]["Testgenus"]
    sc = model.predict(fasta).get_scores()
    assert sc["contig_long"]["Testgenus"] == 1.0 and sc["junk0"]["Testgenus"] < 0.2
    with pytest.raises(ValueError):
        model.calculate_hits(Seq("ACGT"))


# ------------------------------------------------------------------------------ fused two-stage pipeline
def test_genus_then_species_matches_the_file_based_stages(world, oracle, tmp_path):
    """pipeline.genus_then_species == filter_genus (threshold on rounded scores) followed by species predict on
    the kept records (main.py:108-145), without writing the filtered FASTA."""
    from xspect2_b200.models.probabilistic_filter_svm_model import ProbabilisticFilterSVMModel
    from xspect2_b200.models.probabilistic_single_filter_model import ProbabilisticSingleFilterModel
    from xspect2_b200.pipeline import genus_then_species
    from xspect2_b200.seqio import Seq, SeqRecord
    genus = ProbabilisticSingleFilterModel.load(world["ge_json"])
    species = ProbabilisticFilterSVMModel.load(world["sp_json"])
    rng = np.random.default_rng(4)
    gl = list(world["genomes"].values())
    recs = []
    for i in range(400):
        L = int(rng.integers(40, 300))
        if i % 3 == 0:
            s = synth.random_dna(rng, L)
        else:
            g = gl[1]
            st = int(rng.integers(0, g.size - L))
            s = synth.mutate(rng, g[st:st + L], sub=float(rng.choice([0.0, 0.01, 0.03, 0.06])))
        recs.append((f"r{i}", s.tobytes().decode()))
    fq = tmp_path / "in.fastq"
    mf.write_fastq(fq, recs)
    for thr, step in ((0.7, 1), (0.3, 2), (1.0, 1)):
        out = genus_then_species(genus, species, fq, threshold=thr, step=step)
        # the same input through many small blocks (three in flight, slots reused): identical results
        small = genus_then_species(genus, species, fq, threshold=thr, step=step, block_records=37)
        for key in ("genus_hits", "kept", "kept_index", "best", "best_hits", "ambiguous", "num_kmers"):
            assert np.array_equal(out[key], small[key]), key
        assert (out["total_hits"], out["total_kmers"], out["prediction"]) == (small["total_hits"], small["total_kmers"], small["prediction"])
        gres = genus.predict(fq, step=step)
        kept_ids = gres.get_filtered_subsequence_labels("Testgenus", thr)
        assert [rid for rid, keep in zip(out["batch"].ids, out["kept"]) if keep] == kept_ids
        assert {rid: {"Testgenus": int(h)} for rid, h in zip(out["batch"].ids, out["genus_hits"])} == gres.hits
        if not kept_ids:
            assert out["prediction"] is None
            continue
        kept_recs = [SeqRecord(Seq(s), rid) for rid, s in recs if rid in set(kept_ids)]
        sres = species.predict(kept_recs, step=step)
        assert out["total_hits"] == {lab: sres.get_total_hits()[lab] for lab in out["labels"]}
        assert out["total_scores"] == {lab: sres.get_scores()["total"][lab] for lab in out["labels"]}
        assert out["prediction"] == sres.prediction
        for j, rid in enumerate(kept_ids[:100]):
            h = sres.hits[rid]
            assert int(out["best_hits"][j]) == max(h.values())
    assert len(kept_ids) == 0 or thr < 1.0 or all(gres.get_scores()[r]["Testgenus"] == 1.0 for r in kept_ids)


# ------------------------------------------------------------------------------ MLST model
class FakePubMLST:
    def __init__(self):
        self.calls = []

    def get_strain_type_name(self, highest_results, post_url):
        self.calls.append((highest_results, post_url))
        return {"ST": "2"}


def _mlst_reference(oracle, model, sequence: str, step=1, limit=False, limit_number=5):
    """calculate_hits' result structure from the oracle (probabilistic_filter_mlst_model.py:230-303)."""
    result_dict, highest = {}, {}
    for counter, locus in enumerate(model.loci):
        orc = oracle.CobsOracle(model.get_cobs_index_path(locus), load_complete=False)
        sc = oracle.mlst_locus_scores(orc, sequence, model.avg_locus_bp_size[counter], step)
        if len(sequence) >= 10000:
            if limit:
                sc = dict(list(sc.items())[:limit_number])
            if not sc:
                result_dict = "A Strain type could not be detected because of no kmer matches!"
                highest[locus] = {"N/A": 0}
                continue
        elif limit:
            sc = dict(sorted(sc.items(), key=lambda x: -x[1])[:limit_number])
        result_dict[locus] = sc
        first = next(iter(sc))
        highest[locus] = {first: sc[first]}
    return highest, result_dict


def test_mlst_calculate_hits_matches_reference_flow(world, oracle):
    from xspect2_b200.models.probabilistic_filter_mlst_model import ProbabilisticFilterMlstSchemeModel
    from xspect2_b200.seqio import Seq, SeqRecord
    model = ProbabilisticFilterMlstSchemeModel.load(world["ml_json"])
    assert model.slug() == "abaumannii-oxford-mlst" and len(model.indices) == 3
    fake = FakePubMLST()
    model.pubmlst_handler = fake
    rng = np.random.default_rng(31)
    picks = {locus: f"Allele_ID_{4 + 3 * i}" for i, locus in enumerate(model.loci)}
    parts = [synth.random_dna(rng, 7000)]
    for locus, name in picks.items():
        parts += [world["alleles"][locus][name], synth.random_dna(rng, 9000)]
    genome = np.concatenate(parts).tobytes().decode()
    for n in (len(genome), len(genome) - 7):                  # second length exercises the glued remainder
        g = genome[:n]
        for limit in (False, True):
            out = model.calculate_hits(Seq(g), limit=limit)
            highest, allres = _mlst_reference(oracle, model, g, limit=limit)
            got_high = dict(out[0]["Strain type"])
            assert got_high.pop("ST_Name") == {"ST": "2"}
            assert got_high == highest and list(got_high.items()) == list(highest.items())
            assert out[1]["All results"] == allres
            assert [list(v.items()) for v in out[1]["All results"].values()] == [list(v.items()) for v in allres.values()]
    assert {locus: next(iter(v)) for locus, v in highest.items()} == picks
    assert fake.calls[-1][0] == {locus: int(name.split("_")[-1]) for locus, name in picks.items()}
    # short branch (< 10000 bp): single allele, G9 shape (tests/test_probabilistic_filter_mlst_model.py:82-99)
    locus0 = next(iter(model.loci))
    allele = world["alleles"][locus0]["Allele_ID_4"].tobytes().decode()
    out = model.predict(SeqRecord(Seq(allele), "<unknown id>"))
    st = out.hits["test"][0]["Strain type"]
    assert st[locus0] == {"Allele_ID_4": len(allele) - model.k + 1}
    highest, allres = _mlst_reference(oracle, model, allele)
    assert out.hits["test"][1]["All results"] == allres
    # nothing matches: message + unreliable flag
    out = model.calculate_hits(Seq(synth.random_dna(rng, 12000).tobytes().decode()))
    assert out[1]["All results"] == "A Strain type could not be detected because of no kmer matches!"
    assert out[0]["Strain type"]["Attention:"].startswith("This strain type is not reliable")
    with pytest.raises(ValueError):
        model.calculate_hits(Seq("ACGT"))
    with pytest.raises(ValueError):
        model.predict("x")


def test_mlst_predict_file_batches_records(world, oracle, tmp_path):
    """predict(Path) scores all records of a file in one query per locus; per-record results equal the
    single-record path and the oracle's chunk loop."""
    from xspect2_b200.models.probabilistic_filter_mlst_model import ProbabilisticFilterMlstSchemeModel
    from xspect2_b200.seqio import Seq
    model = ProbabilisticFilterMlstSchemeModel.load(world["ml_json"])
    model.pubmlst_handler = FakePubMLST()
    rng = np.random.default_rng(77)
    loci = list(model.loci)
    al = world["alleles"]
    a1 = al[loci[1]]["Allele_ID_2"]

    def contig(ids, pads):
        parts = [synth.random_dna(rng, pads[0])]
        for locus, aid, pad in zip(loci, ids, pads[1:]):
            parts += [al[locus][f"Allele_ID_{aid}"], synth.random_dna(rng, pad)]
        return np.concatenate(parts).tobytes().decode()

    recs = [
        ("contig1", contig((9, 3, 17), (6000, 2500, 1800, 8000))),
        ("allele_only", a1.tobytes().decode()),
        ("contig2", contig((30, 2, 5), (11000, 700, 5000, 3003))),
        ("short_random", synth.random_dna(rng, 900).tobytes().decode()),
    ]
    fa = tmp_path / "asm.fna"
    mf.write_fasta(fa, recs)
    res = model.predict(fa)
    assert list(res.hits) == [r[0] for r in recs] and res.to_dict()["Scheme"] == "Oxford"
    for rid, s in recs:
        single = model.calculate_hits(Seq(s))
        assert res.hits[rid] == single
        highest, allres = _mlst_reference(oracle, model, s)
        got_high = {key: v for key, v in res.hits[rid][0]["Strain type"].items() if key not in ("ST_Name", "Attention:")}
        assert got_high == highest and list(got_high.items()) == list(highest.items())
        assert res.hits[rid][1]["All results"] == allres
    with pytest.raises(ValueError, match="longer than k"):
        bad = tmp_path / "bad.fna"
        mf.write_fasta(bad, recs[:1] + [("tiny", "ACGTACGT")])
        model.predict(bad)
    # bug-compatible: a long record that matches one locus well and another not at all makes the reference's
    # ST lookup fail on int("N/A") (probabilistic_filter_mlst_model.py:259-260,296-299)
    lonely = np.concatenate([synth.random_dna(rng, 9000), al[loci[0]]["Allele_ID_9"], synth.random_dna(rng, 9000)]).tobytes().decode()
    with pytest.raises(ValueError, match="invalid literal"):
        model.calculate_hits(Seq(lonely))


def test_mlst_query_many_hot_chunks_and_steps(world, oracle):
    """xs_mlst_query keeps 16 hot chunk rows per record and locus on the device path; a record with more (a locus
    repeated many times) takes the fallback that reads the flagged rows back.  Both must equal the oracle's chunk loop,
    order included, also with sparse sampling."""
    from xspect2_b200 import engine
    from xspect2_b200.models.probabilistic_filter_mlst_model import ProbabilisticFilterMlstSchemeModel
    model = ProbabilisticFilterMlstSchemeModel.load(world["ml_json"])
    rng = np.random.default_rng(91)
    loci = list(model.loci)
    al = world["alleles"]
    parts = []
    for rep in range(23):                                        # 23 copies of locus 0 alleles (> 16 hot chunks), 2 of locus 1
        parts += [synth.random_dna(rng, 1500 + 37 * rep), al[loci[0]][f"Allele_ID_{1 + rep % 5}"]]
    parts += [synth.random_dna(rng, 4000), al[loci[1]]["Allele_ID_7"], synth.random_dna(rng, 2600), al[loci[1]]["Allele_ID_8"]]
    many = np.concatenate(parts)
    short = al[loci[2]]["Allele_ID_3"]
    plain = np.concatenate([synth.random_dna(rng, 8000), al[loci[2]]["Allele_ID_11"], synth.random_dna(rng, 4000)])
    seqs = [many, short, plain, many[:-5]]
    sizes = np.array([x.size for x in seqs], np.uint64)
    end = np.cumsum(sizes, dtype=np.uint64)
    indices = [srch.index for srch in model.indices]
    for step in (1, 3):
        res = engine.mlst_query(indices, model.avg_locus_bp_size, np.concatenate(seqs), end - sizes, end, step)
        for li, locus in enumerate(loci):
            orc = oracle.CobsOracle(model.get_cobs_index_path(locus), load_complete=False)
            for ri, sq in enumerate(seqs):
                exp = oracle.mlst_locus_scores(orc, sq.tobytes().decode(), model.avg_locus_bp_size[li], step)
                docs, scores = res[li][ri]
                got = [(indices[li].names[d], v) for d, v in zip(docs.tolist(), scores.tolist())]
                assert got == list(exp.items()), (step, locus, ri)
    assert len(res[0][0][0]) >= 5                                # the repeated locus really produced several alleles


# ------------------------------------------------------------------------------ workflows + CLI
def test_workflows_and_cli(world, oracle, tmp_path, monkeypatch):
    monkeypatch.setenv("HOME", str(world["root"].parent))
    home_root = world["root"].parent / "xspect-data"
    if not home_root.exists():
        home_root.symlink_to(world["root"])
    from xspect2_b200 import classify, definitions, filter_sequences
    assert definitions.get_xspect_model_path() == home_root / "models"
    gl = world["genomes"]
    tid0 = next(iter(gl))
    rng = np.random.default_rng(8)
    recs = [(f"c{i}", gl[tid0][i * 1000:(i + 1) * 1000 + 200].tobytes().decode()) for i in range(5)]
    recs += [(f"junk{i}", synth.random_dna(rng, 600).tobytes().decode()) for i in range(3)]
    indir = tmp_path / "in"
    indir.mkdir()
    mf.write_fasta(indir / "sample.fna", recs)
    # classify species (SVM model picked through model_management.is_svm_model)
    classify.classify_species("Testgenus", indir / "sample.fna", tmp_path / "sp.json", exclude_ids=None)
    sp = json.loads((tmp_path / "sp.json").read_text())
    assert sp["input_source"] == "sample.fna" and sp["model_slug"] == "testgenus-species" and sp["prediction"] == tid0
    assert set(sp["hits"]) == {r[0] for r in recs} and set(sp["scores"]) == set(sp["hits"]) | {"total"}
    # directory input: numbered outputs
    classify.classify_genus("Testgenus", indir, tmp_path / "ge.json", step=2)
    ge = json.loads((tmp_path / "ge_1.json").read_text())
    assert ge["sparse_sampling_step"] == 2 and ge["scores"]["c0"]["Testgenus"] == 1.0
    # genus filter keeps the genome-derived contigs only
    filter_sequences.filter_genus("Testgenus", indir / "sample.fna", tmp_path / "kept.fasta", 0.7, tmp_path / "kept.json")
    from xspect2_b200.file_io import get_record_iterator
    assert [r.id for r in get_record_iterator(tmp_path / "kept.fasta")] == [f"c{i}" for i in range(5)]
    filter_sequences.filter_species("Testgenus", tid0, indir / "sample.fna", tmp_path / "spk.fasta", -1)
    assert {r.id for r in get_record_iterator(tmp_path / "spk.fasta")} >= {f"c{i}" for i in range(5)}
    # CLI (click choices are evaluated at import: import after HOME points at the synthetic root)
    from click.testing import CliRunner
    import xspect2_b200.main as main
    main = importlib.reload(main)
    runner = CliRunner()
    r = runner.invoke(main.cli, ["models", "list"])
    assert r.exit_code == 0 and "Testgenus" in r.output and "Species:" in r.output
    r = runner.invoke(main.cli, ["classify", "species", "-g", "Testgenus", "-i", str(indir / "sample.fna"), "-o",
                                 str(tmp_path / "cli.json"), "--sparse-sampling-step", "2", "-n", "--exclude-species", "999"])
    assert r.exit_code == 0, r.output
    cli = json.loads((tmp_path / "cli.json").read_text())
    assert cli["prediction"] == tid0 and any(" - species" in key for key in cli["hits"]["c0"])
    r = runner.invoke(main.cli, ["all", "-g", "Testgenus", "-i", str(indir / "sample.fna"), "-o", str(tmp_path / "all")])
    assert r.exit_code == 0, r.output
    assert "Pipeline completed successfully" in r.output
    outs = sorted(p.name for p in (tmp_path / "all").iterdir())
    assert any(n.startswith("genus_classification_") for n in outs) and any(n.startswith("species_classification_") for n in outs)
    assert len(list((tmp_path / "all" / "filtered_sequences").glob("genus_filtered_*.fasta"))) == 1
