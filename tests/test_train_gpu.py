"""Training side on the GPU: `fit` of the four model classes builds the reference's model
files with xs_cobs_build / xs_bloom_build.  The files must be byte-identical to the oracle's format-faithful
writers on the same inputs, and the trained models must reproduce the shapes of the reference's known answers
(tests/test_probabilistic_filter_model.py:73-118, test_probabilistic_single_filter_model.py:42-56,
test_probabilistic_filter_svm_model.py:36-59, test_probabilistic_filter_mlst_model.py:82-99)."""
import csv
from pathlib import Path

import numpy as np
import pytest

from tests import model_fixtures as mf
from tests import synth

pytestmark = pytest.mark.gpu


def _assemblies(tmp_path, rng, n=3, contigs=3, length=6000):
    d = tmp_path / "assemblies"
    d.mkdir()
    genomes = {}
    anc = synth.random_dna(rng, length * contigs)
    for i in range(n):
        g = anc.copy()
        m = rng.random(g.size) < 0.35
        g[m] = synth.ACGT[rng.integers(0, 4, size=int(m.sum()))]
        g[int(rng.integers(0, g.size))] = ord("N")      # one N: 21 of ~18 000 windows cannot hit, the total still rounds to 1.0
        stem = f"GCF_00000{i}945.{i + 1}_ASM{i}v2_genomic"
        recs = [(f"NC_{i}{c} contig {c}", g[c * length:(c + 1) * length]) for c in range(contigs)]
        mf.write_fasta(d / f"{stem}.fna", recs, wrap=70)
        genomes[stem] = recs
    (d / "notes.txt").write_text("ignored")
    return d, genomes


def test_species_fit_builds_the_reference_layout(gpu, oracle, tmp_path):
    from xspect2_b200.models.probabilistic_filter_model import ProbabilisticFilterModel
    from xspect2_b200.seqio import Seq, SeqRecord
    rng = np.random.default_rng(1)
    d, genomes = _assemblies(tmp_path, rng)
    base = tmp_path / "xspect_data"
    base.mkdir()
    model = ProbabilisticFilterModel(21, "Test Filter", "John Doe", "john.doe@example.com", "Species", base)
    stems = sorted(genomes)
    model.fit(d, display_names={stems[0]: "first species"})
    names = [s.split(".")[0] for s in stems]
    assert model.display_names == {names[0]: "first species", names[1]: stems[1], names[2]: stems[2]}
    idx = base / "test-filter-species" / "index.cobs_classic"
    ref = tmp_path / "ref.cobs_classic"
    oracle.write_classic(ref, {n: [s for _, s in genomes[st]] for n, st in zip(names, stems)}, k=21, num_hashes=7, fpr=0.01)
    assert idx.read_bytes() == ref.read_bytes()
    # G1 / G2 shape: an 80-bp substring of one training genome scores 60 / step on it
    q = genomes[stems[0]][1][1][1000:1080].tobytes().decode()
    for step in (1, 2, 3, 4):
        hits = model.calculate_hits(Seq(q), step=step)
        assert hits[names[0]] == -(-60 // step) and list(hits)[0] == names[0]
    # G3 shape: every training file scores 1.0 on its own document
    for st, n in zip(stems, names):
        total = model.predict(d / f"{st}.fna").get_total_scores()
        assert total[n] == 1.0 and all(v < 1.0 for key, v in total.items() if key != n)
    model.save()
    again = ProbabilisticFilterModel.load(base / "test-filter-species.json")
    assert again.to_dict() == model.to_dict()
    assert again.predict(SeqRecord(Seq(q), "r")).hits == model.predict(SeqRecord(Seq(q), "r")).hits
    with pytest.raises(ValueError):
        model.fit(tmp_path / "missing")
    empty = tmp_path / "empty"
    empty.mkdir()
    with pytest.raises(ValueError, match="No valid files"):
        model.fit(empty)


def test_svm_fit_writes_scores_csv(gpu, oracle, tmp_path):
    from xspect2_b200.models.probabilistic_filter_svm_model import ProbabilisticFilterSVMModel
    rng = np.random.default_rng(2)
    d, genomes = _assemblies(tmp_path, rng)
    stems = sorted(genomes)
    names = [s.split(".")[0] for s in stems]
    svm_dir = tmp_path / "svm"
    for st, n in zip(stems, names):
        (svm_dir / n).mkdir(parents=True)
        mf.write_fasta(svm_dir / n / f"{st}.fna", genomes[st])            # the training genome itself: diagonal 1.0
    base = tmp_path / "xspect_data"
    base.mkdir()
    model = ProbabilisticFilterSVMModel(21, "Test Filter", "John Doe", "john.doe@example.com", "Species", base, "linear", 1.0)
    model.fit(d, svm_dir)
    scores_file = base / "test-filter-species" / "scores.csv"
    rows = list(csv.DictReader(open(scores_file)))
    assert len(rows) == 3 and list(rows[0])[0] == "file" and list(rows[0])[-1] == "label_id"
    for row in rows:
        label = row["label_id"]
        assert row[label] == "1.0"
        assert all(float(v) < 1 for key, v in row.items() if key not in ("label_id", "file", label))
    model.save()
    loaded = ProbabilisticFilterSVMModel.load(base / "test-filter-species.json")
    for st, n in zip(stems, names):
        assert loaded.predict(d / f"{st}.fna", step=50).prediction == n       # step-sampled prediction = folder name


def test_genus_fit_builds_the_bloom_file(gpu, oracle, tmp_path):
    from xspect2_b200.models.probabilistic_single_filter_model import ProbabilisticSingleFilterModel
    from xspect2_b200.seqio import Seq
    rng = np.random.default_rng(3)
    recs = [(f"c{i}", synth.mutate(rng, synth.random_dna(rng, 5000), n_rate=0.001, lower=0.002)) for i in range(3)]
    fa = tmp_path / "Acinetobacter.fasta"
    mf.write_fasta(fa, recs)
    base = tmp_path / "xspect_data"
    base.mkdir()
    model = ProbabilisticSingleFilterModel(21, "Acinetobacter", None, None, "Genus", base)
    model.fit(fa, "Acinetobacter")
    ref = tmp_path / "ref.bloom"
    oracle.write_bloom(ref, [s for _, s in recs], k=21, fpr=0.01)
    assert (base / "acinetobacter-genus" / "filter.bloom").read_bytes() == ref.read_bytes()
    assert model.display_names == {"Acinetobacter": "Acinetobacter"}
    seq = recs[0][1]
    assert model.calculate_hits(Seq(seq[100:122].tobytes().decode())) == {"Acinetobacter": 2}     # G5 shape
    model.save()
    loaded = ProbabilisticSingleFilterModel.load(base / "acinetobacter-genus.json")
    assert loaded.calculate_hits(Seq(seq[100:122].tobytes().decode())) == {"Acinetobacter": 2}
    assert loaded.predict(fa).get_total_scores() == {"Acinetobacter": 1.0}


def test_mlst_fit_builds_compact_indices(gpu, oracle, tmp_path):
    from xspect2_b200.models.probabilistic_filter_mlst_model import ProbabilisticFilterMlstSchemeModel
    from xspect2_b200.seqio import Seq, SeqRecord
    rng = np.random.default_rng(4)
    scheme = tmp_path / "scheme"
    alleles = {}
    for li, locus in enumerate(("Oxf_cpn60", "Oxf_gdhB")):
        (scheme / locus).mkdir(parents=True)
        L = 405 + 30 * li
        cons = synth.random_dna(rng, L)
        alleles[locus] = {}
        for a in range(1, 71):
            s = cons.copy()
            for pos in rng.integers(0, L, size=int(rng.integers(1, 5))):
                s[pos] = synth.ACGT[rng.integers(0, 4)]
            s = s[: L - int(rng.integers(0, 9))]
            alleles[locus][f"Allele_ID_{a}"] = s
            mf.write_fasta(scheme / locus / f"Allele_ID_{a}.fasta", [(f"{locus}_{a}", s)])
    base = tmp_path / "xspect_data"
    base.mkdir()
    model = ProbabilisticFilterMlstSchemeModel(21, "Oxford", base, "https://rest.pubmlst.org/x", "abaumannii")
    model.fit(scheme)
    assert model.loci == {"Oxf_cpn60": 70, "Oxf_gdhB": 70}
    for locus in model.loci:
        files = sorted(p.stem for p in (scheme / locus).iterdir())
        ref = tmp_path / f"{locus}.ref"
        oracle.write_compact(ref, {n: [alleles[locus][n]] for n in files}, k=21, num_hashes=1, fpr=0.001)
        assert model.get_cobs_index_path(locus).read_bytes() == ref.read_bytes()
    first = sorted(p.stem for p in (scheme / "Oxf_cpn60").iterdir())[0]
    assert model.avg_locus_bp_size[0] == alleles["Oxf_cpn60"][first].size
    model.save()
    loaded = ProbabilisticFilterMlstSchemeModel.load(base / "abaumannii-oxford-mlst.json")
    assert loaded.loci == model.loci and loaded.avg_locus_bp_size == model.avg_locus_bp_size

    class NoNet:
        def get_strain_type_name(self, hr, url):
            return "offline"

    loaded.pubmlst_handler = NoNet()
    a4 = alleles["Oxf_cpn60"]["Allele_ID_4"]
    res = loaded.predict(SeqRecord(Seq(a4.tobytes().decode()), "a4"))
    best = res.hits["a4"][0]["Strain type"]["Oxf_cpn60"]
    assert list(best.values()) == [a4.size - 21 + 1]                      # G9 shape: every k-mer of the allele hits
    assert res.hits["a4"][1]["All results"]["Oxf_cpn60"]["Allele_ID_4"] == a4.size - 21 + 1


def test_fit_models_then_cli(gpu, oracle, tmp_path, monkeypatch):
    """The model classes' `fit` (GPU index / filter construction) on local data, saved where the CLI looks for
    models, then the trained models through the CLI: `models list`, `classify species`, `classify genus`.  (The
    training workflows of train.py are out of scope; `fit` is the 8(f)-4 row.)"""
    import importlib
    import json
    from click.testing import CliRunner
    monkeypatch.setenv("HOME", str(tmp_path / "home"))
    (tmp_path / "home" / "xspect-data").mkdir(parents=True)
    rng = np.random.default_rng(6)
    data = tmp_path / "training"
    anc = synth.random_dna(rng, 9000)
    genomes = {}
    for sp in ("470", "471", "48296"):
        g = anc.copy()
        m = rng.random(g.size) < 0.4
        g[m] = synth.ACGT[rng.integers(0, 4, size=int(m.sum()))]
        genomes[sp] = g
        for part, folder in (("cobs", data / "cobs" / sp), ("svm", data / "svm" / sp)):
            folder.mkdir(parents=True)
            for rep in range(2):
                seq = g if part == "cobs" else synth.mutate(rng, g, sub=0.01 * (rep + 1))
                mf.write_fasta(folder / f"GCF_{sp}_{rep}.fna", [(f"{sp}_{rep}_c{c}", seq[c * 3000:(c + 1) * 3000]) for c in range(3)])
    import xspect2_b200.main as main
    main = importlib.reload(main)
    runner = CliRunner()
    from xspect2_b200.definitions import get_xspect_model_path
    from xspect2_b200.file_io import concatenate_metagenome, concatenate_species_fasta_files
    from xspect2_b200.models.probabilistic_filter_svm_model import ProbabilisticFilterSVMModel
    from xspect2_b200.models.probabilistic_single_filter_model import ProbabilisticSingleFilterModel
    species_dir = tmp_path / "species"
    species_dir.mkdir()
    concatenate_species_fasta_files(sorted((data / "cobs").iterdir()), species_dir)
    sp_model = ProbabilisticFilterSVMModel(k=21, model_display_name="Trained", author=None, author_email=None,
                                           model_type="Species", base_path=get_xspect_model_path(), kernel="rbf", c=1.0)
    sp_model.fit(species_dir, data / "svm", svm_step=2)
    sp_model.save()
    concatenate_metagenome(species_dir, tmp_path / "Trained.fasta")
    ge_model = ProbabilisticSingleFilterModel(k=21, model_display_name="Trained", author=None, author_email=None,
                                              model_type="Genus", base_path=get_xspect_model_path())
    ge_model.fit(tmp_path / "Trained.fasta", "Trained")
    ge_model.save()
    models = tmp_path / "home" / "xspect-data" / "models"
    assert (models / "trained-species.json").is_file() and (models / "trained-species" / "index.cobs_classic").is_file()
    assert (models / "trained-species" / "scores.csv").is_file() and (models / "trained-genus" / "filter.bloom").is_file()
    meta = json.loads((models / "trained-species.json").read_text())
    assert meta["model_class"] == "ProbabilisticFilterSVMModel" and sorted(meta["display_names"]) == ["470", "471", "48296"]
    main = importlib.reload(main)        # click choices are read at import
    r = runner.invoke(main.cli, ["models", "list"])
    assert "Trained" in r.output and "Genus:" in r.output and "Species:" in r.output
    sample = tmp_path / "sample.fna"
    mf.write_fasta(sample, [("contig1", synth.mutate(rng, genomes["471"], sub=0.005))])
    r = runner.invoke(main.cli, ["classify", "species", "-g", "Trained", "-i", str(sample), "-o", str(tmp_path / "sp.json")])
    assert r.exit_code == 0, r.output
    assert json.loads((tmp_path / "sp.json").read_text())["prediction"] == "471"
    r = runner.invoke(main.cli, ["classify", "genus", "-g", "Trained", "-i", str(sample), "-o", str(tmp_path / "ge.json")])
    assert r.exit_code == 0, r.output
    assert json.loads((tmp_path / "ge.json").read_text())["scores"]["total"]["Trained"] >= 0.8
