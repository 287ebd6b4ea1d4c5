"""Golden vectors produced by the reference's OWN Python code (tests/golden/make_reference_flows.py: XspecT's real
models/*.py, result.py, mlst_result.py, file_io.py from /root/reference, with its un-installable native dependencies
replaced by oracle-backed stand-ins of the same call shape) against

* CPU (`-m "not gpu"`): the oracle's restated per-record loops (oracle.reference_predict*, mlst_locus_scores — what every
  GPU model test is compared with), this package's ModelResult arithmetic and its bug-compatible SVM fit;
* GPU (`-m gpu`): the CUDA-backed model classes themselves, through their public predict / calculate_hits API.

The model directories and inputs are rebuilt from the generator's seeds.  Dict ORDER is part of the contract (cobs
result order -> ModelResult.hits -> JSON), so orders are compared, not only contents."""
import json
from pathlib import Path

import numpy as np
import pytest

from tests import model_fixtures as mf

GOLD = json.loads((Path(__file__).resolve().parent / "golden" / "reference_flows.json").read_text())
SPECIES_CASES = {"step1": {}, "step3": {"step": 3}, "exclude": {"exclude": [1, 4]}, "display": {"display_name": True},
                 "exclude_display_step2": {"exclude": [0], "display_name": True, "step": 2}}


@pytest.fixture(scope="module")
def world(tmp_path_factory, oracle):
    from tests.golden.make_reference_flows import build_world
    return build_world(tmp_path_factory.mktemp("golden-world"))


def _kw(world, spec):
    ids = list(world["genomes"])
    kw = {k: v for k, v in spec.items() if k != "exclude"}
    if "exclude" in spec:
        kw["exclude_ids"] = [ids[i] for i in spec["exclude"]]
    return kw


def _pairs(hits):
    return {rid: [[k, v] for k, v in h.items()] for rid, h in hits.items()}


# ------------------------------------------------------------------------------------------ CPU: restated loops, host classes
def _display(model_json: dict, hits: dict) -> dict:
    """The display-name re-keying of predict (probabilistic_filter_model.py:296-306), for the restated loop's output."""
    names, genus = model_json["display_names"], model_json["model_display_name"]
    return {rid: {f"{key} -{names.get(key, 'Unknown').replace(genus, '', 1)}": v for key, v in h.items()} for rid, h in hits.items()}


def test_restated_species_loop_equals_reference_code(world, oracle):
    from xspect2_b200.models.result import ModelResult
    meta = json.loads(Path(world["sp_json"]).read_text())
    orc = oracle.CobsOracle(Path(world["sp_json"]).parent / meta["model_slug"] / "index.cobs_classic")
    for name, spec in SPECIES_CASES.items():
        kw = _kw(world, spec)
        gold = GOLD["species"][name]
        hits, nk = oracle.reference_predict(orc, world["recs"], meta["k"], kw.get("exclude_ids"), kw.get("step", 1))
        if kw.get("display_name"):
            hits = _display(meta, hits)
        assert _pairs(hits) == gold["hits_order"], name
        assert nk == gold["to_dict"]["num_kmers"]
        # this package's ModelResult on those hits == the reference's ModelResult (scores, rounding, totals, JSON fields)
        res = ModelResult(meta["model_slug"], hits, nk, sparse_sampling_step=kw.get("step", 1))
        assert res.get_scores() == gold["scores"] and res.get_total_hits() == gold["total_hits"]
        assert list(res.get_total_hits().items()) == list(gold["total_hits"].items())
        assert json.loads(json.dumps(res.to_dict())) == gold["to_dict"]
    hits, nk = oracle.reference_predict(orc, world["recs"], meta["k"])
    res = ModelResult(meta["model_slug"], hits, nk)
    for key, labels in GOLD["species"]["filtered_labels"].items():
        tid, thr = key.split("@")
        assert res.get_filtered_subsequence_labels(tid, float(thr)) == labels
    for key, n in GOLD["species"]["count_kmers"].items():
        length, step = map(int, key.split("/"))
        assert oracle.count_kmers(length, meta["k"], step) == n


def test_svm_fit_and_exclude_quirk_equal_reference_code(world, oracle):
    """_get_svm + svm_input of this package (host code) on the restated loop's totals == the reference's prediction."""
    from xspect2_b200.models.probabilistic_filter_svm_model import ProbabilisticFilterSVMModel as M
    from xspect2_b200.models.result import ModelResult
    meta = json.loads(Path(world["sp_json"]).read_text())
    m = M.__new__(M)
    m.base_path, m.kernel, m.c, m.display_names = Path(world["sp_json"]).parent, meta["kernel"], meta["C"], meta["display_names"]
    m.slug = lambda: meta["model_slug"]
    orc = oracle.CobsOracle(m.base_path / meta["model_slug"] / "index.cobs_classic")
    ids = list(world["genomes"])
    for acc, (tid, g) in list(world["svm_genomes"].items())[::4]:
        gold = GOLD["svm"][acc]
        for ex, pk, sk in ((None, "prediction", "scores_total"), ([ids[2]], "prediction_excluding", "scores_total_excluding")):
            hits, nk = oracle.reference_predict(orc, [(acc, g.tobytes().decode())], meta["k"], ex, 1)
            res = ModelResult(meta["model_slug"], hits, nk)
            assert res.get_scores()["total"] == gold[sk] and list(res.get_scores()["total"].items()) == list(gold[sk].items())
            assert str(m._get_svm(ex).predict(m.svm_input(res))[0]) == gold[pk]
    for step, key in ((1, "prediction"), (3, "prediction_step3")):
        hits, nk = oracle.reference_predict(orc, world["recs"], meta["k"], None, step)
        res = ModelResult(meta["model_slug"], hits, nk, sparse_sampling_step=step)
        assert str(m._get_svm(None).predict(m.svm_input(res))[0]) == GOLD["svm"]["file"][key]


def test_restated_bloom_loop_equals_reference_code(world, oracle):
    from xspect2_b200.models.result import ModelResult
    meta = json.loads(Path(world["ge_json"]).read_text())
    bf = oracle.BloomOracle(Path(world["ge_json"]).parent / meta["model_slug"] / "filter.bloom", meta["k"])
    for name, step in (("step1", 1), ("step4", 4)):
        hits, nk = oracle.reference_predict_bloom(bf, "Testgenus", world["recs"], meta["k"], step)
        gold = GOLD["genus"][name]
        assert _pairs(hits) == gold["hits_order"] and nk == gold["to_dict"]["num_kmers"]
        assert ModelResult(meta["model_slug"], hits, nk, sparse_sampling_step=step).get_scores() == gold["scores"]
    hits, nk = oracle.reference_predict_bloom(bf, "Testgenus", world["recs"], meta["k"], 1)
    res = ModelResult(meta["model_slug"], hits, nk)
    for thr, labels in GOLD["genus"]["filtered_labels"].items():
        assert res.get_filtered_subsequence_labels("Testgenus", float(thr)) == labels


def test_restated_mlst_epilogue_equals_reference_code(world, oracle):
    """oracle.mlst_locus_scores (sequence_splitter chunks, `> 50`, per-allele sums, order) per locus == the reference's
    calculate_hits 'All results' / 'Strain type', chunked (assembly) and unchunked (short record) branches."""
    meta = json.loads(Path(world["ml_json"]).read_text())
    base = Path(world["ml_json"]).parent / meta["model_slug"]
    for name, seq in (("assembly", world["assembly"]), ("short", world["short"])):
        for step in (1, 2):
            gold = GOLD["mlst"][f"{name}_step{step}"]
            strain, allres = dict(gold[0]["Strain type"]), gold[1]["All results"]
            # step 1 ends with the PubMLST name, step 2 with the "not reliable" note (has_sufficient_score is false there)
            assert ("ST_Name" in strain) == (step == 1) and ("Attention:" in strain) == (step == 2)
            strain.pop("ST_Name", None)
            strain.pop("Attention:", None)
            for li, locus in enumerate(meta["loci"]):
                orc = oracle.CobsOracle(base / f"{locus}.cobs_compact", load_complete=False)
                sc = oracle.mlst_locus_scores(orc, seq, meta["average_locus_base_pair_size"][li], step)
                assert list(sc.items()) == list(allres[locus].items()), (name, step, locus)
                first = next(iter(sc))
                assert strain[locus] == {first: sc[first]}
    assert {locus: next(iter(v)) for locus, v in GOLD["mlst"]["assembly_step1"][0]["Strain type"].items() if locus != "ST_Name"} \
        == GOLD["mlst"]["chosen_alleles"]


# ------------------------------------------------------------------------------------------ GPU: the CUDA-backed models
@pytest.mark.gpu
def test_gpu_species_model_equals_reference_code(world, gpu, tmp_path):
    from xspect2_b200.models.probabilistic_filter_model import ProbabilisticFilterModel
    from xspect2_b200.seqio import Seq, SeqRecord
    model = ProbabilisticFilterModel.load(world["sp_json"])
    fasta, fastq = tmp_path / "in.fna", tmp_path / "in.fastq"
    mf.write_fasta(fasta, world["recs"])
    mf.write_fastq(fastq, world["recs"])
    for name, spec in SPECIES_CASES.items():
        res = model.predict(fasta, **_kw(world, spec))
        gold = GOLD["species"][name]
        assert _pairs(res.hits) == gold["hits_order"], name
        assert json.loads(json.dumps(res.to_dict())) == gold["to_dict"]
        assert res.get_scores() == gold["scores"] and list(res.get_total_hits().items()) == list(gold["total_hits"].items())
    rl = [SeqRecord(Seq(s), rid) for rid, s in world["recs"]]
    assert _pairs(model.predict(rl[5]).hits) == GOLD["species"]["single_record"]["hits_order"]
    assert _pairs(model.predict(rl[:7], step=2).hits) == GOLD["species"]["record_list_head"]["hits_order"]
    assert model.predict(fastq).to_dict() == model.predict(fasta).to_dict()
    res = model.predict(fasta)
    for key, labels in GOLD["species"]["filtered_labels"].items():
        tid, thr = key.split("@")
        assert res.get_filtered_subsequence_labels(tid, float(thr)) == labels


@pytest.mark.gpu
def test_gpu_svm_and_genus_models_equal_reference_code(world, gpu, tmp_path):
    from xspect2_b200.models.probabilistic_filter_svm_model import ProbabilisticFilterSVMModel
    from xspect2_b200.models.probabilistic_single_filter_model import ProbabilisticSingleFilterModel
    from xspect2_b200.seqio import Seq, SeqRecord
    svm = ProbabilisticFilterSVMModel.load(world["sp_json"])
    ids = list(world["genomes"])
    fasta = tmp_path / "in.fna"
    mf.write_fasta(fasta, world["recs"])
    for acc, (tid, g) in list(world["svm_genomes"].items())[::4]:
        gold = GOLD["svm"][acc]
        rec = SeqRecord(Seq(g.tobytes().decode()), acc)
        r0, r1 = svm.predict(rec), svm.predict(rec, exclude_ids=[ids[2]])
        assert (r0.prediction, r1.prediction) == (gold["prediction"], gold["prediction_excluding"])
        assert r0.get_scores()["total"] == gold["scores_total"] and r1.get_scores()["total"] == gold["scores_total_excluding"]
    assert svm.predict(fasta).prediction == GOLD["svm"]["file"]["prediction"]
    assert svm.predict(fasta, step=3).prediction == GOLD["svm"]["file"]["prediction_step3"]
    genus = ProbabilisticSingleFilterModel.load(world["ge_json"])
    for name, step in (("step1", 1), ("step4", 4)):
        res = genus.predict(fasta, step=step)
        gold = GOLD["genus"][name]
        assert _pairs(res.hits) == gold["hits_order"] and res.get_scores() == gold["scores"]
        assert json.loads(json.dumps(res.to_dict())) == gold["to_dict"]


@pytest.mark.gpu
def test_gpu_mlst_model_equals_reference_code(world, gpu, tmp_path):
    from xspect2_b200.models.probabilistic_filter_mlst_model import ProbabilisticFilterMlstSchemeModel
    from xspect2_b200.seqio import Seq

    class Handler:
        def get_strain_type_name(self, highest_results, post_url):
            return {"ST": "golden", "received": highest_results}

    model = ProbabilisticFilterMlstSchemeModel.load(world["ml_json"])
    model.pubmlst_handler = Handler()
    for name, seq in (("assembly", world["assembly"]), ("short", world["short"])):
        for step in (1, 2):
            out = model.calculate_hits(Seq(seq), step=step)
            gold = GOLD["mlst"][f"{name}_step{step}"]
            assert json.loads(json.dumps(out)) == gold, (name, step)
            for locus, sc in out[1]["All results"].items():
                assert list(sc.items()) == list(gold[1]["All results"][locus].items())
    asm = tmp_path / "assembly.fna"
    mf.write_fasta(asm, [("asm1", world["assembly"])])
    got = model.predict(asm).to_dict()
    gold = dict(GOLD["mlst"]["predict_file"])
    for d in (got, gold):
        d.pop("Input_source", None)                       # set by the classify workflow, not by predict
    assert json.loads(json.dumps(got)) == gold


# ------------------------------------------------------------------------------------------ workflows (classify.py, filter_sequences.py, ...)
WF = GOLD["workflows"]


@pytest.fixture()
def home(world, monkeypatch):
    """HOME = the directory that holds the synthetic xspect-data root (definitions.get_xspect_root_path)."""
    base = Path(world["models"]).parent.parent
    monkeypatch.setenv("HOME", str(base))
    return base


def test_model_management_and_path_fanout_equal_reference_code(world, home, tmp_path):
    from tests.golden.make_reference_flows import workflow_inputs
    from xspect2_b200 import definitions, model_management as mm
    from xspect2_b200.file_io import prepare_input_output_paths
    assert definitions.get_xspect_model_path() == Path(world["models"])
    rel = lambda p: str(Path(p).relative_to(world["models"]))
    g = WF["model_management"]
    assert rel(mm.get_genus_model_path("Testgenus")) == g["genus_model_path"]
    assert rel(mm.get_species_model_path("Testgenus")) == g["species_model_path"]
    assert rel(mm.get_mlst_model_path("abaumannii", "Oxford")) == g["mlst_model_path"]
    assert mm.is_svm_model("testgenus-species") is g["is_svm_model"] and mm.is_svm_model("testgenus-genus") is g["is_svm_model_genus"]
    assert mm.get_models() == g["models"] and mm.get_model_display_names("testgenus-species") == g["display_names"]
    assert mm.get_available_mlst_schemes() == g["mlst_schemes"]
    assert list(mm.get_model_metadata("testgenus-species").keys()) == g["metadata_keys"]
    inp = workflow_inputs(tmp_path, world)
    paths, get_out = prepare_input_output_paths(inp["dir"])
    assert [p.name for p in paths] == WF["prepare_paths"]["dir_inputs"]
    assert [get_out(i, tmp_path / "res.json").name for i in range(len(paths))] == WF["prepare_paths"]["dir_outputs"]
    assert prepare_input_output_paths(inp["sample"])[1](0, tmp_path / "res.json").name == WF["prepare_paths"]["file_output"]


def test_saved_result_bytes_equal_reference_code(world, oracle, tmp_path):
    """classify_species' output file: the reference's ModelResult.save text, byte for byte, from this package's ModelResult
    (hits of the restated loop, display-name keys, the SVM prediction of this package's host code)."""
    from xspect2_b200.models.probabilistic_filter_svm_model import ProbabilisticFilterSVMModel as M
    from xspect2_b200.models.result import ModelResult
    meta = json.loads(Path(world["sp_json"]).read_text())
    ids = list(world["genomes"])
    ex = [ids[3]]
    orc = oracle.CobsOracle(Path(world["sp_json"]).parent / meta["model_slug"] / "index.cobs_classic")
    hits, nk = oracle.reference_predict(orc, world["recs"], meta["k"], ex, 2)
    m = M.__new__(M)
    m.base_path, m.kernel, m.c, m.display_names = Path(world["sp_json"]).parent, meta["kernel"], meta["C"], meta["display_names"]
    m.slug = lambda: meta["model_slug"]
    plain = ModelResult(meta["model_slug"], hits, nk, sparse_sampling_step=2)
    pred = str(m._get_svm(ex).predict(m.svm_input(plain))[0])
    res = ModelResult(meta["model_slug"], _display(meta, hits), nk, sparse_sampling_step=2, prediction=pred, input_source="sample.fna")
    res.save(tmp_path / "sp.json")
    assert (tmp_path / "sp.json").read_text() == WF["classify_species_text"]


def test_filter_outputs_equal_reference_code(world, oracle, tmp_path):
    """filter_genus / filter_species: the labels this package's ModelResult keeps (threshold and best-score modes) from the
    restated loops' hits, written by its native FASTA filter (host code), equal the reference's filtered files record
    for record — a duplicated id included twice, as the reference's `record.id in included_ids` does."""
    from tests.golden.make_reference_flows import fasta_records, workflow_inputs
    from xspect2_b200.file_io import filter_sequences
    from xspect2_b200.models.result import ModelResult
    inp = workflow_inputs(tmp_path, world)
    ids = list(world["genomes"])
    gm = json.loads(Path(world["ge_json"]).read_text())
    bf = oracle.BloomOracle(Path(world["ge_json"]).parent / gm["model_slug"] / "filter.bloom", gm["k"])
    hits, nk = oracle.reference_predict_bloom(bf, "Testgenus", world["recs"], gm["k"], 2)
    res = ModelResult(gm["model_slug"], hits, nk, sparse_sampling_step=2, input_source="sample.fna")
    assert json.loads(json.dumps(res.to_dict())) == WF["filter_genus"]["classification"]
    filter_sequences(inp["sample"], tmp_path / "kept.fasta", res.get_filtered_subsequence_labels("Testgenus", 0.7))
    assert fasta_records(tmp_path / "kept.fasta") == WF["filter_genus"]["records"]
    sm = json.loads(Path(world["sp_json"]).read_text())
    orc = oracle.CobsOracle(Path(world["sp_json"]).parent / sm["model_slug"] / "index.cobs_classic")
    hits, nk = oracle.reference_predict(orc, world["recs"], sm["k"])
    res = ModelResult(sm["model_slug"], hits, nk)
    for name, thr in (("thr05", 0.5), ("best", -1)):
        filter_sequences(inp["sample"], tmp_path / f"sp_{name}.fasta", res.get_filtered_subsequence_labels(ids[0], thr))
        assert fasta_records(tmp_path / f"sp_{name}.fasta") == WF[f"filter_species_{name}"]["records"], name


@pytest.mark.gpu
def test_gpu_workflows_equal_reference_code(world, gpu, home, tmp_path, monkeypatch):
    """classify_species / classify_genus / classify_mlst / filter_genus / filter_species of this package, run as a user
    would, against the files the reference's workflow functions wrote for the same inputs."""
    import xspect2_b200.models.probabilistic_filter_mlst_model as mlst_mod
    from tests.golden.make_reference_flows import fasta_records, workflow_inputs
    from xspect2_b200 import classify, filter_sequences

    class Handler:
        def get_strain_type_name(self, highest_results, post_url):
            return {"ST": "golden", "received": highest_results}

    monkeypatch.setattr(mlst_mod, "PubMLSTHandler", Handler)
    inp = workflow_inputs(tmp_path, world)
    o = inp["out"]
    ids = list(world["genomes"])
    classify.classify_species("Testgenus", inp["sample"], o / "sp.json", step=2, display_name=True, exclude_ids=[ids[3]])
    assert (o / "sp.json").read_text() == WF["classify_species_text"]                   # byte for byte
    classify.classify_species("Testgenus", inp["dir"], o / "spd.json")
    for n, gold in WF["classify_species_dir"].items():
        assert json.loads((o / n).read_text()) == gold, n
    classify.classify_genus("Testgenus", inp["sample"], o / "ge.json", step=3)
    assert json.loads((o / "ge.json").read_text()) == WF["classify_genus"]
    for limit in (False, True):
        classify.classify_mlst(inp["assembly"], "abaumannii", "Oxford", o / f"ml{int(limit)}.json", limit)
        got, gold = json.loads((o / f"ml{int(limit)}.json").read_text()), WF[f"classify_mlst_limit{int(limit)}"]
        assert got == gold, limit
        for rid, res in got["Results"].items():                                         # allele order inside every locus
            for locus, sc in res[1]["All results"].items():
                assert list(sc.items()) == list(gold["Results"][rid][1]["All results"][locus].items())
    filter_sequences.filter_genus("Testgenus", inp["sample"], o / "kept.fasta", 0.7, o / "kept.json", 2)
    assert fasta_records(o / "kept.fasta") == WF["filter_genus"]["records"]
    assert json.loads((o / "kept.json").read_text()) == WF["filter_genus"]["classification"]
    for name, thr in (("thr05", 0.5), ("best", -1)):
        filter_sequences.filter_species("Testgenus", ids[0], inp["sample"], o / f"sp_{name}.fasta", thr)
        assert fasta_records(o / f"sp_{name}.fasta") == WF[f"filter_species_{name}"]["records"], name
    filter_sequences.filter_genus("Testgenus", inp["dir"], o / "keptd.fasta", 0.99)
    for n, gold in WF["filter_genus_dir"].items():
        assert fasta_records(o / n) == gold, n


# ------------------------------------------------------------------------------------------ CLI: `models list`, `xspect all`
CLI = GOLD["cli"]


def test_cli_models_list_equals_reference_code(world, home):
    """`xspect models list` of the reference's click tree == this package's (click choices are evaluated at import)."""
    import importlib

    from click.testing import CliRunner
    import xspect2_b200.main as main
    main = importlib.reload(main)
    r = CliRunner().invoke(main.cli, ["models", "list"])
    assert r.exit_code == 0 and r.output == CLI["models_list_output"]


def test_all_pipeline_of_the_reference_equals_the_restated_stages(world, oracle, tmp_path):
    """`xspect all` (main.py:84-188, BASELINE config 3's flow) as the reference ran it: genus classification -> records with
    round(hits / num_kmers, 2) >= 0.7 written to the filtered FASTA -> species classification (+ SVM) of that file -> MLST
    of that file because the prediction is 470.  Every file it wrote is reproduced here from the oracle's restated loops,
    this package's ModelResult / SVM host code / FASTA filter — and the kept set equals the exact integer threshold table
    of the fused device pipeline (pipeline.min_hits_table)."""
    from tests.golden.make_reference_flows import fasta_records, workflow_inputs
    from xspect2_b200.file_io import filter_sequences
    from xspect2_b200.models.probabilistic_filter_svm_model import ProbabilisticFilterSVMModel as M
    from xspect2_b200.models.result import ModelResult
    from xspect2_b200.pipeline import min_hits_table
    files = CLI["all_files"]
    inp = workflow_inputs(tmp_path, world)
    # ---- step 1: genus classification and filter
    gm = json.loads(Path(world["ge_json"]).read_text())
    bf = oracle.BloomOracle(Path(world["ge_json"]).parent / gm["model_slug"] / "filter.bloom", gm["k"])
    hits, nk = oracle.reference_predict_bloom(bf, "Testgenus", world["recs"], gm["k"], 1)
    genus = ModelResult(gm["model_slug"], hits, nk, input_source="sample.fna")
    assert json.loads(json.dumps(genus.to_dict())) == files["genus_classification_<run>.json"]
    labels = genus.get_filtered_subsequence_labels("Testgenus", 0.7)
    filter_sequences(inp["sample"], tmp_path / "kept.fasta", labels)
    kept = fasta_records(tmp_path / "kept.fasta")
    assert kept == files["filtered_sequences/genus_filtered_<run>.fasta"]
    table = min_hits_table(max(nk.values()), 0.7)             # the fused pipeline's keep rule, per record id
    assert {rid for rid in hits if hits[rid]["Testgenus"] >= table[nk[rid]]} == set(labels)
    # ---- step 2: species classification of the filtered file (directory input -> numbered output)
    sm = json.loads(Path(world["sp_json"]).read_text())
    orc = oracle.CobsOracle(Path(world["sp_json"]).parent / sm["model_slug"] / "index.cobs_classic")
    recs = [(rid, seq) for rid, seq in kept]
    shits, snk = oracle.reference_predict(orc, recs, sm["k"])
    m = M.__new__(M)
    m.base_path, m.kernel, m.c, m.display_names = Path(world["sp_json"]).parent, sm["kernel"], sm["C"], sm["display_names"]
    m.slug = lambda: sm["model_slug"]
    plain = ModelResult(sm["model_slug"], shits, snk)
    pred = str(m._get_svm(None).predict(m.svm_input(plain))[0])
    species = ModelResult(sm["model_slug"], shits, snk, prediction=pred, input_source="genus_filtered_<run>.fasta")
    gold = files["species_classification_<run>_1.json"]
    assert json.loads(json.dumps(species.to_dict())) == gold and pred == "470"
    assert [list(v.items()) for v in species.hits.values()] == [list(v.items()) for v in gold["hits"].values()]
    # ---- step 3: MLST of the filtered file (reads: the unchunked branch for every record)
    mm = json.loads(Path(world["ml_json"]).read_text())
    base = Path(world["ml_json"]).parent / mm["model_slug"]
    gold = files["mlst_classification_<run>_1.json"]
    assert gold["Scheme"] == "Oxford" and gold["Input_source"] == "genus_filtered_<run>.fasta" and list(gold["Results"]) == list(dict(recs))
    indices = [oracle.CobsOracle(base / f"{locus}.cobs_compact", load_complete=False) for locus in mm["loci"]]
    for rid, seq in dict(recs).items():
        allres = gold["Results"][rid][1]["All results"]
        strain = gold["Results"][rid][0]["Strain type"]
        for li, locus in enumerate(mm["loci"]):
            sc = oracle.mlst_locus_scores(indices[li], seq, mm["average_locus_base_pair_size"][li], 1)
            assert list(sc.items()) == list(allres[locus].items()), (rid, locus)
            first = next(iter(sc))
            assert strain[locus] == {first: sc[first]}
    assert "Step 3/3: Running MLST classification for abaumannii..." in CLI["all_output"]
