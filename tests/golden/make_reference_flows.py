"""Golden vectors of the model-layer flows, produced by the REFERENCE'S OWN Python code.

Run in the build container (needs /root/reference; the result travels, this script's inputs do not):

    python tests/golden/make_reference_flows.py            # writes tests/golden/reference_flows.json

What runs: XspecT's real modules from /root/reference/src (models/probabilistic_filter_model.py,
probabilistic_single_filter_model.py, probabilistic_filter_svm_model.py, probabilistic_filter_mlst_model.py,
models/result.py, models/mlst_result.py, file_io.py) — their predict dispatch, per-record loops, exclude / display-name
handling, _count_kmers, ModelResult arithmetic, the SVM fit on scores.csv, the MLST chunking / `> 50` / per-allele sums /
has_sufficient_score, MlstResult — on the synthetic model directories of tests/model_fixtures.py.

What does NOT run: the reference's un-vendored native dependencies, which are not installable offline.  They are
replaced by stand-ins with the same call shapes: `cobs_index.Search` -> oracle.CobsOracle, `rbloom.Bloom` ->
oracle.BloomOracle, `Bio` -> xspect2_b200.seqio's Seq / SeqRecord / iterators, `slugify` -> model_management.slugify,
mappy / pysam / loguru / requests -> empty modules (their code is off this path).  The golden file therefore pins the
model layer (SURVEY.md 8(a) rows a2, a4, a5, a7's Python loop, a8, a9, a10) against the real reference code; the
third-party layer stays pinned only as DESIGN.md §2 says.

tests/test_reference_golden.py rebuilds the same model directories and inputs from the same seeds and compares (CPU:
the oracle's restated loops that every GPU test is compared with; GPU: the CUDA-backed models themselves)."""
from __future__ import annotations

import json
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
REF_SRC = Path("/root/reference/src")
OUT = Path(__file__).resolve().parent / "reference_flows.json"


def install_stand_ins(real: bool = False):
    """real=True (a box that has cobs-reloaded, rbloom and Biopython installed): the reference runs on its REAL
    dependencies and only modules that are absent get an empty stand-in; the output must then equal the committed file,
    which pins the third-party layer too (tests/test_live_reference.py::test_reference_flows_with_the_real_wheels)."""
    import importlib.util
    from oracle import oracle
    from xspect2_b200 import seqio
    from xspect2_b200.model_management import slugify

    def mod(name, **attrs):
        if real and importlib.util.find_spec(name.split(".")[0]) is not None:
            return None
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    fasta_io = mod("Bio.SeqIO.FastaIO", FastaIterator=seqio.FastaIterator)
    quality_io = mod("Bio.SeqIO.QualityIO", FastqPhredIterator=seqio.FastqPhredIterator)
    seq_io = mod("Bio.SeqIO", parse=lambda handle, fmt: seqio.parse(Path(handle), fmt), FastaIO=fasta_io, QualityIO=quality_io,
                 write=lambda records, handle, fmt: seqio.write_fasta(records, handle))
    bio_seq = mod("Bio.Seq", Seq=seqio.Seq)
    bio_rec = mod("Bio.SeqRecord", SeqRecord=seqio.SeqRecord)
    mod("Bio", SeqIO=seq_io, Seq=bio_seq, SeqRecord=bio_rec)
    mod("slugify", slugify=slugify)

    class Search:                      # cobs_index.Search(path, load_complete).search(str, step=...) -> [.doc_name, .score]
        def __init__(self, path, load_complete=False):
            self._o = oracle.CobsOracle(str(path))

        def search(self, query, step=1):
            return self._o.search(query, step)

    mod("cobs_index", Search=Search, SearchResult=object, DocumentList=object, ClassicIndexParameters=object,
        CompactIndexParameters=object)

    class Bloom:                       # rbloom.Bloom.load(path, hash_func) ; `kmer in bloom`
        def __init__(self, o):
            self._o = o

        @staticmethod
        def load(path, hash_func=None):
            return Bloom(oracle.BloomOracle(str(path), 0))

        def __contains__(self, kmer):
            return kmer in self._o

    mod("rbloom", Bloom=Bloom)
    for name in ("mappy", "pysam", "loguru", "requests"):
        mod(name, logger=types.SimpleNamespace(info=lambda *a, **k: None, warning=lambda *a, **k: None, error=lambda *a, **k: None))
    sys.path.insert(0, str(REF_SRC))


def build_world(base: Path):
    """The model directories and inputs of tests/test_reference_golden.py (same seeds, same order of RNG draws)."""
    from oracle import oracle
    from tests import model_fixtures as mf, synth
    models = base / "xspect-data" / "models"          # HOME = base makes this the data root of both implementations
    models.mkdir(parents=True)
    rng = np.random.default_rng(2026)
    sp_json, genomes, svm_genomes = mf.species_model(oracle, models, rng)
    ge_json = mf.genus_model(oracle, models, list(genomes.values()))
    ml_json, alleles = mf.mlst_model(oracle, models, rng)
    r2 = np.random.default_rng(77)
    gl = list(genomes.values())
    bases, b, e = synth.sample_reads(r2, gl, 60, (40, 400), sub=0.01, n_rate=0.002)
    recs = [(f"read{i}", bases[int(x):int(y)].tobytes().decode()) for i, (x, y) in enumerate(zip(b, e))]
    recs.append(("contig_long", np.concatenate([gl[0], gl[2][:3000]]).tobytes().decode()))
    recs.append(("read3", gl[1][:500].tobytes().decode()))        # duplicate id: overwrites the earlier read3
    # an "assembly" for the MLST model: random flanks around one allele of every locus (some mutated), long enough for
    # the chunked branch (>= 10 000 bp), plus a short record for the unchunked branch
    r3 = np.random.default_rng(78)
    parts = [synth.random_dna(r3, 4000)]
    chosen = {}
    for li, (locus, al) in enumerate(alleles.items()):
        name = list(al)[int(r3.integers(0, len(al)))]
        chosen[locus] = name
        s = al[name].copy()
        if li == 1:
            s = synth.mutate(r3, s, sub=0.01)
        parts += [s, synth.random_dna(r3, 3000)]
    assembly = np.concatenate(parts).tobytes().decode()
    short = alleles[next(iter(alleles))][next(iter(alleles[next(iter(alleles))]))].tobytes().decode()
    return dict(models=models, sp_json=sp_json, ge_json=ge_json, ml_json=ml_json, genomes=genomes, svm_genomes=svm_genomes,
                alleles=alleles, recs=recs, assembly=assembly, short=short, chosen=chosen)


def workflow_inputs(td: Path, w: dict) -> dict:
    """Input files of the workflow cases (shared with tests/test_reference_golden.py)."""
    from tests import model_fixtures as mf
    d = td / "wf"
    (d / "dir").mkdir(parents=True, exist_ok=True)
    mf.write_fasta(d / "sample.fna", w["recs"])
    mf.write_fasta(d / "dir" / "a.fna", w["recs"][:20])
    mf.write_fastq(d / "dir" / "b.fastq", w["recs"][20:45])
    mf.write_fasta(d / "assembly.fna", [("asm1", w["assembly"]), ("short1", w["short"])])
    return {"sample": d / "sample.fna", "dir": d / "dir", "assembly": d / "assembly.fna", "out": d}


def fasta_records(path: Path) -> list:
    from xspect2_b200.file_io import get_record_iterator
    return [[r.id, str(r.seq)] for r in get_record_iterator(path)] if path.exists() else None


def workflows(td: Path, w: dict) -> dict:
    """classify.py / filter_sequences.py / model_management.py / file_io.prepare_input_output_paths of the reference,
    run as a user would (HOME points at the synthetic data root)."""
    import os
    os.environ["HOME"] = str(td)
    import xspect.classify as cl
    import xspect.filter_sequences as fs
    import xspect.model_management as mm
    import xspect.models.probabilistic_filter_mlst_model as ref_mlst
    from xspect.definitions import get_xspect_model_path
    from xspect.file_io import prepare_input_output_paths

    assert get_xspect_model_path() == w["models"]
    inp = workflow_inputs(td, w)
    o = inp["out"]
    ids = list(w["genomes"])
    rel = lambda p: str(Path(p).relative_to(w["models"]))
    g = {"model_management": {
        "genus_model_path": rel(mm.get_genus_model_path("Testgenus")), "species_model_path": rel(mm.get_species_model_path("Testgenus")),
        "mlst_model_path": rel(mm.get_mlst_model_path("abaumannii", "Oxford")), "is_svm_model": mm.is_svm_model("testgenus-species"),
        "is_svm_model_genus": mm.is_svm_model("testgenus-genus"), "models": mm.get_models(),
        "display_names": mm.get_model_display_names("testgenus-species"), "mlst_schemes": mm.get_available_mlst_schemes(),
        "metadata_keys": list(mm.get_model_metadata("testgenus-species").keys())}}
    paths, get_out = prepare_input_output_paths(inp["dir"])
    g["prepare_paths"] = {"dir_inputs": [p.name for p in paths], "dir_outputs": [get_out(i, o / "res.json").name for i in range(len(paths))],
                          "file_output": prepare_input_output_paths(inp["sample"])[1](0, o / "res.json").name}
    cl.classify_species("Testgenus", inp["sample"], o / "sp.json", step=2, display_name=True, exclude_ids=[ids[3]])
    g["classify_species_text"] = (o / "sp.json").read_text()
    cl.classify_species("Testgenus", inp["dir"], o / "spd.json")
    g["classify_species_dir"] = {n: json.loads((o / n).read_text()) for n in ("spd_1.json", "spd_2.json")}
    cl.classify_genus("Testgenus", inp["sample"], o / "ge.json", step=3)
    g["classify_genus"] = json.loads((o / "ge.json").read_text())

    class Handler:
        def get_strain_type_name(self, highest_results, post_url):
            return {"ST": "golden", "received": highest_results}

    ref_mlst.PubMLSTHandler = Handler
    for limit in (False, True):
        cl.classify_mlst(inp["assembly"], "abaumannii", "Oxford", o / f"ml{int(limit)}.json", limit)
        g[f"classify_mlst_limit{int(limit)}"] = json.loads((o / f"ml{int(limit)}.json").read_text())
    fs.filter_genus("Testgenus", inp["sample"], o / "kept.fasta", 0.7, o / "kept.json", 2)
    g["filter_genus"] = {"records": fasta_records(o / "kept.fasta"), "classification": json.loads((o / "kept.json").read_text())}
    for name, thr in (("thr05", 0.5), ("best", -1)):
        fs.filter_species("Testgenus", ids[0], inp["sample"], o / f"sp_{name}.fasta", thr)
        g[f"filter_species_{name}"] = {"records": fasta_records(o / f"sp_{name}.fasta")}
    fs.filter_genus("Testgenus", inp["dir"], o / "keptd.fasta", 0.99)
    g["filter_genus_dir"] = {n: fasta_records(o / n) for n in ("keptd_1.fasta", "keptd_2.fasta")}
    return g


def cli_runs(td: Path, w: dict) -> dict:
    """The reference's click commands (main.py): `models list` and the full `all` pipeline of BASELINE config 3
    (main.py:84-188: filter_genus -> classify_species on the filtered directory -> MLST when the prediction is 470)."""
    import re
    from click.testing import CliRunner
    import xspect.main as ref_main
    import xspect.models.probabilistic_filter_mlst_model as ref_mlst

    class Handler:
        def get_strain_type_name(self, highest_results, post_url):
            return {"ST": "golden", "received": highest_results}

    ref_mlst.PubMLSTHandler = Handler
    inp = workflow_inputs(td, w)
    runner = CliRunner()
    g = {}
    r = runner.invoke(ref_main.cli, ["models", "list"])
    assert r.exit_code == 0, r.output
    g["models_list_output"] = r.output
    outdir = inp["out"] / "all"
    r = runner.invoke(ref_main.cli, ["all", "-g", "Testgenus", "-i", str(inp["sample"]), "-o", str(outdir), "-t", "0.7"])
    assert r.exit_code == 0, r.output
    norm = lambda t: re.sub(r"[0-9a-f]{8}-[0-9a-f]{4}-[0-9a-f]{4}-[0-9a-f]{4}-[0-9a-f]{12}", "<run>", t)
    g["all_output"] = norm(r.output).replace(str(outdir), "<outdir>")
    files = {}
    for p in sorted(outdir.rglob("*")):
        if p.is_file():
            key = norm(str(p.relative_to(outdir)))
            files[key] = fasta_records(p) if p.suffix == ".fasta" else json.loads(norm(p.read_text()))
    g["all_files"] = files
    return g


def result_dict(res) -> dict:
    d = res.to_dict()
    # JSON object order is the dict order the reference produced: keep it as lists of pairs where order is part of the contract
    return {"to_dict": d, "hits_order": {rid: list(h.items()) for rid, h in res.hits.items()}}


def main() -> None:
    import argparse
    global OUT, REF_SRC
    ap = argparse.ArgumentParser()
    ap.add_argument("--real", action="store_true", help="use the installed cobs_index / rbloom / Bio instead of the stand-ins")
    ap.add_argument("--out", default=str(OUT))
    ap.add_argument("--reference-src", default=str(REF_SRC))
    args = ap.parse_args()
    OUT, REF_SRC = Path(args.out), Path(args.reference_src)
    install_stand_ins(real=args.real)
    from Bio.Seq import Seq
    from Bio.SeqRecord import SeqRecord
    from tests import model_fixtures as mf
    from xspect.models.probabilistic_filter_model import ProbabilisticFilterModel
    from xspect.models.probabilistic_filter_mlst_model import ProbabilisticFilterMlstSchemeModel
    from xspect.models.probabilistic_filter_svm_model import ProbabilisticFilterSVMModel
    from xspect.models.probabilistic_single_filter_model import ProbabilisticSingleFilterModel
    import xspect.models.probabilistic_filter_mlst_model as ref_mlst

    out = {"_generator": "tests/golden/make_reference_flows.py", "_reference": "XspecT, /root/reference/src/xspect (models/*.py, file_io.py)",
           "_stand_ins": "cobs_index -> oracle.CobsOracle, rbloom -> oracle.BloomOracle, Bio -> xspect2_b200.seqio"}
    # (the header is the same with --real: the file must not depend on which implementation of the natives ran)
    with tempfile.TemporaryDirectory() as td:
        w = build_world(Path(td))
        fasta, fastq = Path(td) / "in.fna", Path(td) / "in.fastq"
        mf.write_fasta(fasta, w["recs"])
        mf.write_fastq(fastq, w["recs"])
        rl = [SeqRecord(Seq(s), rid) for rid, s in w["recs"]]
        ids = list(w["genomes"])

        # ---- species model: predict over the input kinds, steps, exclude ids, display names; ModelResult arithmetic
        sp = ProbabilisticFilterModel.load(w["sp_json"])
        cases = {}
        for name, kw in {"step1": {}, "step3": {"step": 3}, "exclude": {"exclude_ids": [ids[1], ids[4]]},
                         "display": {"display_name": True}, "exclude_display_step2": {"exclude_ids": [ids[0]], "display_name": True, "step": 2}}.items():
            res = sp.predict(fasta, **kw)
            cases[name] = result_dict(res)
            cases[name]["scores"] = res.get_scores()
            cases[name]["total_hits"] = res.get_total_hits()
        cases["single_record"] = result_dict(sp.predict(rl[5]))
        cases["record_list_head"] = result_dict(sp.predict(rl[:7], step=2))
        cases["fastq_equals_fasta"] = sp.predict(fastq).to_dict() == sp.predict(fasta).to_dict()
        cases["count_kmers"] = {f"{len(r.seq)}/{st}": sp._count_kmers(r, st) for r in rl[:12] for st in (1, 2, 3, 7)}
        res = sp.predict(fasta)
        cases["filtered_labels"] = {f"{tid}@{thr}": res.get_filtered_subsequence_labels(tid, thr) for tid in ids[:3] for thr in (0.3, 0.7, 0.99)}
        out["species"] = cases

        # ---- species + SVM: prediction from scores.csv (bug-compatible exclude ids)
        svm = ProbabilisticFilterSVMModel.load(w["sp_json"])
        sv = {}
        for acc, (tid, g) in list(w["svm_genomes"].items())[::4]:
            rec = SeqRecord(Seq(g.tobytes().decode()), acc)
            r0 = svm.predict(rec)
            r1 = svm.predict(rec, exclude_ids=[ids[2]])
            sv[acc] = {"label": tid, "prediction": r0.prediction, "prediction_excluding": r1.prediction,
                       "scores_total": r0.get_scores()["total"], "scores_total_excluding": r1.get_scores()["total"]}
        sv["file"] = {"prediction": svm.predict(fasta).prediction, "prediction_step3": svm.predict(fasta, step=3).prediction}
        out["svm"] = sv

        # ---- genus model (Bloom filter): the per-k-mer Python loop
        ge = ProbabilisticSingleFilterModel.load(w["ge_json"])
        g = {}
        for name, kw in {"step1": {}, "step4": {"step": 4}}.items():
            res = ge.predict(fasta, **kw)
            g[name] = result_dict(res)
            g[name]["scores"] = res.get_scores()
        res = ge.predict(fasta)
        g["filtered_labels"] = {str(thr): res.get_filtered_subsequence_labels("Testgenus", thr) for thr in (0.3, 0.7, 1.0)}
        out["genus"] = g

        # ---- MLST scheme model: chunked and unchunked branches, per-locus results, sufficient-score flag
        class Handler:                                  # the PubMLST POST is a network side effect inside the path
            def get_strain_type_name(self, highest_results, post_url):
                return {"ST": "golden", "received": highest_results}

        ref_mlst.PubMLSTHandler = Handler
        ml = ProbabilisticFilterMlstSchemeModel.load(w["ml_json"])
        m = {"chosen_alleles": w["chosen"]}
        for name, seq in (("assembly", w["assembly"]), ("short", w["short"])):
            for step in (1, 2):
                hits = ml.calculate_hits(Seq(seq), step=step) if "step" in ml.calculate_hits.__code__.co_varnames else ml.calculate_hits(Seq(seq))
                m[f"{name}_step{step}"] = json.loads(json.dumps(hits, default=lambda o: o.__dict__))
        asm_fa = Path(td) / "assembly.fna"
        mf.write_fasta(asm_fa, [("asm1", w["assembly"])])
        pres = ml.predict(asm_fa)
        m["predict_file"] = pres.to_dict() if hasattr(pres, "to_dict") else json.loads(json.dumps(pres, default=lambda o: o.__dict__))
        out["mlst"] = m
        out["workflows"] = workflows(Path(td), w)
        out["cli"] = cli_runs(Path(td), w)
    OUT.write_text(json.dumps(out, separators=(",", ":"), sort_keys=False))      # compact: a fixture, not a document
    print("wrote", OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
