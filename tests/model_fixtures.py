"""Synthetic XspecT model directories in the reference's on-disk layout (test infrastructure):
<slug>.json + <slug>/index.cobs_classic (+ scores.csv), <slug>/filter.bloom, <slug>/<locus>.cobs_compact.
Index / filter files are written by the oracle's format-faithful writers; scores.csv rows are the oracle's own
rounded totals, which is how the reference's fit produces them (probabilistic_filter_svm_model.py:143-173)."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

from tests import synth


def write_fasta(path: Path, records: list[tuple[str, np.ndarray | str]], wrap: int = 80) -> None:
    with open(path, "w") as f:
        for rid, s in records:
            s = s if isinstance(s, str) else s.tobytes().decode()
            f.write(f">{rid}\n")
            for i in range(0, len(s), wrap):
                f.write(s[i:i + wrap] + "\n")
            if not s:
                f.write("\n")


def write_fastq(path: Path, records: list[tuple[str, np.ndarray | str]]) -> None:
    with open(path, "w") as f:
        for rid, s in records:
            s = s if isinstance(s, str) else s.tobytes().decode()
            f.write(f"@{rid}\n{s}\n+\n{'I' * len(s)}\n")


def species_model(oracle, base: Path, rng, genus: str = "Testgenus", n_species: int = 6, genome_len: int = 6000,
                  k: int = 21, svm: bool = True, kernel: str = "rbf"):
    """Species (+SVM) model; returns (json path, {tax_id: genome}, svm genomes)."""
    from xspect2_b200.model_management import slugify
    ids = [str(470 + 7 * i) for i in range(n_species)]
    docs = synth.make_genomes(rng, n_species, genome_len, shared=0.4)
    docs = {tid: v for tid, v in zip(ids, docs.values())}
    slug = slugify(f"{genus}-Species")
    (base / slug).mkdir(parents=True, exist_ok=True)
    oracle.write_classic(base / slug / "index.cobs_classic", docs, k=k, num_hashes=7, fpr=0.01)
    display = {tid: f"{genus} species{tid}" for tid in ids}
    # display_names deliberately NOT in sorted key order (the reference's exclude_ids quirk depends on it)
    display = dict(sorted(display.items(), key=lambda kv: kv[0][::-1]))
    meta = {"model_slug": slug, "k": k, "model_display_name": genus, "author": "t", "author_email": "t@e",
            "model_type": "Species", "model_class": "ProbabilisticFilterSVMModel" if svm else "ProbabilisticFilterModel",
            "display_names": display, "fpr": 0.01, "num_hashes": 7, "training_accessions": None}
    svm_genomes = {}
    if svm:
        meta |= {"kernel": kernel, "C": 1.0, "svm_accessions": None}
        orc = oracle.CobsOracle(base / slug / "index.cobs_classic")
        rows = []
        for tid in ids:
            for rep in range(3):
                g = synth.mutate(rng, docs[tid][0], sub=0.02 * (rep + 1))
                svm_genomes[f"acc_{tid}_{rep}"] = (tid, g)
                c = orc.counts(g)
                nk = g.size - k + 1
                sc = {name: round(int(v) / nk, 2) for name, v in zip(orc.names, c)}
                rows.append(f"acc_{tid}_{rep}," + ",".join(str(sc[n]) for n in sorted(sc)) + f",{tid}")
        (base / slug / "scores.csv").write_text("file," + ",".join(sorted(display)) + ",label_id\n" + "\n".join(rows))
    p = base / f"{slug}.json"
    p.write_text(json.dumps(meta, indent=4))
    return p, {tid: docs[tid][0] for tid in ids}, svm_genomes


def genus_model(oracle, base: Path, genomes: list[np.ndarray], genus: str = "Testgenus", k: int = 21):
    from xspect2_b200.model_management import slugify
    slug = slugify(f"{genus}-Genus")
    (base / slug).mkdir(parents=True, exist_ok=True)
    oracle.write_bloom(base / slug / "filter.bloom", genomes, k=k, fpr=0.01)
    meta = {"model_slug": slug, "k": k, "model_display_name": genus, "author": None, "author_email": None,
            "model_type": "Genus", "model_class": "ProbabilisticSingleFilterModel", "display_names": {genus: genus},
            "fpr": 0.01, "num_hashes": 1, "training_accessions": None}
    p = base / f"{slug}.json"
    p.write_text(json.dumps(meta, indent=4))
    return p


def mlst_model(oracle, base: Path, rng, organism: str = "abaumannii", scheme: str = "Oxford", n_loci: int = 3,
               n_alleles: int = 40, k: int = 21):
    """MLST scheme model (one compact index per locus); returns (json path, {locus: {allele name: sequence}})."""
    from xspect2_b200.model_management import slugify
    slug = slugify(f"{organism}-{scheme}-MLST")
    (base / slug).mkdir(parents=True, exist_ok=True)
    loci, sizes, alleles = {}, [], {}
    for li in range(n_loci):
        locus = f"Oxf_locus{li}"
        L = 380 + 40 * li
        cons = synth.random_dna(rng, L)
        docs = {}
        for a in range(n_alleles):
            s = cons.copy()
            for pos in rng.integers(0, L, size=int(rng.integers(1, 5))):
                s[pos] = synth.ACGT[rng.integers(0, 4)]
            docs[f"Allele_ID_{a + 1}"] = [s[: L - int(rng.integers(0, 12))]]
        oracle.write_compact(base / slug / f"{locus}.cobs_compact", docs, k=k, num_hashes=1, fpr=0.001)
        loci[locus] = n_alleles
        sizes.append(int(np.mean([v[0].size for v in docs.values()])))
        alleles[locus] = {n: v[0] for n, v in docs.items()}
    meta = {"model_slug": slug, "k": k, "model_display_name": scheme, "author": None, "author_email": None,
            "model_type": "MLST", "model_class": "ProbabilisticFilterMlstSchemeModel", "display_names": {}, "fpr": 0.001,
            "num_hashes": 1, "training_accessions": None, "organism": organism,
            "scheme_url": "https://rest.pubmlst.org/db/pubmlst_abaumannii_seqdef/schemes/1", "loci": loci,
            "average_locus_base_pair_size": sizes}
    p = base / f"{slug}.json"
    p.write_text(json.dumps(meta, indent=4))
    return p, alleles
