"""Side measurements of the other BASELINE.json configs (not the bench headline): genus Bloom stage (config 3,
stage 1), MLST on assembled genomes (config 4), a wide-row index (config 5 geometry, single-GPU slice).
Test infrastructure: fixtures are built with the oracle's writers.  Run on a GPU box:
    python tests/perf_other_configs.py [bloom] [mlst] [wide]
Prints one JSON line per config; numbers are recorded in profiles/."""
import json
import os
import struct
import sys
import tempfile
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle  # noqa: E402
from tests import model_fixtures as mf, synth as tsynth  # noqa: E402
from xspect2_b200 import engine, synth  # noqa: E402
from xspect2_b200._abi import XS_U8  # noqa: E402

PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
dev = torch.device("cuda", 0)


def timed(fn, steps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def bloom(td: Path):
    n_reads, L, k = 10_000_000, 150, 21
    n_bytes = 13_800_000_008 // 8
    genome = synth.synth_genome(4_060_000, seed=1, n_rate=0.0001)
    gen = torch.Generator(device=dev).manual_seed(11)
    bits = torch.randint(0, 256, (n_bytes,), generator=gen, device=dev, dtype=torch.uint8).cpu().numpy()
    bt = oracle._BloomT(bits.ctypes.data, n_bytes * 8, 6, k)
    oracle.lib().xso_bloom_insert(oracle.C.byref(bt), bits.ctypes.data, genome.ctypes.data, genome.size)
    p = td / "filter.bloom"
    with open(p, "wb") as f:
        f.write(struct.pack("<Q", 6))
        f.write(bits.tobytes())
    del bits
    bf = engine.BloomFilter(p, k)
    frac = float(os.environ.get("XS_BLOOM_FRAC", 0.6))          # share of reads drawn from the genome in the filter
    reads = synth.synth_reads(genome, n_reads, L, seed=4, frac_genome=frac, device=dev)
    hb, he = synth.fixed_offsets(n_reads, L)
    d_b = torch.from_numpy(hb.view(np.int64)).to(dev)
    d_e = torch.from_numpy(he.view(np.int64)).to(dev)
    out = torch.empty(n_reads, dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    run = lambda: bf.query_device(reads.data_ptr(), n_reads * L, d_b.data_ptr(), d_e.data_ptr(), n_reads, 1, out.data_ptr(), s)
    sample = 20000
    exp = oracle.BloomOracle(p, k).hits_batch(reads[: sample * L].cpu().numpy(), hb[:sample], he[:sample], 1, threads=8)
    lookups = n_reads * (L - k + 1)
    ref_out = None
    # the direct kernel (k_bloom), then the bucketed kernels (k_bbucket_*: the default for batches >= 32 Mi windows)
    # ... "adaptive" = the default: a sampling kernel estimates the member fraction and the device picks the path
    for path in ("direct", "bucketed", "adaptive"):
        bf.set_bucketed(path != "direct", member_pct=0 if path == "bucketed" else 35)
        out.zero_()
        engine.profile_enable(True)
        engine.profile_read()
        ms = timed(run)
        phases, launches = engine.profile_read_phases()
        engine.profile_enable(False)
        assert np.array_equal(out[:sample].cpu().numpy().astype(np.uint32), exp)
        if ref_out is None:
            ref_out = out.clone()
        else:
            assert torch.equal(ref_out, out)
        frac_member = float(out.sum().item()) / lookups
        steps = max(launches[0], 1) if path == "direct" else max(launches[1], 1)
        print(json.dumps({"config": "cfg3 stage 1: genus Bloom, 13.8e9 bits, k_bloom=6, 10M x 150bp reads", "path": path, "ms_per_step": ms,
                          "kernel_ms[direct,emit,fetch,reduce] per pass": [x / 7 for x in phases], "launches": launches,
                          "lookups_per_sec": lookups / ms * 1e3, "reads_per_sec": n_reads / ms * 1e3,
                          "member_fraction": frac_member, "parity_sample_reads": sample, "identical_outputs": True}), flush=True)


def mlst(td: Path):
    rng = np.random.default_rng(5)
    models = td / "models"
    models.mkdir()
    p, alleles = mf.mlst_model(oracle, models, rng, n_loci=7, n_alleles=600, k=31)
    from xspect2_b200.models.probabilistic_filter_mlst_model import ProbabilisticFilterMlstSchemeModel
    from xspect2_b200.seqio import Seq

    class NoNet:
        def get_strain_type_name(self, hr, url):
            return "offline"

    model = ProbabilisticFilterMlstSchemeModel.load(p)
    model.pubmlst_handler = NoNet()
    genomes = []
    for g in range(4):
        parts = [tsynth.random_dna(rng, 500_000)]
        for locus in model.loci:
            parts += [alleles[locus][f"Allele_ID_{1 + (g * 37) % 600}"], tsynth.random_dna(rng, 500_000)]
        genomes.append(np.concatenate(parts).tobytes().decode())
    model.calculate_hits(Seq(genomes[0]))
    t0 = time.perf_counter()
    outs = [model.calculate_hits(Seq(g)) for g in genomes]
    dt = (time.perf_counter() - t0) / len(genomes)
    # parity of one locus of one genome against the oracle's chunk loop
    l0 = next(iter(model.loci))
    ref = oracle.mlst_locus_scores(oracle.CobsOracle(model.get_cobs_index_path(l0), load_complete=False), genomes[0],
                                   model.avg_locus_bp_size[0], 1)
    assert list(outs[0][1]["All results"][l0].items()) == list(ref.items())
    lookups = 7 * (len(genomes[0]) - 30)
    # all genomes and all loci in one xs_mlst_query call (what predict(Path) does for a multi-record file)
    arrs = [np.frombuffer(g.encode(), np.uint8) for g in genomes] * 2          # 8 assemblies
    sizes = np.array([a.size for a in arrs], np.uint64)
    end = np.cumsum(sizes, dtype=np.uint64)
    cat = engine.pinned_empty(int(end[-1]), np.uint8)
    cat[:] = np.concatenate(arrs)
    idx = [srch.index for srch in model.indices]
    engine.mlst_query(idx, model.avg_locus_bp_size, cat, end - sizes, end, 1)
    t0 = time.perf_counter()
    for _ in range(3):
        res = engine.mlst_query(idx, model.avg_locus_bp_size, cat, end - sizes, end, 1)
    dtb = (time.perf_counter() - t0) / 3 / len(arrs)
    got0 = [(idx[0].names[d], v) for d, v in zip(res[0][0][0].tolist(), res[0][0][1].tolist())]
    assert got0 == list(ref.items())
    print(json.dumps({"config": "cfg4: MLST 7 loci x 600 alleles compact k=31, 4 Mbp genomes",
                      "calculate_hits_s_per_genome": dt, "calculate_hits_lookups_per_sec": lookups / dt,
                      "xs_mlst_query_8_genomes_s_per_genome": dtb, "xs_mlst_query_lookups_per_sec": lookups / dtb,
                      "genome_bp": len(genomes[0]),
                      "note": "index ~ tens of MB: L2 resident; host buffers in, ordered allele lists out; chunk threshold/sum epilogue on the device"}), flush=True)


def wide(td: Path):
    # XS_WIDE_D / XS_WIDE_S: other geometries of the wide-row kernel (e.g. D=300 -> 38-byte rows at a 64-byte stride)
    D, h, k, S = int(os.environ.get("XS_WIDE_D", 10_000)), 7, 21, int(os.environ.get("XS_WIDE_S", 4_000_000))
    row = (D + 7) // 8
    names = [f"d{i}" for i in range(D)]
    gen = torch.Generator(device=dev).manual_seed(6)
    p = td / "index.cobs_classic"
    with open(p, "wb") as f:
        f.write(synth.classic_header(k, 1, names, S, h))
        for r0 in range(0, S, 1 << 19):
            n = min(1 << 19, S - r0)
            a = torch.randint(0, 256, (n, row), generator=gen, device=dev, dtype=torch.uint8)
            b = torch.randint(0, 256, (n, row), generator=gen, device=dev, dtype=torch.uint8)
            f.write((a & b).cpu().numpy().tobytes())          # fill 0.25
    n_reads, L = int(os.environ.get('XS_PERF_READS', 1_000_000)), 150
    genome = synth.synth_genome(1_000_000, seed=7)
    reads = synth.synth_reads(genome, n_reads, L, seed=8, device=dev)
    hb, he = synth.fixed_offsets(n_reads, L)
    d_b = torch.from_numpy(hb.view(np.int64)).to(dev)
    d_e = torch.from_numpy(he.view(np.int64)).to(dev)
    out = torch.empty((n_reads, D), dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    sample = 300
    exp = oracle.CobsOracle(p, load_complete=False).counts_batch(reads[: sample * L].cpu().numpy(), hb[:sample], he[:sample], 1, threads=8)
    lookups = n_reads * (L - k + 1)
    algo = lookups * h * row + n_reads * L * 3 // 8 + n_reads * D
    # rows of 17 .. 128 bytes: k_cobs_mid (default) and k_cobs_wide (XS_NO_MID_KERNEL=1, read at open time) on the same file
    for no_mid in (("0", "1") if row <= 128 else ("0",)):
        os.environ["XS_NO_MID_KERNEL"] = no_mid
        ix = engine.CobsIndex(p)
        engine.profile_enable(True)
        engine.profile_read()
        ms = timed(lambda: ix.query_device(reads.data_ptr(), n_reads * L, d_b.data_ptr(), d_e.data_ptr(), n_reads, 1, XS_U8, out.data_ptr(), s), steps=3, warm=1)
        kms, kn = engine.profile_read()
        engine.profile_enable(False)
        assert np.array_equal(out[:sample].cpu().numpy().astype(np.uint32), np.minimum(exp, 255))
        kernel_ms = kms / max(kn, 1)
        print(json.dumps({"config": f"rows on one GPU: D={D} h={h} S={S} (row {row} B, stride {ix.info.row_stride} B), {n_reads} x 150bp reads",
                          "kernel": ix.kernel, "ms_per_step": ms, "kernel_ms": kernel_ms, "lookups_per_sec": lookups / ms * 1e3,
                          "achieved_GBps": algo / kernel_ms / 1e6, "frac_of_hbm_peak": algo / kernel_ms / 1e6 / PEAK,
                          "dram_fetch_bound_frac": row / max(128, ix.info.row_stride),
                          "checksum": int(out.to(torch.int64).sum().item()), "parity_sample_reads": sample}), flush=True)
        ix.close()
    os.environ.pop("XS_NO_MID_KERNEL", None)


def twostage(td: Path):
    """Config 3 on one GPU: genus Bloom -> threshold 0.7 -> species COBS (+SVM) on 10 M reads, fused on the device."""
    import bench
    from xspect2_b200.models.probabilistic_filter_svm_model import ProbabilisticFilterSVMModel
    from xspect2_b200.models.probabilistic_single_filter_model import ProbabilisticSingleFilterModel
    from xspect2_b200.pipeline import genus_then_species
    from xspect2_b200.seqio import SequenceBatch
    n_reads, L, k = 10_000_000, 150, 21
    models = td / "models"
    (models / "synth-species").mkdir(parents=True)
    (models / "synth-genus").mkdir(parents=True)
    genome = synth.synth_genome(bench.GENOME_LEN, seed=1, n_rate=0.0001)
    rows, valid = engine.kmer_rows(genome, bench.K, bench.H, bench.SIG_SIZE)
    rows = rows[valid.astype(bool)]
    names = synth.write_classic_index(models / "synth-species" / "index.cobs_classic", n_docs=bench.D, k=k, num_hashes=bench.H,
                                      sig_size=bench.SIG_SIZE, seed=2, plant={0: rows.reshape(-1), 1: rows[: int(rows.shape[0] * 0.63)].reshape(-1)},
                                      device=dev)
    rng = np.random.default_rng(3)
    with open(models / "synth-species" / "scores.csv", "w") as f:
        f.write("file," + ",".join(sorted(names)) + ",label_id\n")
        for i, lab in enumerate(sorted(names)):
            for rep in range(4):
                x = np.round(rng.uniform(0, 0.2, size=len(names)), 2)
                x[i] = round(1.0 - 0.05 * rep, 2)
                f.write(f"acc{i}_{rep}," + ",".join(str(v) for v in x) + f",{lab}\n")
    meta = {"model_slug": "synth-species", "k": k, "model_display_name": "Synth", "author": None, "author_email": None,
            "model_type": "Species", "model_class": "ProbabilisticFilterSVMModel", "display_names": {n: f"Synth sp{n}" for n in names},
            "fpr": 0.01, "num_hashes": 7, "training_accessions": None, "kernel": "rbf", "C": 1.0, "svm_accessions": None}
    (models / "synth-species.json").write_text(json.dumps(meta))
    n_bytes = 13_800_000_008 // 8
    gen = torch.Generator(device=dev).manual_seed(11)
    bits = torch.randint(0, 256, (n_bytes,), generator=gen, device=dev, dtype=torch.uint8).cpu().numpy()
    bt = oracle._BloomT(bits.ctypes.data, n_bytes * 8, 6, k)
    oracle.lib().xso_bloom_insert(oracle.C.byref(bt), bits.ctypes.data, genome.ctypes.data, genome.size)
    with open(models / "synth-genus" / "filter.bloom", "wb") as f:
        f.write(struct.pack("<Q", 6))
        f.write(bits.tobytes())
    del bits
    gmeta = {"model_slug": "synth-genus", "k": k, "model_display_name": "Synth", "author": None, "author_email": None,
             "model_type": "Genus", "model_class": "ProbabilisticSingleFilterModel", "display_names": {"Synth": "Synth"},
             "fpr": 0.01, "num_hashes": 1, "training_accessions": None}
    (models / "synth-genus.json").write_text(json.dumps(gmeta))
    genus = ProbabilisticSingleFilterModel.load(models / "synth-genus.json")
    species = ProbabilisticFilterSVMModel.load(models / "synth-species.json")
    reads = synth.synth_reads(genome, n_reads, L, seed=4, device=dev)
    h_bases = engine.pinned_empty(n_reads * L, np.uint8)
    torch.from_numpy(h_bases).copy_(reads)
    del reads
    hb, he = synth.fixed_offsets(n_reads, L)
    batch = SequenceBatch(None, h_bases, hb, he, None, np.zeros(0, np.uint8), np.zeros(n_reads, np.uint64))
    genus_then_species(genus, species, batch, 0.7, 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = genus_then_species(genus, species, batch, 0.7, 1)
    dt = time.perf_counter() - t0
    kept = int(out["kept"].sum())
    # parity of both stages on a sample
    sample = 20000
    eg = oracle.BloomOracle(models / "synth-genus" / "filter.bloom", k).hits_batch(h_bases, hb[:sample], he[:sample], 1, threads=8)
    assert np.array_equal(out["genus_hits"][:sample], eg)
    ki = out["kept_index"][:2000]
    es = oracle.CobsOracle(models / "synth-species" / "index.cobs_classic").counts_batch(h_bases, hb[ki], he[ki], 1, threads=8)
    assert np.array_equal(out["best_hits"][:2000], es.max(axis=1)) and np.array_equal(out["best"][:2000], es.argmax(axis=1))
    print(json.dumps({"config": "cfg3 on one GPU: 10M x 150bp reads, genus Bloom (13.8e9 bits) -> keep score >= 0.7 -> species COBS D=90 + SVM, fused",
                      "s_per_pass": dt, "reads_per_sec": n_reads / dt, "kept_reads": kept, "prediction": out["prediction"],
                      "lookups_per_sec": (n_reads + kept) * (L - k + 1) / dt, "note": "host (pinned) reads in, calls/totals/prediction out"}), flush=True)


def assembly(td: Path):
    """Config 1: one ~4 Mbp assembly (4 contigs, FASTA on disk) through the species model's predict, next to the CPU
    oracle driven like the reference (one search per contig, one host thread)."""
    import bench
    from tests import model_fixtures as mfx
    from xspect2_b200.models.probabilistic_filter_model import ProbabilisticFilterModel
    models = td / "models"
    (models / "synth-species").mkdir(parents=True)
    genome = synth.synth_genome(bench.GENOME_LEN, seed=1, n_rate=0.0001)
    rows, valid = engine.kmer_rows(genome, bench.K, bench.H, bench.SIG_SIZE)
    rows = rows[valid.astype(bool)]
    names = synth.write_classic_index(models / "synth-species" / "index.cobs_classic", n_docs=bench.D, k=bench.K, num_hashes=bench.H,
                                      sig_size=bench.SIG_SIZE, seed=2, plant={0: rows[: int(rows.shape[0] * 0.3)].reshape(-1)}, device=dev)
    meta = {"model_slug": "synth-species", "k": bench.K, "model_display_name": "Synth", "author": None, "author_email": None,
            "model_type": "Species", "model_class": "ProbabilisticFilterModel", "display_names": {n: f"Synth sp{n}" for n in names},
            "fpr": 0.01, "num_hashes": 7, "training_accessions": None}
    (models / "synth-species.json").write_text(json.dumps(meta))
    cuts = [0, 3_900_000, 4_000_000, 4_050_000, bench.GENOME_LEN]
    recs = [(f"contig_{i + 1}", genome[a:b]) for i, (a, b) in enumerate(zip(cuts, cuts[1:]))]
    fa = td / "assembly.fna"
    mfx.write_fasta(fa, recs)
    t0 = time.perf_counter()
    model = ProbabilisticFilterModel.load(models / "synth-species.json")
    t_load = time.perf_counter() - t0
    model.predict(fa)
    t0 = time.perf_counter()
    res = model.predict(fa)
    res.input_source = fa.name
    res.save(td / "assembly.json")
    t_gpu = time.perf_counter() - t0
    orc = oracle.CobsOracle(models / "synth-species" / "index.cobs_classic")
    t0 = time.perf_counter()
    ref_hits, ref_nk = oracle.reference_predict(orc, [(rid, s.tobytes().decode()) for rid, s in recs], bench.K)
    t_cpu = time.perf_counter() - t0
    assert res.hits == ref_hits and res.num_kmers == ref_nk
    lookups = sum(ref_nk.values())
    print(json.dumps({"config": "cfg1: one ~4 Mbp assembly (4 contigs) x D=90 index (S=150000001), predict + save JSON",
                      "model_load_s": t_load, "gpu_predict_save_s": t_gpu, "gpu_lookups_per_sec": lookups / t_gpu,
                      "cpu_oracle_1thread_s": t_cpu, "cpu_lookups_per_sec": lookups / t_cpu, "lookups": lookups,
                      "parity": "hits and num_kmers of all contigs equal the oracle (dict order included)"}), flush=True)


def workflow(td: Path):
    """The reference's `xspect all` stages through the workflow functions (files between stages, main.py:108-145):
    filter_genus (genus JSON + filtered FASTA) -> classify_species (species JSON) on a 1 M-read FASTQ."""
    import os
    rng = np.random.default_rng(12)
    root = td / "home" / "xspect-data" / "models"
    root.mkdir(parents=True)
    sp_json, genomes, _ = mf.species_model(oracle, root, rng, n_species=90, genome_len=20000, svm=True)
    mf.genus_model(oracle, root, list(genomes.values())[:60])
    os.environ["HOME"] = str(td / "home")
    from xspect2_b200 import classify, filter_sequences
    n_reads, L = 1_000_000, 150
    g = np.concatenate(list(genomes.values()))
    reads = synth.synth_reads(g, n_reads, L, seed=13, device=dev).cpu().numpy().reshape(n_reads, L)
    fq = td / "reads.fastq"
    qual = b"I" * L
    with open(fq, "wb") as f:
        for c0 in range(0, n_reads, 50000):
            f.write(b"".join(b"@read%d/1\n" % i + reads[i].tobytes() + b"\n+\n" + qual + b"\n" for i in range(c0, min(n_reads, c0 + 50000))))
    out = td / "out"
    (out / "filtered_sequences").mkdir(parents=True)
    t0 = time.perf_counter()
    filter_sequences.filter_genus("Testgenus", fq, out / "filtered_sequences" / "genus_filtered.fasta", 0.7, out / "genus.json")
    t1 = time.perf_counter()
    classify.classify_species("Testgenus", out / "filtered_sequences", out / "species.json")
    t2 = time.perf_counter()
    kept = sum(1 for line in open(out / "filtered_sequences" / "genus_filtered.fasta", "rb") if line.startswith(b">"))
    with open(out / "species_1.json") as fh:
        head = fh.read(200)
    assert "testgenus-species" in head
    print(json.dumps({"config": "workflow: filter_genus + classify_species (SVM) on a 1M-read FASTQ, files between the stages",
                      "filter_genus_s": t1 - t0, "classify_species_s": t2 - t1, "kept_reads": kept,
                      "genus_json_MB": (out / "genus.json").stat().st_size / 1e6, "species_json_MB": (out / "species_1.json").stat().st_size / 1e6,
                      "reads_per_sec_overall": n_reads / (t2 - t0)}), flush=True)


def train(td: Path):
    """Construction on the GPU at a fraction of the real model sizes: classic index of 90 documents x 4 Mbp
    (S = 38.4 M rows) and a genus Bloom filter over 360 Mbp."""
    n_docs, doc_len, k = 90, 4_000_000, 21
    gen = torch.Generator(device=dev).manual_seed(21)
    acgt = torch.from_numpy(synth.ACGT.copy()).to(dev)
    bases = acgt[torch.randint(0, 4, (n_docs * doc_len,), generator=gen, device=dev)].cpu().numpy()
    begin = np.arange(n_docs, dtype=np.uint64) * np.uint64(doc_len)
    end = begin + np.uint64(doc_len)
    names = [str(1000 + d) for d in range(n_docs)]
    t0 = time.perf_counter()
    engine.build_cobs(td / "index.cobs_classic", 1, k, 7, 0.01, names, bases, begin, end, np.arange(n_docs, dtype=np.uint32))
    t_cobs = time.perf_counter() - t0
    t0 = time.perf_counter()
    engine.build_bloom(td / "filter.bloom", k, n_docs * doc_len - k + 1, 0.01, bases, begin, end)
    t_bloom = time.perf_counter() - t0
    # self-check: every document scores 100 % on its own column, members are found
    ix = engine.CobsIndex(td / "index.cobs_classic")
    c = ix.counts(bases[5 * doc_len: 5 * doc_len + 100_000])
    assert int(c[5]) == 100_000 - k + 1 and int(np.delete(c, 5).max()) < 3000
    bf = engine.BloomFilter(td / "filter.bloom", k)
    assert bf.hits(bases[7 * doc_len: 7 * doc_len + 50_000]) == 50_000 - k + 1
    print(json.dumps({"config": "training: classic index 90 docs x 4 Mbp (h=7, fpr=0.01) and genus Bloom over 360 Mbp, host arrays -> files",
                      "cobs_build_s": t_cobs, "cobs_file_GB": (td / "index.cobs_classic").stat().st_size / 1e9,
                      "cobs_kmers_per_sec": n_docs * (doc_len - k + 1) / t_cobs,
                      "bloom_build_s": t_bloom, "bloom_file_GB": (td / "filter.bloom").stat().st_size / 1e9,
                      "bloom_kmers_per_sec": n_docs * (doc_len - k + 1) / t_bloom}), flush=True)


def api(td: Path):
    """FASTQ on disk -> per-read counts through the model API (native reader + batched query), vs the numbers above."""
    rng = np.random.default_rng(9)
    models = td / "models"
    models.mkdir()
    sp_json, genomes, _ = mf.species_model(oracle, models, rng, n_species=90, genome_len=20000, svm=False)
    from xspect2_b200.models.probabilistic_filter_model import ProbabilisticFilterModel
    model = ProbabilisticFilterModel.load(sp_json)
    n_reads, L = 2_000_000, 150
    g = np.concatenate(list(genomes.values()))
    reads = synth.synth_reads(g, n_reads, L, seed=10, device=dev).cpu().numpy().reshape(n_reads, L)
    fq = td / "reads.fastq"
    qual = b"I" * L
    with open(fq, "wb") as f:
        for c0 in range(0, n_reads, 50000):
            f.write(b"".join(b"@read%d/1\n" % i + reads[i].tobytes() + b"\n+\n" + qual + b"\n" for i in range(c0, min(n_reads, c0 + 50000))))
    from xspect2_b200.seqio import SequenceBatch
    model.predict_arrays(fq)                      # warm up (page cache, pools)
    t0 = time.perf_counter()
    batch = SequenceBatch.from_file(fq)
    t1 = time.perf_counter()
    res = model.predict_arrays(batch)
    t2 = time.perf_counter()
    best, tie = res.argmax()
    totals = res.total_hits()
    t3 = time.perf_counter()
    summ = model.predict_summary(batch)
    t4 = time.perf_counter()
    assert np.array_equal(np.asarray(summ["best"]), best.astype(np.uint32)) and summ["total_hits"] == totals
    t5 = time.perf_counter()
    full = model.predict(fq)                      # the reference's API call: file -> ModelResult
    t6 = time.perf_counter()
    full.input_source = fq.name
    full.save(td / "result.json")                 # ... -> the reference's JSON, through the native writer
    t7 = time.perf_counter()
    import json as _json
    with open(td / "result.json") as fh:
        head = fh.read(400)
    assert head.startswith('{\n    "model_slug": "testgenus-species"')
    sample = 5000
    hb, he = batch.begin[:sample], batch.end[:sample]
    exp = oracle.CobsOracle(model.get_cobs_index_path()).counts_batch(batch.bases, hb, he, 1, threads=8)
    assert np.array_equal(np.asarray(res.counts[:sample]).astype(np.uint32), exp)
    print(json.dumps({"config": "API: 2M-read FASTQ file -> ProbabilisticFilterModel.predict_arrays (D=90 test model)",
                      "file_MB": fq.stat().st_size / 1e6, "parse_s": t1 - t0, "query_s": t2 - t1, "argmax_totals_host_s": t3 - t2,
                      "reads_per_sec_file_to_counts": n_reads / (t2 - t0),
                      "summary_on_device_s": t4 - t3, "predict_s": t6 - t5, "save_json_s": t7 - t6,
                      "json_MB": (td / "result.json").stat().st_size / 1e6, "reads_per_sec_file_to_json": n_reads / (t7 - t5), "reads_per_sec_file_to_calls": n_reads / ((t1 - t0) + (t4 - t3)), "parity_sample_reads": sample}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["bloom", "mlst", "wide", "api", "twostage", "assembly"]
    with tempfile.TemporaryDirectory() as td:
        for w in which:
            sub = Path(td) / w
            sub.mkdir()
            {"bloom": bloom, "mlst": mlst, "wide": wide, "api": api, "twostage": twostage, "assembly": assembly, "workflow": workflow, "train": train}[w](sub)
