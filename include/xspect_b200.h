/* xspect_b200.h — C ABI of libxspect_b200.so: the B200-native replacement for the k-mer
 * scoring step behind XspecT's ProbabilisticFilterModel / ProbabilisticSingleFilterModel /
 * ProbabilisticFilterSVMModel / ProbabilisticFilterMlstSchemeModel prediction.
 *
 * The reference reaches this path through two third-party Python extensions
 * (cobs_index from cobs-reloaded, rbloom); every entry point below names the reference
 * call site (file:line under /root/reference) it replaces.  INTEGRATION.md shows the
 * ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; no CUDA, torch or C++ types in any signature
 *     (`void* stream` is a cudaStream_t passed as an opaque pointer; NULL = default stream)
 *   - every function returns an int: XS_OK (0) or a negative xs_status; the message for the
 *     calling thread is returned by xs_last_error()
 *   - handles own their device (HBM) memory until xs_*_close; callers own all input and
 *     output buffers; the index data is immutable after open, so queries from several host
 *     threads on one handle are safe.  How much they overlap depends on the path: small batches
 *     (direct kernels) run concurrently on their own streams; large batches that take the bucketed
 *     kernels (xs_cobs_set_bucketed / xs_bloom_set_bucketed) are serialised per handle — a mutex
 *     while they are enqueued plus an event chain on the device — because two bucketed sweeps
 *     would evict each other's L2-resident row range.  Two workers that need large batches in
 *     parallel on one GPU open the model twice (2.4 GB per copy for the species model).
 *   - the only library kernel on the path is cub::DeviceScan::ExclusiveSum over the windows-per-
 *     record array (0.03 % of a step); everything else is this library's own kernels
 *   - there is NO CPU fallback: every query needs a CUDA device and fails with XS_ERR_CUDA
 *     otherwise
 *   - sequences are described by two arrays seq_begin[i], seq_end[i] (byte offsets into
 *     `bases`); a contiguous batch passes the same offsets array twice, shifted by one
 *     (seq_end = seq_begin + 1); overlapping segments (MLST chunks) are allowed
 */
#ifndef XSPECT_B200_H
#define XSPECT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct xs_cobs xs_cobs;   /* a COBS classic/compact index resident in HBM */
typedef struct xs_bloom xs_bloom; /* an rbloom bit array resident in HBM */
typedef struct xs_comm xs_comm;   /* one rank's NCCL communicator of a document-column sharded index */

enum xs_status {
    XS_OK = 0,
    XS_ERR_ARG = -1,         /* bad argument (NULL, k out of range, step == 0, ...) */
    XS_ERR_IO = -2,          /* file missing / unreadable  -> FileNotFoundError / OSError */
    XS_ERR_FORMAT = -3,      /* not a COBS / rbloom file, size identity violated */
    XS_ERR_CUDA = -4,        /* CUDA runtime error, or no device */
    XS_ERR_NOMEM = -5,       /* host or device allocation failed */
    XS_ERR_UNSUPPORTED = -6, /* valid input this build does not handle (k > 32, ...) */
    XS_ERR_NCCL = -7         /* libnccl missing or an NCCL call failed (document-column sharded exchange) */
};

/* element type of the per-document count matrix written by xs_cobs_query* */
enum xs_dtype { XS_U8 = 1, XS_U16 = 2, XS_U32 = 4 };

/* what a COBS query does with a window that touches a byte outside upper-case ACGT
 * (third-party behaviour, SURVEY.md A.2.5a — one switch): */
enum xs_nonacgt_policy {
    XS_NONACGT_SKIP = 0,   /* window contributes no hits (default) */
    XS_NONACGT_LITERAL = 1 /* complement of such a byte is 0x00; hash min(literal, mapped revcomp) */
};

enum xs_cobs_kind { XS_COBS_CLASSIC = 1, XS_COBS_COMPACT = 2 };

typedef struct {
    uint32_t kind;          /* xs_cobs_kind */
    uint32_t term_size;     /* k */
    uint32_t canonicalize;  /* header flag */
    uint32_t num_hashes;    /* h */
    uint32_t n_docs_total;  /* documents in the file */
    uint32_t doc_begin;     /* this handle's document-column shard [doc_begin, doc_end) */
    uint32_t doc_end;
    uint32_t n_pages;       /* 1 for classic */
    uint64_t page_bytes;    /* bytes per row per page in the FILE */
    uint64_t row_stride;    /* bytes per row per page in HBM (re-strided, power of two or x16) */
    uint64_t sig_size_max;  /* largest signature_size over pages */
    uint64_t hbm_bytes;     /* device bytes held */
    int32_t device;
    int32_t policy;         /* xs_nonacgt_policy */
} xs_cobs_info_t;

typedef struct {
    uint64_t n_bits;
    uint64_t k_hashes;   /* rbloom's k (number of LCG probes) */
    uint32_t term_size;  /* k-mer length this handle hashes */
    int32_t device;
    uint64_t hbm_bytes;
} xs_bloom_info_t;

/* ---- library ------------------------------------------------------------------------ */
int xs_version(void);
const char* xs_last_error(void);           /* thread-local, never NULL */
int xs_device_count(int* n);
/* number of CUDA kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t xs_launch_count(void);
/* measurement support (bench.py's roofline): when enabled, every query brackets its dominant kernel
 * (k_cobs_narrow / k_cobs_wide / k_bloom) with CUDA events on the stream it is launched on;
 * xs_profile_read synchronises those events, returns the summed kernel time and launch count since
 * the last read, and clears them. */
/* Query workspaces and the bucketed path's scratch come from the device's stream-ordered memory pool and stay cached
 * there between queries (up to 24 GiB after a large batch).  xs_device_trim synchronises the device and returns the
 * cached memory to the driver — call it before loading another large index. */
int xs_device_trim(int device);
int xs_profile_enable(int on);
int xs_profile_read(double* kernel_ms, uint64_t* launches);
/* the same, split by kernel: [0] k_cobs_narrow / k_cobs_wide / k_bloom, [1] k_bucket_emit, [2] k_bucket_fetch,
 * [3] k_bucket_reduce (four entries each) */
int xs_profile_read_phases(double* kernel_ms, uint64_t* launches);
/* pinned host memory for the host-buffer query entry points (pageable memory also works,
 * but is copied through the driver's bounce buffer) */
int xs_host_alloc(uint64_t bytes, void** out);
int xs_host_free(void* p);

/* ---- COBS index ------------------------------------------------------------------------
 * Replaces cobs_index.Search(path, load_complete)
 *   probabilistic_filter_model.py:194,389; probabilistic_filter_svm_model.py:313;
 *   probabilistic_filter_mlst_model.py:142,188
 * Loads index.cobs_classic / <locus>.cobs_compact unchanged into HBM.  [doc_begin, doc_end)
 * selects a document-column shard (multiples of 8; (0, 0) = all documents; classic only).
 * Missing file -> XS_ERR_IO. */
int xs_cobs_open(const char* path, int device, uint32_t doc_begin, uint32_t doc_end, xs_cobs** out);
int xs_cobs_info(const xs_cobs* ix, xs_cobs_info_t* info);
/* document names of the whole file, '\n'-separated, no trailing NUL counted in *needed */
int xs_cobs_doc_names(const xs_cobs* ix, char* buf, uint64_t cap, uint64_t* needed);
/* Which reading of the file header matched.  The header layout of cobs-reloaded is restated from its published
 * source (not checkable offline); xs_cobs_open tries the documented field order first and a few neighbouring orders /
 * widths after it, and accepts one only when the end magic and the size identity
 * (file_size - data_offset == sum signature_size x row bytes) hold.  Compact files: the zero padding before the
 * end magic may be 0 or one whole page when the data is already aligned.  Never NULL. */
const char* xs_cobs_header_layout(const xs_cobs* ix);
/* The scoring kernel this handle's row layout selects for direct (non-bucketed) batches: "k_cobs_narrow" (rows of
 * <= 16 bytes), "k_cobs_pages" (compact index, several narrow pages), "k_cobs_mid" (one page of 32 / 64 / 128-byte
 * rows, 129 .. 1024 documents), "k_cobs_wide" (everything else; XS_FORCE_WIDE=1 / XS_NO_MID_KERNEL=1 at open time
 * force it for measurements).  Results do not depend on the choice.  Never NULL. */
const char* xs_cobs_kernel(const xs_cobs* ix);
/* The same header parse without a device (no CUDA call): what a file would load as. */
typedef struct {
    uint32_t kind, term_size, canonicalize, num_hashes, n_docs, n_pages;
    uint64_t page_bytes, sig_size_max, data_offset, file_size;
    char layout[64];
} xs_cobs_header_t;
int xs_cobs_probe_header(const char* path, xs_cobs_header_t* out);
/* Structural self-check of a loaded index: per local document, the fraction of set bits over `sample_rows` evenly
 * spaced rows of its page (0 = 65536).  An index built at its design load has max fill ~ fpr^(1/h) (0.518 for
 * h = 7, fpr = 0.01; cobs calc_signature_size), and a k-mer absent from a document scores fill^h there: a known
 * member sequence that scores near fill^h instead of near 1 means the hash / bit layout does not match the file
 * (engine.CobsIndex.selfcheck).  fill: [doc_end - doc_begin] doubles, host memory. */
int xs_cobs_doc_fill(const xs_cobs* ix, uint64_t sample_rows, double* fill);
int xs_cobs_set_policy(xs_cobs* ix, int policy);
/* Large batches against a large narrow-row classic index (16-byte rows, 256 MB .. 8.6 GB in HBM) are scored by the
 * bucketed kernels: probe records grouped by L2-sized row ranges, rows fetched from L2, ANDed per window in shared
 * memory.  Results are identical to the direct-gather kernel's.  enabled: 0/1; min_windows: smallest batch (sampled
 * windows) that takes this path (0 = keep); scratch_bytes: upper bound of the per-query device scratch (0 = keep);
 * bucket_shift: log2 rows per bucket, 0 = automatic (test hook: lets small indices exercise the path).
 * Defaults: enabled, 32 Mi windows, 24 GiB (never more than half of the free device memory). */
int xs_cobs_set_bucketed(xs_cobs* ix, int enabled, uint64_t min_windows, uint64_t scratch_bytes, uint32_t bucket_shift);
/* number of queries (device batches) of this handle that went through the bucketed kernels */
int xs_cobs_bucketed_queries(const xs_cobs* ix, uint64_t* n);
int xs_cobs_close(xs_cobs* ix);

/* Replaces the per-record loop around cobs Search.search(str(sequence), step=step)
 *   probabilistic_filter_model.py:227 (called from :291-297);
 *   probabilistic_filter_mlst_model.py:242,274
 * For every sequence i and local document d:
 *   out[i*n_local + d] = #{sampled windows p = 0, step, 2*step, ... <= len_i - k :
 *                          AND_{j<h} bit_d(row[XXH64(term_p, k, seed=j) % signature_size])}
 * Sequences shorter than k produce zeros (the len > k ValueError of
 * probabilistic_filter_model.py:224-225 is raised by the Python layer before the call).
 * XS_U8 / XS_U16 outputs saturate at 255 / 65535.
 * Host variant: all pointers are host memory; copies are pipelined inside; returns when
 * `out` is complete.  Device variant: all pointers are device memory on the handle's device;
 * asynchronous on `stream`. */
int xs_cobs_query(xs_cobs* ix, const uint8_t* bases, uint64_t n_bases, const uint64_t* seq_begin,
                  const uint64_t* seq_end, uint64_t n_seq, uint32_t step, int out_dtype, void* out);
int xs_cobs_query_device(xs_cobs* ix, const uint8_t* d_bases, uint64_t n_bases,
                         const uint64_t* d_seq_begin, const uint64_t* d_seq_end, uint64_t n_seq,
                         uint32_t step, int out_dtype, void* d_out, void* stream);

/* The same query with the score epilogue applied on the device (host buffers in and out): per record the first
 * document holding the maximum count, that count and how many documents share it, plus totals[d] = sum of the
 * counts of document d over all records (uint64 [n_local]).  Counts are kept as uint32 on the device; only 12
 * bytes per record come back instead of the count matrix.  Replaces the host loops that derive read-level
 * calls (scripts/benchmark/main.nf:417-436) and ModelResult.get_total_hits (models/result.py:76-90). */
int xs_cobs_classify(xs_cobs* ix, const uint8_t* bases, uint64_t n_bases, const uint64_t* seq_begin,
                     const uint64_t* seq_end, uint64_t n_seq, uint32_t step, uint32_t* best, uint32_t* best_count,
                     uint32_t* n_best, uint64_t* totals);

/* Replaces, for ALL loci of an MLST scheme and all records of an input in one call, the chunk loop of
 * ProbabilisticFilterMlstSchemeModel.calculate_hits:
 *   probabilistic_filter_mlst_model.py:236-256  records of >= chunk_from_len (10000) bases: sequence_splitter
 *       (:382-426) chunks of allele_len[l] bases (x10 from 1 Mbp, x100 from 10 Mbp, k-1 overlap, short remainder glued
 *       to the last chunk), one search per chunk, scores > min_chunk_score (50) kept (:362-380), summed per allele,
 *       dict in first-appearance order then stably sorted by descending sum
 *   probabilistic_filter_mlst_model.py:272-286  shorter records: one search, every allele in cobs result order
 * loci[l]: the locus' <locus>.cobs_compact handle (same device and k for all).  The bases are uploaded and packed
 * once; every locus' chunk count matrix stays on the device, where the chunks with a score above the threshold are
 * found and compacted; only those rows (normally one or two per record and locus) come back, and the library orders
 * them with the same std::partial_sort cobs uses for its result list.
 * Output, for locus l and record i (D_l = documents of locus l, off_l = n_seq * sum_{l' < l} D_l'):
 *   out_n[l * n_seq + i]                     number of (allele, score) pairs
 *   out_doc / out_score[off_l + i * D_l + j] document index / summed score of the j-th pair, j < out_n[...]
 * All pointers are host memory. */
int xs_mlst_query(xs_cobs* const* loci, uint32_t n_loci, const uint32_t* allele_len, const uint8_t* bases, uint64_t n_bases,
                  const uint64_t* seq_begin, const uint64_t* seq_end, uint64_t n_seq, uint32_t step, uint32_t min_chunk_score,
                  uint64_t chunk_from_len, uint32_t* out_n, uint32_t* out_doc, uint32_t* out_score);

/* Result order of cobs Search.search (what _convert_cobs_result_to_dict iterates,
 * probabilistic_filter_model.py:393-409; MLST tie-breaks, probabilistic_filter_mlst_model.py:254-256,284):
 * all documents, std::partial_sort by score descending.  Host-only helper. */
int xs_cobs_result_order(const uint32_t* scores, uint32_t n_docs, uint32_t* order);
/* the same for n_seq rows of a [n_seq x n_docs] uint32 count matrix (one call per predict batch) */
int xs_cobs_result_order_batch(const uint32_t* scores, uint64_t n_seq, uint32_t n_docs, uint32_t* order);

/* Score epilogue on the device, for callers that want read-level calls and file-level totals instead of the
 * whole count matrix: for every record the first document holding the maximum count (best), that count
 * (best_count) and the number of documents sharing it (n_best > 1 = the "ambiguous" tie rule of
 * scripts/benchmark/main.nf:417-436), and totals[d] += sum over records (ModelResult.get_total_hits,
 * models/result.py:76-90; the caller zeroes totals).  Any output pointer may be NULL.  All pointers are device
 * memory on `device`; asynchronous on `stream`. */
int xs_scores_reduce_device(const void* d_counts, uint64_t n_seq, uint32_t n_docs, int dtype, int device,
                            uint32_t* d_best, uint32_t* d_best_count, uint32_t* d_n_best, uint64_t* d_totals,
                            void* stream);

/* ---- Bloom filter ------------------------------------------------------------------------
 * Replaces rbloom.Bloom.load(path, hash_func=xxh3_64_intdigest)
 *   probabilistic_single_filter_model.py:155-158
 * term_size is the model's k (the .bloom file does not store it). */
int xs_bloom_open(const char* path, uint32_t term_size, int device, xs_bloom** out);
int xs_bloom_info(const xs_bloom* bf, xs_bloom_info_t* info);
/* The same bucketed path for the Bloom filter (probes grouped by 16 MB ranges of the bit array; results identical).
 * The bucketed kernels make all k probes of a window while the direct kernel stops at the first zero bit, so they win
 * only when enough windows are members: a sampling kernel estimates the member fraction on the device and the batch
 * goes through the bucketed kernels when it is >= member_pct per cent (default 35, the measured crossover; 0 = always;
 * -1 = keep).  Other arguments as xs_cobs_set_bucketed; bucket_shift = log2 bits per bucket. */
int xs_bloom_set_bucketed(xs_bloom* bf, int enabled, uint64_t min_windows, uint64_t scratch_bytes, uint32_t bucket_shift,
                          int member_pct);
int xs_bloom_bucketed_queries(const xs_bloom* bf, uint64_t* n);
int xs_bloom_close(xs_bloom* bf);

/* Replaces sum(1 for kmer in _generate_kmers(sequence, step) if kmer in self.bf)
 *   probabilistic_single_filter_model.py:122-124 with :161-180
 * out_hits[i] = #{sampled windows whose min(kmer, revcomp) (Biopython complement, literal
 * bytes, no filtering) passes all k_hashes LCG probes seeded by XXH3-64}. */
int xs_bloom_query(xs_bloom* bf, const uint8_t* bases, uint64_t n_bases, const uint64_t* seq_begin,
                   const uint64_t* seq_end, uint64_t n_seq, uint32_t step, uint32_t* out_hits);
/* Replaces `obj in bf` (rbloom.Bloom.__contains__) on caller-supplied terms: terms holds n_terms strings of term_size
 * bytes back to back, each hashed as it is (no canonicalisation — the reference canonicalises in _generate_kmers,
 * probabilistic_single_filter_model.py:175-180, before it asks the filter); out[i] = 1 for members. */
int xs_bloom_contains(xs_bloom* bf, const uint8_t* terms, uint64_t n_terms, uint8_t* out);
int xs_bloom_query_device(xs_bloom* bf, const uint8_t* d_bases, uint64_t n_bases,
                          const uint64_t* d_seq_begin, const uint64_t* d_seq_end, uint64_t n_seq,
                          uint32_t step, uint32_t* d_out_hits, void* stream);

/* ---- construction (training side) ------------------------------------------------------------------
 * xs_cobs_build replaces cobs.classic_construct_list / cobs.compact_construct_list
 *   (probabilistic_filter_model.py:186-192; probabilistic_filter_mlst_model.py:132-141): every sequence i
 *   (bases[seq_begin[i]:seq_end[i]]) belongs to document seq_doc[i]; all windows of a document are hashed like the
 *   query does and OR-ed into its column.  sig_size = 0 derives the signature size from fpr
 *   (ceil(max k-mers per document * -h / ln(1 - fpr^(1/h)))); compact: documents sorted by size, 8 * page_size
 *   per page (page_size = 0 -> floor(sqrt(n_docs / 8)), at least 1).  Writes the index file in the layout of
 *   SURVEY.md A.1 / A.3 ([UNVERIFIED-3P] like the reader).
 * xs_bloom_build replaces Bloom(expected_items, fpr, hash_func=xxh3_64_intdigest) + add per k-mer + save
 *   (probabilistic_single_filter_model.py:83-96). */
int xs_cobs_build(const char* out_path, int device, int kind, uint32_t k, uint32_t num_hashes, double fpr,
                  uint32_t canonicalize, uint64_t sig_size, uint64_t page_size, const char* names, uint32_t n_docs,
                  const uint8_t* bases, uint64_t n_bases, const uint64_t* seq_begin, const uint64_t* seq_end,
                  const uint32_t* seq_doc, uint64_t n_seq);
int xs_bloom_build(const char* out_path, int device, uint32_t k, uint64_t expected_items, double fpr,
                   const uint8_t* bases, uint64_t n_bases, const uint64_t* seq_begin, const uint64_t* seq_end,
                   uint64_t n_seq);

/* ---- FASTA / FASTQ ingest (host only) -----------------------------------------------------------
 * Replaces the Biopython record iteration that feeds the path (SeqIO.parse behind
 * file_io.get_record_iterator, file_io.py:47-79; consumed at probabilistic_filter_model.py:291-310) with one
 * pass that lays a whole file out as query input.  format: 1 = fasta, 2 = fastq (chosen by file extension on the
 * Python side, like the reference).  xs_fastx_open parses once to size the buffers; xs_fastx_read fills
 *   bases[n_bases], seq_begin/seq_end[n_records] (offsets into bases), ids[n_id_bytes] (record ids back to back:
 *   the first whitespace-delimited word of each title line, Biopython's record.id) and id_end[n_records].
 * Sequence bytes are kept as they are; FASTA line breaks, '\r' and blanks are removed; wrapped FASTQ is accepted;
 * malformed FASTQ -> XS_ERR_FORMAT with Biopython's message. */
typedef struct xs_fastx xs_fastx;
int xs_fastx_open(const char* path, int format, xs_fastx** out);
int xs_fastx_stats(const xs_fastx* fx, uint64_t* n_records, uint64_t* n_bases, uint64_t* n_id_bytes);
int xs_fastx_read(const xs_fastx* fx, uint8_t* bases, uint64_t* seq_begin, uint64_t* seq_end, char* ids, uint64_t* id_end);
/* file_io.filter_sequences (file_io.py:166-191): write the records with keep[i] != 0 to out_path as FASTA the way
 * Bio.SeqIO.write does for parsed records ('>' + title line, sequence wrapped at 60 columns). */
int xs_fastx_filter_fasta(const xs_fastx* fx, const uint8_t* keep, const char* out_path);
int xs_fastx_close(xs_fastx* fx);

/* File -> read-level calls in one streamed call (the loop of probabilistic_filter_model.py:291-310 over
 * file_io.get_record_iterator(path), followed by the per-read argmax / tie rule of scripts/benchmark/main.nf:417-436):
 * blocks of ~block_bytes (0 = 96 MiB) of the FASTA / FASTQ file are cut at record starts, parsed by all host threads
 * straight into page-locked staging buffers, and copied, scored (xs_cobs_query_device) and reduced
 * (xs_scores_reduce_device) on three streams while the next block is being parsed; only 12 bytes per record come
 * back.  Wrapped (multi-line) FASTQ is not accepted here (XS_ERR_FORMAT) — use xs_fastx_open + xs_cobs_classify.
 * Results: per record the first best document, its count, the number of documents sharing it, the sequence length
 * (num_kmers = ceil((len - k + 1) / step) is the caller's arithmetic; n_short counts records with len <= k, which
 * make the reference raise ValueError), record ids, and per-document totals. */
typedef struct xs_file_calls xs_file_calls;
int xs_cobs_classify_file(xs_cobs* ix, const char* path, int format, uint32_t step, uint64_t block_bytes, xs_file_calls** out);
int xs_file_calls_info(const xs_file_calls* r, uint64_t* n_records, uint64_t* n_bases, uint64_t* n_id_bytes, uint64_t* n_short,
                       uint64_t* n_docs, double* parse_s, double* total_s);
int xs_file_calls_read(const xs_file_calls* r, uint32_t* best, uint32_t* best_count, uint32_t* n_best, uint64_t* seq_len,
                       char* ids, uint64_t* id_end, uint64_t* totals);
/* the same arrays without a copy: pointers into the result, valid until xs_file_calls_free */
int xs_file_calls_view(const xs_file_calls* r, const uint32_t** best, const uint32_t** best_count, const uint32_t** n_best,
                       const uint64_t** seq_len, const char** ids, const uint64_t** id_end, const uint64_t** totals);
int xs_file_calls_free(xs_file_calls* r);

/* ---- result writer (host only) -------------------------------------------------------------------
 * Replaces ModelResult.save = json.dumps(self.to_dict(), indent=4) (models/result.py:151-189) for results held as
 * a count matrix: writes byte-identical JSON without building the nested dictionaries.  counts is [* x n_docs]
 * uint32; rec_index[i] is the matrix row of the i-th emitted record (dict semantics for duplicate ids are resolved
 * by the caller); rec_keys / doc_keys are the JSON-escaped keys back to back with their end offsets; documents
 * with doc_include[d] == 0 are left out (exclude_ids); per record the documents appear in cobs result order;
 * prefix / suffix are the JSON text before "hits" and after "num_kmers". */
int xs_result_write_json(const char* path, const char* prefix, const char* suffix, const uint32_t* counts,
                         uint32_t n_docs, const uint64_t* rec_index, uint64_t n_emit, const char* rec_keys,
                         const uint64_t* rec_key_end, const uint64_t* num_kmers, const char* doc_keys,
                         const uint64_t* doc_key_end, const uint8_t* doc_include);

/* ---- document-column sharded index: the exchange step (BASELINE config 5) ----------------
 * An index too large for one GPU is split by document columns: rank g opens its range with
 * xs_cobs_open(path, dev, doc_begin_g, doc_end_g), every rank scores every record against its columns, and the
 * per-record score rows are combined with an NCCL all-gather over NVLink.  The reference has no counterpart (one
 * process, one cobs Search over the whole file, probabilistic_filter_model.py:389); the consumer of the combined
 * rows is what the reference's benchmark does with them — per-read argmax with ties = ambiguous
 * (scripts/benchmark/main.nf:417-436).
 *
 * libnccl.so.2 is bound at run time (the one already loaded in the process — torch's — else the system's, else
 * $XSPECT_B200_NCCL); no NCCL, no sharded exchange (XS_ERR_NCCL) — everything else works without it.
 * Rank 0 creates an id (xs_comm_unique_id, 128 bytes), hands it to the other ranks by any means (a file, MPI,
 * torch.distributed), and every rank calls xs_comm_init(id, rank, world, device). */
int xs_comm_unique_id(uint8_t* id128);
int xs_comm_init(const uint8_t* id128, int rank, int world, int device, xs_comm** out);
int xs_comm_info(const xs_comm* c, int* rank, int* world, int* nccl_version);
int xs_comm_destroy(xs_comm* c);
/* like xs_cobs_query_device with an explicit output row length ld >= local documents (the padding columns are
 * zero): every rank writes score tiles of one common width, so the all-gather needs no pad or concat pass */
int xs_cobs_query_device_ld(xs_cobs* ix, const uint8_t* d_bases, uint64_t n_bases, const uint64_t* d_seq_begin,
                            const uint64_t* d_seq_end, uint64_t n_seq, uint32_t step, int out_dtype, uint64_t ld, void* d_out,
                            void* stream);
/* ncclAllGather of one score tile: d_local = this rank's [n_seq x row_bytes] block, d_all = [world][n_seq][row_bytes]
 * (rank-major), asynchronous on `stream` */
int xs_allgather_scores(xs_comm* c, const void* d_local, uint64_t n_seq, uint64_t row_bytes, void* d_all, void* stream);
/* ncclAllReduce(sum) of per-document uint64 totals, in place (file-level scores of a read-sharded or column-sharded job) */
int xs_allreduce_totals(xs_comm* c, uint64_t* d_totals, uint64_t n, void* stream);
/* the consumer, in place on what the all-gather delivered: d_all = [world][n_seq][w] counts of `dtype`, rank g's
 * block holding widths[g] <= w documents per record.  Per record: first document (global index) with the maximum
 * count, that count, the number of documents sharing it (> 1 = ambiguous, main.nf:417-436); d_totals (optional,
 * uint64 [sum widths], caller-zeroed) accumulates per-document totals.  Any output pointer may be NULL. */
int xs_sharded_reduce_device(const void* d_all, uint64_t n_seq, int dtype, int device, uint32_t world, uint32_t w,
                             const uint32_t* widths, uint32_t* d_best, uint32_t* d_best_count, uint32_t* d_n_best,
                             uint64_t* d_totals, void* stream);
/* A classic index whose rows come from a counter-based generator instead of a file (measurement support for
 * BASELINE config 5: D = 10 000 x S = 96 000 000 is 120 GB; each rank generates its column shard straight into HBM,
 * the oracle regenerates the rows its parity sample touches).  32 documents of row r, word v:
 * m = mix64(seed ^ r * 0x9E3779B97F4A7C15 ^ v * 0xD1B54A32D192ED03), bits = hi32(m) & lo32(m) (splitmix64 finaliser;
 * fill 0.25).  doc_begin must be a multiple of 32.  Queries go through the same kernels as a loaded file. */
int xs_cobs_create_synthetic(int device, uint32_t n_docs, uint32_t doc_begin, uint32_t doc_end, uint64_t sig_size,
                             uint32_t term_size, uint32_t num_hashes, uint64_t seed, xs_cobs** out);

/* ---- single stages (host buffers; used by the parity tests to pin each kernel alone) ----- */
/* 2-bit packing: packed[w] holds bases [32w, 32w+32), base j in bits [2j, 2j+1], A=0 C=1 G=2 T=3;
 * invalid[w] bit j = 1 when the byte is not one of upper-case ACGT.  n_words = n_bases/32 + 1. */
int xs_pack_2bit(const uint8_t* bases, uint64_t n_bases, int device, uint64_t* packed, uint32_t* invalid);
/* canonical k-mer of every window p = 0..n_bases-k of one sequence: codes[p] = 2-bit code of
 * min(kmer, revcomp), first base in the most significant position; valid[p] = 0 when the
 * window touches a non-ACGT byte. */
int xs_canonical_kmers(const uint8_t* bases, uint64_t n_bases, uint32_t k, int device,
                       uint64_t* codes, uint8_t* valid);
/* row ids of one sequence's sampled windows: rows[(w*h + j)*n_pages + page], valid[w] */
int xs_cobs_rows(xs_cobs* ix, const uint8_t* bases, uint64_t n_bases, uint32_t step,
                 uint64_t* rows, uint8_t* valid);
/* the same without an index handle (classic geometry given explicitly): rows[w*h + j].  Used by the
 * synthetic-workload generator to plant k-mers into an index before it is written to disk. */
int xs_kmer_rows(const uint8_t* bases, uint64_t n_bases, uint32_t k, uint32_t canonicalize, uint32_t num_hashes,
                 uint64_t sig_size, uint32_t step, int device, uint64_t* rows, uint8_t* valid);
/* XXH3-64 of the Bloom term of every sampled window of one sequence */
int xs_bloom_hashes(xs_bloom* bf, const uint8_t* bases, uint64_t n_bases, uint32_t step, uint64_t* hashes);

#ifdef __cplusplus
}
#endif
#endif /* XSPECT_B200_H */
