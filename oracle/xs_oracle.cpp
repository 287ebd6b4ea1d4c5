// xs_oracle.cpp — CPU restatement of the XspecT k-mer scoring hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under xspect2_b200/ (the product) may
// import, link or execute this file.  It is used by tests/, by
// __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
// legs as the checker and the reported host-core baseline.
//
// PARITY STATUS: "parity unpinned" for the third-party pieces.  The arithmetic
// of this path lives in un-vendored dependencies of the reference
// (pyproject.toml:8-25, all unpinned, no lock file):
//   * cobs-reloaded (module cobs_index; fork of bingmann/cobs)  — COBS classic
//     / compact index file layout and ClassicSearch::search
//   * rbloom (KenanHanke/rbloom)                                — Bloom file + LCG probes
//   * xxhash (python-xxhash; XXH64 / XXH3-64)                   — pinned HERE against
//     python-xxhash 3.7.0 known answers (tests/golden/xxhash_kat.json)
// None of the first two is installable offline, so their published algorithms
// are restated from SURVEY.md Appendix A; every assumption is one switch
// (see XS_POLICY_* and the notes at each function).  What IS pinned: XXH64 /
// XXH3 bit-exactness, and all in-tree reference semantics (call sites below).
//
// Reference call sites this follows (relative to /root/reference):
//   src/xspect/models/probabilistic_filter_model.py:196-235   calculate_hits → cobs Search.search(str, step)
//   src/xspect/models/probabilistic_filter_model.py:411-469   _count_kmers  = ceil((len-k+1)/step)
//   src/xspect/models/probabilistic_single_filter_model.py:98-125,161-180  Bloom hits, canonical min(kmer, revcomp)
//   src/xspect/models/probabilistic_filter_mlst_model.py:236-303,362-380    MLST chunk / >50 epilogue (python side in oracle.py)
//
// Build: g++ -O2 -pthread -shared -fPIC (see oracle/Makefile).

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

#include <atomic>
#include <thread>

// Parallel loop over sequences with std::thread (libgomp is not reliably linkable in this image).
// The reference itself is single-threaded Python; threads here only serve the reported
// host-core baseline.
template <typename F>
static void parallel_for(uint64_t n, int n_threads, F f) {
    if (n_threads <= 1 || n < 2) { for (uint64_t i = 0; i < n; ++i) f(i); return; }
    std::atomic<uint64_t> next(0);
    const uint64_t grain = 64;
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t)
        th.emplace_back([&]() {
            for (;;) {
                uint64_t b = next.fetch_add(grain);
                if (b >= n) break;
                uint64_t e = std::min(n, b + grain);
                for (uint64_t i = b; i < e; ++i) f(i);
            }
        });
    for (auto& t : th) t.join();
}

extern "C" {

// ----------------------------------------------------------------------------
// XXH64 (xxHash, Yann Collet, BSD-2).  Restated from the published algorithm;
// pinned against python-xxhash (tests/test_oracle_hash.py).
// ----------------------------------------------------------------------------
static const uint64_t P64_1 = 0x9E3779B185EBCA87ULL;
static const uint64_t P64_2 = 0xC2B2AE3D27D4EB4FULL;
static const uint64_t P64_3 = 0x165667B19E3779F9ULL;
static const uint64_t P64_4 = 0x85EBCA77C2B2AE63ULL;
static const uint64_t P64_5 = 0x27D4EB2F165667C5ULL;

static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }

static inline uint64_t xxh64_round(uint64_t acc, uint64_t in) {
    acc += in * P64_2;
    acc = rotl64(acc, 31);
    return acc * P64_1;
}
static inline uint64_t xxh64_merge(uint64_t acc, uint64_t v) {
    acc ^= xxh64_round(0, v);
    return acc * P64_1 + P64_4;
}
static inline uint64_t xxh64_avalanche(uint64_t h) {
    h ^= h >> 33; h *= P64_2;
    h ^= h >> 29; h *= P64_3;
    h ^= h >> 32;
    return h;
}

uint64_t xso_xxh64(const uint8_t* p, uint64_t len, uint64_t seed) {
    const uint8_t* end = p + len;
    uint64_t h;
    if (len >= 32) {
        uint64_t v1 = seed + P64_1 + P64_2, v2 = seed + P64_2, v3 = seed, v4 = seed - P64_1;
        do {
            v1 = xxh64_round(v1, rd64(p));      v2 = xxh64_round(v2, rd64(p + 8));
            v3 = xxh64_round(v3, rd64(p + 16)); v4 = xxh64_round(v4, rd64(p + 24));
            p += 32;
        } while (p + 32 <= end);
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h = xxh64_merge(h, v1); h = xxh64_merge(h, v2);
        h = xxh64_merge(h, v3); h = xxh64_merge(h, v4);
    } else {
        h = seed + P64_5;
    }
    h += len;
    while (p + 8 <= end) {
        h ^= xxh64_round(0, rd64(p));
        h = rotl64(h, 27) * P64_1 + P64_4;
        p += 8;
    }
    if (p + 4 <= end) {
        h ^= (uint64_t)rd32(p) * P64_1;
        h = rotl64(h, 23) * P64_2 + P64_3;
        p += 4;
    }
    while (p < end) {
        h ^= (uint64_t)(*p) * P64_5;
        h = rotl64(h, 11) * P64_1;
        ++p;
    }
    return xxh64_avalanche(h);
}

// ----------------------------------------------------------------------------
// XXH3-64, seed 0, default secret, lengths 0..128 (k-mers are <= 32 bytes).
// This is what `xxhash.xxh3_64_intdigest(str)` computes for the Bloom path
// (probabilistic_single_filter_model.py:88,155-158).
// ----------------------------------------------------------------------------
static const uint8_t XXH3_SECRET[192] = {
    0xb8, 0xfe, 0x6c, 0x39, 0x23, 0xa4, 0x4b, 0xbe, 0x7c, 0x01, 0x81, 0x2c, 0xf7, 0x21, 0xad, 0x1c,
    0xde, 0xd4, 0x6d, 0xe9, 0x83, 0x90, 0x97, 0xdb, 0x72, 0x40, 0xa4, 0xa4, 0xb7, 0xb3, 0x67, 0x1f,
    0xcb, 0x79, 0xe6, 0x4e, 0xcc, 0xc0, 0xe5, 0x78, 0x82, 0x5a, 0xd0, 0x7d, 0xcc, 0xff, 0x72, 0x21,
    0xb8, 0x08, 0x46, 0x74, 0xf7, 0x43, 0x24, 0x8e, 0xe0, 0x35, 0x90, 0xe6, 0x81, 0x3a, 0x26, 0x4c,
    0x3c, 0x28, 0x52, 0xbb, 0x91, 0xc3, 0x00, 0xcb, 0x88, 0xd0, 0x65, 0x8b, 0x1b, 0x53, 0x2e, 0xa3,
    0x71, 0x64, 0x48, 0x97, 0xa2, 0x0d, 0xf9, 0x4e, 0x38, 0x19, 0xef, 0x46, 0xa9, 0xde, 0xac, 0xd8,
    0xa8, 0xfa, 0x76, 0x3f, 0xe3, 0x9c, 0x34, 0x3f, 0xf9, 0xdc, 0xbb, 0xc7, 0xc7, 0x0b, 0x4f, 0x1d,
    0x8a, 0x51, 0xe0, 0x4b, 0xcd, 0xb4, 0x59, 0x31, 0xc8, 0x9f, 0x7e, 0xc9, 0xd9, 0x78, 0x73, 0x64,
    0xea, 0xc5, 0xac, 0x83, 0x34, 0xd3, 0xeb, 0xc3, 0xc5, 0x81, 0xa0, 0xff, 0xfa, 0x13, 0x63, 0xeb,
    0x17, 0x0d, 0xdd, 0x51, 0xb7, 0xf0, 0xda, 0x49, 0xd3, 0x16, 0x55, 0x26, 0x29, 0xd4, 0x68, 0x9e,
    0x2b, 0x16, 0xbe, 0x58, 0x7d, 0x47, 0xa1, 0xfc, 0x8f, 0xf8, 0xb8, 0xd1, 0x7a, 0xd0, 0x31, 0xce,
    0x45, 0xcb, 0x3a, 0x8f, 0x95, 0x16, 0x04, 0x28, 0xaf, 0xd7, 0xfb, 0xca, 0xbb, 0x4b, 0x40, 0x7e,
};
static const uint64_t PRIME_MX1 = 0x165667919E3779F9ULL;
static const uint64_t PRIME_MX2 = 0x9FB21C651E98DF25ULL;
static const uint32_t P32_1 = 0x9E3779B1U, P32_2 = 0x85EBCA77U, P32_3 = 0xC2B2AE3DU;

static inline uint64_t mul128_fold64(uint64_t a, uint64_t b) {
    unsigned __int128 p = (unsigned __int128)a * b;
    return (uint64_t)p ^ (uint64_t)(p >> 64);
}
static inline uint64_t xxh3_avalanche(uint64_t h) {
    h ^= h >> 37; h *= PRIME_MX1; h ^= h >> 32; return h;
}
static inline uint64_t xxh3_mix16(const uint8_t* in, const uint8_t* sec) {
    return mul128_fold64(rd64(in) ^ rd64(sec), rd64(in + 8) ^ rd64(sec + 8));
}
static inline uint64_t bswap64(uint64_t x) { return __builtin_bswap64(x); }

// returns 0 and sets *ok=0 for len > 128 (not needed on this path)
uint64_t xso_xxh3_64(const uint8_t* in, uint64_t len) {
    const uint8_t* s = XXH3_SECRET;
    (void)P32_1; (void)P32_2; (void)P32_3;
    if (len == 0) return xxh64_avalanche(rd64(s + 56) ^ rd64(s + 64));
    if (len <= 3) {
        uint8_t c1 = in[0], c2 = in[len >> 1], c3 = in[len - 1];
        uint32_t combined = ((uint32_t)c1 << 16) | ((uint32_t)c2 << 24) | (uint32_t)c3 | ((uint32_t)len << 8);
        uint64_t bitflip = (uint64_t)(rd32(s) ^ rd32(s + 4));
        return xxh64_avalanche((uint64_t)combined ^ bitflip);
    }
    if (len <= 8) {
        uint32_t in1 = rd32(in), in2 = rd32(in + len - 4);
        uint64_t bitflip = rd64(s + 8) ^ rd64(s + 16);
        uint64_t h = ((uint64_t)in2 + ((uint64_t)in1 << 32)) ^ bitflip;
        h ^= rotl64(h, 49) ^ rotl64(h, 24);
        h *= PRIME_MX2;
        h ^= (h >> 35) + len;
        h *= PRIME_MX2;
        return h ^ (h >> 28);
    }
    if (len <= 16) {
        uint64_t bf1 = rd64(s + 24) ^ rd64(s + 32), bf2 = rd64(s + 40) ^ rd64(s + 48);
        uint64_t lo = rd64(in) ^ bf1, hi = rd64(in + len - 8) ^ bf2;
        uint64_t acc = len + bswap64(lo) + hi + mul128_fold64(lo, hi);
        return xxh3_avalanche(acc);
    }
    if (len <= 128) {
        uint64_t acc = len * P64_1;
        if (len > 32) {
            if (len > 64) {
                if (len > 96) {
                    acc += xxh3_mix16(in + 48, s + 96);
                    acc += xxh3_mix16(in + len - 64, s + 112);
                }
                acc += xxh3_mix16(in + 32, s + 64);
                acc += xxh3_mix16(in + len - 48, s + 80);
            }
            acc += xxh3_mix16(in + 16, s + 32);
            acc += xxh3_mix16(in + len - 32, s + 48);
        }
        acc += xxh3_mix16(in, s);
        acc += xxh3_mix16(in + len - 16, s + 16);
        return xxh3_avalanche(acc);
    }
    return 0;  // unsupported length on this path
}

// ----------------------------------------------------------------------------
// Canonical k-mers.
// ----------------------------------------------------------------------------
// COBS (third-party, [UNVERIFIED-3P] SURVEY A.2.2/A.2.5): term = lexicographic
// min of the k-mer and its reverse complement over upper-case ACGT when the
// index header has canonicalize == 1.  Windows touching any other byte:
//   XS_POLICY_SKIP    (default)  window contributes nothing
//   XS_POLICY_LITERAL            complement of a non-ACGT byte is 0x00; hash the
//                                lexicographic min of literal bytes / mapped revcomp
enum { XS_POLICY_SKIP = 0, XS_POLICY_LITERAL = 1 };

static inline uint8_t cobs_comp(uint8_t c) {
    switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return 0; }
}

// writes the term to hash into out[k]; returns 1 if the window is to be hashed, 0 if skipped
int xso_cobs_term(const uint8_t* w, uint32_t k, int canonicalize, int policy, uint8_t* out) {
    if (!canonicalize) { memcpy(out, w, k); return 1; }
    uint8_t rc[256];
    int good = 1;
    for (uint32_t i = 0; i < k; ++i) {
        uint8_t c = cobs_comp(w[k - 1 - i]);
        if (c == 0) good = 0;
        rc[i] = c;
    }
    if (!good && policy == XS_POLICY_SKIP) return 0;
    if (memcmp(w, rc, k) <= 0) memcpy(out, w, k); else memcpy(out, rc, k);
    return 1;
}

// Bloom path (in-tree, exact: probabilistic_single_filter_model.py:175-180):
//   minimizer = min(kmer, str(kmer.reverse_complement()))
// Python min(a, b) returns b only if b < a; Seq vs str compares raw bytes; Biopython's
// DNA complement table is IUPAC-aware, case-preserving, U->A, everything else identity.
static uint8_t BIO_COMP[256];
static int bio_comp_ready = 0;
static void bio_comp_init() {
    if (bio_comp_ready) return;
    for (int i = 0; i < 256; ++i) BIO_COMP[i] = (uint8_t)i;
    const char* from = "ACGTMRWSYKVHDBXN";
    const char* to   = "TGCAKYWSRMBDHVXN";
    for (int i = 0; from[i]; ++i) {
        BIO_COMP[(uint8_t)from[i]] = (uint8_t)to[i];
        BIO_COMP[(uint8_t)(from[i] + 32)] = (uint8_t)(to[i] + 32);
    }
    BIO_COMP[(uint8_t)'U'] = 'A';
    BIO_COMP[(uint8_t)'u'] = 'a';
    bio_comp_ready = 1;
}
void xso_bio_complement_table(uint8_t* out256) { bio_comp_init(); memcpy(out256, BIO_COMP, 256); }

void xso_bloom_term(const uint8_t* w, uint32_t k, uint8_t* out) {
    bio_comp_init();
    uint8_t rc[256];
    for (uint32_t i = 0; i < k; ++i) rc[i] = BIO_COMP[w[k - 1 - i]];
    if (memcmp(rc, w, k) < 0) memcpy(out, rc, k); else memcpy(out, w, k);
}

// ----------------------------------------------------------------------------
// COBS query.  An index (classic or compact) is described as n_pages pages;
// page i is a row-major [sig_size[i] x page_bytes] byte matrix at data + page_off[i];
// document d lives in page d / (8*page_bytes), byte (d % (8*page_bytes)) / 8, bit d % 8
// (LSB first).  A classic index is the one-page case with page_bytes = ceil(D/8).
// [UNVERIFIED-3P] SURVEY A.1-A.3.
//   counts[d] += AND_j row_j bit d, r_j = XXH64(term, k, seed=j) % sig_size[page]
// sampled positions p = 0, step, 2*step, ... <= L-k   (A.2.1; _count_kmers)
// ----------------------------------------------------------------------------
typedef struct {
    const uint8_t* data;
    uint32_t n_pages;
    uint64_t page_bytes;
    const uint64_t* sig_size;   // [n_pages]
    const uint64_t* page_off;   // [n_pages] byte offset of page i within data
    uint32_t n_docs;
    uint32_t num_hashes;
    uint32_t k;
    int canonicalize;
    int policy;
} xso_cobs_t;

void xso_cobs_query_one(const xso_cobs_t* ix, const uint8_t* seq, uint64_t len, uint32_t step, uint32_t* counts) {
    for (uint32_t d = 0; d < ix->n_docs; ++d) counts[d] = 0;
    if (len < ix->k || step == 0) return;
    uint8_t term[256];
    std::vector<uint8_t> acc(ix->page_bytes);
    uint64_t hashes[64];
    for (uint64_t p = 0; p + ix->k <= len; p += step) {
        if (!xso_cobs_term(seq + p, ix->k, ix->canonicalize, ix->policy, term)) continue;
        for (uint32_t j = 0; j < ix->num_hashes; ++j) hashes[j] = xso_xxh64(term, ix->k, j);
        for (uint32_t pg = 0; pg < ix->n_pages; ++pg) {
            const uint8_t* base = ix->data + ix->page_off[pg];
            const uint8_t* r0 = base + (hashes[0] % ix->sig_size[pg]) * ix->page_bytes;
            memcpy(acc.data(), r0, ix->page_bytes);
            for (uint32_t j = 1; j < ix->num_hashes; ++j) {
                const uint8_t* r = base + (hashes[j] % ix->sig_size[pg]) * ix->page_bytes;
                for (uint64_t b = 0; b < ix->page_bytes; ++b) acc[b] &= r[b];
            }
            uint64_t d0 = (uint64_t)pg * 8 * ix->page_bytes;
            for (uint64_t b = 0; b < ix->page_bytes; ++b) {
                uint8_t v = acc[b];
                while (v) {
                    int bit = __builtin_ctz(v);
                    v &= (uint8_t)(v - 1);
                    uint64_t d = d0 + b * 8 + bit;
                    if (d < ix->n_docs) counts[d]++;
                }
            }
        }
    }
}

// batch: counts is [n_seq x n_docs] uint32, threads over sequences.
void xso_cobs_query_batch(const xso_cobs_t* ix, const uint8_t* bases, const uint64_t* seq_begin,
                          const uint64_t* seq_end, uint64_t n_seq, uint32_t step, uint32_t* counts,
                          int n_threads) {
    parallel_for(n_seq, n_threads, [&](uint64_t i) {
        xso_cobs_query_one(ix, bases + seq_begin[i], seq_end[i] - seq_begin[i], step,
                           counts + i * ix->n_docs);
    });
}

// The same query against the synthetic index of BASELINE config 5 (bench.py's column-sharded leg): a classic index
// of n_docs documents x sig_size rows that exists nowhere as a file; 32 documents of row r, word v, are
//   m = splitmix64_finalise(seed ^ r * 0x9E3779B97F4A7C15 ^ v * 0xD1B54A32D192ED03);  bits = hi32(m) & lo32(m)
// (include/xspect_b200.h, xs_cobs_create_synthetic).  Rows are regenerated on the fly for every probe.
static inline uint64_t synth_mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL; x ^= x >> 27; x *= 0x94D049BB133111EBULL; x ^= x >> 31; return x;
}
static inline uint32_t synth_row_word(uint64_t seed, uint64_t row, uint32_t word) {
    const uint64_t w = synth_mix64(seed ^ (row * 0x9E3779B97F4A7C15ULL) ^ ((uint64_t)word * 0xD1B54A32D192ED03ULL));
    return (uint32_t)(w >> 32) & (uint32_t)w;
}
void xso_synth_row(uint64_t seed, uint64_t row, uint32_t n_docs, uint32_t* words_out) {
    const uint32_t nw = (n_docs + 31) / 32;
    for (uint32_t v = 0; v < nw; ++v) {
        uint32_t x = synth_row_word(seed, row, v);
        if (v * 32 + 32 > n_docs) x &= (1u << (n_docs - v * 32)) - 1u;
        words_out[v] = x;
    }
}
void xso_cobs_query_batch_synth(uint64_t seed, uint32_t n_docs, uint64_t sig_size, uint32_t num_hashes, uint32_t k,
                                int canonicalize, int policy, const uint8_t* bases, const uint64_t* seq_begin,
                                const uint64_t* seq_end, uint64_t n_seq, uint32_t step, uint32_t* counts, int n_threads) {
    const uint32_t nw = (n_docs + 31) / 32;
    parallel_for(n_seq, n_threads, [&](uint64_t i) {
        uint32_t* out = counts + i * n_docs;
        for (uint32_t d = 0; d < n_docs; ++d) out[d] = 0;
        const uint8_t* seq = bases + seq_begin[i];
        const uint64_t len = seq_end[i] - seq_begin[i];
        if (len < k || step == 0) return;
        uint8_t term[256];
        std::vector<uint32_t> acc(nw), row(nw);
        for (uint64_t p = 0; p + k <= len; p += step) {
            if (!xso_cobs_term(seq + p, k, canonicalize, policy, term)) continue;
            for (uint32_t j = 0; j < num_hashes; ++j) {
                xso_synth_row(seed, xso_xxh64(term, k, j) % sig_size, n_docs, j == 0 ? acc.data() : row.data());
                if (j) for (uint32_t v = 0; v < nw; ++v) acc[v] &= row[v];
            }
            for (uint32_t v = 0; v < nw; ++v) {
                uint32_t x = acc[v];
                while (x) { const int b = __builtin_ctz(x); x &= x - 1; out[v * 32 + b]++; }
            }
        }
    });
}

// row ids for every sampled window of one sequence: rows[(w*num_hashes + j)*n_pages + pg];
// valid[w] = 0 for skipped windows.  Used to pin the hash stage of the CUDA path on its own.
void xso_cobs_rows(const xso_cobs_t* ix, const uint8_t* seq, uint64_t len, uint32_t step,
                   uint64_t* rows, uint8_t* valid) {
    uint8_t term[256];
    uint64_t w = 0;
    for (uint64_t p = 0; p + ix->k <= len; p += step, ++w) {
        int ok = xso_cobs_term(seq + p, ix->k, ix->canonicalize, ix->policy, term);
        valid[w] = (uint8_t)ok;
        for (uint32_t j = 0; j < ix->num_hashes; ++j) {
            uint64_t h = ok ? xso_xxh64(term, ix->k, j) : 0;
            for (uint32_t pg = 0; pg < ix->n_pages; ++pg)
                rows[(w * ix->num_hashes + j) * ix->n_pages + pg] = ok ? h % ix->sig_size[pg] : 0;
        }
    }
}

// Construction (test fixtures only): OR document d's k-mers into the index.  Mirrors COBS
// classic construction: every window of the document, canonicalised with the same rule.
void xso_cobs_insert(xso_cobs_t* ix, uint8_t* data_rw, uint32_t doc, const uint8_t* seq, uint64_t len) {
    uint8_t term[256];
    uint32_t pg = (uint32_t)(doc / (8 * ix->page_bytes));
    uint64_t within = doc % (8 * ix->page_bytes);
    for (uint64_t p = 0; p + ix->k <= len; ++p) {
        if (!xso_cobs_term(seq + p, ix->k, ix->canonicalize, ix->policy, term)) continue;
        for (uint32_t j = 0; j < ix->num_hashes; ++j) {
            uint64_t r = xso_xxh64(term, ix->k, j) % ix->sig_size[pg];
            data_rw[ix->page_off[pg] + r * ix->page_bytes + within / 8] |= (uint8_t)(1u << (within % 8));
        }
    }
}

// Result order of cobs Search.search ([UNVERIFIED-3P] SURVEY A.2.5c): all documents,
// std::partial_sort over iota indices with comparator score[a] > score[b], num_results = D.
void xso_cobs_result_order(const uint32_t* scores, uint32_t n_docs, uint32_t* order) {
    std::vector<uint32_t> idx(n_docs);
    std::iota(idx.begin(), idx.end(), 0u);
    std::partial_sort(idx.begin(), idx.end(), idx.end(),
                      [&](uint32_t a, uint32_t b) { return scores[a] > scores[b]; });
    for (uint32_t i = 0; i < n_docs; ++i) order[i] = idx[i];
}

// ----------------------------------------------------------------------------
// rbloom membership ([UNVERIFIED-3P] SURVEY A.4): file = u64 LE k ‖ bit array;
// state(u128) = xxh3_64(term); k times: state = state*M + 1; idx = (u64)(state >> 32) % nbits;
// member iff all bits set (bit idx%8 of byte idx/8).
// ----------------------------------------------------------------------------
typedef struct {
    const uint8_t* bits;
    uint64_t n_bits;
    uint64_t k_hashes;
    uint32_t k;   // k-mer length
} xso_bloom_t;

static const unsigned __int128 LCG_M =
    ((unsigned __int128)0x2360ED051FC65DA4ULL << 64) | (unsigned __int128)0x4385DF649FCB5CEDULL;

int xso_bloom_contains_hash(const xso_bloom_t* bf, uint64_t h0) {
    unsigned __int128 st = (unsigned __int128)h0;
    for (uint64_t i = 0; i < bf->k_hashes; ++i) {
        st = st * LCG_M + 1;
        uint64_t idx = (uint64_t)(st >> 32) % bf->n_bits;
        if (!(bf->bits[idx >> 3] & (1u << (idx & 7)))) return 0;
    }
    return 1;
}
void xso_bloom_indexes(const xso_bloom_t* bf, uint64_t h0, uint64_t* out) {
    unsigned __int128 st = (unsigned __int128)h0;
    for (uint64_t i = 0; i < bf->k_hashes; ++i) {
        st = st * LCG_M + 1;
        out[i] = (uint64_t)(st >> 32) % bf->n_bits;
    }
}
void xso_bloom_add_hash(const xso_bloom_t* bf, uint8_t* bits_rw, uint64_t h0) {
    unsigned __int128 st = (unsigned __int128)h0;
    for (uint64_t i = 0; i < bf->k_hashes; ++i) {
        st = st * LCG_M + 1;
        uint64_t idx = (uint64_t)(st >> 32) % bf->n_bits;
        bits_rw[idx >> 3] |= (uint8_t)(1u << (idx & 7));
    }
}

uint32_t xso_bloom_hits_one(const xso_bloom_t* bf, const uint8_t* seq, uint64_t len, uint32_t step) {
    if (len < bf->k || step == 0) return 0;
    uint8_t term[256];
    uint32_t hits = 0;
    for (uint64_t p = 0; p + bf->k <= len; p += step) {
        xso_bloom_term(seq + p, bf->k, term);
        hits += (uint32_t)xso_bloom_contains_hash(bf, xso_xxh3_64(term, bf->k));
    }
    return hits;
}
void xso_bloom_hits_batch(const xso_bloom_t* bf, const uint8_t* bases, const uint64_t* seq_begin,
                          const uint64_t* seq_end, uint64_t n_seq, uint32_t step, uint32_t* hits,
                          int n_threads) {
    parallel_for(n_seq, n_threads, [&](uint64_t i) {
        hits[i] = xso_bloom_hits_one(bf, bases + seq_begin[i], seq_end[i] - seq_begin[i], step);
    });
}
// fixture construction: add every window (step 1) like ProbabilisticSingleFilterModel.fit (:89-92)
void xso_bloom_insert(const xso_bloom_t* bf, uint8_t* bits_rw, const uint8_t* seq, uint64_t len) {
    uint8_t term[256];
    for (uint64_t p = 0; p + bf->k <= len; ++p) {
        xso_bloom_term(seq + p, bf->k, term);
        xso_bloom_add_hash(bf, bits_rw, xso_xxh3_64(term, bf->k));
    }
}

int xso_max_threads() {
    unsigned n = std::thread::hardware_concurrency();
    return n ? (int)n : 1;
}

}  // extern "C"
