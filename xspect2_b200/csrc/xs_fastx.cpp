// xs_fastx.cpp — native FASTA / FASTQ ingest for the batched scoring path (host only, no CUDA).
//
// Replaces the record iteration the reference does with Biopython (SeqIO.parse behind
// file_io.get_record_iterator, file_io.py:47-79, consumed record by record at
// probabilistic_filter_model.py:291-310) by one pass that lays all records out as the C ABI's query input:
// one contiguous base buffer + seq_begin/seq_end offsets, plus the record ids (first word of the title line,
// Biopython's `record.id`).  Sequence bytes are copied as they are (case, N, IUPAC preserved); FASTA line
// breaks, '\r' and blanks are removed like Bio.SeqIO.FastaIO does; FASTQ may be wrapped.
#include <algorithm>
#include <cerrno>
#include <cstdint>
#include <cstring>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/xspect_b200.h"

int xs_set_error(int code, const std::string& msg);   // xs_lib.cu

struct Checkpoint { uint64_t off, rec, base, id; };   // a record start: file offset and output cursors there

struct xs_fastx {
    const uint8_t* data = nullptr;
    uint64_t size = 0;
    int fd = -1;
    int format = 0;          // 1 fasta, 2 fastq
    uint64_t n_records = 0, n_bases = 0, n_id_bytes = 0;
    std::vector<Checkpoint> cps;   // every CP_EVERY records (sizing pass) -> segments the fill pass runs in parallel
};
static const uint64_t CP_EVERY = 1 << 15;

namespace {

inline const uint8_t* find_nl(const uint8_t* p, const uint8_t* end) {
    const void* q = memchr(p, '\n', (size_t)(end - p));
    return q ? (const uint8_t*)q : end;
}
inline bool is_space(uint8_t c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

// first whitespace-delimited word of [p, e)
inline void first_word(const uint8_t* p, const uint8_t* e, const uint8_t** wb, const uint8_t** we) {
    while (p < e && is_space(*p)) ++p;
    const uint8_t* q = p;
    while (q < e && !is_space(*q)) ++q;
    *wb = p; *we = q;
}

struct Sink {   // pass 1 counts, pass 2 writes
    uint8_t* bases = nullptr;
    uint64_t* seq_begin = nullptr;
    uint64_t* seq_end = nullptr;
    char* ids = nullptr;
    uint64_t* id_end = nullptr;
    uint64_t n_rec = 0, n_bases = 0, n_id = 0;
    bool write = false;
    std::vector<Checkpoint>* cps = nullptr;
    const uint8_t* file_base = nullptr;
    // filter mode: write the kept records as FASTA (title line as read, sequence wrapped at 60 columns)
    FILE* filter_out = nullptr;
    const uint8_t* keep = nullptr;
    bool keeping = false;
    std::string seq_buf;

    inline void begin_record(const uint8_t* rec_start, const uint8_t* title_b, const uint8_t* title_e) {
        if (cps && (n_rec % CP_EVERY) == 0) cps->push_back({(uint64_t)(rec_start - file_base), n_rec, n_bases, n_id});
        if (filter_out) {
            keeping = keep[n_rec] != 0;
            if (keeping) {
                seq_buf.clear();
                fputc('>', filter_out);
                fwrite(title_b, 1, (size_t)(title_e - title_b), filter_out);
                fputc('\n', filter_out);
            }
        }
        const uint8_t *wb, *we;
        first_word(title_b, title_e, &wb, &we);
        if (write) {
            memcpy(ids + n_id, wb, (size_t)(we - wb));
            seq_begin[n_rec] = n_bases;
        }
        n_id += (uint64_t)(we - wb);
        if (write) id_end[n_rec] = n_id;
    }
    inline void end_record() {
        if (write) seq_end[n_rec] = n_bases;
        if (filter_out && keeping) {
            for (size_t i = 0; i < seq_buf.size(); i += 60) {
                fwrite(seq_buf.data() + i, 1, std::min<size_t>(60, seq_buf.size() - i), filter_out);
                fputc('\n', filter_out);
            }
        }
        ++n_rec;
    }
    // append a FASTA line: drop ' ' and '\r' (Bio.SeqIO.FastaIO: "".join(lines).replace(" ", "").replace("\r", ""))
    inline void append_fasta(const uint8_t* p, const uint8_t* e) {
        if (!memchr(p, ' ', (size_t)(e - p)) && !memchr(p, '\r', (size_t)(e - p))) {
            append_raw(p, e);
            return;
        }
        for (; p < e; ++p)
            if (*p != ' ' && *p != '\r') {
                if (write) bases[n_bases] = *p;
                if (filter_out && keeping) seq_buf.push_back((char)*p);
                ++n_bases;
            }
    }
    inline void append_raw(const uint8_t* p, const uint8_t* e) {
        if (write) memcpy(bases + n_bases, p, (size_t)(e - p));
        if (filter_out && keeping) seq_buf.append((const char*)p, (size_t)(e - p));
        n_bases += (uint64_t)(e - p);
    }
};

int parse_fasta(const uint8_t* p, const uint8_t* end, Sink& sk) {
    bool in_record = false;
    while (p < end) {
        const uint8_t* nl = find_nl(p, end);
        if (*p == '>') {
            if (in_record) sk.end_record();
            const uint8_t* te = nl;
            if (te > p + 1 && te[-1] == '\r') --te;
            sk.begin_record(p, p + 1, te);
            in_record = true;
        } else if (in_record) {
            sk.append_fasta(p, nl);
        }
        p = nl + 1;
    }
    if (in_record) sk.end_record();
    return XS_OK;
}

inline const uint8_t* rstrip(const uint8_t* p, const uint8_t* e) {
    while (e > p && is_space(e[-1])) --e;
    return e;
}

int parse_fastq(const uint8_t* p, const uint8_t* end, Sink& sk) {
    while (p < end) {
        const uint8_t* nl = find_nl(p, end);
        if (rstrip(p, nl) == p) { p = nl + 1; continue; }           // blank line between records
        if (*p != '@') return xs_set_error(XS_ERR_FORMAT, "Records in Fastq files should start with '@' character");
        sk.begin_record(p, p + 1, rstrip(p + 1, nl));
        p = nl + 1;
        uint64_t seq_len = 0;
        bool plus = false;
        while (p < end) {
            nl = find_nl(p, end);
            if (*p == '+') { plus = true; p = nl + 1; break; }
            const uint8_t* e = rstrip(p, nl);
            if (memchr(p, ' ', (size_t)(e - p)) || memchr(p, '\t', (size_t)(e - p)))
                return xs_set_error(XS_ERR_FORMAT, "Whitespace is not allowed in the sequence.");
            sk.append_raw(p, e);
            seq_len += (uint64_t)(e - p);
            p = nl + 1;
        }
        if (!plus) return xs_set_error(XS_ERR_FORMAT, "End of file without quality information.");
        uint64_t qual_len = 0;
        while (p < end && qual_len < seq_len) {
            nl = find_nl(p, end);
            qual_len += (uint64_t)(rstrip(p, nl) - p);
            p = nl + 1;
        }
        if (seq_len == 0 && p < end) {                                // empty record: one (empty) quality line
            nl = find_nl(p, end);
            if (rstrip(p, nl) == p) p = nl + 1;
        }
        if (qual_len != seq_len)
            return xs_set_error(XS_ERR_FORMAT, "Lengths of sequence and quality values differs (" + std::to_string(seq_len) + " and " +
                                                   std::to_string(qual_len) + ").");
        sk.end_record();
    }
    return XS_OK;
}

int run(const xs_fastx* fx, const uint8_t* p, const uint8_t* end, Sink& sk) {
    return fx->format == 2 ? parse_fastq(p, end, sk) : parse_fasta(p, end, sk);
}

// line [p, nl) starting at p; returns the start of the next line (end when there is none)
inline const uint8_t* next_line(const uint8_t* p, const uint8_t* end) {
    const uint8_t* nl = find_nl(p, end);
    return nl < end ? nl + 1 : end;
}

// Is `p` (the first byte of a line) the start of a FASTQ record?  '@' alone is not enough — it is also a quality
// character — so the 4-line shape is checked: '@' title, sequence, '+' line, quality of the sequence's length, and
// then another '@' (or the end).  Wrapped FASTQ does not pass; the streaming reader then falls back to the two-pass
// reader (xs_fastx_open), which parses serially from the start.
bool fastq_record_at(const uint8_t* p, const uint8_t* end) {
    if (p >= end || *p != '@') return false;
    const uint8_t* l1 = next_line(p, end);
    if (l1 >= end) return false;
    const uint8_t* l2 = next_line(l1, end);
    if (l2 >= end || *l2 != '+') return false;
    const uint8_t* l3 = next_line(l2, end);
    const uint8_t* l4 = next_line(l3, end);
    const uint64_t seq_len = (uint64_t)(rstrip(l1, l2) - l1), qual_len = (uint64_t)(rstrip(l3, l4) - l3);
    if (seq_len != qual_len) return false;
    const uint8_t* q = l4;                                  // blank lines between records are allowed (parse_fastq)
    while (q < end && rstrip(q, find_nl(q, end)) == q) q = next_line(q, end);
    return q >= end || *q == '@';
}

}  // namespace

// ---- streaming access for xs_cobs_classify_file (xs_lib.cu): blocks of the file parsed by all host threads ----
// first record start at or after `off` (a line start): FASTA '>' / FASTQ 4-line shape; fx->size when there is none
uint64_t xs_fastx_sync(const xs_fastx* fx, uint64_t off) {
    const uint8_t* base = fx->data;
    const uint8_t* end = base + fx->size;
    const uint8_t* p = base + std::min(off, fx->size);
    if (p > base && p[-1] != '\n') p = next_line(p, end);     // move to a line start
    while (p < end) {
        if (fx->format == 1 ? *p == '>' : fastq_record_at(p, end)) return (uint64_t)(p - base);
        p = next_line(p, end);
    }
    return fx->size;
}

// Parses [a, b) — both record starts (or the file's end) — with n_thr threads: a counting pass over sub-ranges, a
// prefix sum, then a fill pass that writes every sub-range at its place.  `sizes` receives {records, bases, id bytes};
// with bases == nullptr only the counting pass runs (the caller sizes its staging buffers from it).
// begin / end are offsets into `bases` of this block; max_len = longest record, n_short = records not longer than k.
int xs_fastx_parse_block(const xs_fastx* fx, uint64_t a, uint64_t b, unsigned n_thr, std::vector<uint64_t>& cuts,
                         std::vector<Checkpoint>& cps, uint8_t* bases, uint64_t* seq_begin, uint64_t* seq_end, char* ids,
                         uint64_t* id_end, uint64_t sizes[3]) {
    const bool count_only = bases == nullptr && seq_begin == nullptr;
    if (count_only) {
        cuts.clear();
        cuts.push_back(a);
        n_thr = std::max(1u, n_thr);
        for (unsigned t = 1; t < n_thr; ++t) {
            const uint64_t c = xs_fastx_sync(fx, a + (b - a) * t / n_thr);
            if (c > cuts.back() && c < b) cuts.push_back(c);
        }
        cuts.push_back(b);
        cps.assign(cuts.size(), Checkpoint{0, 0, 0, 0});
    }
    const size_t n_seg = cuts.size() - 1;
    std::atomic<size_t> next(0);
    std::atomic<int> status(XS_OK);
    auto work = [&]() {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= n_seg) break;
            Sink sk;
            if (!count_only) {
                sk.bases = bases; sk.seq_begin = seq_begin; sk.seq_end = seq_end; sk.ids = ids; sk.id_end = id_end; sk.write = true;
                sk.n_rec = cps[i].rec; sk.n_bases = cps[i].base; sk.n_id = cps[i].id;
            }
            const int rc = run(fx, fx->data + cuts[i], fx->data + cuts[i + 1], sk);
            if (rc != XS_OK) { status.store(rc); continue; }
            if (count_only) cps[i + 1] = Checkpoint{cuts[i + 1], sk.n_rec, sk.n_bases, sk.n_id};   // this segment's own counts
        }
    };
    if (n_seg <= 1 || n_thr <= 1) work();
    else {
        std::vector<std::thread> th;
        for (size_t t = 0; t < std::min<size_t>(n_thr, n_seg); ++t) th.emplace_back(work);
        for (auto& t : th) t.join();
    }
    if (status.load() != XS_OK) return status.load();
    if (count_only) {
        cps[0] = Checkpoint{cuts[0], 0, 0, 0};
        for (size_t i = 1; i <= n_seg; ++i) {               // per-segment counts -> running cursors
            cps[i].rec += cps[i - 1].rec; cps[i].base += cps[i - 1].base; cps[i].id += cps[i - 1].id;
        }
        sizes[0] = cps[n_seg].rec; sizes[1] = cps[n_seg].base; sizes[2] = cps[n_seg].id;
    }
    return XS_OK;
}

// Single-pass variant for the streaming classifier: sub-range i of [a, b) is parsed straight into the staging buffer
// at the offset it has in the file block (a record's bases never outnumber its bytes, so the regions cannot
// overlap); begin / end offsets, ids and id ends go to per-thread arrays that the caller compacts (16 bytes per
// record against ~300 bytes of file).  seg_* describe what each sub-range produced.
struct FastxSegOut {
    uint64_t base0 = 0, n_bases = 0;      // region of the staging buffer this segment filled
    uint64_t n_rec = 0, n_id = 0;
    uint64_t* begin = nullptr; uint64_t* end = nullptr; uint64_t* id_end = nullptr; char* ids = nullptr;
    uint64_t cap_rec = 0, cap_id = 0;
    uint64_t max_len = 0, n_short = 0;    // longest record; records not longer than k_short
};

int xs_fastx_parse_block_1pass(const xs_fastx* fx, uint64_t a, uint64_t b, unsigned n_thr, uint8_t* staging,
                               std::vector<FastxSegOut>& segs, uint64_t k_short) {
    std::vector<uint64_t> cuts;
    cuts.push_back(a);
    n_thr = std::max(1u, n_thr);
    for (unsigned t = 1; t < n_thr; ++t) {
        const uint64_t c = xs_fastx_sync(fx, a + (b - a) * t / n_thr);
        if (c > cuts.back() && c < b) cuts.push_back(c);
    }
    cuts.push_back(b);
    const size_t n_seg = cuts.size() - 1;
    if (segs.size() < n_seg) segs.resize(n_seg);
    const uint64_t min_rec = fx->format == 2 ? 6 : 2;      // "@\n\n+\n\n" / ">\n"
    for (size_t i = 0; i < n_seg; ++i) {
        FastxSegOut& so = segs[i];
        const uint64_t bytes = cuts[i + 1] - cuts[i];
        const uint64_t need_rec = bytes / min_rec + 2, need_id = bytes + 1;
        if (so.cap_rec < need_rec) {
            delete[] so.begin; delete[] so.end; delete[] so.id_end;
            so.begin = new uint64_t[need_rec]; so.end = new uint64_t[need_rec]; so.id_end = new uint64_t[need_rec];   // untouched pages cost nothing
            so.cap_rec = need_rec;
        }
        if (so.cap_id < need_id) { delete[] so.ids; so.ids = new char[need_id]; so.cap_id = need_id; }
        so.base0 = cuts[i] - a; so.n_bases = so.n_rec = so.n_id = 0; so.max_len = so.n_short = 0;
    }
    for (size_t i = n_seg; i < segs.size(); ++i) segs[i].n_rec = segs[i].n_bases = segs[i].n_id = segs[i].max_len = segs[i].n_short = 0;
    std::atomic<size_t> next(0);
    std::atomic<int> status(XS_OK);
    auto work = [&]() {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= n_seg) break;
            FastxSegOut& so = segs[i];
            {   // map this sub-range's pages with one call instead of one minor fault per 64 KB (faults from many
                // threads contend on the address space's locks); ignored where the kernel does not know the advice
#ifndef MADV_POPULATE_READ
#define MADV_POPULATE_READ 22
#endif
                const uintptr_t lo = (uintptr_t)(fx->data + cuts[i]) & ~(uintptr_t)4095;
                const uintptr_t hi = (uintptr_t)(fx->data + cuts[i + 1]);
                if (hi > lo) (void)madvise((void*)lo, (size_t)(hi - lo), MADV_POPULATE_READ);
            }
            Sink sk;
            sk.bases = staging; sk.seq_begin = so.begin; sk.seq_end = so.end; sk.ids = so.ids; sk.id_end = so.id_end; sk.write = true;
            sk.n_rec = 0; sk.n_bases = so.base0; sk.n_id = 0;
            const int rc = run(fx, fx->data + cuts[i], fx->data + cuts[i + 1], sk);
            if (rc != XS_OK) { status.store(rc); continue; }
            so.n_rec = sk.n_rec; so.n_bases = sk.n_bases - so.base0; so.n_id = sk.n_id;
            uint64_t mx = 0, ns = 0;
            for (uint64_t r = 0; r < so.n_rec; ++r) {
                const uint64_t len = so.end[r] - so.begin[r];
                mx = len > mx ? len : mx;
                ns += len <= k_short;
            }
            so.max_len = mx; so.n_short = ns;
        }
    };
    if (n_seg <= 1 || n_thr <= 1) work();
    else {
        std::vector<std::thread> th;
        for (size_t t = 0; t < std::min<size_t>(n_thr, n_seg); ++t) th.emplace_back(work);
        for (auto& t : th) t.join();
    }
    return status.load();
}

void xs_fastx_seg_free(std::vector<FastxSegOut>& segs) {
    for (FastxSegOut& so : segs) { delete[] so.begin; delete[] so.end; delete[] so.id_end; delete[] so.ids; so = FastxSegOut(); }
}

extern "C" {

static int fastx_open_impl(const char* path, int format, bool sizing_pass, xs_fastx** out);
int xs_fastx_open(const char* path, int format, xs_fastx** out) { return fastx_open_impl(path, format, true, out); }
}  // extern "C"
// maps the file without the sizing pass (streaming reader)
int xs_fastx_open_stream(const char* path, int format, xs_fastx** out) { return fastx_open_impl(path, format, false, out); }
uint64_t xs_fastx_file_size(const xs_fastx* fx) { return fx->size; }
// true when [a, b) holds only whitespace (FASTQ: nothing but blank lines may precede the first record the streaming
// reader found; anything else means the 4-line shape check skipped real data, e.g. wrapped FASTQ)
bool xs_fastx_blank(const xs_fastx* fx, uint64_t a, uint64_t b) {
    for (uint64_t i = a; i < b && i < fx->size; ++i)
        if (!is_space(fx->data[i])) return false;
    return true;
}
int xs_fastx_format(const xs_fastx* fx) { return fx->format; }
extern "C" {

static int fastx_open_impl(const char* path, int format, bool sizing_pass, xs_fastx** out) {
    if (!path || !out) return xs_set_error(XS_ERR_ARG, "path/out is NULL");
    *out = nullptr;
    if (format != 1 && format != 2) return xs_set_error(XS_ERR_ARG, "format must be 1 (fasta) or 2 (fastq)");
    int fd = open(path, O_RDONLY);
    if (fd < 0) return xs_set_error(XS_ERR_IO, std::string(path) + ": " + strerror(errno));
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return xs_set_error(XS_ERR_IO, std::string(path) + ": " + strerror(errno)); }
    xs_fastx* fx = new xs_fastx();
    fx->fd = fd; fx->size = (uint64_t)st.st_size; fx->format = format;
    if (fx->size) {
        void* m = mmap(nullptr, fx->size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { close(fd); delete fx; return xs_set_error(XS_ERR_IO, std::string(path) + ": mmap failed"); }
        madvise(m, fx->size, MADV_SEQUENTIAL);
        fx->data = (const uint8_t*)m;
    }
    if (!sizing_pass) { *out = fx; return XS_OK; }
    Sink sk;
    sk.cps = &fx->cps; sk.file_base = fx->data;
    int rc = run(fx, fx->data, fx->data + fx->size, sk);
    if (rc != XS_OK) { xs_fastx_close(fx); return rc; }
    fx->n_records = sk.n_rec; fx->n_bases = sk.n_bases; fx->n_id_bytes = sk.n_id;
    *out = fx;
    return XS_OK;
}

int xs_fastx_stats(const xs_fastx* fx, uint64_t* n_records, uint64_t* n_bases, uint64_t* n_id_bytes) {
    if (!fx) return xs_set_error(XS_ERR_ARG, "NULL reader");
    if (n_records) *n_records = fx->n_records;
    if (n_bases) *n_bases = fx->n_bases;
    if (n_id_bytes) *n_id_bytes = fx->n_id_bytes;
    return XS_OK;
}

int xs_fastx_read(const xs_fastx* fx, uint8_t* bases, uint64_t* seq_begin, uint64_t* seq_end, char* ids, uint64_t* id_end) {
    if (!fx) return xs_set_error(XS_ERR_ARG, "NULL reader");
    if ((fx->n_bases && !bases) || (fx->n_records && (!seq_begin || !seq_end || !id_end)) || (fx->n_id_bytes && !ids))
        return xs_set_error(XS_ERR_ARG, "NULL output buffer");
    // fill pass: the segments between checkpoints are independent, run them on a few host threads
    const size_t n_seg = fx->cps.size();
    unsigned hw = std::thread::hardware_concurrency();
    const size_t n_thr = std::max<size_t>(1, std::min<size_t>(std::min<size_t>(hw ? hw : 1, 16), n_seg));
    std::atomic<size_t> next(0);
    std::atomic<int> status(XS_OK);
    auto work = [&]() {
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= n_seg) break;
            const Checkpoint& c = fx->cps[i];
            const uint8_t* b = fx->data + c.off;
            const uint8_t* e = i + 1 < n_seg ? fx->data + fx->cps[i + 1].off : fx->data + fx->size;
            Sink sk;
            sk.bases = bases; sk.seq_begin = seq_begin; sk.seq_end = seq_end; sk.ids = ids; sk.id_end = id_end; sk.write = true;
            sk.n_rec = c.rec; sk.n_bases = c.base; sk.n_id = c.id;
            int rc = run(fx, b, e, sk);
            if (rc != XS_OK) status.store(rc);
        }
    };
    if (n_thr <= 1) work();
    else {
        std::vector<std::thread> th;
        for (size_t t = 0; t < n_thr; ++t) th.emplace_back(work);
        for (auto& t : th) t.join();
    }
    return status.load();
}

int xs_fastx_filter_fasta(const xs_fastx* fx, const uint8_t* keep, const char* out_path) {
    if (!fx || !out_path || (fx->n_records && !keep)) return xs_set_error(XS_ERR_ARG, "NULL argument");
    FILE* f = fopen(out_path, "wb");
    if (!f) return xs_set_error(XS_ERR_IO, std::string(out_path) + ": " + strerror(errno));
    std::vector<char> big(8 << 20);
    setvbuf(f, big.data(), _IOFBF, big.size());
    Sink sk;
    sk.filter_out = f; sk.keep = keep;
    int rc = run(fx, fx->data, fx->data + fx->size, sk);
    if (fclose(f) != 0 && rc == XS_OK) rc = xs_set_error(XS_ERR_IO, std::string(out_path) + ": write failed");
    return rc;
}

int xs_fastx_close(xs_fastx* fx) {
    if (!fx) return XS_OK;
    if (fx->data) munmap((void*)fx->data, fx->size);
    if (fx->fd >= 0) close(fx->fd);
    delete fx;
    return XS_OK;
}

}  // extern "C"
