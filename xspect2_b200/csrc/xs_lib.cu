// xs_lib.cu — host side of libxspect_b200.so (C ABI declared in include/xspect_b200.h).
//
// Loads the reference's own model files (index.cobs_classic, <locus>.cobs_compact,
// filter.bloom) unchanged into HBM and runs the k-mer scoring kernels of xs_kernels.cuh.
// There is no CPU path in this file: every query launches CUDA kernels or fails.
#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <mutex>
#include <numeric>
#include <thread>
#include <string>
#include <vector>

#include <sys/mman.h>

#include <cub/device/device_scan.cuh>
#include <cuda.h>      // CUtensorMap + cuTensorMapEncodeTiled's signature; the entry point is fetched at run time

#include "../../include/xspect_b200.h"
#include "xs_kernels.cuh"

using namespace xs;

// streaming access to the native FASTA / FASTQ reader (xs_fastx.cpp), C++ linkage
struct Checkpoint { uint64_t off, rec, base, id; };
struct xs_fastx;
int xs_fastx_open_stream(const char* path, int format, xs_fastx** out);
uint64_t xs_fastx_file_size(const xs_fastx* fx);
uint64_t xs_fastx_sync(const xs_fastx* fx, uint64_t off);
bool xs_fastx_blank(const xs_fastx* fx, uint64_t a, uint64_t b);
int xs_fastx_format(const xs_fastx* fx);
struct FastxSegOut {
    uint64_t base0 = 0, n_bases = 0;
    uint64_t n_rec = 0, n_id = 0;
    uint64_t* begin = nullptr; uint64_t* end = nullptr; uint64_t* id_end = nullptr; char* ids = nullptr;
    uint64_t cap_rec = 0, cap_id = 0;
    uint64_t max_len = 0, n_short = 0;
};
int xs_fastx_parse_block_1pass(const xs_fastx* fx, uint64_t a, uint64_t b, unsigned n_thr, uint8_t* staging, std::vector<FastxSegOut>& segs,
                               uint64_t k_short);
void xs_fastx_seg_free(std::vector<FastxSegOut>& segs);
int xs_fastx_parse_block(const xs_fastx* fx, uint64_t a, uint64_t b, unsigned n_thr, std::vector<uint64_t>& cuts,
                         std::vector<Checkpoint>& cps, uint8_t* bases, uint64_t* seq_begin, uint64_t* seq_end, char* ids,
                         uint64_t* id_end, uint64_t sizes[3]);

// ----------------------------------------------------------------------------------------
// errors
// ----------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};

static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
int xs_set_error(int code, const std::string& msg) { return fail(code, msg); }   // for xs_fastx.cpp
#define XS_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(e__ == cudaErrorMemoryAllocation ? XS_ERR_NOMEM : XS_ERR_CUDA,             \
                        std::string(#expr) + ": " + cudaGetErrorString(e__));                      \
    } while (0)
#define XS_TRY(expr)                 \
    do {                             \
        int rc__ = (expr);           \
        if (rc__ != XS_OK) return rc__; \
    } while (0)

static int launch_ok(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(XS_ERR_CUDA, std::string(what) + " launch: " + cudaGetErrorString(e));
    return XS_OK;
}

// optional CUDA-event bracketing of the dominant kernel (xs_profile_enable / xs_profile_read)
#include <mutex>
static std::atomic<int> g_profile{0};
static std::mutex g_prof_mu;
struct ProfEvent { cudaEvent_t a, b; int tag; };
static std::vector<ProfEvent> g_prof_events;
enum { PROF_DIRECT = 0, PROF_EMIT = 1, PROF_FETCH = 2, PROF_REDUCE = 3, PROF_TAGS = 4 };
struct KernelTimer {
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t s;
    int tag;
    explicit KernelTimer(cudaStream_t st, int tg = PROF_DIRECT) : s(st), tag(tg) {
        if (!g_profile.load(std::memory_order_relaxed)) return;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) { a = b = nullptr; return; }
        cudaEventRecord(a, s);
    }
    ~KernelTimer() {
        if (!a) return;
        cudaEventRecord(b, s);
        std::lock_guard<std::mutex> lk(g_prof_mu);
        g_prof_events.push_back(ProfEvent{a, b, tag});
    }
};

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

static uint64_t magic_of(uint64_t m) {
    return m <= 1 ? ~0ULL : (uint64_t)((((unsigned __int128)1) << 64) / m);
}

// ----------------------------------------------------------------------------------------
// handles
// ----------------------------------------------------------------------------------------
// Per-handle state of the bucketed path (shared by the COBS and the Bloom handle).
struct BucketState {
    int enabled = 1;                       // large batches go through the bucketed kernels
    uint64_t min_windows = 32ULL << 20;    // measured crossover: ~25 M windows
    uint64_t scratch_bytes = 24ULL << 30;  // upper bound of the scratch of one query
    uint32_t shift = 0;                    // log2 rows / bits per bucket, 0 = automatic
    std::atomic<uint64_t> queries{0};
    std::atomic<uint64_t> budget{0};       // scratch bytes per query actually used; 0 = not asked yet
    // bucketed queries of one handle run one after the other even when they are enqueued on different streams (the
    // host pipeline uses three): their kernels compete for the same L2 slice and LSU path when they overlap
    std::mutex mu;
    cudaEvent_t done = nullptr;
    bool prev = false;
    cudaStream_t side = nullptr;           // hash-ahead on a side stream (XS_BK_HASH_AHEAD=2)
    cudaEvent_t ev_emit = nullptr, ev_hash = nullptr;
    void configure(int on, uint64_t min_w, uint64_t scratch, uint32_t sh) {
        enabled = on ? 1 : 0;
        if (min_w) min_windows = min_w;
        if (scratch) scratch_bytes = scratch;
        budget.store(0);
        shift = sh;
    }
    void destroy() {
        if (done) cudaEventDestroy(done);
        if (side) cudaStreamDestroy(side);
        if (ev_emit) cudaEventDestroy(ev_emit);
        if (ev_hash) cudaEventDestroy(ev_hash);
        done = ev_emit = ev_hash = nullptr; side = nullptr;
    }
};

struct xs_cobs {
    xs_cobs_info_t info{};
    std::vector<PageDesc> pages;
    std::vector<ColBlock> blocks;
    std::string names;  // '\n' separated
    std::string layout; // which candidate reading of the header matched (parse_cobs_header)
    uint8_t* d_data = nullptr;
    PageDesc* d_pages = nullptr;
    ColBlock* d_blocks = nullptr;
    bool narrow = true;
    TensorMap128 tmap{};         // rows of a one-page narrow index as a [sig_size][4 x u32] tensor (tile::gather4 fetch variant)
    bool has_tmap = false;
    bool pages_kernel = false;   // several narrow pages: k_cobs_pages (one pass, windows hashed once) instead of k_cobs_narrow per page
    int n_sm = 148;
    int force_wide = 0;
    bool mid = false;            // one page of 32 / 64 / 128-byte rows: k_cobs_mid (warp walk, lane groups per row) instead of k_cobs_wide
    BucketState bk;
};

struct xs_bloom {
    xs_bloom_info_t info{};
    uint8_t* d_bits = nullptr;
    int n_sm = 148;
    BucketState bk;                   // bucketed probing of large batches (k_bbucket_emit / fetch / reduce)
    // the bucketed kernels make all k probes of a window, k_bloom stops at the first zero bit: bucketed wins when at
    // least about a third of the windows are members (measured crossover); decided on the device from a sample
    uint32_t bucket_member_pct = 35;  // 0 = always bucketed
};

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

struct StreamSlot {
    uint8_t* h = nullptr; size_t h_cap = 0;          // pinned staging: bases | begin | end
    uint8_t* d = nullptr; size_t d_cap = 0;          // device: bases | begin | end | counts | best | cnt | nb
    uint32_t* h_res = nullptr; size_t r_cap = 0;     // pinned results: best | cnt | nb
    cudaStream_t s = nullptr; cudaEvent_t done = nullptr;
    uint64_t rec0 = 0, n_rec = 0; bool busy = false;
};
static const int STREAM_NS = 3;
static std::mutex g_stream_mu;
static StreamSlot g_stream_slot[STREAM_NS];
static int g_stream_dev = -1;
static void stream_slots_release() {       // caller holds g_stream_mu (or is single-threaded)
    if (g_stream_dev < 0) return;
    DeviceGuard guard(g_stream_dev);
    for (StreamSlot& sl : g_stream_slot) {
        if (sl.s) cudaStreamSynchronize(sl.s);
        if (sl.h) cudaFreeHost(sl.h);
        if (sl.h_res) cudaFreeHost(sl.h_res);
        if (sl.d) cudaFree(sl.d);
        if (sl.s) cudaStreamDestroy(sl.s);
        if (sl.done) cudaEventDestroy(sl.done);
        sl = StreamSlot();
    }
    g_stream_dev = -1;
}

// growable buffer without value-initialisation (std::vector::resize would touch every new page once more)
template <typename T>
struct RawBuf {
    T* p = nullptr; size_t n = 0, cap = 0;
    ~RawBuf() { free(p); }
    bool grow(size_t want) {
        if (want <= cap) { n = want; return true; }
        size_t c = std::max<size_t>(want, cap + cap / 2 + 1024);
        T* q = (T*)realloc(p, c * sizeof(T));
        if (!q) return false;
        p = q; cap = c; n = want;
        return true;
    }
    bool reserve(size_t c) {
        if (c <= cap) return true;
        T* q = (T*)realloc(p, c * sizeof(T));
        if (!q) return false;
        p = q; cap = c;
        // large result arrays are first touched by the copies that fill them: ask for huge pages so that costs a few
        // hundred faults instead of tens of thousands (a hint; ignored where transparent huge pages are off)
        const uintptr_t lo = ((uintptr_t)p + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1);
        const uintptr_t hi = ((uintptr_t)p + c * sizeof(T)) & ~(uintptr_t)((2u << 20) - 1);
        if (hi > lo) (void)madvise((void*)lo, hi - lo, MADV_HUGEPAGE);
        return true;
    }
    T* data() { return p; }
    const T* data() const { return p; }
    size_t size() const { return n; }
};

static std::atomic<int> g_home_device{-1};

static int device_setup(int device, int* n_sm) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(XS_ERR_CUDA, std::string("no CUDA device (there is no CPU fallback): ") +
                                     (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    if (device < 0 || device >= n) return fail(XS_ERR_ARG, "device index out of range");
    XS_CUDA(cudaSetDevice(device));
    { int none = -1; g_home_device.compare_exchange_strong(none, device); }
    XS_CUDA(cudaDeviceGetAttribute(n_sm, cudaDevAttrMultiProcessorCount, device));
    // random 16-byte row gathers: fetch single 32-byte sectors from HBM, not 64/128-byte groups
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
    // keep stream-ordered workspace memory cached between queries
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = ~0ULL;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    return XS_OK;
}

// ----------------------------------------------------------------------------------------
// file parsing
// ----------------------------------------------------------------------------------------
struct CobsFile {
    int kind = 0;
    uint32_t k = 0, canonicalize = 0, n_docs = 0, n_pages = 0;
    uint64_t page_bytes = 0, num_hashes = 0;
    std::vector<uint64_t> sig;
    std::string names;
    uint64_t data_off = 0, file_size = 0;
};

template <typename T>
static bool rd(FILE* f, T* v) { return fread(v, sizeof(T), 1, f) == 1; }

// One candidate reading of the header fields that follow "COBS:<MAGIC>".  The layout of cobs-reloaded's headers is
// restated from its published source ([UNVERIFIED-3P]); the documented field order is
// tried first and a few neighbouring orders / widths after it (Appendix A.1.3).  A candidate is accepted only when
// every structural check holds: version 1, plausible field values, the end magic right after the document list (and
// the compact padding), and the size identity  file_size - data_offset == sum(signature_size x row bytes).
struct HeaderLayout {
    const char* name;
    bool canon_u32;      // canonicalize stored as uint32 instead of uint8
    bool hashes_first;   // classic: num_hashes before signature_size; compact: per page {num_hashes, signature_size}
    bool hashes_u32;     // classic: num_hashes stored as uint32
    bool docs_first;     // compact: num_documents before num_pages
};
static const HeaderLayout kHeaderLayouts[] = {
    {"cobs v1 (documented)", false, false, false, false},
    {"num_hashes before signature_size", false, true, false, false},
    {"32-bit num_hashes", false, false, true, false},
    {"32-bit canonicalize", true, false, false, false},
    {"documents before pages", false, false, false, true},
};

static int parse_cobs_header_as(FILE* f, const char* path, const HeaderLayout& L, bool classic, CobsFile& cf) {
    cf = CobsFile();
    if (fseek(f, 18, SEEK_SET) != 0) return fail(XS_ERR_IO, "seek failed");
    uint32_t version = 0;
    if (!rd(f, &version) || !rd(f, &cf.k)) return fail(XS_ERR_FORMAT, "truncated COBS header");
    if (L.canon_u32) { uint32_t c = 0; if (!rd(f, &c)) return fail(XS_ERR_FORMAT, "truncated COBS header"); cf.canonicalize = c; }
    else { uint8_t c = 0; if (!rd(f, &c)) return fail(XS_ERR_FORMAT, "truncated COBS header"); cf.canonicalize = c; }
    if (version != 1) return fail(XS_ERR_FORMAT, "unsupported COBS index version " + std::to_string(version));
    if (cf.canonicalize > 1) return fail(XS_ERR_FORMAT, "canonicalize flag is not 0/1");
    uint64_t nh = 0;
    if (classic) {
        uint64_t sig = 0;
        uint32_t nh32 = 0;
        if (!rd(f, &cf.n_docs)) return fail(XS_ERR_FORMAT, "truncated COBS header");
        bool ok = L.hashes_first ? (rd(f, &nh) && rd(f, &sig))
                                 : (rd(f, &sig) && (L.hashes_u32 ? rd(f, &nh32) : rd(f, &nh)));
        if (!ok) return fail(XS_ERR_FORMAT, "truncated COBS header");
        if (L.hashes_u32) nh = nh32;
        cf.kind = XS_COBS_CLASSIC; cf.n_pages = 1; cf.sig.push_back(sig);
        cf.page_bytes = ((uint64_t)cf.n_docs + 7) / 8;
    } else {
        bool ok = L.docs_first ? (rd(f, &cf.n_docs) && rd(f, &cf.n_pages)) : (rd(f, &cf.n_pages) && rd(f, &cf.n_docs));
        if (!ok || !rd(f, &cf.page_bytes)) return fail(XS_ERR_FORMAT, "truncated COBS header");
        cf.kind = XS_COBS_COMPACT;
        if (cf.n_pages == 0 || cf.n_pages > (1u << 24)) return fail(XS_ERR_FORMAT, "implausible compact page count");
        if (cf.page_bytes == 0 || cf.page_bytes > (1u << 20)) return fail(XS_ERR_FORMAT, "implausible compact page size");
        for (uint32_t i = 0; i < cf.n_pages; ++i) {
            uint64_t s = 0, h = 0;
            if (!(L.hashes_first ? (rd(f, &h) && rd(f, &s)) : (rd(f, &s) && rd(f, &h)))) return fail(XS_ERR_FORMAT, "truncated COBS header");
            if (i == 0) nh = h;
            else if (h != nh) return fail(XS_ERR_UNSUPPORTED, "compact pages with differing num_hashes");
            cf.sig.push_back(s);
        }
    }
    cf.num_hashes = nh;
    if (cf.k == 0) return fail(XS_ERR_FORMAT, "term_size 0");
    if (cf.k > 32) return fail(XS_ERR_UNSUPPORTED, "term_size > 32 is not supported");
    if (cf.num_hashes == 0 || cf.num_hashes > 64) return fail(XS_ERR_UNSUPPORTED, "num_hashes must be in 1..64");
    if (cf.n_docs == 0 || cf.n_docs > (1u << 28)) return fail(XS_ERR_FORMAT, "implausible document count");
    for (uint64_t sg : cf.sig)
        if (sg == 0 || sg > (1ULL << 62)) return fail(XS_ERR_FORMAT, "implausible signature_size");
    for (uint32_t i = 0; i < cf.n_docs; ++i) {
        int c;
        size_t len = 0;
        while ((c = fgetc(f)) != EOF && c != '\n') { cf.names.push_back((char)c); if (++len > 4096) return fail(XS_ERR_FORMAT, "document name too long"); }
        if (c == EOF) return fail(XS_ERR_FORMAT, "truncated COBS document list");
        cf.names.push_back('\n');
    }
    const long pos = ftell(f);
    if (fseek(f, 0, SEEK_END) != 0) return fail(XS_ERR_IO, "seek failed");
    cf.file_size = (uint64_t)ftell(f);
    unsigned __int128 need = 0;
    for (uint64_t sg : cf.sig) need += (unsigned __int128)sg * cf.page_bytes;
    // compact: zero padding so that the data starts on a multiple of page_size.  cobs computes it as
    // page_size - ((pos + 13) % page_size); whether a remainder of 0 pads nothing or a whole page is not checkable
    // offline, so both are accepted — the end magic and the size identity decide.
    uint64_t pads[2] = {0, 0};
    int n_pads = 1;
    if (!classic) {
        const uint64_t r = (cf.page_bytes - (((uint64_t)pos + 13) % cf.page_bytes)) % cf.page_bytes;
        pads[0] = r;
        if (r == 0) { pads[1] = cf.page_bytes; n_pads = 2; }
    }
    for (int i = 0; i < n_pads; ++i) {
        char endm[13];
        if (fseek(f, pos + (long)pads[i], SEEK_SET) != 0 || fread(endm, 1, 13, f) != 13) continue;
        if (memcmp(endm, classic ? "CLASSIC_INDEX" : "COMPACT_INDEX", 13) != 0) continue;
        const uint64_t off = (uint64_t)pos + pads[i] + 13;
        if (need != (unsigned __int128)(cf.file_size - off)) {
            fail(XS_ERR_FORMAT, std::string(path) + ": size identity violated (data bytes != sum signature_size x row bytes)");
            continue;
        }
        cf.data_off = off;
        return XS_OK;
    }
    if (g_err.find("size identity") != std::string::npos) return XS_ERR_FORMAT;
    return fail(XS_ERR_FORMAT, std::string(path) + ": header end magic missing");
}

static int parse_cobs_header(FILE* f, const char* path, CobsFile& cf, const char** layout_name = nullptr) {
    char magic[18];
    if (fseek(f, 0, SEEK_SET) != 0 || fread(magic, 1, 18, f) != 18 || memcmp(magic, "COBS:", 5) != 0)
        return fail(XS_ERR_FORMAT, std::string(path) + ": not a COBS index (missing 'COBS:')");
    const bool classic = memcmp(magic + 5, "CLASSIC_INDEX", 13) == 0;
    const bool compact = memcmp(magic + 5, "COMPACT_INDEX", 13) == 0;
    if (!classic && !compact) return fail(XS_ERR_FORMAT, std::string(path) + ": unknown COBS magic word");
    int first_rc = XS_OK;
    std::string first_err;
    for (const HeaderLayout& L : kHeaderLayouts) {
        if (classic && L.docs_first) continue;
        if (compact && L.hashes_u32) continue;
        const int rc = parse_cobs_header_as(f, path, L, classic, cf);
        if (rc == XS_OK) {
            if (layout_name) *layout_name = L.name;
            return XS_OK;
        }
        if (first_rc == XS_OK) { first_rc = rc; first_err = g_err; }   // report what the documented layout said
    }
    return fail(first_rc, first_err);
}

// stream rows [n_rows x src_row] of the file into dst rows [dst_stride], keeping bytes [col0, col0+n_col)
static int upload_rows(FILE* f, uint64_t file_off, uint64_t n_rows, uint32_t src_row, uint32_t col0, uint32_t n_col,
                       uint8_t* d_dst, uint32_t dst_stride, int n_sm) {
    const uint64_t STAGE = 32ULL << 20;
    uint64_t rows_per = std::max<uint64_t>(1, STAGE / src_row);
    uint64_t stage_bytes = rows_per * src_row;
    uint8_t* h_stage[2] = {nullptr, nullptr};
    uint8_t* d_stage[2] = {nullptr, nullptr};
    cudaEvent_t ev[2];
    cudaStream_t st;
    XS_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    int rc = XS_OK;
    bool direct = (src_row == dst_stride && col0 == 0 && n_col == src_row);
    for (int i = 0; i < 2 && rc == XS_OK; ++i) {
        if (cudaMallocHost(&h_stage[i], stage_bytes) != cudaSuccess) rc = fail(XS_ERR_NOMEM, "pinned staging allocation failed");
        if (rc == XS_OK && !direct && cudaMalloc(&d_stage[i], stage_bytes) != cudaSuccess) rc = fail(XS_ERR_NOMEM, "device staging allocation failed");
        if (rc == XS_OK && cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) rc = fail(XS_ERR_CUDA, "event creation failed");
    }
    if (rc == XS_OK && fseek(f, (long)file_off, SEEK_SET) != 0) rc = fail(XS_ERR_IO, "seek failed");
    uint64_t done = 0; int b = 0;
    while (rc == XS_OK && done < n_rows) {
        uint64_t n = std::min(rows_per, n_rows - done);
        cudaEventSynchronize(ev[b]);
        if (fread(h_stage[b], 1, n * src_row, f) != n * src_row) { rc = fail(XS_ERR_IO, "short read of index data"); break; }
        cudaError_t e;
        if (direct) {
            e = cudaMemcpyAsync(d_dst + done * dst_stride, h_stage[b], n * src_row, cudaMemcpyHostToDevice, st);
        } else {
            e = cudaMemcpyAsync(d_stage[b], h_stage[b], n * src_row, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) {
                k_restride<<<n_sm * 8, 256, 0, st>>>(d_stage[b], n, src_row, col0, n_col, d_dst + done * dst_stride, dst_stride);
                rc = launch_ok("k_restride");
            }
        }
        if (e != cudaSuccess) { rc = fail(XS_ERR_CUDA, std::string("index upload: ") + cudaGetErrorString(e)); break; }
        cudaEventRecord(ev[b], st);
        done += n; b ^= 1;
    }
    cudaError_t e = cudaStreamSynchronize(st);
    if (rc == XS_OK && e != cudaSuccess) rc = fail(XS_ERR_CUDA, std::string("index upload: ") + cudaGetErrorString(e));
    for (int i = 0; i < 2; ++i) {
        if (h_stage[i]) cudaFreeHost(h_stage[i]);
        if (d_stage[i]) cudaFree(d_stage[i]);
        if (h_stage[i]) cudaEventDestroy(ev[i]);
    }
    cudaStreamDestroy(st);
    return rc;
}

// ----------------------------------------------------------------------------------------
// workspace of one device-side query: 2-bit stream, bitmap, window prefix, scan temp
// ----------------------------------------------------------------------------------------
struct Workspace {
    cudaStream_t stream = nullptr;
    ~Workspace() { if (base) cudaFreeAsync(base, stream); }   // also on the error paths
    uint8_t* base = nullptr;
    uint64_t* packed = nullptr;
    uint32_t* invalid = nullptr;
    uint64_t* prefix = nullptr;
    uint64_t* prefix2 = nullptr;
    unsigned long long* counters = nullptr;
    void* scan_tmp = nullptr;
    size_t scan_bytes = 0;
};

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// windows (or work items) per sequence, then an in-place exclusive scan (CUB: plumbing only)
static int scan_windows(const WindowCountOp& op, uint64_t* d_prefix, void* tmp, size_t tmp_bytes, int n_sm, cudaStream_t s) {
    uint64_t n = op.n_seq + 1;
    uint64_t grid = std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, (uint64_t)n_sm * 16));
    k_count_windows<<<(unsigned)grid, 256, 0, s>>>(op, d_prefix);
    XS_TRY(launch_ok("k_count_windows"));
    XS_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_prefix, d_prefix, (int64_t)n, s));
    g_launches.fetch_add(2, std::memory_order_relaxed);  // init + scan kernels
    return XS_OK;
}

static const size_t N_COUNTERS = 4096;   // one dynamic tile counter per grid.y slice (pages / column blocks)

static int workspace_alloc(Workspace& ws, uint64_t n_bases, uint64_t n_seq, bool two_prefix, cudaStream_t s) {
    uint64_t n_words = n_bases / 32 + 2;
    size_t tb = 0;
    XS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, (uint64_t*)nullptr, (uint64_t*)nullptr, (int64_t)(n_seq + 1), s));
    size_t o_packed = 0;
    size_t o_inv = o_packed + align256(n_words * 8);
    size_t o_pre = o_inv + align256(n_words * 4);
    size_t o_pre2 = o_pre + align256((n_seq + 1) * 8);
    size_t o_cnt = o_pre2 + (two_prefix ? align256((n_seq + 1) * 8) : 0);
    size_t o_tmp = o_cnt + align256(N_COUNTERS * 8);
    size_t total = o_tmp + align256(tb ? tb : 1);
    XS_CUDA(cudaMallocAsync((void**)&ws.base, total, s));
    ws.stream = s;
    ws.packed = reinterpret_cast<uint64_t*>(ws.base + o_packed);
    ws.invalid = reinterpret_cast<uint32_t*>(ws.base + o_inv);
    ws.prefix = reinterpret_cast<uint64_t*>(ws.base + o_pre);
    ws.prefix2 = two_prefix ? reinterpret_cast<uint64_t*>(ws.base + o_pre2) : nullptr;
    ws.counters = reinterpret_cast<unsigned long long*>(ws.base + o_cnt);
    ws.scan_tmp = ws.base + o_tmp;
    ws.scan_bytes = tb;
    XS_CUDA(cudaMemsetAsync(ws.counters, 0, N_COUNTERS * 8, s));
    return XS_OK;
}

static int prepare_batch(Workspace& ws, SeqBatch& sb, const uint8_t* d_bases, uint64_t n_bases, const uint64_t* d_begin,
                         const uint64_t* d_end, uint64_t n_seq, uint64_t base_shift, uint32_t k, uint32_t step,
                         uint32_t chunk, int n_sm, cudaStream_t s) {
    XS_TRY(workspace_alloc(ws, n_bases, n_seq, chunk != 0, s));
    uint64_t n_words = n_bases / 32 + 2;
    uint64_t grid = std::min<uint64_t>((n_words + 255) / 256, (uint64_t)n_sm * 16);
    k_pack2bit<<<(unsigned)std::max<uint64_t>(grid, 1), 256, 0, s>>>(d_bases, n_bases, ws.packed, ws.invalid, n_words);
    XS_TRY(launch_ok("k_pack2bit"));
    WindowCountOp op{d_begin, d_end, n_seq, base_shift, n_bases, k, step, 0};
    XS_TRY(scan_windows(op, ws.prefix, ws.scan_tmp, ws.scan_bytes, n_sm, s));
    if (chunk) {
        op.chunk = chunk;
        XS_TRY(scan_windows(op, ws.prefix2, ws.scan_tmp, ws.scan_bytes, n_sm, s));
    }
    sb.packed = ws.packed; sb.invalid = ws.invalid; sb.bases = d_bases;
    sb.seq_begin = d_begin; sb.seq_end = d_end; sb.win_prefix = ws.prefix; sb.tile_counter = ws.counters;
    sb.n_seq = n_seq; sb.n_bases = n_bases; sb.base_shift = base_shift; sb.step = step; sb.k = k;
    return XS_OK;
}

// ----------------------------------------------------------------------------------------
// kernel dispatch
// ----------------------------------------------------------------------------------------
template <int K, int H>
static void launch_narrow_t(const CobsParams& p, dim3 grid, int dt, cudaStream_t s) {
    if (dt == XS_U8) k_cobs_narrow<K, H, uint8_t><<<grid, NARROW_NT, 0, s>>>(p);
    else if (dt == XS_U16) k_cobs_narrow<K, H, uint16_t><<<grid, NARROW_NT, 0, s>>>(p);
    else k_cobs_narrow<K, H, uint32_t><<<grid, NARROW_NT, 0, s>>>(p);
}
static void launch_narrow(const CobsParams& p, dim3 grid, int dt, cudaStream_t s) {
    if (p.sb.k == 21 && p.num_hashes == 7) launch_narrow_t<21, 7>(p, grid, dt, s);
    else if (p.sb.k == 31 && p.num_hashes == 1) launch_narrow_t<31, 1>(p, grid, dt, s);
    else launch_narrow_t<0, 0>(p, grid, dt, s);
}

// several pages of narrow rows (compact indices): one pass, every window hashed once (k_cobs_pages)
static size_t pages_smem(size_t n_pages) { return (size_t)(PAGES_NT / 32) * n_pages * 128 * 4; }
static const size_t PAGES_SMEM_MAX = 160 * 1024;

template <int K, int H, typename T>
static cudaError_t launch_pages_tt(const CobsParams& p, int n_sm, cudaStream_t s) {
    const size_t smem = pages_smem(p.n_pages);
    cudaError_t e = cudaFuncSetAttribute(k_cobs_pages<K, H, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_cobs_pages<K, H, T>, PAGES_NT, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorInvalidConfiguration;
    k_cobs_pages<K, H, T><<<n_sm * std::min(occ, 6), PAGES_NT, smem, s>>>(p);
    return cudaSuccess;
}
template <int K, int H>
static cudaError_t launch_pages_t(const CobsParams& p, int n_sm, int dt, cudaStream_t s) {
    if (dt == XS_U8) return launch_pages_tt<K, H, uint8_t>(p, n_sm, s);
    if (dt == XS_U16) return launch_pages_tt<K, H, uint16_t>(p, n_sm, s);
    return launch_pages_tt<K, H, uint32_t>(p, n_sm, s);
}
static cudaError_t launch_pages(const CobsParams& p, int n_sm, int dt, cudaStream_t s) {
    if (p.sb.k == 31 && p.num_hashes == 1) return launch_pages_t<31, 1>(p, n_sm, dt, s);
    if (p.sb.k == 21 && p.num_hashes == 1) return launch_pages_t<21, 1>(p, n_sm, dt, s);
    return launch_pages_t<0, 0>(p, n_sm, dt, s);
}

// one page of mid-width rows (k_cobs_mid): lanes per row from the stride
template <int K, int H, typename T>
static void launch_mid_tt(const CobsParams& p, uint32_t stride, int n_sm, cudaStream_t s) {
    static int occ[3] = {0, 0, 0};
    auto go = [&](auto kern, int slot) {
        if (!occ[slot]) {
            int o = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, MID_NT, 0) != cudaSuccess || o < 1) o = 1;
            occ[slot] = o;
        }
        kern<<<n_sm * occ[slot], MID_NT, 0, s>>>(p);
    };
    // 128-byte rows gather only 4 windows per warp instruction and are bound by dependent DRAM round trips: 64
    // registers (a few spilled words in the hash phase) for 4 CTAs per SM is worth +10 % there (18.7 against 20.7 ms,
    // profiles/r2_mid_rows.md); 32 / 64-byte rows sit on the DRAM fetch rate at 80 registers / 3 CTAs per SM
    static const int occ_sel = env_int("XS_MID_OCC", 0);      // measurement switch: 3 or 4 for every stride
    if (stride == 32) { if (occ_sel == 4) go(k_cobs_mid<K, H, T, 2, 4>, 0); else go(k_cobs_mid<K, H, T, 2, 3>, 0); }
    else if (stride == 64) { if (occ_sel == 4) go(k_cobs_mid<K, H, T, 4, 4>, 1); else go(k_cobs_mid<K, H, T, 4, 3>, 1); }
    else { if (occ_sel == 3) go(k_cobs_mid<K, H, T, 8, 3>, 2); else go(k_cobs_mid<K, H, T, 8, 4>, 2); }
}
template <int K, int H>
static void launch_mid_t(const CobsParams& p, uint32_t stride, int n_sm, int dt, cudaStream_t s) {
    if (dt == XS_U8) launch_mid_tt<K, H, uint8_t>(p, stride, n_sm, s);
    else if (dt == XS_U16) launch_mid_tt<K, H, uint16_t>(p, stride, n_sm, s);
    else launch_mid_tt<K, H, uint32_t>(p, stride, n_sm, s);
}
static void launch_mid(const CobsParams& p, uint32_t stride, int n_sm, int dt, cudaStream_t s) {
    if (p.sb.k == 21 && p.num_hashes == 7) launch_mid_t<21, 7>(p, stride, n_sm, dt, s);
    else if (p.sb.k == 31 && p.num_hashes == 1) launch_mid_t<31, 1>(p, stride, n_sm, dt, s);
    else launch_mid_t<0, 0>(p, stride, n_sm, dt, s);
}
// which kernel family a handle's rows take: the wide kernel's work items need the per-sequence chunk prefix
static bool uses_wide(const xs_cobs* ix) { return (!ix->narrow && !ix->mid) || ix->force_wide; }
// k_cobs_mid: a classic index (or a column shard of one) whose rows sit at a 32 / 64 / 128-byte stride
static bool mid_eligible(const xs_cobs* ix) {
    const char* off = getenv("XS_NO_MID_KERNEL");
    if (off && off[0] == '1') return false;
    if (ix->pages.size() != 1) return false;
    const PageDesc& pg = ix->pages[0];
    return (pg.row_stride == 32 || pg.row_stride == 64 || pg.row_stride == 128) && pg.sig_size <= 0xFFFFFFFFull &&
           ix->info.num_hashes <= (uint32_t)MID_MAXH;
}

template <int K, int H, typename T>
static cudaError_t launch_wide_tt(const WideParams& p, dim3 grid, size_t smem, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(k_cobs_wide<K, H, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_cobs_wide<K, H, T>, WIDE_NT, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorInvalidConfiguration;
    grid.x *= (unsigned)occ;     // grid.x arrives as the SM count
    k_cobs_wide<K, H, T><<<grid, WIDE_NT, smem, s>>>(p);
    return cudaSuccess;
}
template <int K, int H>
static cudaError_t launch_wide_t(const WideParams& p, dim3 grid, size_t smem, int dt, cudaStream_t s) {
    if (dt == XS_U8) return launch_wide_tt<K, H, uint8_t>(p, grid, smem, s);
    if (dt == XS_U16) return launch_wide_tt<K, H, uint16_t>(p, grid, smem, s);
    return launch_wide_tt<K, H, uint32_t>(p, grid, smem, s);
}
static cudaError_t launch_wide(const WideParams& p, dim3 grid, size_t smem, int dt, cudaStream_t s) {
    if (p.cp.sb.k == 21 && p.cp.num_hashes == 7) return launch_wide_t<21, 7>(p, grid, smem, dt, s);
    if (p.cp.sb.k == 31 && p.cp.num_hashes == 1) return launch_wide_t<31, 1>(p, grid, smem, dt, s);
    return launch_wide_t<0, 0>(p, grid, smem, dt, s);
}

// ---- bucketed probing: what the COBS and the Bloom path share on the host ------------------------------------
struct SubBatches {
    uint64_t nc_total = 0;   // chunks covering the upper bound of the window count
    uint64_t nc_sub = 0;     // chunks per sub-batch (= per scratch buffer)
    uint64_t n_sub = 0;
};

// No host read of the window count: chunking and scratch are sized from an upper bound that holds whenever the
// sequences do not overlap (sum of ((len - k) / step + 1) <= n_bases / step + n_seq); the kernels read the true count
// on the device and skip chunks beyond it, and windows beyond the bound (overlapping segments) are scored by a tail
// launch of the direct kernel that normally finds nothing to do.  Sub-batches of nc_sub chunks go through emit ->
// fetch -> reduce back to back on the caller's stream.  (Running the kernels of neighbouring sub-batches concurrently
// on several streams was measured and lost: they compete for issue slots, the LSU path and L2,
// profiles/r1_bucketed_notes.md.)  *ok = false: the batch is too small or there is too little memory.
static int plan_sub_batches(BucketState& bk, const SeqBatch& sb, size_t per_chunk, SubBatches& sub, bool* ok) {
    *ok = false;
    uint64_t budget = bk.budget.load(std::memory_order_relaxed);
    if (budget == 0) {     // asked once per handle (and again after a failed allocation): at most half of what is free
        size_t free_b = 0, total_b = 0;
        XS_CUDA(cudaMemGetInfo(&free_b, &total_b));
        budget = std::max<uint64_t>(1, std::min<uint64_t>(bk.scratch_bytes, free_b / 2));
        bk.budget.store(budget, std::memory_order_relaxed);
    }
    const uint64_t bound = sb.n_bases / sb.step + sb.n_seq;
    sub.nc_total = (bound + BK_CH - 1) / BK_CH;
    sub.nc_sub = std::min<uint64_t>(sub.nc_total, budget / per_chunk);
    if (sub.nc_sub == 0 || sub.nc_sub * BK_CH < bk.min_windows / 2) return XS_OK;    // too little memory for L2 re-use
    sub.n_sub = (sub.nc_total + sub.nc_sub - 1) / sub.nc_sub;
    sub.nc_sub = (sub.nc_total + sub.n_sub - 1) / sub.n_sub;
    sub.nc_sub = std::min<uint64_t>(sub.nc_sub, 0xFFFFFFFFu / BK_MAX_BUCKETS);
    sub.n_sub = (sub.nc_total + sub.nc_sub - 1) / sub.nc_sub;
    *ok = true;
    return XS_OK;
}

// holds the handle's lock while a bucketed query is enqueued and chains it behind the previous one
struct BucketSerial {
    BucketState& bk;
    cudaStream_t s;
    std::lock_guard<std::mutex> lock;
    BucketSerial(BucketState& b, cudaStream_t st) : bk(b), s(st), lock(b.mu) {
        if (!bk.done && cudaEventCreateWithFlags(&bk.done, cudaEventDisableTiming) != cudaSuccess) bk.done = nullptr;
        if (bk.done && bk.prev) cudaStreamWaitEvent(s, bk.done, 0);
    }
    ~BucketSerial() { if (bk.done && cudaEventRecord(bk.done, s) == cudaSuccess) bk.prev = true; }
};

static unsigned chunk_seq_grid(uint64_t nc_total, int n_sm) {
    return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((nc_total + 255) / 256, (uint64_t)n_sm * 8));
}

static bool bucket_prefetch_enabled() {
    const char* v = getenv("XS_BK_PREFETCH");      // measurement switch (profiles/experiments/bucketed_phases.py)
    return !(v && *v) || atoi(v) != 0;
}

// ---- bucketed probing, COBS (k_bucket_emit / k_bucket_fetch / k_bucket_reduce) ------------------------------
struct BucketGeom {
    uint32_t nb = 0, bshift = 0, cap = 0;
    size_t smem = 0;        // dynamic shared memory of k_bucket_emit
    size_t per_chunk = 0;   // scratch bytes per chunk
};

// geometry for an index of `sig` rows probed with `h` hashes; false when the bucketed path does not apply
static bool bucket_geometry(uint64_t sig, uint32_t h, uint32_t shift_override, BucketGeom& g) {
    if (h == 0 || h > 16) return false;
    if (shift_override) {
        if (shift_override > BK_MAX_SHIFT) return false;
        g.bshift = shift_override;
        if (((sig - 1) >> g.bshift) + 1 > BK_MAX_BUCKETS) return false;
        g.nb = (uint32_t)(((sig - 1) >> g.bshift) + 1);
    } else {
        g.bshift = 20;
        if (((sig - 1) >> g.bshift) + 1 > BK_MAX_BUCKETS) g.bshift = BK_MAX_SHIFT;
        if (((sig - 1) >> g.bshift) + 1 > BK_MAX_BUCKETS) return false;
        g.nb = (uint32_t)(((sig - 1) >> g.bshift) + 1);
        if (g.nb < 16) return false;                                  // 256 MB .. 8.6 GB of 16-byte rows
    }
    const double mean = (double)BK_CH * h / (double)g.nb;
    g.cap = ((uint32_t)std::ceil(mean + 2.8 * std::sqrt(mean)) + 7) & ~7u;
    if (g.cap > 60000) return false;       // block counts are 16-bit
    g.smem = ((size_t)g.nb * g.cap + g.nb + 2 * (BK_CH / 32) + BK_CH + BK_EMIT_NT / 32) * 4;
    if (g.smem > 200 * 1024) return false;
    g.per_chunk = (size_t)g.nb * g.cap * 20 + (size_t)g.nb * 4 + 2 * (BK_CH / 32) * 4;
    return true;
}

template <int K, int H, bool PRE>
static cudaError_t launch_bucket_emit(const BucketParams& bp, const BucketGeom& g, int n_sm, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(k_bucket_emit<K, H, PRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_bucket_emit<K, H, PRE>, BK_EMIT_NT, g.smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorInvalidConfiguration;
    KernelTimer kt(s, PROF_EMIT);
    k_bucket_emit<K, H, PRE><<<n_sm * occ, BK_EMIT_NT, g.smem, s>>>(bp);
    return cudaSuccess;
}

template <int K, int H>
static cudaError_t launch_bucket_fetch(const BucketParams& bp, int n_sm, cudaStream_t s) {
    KernelTimer kt(s, PROF_FETCH);
    static const int fetch_ctas = env_int("XS_BK_FETCH_CTAS", 8);
    if (bp.use_tma && !(bp.next_nc && bp.next_rows) && !bp.fetch_off) {       // part of the gathers through the TMA unit
        static const int var = env_int("XS_BK_TMA_VAR", 0);
        if (bp.use_tma >= 2) k_bucket_fetch_tma<2, 0><<<n_sm * fetch_ctas, BK_NT, 0, s>>>(bp);
        else if (var == 1) k_bucket_fetch_tma<1, 1><<<n_sm * fetch_ctas, BK_NT, 0, s>>>(bp);
        else if (var == 2) k_bucket_fetch_tma<1, 2><<<n_sm * fetch_ctas, BK_NT, 0, s>>>(bp);
        else if (var == 3) k_bucket_fetch_tma<1, 3><<<n_sm * fetch_ctas, BK_NT, 0, s>>>(bp);
        else k_bucket_fetch_tma<1, 0><<<n_sm * fetch_ctas, BK_NT, 0, s>>>(bp);
        return cudaSuccess;
    }
    if (bp.next_nc && !bp.next_rows) {   // side-stream mode: plain fetch, the grid leaves room for k_bucket_hash's CTAs
        k_bucket_fetch<K, H, false><<<n_sm * fetch_ctas, BK_NT, 0, s>>>(bp);
        return cudaSuccess;
    }
    if (bp.next_nc) {
        int occ = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_bucket_fetch<K, H, true>, BK_NT, 0);
        if (e != cudaSuccess) return e;
        if (occ < 1) return cudaErrorInvalidConfiguration;
        k_bucket_fetch<K, H, true><<<n_sm * std::min(occ, 8), BK_NT, 0, s>>>(bp);
    } else {
        k_bucket_fetch<K, H, false><<<n_sm * 8, BK_NT, 0, s>>>(bp);
    }
    return cudaSuccess;
}

template <int K, int H>
static void launch_bucket_reduce(const BucketParams& bp, int n_sm, int dt, cudaStream_t s) {
        KernelTimer kt(s, PROF_REDUCE);
        const bool pk = bp.pack_id != 0;
        if (dt == XS_U8) { if (pk) k_bucket_reduce<K, H, uint8_t, true><<<n_sm * 4, BK_NT, 0, s>>>(bp); else k_bucket_reduce<K, H, uint8_t, false><<<n_sm * 4, BK_NT, 0, s>>>(bp); }
        else if (dt == XS_U16) { if (pk) k_bucket_reduce<K, H, uint16_t, true><<<n_sm * 4, BK_NT, 0, s>>>(bp); else k_bucket_reduce<K, H, uint16_t, false><<<n_sm * 4, BK_NT, 0, s>>>(bp); }
        else { if (pk) k_bucket_reduce<K, H, uint32_t, true><<<n_sm * 4, BK_NT, 0, s>>>(bp); else k_bucket_reduce<K, H, uint32_t, false><<<n_sm * 4, BK_NT, 0, s>>>(bp); }
}

// side-stream hash-ahead: emit(i) [scatter of precomputed row ids] -> { fetch(i) -> reduce(i) } next to k_bucket_hash(i + 1)
template <int K, int H>
static cudaError_t launch_bucket_side(const BucketParams& bp, const BucketGeom& g, int n_sm, int dt, cudaStream_t s, cudaStream_t side,
                                      cudaEvent_t ev_emit, cudaEvent_t ev_hash) {
    static const int hash_ctas = env_int("XS_BK_HASH_CTAS", 1);
    cudaError_t e = launch_bucket_emit<K, H, true>(bp, g, n_sm, s);
    if (e != cudaSuccess) return e;
    if (bp.next_nc) {
        cudaEventRecord(ev_emit, s);                 // the row-id buffer is free once emit has read it
        cudaStreamWaitEvent(side, ev_emit, 0);
        BucketParams hb = bp;
        hb.next_rows = bp.pre_rows; hb.next_skp = bp.pre_skp;
        hb.counter = bp.counter + 4;                 // the next sub-batch's counters
        k_bucket_hash<K, H><<<n_sm * hash_ctas, BK_NT, 0, side>>>(hb);
        cudaEventRecord(ev_hash, side);
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    BucketParams fb = bp;
    fb.next_rows = nullptr;                          // plain fetch
    e = launch_bucket_fetch<K, H>(fb, n_sm, s);
    if (e != cudaSuccess) return e;
    launch_bucket_reduce<K, H>(bp, n_sm, dt, s);
    if (bp.next_nc) cudaStreamWaitEvent(s, ev_hash, 0);
    g_launches.fetch_add(3, std::memory_order_relaxed);
    return cudaGetLastError();
}

template <int K, int H>
static cudaError_t launch_bucket_t(const BucketParams& bp, const BucketGeom& g, int n_sm, int dt, cudaStream_t s) {
    cudaError_t e = bp.pre_rows ? launch_bucket_emit<K, H, true>(bp, g, n_sm, s) : launch_bucket_emit<K, H, false>(bp, g, n_sm, s);
    if (e != cudaSuccess) return e;
    e = launch_bucket_fetch<K, H>(bp, n_sm, s);
    if (e != cudaSuccess) return e;
    {
        KernelTimer kt(s, PROF_REDUCE);
        const bool pk = bp.pack_id != 0;
        if (dt == XS_U8) { if (pk) k_bucket_reduce<K, H, uint8_t, true><<<n_sm * 4, BK_NT, 0, s>>>(bp); else k_bucket_reduce<K, H, uint8_t, false><<<n_sm * 4, BK_NT, 0, s>>>(bp); }
        else if (dt == XS_U16) { if (pk) k_bucket_reduce<K, H, uint16_t, true><<<n_sm * 4, BK_NT, 0, s>>>(bp); else k_bucket_reduce<K, H, uint16_t, false><<<n_sm * 4, BK_NT, 0, s>>>(bp); }
        else { if (pk) k_bucket_reduce<K, H, uint32_t, true><<<n_sm * 4, BK_NT, 0, s>>>(bp); else k_bucket_reduce<K, H, uint32_t, false><<<n_sm * 4, BK_NT, 0, s>>>(bp); }
    }
    g_launches.fetch_add(3, std::memory_order_relaxed);
    return cudaGetLastError();
}

// measurement switch: 0 = k_bucket_emit hashes its own windows (round 1); 1 = k_bucket_fetch of sub-batch i also hashes
// the windows of sub-batch i + 1; 2 = k_bucket_hash of sub-batch i + 1 on a side stream next to fetch / reduce of sub-batch i
static int bucket_hash_ahead_mode() {
    const char* v = getenv("XS_BK_HASH_AHEAD");
    return (v && *v) ? atoi(v) : 0;
}


// returns XS_OK with *handled = false when the batch should go through k_cobs_narrow instead
static int cobs_launch_bucketed(xs_cobs* ix, const CobsParams& p, int dt, cudaStream_t s, bool* handled) {
    *handled = false;
    BucketState& bk = ix->bk;
    if (!bk.enabled || !ix->narrow || ix->force_wide || ix->pages.size() != 1) return XS_OK;
    if (p.sb.n_bases / p.sb.step < bk.min_windows) return XS_OK;
    BucketGeom g;
    if (!bucket_geometry(ix->pages[0].sig_size, ix->info.num_hashes, bk.shift, g)) return XS_OK;
    // hash-ahead: k_bucket_fetch of sub-batch i also computes the row ids of sub-batch i + 1 (XXH64 in the issue slots its
    // warps leave idle while they wait on L2 gathers); k_bucket_emit then only scatters them
    const uint32_t nh = ix->info.num_hashes;
    const int ahead_mode = bucket_hash_ahead_mode();
    const bool ahead = ahead_mode != 0;
    const size_t pre_chunk = ahead ? (size_t)nh * BK_CH * 4 + (BK_CH / 32) * 4 : 0;
    SubBatches sub;
    bool ok = false;
    XS_TRY(plan_sub_batches(bk, p.sb, g.per_chunk + pre_chunk, sub, &ok));
    if (!ok) return XS_OK;

    const size_t o_rows = 0;
    const size_t o_rec = o_rows + align256(sub.nc_sub * g.nb * g.cap * 16);
    const size_t o_bc = o_rec + align256(sub.nc_sub * g.nb * g.cap * 4);
    const size_t o_cb = o_bc + align256(sub.nc_sub * g.nb * 2);
    const size_t o_ovf = o_cb + align256(sub.nc_sub * g.nb * 2);
    const size_t o_ctr = o_ovf + align256(sub.nc_sub * 2 * (BK_CH / 32) * 4);
    const size_t o_seq = o_ctr + align256((sub.n_sub + 1) * 4 * 8);
    const size_t o_pre = o_seq + align256(sub.nc_total * 8);
    const size_t o_pskp = o_pre + (ahead ? align256(sub.nc_sub * nh * BK_CH * 4) : 0);
    const size_t bytes = o_pskp + (ahead ? align256(sub.nc_sub * (BK_CH / 32) * 4) : 0);
    uint8_t* d = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&d, bytes, s);
    if (e != cudaSuccess) {                                       // no room for the scratch: direct gathers
        cudaGetLastError();
        bk.budget.store(0, std::memory_order_relaxed);
        return XS_OK;
    }
    {
        BucketSerial serial(bk, s);
        e = cudaMemsetAsync(d + o_ctr, 0, (sub.n_sub + 1) * 4 * 8, s);
        k_bucket_chunk_seq<<<chunk_seq_grid(sub.nc_total, ix->n_sm), 256, 0, s>>>(p.sb, sub.nc_total, reinterpret_cast<uint64_t*>(d + o_seq));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        const bool prefetch = bucket_prefetch_enabled();
        const bool k21h7 = p.sb.k == 21 && p.num_hashes == 7;
        static const int tma_q = env_int("XS_BK_TMA", 0);      // records per lane and pass that go through tile::gather4 (0, 1, 2)
        auto fill = [&](BucketParams& bp, uint64_t i) {
            bp.cp = p;
            if (ix->has_tmap && tma_q) { bp.tmap = ix->tmap; bp.use_tma = (uint32_t)tma_q; }
            bp.rows = reinterpret_cast<uint4*>(d + o_rows); bp.rec = reinterpret_cast<uint32_t*>(d + o_rec);
            bp.cnt_bc = reinterpret_cast<uint16_t*>(d + o_bc); bp.cnt_cb = reinterpret_cast<uint16_t*>(d + o_cb);
            bp.ovf = reinterpret_cast<uint32_t*>(d + o_ovf);
            bp.chunk_seq = reinterpret_cast<const uint64_t*>(d + o_seq);
            bp.counter = reinterpret_cast<unsigned long long*>(d + o_ctr) + 4 * i;
            bp.chunk0 = std::min<uint64_t>(i, sub.n_sub) * sub.nc_sub;
            bp.nc = i < sub.n_sub ? (uint32_t)std::min<uint64_t>(sub.nc_sub, sub.nc_total - bp.chunk0) : 0u;
            bp.n_buckets = g.nb; bp.bshift = g.bshift; bp.cap = g.cap;
            bp.pack_id = ix->pages[0].n_docs <= 96 ? 1u : 0u;
            bp.prefetch = prefetch ? 1u : 0u;
        };
        if (ahead_mode == 2 && !bk.side) {
            if (cudaStreamCreateWithFlags(&bk.side, cudaStreamNonBlocking) != cudaSuccess ||
                cudaEventCreateWithFlags(&bk.ev_emit, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&bk.ev_hash, cudaEventDisableTiming) != cudaSuccess)
                e = cudaErrorUnknown;
        }
        if (ahead && e == cudaSuccess) {      // prologue: the row ids of sub-batch 0
            BucketParams bp{};
            fill(bp, sub.n_sub);
            bp.fetch_off = 1;
            bp.next_rows = reinterpret_cast<uint32_t*>(d + o_pre); bp.next_skp = reinterpret_cast<uint32_t*>(d + o_pskp);
            bp.next_chunk0 = 0;
            bp.next_nc = (uint32_t)std::min<uint64_t>(sub.nc_sub, sub.nc_total);
            if (ahead_mode == 2) {
                if (k21h7) k_bucket_hash<21, 7><<<ix->n_sm * 4, BK_NT, 0, s>>>(bp); else k_bucket_hash<0, 0><<<ix->n_sm * 4, BK_NT, 0, s>>>(bp);
                e = cudaGetLastError();
            } else {
                e = k21h7 ? launch_bucket_fetch<21, 7>(bp, ix->n_sm, s) : launch_bucket_fetch<0, 0>(bp, ix->n_sm, s);
            }
            g_launches.fetch_add(1, std::memory_order_relaxed);
        }
        for (uint64_t i = 0; i < sub.n_sub && e == cudaSuccess; ++i) {
            BucketParams bp{};
            fill(bp, i);
            if (ahead) {
                bp.pre_rows = reinterpret_cast<uint32_t*>(d + o_pre); bp.pre_skp = reinterpret_cast<uint32_t*>(d + o_pskp);
                if (i + 1 < sub.n_sub) {
                    bp.next_rows = bp.pre_rows; bp.next_skp = bp.pre_skp;
                    bp.next_chunk0 = (i + 1) * sub.nc_sub;
                    bp.next_nc = (uint32_t)std::min<uint64_t>(sub.nc_sub, sub.nc_total - bp.next_chunk0);
                }
            }
            if (ahead_mode == 2)
                e = k21h7 ? launch_bucket_side<21, 7>(bp, g, ix->n_sm, dt, s, bk.side, bk.ev_emit, bk.ev_hash)
                          : launch_bucket_side<0, 0>(bp, g, ix->n_sm, dt, s, bk.side, bk.ev_emit, bk.ev_hash);
            else
                e = k21h7 ? launch_bucket_t<21, 7>(bp, g, ix->n_sm, dt, s) : launch_bucket_t<0, 0>(bp, g, ix->n_sm, dt, s);
        }
        if (e == cudaSuccess) {
            CobsParams tail = p;
            tail.win_begin = sub.nc_total * BK_CH;
            KernelTimer kt(s, PROF_DIRECT);
            launch_narrow(tail, dim3((unsigned)(ix->n_sm * 4), 1), dt, s);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            e = cudaGetLastError();
        }
    }
    cudaFreeAsync(d, s);
    if (e != cudaSuccess) return fail(XS_ERR_CUDA, std::string("bucketed query: ") + cudaGetErrorString(e));
    *handled = true;
    bk.queries.fetch_add(1, std::memory_order_relaxed);
    return XS_OK;
}

static int dtype_size(int dt) { return (dt == XS_U8 || dt == XS_U16 || dt == XS_U32) ? dt : 0; }

// all pointers device; asynchronous on s
// the scoring kernel of a prepared batch (narrow or wide rows)
static int cobs_launch(xs_cobs* ix, const SeqBatch& sb, const uint64_t* chunk_prefix, int dt, void* d_out, cudaStream_t s,
                       uint64_t ld_override = 0) {
    const uint64_t ld = ld_override ? ld_override : ix->info.doc_end - ix->info.doc_begin;
    const bool wide = uses_wide(ix);
    CobsParams p{};
    p.sb = sb; p.pages = ix->d_pages; p.n_pages = (uint32_t)ix->pages.size();
    p.num_hashes = ix->info.num_hashes; p.canonicalize = ix->info.canonicalize; p.policy = ix->info.policy;
    p.out = d_out; p.ld = ld; p.seq0 = 0;
    if (!wide) {
        if (ix->pages_kernel) {
            KernelTimer kt(s);
            cudaError_t e = launch_pages(p, ix->n_sm, dt, s);
            if (e != cudaSuccess) return fail(XS_ERR_CUDA, std::string("k_cobs_pages: ") + cudaGetErrorString(e));
            return launch_ok("k_cobs_pages");
        }
        if (ix->mid) {
            static const bool all_rows = [] { const char* v = getenv("XS_MID_ALL_ROWS"); return v && v[0] == '1'; }();
            p.all_rows = all_rows ? 1u : 0u;
            KernelTimer kt(s);
            launch_mid(p, ix->pages[0].row_stride, ix->n_sm, dt, s);
            return launch_ok("k_cobs_mid");
        }
        bool handled = false;
        XS_TRY(cobs_launch_bucketed(ix, p, dt, s, &handled));
        if (handled) return XS_OK;
        dim3 grid((unsigned)(ix->n_sm * 4), (unsigned)ix->pages.size());
        KernelTimer kt(s);
        launch_narrow(p, grid, dt, s);
        XS_TRY(launch_ok("k_cobs_narrow"));
    } else {
        WideParams wp{};
        wp.cp = p; wp.blocks = ix->d_blocks; wp.chunk_prefix = chunk_prefix;
        uint32_t max_cols = 0;
        for (const ColBlock& b : ix->blocks) max_cols = std::max(max_cols, b.n_cols);
        size_t smem = (size_t)WIDE_CHUNK * ix->info.num_hashes * 8 + (size_t)max_cols * 128 * 4;
        dim3 grid((unsigned)ix->n_sm, (unsigned)ix->blocks.size());
        KernelTimer kt(s);
        cudaError_t e = launch_wide(wp, grid, smem, dt, s);
        if (e != cudaSuccess) return fail(XS_ERR_CUDA, std::string("k_cobs_wide: ") + cudaGetErrorString(e));
        XS_TRY(launch_ok("k_cobs_wide"));
    }
    return XS_OK;
}

static int cobs_query_dev(xs_cobs* ix, const uint8_t* d_bases, uint64_t n_bases, const uint64_t* d_begin,
                          const uint64_t* d_end, uint64_t n_seq, uint64_t base_shift, uint32_t step, int dt,
                          void* d_out, cudaStream_t s, uint64_t ld_override = 0) {
    const uint64_t ld = ld_override ? ld_override : ix->info.doc_end - ix->info.doc_begin;
    XS_CUDA(cudaMemsetAsync(d_out, 0, n_seq * ld * (uint64_t)dt, s));
    if (n_seq == 0) return XS_OK;
    Workspace ws;
    SeqBatch sb{};
    const bool wide = uses_wide(ix);
    XS_TRY(prepare_batch(ws, sb, d_bases, n_bases, d_begin, d_end, n_seq, base_shift, ix->info.term_size, step,
                         wide ? WIDE_CHUNK : 0, ix->n_sm, s));
    return cobs_launch(ix, sb, ws.prefix2, dt, d_out, s, ld_override);   // ws goes back to the stream-ordered pool on scope exit
}

// ---- Bloom, bucketed probing (k_bbucket_emit / k_bbucket_fetch / k_bbucket_reduce) ---------------------------
struct BloomBucketGeom {
    uint32_t nb = 0, bshift = 0, cap = 0;
    size_t smem = 0, per_chunk = 0;
};

static bool bloom_bucket_geometry(uint64_t n_bits, uint32_t k_hashes, uint32_t shift_override, BloomBucketGeom& g) {
    if (k_hashes == 0 || k_hashes > 16 || n_bits == 0) return false;
    if (shift_override) {
        if (shift_override < 3 || shift_override > 31) return false;
        g.bshift = shift_override;
    } else {
        g.bshift = 27;                                                   // 16 MB of the bit array per bucket
        if (((n_bits - 1) >> g.bshift) + 1 > BK_MAX_BUCKETS) g.bshift = 28;
    }
    if (((n_bits - 1) >> g.bshift) + 1 > BK_MAX_BUCKETS) return false;
    g.nb = (uint32_t)(((n_bits - 1) >> g.bshift) + 1);
    if (!shift_override && g.nb < 16) return false;                      // 256 MB .. 8 GB filters
    const double mean = (double)BK_CH * k_hashes / (double)g.nb;
    g.cap = ((uint32_t)std::ceil(mean + 2.8 * std::sqrt(mean)) + 15) & ~15u;
    if (g.cap > 60000) return false;
    const uint32_t nb4 = (g.nb + 3) & ~3u;
    g.smem = ((size_t)g.nb * g.cap + nb4 + BK_CH / 32 + BK_CH + BK_EMIT_NT / 32) * 4 + (size_t)g.nb * g.cap * 2;
    if (g.smem > 200 * 1024) return false;
    g.per_chunk = (size_t)g.nb * g.cap * 7 + (size_t)g.nb * 2 + (BK_CH / 32) * 4;
    return true;
}

template <int K, int KH>
static cudaError_t launch_bbucket_t(const BloomBucketParams& bp, const BloomBucketGeom& g, int n_sm, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(k_bbucket_emit<K, KH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_bbucket_emit<K, KH>, BK_EMIT_NT, g.smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorInvalidConfiguration;
    {
        KernelTimer kt(s, PROF_EMIT);
        k_bbucket_emit<K, KH><<<n_sm * occ, BK_EMIT_NT, g.smem, s>>>(bp);
    }
    {
        KernelTimer kt(s, PROF_FETCH);
        k_bbucket_fetch<<<n_sm * 8, BK_NT, 0, s>>>(bp);
    }
    {
        KernelTimer kt(s, PROF_REDUCE);
        k_bbucket_reduce<K><<<n_sm * 4, BK_NT, 0, s>>>(bp);
    }
    g_launches.fetch_add(3, std::memory_order_relaxed);
    return cudaGetLastError();
}

static void launch_bloom_kernel(const BloomParams& p, uint32_t k, int n_sm, cudaStream_t s) {
    dim3 grid((unsigned)(n_sm * 4));
    if (k == 21) k_bloom<21><<<grid, BLOOM_NT, 0, s>>>(p);
    else if (k == 31) k_bloom<31><<<grid, BLOOM_NT, 0, s>>>(p);
    else k_bloom<0><<<grid, BLOOM_NT, 0, s>>>(p);
}

// returns XS_OK with *handled = false when the batch should go through k_bloom instead (same contract as
// cobs_launch_bucketed)
static int bloom_launch_bucketed(xs_bloom* bf, const BloomParams& p, cudaStream_t s, bool* handled) {
    *handled = false;
    BucketState& bk = bf->bk;
    if (!bk.enabled || p.literal) return XS_OK;
    if (p.sb.n_bases / p.sb.step < bk.min_windows) return XS_OK;
    BloomBucketGeom g;
    if (!bloom_bucket_geometry(bf->info.n_bits, (uint32_t)bf->info.k_hashes, bk.shift, g)) return XS_OK;
    SubBatches sub;
    bool ok = false;
    XS_TRY(plan_sub_batches(bk, p.sb, g.per_chunk, sub, &ok));
    if (!ok) return XS_OK;

    const size_t o_pos = 0;
    const size_t o_wid = o_pos + align256(sub.nc_sub * g.nb * g.cap * 4);
    const size_t o_res = o_wid + align256(sub.nc_sub * g.nb * g.cap * 2);
    const size_t o_bc = o_res + align256(sub.nc_sub * g.nb * g.cap);
    const size_t o_ovf = o_bc + align256(sub.nc_sub * g.nb * 2);
    const size_t o_ctr = o_ovf + align256(sub.nc_sub * (BK_CH / 32) * 4);
    const size_t ctr_bytes = sub.n_sub * 3 * 8 + 16;             // work counters + {sampled, members}
    const size_t o_seq = o_ctr + align256(ctr_bytes);
    const size_t bytes = o_seq + align256(sub.nc_total * 8);
    uint8_t* d = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&d, bytes, s);
    if (e != cudaSuccess) {
        cudaGetLastError();
        bk.budget.store(0, std::memory_order_relaxed);
        return XS_OK;
    }
    BloomParams pa = p;                                           // the query with the device-side choice attached
    unsigned long long* d_adapt = reinterpret_cast<unsigned long long*>(d + o_ctr) + 3 * sub.n_sub;
    if (bf->bucket_member_pct) { pa.adapt = d_adapt; pa.adapt_pct = bf->bucket_member_pct; }
    const uint32_t k = bf->info.term_size;
    {
        BucketSerial serial(bk, s);
        e = cudaMemsetAsync(d + o_ctr, 0, ctr_bytes, s);
        k_bucket_chunk_seq<<<chunk_seq_grid(sub.nc_total, bf->n_sm), 256, 0, s>>>(p.sb, sub.nc_total, reinterpret_cast<uint64_t*>(d + o_seq));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (pa.adapt) {                                           // ~256 k windows spread over the batch
            const uint64_t stride = std::max<uint64_t>(1, (sub.nc_total * BK_CH) >> 18);
            if (k == 21) k_bloom_sample<21><<<bf->n_sm * 2, 256, 0, s>>>(p, stride, d_adapt);
            else if (k == 31) k_bloom_sample<31><<<bf->n_sm * 2, 256, 0, s>>>(p, stride, d_adapt);
            else k_bloom_sample<0><<<bf->n_sm * 2, 256, 0, s>>>(p, stride, d_adapt);
            g_launches.fetch_add(1, std::memory_order_relaxed);
        }
        const bool prefetch = bucket_prefetch_enabled();
        for (uint64_t i = 0; i < sub.n_sub && e == cudaSuccess; ++i) {
            BloomBucketParams bp{};
            bp.bl = pa;
            bp.pos = reinterpret_cast<uint32_t*>(d + o_pos); bp.wid = reinterpret_cast<uint16_t*>(d + o_wid); bp.res = d + o_res;
            bp.cnt_bc = reinterpret_cast<uint16_t*>(d + o_bc); bp.ovf = reinterpret_cast<uint32_t*>(d + o_ovf);
            bp.chunk_seq = reinterpret_cast<const uint64_t*>(d + o_seq);
            bp.counter = reinterpret_cast<unsigned long long*>(d + o_ctr) + 3 * i;
            bp.chunk0 = i * sub.nc_sub;
            bp.nc = (uint32_t)std::min<uint64_t>(sub.nc_sub, sub.nc_total - bp.chunk0);
            bp.n_buckets = g.nb; bp.bshift = g.bshift; bp.cap = g.cap; bp.prefetch = prefetch ? 1u : 0u;
            if (k == 21 && bf->info.k_hashes == 6) e = launch_bbucket_t<21, 6>(bp, g, bf->n_sm, s);
            else e = launch_bbucket_t<0, 0>(bp, g, bf->n_sm, s);
        }
        if (e == cudaSuccess) {
            BloomParams tail = pa;
            tail.win_begin = sub.nc_total * BK_CH;
            KernelTimer kt(s, PROF_DIRECT);
            launch_bloom_kernel(tail, k, bf->n_sm, s);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            e = cudaGetLastError();
        }
    }
    cudaFreeAsync(d, s);
    if (e != cudaSuccess) return fail(XS_ERR_CUDA, std::string("bucketed Bloom query: ") + cudaGetErrorString(e));
    *handled = true;
    bk.queries.fetch_add(1, std::memory_order_relaxed);
    return XS_OK;
}

static int bloom_launch(xs_bloom* bf, const SeqBatch& sb, uint32_t* d_out, cudaStream_t s, bool literal = false) {
    BloomParams p{};
    p.sb = sb;
    p.literal = literal ? 1u : 0u;
    p.bits = bf->d_bits; p.n_bits = bf->info.n_bits; p.magic = magic_of(bf->info.n_bits);
    p.k_hashes = (uint32_t)bf->info.k_hashes; p.out = d_out; p.seq0 = 0;
    bool handled = false;
    XS_TRY(bloom_launch_bucketed(bf, p, s, &handled));
    if (handled) return XS_OK;
    KernelTimer kt(s);
    launch_bloom_kernel(p, bf->info.term_size, bf->n_sm, s);
    XS_TRY(launch_ok("k_bloom"));
    return XS_OK;
}

static int bloom_query_dev(xs_bloom* bf, const uint8_t* d_bases, uint64_t n_bases, const uint64_t* d_begin,
                           const uint64_t* d_end, uint64_t n_seq, uint64_t base_shift, uint32_t step,
                           uint32_t* d_out, cudaStream_t s) {
    XS_CUDA(cudaMemsetAsync(d_out, 0, n_seq * 4, s));
    if (n_seq == 0) return XS_OK;
    Workspace ws;
    SeqBatch sb{};
    XS_TRY(prepare_batch(ws, sb, d_bases, n_bases, d_begin, d_end, n_seq, base_shift, bf->info.term_size, step, 0, bf->n_sm, s));
    return bloom_launch(bf, sb, d_out, s);
}

// ----------------------------------------------------------------------------------------
// low-latency path for small host batches (single records through the Search / Bloom shims, MLST chunk sets):
// one pinned staging buffer, one H2D copy (offsets + host-computed window prefix + bases), one memset, the
// packing kernel, the scoring kernel, one D2H copy — on a stream and buffers cached per host thread and device.
// ----------------------------------------------------------------------------------------
static const uint64_t SMALL_MAX_SEQ = 4096, SMALL_MAX_SPAN = 1 << 20, SMALL_MAX_OUT = 1 << 20;

struct SmallCtx {
    int device = -1;
    cudaStream_t stream = nullptr;
    uint8_t* h = nullptr;   // pinned staging
    uint8_t* d = nullptr;   // device mirror + workspace
    size_t h_bytes = 0, d_bytes = 0;
    ~SmallCtx() {
        if (device < 0) return;
        int prev = -1;
        if (cudaGetDevice(&prev) != cudaSuccess) return;    // runtime already shut down
        if (cudaSetDevice(device) == cudaSuccess) {
            if (h) cudaFreeHost(h);
            if (d) cudaFree(d);
            if (stream) cudaStreamDestroy(stream);
        }
        cudaSetDevice(prev);
    }
};

static SmallCtx* small_ctx(int device) {
    static thread_local std::vector<SmallCtx*> ctxs;   // leaked on purpose at thread exit only if CUDA is gone
    for (SmallCtx* c : ctxs) if (c->device == device) return c;
    SmallCtx* c = new SmallCtx();
    c->h_bytes = 2 * SMALL_MAX_SEQ * 8 * 2 + SMALL_MAX_SPAN + SMALL_MAX_OUT + 4096;
    c->d_bytes = c->h_bytes + SMALL_MAX_SPAN / 2 + N_COUNTERS * 8 + 8192;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess || cudaMallocHost((void**)&c->h, c->h_bytes) != cudaSuccess ||
        cudaMalloc((void**)&c->d, c->d_bytes) != cudaSuccess) {
        if (c->h) cudaFreeHost(c->h);
        if (c->d) cudaFree(c->d);
        if (c->stream) cudaStreamDestroy(c->stream);
        delete c;
        return nullptr;
    }
    c->device = device;
    ctxs.push_back(c);
    return c;
}

// returns XS_OK and *handled = true when the batch went through the small path
template <typename LaunchFn>
static int small_query(int device, int n_sm, uint32_t k, uint32_t step, uint32_t chunk, uint32_t n_counters, const uint8_t* bases,
                       uint64_t n_bases, const uint64_t* seq_begin, const uint64_t* seq_end, uint64_t n_seq, uint64_t out_row_bytes,
                       void* out, bool* handled, LaunchFn launch) {
    *handled = false;
    if (n_seq == 0 || n_seq > SMALL_MAX_SEQ || n_seq * out_row_bytes > SMALL_MAX_OUT) return XS_OK;
    uint64_t lo = ~0ULL, hi = 0;
    for (uint64_t i = 0; i < n_seq; ++i) {
        if (seq_end[i] < seq_begin[i]) continue;
        lo = std::min(lo, seq_begin[i]); hi = std::max(hi, seq_end[i]);
    }
    if (lo == ~0ULL) { lo = 0; hi = 0; }
    if (hi > n_bases) return fail(XS_ERR_ARG, "sequence offsets exceed n_bases");
    const uint64_t span = hi - lo;
    if (span > SMALL_MAX_SPAN) return XS_OK;
    SmallCtx* c = small_ctx(device);
    if (!c) return XS_OK;   // fall through to the pipelined path, which reports allocation problems
    cudaStream_t s = c->stream;
    // staging layout (host and device identical): begin | end | prefix | prefix2 | bases
    const size_t o_b = 0, o_e = align256(n_seq * 8), o_p = o_e + align256(n_seq * 8), o_p2 = o_p + align256((n_seq + 1) * 8);
    const size_t o_bases = o_p2 + align256((n_seq + 1) * 8);
    const size_t up_bytes = o_bases + align256(span + 64);
    uint64_t* hb = reinterpret_cast<uint64_t*>(c->h + o_b);
    uint64_t* he = reinterpret_cast<uint64_t*>(c->h + o_e);
    uint64_t* hp = reinterpret_cast<uint64_t*>(c->h + o_p);
    uint64_t* hp2 = reinterpret_cast<uint64_t*>(c->h + o_p2);
    uint64_t acc = 0, acc2 = 0;
    for (uint64_t i = 0; i < n_seq; ++i) {
        hb[i] = seq_begin[i]; he[i] = seq_end[i];
        uint64_t nw = windows_of(seq_begin[i], seq_end[i], lo, span, k, step);
        hp[i] = acc; acc += nw;
        hp2[i] = acc2; acc2 += chunk ? (nw + chunk - 1) / chunk : 0;
    }
    hp[n_seq] = acc; hp2[n_seq] = acc2;
    if (span) memcpy(c->h + o_bases, bases + lo, span);
    // device-only regions: 2-bit stream | bitmap | counters | out   (counters and out are cleared by one memset)
    const uint64_t n_words = span / 32 + 2;
    const size_t o_packed = up_bytes, o_inv = o_packed + align256(n_words * 8), o_cnt = o_inv + align256(n_words * 4);
    const size_t o_out = o_cnt + align256((size_t)n_counters * 8);
    const size_t out_bytes = n_seq * out_row_bytes;
    if (o_out + out_bytes > c->d_bytes || up_bytes + out_bytes > c->h_bytes) return XS_OK;
    XS_CUDA(cudaMemcpyAsync(c->d, c->h, up_bytes, cudaMemcpyHostToDevice, s));
    XS_CUDA(cudaMemsetAsync(c->d + o_cnt, 0, (o_out - o_cnt) + out_bytes, s));
    uint64_t grid = std::max<uint64_t>(1, std::min<uint64_t>((n_words + 255) / 256, (uint64_t)n_sm * 16));
    k_pack2bit<<<(unsigned)grid, 256, 0, s>>>(c->d + o_bases, span, reinterpret_cast<uint64_t*>(c->d + o_packed),
                                              reinterpret_cast<uint32_t*>(c->d + o_inv), n_words);
    XS_TRY(launch_ok("k_pack2bit"));
    SeqBatch sb{};
    sb.packed = reinterpret_cast<const uint64_t*>(c->d + o_packed); sb.invalid = reinterpret_cast<const uint32_t*>(c->d + o_inv);
    sb.bases = c->d + o_bases; sb.seq_begin = reinterpret_cast<const uint64_t*>(c->d + o_b);
    sb.seq_end = reinterpret_cast<const uint64_t*>(c->d + o_e); sb.win_prefix = reinterpret_cast<const uint64_t*>(c->d + o_p);
    sb.tile_counter = reinterpret_cast<unsigned long long*>(c->d + o_cnt);
    sb.n_seq = n_seq; sb.n_bases = span; sb.base_shift = lo; sb.step = step; sb.k = k;
    XS_TRY(launch(sb, reinterpret_cast<const uint64_t*>(c->d + o_p2), c->d + o_out, s));
    uint8_t* h_out = c->h + up_bytes;
    XS_CUDA(cudaMemcpyAsync(h_out, c->d + o_out, out_bytes, cudaMemcpyDeviceToHost, s));
    XS_CUDA(cudaStreamSynchronize(s));
    memcpy(out, h_out, out_bytes);
    *handled = true;
    return XS_OK;
}

// ----------------------------------------------------------------------------------------
// host-buffer pipeline: batches of sequences, three streams, copies overlapped with kernels
// ----------------------------------------------------------------------------------------
struct HostBatchPlan {
    uint64_t i0, i1;   // sequences [i0, i1)
    uint64_t lo, hi;   // byte span of `bases` they touch
};

// next batch starting at i0: stops at max_bases of span or max_seq sequences
static HostBatchPlan plan_batch(const uint64_t* b, const uint64_t* e, uint64_t n_seq, uint64_t i0, uint64_t max_span,
                                uint64_t max_seq) {
    HostBatchPlan pl{i0, i0, ~0ULL, 0};
    uint64_t i = i0;
    for (; i < n_seq && i - i0 < max_seq; ++i) {
        uint64_t lo = std::min(pl.lo, b[i]), hi = std::max(pl.hi, std::max(e[i], b[i]));
        if (i > i0 && hi - lo > max_span) break;
        pl.lo = lo; pl.hi = hi;
    }
    pl.i1 = i;
    if (pl.lo == ~0ULL) { pl.lo = 0; pl.hi = 0; }
    return pl;
}

template <typename QueryFn>
static int host_pipeline(const uint8_t* bases, uint64_t n_bases, const uint64_t* seq_begin, const uint64_t* seq_end,
                         uint64_t n_seq, uint64_t out_row_bytes, QueryFn query) {
    // query(d_bases, span, d_begin, d_end, n, base_shift, d_out /* n * out_row_bytes */, first_seq, stream) enqueues the
    // kernels of one batch and the device->host copies of its results
    const int NS = 3;
    const uint64_t MAX_SPAN = 64ULL << 20;
    const uint64_t MAX_OUT = 256ULL << 20;
    const uint64_t max_seq = std::max<uint64_t>(1, std::min<uint64_t>(MAX_OUT / std::max<uint64_t>(out_row_bytes, 1), 1ULL << 22));
    cudaStream_t st[NS];
    for (int i = 0; i < NS; ++i) XS_CUDA(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
    int rc = XS_OK;
    uint64_t i0 = 0; int bi = 0;
    while (rc == XS_OK && i0 < n_seq) {
        HostBatchPlan pl = plan_batch(seq_begin, seq_end, n_seq, i0, MAX_SPAN, max_seq);
        if (pl.hi > n_bases) { rc = fail(XS_ERR_ARG, "sequence offsets exceed n_bases"); break; }
        cudaStream_t s = st[bi % NS];
        uint64_t ns = pl.i1 - pl.i0, span = pl.hi - pl.lo;
        uint8_t* d = nullptr;
        size_t o_b = 0, o_e = align256(span + 64), o_o = o_e + 2 * align256(ns * 8);
        size_t total = o_o + align256(ns * out_row_bytes + 4);
        cudaError_t e = cudaMallocAsync((void**)&d, total, s);
        if (e != cudaSuccess) { rc = fail(XS_ERR_NOMEM, std::string("batch buffers: ") + cudaGetErrorString(e)); break; }
        uint64_t* d_b = reinterpret_cast<uint64_t*>(d + o_e);
        uint64_t* d_e = reinterpret_cast<uint64_t*>(d + o_e + align256(ns * 8));
        if (span) e = cudaMemcpyAsync(d + o_b, bases + pl.lo, span, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_b, seq_begin + pl.i0, ns * 8, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_e, seq_end + pl.i0, ns * 8, cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) { rc = fail(XS_ERR_CUDA, std::string("H2D copy: ") + cudaGetErrorString(e)); cudaFreeAsync(d, s); break; }
        rc = query(d + o_b, span, d_b, d_e, ns, pl.lo, d + o_o, pl.i0, s);
        if (rc != XS_OK) { cudaFreeAsync(d, s); break; }
        e = cudaFreeAsync(d, s);
        if (e != cudaSuccess) { rc = fail(XS_ERR_CUDA, std::string("batch buffers: ") + cudaGetErrorString(e)); break; }
        i0 = pl.i1; ++bi;
    }
    for (int i = 0; i < NS; ++i) {
        cudaError_t e = cudaStreamSynchronize(st[i]);
        if (rc == XS_OK && e != cudaSuccess) rc = fail(XS_ERR_CUDA, std::string("query failed: ") + cudaGetErrorString(e));
        cudaStreamDestroy(st[i]);
    }
    return rc;
}

// ----------------------------------------------------------------------------------------
// C ABI
// ----------------------------------------------------------------------------------------
extern "C" {

int xs_version(void) { return 100; }
const char* xs_last_error(void) { return g_err.c_str(); }
uint64_t xs_launch_count(void) { return g_launches.load(); }

int xs_profile_enable(int on) {
    g_profile.store(on ? 1 : 0);
    return XS_OK;
}
static int profile_collect(double* ms, uint64_t* launches) {   // [PROF_TAGS] each
    std::vector<ProfEvent> ev;
    {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        ev.swap(g_prof_events);
    }
    for (int t = 0; t < PROF_TAGS; ++t) { ms[t] = 0; launches[t] = 0; }
    int rc = XS_OK;
    for (auto& pr : ev) {
        float t = 0;
        cudaError_t e = cudaEventSynchronize(pr.b);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&t, pr.a, pr.b);
        if (e != cudaSuccess) rc = fail(XS_ERR_CUDA, std::string("profile events: ") + cudaGetErrorString(e));
        ms[pr.tag] += t; launches[pr.tag] += 1;
        cudaEventDestroy(pr.a); cudaEventDestroy(pr.b);
    }
    return rc;
}
int xs_profile_read(double* kernel_ms, uint64_t* launches) {
    double ms[PROF_TAGS]; uint64_t n[PROF_TAGS];
    int rc = profile_collect(ms, n);
    if (kernel_ms) { *kernel_ms = 0; for (int t = 0; t < PROF_TAGS; ++t) *kernel_ms += ms[t]; }
    if (launches) { *launches = 0; for (int t = 0; t < PROF_TAGS; ++t) *launches += n[t]; }
    return rc;
}
int xs_profile_read_phases(double* kernel_ms, uint64_t* launches) {
    if (!kernel_ms || !launches) return fail(XS_ERR_ARG, "NULL argument");
    return profile_collect(kernel_ms, launches);
}

int xs_device_count(int* n) {
    if (!n) return fail(XS_ERR_ARG, "n is NULL");
    cudaError_t e = cudaGetDeviceCount(n);
    if (e != cudaSuccess) { *n = 0; return fail(XS_ERR_CUDA, cudaGetErrorString(e)); }
    return XS_OK;
}

// device whose context page-locked host memory is allocated under: the first device a handle was opened on, else
// XSPECT_B200_DEVICE / LOCAL_RANK, else 0 — so that a rank of a multi-GPU job never creates a context on GPU 0 just
// to pin memory.  The allocation is portable (usable from every context).
static int home_device() {
    int d = g_home_device.load(std::memory_order_relaxed);
    if (d >= 0) return d;
    const char* v = getenv("XSPECT_B200_DEVICE");
    if (!v || !*v) v = getenv("LOCAL_RANK");
    return (v && *v) ? atoi(v) : 0;
}

int xs_host_alloc(uint64_t bytes, void** out) {
    if (!out) return fail(XS_ERR_ARG, "out is NULL");
    DeviceGuard guard(home_device());
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? XS_ERR_NOMEM : XS_ERR_CUDA, cudaGetErrorString(e));
    return XS_OK;
}
int xs_device_trim(int device) {
    DeviceGuard guard(device);
    if (!guard.ok) return fail(XS_ERR_CUDA, "cannot select the device");
    XS_CUDA(cudaDeviceSynchronize());
    {   // the streaming file reader's cached staging buffers
        std::lock_guard<std::mutex> lk(g_stream_mu);
        if (g_stream_dev == device) stream_slots_release();
    }
    cudaMemPool_t pool;
    XS_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    XS_CUDA(cudaMemPoolTrimTo(pool, 0));
    return XS_OK;
}
int xs_host_free(void* p) {
    if (p) XS_CUDA(cudaFreeHost(p));
    return XS_OK;
}

// 2-D tensor map over 16-byte rows for cp.async.bulk.tensor tile::gather4 (k_bucket_fetch_tma); false when the driver
// entry point is missing or the geometry does not fit (rows >= 2^31)
static bool make_row_tensor_map(const uint8_t* d_rows, uint64_t n_rows, TensorMap128* out) {
    typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static_assert(sizeof(CUtensorMap) == sizeof(TensorMap128), "CUtensorMap is 128 bytes");
    if (n_rows == 0 || n_rows >= (1ULL << 31)) return false;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) { cudaGetLastError(); return false; }
    cuuint64_t gdim[2] = {4, n_rows};
    cuuint64_t gstr[1] = {16};
    cuuint32_t box[2] = {4, 1};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMap tm;
    CUresult r = ((EncodeTiled)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void*)d_rows, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    memcpy(out, &tm, sizeof(tm));
    return true;
}

int xs_cobs_open(const char* path, int device, uint32_t doc_begin, uint32_t doc_end, xs_cobs** out) {
    if (!path || !out) return fail(XS_ERR_ARG, "path/out is NULL");
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) return fail(XS_ERR_IO, std::string(path) + ": " + strerror(errno));
    CobsFile cf;
    const char* layout = "";
    int rc = parse_cobs_header(f, path, cf, &layout);
    if (rc != XS_OK) { fclose(f); return rc; }
    if (doc_begin == 0 && doc_end == 0) doc_end = cf.n_docs;
    if (doc_end > cf.n_docs || doc_begin >= doc_end || (doc_begin % 8) != 0 || (doc_end % 8 != 0 && doc_end != cf.n_docs)) {
        fclose(f);
        return fail(XS_ERR_ARG, "document shard must be [multiple of 8, multiple of 8 or n_docs) within the index");
    }
    if (cf.kind == XS_COBS_COMPACT && !(doc_begin == 0 && doc_end == cf.n_docs)) {
        fclose(f);
        return fail(XS_ERR_UNSUPPORTED, "document-column shards are supported for classic indices only");
    }
    DeviceGuard guard(device);
    int n_sm = 0;
    rc = device_setup(device, &n_sm);
    if (rc != XS_OK) { fclose(f); return rc; }

    xs_cobs* ix = new xs_cobs();
    ix->info.device = device;
    ix->n_sm = n_sm;
    ix->names = cf.names;
    ix->layout = layout;
    const char* fw = getenv("XS_FORCE_WIDE");
    ix->force_wide = (fw && fw[0] == '1') ? 1 : 0;
    if (const char* v = getenv("XS_BUCKETED")) ix->bk.enabled = v[0] != '0';
    if (const char* v = getenv("XS_BUCKET_MIN_WINDOWS")) ix->bk.min_windows = std::max<uint64_t>(1, strtoull(v, nullptr, 10));
    if (const char* v = getenv("XS_BUCKET_SCRATCH_MB")) ix->bk.scratch_bytes = strtoull(v, nullptr, 10) << 20;
    uint32_t col0 = doc_begin / 8;
    uint32_t n_col = cf.kind == XS_COBS_CLASSIC ? (doc_end - doc_begin + 7) / 8 : (uint32_t)cf.page_bytes;
    // HBM row stride: a row never straddles a 128-byte DRAM fetch (power of two up to 128 B, then multiples of 128 B)
    uint32_t stride = 16;
    while (stride < n_col && stride < 128) stride <<= 1;
    if (n_col > 128) stride = (n_col + 127) & ~127u;
    ix->narrow = stride == 16;
    // compact indices (several pages of narrow rows) are scored by k_cobs_pages; their rows of <= 8 bytes sit at an
    // 8-byte stride (half the L2 footprint: one MLST locus of 600 alleles is 34 MB instead of 67 MB)
    const char* no_pages = getenv("XS_NO_PAGES_KERNEL");
    ix->pages_kernel = ix->narrow && !ix->force_wide && cf.kind == XS_COBS_COMPACT && cf.n_pages > 1 &&
                       pages_smem(cf.n_pages) <= PAGES_SMEM_MAX && !(no_pages && no_pages[0] == '1');
    if (ix->pages_kernel && n_col <= 8) stride = 8;
    const uint32_t n_chunks = (n_col + 15) / 16;     // 16-byte chunks that hold documents

    uint64_t total = 0;
    std::vector<uint64_t> offs;
    for (uint64_t s : cf.sig) { offs.push_back(total); total += s * stride; }
    cudaError_t e = cudaMalloc((void**)&ix->d_data, total + 256);
    if (e != cudaSuccess) {
        fclose(f); delete ix;
        return fail(XS_ERR_NOMEM, "index of " + std::to_string(total) + " bytes does not fit in HBM: " + cudaGetErrorString(e));
    }
    uint64_t file_off = cf.data_off;
    for (uint32_t pg = 0; pg < cf.n_pages && rc == XS_OK; ++pg) {
        rc = upload_rows(f, file_off, cf.sig[pg], (uint32_t)cf.page_bytes, col0, n_col, ix->d_data + offs[pg], stride, n_sm);
        file_off += cf.sig[pg] * cf.page_bytes;
        PageDesc pd{};
        pd.data = ix->d_data + offs[pg];
        pd.sig_size = cf.sig[pg];
        pd.magic = magic_of(cf.sig[pg]);
        pd.row_stride = stride;
        if (cf.kind == XS_COBS_CLASSIC) { pd.n_docs = doc_end - doc_begin; pd.doc_off = 0; }
        else {
            uint64_t d0 = (uint64_t)pg * 8 * cf.page_bytes;
            pd.n_docs = d0 >= cf.n_docs ? 0 : (uint32_t)std::min<uint64_t>(8 * cf.page_bytes, cf.n_docs - d0);
            pd.doc_off = (uint32_t)d0;
        }
        ix->pages.push_back(pd);
        uint32_t C = n_chunks;
        uint32_t nb = (C + WIDE_MAX_COLS - 1) / WIDE_MAX_COLS;
        uint32_t per = (C + nb - 1) / nb;
        for (uint32_t c0 = 0; c0 < C; c0 += per) {
            ColBlock cb{pg, c0, std::min(per, C - c0), 0};
            uint64_t dlo = (uint64_t)c0 * 128, dhi = std::min<uint64_t>((uint64_t)(c0 + cb.n_cols) * 128, pd.n_docs);
            cb.n_docs = dhi > dlo ? (uint32_t)(dhi - dlo) : 0;
            if (cb.n_docs) ix->blocks.push_back(cb);
        }
    }
    fclose(f);
    if (rc == XS_OK && (ix->pages.size() > N_COUNTERS || ix->blocks.size() > N_COUNTERS))
        rc = fail(XS_ERR_UNSUPPORTED, "more than 4096 pages / column blocks in one index");
    if (rc == XS_OK) {
        if (cudaMalloc((void**)&ix->d_pages, ix->pages.size() * sizeof(PageDesc)) != cudaSuccess ||
            cudaMalloc((void**)&ix->d_blocks, std::max<size_t>(1, ix->blocks.size()) * sizeof(ColBlock)) != cudaSuccess)
            rc = fail(XS_ERR_NOMEM, "descriptor allocation failed");
    }
    if (rc == XS_OK) {
        cudaMemcpy(ix->d_pages, ix->pages.data(), ix->pages.size() * sizeof(PageDesc), cudaMemcpyHostToDevice);
        if (!ix->blocks.empty())
            cudaMemcpy(ix->d_blocks, ix->blocks.data(), ix->blocks.size() * sizeof(ColBlock), cudaMemcpyHostToDevice);
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) rc = fail(XS_ERR_CUDA, std::string("index upload: ") + cudaGetErrorString(e));
    }
    if (rc != XS_OK) {
        std::string keep = g_err;
        xs_cobs_close(ix);
        g_err = keep;
        return rc;
    }
    xs_cobs_info_t& in = ix->info;
    in.kind = (uint32_t)cf.kind; in.term_size = cf.k; in.canonicalize = cf.canonicalize;
    in.num_hashes = (uint32_t)cf.num_hashes; in.n_docs_total = cf.n_docs;
    in.doc_begin = doc_begin; in.doc_end = doc_end; in.n_pages = cf.n_pages;
    in.page_bytes = cf.page_bytes; in.row_stride = stride;
    in.sig_size_max = *std::max_element(cf.sig.begin(), cf.sig.end());
    in.hbm_bytes = total; in.device = device; in.policy = XS_NONACGT_SKIP;
    ix->mid = cf.kind == XS_COBS_CLASSIC && mid_eligible(ix);
    if (ix->narrow && ix->pages.size() == 1 && ix->pages[0].row_stride == 16)
        ix->has_tmap = make_row_tensor_map(ix->pages[0].data, ix->pages[0].sig_size, &ix->tmap);
    *out = ix;
    return XS_OK;
}

// A classic index whose rows come from the counter-based generator of k_synth_rows instead of a file (BASELINE
// config 5: D = 10 000, S = 96 000 000 is a 120 GB index; the bench generates each rank's column shard straight into
// HBM and the oracle regenerates the rows its parity sample touches).  Same handle, same kernels as xs_cobs_open.
int xs_cobs_create_synthetic(int device, uint32_t n_docs, uint32_t doc_begin, uint32_t doc_end, uint64_t sig_size,
                             uint32_t term_size, uint32_t num_hashes, uint64_t seed, xs_cobs** out) {
    if (!out) return fail(XS_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (doc_begin == 0 && doc_end == 0) doc_end = n_docs;
    if (n_docs == 0 || doc_end > n_docs || doc_begin >= doc_end || (doc_begin % 32) != 0 || (doc_end % 8 != 0 && doc_end != n_docs))
        return fail(XS_ERR_ARG, "document shard must be [multiple of 32, multiple of 8 or n_docs) within the index");
    if (sig_size == 0 || term_size == 0 || term_size > 32 || num_hashes == 0 || num_hashes > 64)
        return fail(XS_ERR_ARG, "bad synthetic index geometry");
    DeviceGuard guard(device);
    int n_sm = 0;
    XS_TRY(device_setup(device, &n_sm));
    xs_cobs* ix = new xs_cobs();
    ix->info.device = device;
    ix->n_sm = n_sm;
    ix->layout = "synthetic rows (k_synth_rows)";
    for (uint32_t d = 0; d < n_docs; ++d) { ix->names += "d" + std::to_string(d); ix->names.push_back('\n'); }
    const uint32_t col0 = doc_begin / 8, n_col = (doc_end - doc_begin + 7) / 8;
    uint32_t stride = 16;
    while (stride < n_col && stride < 128) stride <<= 1;
    if (n_col > 128) stride = (n_col + 127) & ~127u;
    ix->narrow = stride == 16;
    const uint64_t total = sig_size * stride;
    cudaError_t e = cudaMalloc((void**)&ix->d_data, total + 256);
    if (e != cudaSuccess) { delete ix; return fail(XS_ERR_NOMEM, "synthetic index of " + std::to_string(total) + " bytes does not fit in HBM: " + cudaGetErrorString(e)); }
    k_synth_rows<<<n_sm * 16, 256>>>(ix->d_data, sig_size, stride, col0, n_col, n_docs, seed);
    int rc = launch_ok("k_synth_rows");
    PageDesc pd{};
    pd.data = ix->d_data; pd.sig_size = sig_size; pd.magic = magic_of(sig_size); pd.row_stride = stride;
    pd.n_docs = doc_end - doc_begin; pd.doc_off = 0;
    ix->pages.push_back(pd);
    const uint32_t C = (n_col + 15) / 16, nb = (C + WIDE_MAX_COLS - 1) / WIDE_MAX_COLS, per = (C + nb - 1) / nb;
    for (uint32_t c0 = 0; c0 < C; c0 += per) {
        ColBlock cb{0, c0, std::min(per, C - c0), 0};
        const uint64_t dlo = (uint64_t)c0 * 128, dhi = std::min<uint64_t>((uint64_t)(c0 + cb.n_cols) * 128, pd.n_docs);
        cb.n_docs = dhi > dlo ? (uint32_t)(dhi - dlo) : 0;
        if (cb.n_docs) ix->blocks.push_back(cb);
    }
    if (rc == XS_OK && (cudaMalloc((void**)&ix->d_pages, sizeof(PageDesc)) != cudaSuccess ||
                        cudaMalloc((void**)&ix->d_blocks, std::max<size_t>(1, ix->blocks.size()) * sizeof(ColBlock)) != cudaSuccess))
        rc = fail(XS_ERR_NOMEM, "descriptor allocation failed");
    if (rc == XS_OK) {
        cudaMemcpy(ix->d_pages, ix->pages.data(), sizeof(PageDesc), cudaMemcpyHostToDevice);
        if (!ix->blocks.empty()) cudaMemcpy(ix->d_blocks, ix->blocks.data(), ix->blocks.size() * sizeof(ColBlock), cudaMemcpyHostToDevice);
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) rc = fail(XS_ERR_CUDA, std::string("synthetic index: ") + cudaGetErrorString(e));
    }
    if (rc != XS_OK) { std::string keep = g_err; xs_cobs_close(ix); g_err = keep; return rc; }
    xs_cobs_info_t& in = ix->info;
    in.kind = XS_COBS_CLASSIC; in.term_size = term_size; in.canonicalize = 1; in.num_hashes = num_hashes; in.n_docs_total = n_docs;
    in.doc_begin = doc_begin; in.doc_end = doc_end; in.n_pages = 1; in.page_bytes = ((uint64_t)n_docs + 7) / 8; in.row_stride = stride;
    in.sig_size_max = sig_size; in.hbm_bytes = total; in.device = device; in.policy = XS_NONACGT_SKIP;
    ix->mid = mid_eligible(ix);
    *out = ix;
    return XS_OK;
}

// ----------------------------------------------------------------------------------------
// NCCL exchange of the document-column sharded path.  libnccl is bound at run time (dlopen): the
// library has no link-time dependency on it, single-GPU users never load it, and inside a torch process the NCCL
// torch already loaded is the one used.
// ----------------------------------------------------------------------------------------
#include <dlfcn.h>
namespace {
typedef struct { char internal[128]; } nccl_uid;
typedef void* nccl_comm;
struct NcclApi {
    int (*GetUniqueId)(nccl_uid*) = nullptr;
    int (*CommInitRank)(nccl_comm*, int, nccl_uid, int) = nullptr;
    int (*CommDestroy)(nccl_comm) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    void* handle = nullptr;
    std::string err;
};
NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* env = getenv("XSPECT_B200_NCCL");
        void* h = nullptr;
        if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);    // already in the process (torch)
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { const char* m = dlerror(); api.err = std::string("libnccl.so.2 cannot be loaded: ") + (m ? m : "?"); return; }
        api.handle = h;
        auto sym = [&](const char* n) { void* p = dlsym(h, n); if (!p && api.err.empty()) api.err = std::string("NCCL symbol missing: ") + n; return p; };
        api.GetUniqueId = (int (*)(nccl_uid*))sym("ncclGetUniqueId");
        api.CommInitRank = (int (*)(nccl_comm*, int, nccl_uid, int))sym("ncclCommInitRank");
        api.CommDestroy = (int (*)(nccl_comm))sym("ncclCommDestroy");
        api.AllGather = (int (*)(const void*, void*, size_t, int, nccl_comm, cudaStream_t))sym("ncclAllGather");
        api.AllReduce = (int (*)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t))sym("ncclAllReduce");
        api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
        api.GetVersion = (int (*)(int*))sym("ncclGetVersion");
    });
    return &api;
}
}  // namespace

struct xs_comm {
    nccl_comm comm = nullptr;
    int rank = 0, world = 1, device = 0;
};

static int nccl_fail(NcclApi* a, int r, const char* what) {
    return fail(XS_ERR_NCCL, std::string(what) + ": " + (a->GetErrorString ? a->GetErrorString(r) : "NCCL error"));
}

int xs_comm_unique_id(uint8_t* id128) {
    if (!id128) return fail(XS_ERR_ARG, "NULL argument");
    NcclApi* a = nccl_api();
    if (!a->err.empty()) return fail(XS_ERR_NCCL, a->err);
    nccl_uid id;
    int r = a->GetUniqueId(&id);
    if (r != 0) return nccl_fail(a, r, "ncclGetUniqueId");
    memcpy(id128, id.internal, 128);
    return XS_OK;
}

int xs_comm_init(const uint8_t* id128, int rank, int world, int device, xs_comm** out) {
    if (!id128 || !out) return fail(XS_ERR_ARG, "NULL argument");
    *out = nullptr;
    if (world < 1 || world > 16 || rank < 0 || rank >= world) return fail(XS_ERR_ARG, "rank/world out of range (world <= 16)");
    NcclApi* a = nccl_api();
    if (!a->err.empty()) return fail(XS_ERR_NCCL, a->err);
    DeviceGuard guard(device);
    if (!guard.ok) return fail(XS_ERR_CUDA, "cannot select the device");
    nccl_uid id;
    memcpy(id.internal, id128, 128);
    xs_comm* c = new xs_comm();
    c->rank = rank; c->world = world; c->device = device;
    int r = a->CommInitRank(&c->comm, world, id, rank);
    if (r != 0) { delete c; return nccl_fail(a, r, "ncclCommInitRank"); }
    *out = c;
    return XS_OK;
}

int xs_comm_destroy(xs_comm* c) {
    if (!c) return XS_OK;
    NcclApi* a = nccl_api();
    if (c->comm && a->CommDestroy) { DeviceGuard guard(c->device); a->CommDestroy(c->comm); }
    delete c;
    return XS_OK;
}

int xs_comm_info(const xs_comm* c, int* rank, int* world, int* nccl_version) {
    if (!c) return fail(XS_ERR_ARG, "NULL communicator");
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    if (nccl_version) { *nccl_version = 0; NcclApi* a = nccl_api(); if (a->GetVersion) a->GetVersion(nccl_version); }
    return XS_OK;
}

int xs_allgather_scores(xs_comm* c, const void* d_local, uint64_t n_seq, uint64_t row_bytes, void* d_all, void* stream) {
    if (!c || (n_seq && row_bytes && (!d_local || !d_all))) return fail(XS_ERR_ARG, "NULL argument");
    if (n_seq == 0 || row_bytes == 0) return XS_OK;
    NcclApi* a = nccl_api();
    DeviceGuard guard(c->device);
    int r = a->AllGather(d_local, d_all, (size_t)(n_seq * row_bytes), /*ncclUint8*/ 1, c->comm, (cudaStream_t)stream);
    if (r != 0) return nccl_fail(a, r, "ncclAllGather");
    return XS_OK;
}

int xs_allreduce_totals(xs_comm* c, uint64_t* d_totals, uint64_t n, void* stream) {
    if (!c || (n && !d_totals)) return fail(XS_ERR_ARG, "NULL argument");
    if (n == 0) return XS_OK;
    NcclApi* a = nccl_api();
    DeviceGuard guard(c->device);
    int r = a->AllReduce(d_totals, d_totals, (size_t)n, /*ncclUint64*/ 5, /*ncclSum*/ 0, c->comm, (cudaStream_t)stream);
    if (r != 0) return nccl_fail(a, r, "ncclAllReduce");
    return XS_OK;
}

int xs_sharded_reduce_device(const void* d_all, uint64_t n_seq, int dtype, int device, uint32_t world, uint32_t w,
                             const uint32_t* widths, uint32_t* d_best, uint32_t* d_best_count, uint32_t* d_n_best,
                             uint64_t* d_totals, void* stream) {
    if (!dtype_size(dtype)) return fail(XS_ERR_ARG, "dtype must be XS_U8, XS_U16 or XS_U32");
    if (world == 0 || world > 16 || !widths) return fail(XS_ERR_ARG, "world must be 1..16");
    if (n_seq == 0) return XS_OK;
    if (!d_all) return fail(XS_ERR_ARG, "NULL count buffer");
    ShardLayout lay{};
    lay.world = world; lay.w = w;
    uint32_t d0 = 0;
    for (uint32_t g = 0; g < world; ++g) {
        if (widths[g] > w) return fail(XS_ERR_ARG, "a shard is wider than the padded block");
        lay.width[g] = widths[g]; lay.doc0[g] = d0; d0 += widths[g];
    }
    DeviceGuard guard(device);
    if (!guard.ok) return fail(XS_ERR_CUDA, "cannot select the device");
    int n_sm = 0;
    XS_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n_seq + 7) / 8, (uint64_t)n_sm * 8));
    unsigned long long* tot = reinterpret_cast<unsigned long long*>(d_totals);
    if (dtype == XS_U8) k_sharded_reduce<uint8_t><<<grid, 256, 0, s>>>((const uint8_t*)d_all, n_seq, lay, d_best, d_best_count, d_n_best, tot);
    else if (dtype == XS_U16) k_sharded_reduce<uint16_t><<<grid, 256, 0, s>>>((const uint16_t*)d_all, n_seq, lay, d_best, d_best_count, d_n_best, tot);
    else k_sharded_reduce<uint32_t><<<grid, 256, 0, s>>>((const uint32_t*)d_all, n_seq, lay, d_best, d_best_count, d_n_best, tot);
    return launch_ok("k_sharded_reduce");
}

int xs_cobs_info(const xs_cobs* ix, xs_cobs_info_t* info) {
    if (!ix || !info) return fail(XS_ERR_ARG, "NULL argument");
    *info = ix->info;
    return XS_OK;
}

const char* xs_cobs_header_layout(const xs_cobs* ix) { return ix ? ix->layout.c_str() : ""; }

const char* xs_cobs_kernel(const xs_cobs* ix) {
    if (!ix) return "";
    if (uses_wide(ix)) return "k_cobs_wide";
    if (ix->pages_kernel) return "k_cobs_pages";
    return ix->mid ? "k_cobs_mid" : "k_cobs_narrow";
}

int xs_cobs_probe_header(const char* path, xs_cobs_header_t* out) {
    if (!path || !out) return fail(XS_ERR_ARG, "path/out is NULL");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(XS_ERR_IO, std::string(path) + ": " + strerror(errno));
    CobsFile cf;
    const char* layout = "";
    const int rc = parse_cobs_header(f, path, cf, &layout);
    fclose(f);
    if (rc != XS_OK) return rc;
    memset(out, 0, sizeof(*out));
    out->kind = (uint32_t)cf.kind; out->term_size = cf.k; out->canonicalize = cf.canonicalize;
    out->num_hashes = (uint32_t)cf.num_hashes; out->n_docs = cf.n_docs; out->n_pages = cf.n_pages;
    out->page_bytes = cf.page_bytes; out->sig_size_max = *std::max_element(cf.sig.begin(), cf.sig.end());
    out->data_offset = cf.data_off; out->file_size = cf.file_size;
    strncpy(out->layout, layout, sizeof(out->layout) - 1);
    return XS_OK;
}

int xs_cobs_doc_fill(const xs_cobs* ix, uint64_t sample_rows, double* fill) {
    if (!ix || !fill) return fail(XS_ERR_ARG, "NULL argument");
    if (sample_rows == 0) sample_rows = 1 << 16;
    DeviceGuard guard(ix->info.device);
    const uint32_t n_local = ix->info.doc_end - ix->info.doc_begin;
    uint32_t* d_cnt = nullptr;
    XS_CUDA(cudaMalloc((void**)&d_cnt, (size_t)n_local * 4));
    cudaError_t e = cudaMemset(d_cnt, 0, (size_t)n_local * 4);
    std::vector<uint64_t> n_of(n_local, 1);
    for (const PageDesc& pg : ix->pages) {
        if (e != cudaSuccess || pg.n_docs == 0) continue;
        const uint64_t n = std::min<uint64_t>(sample_rows, pg.sig_size);
        k_doc_fill<<<ix->n_sm * 4, 256>>>(pg, n, d_cnt);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        e = cudaGetLastError();
        for (uint32_t d = 0; d < pg.n_docs; ++d) n_of[pg.doc_off + d] = n;
    }
    std::vector<uint32_t> h(n_local);
    if (e == cudaSuccess) e = cudaMemcpy(h.data(), d_cnt, (size_t)n_local * 4, cudaMemcpyDeviceToHost);
    cudaFree(d_cnt);
    if (e != cudaSuccess) return fail(XS_ERR_CUDA, std::string("xs_cobs_doc_fill: ") + cudaGetErrorString(e));
    for (uint32_t d = 0; d < n_local; ++d) fill[d] = (double)h[d] / (double)n_of[d];
    return XS_OK;
}

int xs_cobs_doc_names(const xs_cobs* ix, char* buf, uint64_t cap, uint64_t* needed) {
    if (!ix) return fail(XS_ERR_ARG, "NULL index");
    if (needed) *needed = ix->names.size();
    if (buf && cap) memcpy(buf, ix->names.data(), std::min<uint64_t>(cap, ix->names.size()));
    return XS_OK;
}

int xs_cobs_set_policy(xs_cobs* ix, int policy) {
    if (!ix) return fail(XS_ERR_ARG, "NULL index");
    if (policy != XS_NONACGT_SKIP && policy != XS_NONACGT_LITERAL) return fail(XS_ERR_ARG, "unknown policy");
    ix->info.policy = policy;
    return XS_OK;
}

int xs_cobs_set_bucketed(xs_cobs* ix, int enabled, uint64_t min_windows, uint64_t scratch_bytes, uint32_t bucket_shift) {
    if (!ix) return fail(XS_ERR_ARG, "NULL index");
    if (bucket_shift > BK_MAX_SHIFT) return fail(XS_ERR_ARG, "bucket_shift must be <= 21");
    ix->bk.configure(enabled, min_windows, scratch_bytes, bucket_shift);
    return XS_OK;
}

int xs_cobs_bucketed_queries(const xs_cobs* ix, uint64_t* n) {
    if (!ix || !n) return fail(XS_ERR_ARG, "NULL argument");
    *n = ix->bk.queries.load();
    return XS_OK;
}

int xs_cobs_close(xs_cobs* ix) {
    if (!ix) return XS_OK;
    DeviceGuard guard(ix->info.device);
    if (ix->d_data) cudaFree(ix->d_data);
    if (ix->d_pages) cudaFree(ix->d_pages);
    if (ix->d_blocks) cudaFree(ix->d_blocks);
    ix->bk.destroy();
    delete ix;
    return XS_OK;
}

int xs_cobs_query_device(xs_cobs* ix, const uint8_t* d_bases, uint64_t n_bases, const uint64_t* d_seq_begin,
                         const uint64_t* d_seq_end, uint64_t n_seq, uint32_t step, int out_dtype, void* d_out,
                         void* stream) {
    if (!ix || (!d_out && n_seq) || (n_seq && (!d_seq_begin || !d_seq_end))) return fail(XS_ERR_ARG, "NULL argument");
    if (step == 0) return fail(XS_ERR_ARG, "step must be >= 1");
    if (!dtype_size(out_dtype)) return fail(XS_ERR_ARG, "out_dtype must be XS_U8, XS_U16 or XS_U32");
    DeviceGuard guard(ix->info.device);
    if (!guard.ok) return fail(XS_ERR_CUDA, "cannot select the index's device");
    return cobs_query_dev(ix, d_bases, n_bases, d_seq_begin, d_seq_end, n_seq, 0, step, out_dtype, d_out,
                          (cudaStream_t)stream);
}

int xs_cobs_query_device_ld(xs_cobs* ix, const uint8_t* d_bases, uint64_t n_bases, const uint64_t* d_seq_begin,
                            const uint64_t* d_seq_end, uint64_t n_seq, uint32_t step, int out_dtype, uint64_t ld, void* d_out,
                            void* stream) {
    if (!ix || (!d_out && n_seq) || (n_seq && (!d_seq_begin || !d_seq_end))) return fail(XS_ERR_ARG, "NULL argument");
    if (step == 0) return fail(XS_ERR_ARG, "step must be >= 1");
    if (!dtype_size(out_dtype)) return fail(XS_ERR_ARG, "out_dtype must be XS_U8, XS_U16 or XS_U32");
    if (ld < ix->info.doc_end - ix->info.doc_begin) return fail(XS_ERR_ARG, "ld is smaller than the number of local documents");
    DeviceGuard guard(ix->info.device);
    if (!guard.ok) return fail(XS_ERR_CUDA, "cannot select the index's device");
    return cobs_query_dev(ix, d_bases, n_bases, d_seq_begin, d_seq_end, n_seq, 0, step, out_dtype, d_out, (cudaStream_t)stream, ld);
}

int xs_cobs_query(xs_cobs* ix, const uint8_t* bases, uint64_t n_bases, const uint64_t* seq_begin,
                  const uint64_t* seq_end, uint64_t n_seq, uint32_t step, int out_dtype, void* out) {
    if (!ix || (!out && n_seq) || (n_seq && (!seq_begin || !seq_end)) || (n_bases && !bases))
        return fail(XS_ERR_ARG, "NULL argument");
    if (step == 0) return fail(XS_ERR_ARG, "step must be >= 1");
    if (!dtype_size(out_dtype)) return fail(XS_ERR_ARG, "out_dtype must be XS_U8, XS_U16 or XS_U32");
    DeviceGuard guard(ix->info.device);
    if (!guard.ok) return fail(XS_ERR_CUDA, "cannot select the index's device");
    const uint64_t ld = ix->info.doc_end - ix->info.doc_begin;
    const uint64_t row = ld * (uint64_t)out_dtype;
    {
        bool handled = false;
        const bool wide = uses_wide(ix);
        const uint32_t n_counters = (uint32_t)std::max(ix->pages.size(), ix->blocks.size());
        XS_TRY(small_query(ix->info.device, ix->n_sm, ix->info.term_size, step, wide ? WIDE_CHUNK : 0, n_counters, bases, n_bases,
                           seq_begin, seq_end, n_seq, row, out, &handled,
                           [&](const SeqBatch& sb, const uint64_t* chunk_prefix, uint8_t* d_o, cudaStream_t s) {
                               return cobs_launch(ix, sb, chunk_prefix, out_dtype, d_o, s);
                           }));
        if (handled) return XS_OK;
    }
    return host_pipeline(bases, n_bases, seq_begin, seq_end, n_seq, row,
                         [&](const uint8_t* db, uint64_t span, const uint64_t* d_b, const uint64_t* d_e, uint64_t ns,
                             uint64_t shift, uint8_t* d_o, uint64_t i0, cudaStream_t s) {
                             XS_TRY(cobs_query_dev(ix, db, span, d_b, d_e, ns, shift, step, out_dtype, d_o, s));
                             XS_CUDA(cudaMemcpyAsync((uint8_t*)out + i0 * row, d_o, ns * row, cudaMemcpyDeviceToHost, s));
                             return (int)XS_OK;
                         });
}

int xs_cobs_classify(xs_cobs* ix, const uint8_t* bases, uint64_t n_bases, const uint64_t* seq_begin,
                     const uint64_t* seq_end, uint64_t n_seq, uint32_t step, uint32_t* best, uint32_t* best_count,
                     uint32_t* n_best, uint64_t* totals) {
    if (!ix || (n_seq && (!seq_begin || !seq_end || !best || !best_count || !n_best)) || (n_bases && !bases) || !totals)
        return fail(XS_ERR_ARG, "NULL argument");
    if (step == 0) return fail(XS_ERR_ARG, "step must be >= 1");
    DeviceGuard guard(ix->info.device);
    if (!guard.ok) return fail(XS_ERR_CUDA, "cannot select the index's device");
    const uint64_t ld = ix->info.doc_end - ix->info.doc_begin;
    // counts stay on the device as uint32 (no saturation whatever the record length); only 12 bytes per record
    // and the per-document totals come back
    uint64_t* d_tot = nullptr;
    XS_CUDA(cudaMalloc((void**)&d_tot, ld * 8));
    {   // the pipeline's streams are non-blocking (they do not order against the legacy stream): make the zeroing
        // visible to all of them before any is created
        cudaError_t e = cudaMemset(d_tot, 0, ld * 8);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            cudaFree(d_tot);
            return fail(XS_ERR_CUDA, std::string("totals clear: ") + cudaGetErrorString(e));
        }
    }
    const uint64_t row = ld * 4 + 12;
    int rc = host_pipeline(bases, n_bases, seq_begin, seq_end, n_seq, row,
                           [&](const uint8_t* db, uint64_t span, const uint64_t* d_b, const uint64_t* d_e, uint64_t ns,
                               uint64_t shift, uint8_t* d_o, uint64_t i0, cudaStream_t s) {
                               uint32_t* d_best = reinterpret_cast<uint32_t*>(d_o + ns * ld * 4);
                               uint32_t* d_cnt = d_best + ns;
                               uint32_t* d_nb = d_cnt + ns;
                               XS_TRY(cobs_query_dev(ix, db, span, d_b, d_e, ns, shift, step, XS_U32, d_o, s));
                               unsigned grid = (unsigned)std::min<uint64_t>((ns + REDUCE_ROWS - 1) / REDUCE_ROWS, (uint64_t)ix->n_sm * 8);
                               k_scores_reduce<uint32_t><<<grid, REDUCE_NT, 0, s>>>((const uint32_t*)d_o, ns, (uint32_t)ld, d_best, d_cnt, d_nb,
                                                                                    reinterpret_cast<unsigned long long*>(d_tot));
                               XS_TRY(launch_ok("k_scores_reduce"));
                               XS_CUDA(cudaMemcpyAsync(best + i0, d_best, ns * 4, cudaMemcpyDeviceToHost, s));
                               XS_CUDA(cudaMemcpyAsync(best_count + i0, d_cnt, ns * 4, cudaMemcpyDeviceToHost, s));
                               XS_CUDA(cudaMemcpyAsync(n_best + i0, d_nb, ns * 4, cudaMemcpyDeviceToHost, s));
                               return (int)XS_OK;
                           });
    if (rc == XS_OK) {
        cudaError_t e = cudaMemcpy(totals, d_tot, ld * 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(XS_ERR_CUDA, std::string("totals copy: ") + cudaGetErrorString(e));
    }
    cudaFree(d_tot);
    return rc;
}

// ----------------------------------------------------------------------------------------
// File -> read-level calls, streamed: blocks of the FASTA / FASTQ file are parsed by all host threads straight into
// page-locked staging buffers while the previous blocks are copied, scored and reduced on the device.
// ----------------------------------------------------------------------------------------

struct xs_file_calls {
    RawBuf<uint32_t> best, best_count, n_best;
    RawBuf<uint64_t> seq_len, id_end;
    RawBuf<char> ids;
    std::vector<uint64_t> totals;
    uint64_t n_bases = 0, n_short = 0, n_blocks = 0;
    double parse_s = 0, total_s = 0;
};

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int xs_cobs_classify_file(xs_cobs* ix, const char* path, int format, uint32_t step, uint64_t block_bytes, xs_file_calls** out) {
    if (!ix || !path || !out) return fail(XS_ERR_ARG, "NULL argument");
    *out = nullptr;
    if (step == 0) return fail(XS_ERR_ARG, "step must be >= 1");
    xs_fastx* fx = nullptr;
    XS_TRY(xs_fastx_open_stream(path, format, &fx));
    DeviceGuard guard(ix->info.device);
    if (!guard.ok) { xs_fastx_close(fx); return fail(XS_ERR_CUDA, "cannot select the index's device"); }
    const double t_start = now_s();
    const uint64_t fsize = xs_fastx_file_size(fx);
    if (block_bytes == 0) block_bytes = 96ULL << 20;            // ~300 k 150-bp FASTQ reads: enough windows for the bucketed kernels
    const unsigned hw = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    const uint64_t ld = ix->info.doc_end - ix->info.doc_begin;
    const uint32_t k = ix->info.term_size;
    xs_file_calls* res = new xs_file_calls();
    res->totals.assign(ld, 0);
    // staging buffers, streams and events are kept between calls (page-locking ~100 MB per slot costs more than
    // parsing a block); calls on one process take turns
    std::lock_guard<std::mutex> stream_lock(g_stream_mu);
    if (g_stream_dev != ix->info.device) { stream_slots_release(); g_stream_dev = ix->info.device; }
    const int NS = STREAM_NS;
    StreamSlot* slot = g_stream_slot;
    for (int i = 0; i < NS; ++i) { slot[i].busy = false; slot[i].rec0 = slot[i].n_rec = 0; }
    uint64_t* d_tot = nullptr;
    int rc = XS_OK;
    auto cuda_ok = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && rc == XS_OK) rc = fail(e == cudaErrorMemoryAllocation ? XS_ERR_NOMEM : XS_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
        return e == cudaSuccess;
    };
    cuda_ok(cudaMalloc((void**)&d_tot, std::max<uint64_t>(ld, 1) * 8), "totals");
    if (rc == XS_OK) cuda_ok(cudaMemset(d_tot, 0, ld * 8), "totals");
    if (rc == XS_OK) cuda_ok(cudaDeviceSynchronize(), "totals");
    for (int i = 0; i < NS && rc == XS_OK; ++i) {
        if (!slot[i].s) cuda_ok(cudaStreamCreateWithFlags(&slot[i].s, cudaStreamNonBlocking), "stream");
        if (!slot[i].done) cuda_ok(cudaEventCreateWithFlags(&slot[i].done, cudaEventDisableTiming), "event");
    }
    auto retire = [&](StreamSlot& sl) {       // results of the slot's block -> the result vectors
        if (!sl.busy) return;
        cuda_ok(cudaEventSynchronize(sl.done), "block");
        if (rc == XS_OK) {
            memcpy(res->best.data() + sl.rec0, sl.h_res, sl.n_rec * 4);
            memcpy(res->best_count.data() + sl.rec0, sl.h_res + sl.n_rec, sl.n_rec * 4);
            memcpy(res->n_best.data() + sl.rec0, sl.h_res + 2 * sl.n_rec, sl.n_rec * 4);
        }
        sl.busy = false;
    };
    std::vector<FastxSegOut> segs;
    uint64_t a = rc == XS_OK ? xs_fastx_sync(fx, 0) : fsize;
    if (rc == XS_OK && format == 2 && !xs_fastx_blank(fx, 0, a))
        rc = fail(XS_ERR_FORMAT, "the streaming reader accepts 4-line FASTQ only (wrapped or malformed records found)");
    uint64_t rec_total = 0;
    int bi = 0;
    while (rc == XS_OK && a < fsize) {
        uint64_t b = a + block_bytes >= fsize ? fsize : xs_fastx_sync(fx, a + block_bytes);
        if (b <= a) b = fsize;
        StreamSlot& sl = slot[bi % NS];
        retire(sl);
        if (rc != XS_OK) break;
        const double tp = now_s();
        // staging: the block's bytes as an upper bound of its bases (every parser thread fills its own region), then
        // begin | end, sized after the parse
        const uint64_t span = b - a;
        const size_t o_b = align256(span + 64);
        const uint64_t rec_cap = span / (format == 2 ? 6 : 2) + 2;
        size_t need_h = o_b + 2 * align256(std::min<uint64_t>(rec_cap, span / 64 + 4096) * 8);     // typical; regrown below if short
        auto grow_h = [&](size_t need) {
            if (need <= sl.h_cap) return true;
            uint8_t* nh = nullptr;
            if (!cuda_ok(cudaHostAlloc((void**)&nh, need + need / 8, cudaHostAllocPortable), "pinned staging")) return false;
            if (sl.h) cudaFreeHost(sl.h);
            sl.h = nh; sl.h_cap = need + need / 8;
            return true;
        };
        if (!grow_h(need_h)) break;
        rc = xs_fastx_parse_block_1pass(fx, a, b, hw, sl.h, segs, k);
        if (rc != XS_OK) break;
        uint64_t nr = 0, nb = 0, nid = 0;
        for (const FastxSegOut& so : segs) { nr += so.n_rec; nb += so.n_bases; nid += so.n_id; }
        if (bi == 0 && b < fsize) {      // size the result arrays once from the first block's record density
            const double scale = (double)fsize / (double)(b - a) * 1.03;
            const size_t er = (size_t)(nr * scale) + 4096, ei = (size_t)(nid * scale) + 65536;
            res->best.reserve(er); res->best_count.reserve(er); res->n_best.reserve(er);
            res->seq_len.reserve(er); res->id_end.reserve(er); res->ids.reserve(ei);
        }
        need_h = o_b + 2 * align256(nr * 8);
        if (need_h > sl.h_cap) {       // more records than the typical bound: keep the parsed bases, move to a larger buffer
            uint8_t* old = sl.h; sl.h = nullptr; const size_t old_cap = sl.h_cap; sl.h_cap = 0;
            if (!grow_h(need_h)) { cudaFreeHost(old); break; }
            memcpy(sl.h, old, std::min<size_t>(old_cap, o_b));
            cudaFreeHost(old);
        }
        if (nr * 12 > sl.r_cap) {
            if (sl.h_res) cudaFreeHost(sl.h_res);
            sl.h_res = nullptr; sl.r_cap = 0;
            if (!cuda_ok(cudaHostAlloc((void**)&sl.h_res, nr * 12 + nr, cudaHostAllocPortable), "pinned results")) break;
            sl.r_cap = nr * 12 + nr;
        }
        uint64_t* h_b = reinterpret_cast<uint64_t*>(sl.h + o_b);
        uint64_t* h_e = reinterpret_cast<uint64_t*>(sl.h + o_b + align256(nr * 8));
        // compaction of the per-thread record arrays (16 bytes per record) + ids
        const size_t id0 = res->ids.size();
        if (!res->ids.grow(id0 + nid) || !res->id_end.grow(rec_total + nr) || !res->seq_len.grow(rec_total + nr) ||
            !res->best.grow(rec_total + nr) || !res->best_count.grow(rec_total + nr) || !res->n_best.grow(rec_total + nr)) {
            rc = fail(XS_ERR_NOMEM, "result arrays");
            break;
        }
        uint64_t max_len = 0, n_short = 0, r0 = 0, i0 = 0;
        for (const FastxSegOut& so : segs) {
            if (!so.n_rec) continue;
            memcpy(h_b + r0, so.begin, so.n_rec * 8);
            memcpy(h_e + r0, so.end, so.n_rec * 8);
            memcpy(res->ids.data() + id0 + i0, so.ids, so.n_id);
            uint64_t* sl_out = res->seq_len.data() + rec_total + r0;
            uint64_t* ie_out = res->id_end.data() + rec_total + r0;
            const uint64_t id_shift = id0 + i0;
            for (uint64_t i = 0; i < so.n_rec; ++i) { sl_out[i] = so.end[i] - so.begin[i]; ie_out[i] = id_shift + so.id_end[i]; }
            max_len = std::max(max_len, so.max_len);
            n_short += so.n_short;
            r0 += so.n_rec; i0 += so.n_id;
        }
        res->n_short += n_short;
        res->n_bases += nb;
        res->parse_s += now_s() - tp;
        // device side of the block
        const uint64_t max_win = max_len >= k ? (max_len - k) / step + 1 : 0;
        const int dt = max_win <= 255 ? XS_U8 : max_win <= 65535 ? XS_U16 : XS_U32;
        const size_t o_cnt = need_h, o_res = o_cnt + align256(nr * ld * (uint64_t)dt + 16), need_d = o_res + align256(nr * 12 + 16);
        if (need_d > sl.d_cap) {
            if (sl.d) cudaFree(sl.d);
            sl.d = nullptr; sl.d_cap = 0;
            if (!cuda_ok(cudaMalloc((void**)&sl.d, need_d + need_d / 8), "device block buffers")) break;
            sl.d_cap = need_d + need_d / 8;
        }
        if (nr) {
            for (const FastxSegOut& so : segs)      // only the filled part of every parser region travels
                if (so.n_bases) cuda_ok(cudaMemcpyAsync(sl.d + so.base0, sl.h + so.base0, so.n_bases, cudaMemcpyHostToDevice, sl.s), "H2D copy");
            cuda_ok(cudaMemcpyAsync(sl.d + o_b, sl.h + o_b, 2 * align256(nr * 8), cudaMemcpyHostToDevice, sl.s), "H2D copy");
            uint64_t* d_b = reinterpret_cast<uint64_t*>(sl.d + o_b);
            uint64_t* d_e = reinterpret_cast<uint64_t*>(sl.d + o_b + align256(nr * 8));
            uint32_t* d_best = reinterpret_cast<uint32_t*>(sl.d + o_res);
            if (rc == XS_OK) rc = cobs_query_dev(ix, sl.d, span, d_b, d_e, nr, 0, step, dt, sl.d + o_cnt, sl.s);
            if (rc == XS_OK) {
                const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((nr + REDUCE_ROWS - 1) / REDUCE_ROWS, (uint64_t)ix->n_sm * 8));
                unsigned long long* tot = reinterpret_cast<unsigned long long*>(d_tot);
                if (dt == XS_U8) k_scores_reduce<uint8_t><<<grid, REDUCE_NT, 0, sl.s>>>((const uint8_t*)(sl.d + o_cnt), nr, (uint32_t)ld, d_best, d_best + nr, d_best + 2 * nr, tot);
                else if (dt == XS_U16) k_scores_reduce<uint16_t><<<grid, REDUCE_NT, 0, sl.s>>>((const uint16_t*)(sl.d + o_cnt), nr, (uint32_t)ld, d_best, d_best + nr, d_best + 2 * nr, tot);
                else k_scores_reduce<uint32_t><<<grid, REDUCE_NT, 0, sl.s>>>((const uint32_t*)(sl.d + o_cnt), nr, (uint32_t)ld, d_best, d_best + nr, d_best + 2 * nr, tot);
                rc = launch_ok("k_scores_reduce");
            }
            if (rc == XS_OK) cuda_ok(cudaMemcpyAsync(sl.h_res, d_best, nr * 12, cudaMemcpyDeviceToHost, sl.s), "D2H copy");
            if (rc == XS_OK) cuda_ok(cudaEventRecord(sl.done, sl.s), "event");
            sl.rec0 = rec_total; sl.n_rec = nr; sl.busy = rc == XS_OK;
        }
        rec_total += nr;
        a = b; ++bi; ++res->n_blocks;
    }
    xs_fastx_seg_free(segs);
    for (int i = 0; i < NS; ++i) retire(slot[(bi + i) % NS]);
    if (rc == XS_OK && d_tot) {
        cuda_ok(cudaDeviceSynchronize(), "file query");
        if (rc == XS_OK) cuda_ok(cudaMemcpy(res->totals.data(), d_tot, ld * 8, cudaMemcpyDeviceToHost), "totals copy");
    } else {
        cudaDeviceSynchronize();
    }
    if (d_tot) cudaFree(d_tot);
    xs_fastx_close(fx);
    res->total_s = now_s() - t_start;
    if (rc != XS_OK) { delete res; return rc; }
    *out = res;
    return XS_OK;
}

int xs_file_calls_info(const xs_file_calls* r, uint64_t* n_records, uint64_t* n_bases, uint64_t* n_id_bytes, uint64_t* n_short,
                       uint64_t* n_docs, double* parse_s, double* total_s) {
    if (!r) return fail(XS_ERR_ARG, "NULL result");
    if (n_records) *n_records = r->best.size();
    if (n_bases) *n_bases = r->n_bases;
    if (n_id_bytes) *n_id_bytes = r->ids.size();
    if (n_short) *n_short = r->n_short;
    if (n_docs) *n_docs = r->totals.size();
    if (parse_s) *parse_s = r->parse_s;
    if (total_s) *total_s = r->total_s;
    return XS_OK;
}

int xs_file_calls_read(const xs_file_calls* r, uint32_t* best, uint32_t* best_count, uint32_t* n_best, uint64_t* seq_len,
                       char* ids, uint64_t* id_end, uint64_t* totals) {
    if (!r) return fail(XS_ERR_ARG, "NULL result");
    const size_t n = r->best.size();
    if (best) memcpy(best, r->best.data(), n * 4);
    if (best_count) memcpy(best_count, r->best_count.data(), n * 4);
    if (n_best) memcpy(n_best, r->n_best.data(), n * 4);
    if (seq_len) memcpy(seq_len, r->seq_len.data(), n * 8);
    if (ids) memcpy(ids, r->ids.data(), r->ids.size());
    if (id_end) memcpy(id_end, r->id_end.data(), n * 8);
    if (totals) memcpy(totals, r->totals.data(), r->totals.size() * 8);
    return XS_OK;
}

int xs_file_calls_view(const xs_file_calls* r, const uint32_t** best, const uint32_t** best_count, const uint32_t** n_best,
                       const uint64_t** seq_len, const char** ids, const uint64_t** id_end, const uint64_t** totals) {
    if (!r) return fail(XS_ERR_ARG, "NULL result");
    if (best) *best = r->best.data();
    if (best_count) *best_count = r->best_count.data();
    if (n_best) *n_best = r->n_best.data();
    if (seq_len) *seq_len = r->seq_len.data();
    if (ids) *ids = r->ids.data();
    if (id_end) *id_end = r->id_end.data();
    if (totals) *totals = r->totals.data();
    return XS_OK;
}

int xs_file_calls_free(xs_file_calls* r) { delete r; return XS_OK; }

// ----------------------------------------------------------------------------------------
// MLST: chunked scoring of every locus of a scheme (probabilistic_filter_mlst_model.py:236-286)
// ----------------------------------------------------------------------------------------
struct MlstSegs {                 // the sequence_splitter chunks of all records for one locus, as byte segments
    std::vector<uint64_t> begin, end, rec_seg0;
    std::vector<uint8_t> always;  // segment of an unchunked record: kept whatever it scores
    uint64_t max_windows = 0;
};

// sequence_splitter (:382-426) as segments of the record [b, e).  The one non-contiguous chunk (a remainder shorter
// than k glued to the last chunk, repeating the k-1 overlap bases) is materialised in `extra`, which is appended
// to the device copy of the bases at offset extra_base.
static int mlst_segments(const uint8_t* bases, const uint64_t* sb, const uint64_t* se, uint64_t n_seq, uint32_t k, uint32_t step,
                         uint32_t allele_len, uint64_t chunk_from, uint64_t extra_base, std::vector<uint8_t>& extra, MlstSegs& out) {
    out.rec_seg0.assign(1, 0);
    for (uint64_t i = 0; i < n_seq; ++i) {
        const uint64_t b = sb[i], e = se[i], n = e - b;
        auto push = [&](uint64_t x, uint64_t y, uint8_t always) {
            out.begin.push_back(x); out.end.push_back(y); out.always.push_back(always);
            const uint64_t len = y - x;
            if (len >= k) out.max_windows = std::max(out.max_windows, (len - k) / step + 1);
        };
        if (n < chunk_from) {
            push(b, e, 1);
        } else {
            const uint64_t sub = n < 1000000 ? (uint64_t)allele_len : n < 10000000 ? (uint64_t)allele_len * 10 : (uint64_t)allele_len * 100;
            if (sub < k) return fail(XS_ERR_ARG, "MLST chunk length (average allele length) must be at least k");
            const uint64_t stride = sub - k + 1;
            uint64_t start = 0;
            const size_t first = out.begin.size();
            while (start + sub <= n) { push(b + start, b + start + sub, 0); start += stride; }
            if (start < n) {
                if (n - start < k) {
                    if (out.begin.size() == first) return fail(XS_ERR_ARG, "MLST record shorter than k");
                    // last chunk += remainder: bytes [last, last + sub) followed by [start, n)
                    const uint64_t last = out.begin.back();
                    const uint64_t x = extra_base + extra.size();
                    extra.insert(extra.end(), bases + last, bases + last + sub);
                    extra.insert(extra.end(), bases + b + start, bases + e);
                    out.begin.back() = x; out.end.back() = x + sub + (n - start);
                    out.max_windows = std::max(out.max_windows, (sub + (n - start) - k) / step + 1);
                } else {
                    push(b + start, e, 0);
                }
            }
        }
        out.rec_seg0.push_back(out.begin.size());
    }
    return XS_OK;
}

// result-dict order of one record and locus from its kept rows (chunk order): first appearance while walking each
// row in cobs result order, then a stable sort by descending sum (:244-256); an unchunked record returns its single
// row in cobs result order, zeros included (:272-286)
static void mlst_order(const uint32_t* rows, uint32_t n_rows, uint32_t n_docs, bool chunked, uint32_t thr,
                       std::vector<uint32_t>& scratch, uint32_t* out_doc, uint32_t* out_score, uint32_t* out_n) {
    std::vector<uint32_t>& idx = scratch;
    idx.resize(n_docs);
    if (!chunked) {
        std::iota(idx.begin(), idx.end(), 0u);
        const uint32_t* sc = rows;
        std::partial_sort(idx.begin(), idx.end(), idx.end(), [sc](uint32_t a, uint32_t b) { return sc[a] > sc[b]; });
        for (uint32_t j = 0; j < n_docs; ++j) { out_doc[j] = idx[j]; out_score[j] = sc[idx[j]]; }
        *out_n = n_docs;
        return;
    }
    std::vector<uint64_t> sum(n_docs, 0);
    std::vector<uint32_t> seen;
    for (uint32_t r = 0; r < n_rows; ++r) {
        const uint32_t* sc = rows + (uint64_t)r * n_docs;
        std::iota(idx.begin(), idx.end(), 0u);
        std::partial_sort(idx.begin(), idx.end(), idx.end(), [sc](uint32_t a, uint32_t b) { return sc[a] > sc[b]; });
        for (uint32_t j = 0; j < n_docs; ++j) {
            const uint32_t d = idx[j], v = sc[d];
            if (v <= thr) break;             // descending: nothing above the threshold follows
            if (sum[d] == 0) seen.push_back(d);
            sum[d] += v;
        }
    }
    std::stable_sort(seen.begin(), seen.end(), [&](uint32_t a, uint32_t b) { return sum[a] > sum[b]; });
    for (size_t j = 0; j < seen.size(); ++j) { out_doc[j] = seen[j]; out_score[j] = (uint32_t)std::min<uint64_t>(sum[seen[j]], 0xFFFFFFFFu); }
    *out_n = (uint32_t)seen.size();
}

int xs_mlst_query(xs_cobs* const* loci, uint32_t n_loci, const uint32_t* allele_len, const uint8_t* bases, uint64_t n_bases,
                  const uint64_t* seq_begin, const uint64_t* seq_end, uint64_t n_seq, uint32_t step, uint32_t min_chunk_score,
                  uint64_t chunk_from_len, uint32_t* out_n, uint32_t* out_doc, uint32_t* out_score) {
    if (!loci || !n_loci || !allele_len || (n_seq && (!seq_begin || !seq_end || !out_n || !out_doc || !out_score)) || (n_bases && !bases))
        return fail(XS_ERR_ARG, "NULL argument");
    if (step == 0) return fail(XS_ERR_ARG, "step must be >= 1");
    const int device = loci[0]->info.device;
    const uint32_t k = loci[0]->info.term_size;
    for (uint32_t l = 0; l < n_loci; ++l) {
        if (!loci[l]) return fail(XS_ERR_ARG, "NULL locus index");
        if (loci[l]->info.device != device || loci[l]->info.term_size != k)
            return fail(XS_ERR_ARG, "all loci of a scheme must share the device and the k-mer length");
    }
    for (uint64_t i = 0; i < n_seq; ++i)
        if (seq_end[i] < seq_begin[i] || seq_end[i] > n_bases) return fail(XS_ERR_ARG, "sequence offsets out of range");
    if (n_seq == 0) return XS_OK;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(XS_ERR_CUDA, "cannot select the index's device");

    // chunk segments of every locus; glued last chunks are appended behind the bases
    std::vector<MlstSegs> segs(n_loci);
    std::vector<uint8_t> extra;
    for (uint32_t l = 0; l < n_loci; ++l)
        XS_TRY(mlst_segments(bases, seq_begin, seq_end, n_seq, k, step, allele_len[l], chunk_from_len, n_bases, extra, segs[l]));
    const uint64_t total_bases = n_bases + extra.size();

    const uint32_t CAP = 16;        // hot chunks kept per record; more (repeats of a locus) take the fallback below
    struct LocusBuf {
        uint64_t* d_seg = nullptr; uint8_t* d_always = nullptr; void* d_counts = nullptr; uint8_t* d_hot = nullptr;
        uint32_t *d_nhot = nullptr, *d_hseg = nullptr, *d_hrows = nullptr;
        std::vector<uint32_t> nhot, hseg, hrows;
        int dt = XS_U16; uint32_t n_docs = 0; uint64_t n_seg = 0;
    };
    std::vector<LocusBuf> lb(n_loci);
    cudaStream_t s = nullptr;
    uint8_t* d_bases = nullptr;
    int rc = XS_OK;
    auto cuda_ok = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && rc == XS_OK) rc = fail(e == cudaErrorMemoryAllocation ? XS_ERR_NOMEM : XS_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
        return e == cudaSuccess;
    };
    cuda_ok(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "stream");
    if (rc == XS_OK && cuda_ok(cudaMallocAsync((void**)&d_bases, total_bases + 64, s), "bases allocation")) {
        cuda_ok(cudaMemcpyAsync(d_bases, bases, n_bases, cudaMemcpyHostToDevice, s), "bases upload");
        if (!extra.empty()) cuda_ok(cudaMemcpyAsync(d_bases + n_bases, extra.data(), extra.size(), cudaMemcpyHostToDevice, s), "bases upload");
    }
    for (uint32_t l = 0; l < n_loci && rc == XS_OK; ++l) {
        xs_cobs* ix = loci[l];
        LocusBuf& b = lb[l];
        const MlstSegs& sg = segs[l];
        b.n_docs = ix->info.doc_end - ix->info.doc_begin;
        b.n_seg = sg.begin.size();
        b.dt = sg.max_windows <= 65535 ? XS_U16 : XS_U32;
        const uint64_t nsg = b.n_seg;
        if (!cuda_ok(cudaMallocAsync((void**)&b.d_seg, (2 * nsg + n_seq + 1) * 8, s), "segment allocation")) break;
        if (!cuda_ok(cudaMallocAsync((void**)&b.d_always, 2 * nsg + 16, s), "segment allocation")) break;
        b.d_hot = b.d_always + nsg;
        if (!cuda_ok(cudaMallocAsync(&b.d_counts, std::max<uint64_t>(1, nsg * b.n_docs * (uint64_t)b.dt), s), "count matrix allocation")) break;
        if (!cuda_ok(cudaMallocAsync((void**)&b.d_nhot, (n_seq + n_seq * CAP + n_seq * CAP * (uint64_t)b.n_docs) * 4, s), "hot row allocation")) break;
        b.d_hseg = b.d_nhot + n_seq;
        b.d_hrows = b.d_hseg + n_seq * CAP;
        cuda_ok(cudaMemcpyAsync(b.d_seg, sg.begin.data(), nsg * 8, cudaMemcpyHostToDevice, s), "segment upload");
        cuda_ok(cudaMemcpyAsync(b.d_seg + nsg, sg.end.data(), nsg * 8, cudaMemcpyHostToDevice, s), "segment upload");
        cuda_ok(cudaMemcpyAsync(b.d_seg + 2 * nsg, sg.rec_seg0.data(), (n_seq + 1) * 8, cudaMemcpyHostToDevice, s), "segment upload");
        cuda_ok(cudaMemcpyAsync(b.d_always, sg.always.data(), nsg, cudaMemcpyHostToDevice, s), "segment upload");
        if (rc != XS_OK) break;
        rc = cobs_query_dev(ix, d_bases, total_bases, b.d_seg, b.d_seg + nsg, nsg, 0, step, b.dt, b.d_counts, s);
        if (rc != XS_OK) break;
        const unsigned hgrid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((nsg + 7) / 8, (uint64_t)ix->n_sm * 8));
        if (b.dt == XS_U16) {
            k_mlst_hot<uint16_t><<<hgrid, 256, 0, s>>>((const uint16_t*)b.d_counts, nsg, b.n_docs, min_chunk_score, b.d_always, b.d_hot);
            k_mlst_compact<uint16_t><<<(unsigned)n_seq, 256, CAP * 4, s>>>((const uint16_t*)b.d_counts, b.d_seg + 2 * nsg, b.n_docs, b.d_hot, CAP, b.d_nhot, b.d_hseg, b.d_hrows);
        } else {
            k_mlst_hot<uint32_t><<<hgrid, 256, 0, s>>>((const uint32_t*)b.d_counts, nsg, b.n_docs, min_chunk_score, b.d_always, b.d_hot);
            k_mlst_compact<uint32_t><<<(unsigned)n_seq, 256, CAP * 4, s>>>((const uint32_t*)b.d_counts, b.d_seg + 2 * nsg, b.n_docs, b.d_hot, CAP, b.d_nhot, b.d_hseg, b.d_hrows);
        }
        g_launches.fetch_add(1, std::memory_order_relaxed);
        rc = launch_ok("k_mlst_hot / k_mlst_compact");
        if (rc != XS_OK) break;
        b.nhot.resize(n_seq); b.hseg.resize(n_seq * CAP); b.hrows.resize(n_seq * CAP * (uint64_t)b.n_docs);
        cuda_ok(cudaMemcpyAsync(b.nhot.data(), b.d_nhot, n_seq * 4, cudaMemcpyDeviceToHost, s), "result copy");
        cuda_ok(cudaMemcpyAsync(b.hseg.data(), b.d_hseg, n_seq * CAP * 4, cudaMemcpyDeviceToHost, s), "result copy");
        cuda_ok(cudaMemcpyAsync(b.hrows.data(), b.d_hrows, n_seq * CAP * (uint64_t)b.n_docs * 4, cudaMemcpyDeviceToHost, s), "result copy");
    }
    if (s) cuda_ok(cudaStreamSynchronize(s), "MLST query");

    // ordering on the host from the few kept rows (libstdc++ partial_sort = cobs' own result order)
    if (rc == XS_OK) {
        std::vector<uint32_t> scratch, full;
        uint64_t col0 = 0;
        for (uint32_t l = 0; l < n_loci && rc == XS_OK; ++l) {
            LocusBuf& b = lb[l];
            const MlstSegs& sg = segs[l];
            for (uint64_t i = 0; i < n_seq && rc == XS_OK; ++i) {
                const bool chunked = (seq_end[i] - seq_begin[i]) >= chunk_from_len;
                uint32_t* od = out_doc + col0 + i * b.n_docs;
                uint32_t* os = out_score + col0 + i * b.n_docs;
                const uint32_t nh = b.nhot[i];
                const uint32_t* rows = b.hrows.data() + i * CAP * (uint64_t)b.n_docs;
                uint32_t n_rows = nh;
                if (nh > CAP) {
                    // more hot chunks than kept: fetch this record's hot rows from the count matrix still on the device
                    const uint64_t g0 = sg.rec_seg0[i], g1 = sg.rec_seg0[i + 1];
                    std::vector<uint8_t> hot(g1 - g0);
                    if (!cuda_ok(cudaMemcpy(hot.data(), b.d_hot + g0, g1 - g0, cudaMemcpyDeviceToHost), "hot flags copy")) break;
                    full.clear();
                    std::vector<uint8_t> raw((size_t)b.n_docs * b.dt);
                    for (uint64_t g = g0; g < g1; ++g) {
                        if (!hot[g - g0]) continue;
                        if (!cuda_ok(cudaMemcpy(raw.data(), (const uint8_t*)b.d_counts + g * b.n_docs * (uint64_t)b.dt, raw.size(), cudaMemcpyDeviceToHost), "row copy")) break;
                        for (uint32_t d = 0; d < b.n_docs; ++d)
                            full.push_back(b.dt == XS_U16 ? (uint32_t)reinterpret_cast<const uint16_t*>(raw.data())[d] : reinterpret_cast<const uint32_t*>(raw.data())[d]);
                    }
                    rows = full.data();
                    n_rows = (uint32_t)(full.size() / b.n_docs);
                }
                mlst_order(rows, n_rows, b.n_docs, chunked, min_chunk_score, scratch, od, os, out_n + (uint64_t)l * n_seq + i);
            }
            col0 += n_seq * b.n_docs;
        }
    }
    for (LocusBuf& b : lb) {
        if (b.d_seg) cudaFreeAsync(b.d_seg, s);
        if (b.d_always) cudaFreeAsync(b.d_always, s);
        if (b.d_counts) cudaFreeAsync(b.d_counts, s);
        if (b.d_nhot) cudaFreeAsync(b.d_nhot, s);
    }
    if (d_bases) cudaFreeAsync(d_bases, s);
    if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
    return rc;
}

int xs_cobs_result_order(const uint32_t* scores, uint32_t n_docs, uint32_t* order) {
    if ((!scores || !order) && n_docs) return fail(XS_ERR_ARG, "NULL argument");
    std::vector<uint32_t> idx(n_docs);
    std::iota(idx.begin(), idx.end(), 0u);
    std::partial_sort(idx.begin(), idx.end(), idx.end(), [&](uint32_t a, uint32_t b) { return scores[a] > scores[b]; });
    std::copy(idx.begin(), idx.end(), order);
    return XS_OK;
}

int xs_cobs_result_order_batch(const uint32_t* scores, uint64_t n_seq, uint32_t n_docs, uint32_t* order) {
    if ((!scores || !order) && n_seq && n_docs) return fail(XS_ERR_ARG, "NULL argument");
    for (uint64_t i = 0; i < n_seq; ++i) {
        const uint32_t* sc = scores + i * n_docs;
        uint32_t* o = order + i * n_docs;
        std::iota(o, o + n_docs, 0u);
        std::partial_sort(o, o + n_docs, o + n_docs, [sc](uint32_t a, uint32_t b) { return sc[a] > sc[b]; });
    }
    return XS_OK;
}

int xs_scores_reduce_device(const void* d_counts, uint64_t n_seq, uint32_t n_docs, int dtype, int device,
                            uint32_t* d_best, uint32_t* d_best_count, uint32_t* d_n_best, uint64_t* d_totals,
                            void* stream) {
    if (!dtype_size(dtype)) return fail(XS_ERR_ARG, "dtype must be XS_U8, XS_U16 or XS_U32");
    if (n_seq == 0 || n_docs == 0) return XS_OK;
    if (!d_counts) return fail(XS_ERR_ARG, "NULL count matrix");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(XS_ERR_CUDA, "cannot select the device");
    int n_sm = 0;
    XS_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
    cudaStream_t s = (cudaStream_t)stream;
    unsigned grid = (unsigned)std::min<uint64_t>((n_seq + REDUCE_ROWS - 1) / REDUCE_ROWS, (uint64_t)n_sm * 8);
    unsigned long long* tot = reinterpret_cast<unsigned long long*>(d_totals);
    if (dtype == XS_U8) k_scores_reduce<uint8_t><<<grid, REDUCE_NT, 0, s>>>((const uint8_t*)d_counts, n_seq, n_docs, d_best, d_best_count, d_n_best, tot);
    else if (dtype == XS_U16) k_scores_reduce<uint16_t><<<grid, REDUCE_NT, 0, s>>>((const uint16_t*)d_counts, n_seq, n_docs, d_best, d_best_count, d_n_best, tot);
    else k_scores_reduce<uint32_t><<<grid, REDUCE_NT, 0, s>>>((const uint32_t*)d_counts, n_seq, n_docs, d_best, d_best_count, d_n_best, tot);
    return launch_ok("k_scores_reduce");
}

int xs_bloom_open(const char* path, uint32_t term_size, int device, xs_bloom** out) {
    if (!path || !out) return fail(XS_ERR_ARG, "path/out is NULL");
    *out = nullptr;
    if (term_size == 0) return fail(XS_ERR_ARG, "term_size 0");
    if (term_size > 32) return fail(XS_ERR_UNSUPPORTED, "term_size > 32 is not supported");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(XS_ERR_IO, std::string(path) + ": " + strerror(errno));
    uint64_t kh = 0;
    if (!rd(f, &kh)) { fclose(f); return fail(XS_ERR_FORMAT, std::string(path) + ": not an rbloom file (too small)"); }
    fseek(f, 0, SEEK_END);
    uint64_t size = (uint64_t)ftell(f);
    if (size <= 8) { fclose(f); return fail(XS_ERR_FORMAT, std::string(path) + ": rbloom file has no bit array"); }
    if (kh > 4096) { fclose(f); return fail(XS_ERR_FORMAT, std::string(path) + ": implausible number of hash functions"); }
    uint64_t nbytes = size - 8;
    DeviceGuard guard(device);
    int n_sm = 0;
    int rc = device_setup(device, &n_sm);
    if (rc != XS_OK) { fclose(f); return rc; }
    xs_bloom* bf = new xs_bloom();
    bf->n_sm = n_sm;
    if (const char* v = getenv("XS_BUCKETED")) bf->bk.enabled = v[0] != '0';
    cudaError_t e = cudaMalloc((void**)&bf->d_bits, nbytes + 256);
    if (e != cudaSuccess) { fclose(f); delete bf; return fail(XS_ERR_NOMEM, std::string("bloom bit array: ") + cudaGetErrorString(e)); }
    // upload in slices through the row uploader (1 "row" = 1 MiB, remainder separately)
    const uint32_t ROW = 1u << 20;
    uint64_t rows = nbytes / ROW, rem = nbytes % ROW;
    if (rows) rc = upload_rows(f, 8, rows, ROW, 0, ROW, bf->d_bits, ROW, n_sm);
    if (rc == XS_OK && rem) rc = upload_rows(f, 8 + rows * ROW, 1, (uint32_t)rem, 0, (uint32_t)rem, bf->d_bits + rows * ROW, (uint32_t)rem, n_sm);
    fclose(f);
    if (rc != XS_OK) { cudaFree(bf->d_bits); delete bf; return rc; }
    bf->info.n_bits = nbytes * 8; bf->info.k_hashes = kh; bf->info.term_size = term_size;
    bf->info.device = device; bf->info.hbm_bytes = nbytes;
    *out = bf;
    return XS_OK;
}

int xs_bloom_info(const xs_bloom* bf, xs_bloom_info_t* info) {
    if (!bf || !info) return fail(XS_ERR_ARG, "NULL argument");
    *info = bf->info;
    return XS_OK;
}

int xs_bloom_set_bucketed(xs_bloom* bf, int enabled, uint64_t min_windows, uint64_t scratch_bytes, uint32_t bucket_shift,
                          int member_pct) {
    if (!bf) return fail(XS_ERR_ARG, "NULL filter");
    if (bucket_shift > 31 || (bucket_shift && bucket_shift < 3)) return fail(XS_ERR_ARG, "bucket_shift must be 0 or in 3..31");
    if (member_pct > 100) return fail(XS_ERR_ARG, "member_pct must be <= 100");
    if (member_pct >= 0) bf->bucket_member_pct = (uint32_t)member_pct;
    bf->bk.configure(enabled, min_windows, scratch_bytes, bucket_shift);
    return XS_OK;
}

int xs_bloom_bucketed_queries(const xs_bloom* bf, uint64_t* n) {
    if (!bf || !n) return fail(XS_ERR_ARG, "NULL argument");
    *n = bf->bk.queries.load();
    return XS_OK;
}

int xs_bloom_close(xs_bloom* bf) {
    if (!bf) return XS_OK;
    DeviceGuard guard(bf->info.device);
    bf->bk.destroy();
    if (bf->d_bits) cudaFree(bf->d_bits);
    delete bf;
    return XS_OK;
}

int xs_bloom_contains(xs_bloom* bf, const uint8_t* terms, uint64_t n_terms, uint8_t* out) {
    if (!bf || (n_terms && (!terms || !out))) return fail(XS_ERR_ARG, "NULL argument");
    DeviceGuard guard(bf->info.device);
    if (!guard.ok) return fail(XS_ERR_CUDA, "cannot select the filter's device");
    const uint32_t k = bf->info.term_size;
    // every term is one sequence of exactly k bytes = one window, hashed without canonicalisation
    const uint64_t CH = SMALL_MAX_SEQ;
    std::vector<uint64_t> b(std::min(n_terms, CH)), e(b.size());
    std::vector<uint32_t> hits(b.size());
    for (uint64_t i0 = 0; i0 < n_terms; i0 += CH) {
        uint64_t n = std::min(CH, n_terms - i0);
        for (uint64_t i = 0; i < n; ++i) { b[i] = (i0 + i) * k; e[i] = b[i] + k; }
        bool handled = false;
        XS_TRY(small_query(bf->info.device, bf->n_sm, k, 1, 0, 1, terms, n_terms * k, b.data(), e.data(), n, 4, hits.data(), &handled,
                           [&](const SeqBatch& sb, const uint64_t*, uint8_t* d_o, cudaStream_t s) {
                               return bloom_launch(bf, sb, reinterpret_cast<uint32_t*>(d_o), s, true);
                           }));
        if (!handled) return fail(XS_ERR_NOMEM, "membership batch could not be staged");
        for (uint64_t i = 0; i < n; ++i) out[i0 + i] = (uint8_t)(hits[i] != 0);
    }
    return XS_OK;
}

int xs_bloom_query_device(xs_bloom* bf, const uint8_t* d_bases, uint64_t n_bases, const uint64_t* d_seq_begin,
                          const uint64_t* d_seq_end, uint64_t n_seq, uint32_t step, uint32_t* d_out_hits, void* stream) {
    if (!bf || (!d_out_hits && n_seq) || (n_seq && (!d_seq_begin || !d_seq_end))) return fail(XS_ERR_ARG, "NULL argument");
    if (step == 0) return fail(XS_ERR_ARG, "step must be >= 1");
    DeviceGuard guard(bf->info.device);
    if (!guard.ok) return fail(XS_ERR_CUDA, "cannot select the filter's device");
    return bloom_query_dev(bf, d_bases, n_bases, d_seq_begin, d_seq_end, n_seq, 0, step, d_out_hits, (cudaStream_t)stream);
}

int xs_bloom_query(xs_bloom* bf, const uint8_t* bases, uint64_t n_bases, const uint64_t* seq_begin,
                   const uint64_t* seq_end, uint64_t n_seq, uint32_t step, uint32_t* out_hits) {
    if (!bf || (!out_hits && n_seq) || (n_seq && (!seq_begin || !seq_end)) || (n_bases && !bases))
        return fail(XS_ERR_ARG, "NULL argument");
    if (step == 0) return fail(XS_ERR_ARG, "step must be >= 1");
    DeviceGuard guard(bf->info.device);
    if (!guard.ok) return fail(XS_ERR_CUDA, "cannot select the filter's device");
    {
        bool handled = false;
        XS_TRY(small_query(bf->info.device, bf->n_sm, bf->info.term_size, step, 0, 1, bases, n_bases, seq_begin, seq_end, n_seq, 4,
                           out_hits, &handled, [&](const SeqBatch& sb, const uint64_t*, uint8_t* d_o, cudaStream_t s) {
                               return bloom_launch(bf, sb, reinterpret_cast<uint32_t*>(d_o), s);
                           }));
        if (handled) return XS_OK;
    }
    return host_pipeline(bases, n_bases, seq_begin, seq_end, n_seq, 4,
                         [&](const uint8_t* db, uint64_t span, const uint64_t* d_b, const uint64_t* d_e, uint64_t ns,
                             uint64_t shift, uint8_t* d_o, uint64_t i0, cudaStream_t s) {
                             XS_TRY(bloom_query_dev(bf, db, span, d_b, d_e, ns, shift, step, (uint32_t*)d_o, s));
                             XS_CUDA(cudaMemcpyAsync(out_hits + i0, d_o, ns * 4, cudaMemcpyDeviceToHost, s));
                             return (int)XS_OK;
                         });
}

// ---- construction ---------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t n) {
        cudaError_t e = cudaMalloc(&p, n ? n : 1);
        return e == cudaSuccess ? XS_OK : fail(XS_ERR_NOMEM, cudaGetErrorString(e));
    }
};

static uint64_t cobs_signature_size(uint64_t max_doc_kmers, uint32_t num_hashes, double fpr) {
    double ratio = -(double)num_hashes / std::log(1.0 - std::pow(fpr, 1.0 / (double)num_hashes));
    double v = std::ceil((double)std::max<uint64_t>(max_doc_kmers, 1) * ratio);
    return v < 1.0 ? 1 : (uint64_t)v;
}

static bool put_all(FILE* f, const void* p, size_t n) { return n == 0 || fwrite(p, 1, n, f) == n; }

int xs_cobs_build(const char* out_path, int device, int kind, uint32_t k, uint32_t num_hashes, double fpr, uint32_t canonicalize,
                  uint64_t sig_size, uint64_t page_size, const char* names, uint32_t n_docs, const uint8_t* bases, uint64_t n_bases,
                  const uint64_t* seq_begin, const uint64_t* seq_end, const uint32_t* seq_doc, uint64_t n_seq) {
    if (!out_path || !names || n_docs == 0 || (n_seq && (!seq_begin || !seq_end || !seq_doc)) || (n_bases && !bases))
        return fail(XS_ERR_ARG, "NULL argument");
    if (kind != XS_COBS_CLASSIC && kind != XS_COBS_COMPACT) return fail(XS_ERR_ARG, "kind must be classic or compact");
    if (k == 0 || k > 32) return fail(XS_ERR_UNSUPPORTED, "term_size must be in 1..32");
    if (num_hashes == 0 || num_hashes > 64) return fail(XS_ERR_UNSUPPORTED, "num_hashes must be in 1..64");
    if (!(fpr > 0.0 && fpr < 1.0) && sig_size == 0) return fail(XS_ERR_ARG, "false positive rate must be in (0, 1)");
    // document names and sizes (k-mers per document)
    std::vector<std::string> name(n_docs);
    {
        const char* p = names;
        for (uint32_t d = 0; d < n_docs; ++d) {
            const char* e = strchr(p, '\n');
            if (!e) { name[d] = p; p += name[d].size(); } else { name[d].assign(p, e); p = e + 1; }
        }
    }
    std::vector<uint64_t> size(n_docs, 0);
    for (uint64_t i = 0; i < n_seq; ++i) {
        if (seq_doc[i] >= n_docs || seq_end[i] < seq_begin[i] || seq_end[i] > n_bases) return fail(XS_ERR_ARG, "bad sequence / document table");
        uint64_t len = seq_end[i] - seq_begin[i];
        if (len >= k) size[seq_doc[i]] += len - k + 1;
    }
    // page layout: classic = one page of all documents; compact = documents sorted by size, 8 * page_size per page
    std::vector<uint32_t> order(n_docs);
    std::iota(order.begin(), order.end(), 0u);
    uint64_t row_bytes, n_pages;
    if (kind == XS_COBS_CLASSIC) {
        row_bytes = ((uint64_t)n_docs + 7) / 8; n_pages = 1;
    } else {
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return size[a] < size[b]; });
        if (page_size == 0) page_size = std::max<uint64_t>(1, (uint64_t)std::sqrt((double)n_docs / 8.0));
        row_bytes = page_size; n_pages = ((uint64_t)n_docs + 8 * page_size - 1) / (8 * page_size);
    }
    std::vector<uint32_t> rank(n_docs);
    for (uint32_t r = 0; r < n_docs; ++r) rank[order[r]] = r;
    const uint64_t per_page = kind == XS_COBS_CLASSIC ? n_docs : 8 * row_bytes;
    std::vector<uint64_t> sig(n_pages);
    for (uint64_t pg = 0; pg < n_pages; ++pg) {
        uint64_t mx = 0;
        for (uint64_t r = pg * per_page; r < std::min<uint64_t>((pg + 1) * per_page, n_docs); ++r) mx = std::max(mx, size[order[r]]);
        sig[pg] = (sig_size && kind == XS_COBS_CLASSIC) ? sig_size : cobs_signature_size(mx, num_hashes, fpr);
    }
    // header
    std::string head = std::string("COBS:") + (kind == XS_COBS_CLASSIC ? "CLASSIC_INDEX" : "COMPACT_INDEX");
    auto pod = [&](const void* p, size_t n) { head.append((const char*)p, n); };
    uint32_t version = 1; uint8_t canon = (uint8_t)(canonicalize != 0);
    pod(&version, 4); pod(&k, 4); pod(&canon, 1);
    if (kind == XS_COBS_CLASSIC) {
        uint64_t nh = num_hashes;
        pod(&n_docs, 4); pod(&sig[0], 8); pod(&nh, 8);
    } else {
        uint32_t np = (uint32_t)n_pages;
        pod(&np, 4); pod(&n_docs, 4); pod(&row_bytes, 8);
        for (uint64_t pg = 0; pg < n_pages; ++pg) { uint64_t nh = num_hashes; pod(&sig[pg], 8); pod(&nh, 8); }
    }
    for (uint32_t r = 0; r < n_docs; ++r) { head += name[order[r]]; head += '\n'; }
    if (kind == XS_COBS_COMPACT) {
        // cobs pads so that the data starts on a multiple of page_size; when it already does, the two readings of its
        // rule differ (0 bytes, or one whole page with XS_COMPACT_PAD_FULL=1).  Both readers here accept both.
        uint64_t pad = (row_bytes - ((head.size() + 13) % row_bytes)) % row_bytes;
        const char* full = getenv("XS_COMPACT_PAD_FULL");
        if (pad == 0 && full && full[0] == '1') pad = row_bytes;
        head.append(pad, '\0');
    }
    head += kind == XS_COBS_CLASSIC ? "CLASSIC_INDEX" : "COMPACT_INDEX";

    DeviceGuard guard(device);
    int n_sm = 0;
    XS_TRY(device_setup(device, &n_sm));
    cudaStream_t s = nullptr;
    DevBuf d_bases, d_b, d_e, d_doc, d_data;
    XS_TRY(d_bases.alloc(n_bases + 64));
    XS_TRY(d_b.alloc(n_seq * 8)); XS_TRY(d_e.alloc(n_seq * 8)); XS_TRY(d_doc.alloc(n_seq * 4));
    std::vector<uint32_t> doc_rank(n_seq);
    for (uint64_t i = 0; i < n_seq; ++i) doc_rank[i] = rank[seq_doc[i]];
    if (n_bases) XS_CUDA(cudaMemcpy(d_bases.p, bases, n_bases, cudaMemcpyHostToDevice));
    if (n_seq) {
        XS_CUDA(cudaMemcpy(d_b.p, seq_begin, n_seq * 8, cudaMemcpyHostToDevice));
        XS_CUDA(cudaMemcpy(d_e.p, seq_end, n_seq * 8, cudaMemcpyHostToDevice));
        XS_CUDA(cudaMemcpy(d_doc.p, doc_rank.data(), n_seq * 4, cudaMemcpyHostToDevice));
    }
    Workspace ws;
    SeqBatch sb{};
    XS_TRY(prepare_batch(ws, sb, (const uint8_t*)d_bases.p, n_bases, (const uint64_t*)d_b.p, (const uint64_t*)d_e.p, n_seq, 0, k, 1, 0, n_sm, s));
    FILE* f = fopen(out_path, "wb");
    if (!f) return fail(XS_ERR_IO, std::string(out_path) + ": " + strerror(errno));
    int rc = put_all(f, head.data(), head.size()) ? XS_OK : fail(XS_ERR_IO, "write failed");
    std::vector<uint8_t> host;
    for (uint64_t pg = 0; pg < n_pages && rc == XS_OK; ++pg) {
        uint64_t bytes = sig[pg] * row_bytes;
        DevBuf d_page;
        rc = d_page.alloc(bytes + 8);
        if (rc != XS_OK) break;
        cudaMemsetAsync(d_page.p, 0, bytes + 8, s);
        CobsBuildParams bp{};
        bp.sb = sb; bp.seq_doc = (const uint32_t*)d_doc.p; bp.data = (uint8_t*)d_page.p; bp.sig_size = sig[pg]; bp.magic = magic_of(sig[pg]);
        bp.row_bytes = (uint32_t)row_bytes; bp.num_hashes = num_hashes; bp.canonicalize = canonicalize; bp.policy = XS_NONACGT_SKIP;
        bp.doc_lo = (uint32_t)(pg * per_page); bp.doc_hi = (uint32_t)std::min<uint64_t>((pg + 1) * per_page, n_docs);
        k_cobs_build<<<n_sm * 8, 256, 0, s>>>(bp);
        rc = launch_ok("k_cobs_build");
        if (rc != XS_OK) break;
        host.resize(bytes);
        cudaError_t e = cudaMemcpyAsync(host.data(), d_page.p, bytes, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { rc = fail(XS_ERR_CUDA, std::string("index build: ") + cudaGetErrorString(e)); break; }
        if (!put_all(f, host.data(), bytes)) rc = fail(XS_ERR_IO, "write failed");
    }
    if (fclose(f) != 0 && rc == XS_OK) rc = fail(XS_ERR_IO, "write failed");
    cudaStreamSynchronize(s);
    return rc;
}

int xs_bloom_build(const char* out_path, int device, uint32_t k, uint64_t expected_items, double fpr, const uint8_t* bases,
                   uint64_t n_bases, const uint64_t* seq_begin, const uint64_t* seq_end, uint64_t n_seq) {
    if (!out_path || (n_seq && (!seq_begin || !seq_end)) || (n_bases && !bases)) return fail(XS_ERR_ARG, "NULL argument");
    if (k == 0 || k > 32) return fail(XS_ERR_UNSUPPORTED, "term_size must be in 1..32");
    if (expected_items == 0 || !(fpr > 0.0 && fpr < 1.0)) return fail(XS_ERR_ARG, "expected_items must be > 0 and the false positive rate in (0, 1)");
    // rbloom.Bloom(expected_items, fpr): bits = -n ln(fpr) / ln(2)^2 (truncated), k = trunc(bits / n * ln 2)
    double size_in_bits = -1.0 * (double)expected_items * std::log(fpr) / std::pow(std::log(2.0), 2.0);
    uint64_t k_hashes = (uint64_t)((size_in_bits / (double)expected_items) * std::log(2.0));
    uint64_t n_bytes = ((uint64_t)size_in_bits + 7) / 8;
    if (n_bytes == 0) return fail(XS_ERR_ARG, "filter would be empty");
    DeviceGuard guard(device);
    int n_sm = 0;
    XS_TRY(device_setup(device, &n_sm));
    cudaStream_t s = nullptr;
    DevBuf d_bases, d_b, d_e, d_bits;
    XS_TRY(d_bases.alloc(n_bases + 64));
    XS_TRY(d_b.alloc(n_seq * 8)); XS_TRY(d_e.alloc(n_seq * 8)); XS_TRY(d_bits.alloc(n_bytes + 8));
    if (n_bases) XS_CUDA(cudaMemcpy(d_bases.p, bases, n_bases, cudaMemcpyHostToDevice));
    if (n_seq) {
        XS_CUDA(cudaMemcpy(d_b.p, seq_begin, n_seq * 8, cudaMemcpyHostToDevice));
        XS_CUDA(cudaMemcpy(d_e.p, seq_end, n_seq * 8, cudaMemcpyHostToDevice));
    }
    XS_CUDA(cudaMemsetAsync(d_bits.p, 0, n_bytes + 8, s));
    Workspace ws;
    BloomBuildParams bp{};
    XS_TRY(prepare_batch(ws, bp.sb, (const uint8_t*)d_bases.p, n_bases, (const uint64_t*)d_b.p, (const uint64_t*)d_e.p, n_seq, 0, k, 1, 0, n_sm, s));
    bp.bits = (uint8_t*)d_bits.p; bp.n_bits = n_bytes * 8; bp.magic = magic_of(n_bytes * 8); bp.k_hashes = (uint32_t)k_hashes;
    k_bloom_build<<<n_sm * 8, 256, 0, s>>>(bp);
    XS_TRY(launch_ok("k_bloom_build"));
    std::vector<uint8_t> host(n_bytes);
    XS_CUDA(cudaMemcpyAsync(host.data(), d_bits.p, n_bytes, cudaMemcpyDeviceToHost, s));
    XS_CUDA(cudaStreamSynchronize(s));
    FILE* f = fopen(out_path, "wb");
    if (!f) return fail(XS_ERR_IO, std::string(out_path) + ": " + strerror(errno));
    bool ok = put_all(f, &k_hashes, 8) && put_all(f, host.data(), n_bytes);
    if (fclose(f) != 0) ok = false;
    return ok ? XS_OK : fail(XS_ERR_IO, std::string(out_path) + ": write failed");
}

// ---- single stages -----------------------------------------------------------------------
static int stage_pack(const uint8_t* bases, uint64_t n_bases, int n_sm, DevBuf& d_bases, DevBuf& d_packed, DevBuf& d_inv) {
    uint64_t n_words = n_bases / 32 + 2;
    XS_TRY(d_bases.alloc(n_bases + 64));
    XS_TRY(d_packed.alloc(n_words * 8));
    XS_TRY(d_inv.alloc(n_words * 4));
    if (n_bases) XS_CUDA(cudaMemcpy(d_bases.p, bases, n_bases, cudaMemcpyHostToDevice));
    uint64_t grid = std::max<uint64_t>(1, std::min<uint64_t>((n_words + 255) / 256, (uint64_t)n_sm * 16));
    k_pack2bit<<<(unsigned)grid, 256>>>((const uint8_t*)d_bases.p, n_bases, (uint64_t*)d_packed.p, (uint32_t*)d_inv.p, n_words);
    return launch_ok("k_pack2bit");
}

int xs_pack_2bit(const uint8_t* bases, uint64_t n_bases, int device, uint64_t* packed, uint32_t* invalid) {
    if ((n_bases && !bases) || !packed || !invalid) return fail(XS_ERR_ARG, "NULL argument");
    DeviceGuard guard(device);
    int n_sm = 0;
    XS_TRY(device_setup(device, &n_sm));
    DevBuf db, dp, di;
    XS_TRY(stage_pack(bases, n_bases, n_sm, db, dp, di));
    uint64_t n_words = n_bases / 32 + 1;
    XS_CUDA(cudaMemcpy(packed, dp.p, n_words * 8, cudaMemcpyDeviceToHost));
    XS_CUDA(cudaMemcpy(invalid, di.p, n_words * 4, cudaMemcpyDeviceToHost));
    return XS_OK;
}

int xs_canonical_kmers(const uint8_t* bases, uint64_t n_bases, uint32_t k, int device, uint64_t* codes, uint8_t* valid) {
    if (k == 0 || k > 32) return fail(XS_ERR_ARG, "k must be in 1..32");
    if (n_bases < k) return XS_OK;
    if (!bases || !codes || !valid) return fail(XS_ERR_ARG, "NULL argument");
    DeviceGuard guard(device);
    int n_sm = 0;
    XS_TRY(device_setup(device, &n_sm));
    DevBuf db, dp, di, dc, dv;
    XS_TRY(stage_pack(bases, n_bases, n_sm, db, dp, di));
    uint64_t n_win = n_bases - k + 1;
    XS_TRY(dc.alloc(n_win * 8));
    XS_TRY(dv.alloc(n_win));
    k_stage_canonical<<<n_sm * 8, 256>>>((const uint64_t*)dp.p, (const uint32_t*)di.p, n_win, k, (uint64_t*)dc.p, (uint8_t*)dv.p);
    XS_TRY(launch_ok("k_stage_canonical"));
    XS_CUDA(cudaMemcpy(codes, dc.p, n_win * 8, cudaMemcpyDeviceToHost));
    XS_CUDA(cudaMemcpy(valid, dv.p, n_win, cudaMemcpyDeviceToHost));
    return XS_OK;
}

int xs_cobs_rows(xs_cobs* ix, const uint8_t* bases, uint64_t n_bases, uint32_t step, uint64_t* rows, uint8_t* valid) {
    if (!ix) return fail(XS_ERR_ARG, "NULL index");
    if (step == 0) return fail(XS_ERR_ARG, "step must be >= 1");
    uint32_t k = ix->info.term_size;
    if (n_bases < k) return XS_OK;
    if (!bases || !rows || !valid) return fail(XS_ERR_ARG, "NULL argument");
    DeviceGuard guard(ix->info.device);
    DevBuf db, dp, di, dr, dv;
    XS_TRY(stage_pack(bases, n_bases, ix->n_sm, db, dp, di));
    uint64_t n_win = (n_bases - k) / step + 1;
    uint32_t h = ix->info.num_hashes, np = (uint32_t)ix->pages.size();
    XS_TRY(dr.alloc(n_win * h * np * 8));
    XS_TRY(dv.alloc(n_win));
    SeqBatch sb{};
    sb.packed = (const uint64_t*)dp.p; sb.invalid = (const uint32_t*)di.p; sb.bases = (const uint8_t*)db.p;
    sb.n_seq = 1; sb.n_bases = n_bases; sb.step = step; sb.k = k;
    k_stage_cobs_rows<<<ix->n_sm * 8, 256>>>(sb, ix->d_pages, np, h, ix->info.canonicalize, ix->info.policy, n_win,
                                             (uint64_t*)dr.p, (uint8_t*)dv.p);
    XS_TRY(launch_ok("k_stage_cobs_rows"));
    XS_CUDA(cudaMemcpy(rows, dr.p, n_win * h * np * 8, cudaMemcpyDeviceToHost));
    XS_CUDA(cudaMemcpy(valid, dv.p, n_win, cudaMemcpyDeviceToHost));
    return XS_OK;
}

int xs_kmer_rows(const uint8_t* bases, uint64_t n_bases, uint32_t k, uint32_t canonicalize, uint32_t num_hashes,
                 uint64_t sig_size, uint32_t step, int device, uint64_t* rows, uint8_t* valid) {
    if (k == 0 || k > 32 || num_hashes == 0 || num_hashes > 64 || sig_size == 0 || step == 0) return fail(XS_ERR_ARG, "bad geometry");
    if (n_bases < k) return XS_OK;
    if (!bases || !rows || !valid) return fail(XS_ERR_ARG, "NULL argument");
    DeviceGuard guard(device);
    int n_sm = 0;
    XS_TRY(device_setup(device, &n_sm));
    DevBuf db, dp, di, dr, dv, dpg;
    XS_TRY(stage_pack(bases, n_bases, n_sm, db, dp, di));
    uint64_t n_win = (n_bases - k) / step + 1;
    XS_TRY(dr.alloc(n_win * num_hashes * 8));
    XS_TRY(dv.alloc(n_win));
    XS_TRY(dpg.alloc(sizeof(PageDesc)));
    PageDesc pd{};
    pd.sig_size = sig_size; pd.magic = magic_of(sig_size); pd.row_stride = 16;
    XS_CUDA(cudaMemcpy(dpg.p, &pd, sizeof(pd), cudaMemcpyHostToDevice));
    SeqBatch sb{};
    sb.packed = (const uint64_t*)dp.p; sb.invalid = (const uint32_t*)di.p; sb.bases = (const uint8_t*)db.p;
    sb.n_seq = 1; sb.n_bases = n_bases; sb.step = step; sb.k = k;
    k_stage_cobs_rows<<<n_sm * 8, 256>>>(sb, (const PageDesc*)dpg.p, 1, num_hashes, canonicalize, POLICY_SKIP, n_win,
                                         (uint64_t*)dr.p, (uint8_t*)dv.p);
    XS_TRY(launch_ok("k_stage_cobs_rows"));
    XS_CUDA(cudaMemcpy(rows, dr.p, n_win * num_hashes * 8, cudaMemcpyDeviceToHost));
    XS_CUDA(cudaMemcpy(valid, dv.p, n_win, cudaMemcpyDeviceToHost));
    return XS_OK;
}

int xs_bloom_hashes(xs_bloom* bf, const uint8_t* bases, uint64_t n_bases, uint32_t step, uint64_t* hashes) {
    if (!bf) return fail(XS_ERR_ARG, "NULL filter");
    if (step == 0) return fail(XS_ERR_ARG, "step must be >= 1");
    uint32_t k = bf->info.term_size;
    if (n_bases < k) return XS_OK;
    if (!bases || !hashes) return fail(XS_ERR_ARG, "NULL argument");
    DeviceGuard guard(bf->info.device);
    DevBuf db, dp, di, dh;
    XS_TRY(stage_pack(bases, n_bases, bf->n_sm, db, dp, di));
    uint64_t n_win = (n_bases - k) / step + 1;
    XS_TRY(dh.alloc(n_win * 8));
    SeqBatch sb{};
    sb.packed = (const uint64_t*)dp.p; sb.invalid = (const uint32_t*)di.p; sb.bases = (const uint8_t*)db.p;
    sb.n_seq = 1; sb.n_bases = n_bases; sb.step = step; sb.k = k;
    k_stage_bloom_hashes<<<bf->n_sm * 8, 256>>>(sb, n_win, (uint64_t*)dh.p);
    XS_TRY(launch_ok("k_stage_bloom_hashes"));
    XS_CUDA(cudaMemcpy(hashes, dh.p, n_win * 8, cudaMemcpyDeviceToHost));
    return XS_OK;
}

}  // extern "C"
