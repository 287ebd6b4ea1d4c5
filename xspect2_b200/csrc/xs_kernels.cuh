// xs_kernels.cuh — sm_100a kernels of the k-mer scoring path.
//
//   k_pack2bit        ASCII bases -> 2-bit stream + non-ACGT bitmap             (feeder, file_io.py:47-79)
//   k_cobs_narrow     rows <= 16 B (D <= 128 per page): one thread per sampled window; extract ->
//                     canonical -> XXH64 x h -> Barrett mod -> h x 128-bit gathers -> AND ->
//                     per-sequence document counts by warp ballot/popcount
//                     (cobs Search.search behind probabilistic_filter_model.py:227)
//   k_bucket_emit,    the same scoring for large batches against an index much larger than L2: probe records grouped
//   k_bucket_fetch,     by L2-sized row ranges, rows gathered from L2, ANDed per window in shared memory (a random
//   k_bucket_reduce     16-byte row gather from DRAM costs a 128-byte fetch; every row is probed many times per batch)
//   k_cobs_wide       rows > 16 B: lanes own 16-byte column chunks, warp-coalesced row gathers (2h loads in
//                     flight per lane), bit-plane counters, shared-memory count staging
//   k_bloom           XXH3-64 -> 128-bit LCG -> bit probes with early exit
//                     (probabilistic_single_filter_model.py:122-124,161-180)
//   k_bloom_sample,   the bucketed scheme for the Bloom filter, chosen on the device per batch from a sampled member
//   k_bbucket_emit,     fraction (all k probes are made there; k_bloom stops at the first zero bit)
//   k_bbucket_fetch,
//   k_bbucket_reduce
//   k_scores_reduce   per-record best document / tie multiplicity and per-document totals over a count matrix
//   k_cobs_build,     construction: the query hashing with an atomic OR instead of a gather
//   k_bloom_build       (probabilistic_filter_model.py:169-194, probabilistic_single_filter_model.py:63-96)
//   stage kernels     canonical codes / row ids / bloom hashes alone, for the parity tests
//
// Work decomposition.  The sampled windows of all sequences of a batch form one flat index
// space (win_prefix = exclusive scan of windows per sequence).  Persistent warps walk tiles of
// that space (warp_walk), so lanes stay dense whatever the read lengths are; all bookkeeping
// is warp-uniform (no shared memory, no CTA barrier on the narrow and Bloom paths).  Tiles and work
// items are handed out through an atomic counter: a static equal split runs at the pace of the
// slowest SM (profiles/microbench/gather_concurrent.cu).  The bucketed kernels cut the same flat space into
// chunks of 2048 windows and resolve window -> sequence per chunk in shared memory (chunk_seq_table).
#pragma once
#include "xs_device.cuh"

namespace xs {

// ----------------------------------------------------------------------------------------
// shared structures
// ----------------------------------------------------------------------------------------
struct PageDesc {
    const uint8_t* data;   // re-strided rows of this page (this handle's column shard)
    uint64_t sig_size;     // rows
    uint64_t magic;        // floor(2^64 / sig_size)
    uint32_t row_stride;   // bytes per row in HBM (multiple of 16)
    uint32_t n_docs;       // documents (columns) of this page held here
    uint32_t doc_off;      // first output column of this page
    uint32_t pad;
};

struct SeqBatch {
    const uint64_t* packed;      // 2-bit stream of the batch's bases
    const uint32_t* invalid;     // non-ACGT bitmap
    const uint8_t* bases;        // raw bytes (literal path)
    const uint64_t* seq_begin;   // [n_seq] byte offsets (minus base_shift -> index into bases)
    const uint64_t* seq_end;     // [n_seq]
    const uint64_t* win_prefix;  // [n_seq + 1] exclusive scan of windows per sequence
    unsigned long long* tile_counter;  // [grid.y] zeroed per launch: dynamic tile / work-item hand-out
    uint64_t n_seq;
    uint64_t n_bases;
    uint64_t base_shift;
    uint32_t step;
    uint32_t k;
};

XS_HD uint64_t windows_of(uint64_t b, uint64_t e, uint64_t shift, uint64_t n_bases, uint32_t k, uint32_t step) {
    if (e < b || b < shift || e - shift > n_bases) return 0;
    uint64_t len = e - b;
    return len >= k ? (len - k) / step + 1 : 0;
}

struct WindowCountOp {  // windows (chunk == 0) or work items of `chunk` windows per sequence
    const uint64_t* seq_begin; const uint64_t* seq_end;
    uint64_t n_seq, shift, n_bases; uint32_t k, step, chunk;  // chunk > 0: count chunks of that many windows
    XS_HD uint64_t operator()(uint64_t i) const {
        if (i >= n_seq) return 0;
        uint64_t nw = windows_of(seq_begin[i], seq_end[i], shift, n_bases, k, step);
        return chunk ? (nw + chunk - 1) / chunk : nw;
    }
};

__global__ void __launch_bounds__(256) k_count_windows(const WindowCountOp op, uint64_t* __restrict__ counts) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= op.n_seq; i += (uint64_t)gridDim.x * blockDim.x)
        counts[i] = op(i);
}

// largest i in [0, n) with prefix[i] <= g   (prefix[0] = 0, prefix[n] > g)
__device__ __forceinline__ uint64_t seq_of_window(const uint64_t* __restrict__ prefix, uint64_t n, uint64_t g) {
    uint64_t lo = 0, hi = n;
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if (__ldg(prefix + mid) <= g) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ uint4 ldg128(const uint8_t* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <typename OutT> __device__ __forceinline__ uint32_t out_max();
template <> __device__ __forceinline__ uint32_t out_max<uint8_t>() { return 255u; }
template <> __device__ __forceinline__ uint32_t out_max<uint16_t>() { return 65535u; }
template <> __device__ __forceinline__ uint32_t out_max<uint32_t>() { return 0xFFFFFFFFu; }

template <typename OutT>
__device__ __forceinline__ void out_store(OutT* p, uint32_t v) {
    *p = (OutT)(v > out_max<OutT>() ? out_max<OutT>() : v);
}
// saturating add into an element other CTAs may add to as well.  nosat: the record has no more windows than OutT
// can hold, so no partial sum can carry into the neighbouring element and a plain word atomic is exact
template <typename OutT>
__device__ __forceinline__ void out_add(OutT* p, uint32_t v, bool nosat) {
    if (v == 0) return;
    if (sizeof(OutT) == 4) {
        atomicAdd(reinterpret_cast<unsigned int*>(p), v);
    } else {
        uintptr_t a = reinterpret_cast<uintptr_t>(p);
        unsigned int* w = reinterpret_cast<unsigned int*>(a & ~(uintptr_t)3);
        uint32_t sh = (uint32_t)(a & 3) * 8;
        if (nosat) { atomicAdd(w, v << sh); return; }
        uint32_t mx = out_max<OutT>();
        unsigned int old = *w, assumed;
        do {
            assumed = old;
            uint32_t cur = (assumed >> sh) & mx;
            uint32_t nv = cur + v; if (nv > mx) nv = mx;
            unsigned int repl = (assumed & ~(mx << sh)) | (nv << sh);
            if (repl == assumed) break;
            old = atomicCAS(w, assumed, repl);
        } while (old != assumed);
    }
}

// ----------------------------------------------------------------------------------------
// 2-bit packing.  One thread per 32 bases.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void pack_bytes4(uint32_t v, int j0, uint64_t& code, uint32_t& inv) {
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        uint32_t c = (v >> (8 * b)) & 0xFF;
        uint32_t cd = ((c >> 1) ^ (c >> 2)) & 3;
        bool ok = (c == 'A') | (c == 'C') | (c == 'G') | (c == 'T');
        if (ok) code |= (uint64_t)cd << (2 * (j0 + b)); else inv |= 1u << (j0 + b);
    }
}

__global__ void __launch_bounds__(256) k_pack2bit(const uint8_t* __restrict__ bases, uint64_t n_bases,
                                                  uint64_t* __restrict__ packed, uint32_t* __restrict__ invalid,
                                                  uint64_t n_words) {
    const bool aligned = (reinterpret_cast<uintptr_t>(bases) & 15) == 0;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t b0 = w * 32;
        uint64_t code = 0; uint32_t inv = 0;
        if (b0 + 32 <= n_bases && aligned) {
            const uint4* p = reinterpret_cast<const uint4*>(bases + b0);
            uint4 x = __ldg(p), y = __ldg(p + 1);
            pack_bytes4(x.x, 0, code, inv);  pack_bytes4(x.y, 4, code, inv);
            pack_bytes4(x.z, 8, code, inv);  pack_bytes4(x.w, 12, code, inv);
            pack_bytes4(y.x, 16, code, inv); pack_bytes4(y.y, 20, code, inv);
            pack_bytes4(y.z, 24, code, inv); pack_bytes4(y.w, 28, code, inv);
        } else {
            for (int j = 0; j < 32; ++j) {
                uint64_t b = b0 + j;
                if (b < n_bases) {
                    uint32_t c = __ldg(bases + b);
                    uint32_t cd = ((c >> 1) ^ (c >> 2)) & 3;
                    bool ok = (c == 'A') | (c == 'C') | (c == 'G') | (c == 'T');
                    if (ok) code |= (uint64_t)cd << (2 * j); else inv |= 1u << j;
                } else {
                    inv |= 1u << j;  // beyond the buffer: never a valid window
                }
            }
        }
        packed[w] = code;
        invalid[w] = inv;
    }
}

// ----------------------------------------------------------------------------------------
// term of one window: fast 2-bit path or literal bytes.  Returns false when the window is
// skipped (COBS SKIP policy).
// ----------------------------------------------------------------------------------------
enum { POLICY_SKIP = 0, POLICY_LITERAL = 1 };

template <int K>
__device__ __forceinline__ bool cobs_term(const SeqBatch& sb, uint64_t pos, bool canonicalize, int policy, Term& t) {
    const uint32_t k = K ? K : sb.k;
    if (window_invalid(sb.invalid, pos, k)) {
        if (policy == POLICY_SKIP && canonicalize) return false;
        // canonicalize == 0 hashes the literal bytes whatever they are (cobs create_hashes)
        literal_term(sb.bases, pos, k, CobsComp(), canonicalize, t);
        return true;
    }
    uint64_t fr = window_lsb(sb.packed, pos, k);
    uint64_t msb;
    uint64_t c = canonicalize ? canonical_lsb(fr, k, &msb) : fr;
    expand_ascii(c, k, t);
    return true;
}

template <int K>
__device__ __forceinline__ void bloom_term(const SeqBatch& sb, uint64_t pos, Term& t) {
    const uint32_t k = K ? K : sb.k;
    if (window_invalid(sb.invalid, pos, k)) {
        literal_term(sb.bases, pos, k, BioComp(), true, t);
        return;
    }
    uint64_t fr = window_lsb(sb.packed, pos, k);
    uint64_t msb;
    uint64_t c = canonical_lsb(fr, k, &msb);
    expand_ascii(c, k, t);
}

// ----------------------------------------------------------------------------------------
// warp walk: one warp owns a tile [t0, t1) of the flat window space and walks the sequences it
// touches.  Each round packs the next 32 windows into the lanes (a round may span several
// sequences); `compute(has, pos, g)` is called once per lane with the base position and the flat index of
// its window, `accum(segmask)` once per (round, sequence) with the lanes of that sequence,
// `flush(seq, complete, nwin)` when a sequence (or the tile) ends (nwin = all windows of the sequence).  All bookkeeping is warp-uniform: no shared memory, no CTA
// barrier, lanes stay dense whatever the read lengths are.
// ----------------------------------------------------------------------------------------
struct WalkBatch {           // 32 consecutive sequences, one per lane
    uint64_t pre, pre_next;  // win_prefix[s], win_prefix[s + 1]
    uint64_t beg;            // seq_begin[s] - base_shift
};

__device__ __forceinline__ void walk_load(const SeqBatch& sb, uint64_t base, uint32_t lane, WalkBatch& wb) {
    uint64_t s = base + lane;
    uint64_t a = s < sb.n_seq ? s : sb.n_seq, b = s + 1 < sb.n_seq ? s + 1 : sb.n_seq;
    wb.pre = __ldg(sb.win_prefix + a);
    wb.pre_next = __ldg(sb.win_prefix + b);
    wb.beg = s < sb.n_seq ? __ldg(sb.seq_begin + s) - sb.base_shift : 0;
}

template <class Compute, class Accum, class Flush>
__device__ __forceinline__ void warp_walk(const SeqBatch& sb, uint64_t t0, uint64_t t1, uint32_t lane,
                                          Compute compute, Accum accum, Flush flush) {
    const uint32_t FULL = 0xFFFFFFFFu;
    uint64_t base = seq_of_window(sb.win_prefix, sb.n_seq, t0);
    WalkBatch wb;
    walk_load(sb, base, lane, wb);
    uint32_t cur = 0;
    uint64_t f = t0;
    while (f < t1) {
        if (cur == 32) { base += 32; walk_load(sb, base, lane, wb); cur = 0; }
        const uint32_t cur0 = cur;
        const uint64_t f0 = f;
        // ---- pack up to 32 windows into the lanes
        uint32_t fill = 0;
        bool has = false;
        uint64_t pos = 0, gidx = 0;
        while (fill < 32 && f < t1 && cur < 32) {
            uint64_t s_start = __shfl_sync(FULL, wb.pre, cur), s_end = __shfl_sync(FULL, wb.pre_next, cur);
            uint64_t lim = s_end < t1 ? s_end : t1;
            uint32_t avail = lim > f ? (uint32_t)(lim - f > 32 ? 32 : lim - f) : 0;
            if (avail == 0) { ++cur; continue; }
            uint32_t take = avail < 32 - fill ? avail : 32 - fill;
            uint64_t b = __shfl_sync(FULL, wb.beg, cur);
            if (lane >= fill && lane < fill + take) {
                gidx = f + (lane - fill);
                pos = b + (gidx - s_start) * sb.step;
                has = true;
            }
            fill += take; f += take;
            if (f == s_end) ++cur;
        }
        compute(has, pos, gidx);
        // ---- replay the packing to attribute lanes to sequences
        cur = cur0; f = f0; fill = 0;
        while (fill < 32 && f < t1 && cur < 32) {
            uint64_t s_start = __shfl_sync(FULL, wb.pre, cur), s_end = __shfl_sync(FULL, wb.pre_next, cur);
            uint64_t lim = s_end < t1 ? s_end : t1;
            uint32_t avail = lim > f ? (uint32_t)(lim - f > 32 ? 32 : lim - f) : 0;
            if (avail == 0) { ++cur; continue; }
            uint32_t take = avail < 32 - fill ? avail : 32 - fill;
            uint32_t segmask = (take == 32 ? FULL : ((1u << take) - 1u)) << fill;
            accum(segmask);
            fill += take; f += take;
            if (f == s_end || f == t1) flush(base + cur, s_start >= t0 && s_end <= t1, s_end - s_start);
            if (f == s_end) ++cur;
        }
    }
}

__device__ __forceinline__ uint64_t next_tile(unsigned long long* counter, uint32_t lane) {
    unsigned long long t = 0;
    if (lane == 0) t = atomicAdd(counter, 1ULL);
    return __shfl_sync(0xFFFFFFFFu, t, 0);
}

// rounds per warp tile: large enough that few sequences straddle tiles, small enough to fill the machine
__device__ __forceinline__ uint64_t walk_tile_windows(uint64_t total, uint64_t n_warps) {
    uint64_t r = (total + n_warps * 512 - 1) / (n_warps * 512);   // >= ~16 tiles per warp
    r = r < 8 ? 8 : (r > 128 ? 128 : r);
    return r * 32;
}

// ----------------------------------------------------------------------------------------
// COBS, narrow rows (row_stride == 16, <= 128 documents per page)
// ----------------------------------------------------------------------------------------
struct CobsParams {
    SeqBatch sb;
    const PageDesc* pages;
    uint32_t n_pages;
    uint32_t num_hashes;
    uint32_t canonicalize;
    int32_t policy;
    void* out;          // [n_seq x ld] OutT
    uint64_t ld;        // output row length (all local documents)
    uint64_t seq0;      // output row of sequence 0 of this batch
    uint64_t win_begin; // k_cobs_narrow: first flat window to score (the bucketed path's tail launch), normally 0
    uint32_t all_rows;  // k_cobs_mid: gather all h rows even when the AND is already empty (measurement switch)
};

constexpr int NARROW_NT = 256;
#ifndef XS_EARLY_EXIT_AFTER
#define XS_EARLY_EXIT_AFTER 5
#endif

// 128-bit document mask of one window: h row gathers ANDed
template <int K, int H>
__device__ __forceinline__ uint4 cobs_mask16(const CobsParams& p, const PageDesc& pg, uint64_t pos) {
    const SeqBatch& sb = p.sb;
    const uint32_t k = K ? K : sb.k;
    Term t;
    if (!cobs_term<K>(sb, pos, p.canonicalize != 0, p.policy, t)) return make_uint4(0, 0, 0, 0);
    Xxh64Pre pre;
    xxh64_prepare(t, k, pre);
    uint4 m = make_uint4(~0u, ~0u, ~0u, ~0u);
    if (H) {
        // the first H1 rows go out together; when their AND is already empty the remaining gathers cannot change
        // the result and are skipped (every gather is a 128-byte DRAM fetch)
        constexpr int H1 = (H ? H : 1) > 5 ? (XS_EARLY_EXIT_AFTER) : (H ? H : 1);
        const uint8_t* addr[H ? H : 1];
#pragma unroll
        for (int j = 0; j < (H ? H : 1); ++j) {
            uint64_t hv = xxh64_finish(pre, k, (uint64_t)j);
            addr[j] = pg.data + mod_barrett(hv, pg.sig_size, pg.magic) * 16;
        }
#pragma unroll
        for (int j = 0; j < H1; ++j) {
            uint4 v = ldg128(addr[j]);
            m.x &= v.x; m.y &= v.y; m.z &= v.z; m.w &= v.w;
        }
        if (H1 < (H ? H : 1) && (m.x | m.y | m.z | m.w) != 0) {
#pragma unroll
            for (int j = H1; j < (H ? H : 1); ++j) {
                uint4 v = ldg128(addr[j]);
                m.x &= v.x; m.y &= v.y; m.z &= v.z; m.w &= v.w;
            }
        }
    } else {
        for (uint32_t j = 0; j < p.num_hashes; ++j) {
            uint64_t hv = xxh64_finish(pre, k, (uint64_t)j);
            uint4 v = ldg128(pg.data + mod_barrett(hv, pg.sig_size, pg.magic) * 16);
            m.x &= v.x; m.y &= v.y; m.z &= v.z; m.w &= v.w;
        }
    }
    return m;
}

// per-sequence document counts of the masks held by the lanes in `segmask`: lane l keeps the counts of documents
// l, l+32, l+64, l+96
__device__ __forceinline__ void narrow_accum(uint32_t lane, uint32_t segmask, const uint4& m, uint32_t (&cnt)[4]) {
    const bool mine = (segmask >> lane) & 1u;
    uint32_t mw[4] = {mine ? m.x : 0u, mine ? m.y : 0u, mine ? m.z : 0u, mine ? m.w : 0u};
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        uint32_t u = __reduce_or_sync(0xFFFFFFFFu, mw[w]);
        while (u) {
            uint32_t b = __ffs(u) - 1;
            u &= u - 1;
            uint32_t v = __popc(__ballot_sync(0xFFFFFFFFu, (mw[w] >> b) & 1u));
            if (lane == b) cnt[w] += v;
        }
    }
}

template <typename OutT>
__device__ __forceinline__ void narrow_flush(const CobsParams& p, const PageDesc& pg, uint32_t lane, uint64_t seq, bool complete,
                                             uint64_t nwin, uint32_t (&cnt)[4]) {
    OutT* row = reinterpret_cast<OutT*>(p.out) + (p.seq0 + seq) * p.ld + pg.doc_off;
    const bool nosat = nwin <= (uint64_t)out_max<OutT>();
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        uint32_t d = w * 32 + lane;
        if (d < pg.n_docs) {
            if (complete) out_store<OutT>(row + d, cnt[w]); else out_add<OutT>(row + d, cnt[w], nosat);
        }
        cnt[w] = 0;
    }
}

template <int K, int H, typename OutT>
__global__ void __launch_bounds__(NARROW_NT, 4) k_cobs_narrow(const CobsParams p) {
    const SeqBatch& sb = p.sb;
    const PageDesc pg = p.pages[blockIdx.y];
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t n_warps = ((uint64_t)gridDim.x * NARROW_NT) >> 5;

    const uint64_t total = __ldg(sb.win_prefix + sb.n_seq);
    const uint64_t first = p.win_begin < total ? p.win_begin : total;
    const uint64_t W = walk_tile_windows(total - first, n_warps);
    const uint64_t n_tiles = (total - first + W - 1) / W;

    // tiles are handed out dynamically: SMs do not all sustain the same gather rate (two dies, 70/78 SM
    // split), and a static equal split runs at the pace of the slowest one (profiles/microbench)
    for (;;) {
        const uint64_t tile = next_tile(sb.tile_counter + blockIdx.y, lane);
        if (tile >= n_tiles) break;
        const uint64_t t0 = first + tile * W, t1 = t0 + W < total ? t0 + W : total;
        uint4 m = make_uint4(0, 0, 0, 0);
        uint32_t cnt[4] = {0, 0, 0, 0};   // lane l: documents l, l+32, l+64, l+96 of the current sequence
        warp_walk(
            sb, t0, t1, lane,
            [&](bool has, uint64_t pos, uint64_t) { m = has ? cobs_mask16<K, H>(p, pg, pos) : make_uint4(0, 0, 0, 0); },
            [&](uint32_t segmask) { narrow_accum(lane, segmask, m, cnt); },
            [&](uint64_t seq, bool complete, uint64_t nwin) { narrow_flush<OutT>(p, pg, lane, seq, complete, nwin, cnt); });
    }
}

// ----------------------------------------------------------------------------------------
// COBS, narrow rows, several pages (compact indices: one <locus>.cobs_compact of an MLST scheme has ~10 pages of 64
// alleles).  k_cobs_narrow walks the batch once per page (grid.y) and hashes every window again for each of them;
// here a window is extracted, canonicalised and hashed ONCE (XXH64 does not depend on the page, only `% signature_size`
// does) and its rows are gathered from all pages, PAGES_PG pages in flight per lane.  Per-page document counts of the
// current sequence live in shared memory (one private slice per warp, no atomics); rows may sit at an 8-byte stride
// (page_size <= 8: two rows per 16 bytes, which halves the index's L2 footprint — an MLST locus then fits in L2).
// ----------------------------------------------------------------------------------------
constexpr int PAGES_NT = 256;
constexpr int PAGES_PG = 8;

__device__ __forceinline__ uint4 page_row(const PageDesc& pg, uint64_t hv) {
    const uint64_t r = mod_barrett(hv, pg.sig_size, pg.magic);
    if (pg.row_stride == 8) {
        uint2 v;
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(pg.data + r * 8));
        return make_uint4(v.x, v.y, 0u, 0u);
    }
    return ldg128(pg.data + r * 16);
}

template <int K, int H, typename OutT>
__global__ void __launch_bounds__(PAGES_NT) k_cobs_pages(const CobsParams p) {
    extern __shared__ uint32_t s_pages_cnt[];                 // [warps][n_pages * 128]
    const SeqBatch& sb = p.sb;
    const uint32_t k = K ? K : sb.k;
    const uint32_t h = H ? H : p.num_hashes;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_pages = p.n_pages;
    uint32_t* my = s_pages_cnt + (size_t)warp * n_pages * 128;
    for (uint32_t i = lane; i < n_pages * 128; i += 32) my[i] = 0;
    __syncwarp();
    const uint64_t n_warps = ((uint64_t)gridDim.x * PAGES_NT) >> 5;
    const uint64_t total = __ldg(sb.win_prefix + sb.n_seq);
    const uint64_t W = walk_tile_windows(total, n_warps);
    const uint64_t n_tiles = (total + W - 1) / W;
    OutT* out = reinterpret_cast<OutT*>(p.out);
    for (;;) {
        const uint64_t tile = next_tile(sb.tile_counter, lane);
        if (tile >= n_tiles) break;
        const uint64_t t0 = tile * W, t1 = t0 + W < total ? t0 + W : total;
        bool valid = false;
        Xxh64Pre pre;
        uint64_t hv0 = 0;
        warp_walk(
            sb, t0, t1, lane,
            [&](bool has, uint64_t pos, uint64_t) {
                valid = false;
                if (has) {
                    Term t;
                    valid = cobs_term<K>(sb, pos, p.canonicalize != 0, p.policy, t);
                    if (valid) {
                        xxh64_prepare(t, k, pre);
                        hv0 = xxh64_finish(pre, k, 0);
                    }
                }
            },
            [&](uint32_t segmask) {
                const bool mine = valid && ((segmask >> lane) & 1u);
                for (uint32_t p0 = 0; p0 < n_pages; p0 += PAGES_PG) {
                    uint4 m[PAGES_PG];
#pragma unroll
                    for (int q = 0; q < PAGES_PG; ++q) {
                        m[q] = make_uint4(0, 0, 0, 0);
                        if (p0 + q < n_pages && mine) {
                            const PageDesc pg = p.pages[p0 + q];
                            m[q] = page_row(pg, hv0);
                            for (uint32_t j = 1; j < h; ++j) {
                                const uint4 v = page_row(pg, xxh64_finish(pre, k, (uint64_t)j));
                                m[q].x &= v.x; m[q].y &= v.y; m[q].z &= v.z; m[q].w &= v.w;
                            }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < PAGES_PG; ++q) {
                        if (p0 + q >= n_pages) break;
                        const uint32_t mw[4] = {m[q].x, m[q].y, m[q].z, m[q].w};
                        uint32_t* c = my + (size_t)(p0 + q) * 128;
#pragma unroll
                        for (int w = 0; w < 4; ++w) {
                            uint32_t u = __reduce_or_sync(0xFFFFFFFFu, mw[w]);
                            while (u) {
                                const uint32_t b = __ffs(u) - 1;
                                u &= u - 1;
                                const uint32_t v = __popc(__ballot_sync(0xFFFFFFFFu, (mw[w] >> b) & 1u));
                                if (lane == b) c[w * 32 + b] += v;
                            }
                        }
                    }
                }
            },
            [&](uint64_t seq, bool complete, uint64_t nwin) {
                __syncwarp();
                const bool nosat = nwin <= (uint64_t)out_max<OutT>();
                OutT* row = out + (p.seq0 + seq) * p.ld;
                for (uint32_t i = lane; i < n_pages * 128; i += 32) {
                    const uint32_t v = my[i];
                    if (v) {
                        const PageDesc& pg = p.pages[i >> 7];
                        const uint32_t d = i & 127;
                        if (d < pg.n_docs) {
                            if (complete) out_store<OutT>(row + pg.doc_off + d, v); else out_add<OutT>(row + pg.doc_off + d, v, nosat);
                        }
                        my[i] = 0;
                    }
                }
                __syncwarp();
            });
    }
}

// ----------------------------------------------------------------------------------------
// COBS, narrow rows, bucketed probing (large batches against an index much larger than L2).
//
// A random 16-byte row gather costs a 128-byte DRAM fetch, so k_cobs_narrow moves ~10x the bytes it uses.  For large
// batches every index row is probed many times; the three kernels below turn those re-reads into L2 hits:
//
//   k_bucket_emit     window -> canonical -> XXH64 x h -> row ids; probe record (local window, row in bucket) appended
//                     to the block of (chunk of BK_CH windows, bucket of 2^bshift rows), staged in shared memory and
//                     written out in whole sectors
//   k_bucket_fetch    bucket by bucket (all warps sweep one bucket's blocks at a time, its 16-32 MB of rows stay in
//                     L2): row gather per record, written next to the chunk it belongs to
//   k_bucket_reduce   CTA per chunk: AND of the fetched rows into per-window masks in shared memory, then the
//                     per-sequence document counts exactly as k_cobs_narrow does
//
// Blocks have a fixed capacity (mean + ~3 sigma); a window with a probe that does not fit (low-complexity reads put
// thousands of identical k-mers into one block) is flagged and scored with direct gathers by k_bucket_reduce.
// Results are identical to k_cobs_narrow's whatever the record order inside a block.
// ----------------------------------------------------------------------------------------
constexpr int BK_CH = 2048;            // windows per chunk (11-bit local window index)
constexpr int BK_NT = 256;
constexpr int BK_WARP_WIN = BK_CH / (BK_NT / 32);
constexpr uint32_t BK_MAX_BUCKETS = 256;
constexpr uint32_t BK_MAX_SHIFT = 21;  // 11 + 21 = 32 bits per probe record

struct alignas(64) TensorMap128 { uint64_t q[16]; };   // a CUtensorMap (driver API), opaque here

struct BucketParams {
    TensorMap128 tmap;     // 2-D view of the index rows [sig_size][4 x u32] for tile::gather4 (k_bucket_fetch_tma)
    CobsParams cp;
    uint32_t* rec;        // [n_buckets][nc][cap]  (local window << bshift) | (row & (2^bshift - 1))
    uint4* rows;          // [nc][n_buckets][cap]  fetched rows; .w = local window when pack_id
    uint16_t* cnt_bc;     // [n_buckets][nc]       records per block
    uint16_t* cnt_cb;     // [nc][n_buckets]
    uint32_t* ovf;        // [nc][2][BK_CH / 32]   windows to score with direct gathers | windows that are not scored
    const uint64_t* chunk_seq;   // [all chunks of the batch] sequence of each chunk's first window (k_bucket_chunk_seq)
    unsigned long long* counter;   // [4] work hand-out of emit / fetch / reduce / k_bucket_hash, zeroed
    uint64_t chunk0;      // first chunk of this sub-batch in the flat window space
    uint32_t nc;          // chunks of this sub-batch
    uint32_t n_buckets, bshift, cap;
    uint32_t pack_id;     // rows hold <= 96 documents: the fourth word is free for the local window
    uint32_t prefetch;    // k_bucket_fetch prefetches the next bucket's rows into L2
    // hash-ahead: k_bucket_fetch of sub-batch i also hashes the windows of sub-batch i + 1 (its warps wait on L2 gathers
    // most of the time; the XXH64 work fills the idle issue slots), so k_bucket_emit only has to scatter row ids
    uint32_t* pre_rows;        // [nc][H][BK_CH] row ids of this sub-batch's windows (written by the previous fetch), or null
    uint32_t* pre_skp;         // [nc][BK_CH / 32] windows that are not scored
    uint32_t* next_rows;       // the same two arrays of the NEXT sub-batch, filled by this launch of k_bucket_fetch
    uint32_t* next_skp;
    uint64_t next_chunk0;      // first chunk of the next sub-batch
    uint32_t next_nc;          // its chunks (0 = nothing to hash)
    uint32_t fetch_off;        // 1: this launch only hashes (the first sub-batch of a query)
    uint32_t use_tma;          // k_bucket_fetch_tma: records per lane and pass fetched with tile::gather4 (0 = LSU only)
};

__device__ __forceinline__ uint32_t ld_stream32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.cs.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ld_stream128(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
// row gathers of the bucket being swept: keep them in L2 ahead of the record / row streams passing through
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint4 ldg128_keep(const uint8_t* p, uint64_t pol) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void prefetch_l2_keep(const uint8_t* p, uint64_t pol) {
    asm volatile("prefetch.global.L2::evict_last [%0];" :: "l"(p));
    (void)pol;
}
__device__ __forceinline__ void st_stream128(uint4* p, const uint4& v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// chunks of the sub-batch that hold windows
__device__ __forceinline__ uint32_t bucket_live_chunks(const BucketParams& bp, uint64_t total) {
    const uint64_t first = bp.chunk0 * BK_CH;
    if (total <= first) return 0;
    const uint64_t n = (total - first + BK_CH - 1) / BK_CH;
    return n < bp.nc ? (uint32_t)n : bp.nc;
}

// Sequence of every window of the chunk [g0, g1): s_wseq[l] = (sequence of window g0 + l) - s_lo + 1, where s_lo (from
// k_bucket_chunk_seq, also the return value) is the sequence of window g0.  Sequences that start inside the chunk mark their first window, an
// inclusive max-scan carries the marks forward.  CTA-wide (NT threads); s_wseq must be zero on entry; ends with a barrier.  This replaces warp_walk
// in the bucketed kernels: its 64-bit shuffle bookkeeping is hidden by DRAM latency in k_cobs_narrow but was a third
// of k_bucket_emit's and half of k_bucket_reduce's instructions (profiles/r1_bucketed_notes.md).
// sequence of the first window of every chunk: one binary search per chunk, all chunks in parallel (inside the
// kernels below the same search by one thread per chunk stalled the other 511 for ~27 dependent loads)
__global__ void __launch_bounds__(256) k_bucket_chunk_seq(const SeqBatch sb, uint64_t n_chunks, uint64_t* __restrict__ chunk_seq) {
    const uint64_t total = __ldg(sb.win_prefix + sb.n_seq);
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t g0 = c * BK_CH;
        chunk_seq[c] = g0 < total ? seq_of_window(sb.win_prefix, sb.n_seq, g0) : 0;
    }
}

template <int NT>
__device__ __forceinline__ uint64_t chunk_seq_table(const SeqBatch& sb, uint64_t s_lo, uint64_t g0, uint64_t g1, uint32_t* s_wseq,
                                                    uint32_t* s_scan) {
    constexpr int PER = BK_CH / NT;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // s_wseq arrives zeroed and the CTA synchronised (the callers clear it with their other per-chunk state)
    for (uint64_t q = s_lo + tid; q < sb.n_seq; q += NT) {
        const uint64_t p = __ldg(sb.win_prefix + q);
        if (p >= g1) break;
        const uint64_t pn = __ldg(sb.win_prefix + q + 1);
        if (pn > p && pn > g0) s_wseq[p > g0 ? (uint32_t)(p - g0) : 0u] = (uint32_t)(q - s_lo) + 1u;
    }
    __syncthreads();
    uint32_t v[PER];
    uint32_t run = 0;
#pragma unroll
    for (int q = 0; q < PER; ++q) { const uint32_t x = s_wseq[tid * PER + q]; run = x > run ? x : run; v[q] = run; }
    uint32_t x = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= (uint32_t)o && y > x) x = y;
    }
    if (lane == 31) s_scan[warp] = x;
    __syncthreads();
    uint32_t pre = 0;
    for (uint32_t w = 0; w < warp; ++w) { const uint32_t y = s_scan[w]; pre = y > pre ? y : pre; }
    uint32_t excl = __shfl_up_sync(0xFFFFFFFFu, x, 1);
    if (lane == 0) excl = 0;
    excl = excl > pre ? excl : pre;
#pragma unroll
    for (int q = 0; q < PER; ++q) s_wseq[tid * PER + q] = v[q] > excl ? v[q] : excl;
    __syncthreads();
    return s_lo;
}

constexpr int BK_EMIT_NT = 512;

template <int K, int H, bool PRE>
__global__ void __launch_bounds__(BK_EMIT_NT, 2) k_bucket_emit(const BucketParams bp) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    const CobsParams& p = bp.cp;
    const SeqBatch& sb = p.sb;
    const PageDesc pg = p.pages[0];
    const uint32_t k = K ? K : sb.k;
    const uint32_t h = H ? H : p.num_hashes;
    uint32_t* s_rec = reinterpret_cast<uint32_t*>(s_dyn);     // [n_buckets][cap]
    uint32_t* s_cnt = s_rec + (size_t)bp.n_buckets * bp.cap;  // [n_buckets] records offered (may exceed cap)
    uint32_t* s_ovf = s_cnt + bp.n_buckets;                   // [BK_CH / 32] windows with a probe that did not fit
    uint32_t* s_skp = s_ovf + BK_CH / 32;                     // [BK_CH / 32] windows that are not scored (non-ACGT, SKIP)
    uint32_t* s_wseq = s_skp + BK_CH / 32;                    // [BK_CH]
    uint32_t* s_scan = s_wseq + BK_CH;                        // [BK_EMIT_NT / 32]
    __shared__ unsigned long long s_chunk;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t total = __ldg(sb.win_prefix + sb.n_seq);
    const uint32_t nc_live = bucket_live_chunks(bp, total);
    const uint32_t rmask = (1u << bp.bshift) - 1u;
    const uint32_t cap = bp.cap, nb = bp.n_buckets;
    const uint32_t sig32 = (uint32_t)pg.sig_size;

    // the next chunk is requested while the current one is processed (thread 0 keeps the pending ticket)
    unsigned long long ticket = tid == 0 ? atomicAdd(bp.counter + 0, 1ULL) : 0ULL;
    for (;;) {
        if (tid == 0) s_chunk = ticket;
        for (uint32_t i = tid; i < nb + 2 * (BK_CH / 32) + BK_CH; i += BK_EMIT_NT) s_cnt[i] = 0;   // s_cnt, s_ovf, s_skp, s_wseq are contiguous
        __syncthreads();
        const uint64_t c = s_chunk;
        if (c >= nc_live) break;
        if (tid == 0) ticket = atomicAdd(bp.counter + 0, 1ULL);
        const uint64_t g0 = (bp.chunk0 + c) * BK_CH;
        const uint64_t g1 = g0 + BK_CH < total ? g0 + BK_CH : total;
        const uint32_t nwin = (uint32_t)(g1 - g0);
        if (PRE) {
            // row ids and the skip bitmap of this chunk were computed by the previous sub-batch's k_bucket_fetch
            if (tid < BK_CH / 32) s_skp[tid] = __ldg(bp.pre_skp + c * (BK_CH / 32) + tid);
            __syncthreads();
            const uint32_t* rows = bp.pre_rows + (size_t)c * h * BK_CH;
#pragma unroll 1
            for (uint32_t lid = tid; lid < nwin; lid += BK_EMIT_NT) {
                if ((s_skp[lid >> 5] >> (lid & 31)) & 1u) continue;
                bool over = false;
                auto put = [&](uint32_t j) {
                    const uint32_t row = ld_stream32(rows + j * BK_CH + lid);
                    const uint32_t b = row >> bp.bshift;
                    const uint32_t slot = atomicAdd(&s_cnt[b], 1u);
                    if (slot < cap) s_rec[b * cap + slot] = (lid << bp.bshift) | (row & rmask);
                    else over = true;
                };
                if (H) {
#pragma unroll
                    for (int j = 0; j < (H ? H : 1); ++j) put((uint32_t)j);
                } else {
                    for (uint32_t j = 0; j < h; ++j) put(j);
                }
                if (over) atomicOr(&s_ovf[lid >> 5], 1u << (lid & 31));
            }
        } else {
        const uint64_t s_lo = chunk_seq_table<BK_EMIT_NT>(sb, __ldg(bp.chunk_seq + bp.chunk0 + c), g0, g1, s_wseq, s_scan);
#pragma unroll 1
        for (uint32_t lid = tid; lid < nwin; lid += BK_EMIT_NT) {
            const uint64_t seq = s_lo + s_wseq[lid] - 1;
            const uint64_t pos = __ldg(sb.seq_begin + seq) - sb.base_shift + (g0 + lid - __ldg(sb.win_prefix + seq)) * sb.step;
            Term t;
            if (!cobs_term<K>(sb, pos, p.canonicalize != 0, p.policy, t)) {
                atomicOr(&s_skp[lid >> 5], 1u << (lid & 31));
                continue;
            }
            Xxh64Pre pre;
            xxh64_prepare(t, k, pre);
            bool over = false;
            auto emit = [&](uint32_t j) {   // signature_size < 2^29 on this path (bucket_geometry)
                const uint32_t row = mod_barrett_small(xxh64_finish(pre, k, (uint64_t)j), sig32, pg.magic);
                const uint32_t b = row >> bp.bshift;
                const uint32_t slot = atomicAdd(&s_cnt[b], 1u);
                if (slot < cap) s_rec[b * cap + slot] = (lid << bp.bshift) | (row & rmask);
                else over = true;
            };
            if (H) {
#pragma unroll
                for (int j = 0; j < (H ? H : 1); ++j) emit((uint32_t)j);
            } else {
                for (uint32_t j = 0; j < h; ++j) emit(j);
            }
            if (over) atomicOr(&s_ovf[lid >> 5], 1u << (lid & 31));
        }
        }
        __syncthreads();
        // blocks go out in whole 32-byte sectors (cap is a multiple of 8; the slack words are never read)
        for (uint32_t b = warp; b < nb; b += BK_EMIT_NT / 32) {
            const uint32_t n = s_cnt[b] < cap ? s_cnt[b] : cap;
            const uint32_t n4 = ((n + 7) & ~7u) / 4;
            const uint4* src = reinterpret_cast<const uint4*>(s_rec + b * cap);
            uint4* dst = reinterpret_cast<uint4*>(bp.rec + ((uint64_t)b * bp.nc + c) * cap);
            for (uint32_t i = lane; i < n4; i += 32) dst[i] = src[i];
        }
        for (uint32_t b = tid; b < nb; b += BK_EMIT_NT) {
            const uint16_t n = (uint16_t)(s_cnt[b] < cap ? s_cnt[b] : cap);
            bp.cnt_bc[(uint64_t)b * bp.nc + c] = n;
            bp.cnt_cb[c * nb + b] = n;
        }
        for (uint32_t i = tid; i < 2 * (BK_CH / 32); i += BK_EMIT_NT) bp.ovf[c * (2 * (BK_CH / 32)) + i] = s_ovf[i];   // ovf | skp
        __syncthreads();
    }
}

constexpr uint32_t BK_FETCH_SPAN = 8;   // consecutive (bucket, chunk) blocks per warp hand-out
constexpr uint32_t BK_HASH_UNIT = 256;  // windows per hash-ahead unit (8 warp rounds)

// first i >= lo with prefix[i + 1] > g, i.e. the sequence of flat window g, searching upwards from a known lower bound
__device__ __forceinline__ uint64_t seq_of_window_from(const uint64_t* __restrict__ prefix, uint64_t n, uint64_t lo, uint64_t g) {
    uint64_t step = 1, hi = lo + 1;
    while (hi < n && __ldg(prefix + hi) <= g) { lo = hi; step <<= 1; hi = lo + step; }
    if (hi > n) hi = n;
    while (hi - lo > 1) {
        const uint64_t mid = (lo + hi) >> 1;
        if (__ldg(prefix + mid) <= g) lo = mid; else hi = mid;
    }
    return lo;
}

// hash-ahead unit: row ids of BK_HASH_UNIT consecutive windows of the next sub-batch (one warp, 32 windows per round)
template <int K, int H>
__device__ __forceinline__ void bucket_hash_unit(const BucketParams& bp, uint64_t hu, uint64_t total, uint32_t lane) {
    const CobsParams& p = bp.cp;
    const SeqBatch& sb = p.sb;
    const uint32_t k = K ? K : sb.k;
    const uint32_t h = H ? H : p.num_hashes;
    constexpr uint32_t PER = BK_CH / BK_HASH_UNIT;
    const uint64_t cn = hu / PER;                                   // chunk of the next sub-batch
    const uint32_t l0 = (uint32_t)(hu % PER) * BK_HASH_UNIT;        // first local window of the unit
    const uint64_t c_glob = bp.next_chunk0 + cn;
    const uint64_t g_chunk = c_glob * BK_CH;
    if (g_chunk + l0 >= total) return;
    const PageDesc pg = p.pages[0];
    const uint32_t sig32 = (uint32_t)pg.sig_size;
    uint32_t* rows = bp.next_rows + (size_t)cn * h * BK_CH;
    uint64_t s_first = seq_of_window_from(sb.win_prefix, sb.n_seq, __ldg(bp.chunk_seq + c_glob), g_chunk + l0);
    for (uint32_t r = 0; r < BK_HASH_UNIT / 32; ++r) {
        const uint32_t lid = l0 + r * 32 + lane;
        const uint64_t g = g_chunk + lid;
        const bool has = g < total;
        // sequence of window g: prefix of the 32 sequences after s_first in the lanes, 5-step search over shuffles
        const uint64_t qi = s_first + 1 + lane;
        const uint64_t P = __ldg(sb.win_prefix + (qi < sb.n_seq ? qi : sb.n_seq));
        uint32_t lo = 0, hi = 32;                                   // smallest idx in [0, 32] with P[idx] > g (32 = none)
#pragma unroll
        for (int it = 0; it < 6; ++it) {
            const uint32_t mid = (lo + hi) >> 1;
            const uint64_t pm = __shfl_sync(0xFFFFFFFFu, P, mid & 31);
            if (lo < hi) { if (pm <= g) lo = mid + 1; else hi = mid; }
        }
        uint64_t seq = s_first + lo;
        if (has && lo == 32) seq = seq_of_window_from(sb.win_prefix, sb.n_seq, s_first + 32, g);   // > 32 sequences in one round
        bool ok = false;
        if (has) {
            const uint64_t pos = __ldg(sb.seq_begin + seq) - sb.base_shift + (g - __ldg(sb.win_prefix + seq)) * sb.step;
            Term t;
            ok = cobs_term<K>(sb, pos, p.canonicalize != 0, p.policy, t);
            if (ok) {
                Xxh64Pre pre;
                xxh64_prepare(t, k, pre);
                if (H) {
#pragma unroll
                    for (int j = 0; j < (H ? H : 1); ++j)
                        rows[j * BK_CH + lid] = mod_barrett_small(xxh64_finish(pre, k, (uint64_t)j), sig32, pg.magic);
                } else {
                    for (uint32_t j = 0; j < h; ++j)
                        rows[j * BK_CH + lid] = mod_barrett_small(xxh64_finish(pre, k, (uint64_t)j), sig32, pg.magic);
                }
            }
        }
        const uint32_t skip = __ballot_sync(0xFFFFFFFFu, has && !ok);
        if (lane == 0) bp.next_skp[cn * (BK_CH / 32) + (lid >> 5)] = skip;
        // the next round starts in (or after) the sequence of this round's last window
        const uint32_t last = 31 - __clz(__ballot_sync(0xFFFFFFFFu, has) | 1u);
        s_first = __shfl_sync(0xFFFFFFFFu, seq, last);
    }
}

// the hash-ahead work alone, for a side stream next to k_bucket_fetch / k_bucket_reduce of the previous sub-batch
template <int K, int H>
__global__ void __launch_bounds__(BK_NT) k_bucket_hash(const BucketParams bp) {
    const SeqBatch& sb = bp.cp.sb;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t total = __ldg(sb.win_prefix + sb.n_seq);
    const uint64_t first = bp.next_chunk0 * BK_CH;
    uint64_t live = total > first ? (total - first + BK_CH - 1) / BK_CH : 0;
    if (live > bp.next_nc) live = bp.next_nc;
    const uint64_t n_hash = live * (BK_CH / BK_HASH_UNIT);
    for (;;) {
        const uint64_t t = next_tile(bp.counter + 3, lane);
        if (t >= n_hash) break;
        bucket_hash_unit<K, H>(bp, t, total, lane);
    }
}

template <int K, int H, bool HASH>
__global__ void __launch_bounds__(BK_NT, HASH ? 4 : 6) k_bucket_fetch(const BucketParams bp) {
    const SeqBatch& sb = bp.cp.sb;
    const PageDesc pg = bp.cp.pages[0];
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t total = __ldg(sb.win_prefix + sb.n_seq);
    const uint32_t nc_live = bp.fetch_off ? 0u : bucket_live_chunks(bp, total);
    const uint64_t n_units = (uint64_t)bp.n_buckets * nc_live;    // unit = bucket * nc_live + chunk: bucket-major sweep
    const uint32_t rmask = (1u << bp.bshift) - 1u;
    const uint32_t cap = bp.cap;
    const uint64_t keep = l2_policy_evict_last();
    // the units of bucket b prefetch the rows of bucket b + 1 between them, so a sweep does not start on cold rows
    const uint64_t slice_lines = ((16ULL << bp.bshift) + 127) / 128;
    const uint32_t pf_per_unit = nc_live ? (uint32_t)((slice_lines + nc_live - 1) / nc_live) : 0;
    const uint64_t index_lines = (pg.sig_size * 16 + 127) / 128;
    // tickets interleave R fetch spans with one hash-ahead unit, so both queues drain together
    const uint64_t n_spans = (n_units + BK_FETCH_SPAN - 1) / BK_FETCH_SPAN;
    uint64_t n_hash = 0;
    if (HASH && bp.next_nc) {
        const uint64_t first = bp.next_chunk0 * BK_CH;
        uint64_t live = total > first ? (total - first + BK_CH - 1) / BK_CH : 0;
        if (live > bp.next_nc) live = bp.next_nc;
        n_hash = live * (BK_CH / BK_HASH_UNIT);
    }
    uint64_t R = n_hash ? (n_spans + n_hash / 2) / n_hash : 1;
    if (R == 0) R = 1;
    const uint64_t groups_f = (n_spans + R - 1) / R;
    const uint64_t n_tickets = HASH ? (groups_f > n_hash ? groups_f : n_hash) * (R + 1) : n_spans;
    for (;;) {
        const uint64_t t = next_tile(bp.counter + 1, lane);
        if (t >= n_tickets) break;
        uint64_t span = t;
        if (HASH) {
            const uint64_t grp = t / (R + 1), pos = t % (R + 1);
            if (pos == R) {
                if (grp < n_hash) bucket_hash_unit<K, H>(bp, grp, total, lane);
                continue;
            }
            span = grp * R + pos;
            if (span >= n_spans) continue;
        }
        const uint64_t u0 = span * BK_FETCH_SPAN;
        const uint64_t u1 = u0 + BK_FETCH_SPAN < n_units ? u0 + BK_FETCH_SPAN : n_units;
        for (uint64_t u = u0; u < u1; ++u) {
            const uint32_t b = (uint32_t)(u / nc_live), c = (uint32_t)(u % nc_live);
            const uint32_t n = __ldg(bp.cnt_bc + (uint64_t)b * bp.nc + c);
            const uint32_t* src = bp.rec + ((uint64_t)b * bp.nc + c) * cap;
            uint4* dst = bp.rows + ((uint64_t)c * bp.n_buckets + b) * cap;
            const uint8_t* base = pg.data + (((uint64_t)b << bp.bshift) * 16);
            if (bp.prefetch && b + 1 < bp.n_buckets) {
                for (uint32_t l = lane; l < pf_per_unit; l += 32) {
                    const uint64_t in_slice = (uint64_t)c * pf_per_unit + l;
                    const uint64_t line = (uint64_t)(b + 1) * slice_lines + in_slice;
                    if (in_slice < slice_lines && line < index_lines) prefetch_l2_keep(pg.data + line * 128, keep);
                }
            }
            for (uint32_t i0 = 0; i0 < n; i0 += 128) {
                uint32_t r[4];
                uint4 v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t i = i0 + q * 32 + lane;
                    r[q] = i < n ? ld_stream32(src + i) : 0u;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t i = i0 + q * 32 + lane;
                    if (i < n) v[q] = ldg128_keep(base + (uint64_t)(r[q] & rmask) * 16, keep);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t i = i0 + q * 32 + lane;
                    if (i < n) {
                        if (bp.pack_id) v[q].w = r[q] >> bp.bshift;
                        st_stream128(dst + i, v[q]);
                    }
                }
            }
        }
    }
}

// ----------------------------------------------------------------------------------------
// k_bucket_fetch with part of the gathers moved to the TMA unit.  The LSU path tops out at one L1TEX tag lookup per
// gathered row (~168 G rows/s); Blackwell's tile::gather4 tensor copy fetches four rows of a 2-D tensor by row index
// with one instruction and does not pass through L1TEX (profiles/microbench/gather_tma4.cu: 88 G rows/s alone,
// additive next to LSU gathers).  Per pass over a block of <= 128 records a lane takes TMAQ of its four records
// through gather4 (the quad leader issues one copy for the quad's four rows into a 64-byte shared-memory slot,
// completion on a per-warp mbarrier) and the others through ld.global.nc as before.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int TMAQ, int VAR>      // VAR bit 0: no L2 cache hint on the tensor copies; bit 1: one lane polls the mbarrier
__global__ void __launch_bounds__(BK_NT, 6) k_bucket_fetch_tma(const __grid_constant__ BucketParams bp) {
    __shared__ __align__(128) uint8_t s_slot[BK_NT / 32][TMAQ][8][128];   // [warp][q][quad] 64 bytes used of each 128
    __shared__ __align__(8) uint64_t s_bar[BK_NT / 32];
    const SeqBatch& sb = bp.cp.sb;
    const PageDesc pg = bp.cp.pages[0];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t total = __ldg(sb.win_prefix + sb.n_seq);
    const uint32_t nc_live = bucket_live_chunks(bp, total);
    const uint64_t n_units = (uint64_t)bp.n_buckets * nc_live;
    const uint32_t rmask = (1u << bp.bshift) - 1u;
    const uint32_t cap = bp.cap;
    const uint64_t keep = l2_policy_evict_last();
    const uint64_t slice_lines = ((16ULL << bp.bshift) + 127) / 128;
    const uint32_t pf_per_unit = nc_live ? (uint32_t)((slice_lines + nc_live - 1) / nc_live) : 0;
    const uint64_t index_lines = (pg.sig_size * 16 + 127) / 128;
    const uint32_t bar = smem_addr(&s_bar[warp]);
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    uint32_t phase = 0;
    constexpr int LSUQ = 4 - TMAQ;
    for (;;) {
        const uint64_t u0 = next_tile(bp.counter + 1, lane) * BK_FETCH_SPAN;
        if (u0 >= n_units) break;
        const uint64_t u1 = u0 + BK_FETCH_SPAN < n_units ? u0 + BK_FETCH_SPAN : n_units;
        for (uint64_t u = u0; u < u1; ++u) {
            const uint32_t b = (uint32_t)(u / nc_live), c = (uint32_t)(u % nc_live);
            const uint32_t n = __ldg(bp.cnt_bc + (uint64_t)b * bp.nc + c);
            const uint32_t* src = bp.rec + ((uint64_t)b * bp.nc + c) * cap;
            uint4* dst = bp.rows + ((uint64_t)c * bp.n_buckets + b) * cap;
            const uint32_t row0 = b << bp.bshift;
            const uint8_t* base = pg.data + (uint64_t)row0 * 16;
            if (bp.prefetch && b + 1 < bp.n_buckets) {
                for (uint32_t l = lane; l < pf_per_unit; l += 32) {
                    const uint64_t in_slice = (uint64_t)c * pf_per_unit + l;
                    const uint64_t line = (uint64_t)(b + 1) * slice_lines + in_slice;
                    if (in_slice < slice_lines && line < index_lines) prefetch_l2_keep(pg.data + line * 128, keep);
                }
            }
            for (uint32_t i0 = 0; i0 < n; i0 += 128) {
                uint32_t r[4];
                uint4 v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t i = i0 + q * 32 + lane;
                    r[q] = i < n ? ld_stream32(src + i) : 0u;
                }
                // ---- TMA part: records q >= LSUQ
                if (lane == 0)
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(TMAQ * 8 * 64) : "memory");
                __syncwarp();
#pragma unroll
                for (int q = LSUQ; q < 4; ++q) {
                    // slots past the block's end must not all fetch one row: every warp of the GPU sweeps the same bucket, and
                    // thousands of copies of one sector serialise in its L2 slice (measured: 8x slower fetch).  They
                    // re-fetch the lane's first record instead (or a lane-specific row), and the result is ignored.
                    const uint32_t iq = i0 + q * 32 + lane;
                    const uint32_t rr = iq < n ? r[q] : (i0 + lane < n ? r[0] : (threadIdx.x * 2654435761u + blockIdx.x * 40503u));
                    const int32_t my = (int32_t)(row0 + (rr & rmask));
                    const int32_t a0 = __shfl_sync(0xFFFFFFFFu, my, (lane & ~3u) + 0), a1 = __shfl_sync(0xFFFFFFFFu, my, (lane & ~3u) + 1),
                                  a2 = __shfl_sync(0xFFFFFFFFu, my, (lane & ~3u) + 2), a3 = __shfl_sync(0xFFFFFFFFu, my, (lane & ~3u) + 3);
                    if ((lane & 3u) == 0) {
                        if (VAR & 1)
                            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
                                         " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                                         :: "r"(smem_addr(&s_slot[warp][q - LSUQ][lane >> 2][0])), "l"(&bp.tmap), "r"(bar), "r"(0), "r"(a0), "r"(a1),
                                            "r"(a2), "r"(a3) : "memory");
                        else
                            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes.L2::cache_hint"
                                         " [%0], [%1, {%3, %4, %5, %6, %7}], [%2], %8;"
                                         :: "r"(smem_addr(&s_slot[warp][q - LSUQ][lane >> 2][0])), "l"(&bp.tmap), "r"(bar), "r"(0), "r"(a0), "r"(a1),
                                            "r"(a2), "r"(a3), "l"(keep) : "memory");
                    }
                }
                // ---- LSU part
#pragma unroll
                for (int q = 0; q < LSUQ; ++q) {
                    const uint32_t i = i0 + q * 32 + lane;
                    if (i < n) v[q] = ldg128_keep(base + (uint64_t)(r[q] & rmask) * 16, keep);
                }
                if (!(VAR & 2) || lane == 0) {
                    uint32_t done = 0;
                    while (!done)
                        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                     : "=r"(done) : "r"(bar), "r"(phase) : "memory");
                }
                phase ^= 1;
                __syncwarp();
#pragma unroll
                for (int q = LSUQ; q < 4; ++q)
                    v[q] = *reinterpret_cast<const uint4*>(&s_slot[warp][q - LSUQ][lane >> 2][(lane & 3u) * 16]);
                __syncwarp();          // every lane has read its slot before the next pass overwrites it
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t i = i0 + q * 32 + lane;
                    if (i < n) {
                        if (bp.pack_id) v[q].w = r[q] >> bp.bshift;
                        st_stream128(dst + i, v[q]);
                    }
                }
            }
        }
    }
}

// the rare flagged window: direct gathers, kept out of line so the streaming path stays lean
template <int K, int H>
__device__ __noinline__ void cobs_mask16_outline(const CobsParams* p, const PageDesc* pg, uint64_t pos, uint4* m) {
    *m = cobs_mask16<K, H>(*p, *pg, pos);
}

template <int K, int H, typename OutT, bool PACK>
__global__ void __launch_bounds__(BK_NT, 4) k_bucket_reduce(const __grid_constant__ BucketParams bp) {
    __shared__ __align__(16) uint32_t s_m[4][BK_CH];   // window masks, one array per 32 documents
    __shared__ uint32_t s_wseq[BK_CH];
    __shared__ uint32_t s_flag[2 * (BK_CH / 32)];      // ovf | skp bitmaps of the chunk
    __shared__ uint32_t s_scan[BK_NT / 32];
    __shared__ unsigned long long s_chunk;
    const CobsParams& p = bp.cp;
    const SeqBatch& sb = p.sb;
    const PageDesc pg = p.pages[0];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t total = __ldg(sb.win_prefix + sb.n_seq);
    const uint32_t nc_live = bucket_live_chunks(bp, total);
    const uint32_t cap = bp.cap, nb = bp.n_buckets;

    unsigned long long ticket = tid == 0 ? atomicAdd(bp.counter + 2, 1ULL) : 0ULL;
    for (;;) {
        if (tid == 0) s_chunk = ticket;
        {
            uint4* z = reinterpret_cast<uint4*>(&s_m[0][0]);
            const uint4 ones = make_uint4(~0u, ~0u, ~0u, ~0u);
            for (uint32_t i = tid; i < 4 * BK_CH / 4; i += BK_NT) z[i] = ones;
            for (uint32_t i = tid; i < BK_CH; i += BK_NT) s_wseq[i] = 0;
        }
        __syncthreads();
        const uint64_t c = s_chunk;
        if (c >= nc_live) break;
        if (tid == 0) ticket = atomicAdd(bp.counter + 2, 1ULL);
        const uint64_t g0 = (bp.chunk0 + c) * BK_CH;
        const uint64_t g1 = g0 + BK_CH < total ? g0 + BK_CH : total;
        const uint32_t nwin = (uint32_t)(g1 - g0);
        if (tid < 2 * (BK_CH / 32)) s_flag[tid] = __ldg(bp.ovf + c * (2 * (BK_CH / 32)) + tid);
        const uint64_t s_lo = chunk_seq_table<BK_NT>(sb, __ldg(bp.chunk_seq + bp.chunk0 + c), g0, g1, s_wseq, s_scan);

        // ---- AND of the fetched rows into the window masks (warp w takes buckets w, w + 8, ...)
        uint32_t myn = 0;
        {
            const uint32_t b = lane * (BK_NT / 32) + warp;
            if (b < nb) myn = __ldg(bp.cnt_cb + c * nb + b);
        }
        {
            uint32_t* const m0 = &s_m[0][0];
            const uint4* src = bp.rows + (c * nb + warp) * (uint64_t)cap + lane;              // this lane's slot of the block
            const uint32_t* rsrc = bp.rec + ((uint64_t)warp * bp.nc + c) * cap + lane;
            const uint64_t src_step = (uint64_t)(BK_NT / 32) * cap, rsrc_step = (uint64_t)(BK_NT / 32) * bp.nc * cap;
            for (uint32_t t = 0; t * (BK_NT / 32) + warp < nb; ++t, src += src_step, rsrc += rsrc_step) {
                const uint32_t n = __shfl_sync(0xFFFFFFFFu, myn, t);
                for (uint32_t i0 = lane; i0 < n; i0 += 128) {       // i0 = this lane's first record of the pass
                    const uint32_t left = n - i0;                   // records from i0 on; lane takes i0, +32, +64, +96
                    uint4 v[4];
                    uint32_t id[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if ((uint32_t)(q * 32) < left) {
                            v[q] = ld_stream128(src + (i0 - lane) + q * 32);
                            if (!PACK) id[q] = ld_stream32(rsrc + (i0 - lane) + q * 32) >> bp.bshift;
                        }
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if ((uint32_t)(q * 32) < left) {
                            uint32_t* a = m0 + ((PACK ? v[q].w : id[q]) & (BK_CH - 1));
                            atomicAnd(a, v[q].x);
                            atomicAnd(a + BK_CH, v[q].y);
                            atomicAnd(a + 2 * BK_CH, v[q].z);
                            if (!PACK) atomicAnd(a + 3 * BK_CH, v[q].w);
                        }
                }
            }
        }
        __syncthreads();

        // ---- per-sequence document counts: warp w owns windows [w * BK_WARP_WIN, ...) of the chunk; the lanes of a
        // round that belong to one sequence are counted together, a sequence is flushed after its last window here
        const uint32_t w_lo = warp * BK_WARP_WIN;
        const uint32_t w_hi = w_lo + BK_WARP_WIN < nwin ? w_lo + BK_WARP_WIN : nwin;
        uint32_t cnt[4] = {0, 0, 0, 0};
        for (uint32_t r0 = w_lo; r0 < w_hi; r0 += 32) {
            const uint32_t lid = r0 + lane;
            const bool has = lid < w_hi;
            const uint32_t sq = has ? s_wseq[lid] : 0u;
            uint4 m = make_uint4(0, 0, 0, 0);
            if (has) {
                const uint32_t bit = 1u << (lid & 31);
                if (s_flag[lid >> 5] & bit) {
                    const uint64_t seq = s_lo + sq - 1;
                    const uint64_t pos = __ldg(sb.seq_begin + seq) - sb.base_shift + (g0 + lid - __ldg(sb.win_prefix + seq)) * sb.step;
                    cobs_mask16_outline<K, H>(&bp.cp, bp.cp.pages, pos, &m);
                } else if (!(s_flag[BK_CH / 32 + (lid >> 5)] & bit)) {
                    m = make_uint4(s_m[0][lid], s_m[1][lid], s_m[2][lid], PACK ? 0u : s_m[3][lid]);
                }
            }
            uint32_t rem = __ballot_sync(0xFFFFFFFFu, has);
            while (rem) {
                const uint32_t cur = __shfl_sync(0xFFFFFFFFu, sq, __ffs(rem) - 1);
                const uint32_t segmask = __ballot_sync(0xFFFFFFFFu, has && sq == cur);
                narrow_accum(lane, segmask, m, cnt);
                rem &= ~segmask;
                const uint32_t nxt = r0 + (32 - __clz(segmask));   // the window after this sequence's last lane of the round
                if (!(nxt < w_hi && s_wseq[nxt] == cur)) {
                    const uint64_t seq = s_lo + cur - 1;
                    const uint64_t ps = __ldg(sb.win_prefix + seq), pe = __ldg(sb.win_prefix + seq + 1);
                    narrow_flush<OutT>(p, pg, lane, seq, ps >= g0 + w_lo && pe <= g0 + w_hi, pe - ps, cnt);
                }
            }
        }
        __syncthreads();
    }
}

// ----------------------------------------------------------------------------------------
// COBS, wide rows.  One CTA per (sequence, chunk of <= WIDE_CHUNK windows); grid.y walks column
// blocks of <= WIDE_MAX_CHUNKS 16-byte chunks of one page.
// ----------------------------------------------------------------------------------------
constexpr int WIDE_NT = 256;
constexpr int WIDE_CHUNK = 248;         // windows per work item (vertical counters hold 8 bits)
constexpr int WIDE_MAX_COLS = 128;      // 16-byte column chunks per column block (16384 documents)

struct ColBlock {
    uint32_t page;
    uint32_t c0;       // first 16-byte chunk of the row
    uint32_t n_cols;   // chunks in this block
    uint32_t n_docs;   // documents in this block (<= n_cols * 128)
};

struct WideParams {
    CobsParams cp;
    const ColBlock* blocks;
    const uint64_t* chunk_prefix;  // [n_seq + 1] exclusive scan of work items per sequence
};

template <int K, int H, typename OutT>
__global__ void __launch_bounds__(WIDE_NT) k_cobs_wide(const WideParams wp) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    const CobsParams& p = wp.cp;
    const SeqBatch& sb = p.sb;
    const uint32_t k = K ? K : sb.k;
    const uint32_t h = H ? H : p.num_hashes;
    const ColBlock cb = wp.blocks[blockIdx.y];
    const PageDesc pg = p.pages[cb.page];
    uint64_t* s_rows = reinterpret_cast<uint64_t*>(s_dyn);                    // [WIDE_CHUNK * h] byte offsets
    uint32_t* s_cnt = reinterpret_cast<uint32_t*>(s_dyn + (size_t)WIDE_CHUNK * h * 8);  // [n_cols * 128]
    __shared__ uint64_t s_seq, s_chunk, s_item;
    __shared__ uint32_t s_nvalid;

    const int tid = threadIdx.x;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    constexpr uint32_t NWARP = WIDE_NT / 32;
    OutT* out = reinterpret_cast<OutT*>(p.out);

    const uint32_t C = cb.n_cols;
    // column rounds: full rounds of 32 chunks (one window per warp pass) and a tail round of C % 32 chunks in which
    // the warp handles 32 / pow2ceil(tail) windows at once, so lanes stay busy when C is not a multiple of 32
    const uint32_t n_full = C / 32, tail = C % 32;
    uint32_t lpw_tail = 1; while (lpw_tail < tail) lpw_tail <<= 1;
    const uint32_t n_round = n_full + (tail ? 1u : 0u);
    const uint32_t wsplit = NWARP;                             // window slices: every warp gets one unit per round
    const uint32_t n_unit = n_round * wsplit;

    const uint64_t total_items = __ldg(wp.chunk_prefix + sb.n_seq);
    for (;;) {
        if (tid == 0) {
            uint64_t item = atomicAdd(sb.tile_counter + blockIdx.y, 1ULL);
            s_item = item;
            s_nvalid = 0;
            if (item < total_items) {
                uint64_t s = seq_of_window(wp.chunk_prefix, sb.n_seq, item);
                s_seq = s;
                s_chunk = item - __ldg(wp.chunk_prefix + s);
            }
        }
        __syncthreads();
        if (s_item >= total_items) break;
        const uint64_t seq = s_seq;
        const uint64_t sbeg = __ldg(sb.seq_begin + seq), send = __ldg(sb.seq_end + seq);
        const uint64_t nw_seq = windows_of(sbeg, send, sb.base_shift, sb.n_bases, k, sb.step);
        const uint64_t w0 = s_chunk * WIDE_CHUNK;
        const uint32_t nwin = (uint32_t)(nw_seq - w0 < (uint64_t)WIDE_CHUNK ? nw_seq - w0 : (uint64_t)WIDE_CHUNK);
        const bool complete = nw_seq <= (uint64_t)WIDE_CHUNK;
        const bool nosat = nw_seq <= (uint64_t)out_max<OutT>();

        // ---- phase 1: row byte offsets of the valid windows of the item, compacted (order is irrelevant for counts)
        for (uint32_t lw = tid; lw < nwin; lw += WIDE_NT) {
            uint64_t pos = sbeg - sb.base_shift + (w0 + lw) * sb.step;
            Term t;
            if (cobs_term<K>(sb, pos, p.canonicalize != 0, p.policy, t)) {
                Xxh64Pre pre;
                xxh64_prepare(t, k, pre);
                uint32_t slot_w = atomicAdd(&s_nvalid, 1u);
                for (uint32_t j = 0; j < h; ++j) {
                    uint64_t hv = xxh64_finish(pre, k, (uint64_t)j);
                    s_rows[slot_w * h + j] = mod_barrett(hv, pg.sig_size, pg.magic) * pg.row_stride;
                }
            }
        }
        for (uint32_t d = tid; d < C * 128; d += WIDE_NT) s_cnt[d] = 0;
        __syncthreads();
        const uint32_t nv = s_nvalid;

        // ---- phase 2: gather + AND + bit-plane counters.  Unit = (column round, window slice); every warp gets
        // n_round units.  Two windows (2h loads) are in flight per lane; all-zero masks (the common case) skip
        // the counter update; 4 bit planes are spilled into the shared counters every 15 updates.
        for (uint32_t u = warp; u < n_unit; u += NWARP) {
            const uint32_t r = u / wsplit, ws = u % wsplit;
            const uint32_t lpw = r < n_full ? 32u : lpw_tail;     // lanes per window in this round
            const uint32_t slots = 32 / lpw;                      // windows the warp handles at once
            const uint32_t slot = lane / lpw, cl = lane % lpw;
            const uint32_t col = r * 32 + cl;
            const bool active = col < C;
            const uint8_t* colbase = pg.data + (size_t)(cb.c0 + col) * 16;
            uint32_t pl[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) pl[a][b] = 0;
            uint32_t pending = 0;
            auto spill = [&]() {
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    uint32_t any = pl[0][b] | pl[1][b] | pl[2][b] | pl[3][b];
                    while (any) {
                        uint32_t bit = __ffs(any) - 1; any &= any - 1;
                        uint32_t c = ((pl[0][b] >> bit) & 1u) | (((pl[1][b] >> bit) & 1u) << 1) |
                                     (((pl[2][b] >> bit) & 1u) << 2) | (((pl[3][b] >> bit) & 1u) << 3);
                        atomicAdd(&s_cnt[col * 128 + b * 32 + bit], c);
                    }
                    pl[0][b] = pl[1][b] = pl[2][b] = pl[3][b] = 0;
                }
                pending = 0;
            };
            auto add_mask = [&](const uint4& m) {
                if ((m.x | m.y | m.z | m.w) == 0) return;
                uint32_t carry[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        uint32_t tcar = pl[a][b] & carry[b];
                        pl[a][b] ^= carry[b];
                        carry[b] = tcar;
                    }
                if (++pending == 15) spill();
            };
            auto gather = [&](uint32_t w) -> uint4 {     // AND of the h rows of window w (this lane's 16-byte column)
                uint4 m = ldg128(colbase + s_rows[w * h]);
                if (H) {
                    uint4 v[H ? H : 1];
#pragma unroll
                    for (int j = 1; j < (H ? H : 1); ++j) v[j] = ldg128(colbase + s_rows[w * h + j]);
#pragma unroll
                    for (int j = 1; j < (H ? H : 1); ++j) { m.x &= v[j].x; m.y &= v[j].y; m.z &= v[j].z; m.w &= v[j].w; }
                } else {
                    for (uint32_t j = 1; j < h; ++j) {
                        uint4 v = ldg128(colbase + s_rows[w * h + j]);
                        m.x &= v.x; m.y &= v.y; m.z &= v.z; m.w &= v.w;
                    }
                }
                return m;
            };
            if (active) {
                const uint32_t stride = wsplit * slots;
                uint32_t w = ws * slots + slot;
                for (; w + stride < nv; w += 2 * stride) {
                    uint4 m0 = gather(w), m1 = gather(w + stride);
                    add_mask(m0);
                    add_mask(m1);
                }
                if (w < nv) add_mask(gather(w));
                spill();
            }
        }
        __syncthreads();

        // ---- phase 3: coalesced write of the block's document counts
        OutT* row = out + (p.seq0 + seq) * p.ld + pg.doc_off + (size_t)cb.c0 * 128;
        for (uint32_t d = tid; d < cb.n_docs; d += WIDE_NT) {
            uint32_t v = s_cnt[d];
            if (complete) out_store<OutT>(row + d, v); else out_add<OutT>(row + d, v, nosat);
        }
        __syncthreads();
    }
}

// ----------------------------------------------------------------------------------------
// COBS, mid-width rows: one page, row_stride = 32 / 64 / 128 bytes (129 .. 1024 documents — species models of large
// genera, narrow column shards).  k_cobs_wide gives such rows a CTA per (sequence, window chunk): three CTA barriers
// around ~4 windows per warp for a 150-bp read, and lanes idle whenever the row has fewer than 32 column chunks.
// Here the warp keeps k_cobs_narrow's barrier-free walk over the flat window space and changes roles between the two
// halves of a round:
//   hash     one lane per window: canonical k-mer -> XXH64 x h -> row ids (registers)
//   gather   LPW = row_stride / 16 lanes per window, 32 / LPW windows at a time: the row ids travel by shuffle, every
//            lane loads its 16-byte column chunk of the h rows (a row is one 128-byte DRAM line or an aligned part of
//            one), ANDs them and adds the mask to 4 vertical bit-plane counters; for h > 5 a lane whose AND is already
//            empty after five rows skips the rest
// Bit planes spill into a warp-private shared-memory slice of LPW x 128 counters every 15 masks and when the sequence
// ends; the non-zero counters are written (or added, for a sequence that straddles warp tiles) to the zeroed output.
// ----------------------------------------------------------------------------------------
constexpr int MID_NT = 256;
constexpr int MID_MAXH = 8;             // row ids a lane keeps in registers (generic instantiation: h <= 8)

template <int K, int H, typename OutT, int LPW, int OCC>
__global__ void __launch_bounds__(MID_NT, OCC) k_cobs_mid(const CobsParams p) {
    __shared__ uint32_t s_mid_cnt[MID_NT / 32][LPW * 128];
    constexpr uint32_t FULL = 0xFFFFFFFFu;
    constexpr uint32_t SLOTS = 32 / LPW;          // windows the warp gathers at once
    constexpr int HH = H ? H : MID_MAXH;
    constexpr int H1 = HH > 5 ? XS_EARLY_EXIT_AFTER : HH;
    const SeqBatch& sb = p.sb;
    const PageDesc pg = p.pages[0];
    const uint32_t k = K ? K : sb.k;
    const uint32_t h = H ? H : p.num_hashes;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t slot = lane / LPW, cl = lane % LPW;
    uint32_t* my = s_mid_cnt[warp];
    for (uint32_t i = lane; i < LPW * 128; i += 32) my[i] = 0;
    __syncwarp();
    const uint8_t* colbase = pg.data + cl * 16;
    OutT* out = reinterpret_cast<OutT*>(p.out);

    const uint64_t n_warps = ((uint64_t)gridDim.x * MID_NT) >> 5;
    const uint64_t total = __ldg(sb.win_prefix + sb.n_seq);
    const uint64_t W = walk_tile_windows(total, n_warps);
    const uint64_t n_tiles = (total + W - 1) / W;

    uint32_t pl[4][4];                            // [plane][word]: vertical counters of this lane's 128 documents
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) pl[a][b] = 0;
    uint32_t pending = 0;
    auto spill = [&]() {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            uint32_t any = pl[0][b] | pl[1][b] | pl[2][b] | pl[3][b];
            while (any) {
                const uint32_t bit = __ffs(any) - 1; any &= any - 1;
                const uint32_t c = ((pl[0][b] >> bit) & 1u) | (((pl[1][b] >> bit) & 1u) << 1) |
                                   (((pl[2][b] >> bit) & 1u) << 2) | (((pl[3][b] >> bit) & 1u) << 3);
                atomicAdd(&my[cl * 128 + b * 32 + bit], c);      // the SLOTS lanes of one column chunk share counters
            }
            pl[0][b] = pl[1][b] = pl[2][b] = pl[3][b] = 0;
        }
        pending = 0;
    };
    auto add_mask = [&](const uint4& m) {
        if ((m.x | m.y | m.z | m.w) == 0) return;
        uint32_t carry[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const uint32_t tc = pl[a][b] & carry[b];
                pl[a][b] ^= carry[b];
                carry[b] = tc;
            }
        if (++pending == 15) spill();
    };

    for (;;) {
        const uint64_t tile = next_tile(sb.tile_counter, lane);
        if (tile >= n_tiles) break;
        const uint64_t t0 = tile * W, t1 = t0 + W < total ? t0 + W : total;
        uint32_t r[HH];                           // row ids of this lane's window
        int valid = 0;
        warp_walk(
            sb, t0, t1, lane,
            [&](bool has, uint64_t pos, uint64_t) {
                valid = 0;
                if (has) {
                    Term t;
                    if (cobs_term<K>(sb, pos, p.canonicalize != 0, p.policy, t)) {
                        valid = 1;
                        Xxh64Pre pre;
                        xxh64_prepare(t, k, pre);
#pragma unroll
                        for (int j = 0; j < HH; ++j)
                            r[j] = (uint32_t)j < h ? (uint32_t)mod_barrett(xxh64_finish(pre, k, (uint64_t)j), pg.sig_size, pg.magic) : 0u;
                    }
                }
            },
            [&](uint32_t segmask) {
                const uint32_t first = __ffs(segmask) - 1, count = __popc(segmask);
                for (uint32_t sub = 0; sub < count; sub += SLOTS) {
                    const bool in = sub + slot < count;
                    const uint32_t src = in ? first + sub + slot : first;
                    const bool ok = __shfl_sync(FULL, valid, src) != 0 && in;
                    uint32_t rj[HH];
#pragma unroll
                    for (int j = 0; j < HH; ++j) rj[j] = __shfl_sync(FULL, r[j], src);
                    if (!ok) continue;
                    uint4 m = make_uint4(~0u, ~0u, ~0u, ~0u);
                    uint4 v[H1];
#pragma unroll
                    for (int j = 0; j < H1; ++j)
                        if (H || (uint32_t)j < h) v[j] = ldg128(colbase + (size_t)rj[j] * (LPW * 16));
#pragma unroll
                    for (int j = 0; j < H1; ++j)
                        if (H || (uint32_t)j < h) { m.x &= v[j].x; m.y &= v[j].y; m.z &= v[j].z; m.w &= v[j].w; }
                    if (H1 < HH && ((m.x | m.y | m.z | m.w) != 0 || p.all_rows)) {
#pragma unroll
                        for (int j = H1; j < HH; ++j)
                            if (H || (uint32_t)j < h) {
                                const uint4 w = ldg128(colbase + (size_t)rj[j] * (LPW * 16));
                                m.x &= w.x; m.y &= w.y; m.z &= w.z; m.w &= w.w;
                            }
                    }
                    add_mask(m);
                }
            },
            [&](uint64_t seq, bool complete, uint64_t nwin) {
                spill();
                __syncwarp();
                const bool nosat = nwin <= (uint64_t)out_max<OutT>();
                OutT* row = out + (p.seq0 + seq) * p.ld + pg.doc_off;
                for (uint32_t i = lane; i < LPW * 128; i += 32) {
                    const uint32_t c = my[i];
                    if (c) {
                        if (i < pg.n_docs) {
                            if (complete) out_store<OutT>(row + i, c); else out_add<OutT>(row + i, c, nosat);
                        }
                        my[i] = 0;
                    }
                }
                __syncwarp();
            });
    }
}

// ----------------------------------------------------------------------------------------
// Bloom
// ----------------------------------------------------------------------------------------
struct BloomParams {
    SeqBatch sb;
    const uint8_t* bits;
    uint64_t n_bits;
    uint64_t magic;      // floor(2^64 / n_bits)
    uint32_t k_hashes;
    uint32_t literal;    // 1: hash the window bytes as they are (rbloom's `obj in bf` on a caller-supplied k-mer)
    uint32_t* out;       // [n_seq]
    uint64_t seq0;
    uint64_t win_begin;  // k_bloom: first flat window to score (the bucketed path's tail launch), normally 0
    // device-side choice between k_bloom and the bucketed kernels (both are enqueued, the one not chosen returns at
    // once): adapt = {windows sampled, members among them} from k_bloom_sample, nullptr = no choice to make
    const unsigned long long* adapt;
    uint32_t adapt_pct;  // bucketed when members * 100 >= adapt_pct * sampled
};

__device__ __forceinline__ bool bloom_bucketed_chosen(const BloomParams& p) {
    if (!p.adapt) return true;
    const unsigned long long n = p.adapt[0], m = p.adapt[1];
    return m * 100ULL >= (unsigned long long)p.adapt_pct * n;
}

constexpr int BLOOM_NT = 256;

__device__ __forceinline__ uint32_t ldg_byte(const uint8_t* p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ bool bloom_member(const BloomParams& p, uint64_t h0) {
    uint64_t hi = 0, lo = h0;
    // the first two probes go out together (most non-members fail one of them), the rest
    // exit early like rbloom's iterator does
    uint64_t i0 = mod_barrett(lcg_next(hi, lo), p.n_bits, p.magic);
    if (p.k_hashes == 0) return true;
    if (p.k_hashes == 1) return (ldg_byte(p.bits + (i0 >> 3)) >> (i0 & 7)) & 1u;
    uint64_t i1 = mod_barrett(lcg_next(hi, lo), p.n_bits, p.magic);
    uint32_t b0 = ldg_byte(p.bits + (i0 >> 3)), b1 = ldg_byte(p.bits + (i1 >> 3));
    if (!((b0 >> (i0 & 7)) & (b1 >> (i1 & 7)) & 1u)) return false;
    for (uint32_t j = 2; j < p.k_hashes; ++j) {
        uint64_t ix = mod_barrett(lcg_next(hi, lo), p.n_bits, p.magic);
        if (!((ldg_byte(p.bits + (ix >> 3)) >> (ix & 7)) & 1u)) return false;
    }
    return true;
}

template <int K>
__global__ void __launch_bounds__(BLOOM_NT, 4) k_bloom(const BloomParams p) {
    const SeqBatch& sb = p.sb;
    const uint32_t k = K ? K : sb.k;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t n_warps = ((uint64_t)gridDim.x * BLOOM_NT) >> 5;

    const uint64_t total = __ldg(sb.win_prefix + sb.n_seq);
    // the tail launch of the bucketed path scores the windows from win_begin on, or all of them when the sampled
    // member fraction said the bucketed kernels should not run
    const uint64_t begin = (p.adapt && !bloom_bucketed_chosen(p)) ? 0 : p.win_begin;
    const uint64_t first = begin < total ? begin : total;
    const uint64_t W = walk_tile_windows(total - first, n_warps);
    const uint64_t n_tiles = (total - first + W - 1) / W;
    for (;;) {
        const uint64_t tile = next_tile(sb.tile_counter, lane);
        if (tile >= n_tiles) break;
        const uint64_t t0 = first + tile * W, t1 = t0 + W < total ? t0 + W : total;
        uint32_t hits = 0;    // warp-uniform: hits of the current sequence inside this tile
        uint32_t bal = 0;
        warp_walk(
            sb, t0, t1, lane,
            [&](bool has, uint64_t pos, uint64_t) {
                bool hit = false;
                if (has) {
                    Term t;
                    if (p.literal) literal_term(sb.bases, pos, k, BioComp(), false, t);
                    else bloom_term<K>(sb, pos, t);
                    hit = bloom_member(p, xxh3_64(t, k));
                }
                bal = __ballot_sync(0xFFFFFFFFu, hit);
            },
            [&](uint32_t segmask) { hits += __popc(bal & segmask); },
            [&](uint64_t seq, bool complete, uint64_t) {
                if (lane == 0 && hits) {
                    uint32_t* o = p.out + p.seq0 + seq;
                    if (complete) *o = hits; else atomicAdd(o, hits);
                }
                hits = 0;
            });
    }
}

// ----------------------------------------------------------------------------------------
// Bloom, bucketed probing: the scheme of k_bucket_emit / fetch / reduce for the single-filter model.  A probe of
// k_bloom costs a 128-byte DRAM fetch for one bit; here the k_hashes probes of every window are grouped by 16 MB ranges
// of the bit array and served from L2.  All k_hashes probes are emitted (no early exit), which still wins when most
// windows are members or the batch is large.  Record = bit index in the bucket (u32) + window in chunk (u16); fetch adds
// one result byte; reduce marks the windows with a failed probe and counts the others per sequence.
// ----------------------------------------------------------------------------------------
struct BloomBucketParams {
    BloomParams bl;
    uint32_t* pos;        // [n_buckets][nc][cap]  bit index within the bucket
    uint16_t* wid;        // [n_buckets][nc][cap]  window in chunk
    uint8_t* res;         // [n_buckets][nc][cap]  the probed bit (k_bbucket_fetch)
    uint16_t* cnt_bc;     // [n_buckets][nc]       records per block
    uint32_t* ovf;        // [nc][BK_CH / 32]      windows to score with direct probes
    const uint64_t* chunk_seq;
    unsigned long long* counter;   // [3]
    uint64_t chunk0;
    uint32_t nc, n_buckets, bshift, cap, prefetch;   // bshift = log2 bits per bucket
};

// member fraction of every `stride`-th window of the batch (full membership test with early exit, like k_bloom)
template <int K>
__global__ void __launch_bounds__(256) k_bloom_sample(const BloomParams p, uint64_t stride, unsigned long long* __restrict__ counts) {
    const SeqBatch& sb = p.sb;
    const uint32_t k = K ? K : sb.k;
    const uint64_t total = __ldg(sb.win_prefix + sb.n_seq);
    uint32_t n = 0, m = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i * stride < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t g = i * stride;
        const uint64_t q = seq_of_window(sb.win_prefix, sb.n_seq, g);
        const uint64_t pos = __ldg(sb.seq_begin + q) - sb.base_shift + (g - __ldg(sb.win_prefix + q)) * sb.step;
        Term t;
        bloom_term<K>(sb, pos, t);
        ++n;
        m += bloom_member(p, xxh3_64(t, k)) ? 1u : 0u;
    }
    n = __reduce_add_sync(0xFFFFFFFFu, n);
    m = __reduce_add_sync(0xFFFFFFFFu, m);
    if ((threadIdx.x & 31) == 0 && n) { atomicAdd(counts, (unsigned long long)n); atomicAdd(counts + 1, (unsigned long long)m); }
}

__device__ __forceinline__ uint32_t bbucket_live_chunks(const BloomBucketParams& bp, uint64_t total) {
    if (!bloom_bucketed_chosen(bp.bl)) return 0;          // the batch goes through k_bloom
    const uint64_t first = bp.chunk0 * BK_CH;
    if (total <= first) return 0;
    const uint64_t n = (total - first + BK_CH - 1) / BK_CH;
    return n < bp.nc ? (uint32_t)n : bp.nc;
}

template <int K, int KH>   // KH: number of hash functions when known at compile time (6 for fpr 0.01), else 0
__global__ void __launch_bounds__(BK_EMIT_NT, 2) k_bbucket_emit(const BloomBucketParams bp) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    const BloomParams& p = bp.bl;
    const SeqBatch& sb = p.sb;
    const uint32_t k = K ? K : sb.k;
    const uint32_t cap = bp.cap, nb = bp.n_buckets;
    const uint32_t nb4 = (nb + 3) & ~3u;                                  // keeps s_wid 16-byte aligned
    uint32_t* s_pos = reinterpret_cast<uint32_t*>(s_dyn);                 // [nb][cap]
    uint32_t* s_cnt = s_pos + (size_t)nb * cap;                           // [nb4]
    uint32_t* s_ovf = s_cnt + nb4;                                        // [BK_CH / 32]
    uint32_t* s_wseq = s_ovf + BK_CH / 32;                                // [BK_CH]
    uint32_t* s_scan = s_wseq + BK_CH;                                    // [BK_EMIT_NT / 32]
    uint16_t* s_wid = reinterpret_cast<uint16_t*>(s_scan + BK_EMIT_NT / 32);   // [nb][cap]
    __shared__ unsigned long long s_chunk;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t total = __ldg(sb.win_prefix + sb.n_seq);
    const uint32_t nc_live = bbucket_live_chunks(bp, total);
    const uint64_t bmask = (1ULL << bp.bshift) - 1ULL;

    unsigned long long ticket = tid == 0 ? atomicAdd(bp.counter + 0, 1ULL) : 0ULL;
    for (;;) {
        if (tid == 0) s_chunk = ticket;
        for (uint32_t i = tid; i < nb4 + BK_CH / 32 + BK_CH; i += BK_EMIT_NT) s_cnt[i] = 0;   // s_cnt, s_ovf, s_wseq are contiguous
        __syncthreads();
        const uint64_t c = s_chunk;
        if (c >= nc_live) break;
        if (tid == 0) ticket = atomicAdd(bp.counter + 0, 1ULL);
        const uint64_t g0 = (bp.chunk0 + c) * BK_CH;
        const uint64_t g1 = g0 + BK_CH < total ? g0 + BK_CH : total;
        const uint32_t nwin = (uint32_t)(g1 - g0);
        const uint64_t s_lo = chunk_seq_table<BK_EMIT_NT>(sb, __ldg(bp.chunk_seq + bp.chunk0 + c), g0, g1, s_wseq, s_scan);
#pragma unroll 1
        for (uint32_t lid = tid; lid < nwin; lid += BK_EMIT_NT) {
            const uint64_t seq = s_lo + s_wseq[lid] - 1;
            const uint64_t pos = __ldg(sb.seq_begin + seq) - sb.base_shift + (g0 + lid - __ldg(sb.win_prefix + seq)) * sb.step;
            Term t;
            bloom_term<K>(sb, pos, t);
            uint64_t hi = 0, lo = xxh3_64(t, k);
            bool over = false;
            auto probe = [&]() {
                const uint64_t ix = mod_barrett(lcg_next(hi, lo), p.n_bits, p.magic);
                const uint32_t b = (uint32_t)(ix >> bp.bshift);
                const uint32_t slot = atomicAdd(&s_cnt[b], 1u);
                if (slot < cap) { s_pos[b * cap + slot] = (uint32_t)(ix & bmask); s_wid[b * cap + slot] = (uint16_t)lid; }
                else over = true;
            };
            if (KH) {
#pragma unroll
                for (int j = 0; j < (KH ? KH : 1); ++j) probe();
            } else {
                for (uint32_t j = 0; j < p.k_hashes; ++j) probe();
            }
            if (over) atomicOr(&s_ovf[lid >> 5], 1u << (lid & 31));
        }
        __syncthreads();
        // blocks go out in whole sectors (cap is a multiple of 16: 64 bytes of positions, 32 bytes of window ids)
        for (uint32_t b = warp; b < nb; b += BK_EMIT_NT / 32) {
            const uint32_t n = s_cnt[b] < cap ? s_cnt[b] : cap;
            const uint32_t n16 = (n + 15) & ~15u;
            const uint64_t blk = ((uint64_t)b * bp.nc + c) * cap;
            const uint4* src = reinterpret_cast<const uint4*>(s_pos + b * cap);
            uint4* dst = reinterpret_cast<uint4*>(bp.pos + blk);
            for (uint32_t i = lane; i < n16 / 4; i += 32) dst[i] = src[i];
            const uint4* srcw = reinterpret_cast<const uint4*>(s_wid + b * cap);
            uint4* dstw = reinterpret_cast<uint4*>(bp.wid + blk);
            for (uint32_t i = lane; i < n16 / 8; i += 32) dstw[i] = srcw[i];
        }
        for (uint32_t b = tid; b < nb; b += BK_EMIT_NT) bp.cnt_bc[(uint64_t)b * bp.nc + c] = (uint16_t)(s_cnt[b] < cap ? s_cnt[b] : cap);
        for (uint32_t i = tid; i < BK_CH / 32; i += BK_EMIT_NT) bp.ovf[c * (BK_CH / 32) + i] = s_ovf[i];
        __syncthreads();
    }
}

__device__ __forceinline__ uint32_t ldg_byte_keep(const uint8_t* p, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}

__global__ void __launch_bounds__(BK_NT) k_bbucket_fetch(const BloomBucketParams bp) {
    const BloomParams& p = bp.bl;
    const SeqBatch& sb = p.sb;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t total = __ldg(sb.win_prefix + sb.n_seq);
    const uint32_t nc_live = bbucket_live_chunks(bp, total);
    const uint64_t n_units = (uint64_t)bp.n_buckets * nc_live;    // unit = bucket * nc_live + chunk: bucket-major sweep
    const uint32_t cap = bp.cap;
    const uint64_t keep = l2_policy_evict_last();
    const uint64_t slice_bytes = (1ULL << bp.bshift) / 8;
    const uint64_t slice_lines = (slice_bytes + 127) / 128;
    const uint32_t pf_per_unit = nc_live ? (uint32_t)((slice_lines + nc_live - 1) / nc_live) : 0;
    const uint64_t array_lines = (p.n_bits / 8 + 127) / 128;
    for (;;) {
        const uint64_t u0 = next_tile(bp.counter + 1, lane) * BK_FETCH_SPAN;
        if (u0 >= n_units) break;
        const uint64_t u1 = u0 + BK_FETCH_SPAN < n_units ? u0 + BK_FETCH_SPAN : n_units;
        for (uint64_t u = u0; u < u1; ++u) {
            const uint32_t b = (uint32_t)(u / nc_live), c = (uint32_t)(u % nc_live);
            const uint32_t n = __ldg(bp.cnt_bc + (uint64_t)b * bp.nc + c);
            const uint64_t blk = ((uint64_t)b * bp.nc + c) * cap;
            const uint32_t* src = bp.pos + blk + lane;
            uint8_t* dst = bp.res + blk + lane;
            const uint8_t* base = p.bits + (uint64_t)b * slice_bytes;
            if (bp.prefetch && b + 1 < bp.n_buckets) {
                for (uint32_t l = lane; l < pf_per_unit; l += 32) {
                    const uint64_t in_slice = (uint64_t)c * pf_per_unit + l;
                    const uint64_t line = (uint64_t)(b + 1) * slice_lines + in_slice;
                    if (in_slice < slice_lines && line < array_lines) prefetch_l2_keep(p.bits + line * 128, keep);
                }
            }
            for (uint32_t i0 = lane; i0 < n; i0 += 128) {
                const uint32_t left = n - i0;
                uint32_t r[4], v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) if ((uint32_t)(q * 32) < left) r[q] = ld_stream32(src + (i0 - lane) + q * 32);
#pragma unroll
                for (int q = 0; q < 4; ++q) if ((uint32_t)(q * 32) < left) v[q] = ldg_byte_keep(base + (r[q] >> 3), keep);
#pragma unroll
                for (int q = 0; q < 4; ++q) if ((uint32_t)(q * 32) < left) dst[(i0 - lane) + q * 32] = (uint8_t)((v[q] >> (r[q] & 7)) & 1u);
            }
        }
    }
}

template <int K>
__global__ void __launch_bounds__(BK_NT, 4) k_bbucket_reduce(const BloomBucketParams bp) {
    __shared__ uint32_t s_fail[BK_CH / 32];   // windows with a probe that found a zero bit
    __shared__ uint32_t s_ovf[BK_CH / 32];
    __shared__ uint32_t s_wseq[BK_CH];
    __shared__ uint32_t s_scan[BK_NT / 32];
    __shared__ unsigned long long s_chunk;
    const BloomParams& p = bp.bl;
    const SeqBatch& sb = p.sb;
    const uint32_t k = K ? K : sb.k;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t total = __ldg(sb.win_prefix + sb.n_seq);
    const uint32_t nc_live = bbucket_live_chunks(bp, total);
    const uint32_t cap = bp.cap, nb = bp.n_buckets;

    unsigned long long ticket = tid == 0 ? atomicAdd(bp.counter + 2, 1ULL) : 0ULL;
    for (;;) {
        if (tid == 0) s_chunk = ticket;
        if (tid < BK_CH / 32) s_fail[tid] = 0;
        for (uint32_t i = tid; i < BK_CH; i += BK_NT) s_wseq[i] = 0;
        __syncthreads();
        const uint64_t c = s_chunk;
        if (c >= nc_live) break;
        if (tid == 0) ticket = atomicAdd(bp.counter + 2, 1ULL);
        const uint64_t g0 = (bp.chunk0 + c) * BK_CH;
        const uint64_t g1 = g0 + BK_CH < total ? g0 + BK_CH : total;
        const uint32_t nwin = (uint32_t)(g1 - g0);
        if (tid < BK_CH / 32) s_ovf[tid] = __ldg(bp.ovf + c * (BK_CH / 32) + tid);
        const uint64_t s_lo = chunk_seq_table<BK_NT>(sb, __ldg(bp.chunk_seq + bp.chunk0 + c), g0, g1, s_wseq, s_scan);

        // ---- windows with a failed probe.  Warp w takes buckets w, w + 8, ...; the (block, group of 16 records) pairs of
        // all its blocks form one flat item space over the lanes, two items in flight per lane: one dependent load
        // chain per block made this loop latency-bound.
        auto scan16 = [&](const uint4& r, const uint4& wa, const uint4& wb, uint32_t valid) {
            const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
            const uint32_t ww[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t z = ~rw[q] & 0x01010101u;                      // result bytes are 0 or 1
                while (z) {
                    const uint32_t t = (uint32_t)q * 4 + ((__ffs(z) - 1) >> 3);
                    z &= z - 1;
                    if (t < valid) {
                        const uint32_t w = ((ww[t >> 1] >> ((t & 1) * 16)) & 0xFFFFu) & (BK_CH - 1);
                        atomicOr(&s_fail[w >> 5], 1u << (w & 31));
                    }
                }
            }
        };
        {
            const uint32_t G = cap / 16;                                       // groups per block (cap is a multiple of 16)
            const uint32_t n_items = nb > warp ? ((nb - warp + BK_NT / 32 - 1) / (BK_NT / 32)) * G : 0;
            for (uint32_t it0 = lane; it0 < n_items; it0 += 64) {
                uint4 r[2], wa[2], wb[2];
                uint32_t valid[2] = {0, 0};
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const uint32_t it = it0 + u * 32;
                    if (it < n_items) {
                        const uint32_t t = it / G, i16 = it - t * G;
                        const uint32_t b = t * (BK_NT / 32) + warp;
                        const uint32_t n = __ldg(bp.cnt_bc + (uint64_t)b * bp.nc + c);
                        if (i16 * 16 < n) {
                            const uint64_t blk = ((uint64_t)b * bp.nc + c) * cap;
                            const uint4* r4 = reinterpret_cast<const uint4*>(bp.res + blk);
                            const uint4* w4 = reinterpret_cast<const uint4*>(bp.wid + blk);
                            r[u] = ld_stream128(r4 + i16); wa[u] = ld_stream128(w4 + 2 * i16); wb[u] = ld_stream128(w4 + 2 * i16 + 1);
                            valid[u] = n - i16 * 16;
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 2; ++u)
                    if (valid[u]) scan16(r[u], wa[u], wb[u], valid[u]);
            }
        }
        __syncthreads();

        // ---- hits per sequence: warp w owns windows [w * BK_WARP_WIN, ...) of the chunk
        const uint32_t w_lo = warp * BK_WARP_WIN;
        const uint32_t w_hi = w_lo + BK_WARP_WIN < nwin ? w_lo + BK_WARP_WIN : nwin;
        uint32_t hits = 0;
        for (uint32_t r0 = w_lo; r0 < w_hi; r0 += 32) {
            const uint32_t lid = r0 + lane;
            const bool has = lid < w_hi;
            const uint32_t sq = has ? s_wseq[lid] : 0u;
            bool hit = false;
            if (has) {
                const uint32_t bit = 1u << (lid & 31);
                if (s_ovf[lid >> 5] & bit) {
                    const uint64_t seq = s_lo + sq - 1;
                    const uint64_t pos = __ldg(sb.seq_begin + seq) - sb.base_shift + (g0 + lid - __ldg(sb.win_prefix + seq)) * sb.step;
                    Term t;
                    bloom_term<K>(sb, pos, t);
                    hit = bloom_member(p, xxh3_64(t, k));
                } else {
                    hit = !(s_fail[lid >> 5] & bit);
                }
            }
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, hit);
            uint32_t rem = __ballot_sync(0xFFFFFFFFu, has);
            while (rem) {
                const uint32_t cur = __shfl_sync(0xFFFFFFFFu, sq, __ffs(rem) - 1);
                const uint32_t segmask = __ballot_sync(0xFFFFFFFFu, has && sq == cur);
                hits += __popc(bal & segmask);
                rem &= ~segmask;
                const uint32_t nxt = r0 + (32 - __clz(segmask));
                if (!(nxt < w_hi && s_wseq[nxt] == cur)) {
                    const uint64_t seq = s_lo + cur - 1;
                    if (lane == 0 && hits) {
                        const uint64_t ps = __ldg(sb.win_prefix + seq), pe = __ldg(sb.win_prefix + seq + 1);
                        uint32_t* o = p.out + p.seq0 + seq;
                        if (ps >= g0 + w_lo && pe <= g0 + w_hi) *o = hits; else atomicAdd(o, hits);
                    }
                    hits = 0;
                }
            }
        }
        __syncthreads();
    }
}

// ----------------------------------------------------------------------------------------
// score epilogue: per record the best document (first index of the maximum), the maximum and how many
// documents share it (ties = "ambiguous", scripts/benchmark/main.nf:417-436), plus per-document totals
// (ModelResult.get_total_hits, result.py:76-90).  Streams the count matrix once; saves the
// [n_seq x n_docs] device->host transfer when only read-level calls and file-level totals are wanted.
// ----------------------------------------------------------------------------------------
constexpr int REDUCE_NT = 256;
constexpr int REDUCE_ROWS = 256;   // records per CTA pass

template <typename OutT>
__global__ void __launch_bounds__(REDUCE_NT) k_scores_reduce(const OutT* __restrict__ counts, uint64_t n_seq, uint32_t n_docs,
                                                             uint32_t* __restrict__ best, uint32_t* __restrict__ best_count,
                                                             uint32_t* __restrict__ n_best, unsigned long long* __restrict__ totals) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr uint32_t NWARP = REDUCE_NT / 32;
    for (uint64_t r0 = (uint64_t)blockIdx.x * REDUCE_ROWS; r0 < n_seq; r0 += (uint64_t)gridDim.x * REDUCE_ROWS) {
        const uint32_t rows = (uint32_t)(n_seq - r0 < REDUCE_ROWS ? n_seq - r0 : REDUCE_ROWS);
        if (best || best_count || n_best) {
            for (uint32_t r = warp; r < rows; r += NWARP) {
                const OutT* row = counts + (r0 + r) * n_docs;
                uint32_t mx = 0, idx = 0xFFFFFFFFu, cnt = 0;
                for (uint32_t d = lane; d < n_docs; d += 32) {
                    uint32_t v = row[d];
                    if (idx == 0xFFFFFFFFu || v > mx) { mx = v; idx = d; cnt = 1; }
                    else if (v == mx) ++cnt;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    uint32_t omx = __shfl_xor_sync(0xFFFFFFFFu, mx, o), oidx = __shfl_xor_sync(0xFFFFFFFFu, idx, o),
                             ocnt = __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
                    if (oidx != 0xFFFFFFFFu) {
                        if (idx == 0xFFFFFFFFu || omx > mx) { mx = omx; idx = oidx; cnt = ocnt; }
                        else if (omx == mx) { cnt += ocnt; idx = oidx < idx ? oidx : idx; }
                    }
                }
                if (lane == 0) {
                    if (best) best[r0 + r] = idx;
                    if (best_count) best_count[r0 + r] = mx;
                    if (n_best) n_best[r0 + r] = cnt;
                }
            }
        }
        if (totals) {
            for (uint32_t d = threadIdx.x; d < n_docs; d += REDUCE_NT) {
                unsigned long long sum = 0;
                for (uint32_t r = 0; r < rows; ++r) sum += counts[(r0 + r) * n_docs + d];
                if (sum) atomicAdd(totals + d, sum);
            }
        }
    }
}

// ----------------------------------------------------------------------------------------
// construction (training side; probabilistic_filter_model.py:169-194, probabilistic_single_filter_model.py:63-96,
// probabilistic_filter_mlst_model.py:100-142): the query kernels' hashing with an atomic OR instead of a gather
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_or_bit(uint8_t* data, uint64_t byte, uint32_t bit) {
    uintptr_t a = reinterpret_cast<uintptr_t>(data + byte);
    unsigned int* w = reinterpret_cast<unsigned int*>(a & ~(uintptr_t)3);
    atomicOr(w, 1u << ((uint32_t)(a & 3) * 8 + bit));
}

struct CobsBuildParams {
    SeqBatch sb;
    const uint32_t* seq_doc;   // [n_seq] document (after page ordering) of every sequence
    uint8_t* data;             // [sig_size x row_bytes] of this page, zeroed, 4-byte aligned
    uint64_t sig_size, magic;
    uint32_t row_bytes, num_hashes, canonicalize;
    int32_t policy;
    uint32_t doc_lo, doc_hi;   // documents of this page
};

__global__ void __launch_bounds__(256) k_cobs_build(const CobsBuildParams p) {
    const SeqBatch& sb = p.sb;
    const uint64_t total = __ldg(sb.win_prefix + sb.n_seq);
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t s = seq_of_window(sb.win_prefix, sb.n_seq, g);
        uint32_t doc = __ldg(p.seq_doc + s);
        if (doc < p.doc_lo || doc >= p.doc_hi) continue;
        uint64_t pos = __ldg(sb.seq_begin + s) - sb.base_shift + (g - __ldg(sb.win_prefix + s)) * sb.step;
        Term t;
        if (!cobs_term<0>(sb, pos, p.canonicalize != 0, p.policy, t)) continue;
        Xxh64Pre pre;
        xxh64_prepare(t, sb.k, pre);
        const uint32_t within = doc - p.doc_lo;
        for (uint32_t j = 0; j < p.num_hashes; ++j) {
            uint64_t row = mod_barrett(xxh64_finish(pre, sb.k, (uint64_t)j), p.sig_size, p.magic);
            atomic_or_bit(p.data, row * p.row_bytes + within / 8, within % 8);
        }
    }
}

struct BloomBuildParams {
    SeqBatch sb;
    uint8_t* bits;
    uint64_t n_bits, magic;
    uint32_t k_hashes;
};

__global__ void __launch_bounds__(256) k_bloom_build(const BloomBuildParams p) {
    const SeqBatch& sb = p.sb;
    const uint64_t total = __ldg(sb.win_prefix + sb.n_seq);
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t s = seq_of_window(sb.win_prefix, sb.n_seq, g);
        uint64_t pos = __ldg(sb.seq_begin + s) - sb.base_shift + (g - __ldg(sb.win_prefix + s)) * sb.step;
        Term t;
        bloom_term<0>(sb, pos, t);
        uint64_t hi = 0, lo = xxh3_64(t, sb.k);
        for (uint32_t j = 0; j < p.k_hashes; ++j) {
            uint64_t ix = mod_barrett(lcg_next(hi, lo), p.n_bits, p.magic);
            atomic_or_bit(p.bits, ix >> 3, (uint32_t)(ix & 7));
        }
    }
}

// ----------------------------------------------------------------------------------------
// stage kernels (parity tests pin each stage on its own)
// ----------------------------------------------------------------------------------------
__global__ void k_stage_canonical(const uint64_t* __restrict__ packed, const uint32_t* __restrict__ invalid,
                                  uint64_t n_win, uint32_t k, uint64_t* __restrict__ codes, uint8_t* __restrict__ valid) {
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_win; g += (uint64_t)gridDim.x * blockDim.x) {
        bool inv = window_invalid(invalid, g, k);
        uint64_t fr = window_lsb(packed, g, k);
        uint64_t msb;
        canonical_lsb(fr, k, &msb);
        codes[g] = inv ? 0 : msb;
        valid[g] = inv ? 0 : 1;
    }
}

__global__ void k_stage_cobs_rows(const SeqBatch sb, const PageDesc* __restrict__ pages, uint32_t n_pages, uint32_t h,
                                  uint32_t canonicalize, int policy, uint64_t n_win,
                                  uint64_t* __restrict__ rows, uint8_t* __restrict__ valid) {
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_win; w += (uint64_t)gridDim.x * blockDim.x) {
        Term t;
        bool ok = cobs_term<0>(sb, w * sb.step, canonicalize != 0, policy, t);
        valid[w] = ok ? 1 : 0;
        Xxh64Pre pre;
        if (ok) xxh64_prepare(t, sb.k, pre);
        for (uint32_t j = 0; j < h; ++j) {
            uint64_t hv = ok ? xxh64_finish(pre, sb.k, (uint64_t)j) : 0;
            for (uint32_t pgi = 0; pgi < n_pages; ++pgi)
                rows[(w * h + j) * n_pages + pgi] = ok ? mod_barrett(hv, pages[pgi].sig_size, pages[pgi].magic) : 0;
        }
    }
}

__global__ void k_stage_bloom_hashes(const SeqBatch sb, uint64_t n_win, uint64_t* __restrict__ hashes) {
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_win; w += (uint64_t)gridDim.x * blockDim.x) {
        Term t;
        bloom_term<0>(sb, w * sb.step, t);
        hashes[w] = xxh3_64(t, sb.k);
    }
}

// re-stride file rows [src_row_bytes] -> HBM rows [dst_stride], keeping bytes [col0, col0 + n_col)
__global__ void k_restride(const uint8_t* __restrict__ src, uint64_t n_rows, uint32_t src_row_bytes, uint32_t col0,
                           uint32_t n_col, uint8_t* __restrict__ dst, uint32_t dst_stride) {
    uint64_t total = n_rows * dst_stride;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t r = i / dst_stride;
        uint32_t c = (uint32_t)(i - r * dst_stride);
        dst[i] = c < n_col ? src[r * src_row_bytes + col0 + c] : (uint8_t)0;
    }
}

// ----------------------------------------------------------------------------------------
// MLST epilogue (probabilistic_filter_mlst_model.py:236-256): the count matrix [n_seg x n_docs] of one locus stays on
// the device; a chunk matters only when some allele scores above the threshold in it ("hot").  k_mlst_hot flags the
// hot segments (one warp per row); k_mlst_compact keeps, per record and in chunk order, the indices and rows of its
// first `cap` hot segments.  Records below the chunking length have one segment, which is always kept.
// ----------------------------------------------------------------------------------------
template <typename CntT>
__global__ void __launch_bounds__(256) k_mlst_hot(const CntT* __restrict__ counts, uint64_t n_seg, uint32_t n_docs, uint32_t thr,
                                                  const uint8_t* __restrict__ seg_always, uint8_t* __restrict__ hot) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t sg = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; sg < n_seg; sg += n_warps) {
        const CntT* row = counts + sg * n_docs;
        bool any = seg_always[sg] != 0;
        for (uint32_t d = lane; d < n_docs && !any; d += 32) any = (uint32_t)row[d] > thr;
        any = __any_sync(0xFFFFFFFFu, any);
        if (lane == 0) hot[sg] = any ? 1 : 0;
    }
}

template <typename CntT>
__global__ void __launch_bounds__(256) k_mlst_compact(const CntT* __restrict__ counts, const uint64_t* __restrict__ rec_seg0,
                                                      uint32_t n_docs, const uint8_t* __restrict__ hot, uint32_t cap,
                                                      uint32_t* __restrict__ n_hot, uint32_t* __restrict__ hot_seg,
                                                      uint32_t* __restrict__ hot_rows) {
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_base;
    extern __shared__ uint32_t s_list[];   // [cap] local segment indices
    const uint32_t rec = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t g0 = rec_seg0[rec], g1 = rec_seg0[rec + 1];
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (uint64_t t0 = g0; t0 < g1; t0 += 256) {
        const uint64_t sg = t0 + tid;
        const bool h = sg < g1 && hot[sg];
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, h);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        uint32_t pre = s_base;
        for (uint32_t w = 0; w < warp; ++w) pre += s_warp[w];
        const uint32_t pos = pre + __popc(bal & ((1u << lane) - 1u));
        if (h && pos < cap) s_list[pos] = (uint32_t)(sg - g0);
        __syncthreads();
        if (tid == 0) { uint32_t t = 0; for (uint32_t w = 0; w < 8; ++w) t += s_warp[w]; s_base += t; }
        __syncthreads();
    }
    const uint32_t total = s_base, kept = total < cap ? total : cap;
    if (tid == 0) n_hot[rec] = total;
    for (uint32_t i = tid; i < kept; i += 256) hot_seg[(uint64_t)rec * cap + i] = s_list[i];
    for (uint64_t i = tid; i < (uint64_t)kept * n_docs; i += 256) {
        const uint32_t r = (uint32_t)(i / n_docs), d = (uint32_t)(i - (uint64_t)r * n_docs);
        hot_rows[((uint64_t)rec * cap + r) * n_docs + d] = (uint32_t)counts[(g0 + s_list[r]) * n_docs + d];
    }
}

// ----------------------------------------------------------------------------------------
// Synthetic index rows (BASELINE config 5: a 120 GB index that no file holds).  Counter-based: 32 documents of row r
// are   w = mix64(seed ^ r * C1 ^ word * C2);  bits = hi32(w) & lo32(w)   (fill 0.25), so any row can be regenerated
// anywhere (the oracle does, for the parity sample) without materialising the index.
// ----------------------------------------------------------------------------------------
XS_HD uint64_t synth_mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL; x ^= x >> 27; x *= 0x94D049BB133111EBULL; x ^= x >> 31; return x;
}
XS_HD uint32_t synth_row_word(uint64_t seed, uint64_t row, uint32_t word) {
    const uint64_t w = synth_mix64(seed ^ (row * 0x9E3779B97F4A7C15ULL) ^ ((uint64_t)word * 0xD1B54A32D192ED03ULL));
    return (uint32_t)(w >> 32) & (uint32_t)w;
}
// fills rows [0, n_rows) of a column shard: bytes [col0, col0 + n_col) of every row at dst_stride, zero padded;
// documents >= n_docs_total are zero like in a real file's last byte
__global__ void __launch_bounds__(256) k_synth_rows(uint8_t* __restrict__ dst, uint64_t n_rows, uint32_t dst_stride, uint32_t col0,
                                                    uint32_t n_col, uint32_t n_docs_total, uint64_t seed) {
    const uint32_t words = dst_stride / 4;
    const uint64_t total = n_rows * words;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / words;
        const uint32_t w = (uint32_t)(i - r * words);
        uint32_t v = 0;
        if (w * 4 < n_col) {
            const uint32_t gw = (col0 + w * 4) / 4;             // col0 is a multiple of 4 (document shards are 128-aligned)
            v = synth_row_word(seed, r, gw);
            const uint32_t d0 = gw * 32;
            if (d0 + 32 > n_docs_total) v &= d0 >= n_docs_total ? 0u : ((1u << (n_docs_total - d0)) - 1u);
            const uint32_t keep = n_col - w * 4;                  // bytes of this word inside the shard
            if (keep < 4) v &= (1u << (8 * keep)) - 1u;
        }
        reinterpret_cast<uint32_t*>(dst + r * dst_stride)[w] = v;
    }
}

// ----------------------------------------------------------------------------------------
// Epilogue of the document-column sharded exchange (config 5): `all` = what ncclAllGather delivered,
// [world][n_seq][w] counts, rank g's block holding its widths[g] local documents per record (zero padded to w).
// Per record: first document with the maximum count over all shards, that count, how many documents share it, and
// (optionally) per-document totals.  Consumed in place: no concatenation pass.  One warp per record.
// ----------------------------------------------------------------------------------------
struct ShardLayout {
    uint32_t world;
    uint32_t w;                  // padded documents per rank block
    uint32_t width[16];          // documents of rank g
    uint32_t doc0[16];           // first global document of rank g
};

template <typename CntT>
__global__ void __launch_bounds__(256) k_sharded_reduce(const CntT* __restrict__ all, uint64_t n_seq, const ShardLayout lay,
                                                        uint32_t* __restrict__ best, uint32_t* __restrict__ best_count,
                                                        uint32_t* __restrict__ n_best, unsigned long long* __restrict__ totals) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    constexpr uint32_t VEC = 16 / sizeof(CntT);
    uint32_t n_docs_total = 0;
    for (uint32_t g = 0; g < lay.world; ++g) n_docs_total += lay.width[g];
    for (uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_seq; r += n_warps) {
        uint32_t mx = 0, idx = 0xFFFFFFFFu, cnt = 0;
        for (uint32_t g = 0; g < lay.world; ++g) {
            const CntT* row = all + ((uint64_t)g * n_seq + r) * lay.w;
            const uint32_t wd = lay.width[g];
            if (sizeof(CntT) == 1 && (lay.w % VEC) == 0) {
                // uint8 counts, four per word: all-zero words (most of a score row) cost three instructions; the padding
                // columns behind wd are zero and never tie with a positive maximum.  A record whose maximum is 0 is
                // resolved after the loop (every document ties).
                for (uint32_t d0 = lane * 16; d0 < wd; d0 += 32 * 16) {
                    const uint4 q = *reinterpret_cast<const uint4*>(row + d0);
                    const uint32_t wv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t w = wv[j];
                        if (w == 0) continue;
                        uint32_t m = __vmaxu4(w, __byte_perm(w, 0, 0x1032));
                        m = __vmaxu4(m, __byte_perm(m, 0, 0x2301)) & 0xFFu;
                        if (m >= mx) {
                            const uint32_t eq = __vcmpeq4(w, m * 0x01010101u);
                            const uint32_t c = (uint32_t)__popc(eq) >> 3;
                            if (m > mx || idx == 0xFFFFFFFFu) { mx = m; idx = lay.doc0[g] + d0 + j * 4 + ((__ffs(eq) - 1) >> 3); cnt = c; }
                            else cnt += c;
                        }
                        if (totals) {
#pragma unroll
                            for (int b = 0; b < 4; ++b) {
                                const uint32_t v = (w >> (8 * b)) & 0xFFu;
                                if (v) atomicAdd(totals + lay.doc0[g] + d0 + j * 4 + b, (unsigned long long)v);
                            }
                        }
                    }
                }
            } else if ((lay.w % VEC) == 0) {           // 16-byte aligned rows: vector loads
                for (uint32_t d0 = lane * VEC; d0 < wd; d0 += 32 * VEC) {
                    const uint4 q = *reinterpret_cast<const uint4*>(row + d0);
                    const CntT* e = reinterpret_cast<const CntT*>(&q);
#pragma unroll
                    for (uint32_t j = 0; j < VEC; ++j) {
                        if (d0 + j >= wd) break;
                        const uint32_t v = e[j], d = lay.doc0[g] + d0 + j;
                        if (idx == 0xFFFFFFFFu || v > mx) { mx = v; idx = d; cnt = 1; }
                        else if (v == mx) ++cnt;
                        if (totals && v) atomicAdd(totals + d, (unsigned long long)v);
                    }
                }
            } else {
                for (uint32_t d0 = lane; d0 < wd; d0 += 32) {
                    const uint32_t v = row[d0], d = lay.doc0[g] + d0;
                    if (idx == 0xFFFFFFFFu || v > mx) { mx = v; idx = d; cnt = 1; }
                    else if (v == mx) ++cnt;
                    if (totals && v) atomicAdd(totals + d, (unsigned long long)v);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const uint32_t omx = __shfl_xor_sync(0xFFFFFFFFu, mx, o), oidx = __shfl_xor_sync(0xFFFFFFFFu, idx, o),
                           ocnt = __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
            if (oidx != 0xFFFFFFFFu) {
                if (idx == 0xFFFFFFFFu || omx > mx) { mx = omx; idx = oidx; cnt = ocnt; }
                else if (omx == mx) { cnt += ocnt; idx = oidx < idx ? oidx : idx; }
            }
        }
        if (sizeof(CntT) == 1 && (lay.w % VEC) == 0 && (idx == 0xFFFFFFFFu || mx == 0)) { mx = 0; idx = 0; cnt = n_docs_total; }
        if (lane == 0) {
            if (best) best[r] = idx;
            if (best_count) best_count[r] = mx;
            if (n_best) n_best[r] = cnt;
        }
    }
}

// per-document fill of a page estimated from `n_sample` evenly spaced rows (xs_cobs_doc_fill): thread per
// (sampled row, 32-document word), one atomic per set bit
__global__ void __launch_bounds__(256) k_doc_fill(const PageDesc pg, uint64_t n_sample, uint32_t* __restrict__ counts) {
    const uint32_t words = (pg.n_docs + 31) / 32;
    const uint64_t total = n_sample * words;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / words;
        const uint32_t w = (uint32_t)(i - r * words);
        const uint64_t row = (unsigned __int128)r * pg.sig_size / n_sample;
        uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(pg.data + row * pg.row_stride) + w);
        while (v) {
            const uint32_t b = __ffs(v) - 1;
            v &= v - 1;
            const uint32_t d = w * 32 + b;
            if (d < pg.n_docs) atomicAdd(counts + pg.doc_off + d, 1u);
        }
    }
}

}  // namespace xs
