// xs_kernels.cuh — sm_100a kernels of the k-mer scoring path.
//
//   k_pack2bit        ASCII bases -> 2-bit stream + non-ACGT bitmap             (feeder, file_io.py:47-79)
//   k_cobs_narrow     rows <= 16 B (D <= 128 per page): one thread per sampled window; extract ->
//                     canonical -> XXH64 x h -> Barrett mod -> h x 128-bit gathers -> AND ->
//                     per-sequence document counts by warp ballot/popcount
//                     (cobs Search.search behind probabilistic_filter_model.py:227)
//   k_cobs_wide       rows > 16 B: lanes own 16-byte column chunks, warp-coalesced row gathers,
//                     bit-sliced vertical counters, shared-memory count staging
//   k_bloom           XXH3-64 -> 128-bit LCG -> bit probes with early exit
//                     (probabilistic_single_filter_model.py:122-124,161-180)
//   stage kernels     canonical codes / row ids / bloom hashes alone, for the parity tests
//
// Work decomposition.  The sampled windows of all sequences of a batch form one flat index
// space (win_prefix = exclusive scan of windows per sequence).  Persistent CTAs walk tiles of
// that space, so lanes stay dense whatever the read lengths are; the sequences a tile touches
// are resolved once per tile into shared memory.
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
// saturating add into an element other CTAs may add to as well
template <typename OutT>
__device__ __forceinline__ void out_add(OutT* p, uint32_t v) {
    if (v == 0) return;
    if (sizeof(OutT) == 4) {
        atomicAdd(reinterpret_cast<unsigned int*>(p), v);
    } else {
        uintptr_t a = reinterpret_cast<uintptr_t>(p);
        unsigned int* w = reinterpret_cast<unsigned int*>(a & ~(uintptr_t)3);
        uint32_t sh = (uint32_t)(a & 3) * 8;
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
// tile map: which sequences a tile of the flat window space touches
// ----------------------------------------------------------------------------------------
template <int TILE_W>
struct TileMap {
    int32_t rel[TILE_W + 2];     // window offset of local sequence i relative to the tile start, clamped to [0, tile_n]
    uint64_t begin[TILE_W + 1];  // seq_begin - base_shift of local sequence i
    uint64_t first_off;          // tile_start - prefix[seq_lo]
    uint64_t seq_lo;
    uint64_t seq_hi;
    uint32_t ns;                 // sequences in [seq_lo, seq_hi]
    uint32_t tile_n;             // windows in this tile
    int32_t fallback;            // ns > TILE_W: too many (empty) sequences to stage
    int32_t last_complete;       // last local sequence ends inside the tile
};

template <int TILE_W, int NT>
__device__ __forceinline__ void tile_setup(TileMap<TILE_W>& tm, const SeqBatch& sb, uint64_t tile_start, uint64_t total) {
    const int tid = threadIdx.x;
    uint32_t tile_n = (uint32_t)(total - tile_start < (uint64_t)TILE_W ? total - tile_start : (uint64_t)TILE_W);
    if (tid == 0) {
        uint64_t s = seq_of_window(sb.win_prefix, sb.n_seq, tile_start);
        tm.seq_lo = s;
        tm.first_off = tile_start - __ldg(sb.win_prefix + s);
        tm.tile_n = tile_n;
    }
    if (tid == 32) tm.seq_hi = seq_of_window(sb.win_prefix, sb.n_seq, tile_start + tile_n - 1);
    __syncthreads();
    uint64_t ns64 = tm.seq_hi - tm.seq_lo + 1;
    bool fb = ns64 > (uint64_t)TILE_W;
    if (tid == 0) { tm.ns = fb ? 0u : (uint32_t)ns64; tm.fallback = fb ? 1 : 0; }
    if (!fb) {
        uint32_t ns = (uint32_t)ns64;
        for (uint32_t i = tid; i <= ns; i += NT) {
            uint64_t pv = __ldg(sb.win_prefix + tm.seq_lo + i);
            int64_t r = (int64_t)(pv - tile_start);
            if (i == ns) tm.last_complete = (r <= (int64_t)tile_n) ? 1 : 0;
            r = r < 0 ? 0 : (r > (int64_t)tile_n ? (int64_t)tile_n : r);
            tm.rel[i] = (int32_t)r;
            if (i < ns) tm.begin[i] = __ldg(sb.seq_begin + tm.seq_lo + i) - sb.base_shift;
        }
    }
    __syncthreads();
}

// base position of local window lw; *seq_local receives the local sequence index (staged tiles)
template <int TILE_W>
__device__ __forceinline__ uint64_t tile_locate(const TileMap<TILE_W>& tm, const SeqBatch& sb, uint64_t tile_start,
                                                uint32_t lw, uint64_t* seq_global) {
    if (tm.fallback) {
        uint64_t g = tile_start + lw;
        uint64_t s = seq_of_window(sb.win_prefix, sb.n_seq, g);
        *seq_global = s;
        uint64_t wi = g - __ldg(sb.win_prefix + s);
        return __ldg(sb.seq_begin + s) - sb.base_shift + wi * sb.step;
    }
    uint32_t lo = 0, hi = tm.ns;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if ((uint32_t)tm.rel[mid] <= lw) lo = mid; else hi = mid;
    }
    *seq_global = tm.seq_lo + lo;
    uint64_t wi = lo == 0 ? (uint64_t)lw + tm.first_off : (uint64_t)(lw - (uint32_t)tm.rel[lo]);
    return tm.begin[lo] + wi * sb.step;
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
};

constexpr int NARROW_NT = 256;
constexpr int NARROW_R = 4;
constexpr int NARROW_TILE = NARROW_NT * NARROW_R;

// count the documents of masks[lo..hi) into per-lane counters: lane l holds documents l, l+32, l+64, l+96
__device__ __forceinline__ void count_masks(const uint4* __restrict__ masks, uint32_t lo, uint32_t hi, uint32_t lane,
                                            uint32_t cnt[4]) {
    for (uint32_t base = lo; base < hi; base += 32) {
        uint32_t i = base + lane;
        uint4 m = i < hi ? masks[i] : make_uint4(0, 0, 0, 0);
        uint32_t mw[4] = {m.x, m.y, m.z, m.w};
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
}

template <int K, int H, typename OutT>
__global__ void __launch_bounds__(NARROW_NT) k_cobs_narrow(const CobsParams p) {
    __shared__ TileMap<NARROW_TILE> tm;
    __shared__ uint4 s_mask[NARROW_TILE];

    const SeqBatch& sb = p.sb;
    const uint32_t k = K ? K : sb.k;
    const uint32_t h = H ? H : p.num_hashes;
    const PageDesc pg = p.pages[blockIdx.y];
    const int tid = threadIdx.x;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    constexpr int NWARP = NARROW_NT / 32;
    OutT* out = reinterpret_cast<OutT*>(p.out);

    const uint64_t total = __ldg(sb.win_prefix + sb.n_seq);
    const uint64_t n_tiles = (total + NARROW_TILE - 1) / NARROW_TILE;

    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t tile_start = tile * NARROW_TILE;
        tile_setup<NARROW_TILE, NARROW_NT>(tm, sb, tile_start, total);
        const uint32_t tile_n = tm.tile_n;

        // ---- phase A: one window per thread per round -> 128-bit document mask
#pragma unroll 1
        for (int r = 0; r < NARROW_R; ++r) {
            uint32_t lw = r * NARROW_NT + tid;
            uint4 m = make_uint4(0, 0, 0, 0);
            uint64_t seq_g = 0;
            if (lw < tile_n) {
                uint64_t pos = tile_locate(tm, sb, tile_start, lw, &seq_g);
                Term t;
                if (cobs_term<K>(sb, pos, p.canonicalize != 0, p.policy, t)) {
                    Xxh64Pre pre;
                    xxh64_prepare(t, k, pre);
                    m = make_uint4(~0u, ~0u, ~0u, ~0u);
                    if (H) {
                        const uint8_t* addr[H ? H : 1];
#pragma unroll
                        for (int j = 0; j < (H ? H : 1); ++j) {
                            uint64_t hv = xxh64_finish(pre, k, (uint64_t)j);
                            addr[j] = pg.data + mod_barrett(hv, pg.sig_size, pg.magic) * 16;
                        }
#pragma unroll
                        for (int j = 0; j < (H ? H : 1); ++j) {
                            uint4 v = ldg128(addr[j]);
                            m.x &= v.x; m.y &= v.y; m.z &= v.z; m.w &= v.w;
                        }
                    } else {
                        for (uint32_t j = 0; j < h; ++j) {
                            uint64_t hv = xxh64_finish(pre, k, (uint64_t)j);
                            uint4 v = ldg128(pg.data + mod_barrett(hv, pg.sig_size, pg.magic) * 16);
                            m.x &= v.x; m.y &= v.y; m.z &= v.z; m.w &= v.w;
                        }
                    }
                }
                if (tm.fallback) {
                    // too many (empty) sequences in this tile to stage: add set bits directly
                    uint32_t mw[4] = {m.x, m.y, m.z, m.w};
                    for (int w = 0; w < 4; ++w) {
                        uint32_t u = mw[w];
                        while (u) {
                            uint32_t b = __ffs(u) - 1; u &= u - 1;
                            uint32_t d = w * 32 + b;
                            if (d < pg.n_docs) out_add<OutT>(out + (p.seq0 + seq_g) * p.ld + pg.doc_off + d, 1u);
                        }
                    }
                }
            }
            s_mask[lw] = m;
        }
        __syncthreads();

        // ---- phase B: per-sequence document counts
        if (!tm.fallback) {
            const uint32_t ns = tm.ns;
            if (ns >= 4) {
                for (uint32_t i = warp; i < ns; i += NWARP) {
                    uint32_t lo = (uint32_t)tm.rel[i], hi = (uint32_t)tm.rel[i + 1];
                    if (i == 0) lo = 0;
                    if (hi <= lo) continue;
                    uint32_t cnt[4] = {0, 0, 0, 0};
                    count_masks(s_mask, lo, hi, lane, cnt);
                    bool complete = (i > 0 || tm.first_off == 0) && (i + 1 < ns || tm.last_complete);
                    OutT* row = out + (p.seq0 + tm.seq_lo + i) * p.ld + pg.doc_off;
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        uint32_t d = w * 32 + lane;
                        if (d < pg.n_docs) {
                            if (complete) out_store<OutT>(row + d, cnt[w]); else out_add<OutT>(row + d, cnt[w]);
                        }
                    }
                }
            } else {
                // few long sequences: every warp takes a slice of each
                for (uint32_t i = 0; i < ns; ++i) {
                    uint32_t lo = (uint32_t)tm.rel[i], hi = (uint32_t)tm.rel[i + 1];
                    if (i == 0) lo = 0;
                    if (hi <= lo) continue;
                    uint32_t span = (((hi - lo) + NWARP - 1) / NWARP + 31) & ~31u;
                    uint32_t a = lo + warp * span, b = a + span < hi ? a + span : hi;
                    if (a >= hi) continue;
                    uint32_t cnt[4] = {0, 0, 0, 0};
                    count_masks(s_mask, a, b, lane, cnt);
                    OutT* row = out + (p.seq0 + tm.seq_lo + i) * p.ld + pg.doc_off;
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        uint32_t d = w * 32 + lane;
                        if (d < pg.n_docs) out_add<OutT>(row + d, cnt[w]);
                    }
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
constexpr int WIDE_MAX_COLS = 96;       // 16-byte column chunks per column block (12288 documents)

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
    __shared__ uint64_t s_seq, s_chunk;

    const int tid = threadIdx.x;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    constexpr uint32_t NWARP = WIDE_NT / 32;
    OutT* out = reinterpret_cast<OutT*>(p.out);

    const uint32_t C = cb.n_cols;
    uint32_t lpw = 1; while (lpw < C && lpw < 32) lpw <<= 1;   // lanes per window
    const uint32_t slots = 32 / lpw;                           // windows a warp handles at once
    const uint32_t n_round = (C + lpw - 1) / lpw;
    const uint32_t wsplit = NWARP / n_round > 0 ? NWARP / n_round : 1;
    const uint32_t n_unit = n_round * wsplit;

    const uint64_t total_items = __ldg(wp.chunk_prefix + sb.n_seq);
    for (uint64_t item = blockIdx.x; item < total_items; item += gridDim.x) {
        if (tid == 0) {
            uint64_t s = seq_of_window(wp.chunk_prefix, sb.n_seq, item);
            s_seq = s;
            s_chunk = item - __ldg(wp.chunk_prefix + s);
        }
        __syncthreads();
        const uint64_t seq = s_seq;
        const uint64_t sbeg = __ldg(sb.seq_begin + seq), send = __ldg(sb.seq_end + seq);
        const uint64_t nw_seq = windows_of(sbeg, send, sb.base_shift, sb.n_bases, k, sb.step);
        const uint64_t w0 = s_chunk * WIDE_CHUNK;
        const uint32_t nwin = (uint32_t)(nw_seq - w0 < (uint64_t)WIDE_CHUNK ? nw_seq - w0 : (uint64_t)WIDE_CHUNK);
        const bool complete = nw_seq <= (uint64_t)WIDE_CHUNK;

        // ---- phase 1: row byte offsets of every window of the item
        for (uint32_t lw = tid; lw < nwin; lw += WIDE_NT) {
            uint64_t pos = sbeg - sb.base_shift + (w0 + lw) * sb.step;
            Term t;
            if (cobs_term<K>(sb, pos, p.canonicalize != 0, p.policy, t)) {
                Xxh64Pre pre;
                xxh64_prepare(t, k, pre);
                for (uint32_t j = 0; j < h; ++j) {
                    uint64_t hv = xxh64_finish(pre, k, (uint64_t)j);
                    s_rows[lw * h + j] = mod_barrett(hv, pg.sig_size, pg.magic) * pg.row_stride;
                }
            } else {
                s_rows[lw * h] = ~0ULL;
            }
        }
        for (uint32_t d = tid; d < C * 128; d += WIDE_NT) s_cnt[d] = 0;
        __syncthreads();

        // ---- phase 2: gather + AND + vertical counters
        for (uint32_t u = warp; u < n_unit; u += NWARP) {
            const uint32_t r = u % n_round, ws = u / n_round;
            const uint32_t slot = lane / lpw, cl = lane % lpw;
            const uint32_t col = r * lpw + cl;
            const bool active = col < C;
            const uint8_t* colbase = pg.data + (size_t)(cb.c0 + col) * 16;
            uint32_t pl[8][4];
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) pl[a][b] = 0;
            if (active) {
                for (uint32_t w = ws * slots + slot; w < nwin; w += wsplit * slots) {
                    uint64_t r0 = s_rows[w * h];
                    if (r0 == ~0ULL) continue;
                    uint4 m = ldg128(colbase + r0);
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
                    uint32_t carry[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
                    for (int a = 0; a < 8; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            uint32_t tcar = pl[a][b] & carry[b];
                            pl[a][b] ^= carry[b];
                            carry[b] = tcar;
                        }
                }
                // expand the bit planes of this lane's 128 documents
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    uint32_t any = 0;
#pragma unroll
                    for (int a = 0; a < 8; ++a) any |= pl[a][b];
                    while (any) {
                        uint32_t bit = __ffs(any) - 1; any &= any - 1;
                        uint32_t c = 0;
#pragma unroll
                        for (int a = 0; a < 8; ++a) c |= ((pl[a][b] >> bit) & 1u) << a;
                        atomicAdd(&s_cnt[col * 128 + b * 32 + bit], c);
                    }
                }
            }
        }
        __syncthreads();

        // ---- phase 3: coalesced write of the block's document counts
        OutT* row = out + (p.seq0 + seq) * p.ld + pg.doc_off + (size_t)cb.c0 * 128;
        for (uint32_t d = tid; d < cb.n_docs; d += WIDE_NT) {
            uint32_t v = s_cnt[d];
            if (complete) out_store<OutT>(row + d, v); else out_add<OutT>(row + d, v);
        }
        __syncthreads();
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
    uint32_t pad;
    uint32_t* out;       // [n_seq]
    uint64_t seq0;
};

constexpr int BLOOM_NT = 256;
constexpr int BLOOM_R = 4;
constexpr int BLOOM_TILE = BLOOM_NT * BLOOM_R;

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
__global__ void __launch_bounds__(BLOOM_NT) k_bloom(const BloomParams p) {
    __shared__ TileMap<BLOOM_TILE> tm;
    __shared__ uint32_t s_hit[BLOOM_TILE / 32];   // one bit per window

    const SeqBatch& sb = p.sb;
    const uint32_t k = K ? K : sb.k;
    const int tid = threadIdx.x;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    constexpr int NWARP = BLOOM_NT / 32;

    const uint64_t total = __ldg(sb.win_prefix + sb.n_seq);
    const uint64_t n_tiles = (total + BLOOM_TILE - 1) / BLOOM_TILE;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t tile_start = tile * BLOOM_TILE;
        tile_setup<BLOOM_TILE, BLOOM_NT>(tm, sb, tile_start, total);
        const uint32_t tile_n = tm.tile_n;
#pragma unroll 1
        for (int r = 0; r < BLOOM_R; ++r) {
            uint32_t lw = r * BLOOM_NT + tid;
            bool hit = false;
            uint64_t seq_g = 0;
            if (lw < tile_n) {
                uint64_t pos = tile_locate(tm, sb, tile_start, lw, &seq_g);
                Term t;
                bloom_term<K>(sb, pos, t);
                hit = bloom_member(p, xxh3_64(t, k));
                if (tm.fallback && hit) atomicAdd(p.out + p.seq0 + seq_g, 1u);
            }
            uint32_t bal = __ballot_sync(0xFFFFFFFFu, hit);
            if (lane == 0) s_hit[lw >> 5] = bal;
        }
        __syncthreads();
        if (!tm.fallback) {
            const uint32_t ns = tm.ns;
            for (uint32_t i = warp; i < ns; i += NWARP) {
                uint32_t lo = (uint32_t)tm.rel[i], hi = (uint32_t)tm.rel[i + 1];
                if (i == 0) lo = 0;
                if (hi <= lo) continue;
                uint32_t c = 0;
                for (uint32_t wd = (lo >> 5) + lane; wd <= ((hi - 1) >> 5); wd += 32) {
                    uint32_t v = s_hit[wd];
                    uint32_t b0 = wd << 5;
                    if (b0 < lo) v &= ~0u << (lo - b0);
                    if (b0 + 32 > hi) v &= ~0u >> (b0 + 32 - hi);
                    c += __popc(v);
                }
                c = __reduce_add_sync(0xFFFFFFFFu, c);
                if (lane == 0 && c) {
                    bool complete = (i > 0 || tm.first_off == 0) && (i + 1 < ns || tm.last_complete);
                    uint32_t* o = p.out + p.seq0 + tm.seq_lo + i;
                    if (complete) *o = c; else atomicAdd(o, c);
                }
            }
        }
        __syncthreads();
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

}  // namespace xs
