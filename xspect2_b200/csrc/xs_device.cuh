// xs_device.cuh — device-side building blocks of the k-mer scoring path (sm_100a).
//
//   2-bit window extraction + canonical form  (replaces cobs canonicalize_kmer and
//       probabilistic_single_filter_model.py:175-180 for pure-ACGT windows)
//   literal-byte canonical form               (same, for windows touching other bytes)
//   XXH64 with seed j                         (cobs process_hashes / create_hashes)
//   XXH3-64 seed 0                            (xxhash.xxh3_64_intdigest; ...single_filter_model.py:88,155-158)
//   128-bit LCG probe indexes                 (rbloom lcg)
//   64-bit Barrett modulo                     (hash % signature_size, idx % nbits)
//
// A k-mer term (k <= 32) is carried as four little-endian 64-bit words of its ASCII bytes,
// zero padded, which is exactly the view XXH64 / XXH3 take of their input.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define XS_HD __host__ __device__ __forceinline__

namespace xs {

// intrinsic shims so the same functions can be unit-tested on the host (tests/native/)
XS_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t s) {
#ifdef __CUDA_ARCH__
    return __byte_perm(a, b, s);
#else
    uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        uint32_t sel = (s >> (4 * i)) & 0xF;
        uint32_t byte = (uint32_t)(v >> (8 * (sel & 7))) & 0xFF;
        if (sel & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
        r |= byte << (8 * i);
    }
    return r;
#endif
}
XS_HD uint64_t brev64(uint64_t x) {
#ifdef __CUDA_ARCH__
    return __brevll(x);
#else
    uint64_t r = 0;
    for (int i = 0; i < 64; ++i) r |= ((x >> i) & 1ULL) << (63 - i);
    return r;
#endif
}
XS_HD uint64_t umul64hi(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}
template <typename T>
XS_HD T ldg(const T* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

struct Term {
    uint64_t w[4];
};

// ------------------------------------------------------------------ small helpers
XS_HD uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

XS_HD uint64_t mask2k(uint32_t k) {  // low 2k bits
    return k >= 32 ? ~0ULL : ((1ULL << (2 * k)) - 1ULL);
}

// reverse the order of the 32 two-bit groups of x
XS_HD uint64_t group_reverse(uint64_t x) {
    uint64_t y = brev64(x);
    return ((y >> 1) & 0x5555555555555555ULL) | ((y & 0x5555555555555555ULL) << 1);
}

// x mod m with magic = floor(2^64 / m) (m >= 2; magic = ~0 for m == 1); m <= 2^63
XS_HD uint64_t mod_barrett(uint64_t x, uint64_t m, uint64_t magic) {
    uint64_t q = umul64hi(x, magic);
    uint64_t r = x - q * m;
    return r >= m ? r - m : r;
}

// the same for m < 2^31: the remainder before the correction is < 2m < 2^32, so 32-bit arithmetic is exact
XS_HD uint32_t mod_barrett_small(uint64_t x, uint32_t m, uint64_t magic) {
    uint32_t q = (uint32_t)umul64hi(x, magic);
    uint32_t r = (uint32_t)x - q * m;
    return r >= m ? r - m : r;
}

// ------------------------------------------------------------------ 2-bit windows
// packed stream: word i covers bases [32i, 32i+32), base j at bits [2j, 2j+1] (A0 C1 G2 T3);
// invalid stream: uint32 word i bit j = 1 for a non-ACGT byte.  Both have one spare word.
XS_HD uint64_t window_lsb(const uint64_t* __restrict__ packed, uint64_t g, uint32_t k) {
    uint64_t wi = g >> 5;
    uint32_t s = (uint32_t)(g & 31) * 2;
    uint64_t lo = ldg(packed + wi);
    uint64_t hi = ldg(packed + wi + 1);
    uint64_t x = s ? ((lo >> s) | (hi << (64 - s))) : lo;
    return x & mask2k(k);
}

XS_HD bool window_invalid(const uint32_t* __restrict__ invalid, uint64_t g, uint32_t k) {
    uint64_t wi = g >> 5;
    uint32_t s = (uint32_t)(g & 31);
    uint64_t v = (uint64_t)ldg(invalid + wi) | ((uint64_t)ldg(invalid + wi + 1) << 32);
    v >>= s;
    uint64_t m = k >= 32 ? 0xFFFFFFFFULL : ((1ULL << k) - 1ULL);
    return (v & m) != 0;
}

// canonical form of a pure-ACGT window.  fr = window, first base in the LEAST significant
// group.  Returns the canonical k-mer in the same LSB-first form (ready for expansion) and,
// through *msb, in first-base-most-significant form (integer order == lexicographic order).
XS_HD uint64_t canonical_lsb(uint64_t fr, uint32_t k, uint64_t* msb) {
    uint64_t m = mask2k(k);
    uint64_t f = group_reverse(fr) >> (64 - 2 * k);  // forward, MSB-first
    uint64_t rc = (~fr) & m;                         // reverse complement, MSB-first
    bool use_rc = rc < f;
    *msb = use_rc ? rc : f;
    return use_rc ? ((~f) & m) : fr;                 // (~f)&m == reverse complement, LSB-first
}

// expand up to 32 two-bit codes (LSB-first) into ASCII bytes, zero beyond k
XS_HD void expand_ascii(uint64_t c, uint32_t k, Term& t) {
    const uint32_t LUT = 0x54474341u;  // 'A','C','G','T'
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t x = (uint32_t)(c >> (16 * i)) & 0xFFFFu;
        uint32_t e = byte_perm(LUT, 0, x & 0x3333u);         // bases 0,2,4,6
        uint32_t o = byte_perm(LUT, 0, (x >> 2) & 0x3333u);  // bases 1,3,5,7
        uint32_t lo = byte_perm(e, o, 0x5140);
        uint32_t hi = byte_perm(e, o, 0x7362);
        t.w[i] = ((uint64_t)hi << 32) | lo;
    }
    // zero the bytes at and beyond k
    uint32_t full = k >> 3, rem = k & 7;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if ((uint32_t)i > full || ((uint32_t)i == full && rem == 0)) t.w[i] = 0;
        else if ((uint32_t)i == full) t.w[i] &= (1ULL << (8 * rem)) - 1ULL;
    }
}

// ------------------------------------------------------------------ literal-byte path
// complement rules for windows that touch a byte outside upper-case ACGT
struct CobsComp {   // cobs: ACGT only, anything else maps to 0x00 (LITERAL policy)
    XS_HD uint8_t operator()(uint8_t c) const {
        switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return 0; }
    }
};
struct BioComp {    // Biopython DNA complement: IUPAC-aware, case-preserving, U->A, others identity
    XS_HD uint8_t operator()(uint8_t c) const {
        uint8_t u = c & 0xDF;
        if (u < 'A' || u > 'Z') return c;
        uint8_t m;
        switch (u) {
            case 'A': m = 'T'; break; case 'C': m = 'G'; break; case 'G': m = 'C'; break; case 'T': m = 'A'; break;
            case 'M': m = 'K'; break; case 'R': m = 'Y'; break; case 'Y': m = 'R'; break; case 'K': m = 'M'; break;
            case 'V': m = 'B'; break; case 'H': m = 'D'; break; case 'D': m = 'H'; break; case 'B': m = 'V'; break;
            case 'U': m = 'A'; break;
            default: m = u; break;  // W S X N and every other letter map to themselves
        }
        return (uint8_t)(m | (c & 0x20));
    }
};
struct TableComp {  // 256-byte table (host unit tests)
    const uint8_t* t;
    XS_HD uint8_t operator()(uint8_t c) const { return t[c]; }
};

// Builds min(window, revcomp(window)) over literal bytes (the forward k-mer wins ties, the bytes
// are equal then anyway).
template <class Comp>
XS_HD void literal_term(const uint8_t* __restrict__ bases, uint64_t g, uint32_t k, Comp comp,
                        bool canonicalize, Term& t) {
    uint64_t f[4] = {0, 0, 0, 0}, r[4] = {0, 0, 0, 0};
    int cmp = 0;  // sign of (fwd - rc) at the first differing byte
    for (uint32_t i = 0; i < k; ++i) {
        uint8_t a = bases[g + i];
        uint8_t b = comp(bases[g + k - 1 - i]);
        f[i >> 3] |= (uint64_t)a << (8 * (i & 7));
        r[i >> 3] |= (uint64_t)b << (8 * (i & 7));
        if (cmp == 0 && a != b) cmp = a < b ? -1 : 1;
    }
    bool use_rc = canonicalize && cmp > 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) t.w[i] = use_rc ? r[i] : f[i];
}

// ------------------------------------------------------------------ XXH64
#define XS_P64_1 0x9E3779B185EBCA87ULL
#define XS_P64_2 0xC2B2AE3D27D4EB4FULL
#define XS_P64_3 0x165667B19E3779F9ULL
#define XS_P64_4 0x85EBCA77C2B2AE63ULL
#define XS_P64_5 0x27D4EB2F165667C5ULL

XS_HD uint64_t xxh64_round(uint64_t acc, uint64_t in) {
    acc += in * XS_P64_2;
    acc = rotl64(acc, 31);
    return acc * XS_P64_1;
}
XS_HD uint64_t xxh64_avalanche(uint64_t h) {
    h ^= h >> 33; h *= XS_P64_2;
    h ^= h >> 29; h *= XS_P64_3;
    h ^= h >> 32;
    return h;
}

// Seed-independent part of XXH64 over a <= 32-byte term: the per-lane products that do not
// involve the running hash are computed once and shared by all h seeds.
struct Xxh64Pre {
    uint64_t k8[4];  // len < 32: round(0, w_i) for each whole 8-byte lane; len == 32: w_i * P2
    uint64_t k4;     // (4-byte lane) * P1
    uint64_t kb[3];  // (tail byte) * P5
};

XS_HD void xxh64_prepare(const Term& t, uint32_t len, Xxh64Pre& p) {
    if (len == 32) {
#pragma unroll
        for (int i = 0; i < 4; ++i) p.k8[i] = t.w[i] * XS_P64_2;
        p.k4 = 0; p.kb[0] = p.kb[1] = p.kb[2] = 0;
        return;
    }
    uint32_t n8 = len >> 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) p.k8[i] = (uint32_t)i < n8 ? xxh64_round(0, t.w[i]) : 0;
    uint64_t tail = n8 < 4 ? t.w[n8 & 3] : 0;  // bytes after the whole lanes
    uint32_t rem = len & 7;
    p.k4 = 0;
    if (rem >= 4) { p.k4 = (tail & 0xFFFFFFFFULL) * XS_P64_1; tail >>= 32; rem -= 4; }
#pragma unroll
    for (int i = 0; i < 3; ++i) p.kb[i] = (uint32_t)i < rem ? ((tail >> (8 * i)) & 0xFF) * XS_P64_5 : 0;
}

XS_HD uint64_t xxh64_finish(const Xxh64Pre& p, uint32_t len, uint64_t seed) {
    uint64_t h;
    if (len == 32) {
        uint64_t v[4] = {seed + XS_P64_1 + XS_P64_2, seed + XS_P64_2, seed, seed - XS_P64_1};
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = rotl64(v[i] + p.k8[i], 31) * XS_P64_1;
        h = rotl64(v[0], 1) + rotl64(v[1], 7) + rotl64(v[2], 12) + rotl64(v[3], 18);
#pragma unroll
        for (int i = 0; i < 4; ++i) { h ^= xxh64_round(0, v[i]); h = h * XS_P64_1 + XS_P64_4; }
        h += 32;
        return xxh64_avalanche(h);
    }
    h = seed + XS_P64_5 + len;
    uint32_t n8 = len >> 3;
#pragma unroll
    for (int i = 0; i < 3; ++i)
        if ((uint32_t)i < n8) { h ^= p.k8[i]; h = rotl64(h, 27) * XS_P64_1 + XS_P64_4; }
    uint32_t rem = len & 7;
    if (rem >= 4) { h ^= p.k4; h = rotl64(h, 23) * XS_P64_2 + XS_P64_3; rem -= 4; }
#pragma unroll
    for (int i = 0; i < 3; ++i)
        if ((uint32_t)i < rem) { h ^= p.kb[i]; h = rotl64(h, 11) * XS_P64_1; }
    return xxh64_avalanche(h);
}

// ------------------------------------------------------------------ XXH3-64 (seed 0, default secret)
// first 64 bytes of XXH3_kSecret as little-endian words (len <= 32 needs bytes 0..55 only)
#define XS_SEC(i) ::xs::secret_word(i)
XS_HD uint64_t secret_word(int i) {
    switch (i) {
        case 0: return 0xbe4ba423396cfeb8ULL;
        case 1: return 0x1cad21f72c81017cULL;
        case 2: return 0xdb979083e96dd4deULL;
        case 3: return 0x1f67b3b7a4a44072ULL;
        case 4: return 0x78e5c0cc4ee679cbULL;
        case 5: return 0x2172ffcc7dd05a82ULL;
        case 6: return 0x8e2443f7744608b8ULL;
        default: return 0x4c263a81e69035e0ULL;  // 7
    }
}
#define XS_PRIME_MX1 0x165667919E3779F9ULL
#define XS_PRIME_MX2 0x9FB21C651E98DF25ULL

XS_HD uint64_t mul128_fold64(uint64_t a, uint64_t b) {
    return (a * b) ^ umul64hi(a, b);
}
XS_HD uint64_t xxh3_avalanche(uint64_t h) {
    h ^= h >> 37; h *= XS_PRIME_MX1; h ^= h >> 32; return h;
}
// unaligned 64-bit little-endian read at byte offset `off` (0 <= off <= 24) of the term
XS_HD uint64_t term_read64(const Term& t, uint32_t off) {
    uint32_t wi = off >> 3, s = (off & 7) * 8;
    uint64_t lo = t.w[wi & 3];
    uint64_t hi = wi + 1 < 4 ? t.w[(wi + 1) & 3] : 0;
    return s ? ((lo >> s) | (hi << (64 - s))) : lo;
}
XS_HD uint32_t term_read32(const Term& t, uint32_t off) {
    return (uint32_t)term_read64(t, off);  // off <= 24 always holds on the paths below
}

XS_HD uint64_t xxh3_64(const Term& t, uint32_t len) {
    if (len > 16) {  // 17..32
        uint64_t acc = (uint64_t)len * XS_P64_1;
        acc += mul128_fold64(t.w[0] ^ XS_SEC(0), t.w[1] ^ XS_SEC(1));
        acc += mul128_fold64(term_read64(t, len - 16) ^ XS_SEC(2), term_read64(t, len - 8) ^ XS_SEC(3));
        return xxh3_avalanche(acc);
    }
    if (len > 8) {   // 9..16
        uint64_t bf1 = XS_SEC(3) ^ XS_SEC(4), bf2 = XS_SEC(5) ^ XS_SEC(6);
        uint64_t lo = t.w[0] ^ bf1, hi = term_read64(t, len - 8) ^ bf2;
        uint64_t sw = ((uint64_t)byte_perm((uint32_t)lo, 0, 0x0123) << 32) | byte_perm((uint32_t)(lo >> 32), 0, 0x0123);
        uint64_t acc = (uint64_t)len + sw + hi + mul128_fold64(lo, hi);
        return xxh3_avalanche(acc);
    }
    if (len >= 4) {  // 4..8
        uint32_t in1 = (uint32_t)t.w[0], in2 = term_read32(t, len - 4);
        uint64_t bitflip = XS_SEC(1) ^ XS_SEC(2);
        uint64_t h = ((uint64_t)in2 + ((uint64_t)in1 << 32)) ^ bitflip;
        h ^= rotl64(h, 49) ^ rotl64(h, 24);
        h *= XS_PRIME_MX2;
        h ^= (h >> 35) + len;
        h *= XS_PRIME_MX2;
        return h ^ (h >> 28);
    }
    if (len > 0) {   // 1..3
        uint32_t c1 = (uint32_t)(t.w[0] & 0xFF), c2 = (uint32_t)((t.w[0] >> (8 * (len >> 1))) & 0xFF),
                 c3 = (uint32_t)((t.w[0] >> (8 * (len - 1))) & 0xFF);
        uint32_t combined = (c1 << 16) | (c2 << 24) | c3 | (len << 8);
        uint64_t s0 = XS_SEC(0);
        uint64_t bitflip = (uint64_t)((uint32_t)s0 ^ (uint32_t)(s0 >> 32));
        return xxh64_avalanche((uint64_t)combined ^ bitflip);
    }
    return 0;  // len 0 never occurs on this path (k >= 1 is enforced at open)
}

// ------------------------------------------------------------------ rbloom LCG
// state(u128) = state * M + 1; index = (u64)(state >> 32)
#define XS_LCG_MH 0x2360ED051FC65DA4ULL
#define XS_LCG_ML 0x4385DF649FCB5CEDULL
XS_HD uint64_t lcg_next(uint64_t& hi, uint64_t& lo) {
    uint64_t nlo = lo * XS_LCG_ML;
    uint64_t nhi = umul64hi(lo, XS_LCG_ML) + lo * XS_LCG_MH + hi * XS_LCG_ML;
    nlo += 1;
    nhi += (nlo == 0);
    hi = nhi; lo = nlo;
    return (hi << 32) | (lo >> 32);
}

}  // namespace xs
