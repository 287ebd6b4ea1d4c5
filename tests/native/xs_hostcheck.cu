// Host-side harness around the __host__ __device__ functions of xs_device.cuh, so that their
// logic (canonical form, ASCII expansion, XXH64, XXH3, LCG, Barrett) is checked against the
// oracle / python-xxhash on a box without a GPU.  Test infrastructure only.
#include <cstring>
#include "../../xspect2_b200/csrc/xs_device.cuh"

using namespace xs;

static void term_from_bytes(const uint8_t* p, uint32_t len, Term& t) {
    for (int i = 0; i < 4; ++i) t.w[i] = 0;
    for (uint32_t i = 0; i < len; ++i) t.w[i >> 3] |= (uint64_t)p[i] << (8 * (i & 7));
}

extern "C" {

uint64_t hc_xxh64(const uint8_t* p, uint32_t len, uint64_t seed) {
    Term t; term_from_bytes(p, len, t);
    Xxh64Pre pre; xxh64_prepare(t, len, pre);
    return xxh64_finish(pre, len, seed);
}
uint64_t hc_xxh3(const uint8_t* p, uint32_t len) {
    Term t; term_from_bytes(p, len, t);
    return xxh3_64(t, len);
}
// pack a pure-ACGT ascii string with the reference packing, extract window g, canonicalise,
// expand; returns the canonical term bytes in out[k] and the msb-first code
uint64_t hc_canonical(const uint8_t* ascii, uint64_t n, uint64_t g, uint32_t k, uint8_t* out, int* invalid) {
    uint64_t nw = n / 32 + 2;
    uint64_t* packed = new uint64_t[nw]();
    uint32_t* inv = new uint32_t[nw]();
    for (uint64_t i = 0; i < n; ++i) {
        uint8_t c = ascii[i];
        uint32_t code = ((c >> 1) ^ (c >> 2)) & 3;
        bool ok = c == 'A' || c == 'C' || c == 'G' || c == 'T';
        if (ok) packed[i >> 5] |= (uint64_t)code << (2 * (i & 31));
        else inv[i >> 5] |= 1u << (i & 31);
    }
    *invalid = window_invalid(inv, g, k) ? 1 : 0;
    uint64_t fr = window_lsb(packed, g, k);
    uint64_t msb;
    uint64_t c = canonical_lsb(fr, k, &msb);
    Term t; expand_ascii(c, k, t);
    for (uint32_t i = 0; i < k; ++i) out[i] = (uint8_t)(t.w[i >> 3] >> (8 * (i & 7)));
    // bytes beyond k must be zero
    for (uint32_t i = k; i < 32; ++i) if ((uint8_t)(t.w[i >> 3] >> (8 * (i & 7)))) *invalid |= 2;
    delete[] packed; delete[] inv;
    return msb;
}
void hc_literal(const uint8_t* bases, uint64_t g, uint32_t k, const uint8_t* comp, int canonicalize, uint8_t* out) {
    Term t;
    if (comp) literal_term(bases, g, k, TableComp{comp}, canonicalize != 0, t);
    else literal_term(bases, g, k, BioComp(), canonicalize != 0, t);
    for (uint32_t i = 0; i < k; ++i) out[i] = (uint8_t)(t.w[i >> 3] >> (8 * (i & 7)));
}
void hc_lcg(uint64_t h0, uint32_t n, uint64_t* out) {
    uint64_t hi = 0, lo = h0;
    for (uint32_t i = 0; i < n; ++i) out[i] = lcg_next(hi, lo);
}
uint8_t hc_biocomp(uint8_t c) { return BioComp()(c); }
uint8_t hc_cobscomp(uint8_t c) { return CobsComp()(c); }
uint64_t hc_mod(uint64_t x, uint64_t m) {
    uint64_t magic = m == 1 ? ~0ULL : (uint64_t)((((unsigned __int128)1) << 64) / m);
    return mod_barrett(x, m, magic);
}
uint32_t hc_mod_small(uint64_t x, uint32_t m) {   // m < 2^31
    uint64_t magic = m == 1 ? ~0ULL : (uint64_t)((((unsigned __int128)1) << 64) / m);
    return mod_barrett_small(x, m, magic);
}
}
