// xs_result.cpp — native writer of the reference's ModelResult JSON (host only, no CUDA).
//
// ModelResult.save (models/result.py:178-189) is `json.dumps(self.to_dict(), indent=4)`; for a read set that is
// millions of nested Python dicts and `round(hits / num_kmers, 2)` calls (result.py:45-74).  This writer streams the
// same bytes straight from the count matrix: per record the documents in cobs result order
// (std::partial_sort, see xs_cobs_result_order), hits, then scores formatted like Python's repr of the rounded
// double (the shortest round-trip form of round(x, 2) is "%.2f" with trailing zeros cut back to one decimal),
// the "total" row, and num_kmers.  Keys arrive already JSON-escaped from the Python side.
#include <algorithm>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <cstdlib>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/xspect_b200.h"

int xs_set_error(int code, const std::string& msg);   // xs_lib.cu

namespace {

struct Out {
    FILE* f;
    std::vector<char> buf;
    size_t n = 0;
    bool ok = true;
    explicit Out(FILE* fp) : f(fp), buf(8 << 20) {}
    void flush() { if (n && fwrite(buf.data(), 1, n, f) != n) ok = false; n = 0; }
    void put(const char* p, size_t len) {
        if (len > buf.size()) { flush(); if (fwrite(p, 1, len, f) != len) ok = false; return; }
        if (n + len > buf.size()) flush();
        memcpy(buf.data() + n, p, len); n += len;
    }
    void put(const std::string& s) { put(s.data(), s.size()); }
    void put(const char* s) { put(s, strlen(s)); }
    void put_u64(uint64_t v) { char t[24]; int l = snprintf(t, sizeof t, "%llu", (unsigned long long)v); put(t, (size_t)l); }
};

// repr(round(h / n, 2)) for Python floats
std::string score_repr(uint64_t h, uint64_t n) {
    char t[64];
    int l = snprintf(t, sizeof t, "%.2f", (double)h / (double)n);
    while (l > 0 && t[l - 1] == '0' && l >= 2 && t[l - 2] != '.') --l;
    return std::string(t, (size_t)l);
}

struct ScoreCache {   // per num_kmers value: the strings of all possible scores (reads share a handful of lengths)
    std::unordered_map<uint64_t, std::vector<std::string>> tab;
    const std::string& get(uint64_t h, uint64_t n, std::string& scratch) {
        if (n > 4096 || h > n) { scratch = score_repr(h, n); return scratch; }
        auto it = tab.find(n);
        if (it == tab.end()) {
            std::vector<std::string> v(n + 1);
            for (uint64_t i = 0; i <= n; ++i) v[i] = score_repr(i, n);
            it = tab.emplace(n, std::move(v)).first;
        }
        return it->second[h];
    }
};

// one record's entry of the "hits" (mode 0) or "scores" (mode 1) section
struct RecordFormatter {
    const uint32_t* counts; uint32_t n_docs; const uint64_t* rec_index; uint64_t n_emit; const char* rec_keys;
    const uint64_t* rec_key_end; const uint64_t* num_kmers; const std::vector<std::string>* dkey; const uint8_t* doc_include;
    uint16_t* order_cache;   // [n_emit][n_docs] document order of every record, filled by mode 0 and reused by mode 1 (or null)

    void rec_key(uint64_t i, const char** p, size_t* len) const {
        uint64_t b = i ? rec_key_end[i - 1] : 0;
        *p = rec_keys + b; *len = (size_t)(rec_key_end[i] - b);
    }
    static void row_order(const uint32_t* row, std::vector<uint32_t>& order) {
        std::iota(order.begin(), order.end(), 0u);
        std::partial_sort(order.begin(), order.end(), order.end(), [row](uint32_t a, uint32_t b) { return row[a] > row[b]; });
    }
    // records [i0, i1) appended to `out`; mode 0 also adds their hits / k-mers to the running totals.  The text is
    // assembled with memcpy into a buffer grown per record to its worst case (no per-field allocation or printf).
    void format(int mode, uint64_t i0, uint64_t i1, std::string& out, std::vector<uint64_t>& totals, uint64_t& kmers,
                ScoreCache& cache) const {
        std::vector<uint32_t> order(n_docs);
        std::string scratch;
        size_t key_bytes = 0;
        for (uint32_t d = 0; d < n_docs; ++d) key_bytes += (*dkey)[d].size();
        const size_t per_record = 64 + key_bytes + (size_t)n_docs * (14 + 2 + 24);   // + the record key
        size_t used = out.size();
        auto put = [](char*& c, const char* p, size_t len) { memcpy(c, p, len); c += len; };
        auto put_u32 = [](char*& c, uint32_t v) {
            char t[10];
            int l = 0;
            do { t[l++] = (char)('0' + v % 10); v /= 10; } while (v);
            while (l) *c++ = t[--l];
        };
        for (uint64_t i = i0; i < i1; ++i) {
            const uint32_t* row = counts + rec_index[i] * (uint64_t)n_docs;
            if (mode == 1 && order_cache) {
                const uint16_t* oc = order_cache + i * (uint64_t)n_docs;
                for (uint32_t d = 0; d < n_docs; ++d) order[d] = oc[d];
            } else {
                row_order(row, order);
                if (order_cache) {
                    uint16_t* oc = order_cache + i * (uint64_t)n_docs;
                    for (uint32_t d = 0; d < n_docs; ++d) oc[d] = (uint16_t)order[d];
                }
            }
            if (mode == 0) {
                for (uint32_t d = 0; d < n_docs; ++d) totals[d] += row[d];
                kmers += num_kmers[i];
            }
            const char* kp; size_t kl;
            rec_key(i, &kp, &kl);
            if (out.size() < used + per_record + kl) out.resize(std::max(out.size() * 2, used + per_record + kl));
            char* c = &out[used];
            put(c, "        ", 8); put(c, kp, kl); put(c, ": {", 3);
            bool any = false;
            for (uint32_t d : order) {
                if (!doc_include[d]) continue;
                if (any) put(c, ",\n            ", 14); else put(c, "\n            ", 13);
                any = true;
                const std::string& key = (*dkey)[d];
                put(c, key.data(), key.size()); put(c, ": ", 2);
                if (mode == 0) put_u32(c, row[d]);
                else { const std::string& sc = cache.get(row[d], num_kmers[i], scratch); put(c, sc.data(), sc.size()); }
            }
            if (any) put(c, "\n        }", 10); else put(c, "}", 1);
            if (mode == 1 || i + 1 < n_emit) put(c, ",\n", 2);
            used = (size_t)(c - out.data());
        }
        out.resize(used);
    }
};

}  // namespace

extern "C" int xs_result_write_json(const char* path, const char* prefix, const char* suffix, const uint32_t* counts,
                                    uint32_t n_docs, const uint64_t* rec_index, uint64_t n_emit, const char* rec_keys,
                                    const uint64_t* rec_key_end, const uint64_t* num_kmers, const char* doc_keys,
                                    const uint64_t* doc_key_end, const uint8_t* doc_include) {
    if (!path || !prefix || !suffix || (n_emit && (!counts || !rec_index || !rec_keys || !rec_key_end || !num_kmers)) ||
        (n_docs && (!doc_keys || !doc_key_end || !doc_include)))
        return xs_set_error(XS_ERR_ARG, "NULL argument");
    if (n_emit == 0) return xs_set_error(XS_ERR_ARG, "a result without records has no total scores");
    FILE* f = fopen(path, "wb");
    if (!f) return xs_set_error(XS_ERR_IO, std::string(path) + ": " + strerror(errno));
    Out o(f);
    std::vector<std::string> dkey(n_docs);
    for (uint32_t d = 0; d < n_docs; ++d) dkey[d].assign(doc_keys + (d ? doc_key_end[d - 1] : 0), doc_keys + doc_key_end[d]);
    // the cobs result order of a record (std::partial_sort) is needed in both sections: kept between them when it fits
    std::vector<uint16_t> order_cache;
    if (n_docs <= 65535 && n_emit * (uint64_t)n_docs <= (512ULL << 20)) {
        try { order_cache.resize(n_emit * (uint64_t)n_docs); } catch (...) { order_cache.clear(); }
    }
    const RecordFormatter fmt{counts, n_docs, rec_index, n_emit, rec_keys, rec_key_end, num_kmers, &dkey, doc_include,
                              order_cache.empty() ? nullptr : order_cache.data()};
    // label order of the "total" row = the first record's (result.py:86)
    std::vector<uint32_t> first_order;
    {
        std::vector<uint32_t> order(n_docs);
        RecordFormatter::row_order(counts + rec_index[0] * (uint64_t)n_docs, order);
        for (uint32_t d : order) if (doc_include[d]) first_order.push_back(d);
    }
    std::vector<uint64_t> totals(n_docs, 0);
    uint64_t total_kmers = 0;

    // Records are formatted in blocks by a wave of host threads (every record's text is independent of its
    // neighbours); the blocks of a wave are written in order by a writer thread while the next wave is formatted
    // into the second buffer set.  A read set's JSON is gigabytes: ~4 KB per record.
    const uint64_t BLK = 2048;
    unsigned hw = std::thread::hardware_concurrency();
    if (const char* v = getenv("XS_RESULT_THREADS")) hw = (unsigned)std::max(1, atoi(v));
    const unsigned T = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(std::min<unsigned>(hw ? hw : 1, 32), (n_emit + BLK - 1) / BLK));
    std::vector<std::string> bufs[2] = {std::vector<std::string>(T), std::vector<std::string>(T)};
    std::vector<std::vector<uint64_t>> part_tot(T, std::vector<uint64_t>(n_docs, 0));
    std::vector<uint64_t> part_km(T, 0);
    std::vector<ScoreCache> caches(T);
    auto section = [&](int mode) {
        std::thread writer;
        int set = 0;
        for (uint64_t w0 = 0; w0 < n_emit; w0 += (uint64_t)T * BLK, set ^= 1) {
            std::vector<std::string>& cur = bufs[set];
            auto work = [&](unsigned t) {
                const uint64_t i0 = std::min<uint64_t>(n_emit, w0 + (uint64_t)t * BLK), i1 = std::min<uint64_t>(n_emit, i0 + BLK);
                cur[t].clear();
                if (i0 < i1) fmt.format(mode, i0, i1, cur[t], part_tot[t], part_km[t], caches[t]);
            };
            if (T == 1) {
                work(0);
            } else {
                std::vector<std::thread> th;
                th.reserve(T - 1);
                for (unsigned t = 1; t < T; ++t) th.emplace_back(work, t);
                work(0);
                for (auto& x : th) x.join();
            }
            if (writer.joinable()) writer.join();          // the previous wave is on disk; its buffer set is free again
            if (T == 1) { o.put(cur[0]); }
            else writer = std::thread([&o, &cur]() { for (const std::string& b : cur) o.put(b); });
        }
        if (writer.joinable()) writer.join();
    };
    o.put(prefix);
    o.put("    \"hits\": {\n");
    section(0);
    for (unsigned t = 0; t < T; ++t) {
        for (uint32_t d = 0; d < n_docs; ++d) totals[d] += part_tot[t][d];
        total_kmers += part_km[t];
    }
    o.put("\n    },\n    \"scores\": {\n");
    section(1);
    o.put("        \"total\": {");
    bool any = false;
    for (uint32_t d : first_order) {
        o.put(any ? ",\n            " : "\n            ");
        any = true;
        o.put(dkey[d]); o.put(": "); o.put(score_repr(totals[d], total_kmers));
    }
    o.put(any ? "\n        }" : "}");
    o.put("\n    },\n    \"num_kmers\": {\n");
    for (uint64_t i = 0; i < n_emit; ++i) {
        const char* kp; size_t kl;
        fmt.rec_key(i, &kp, &kl);
        o.put("        "); o.put(kp, kl); o.put(": "); o.put_u64(num_kmers[i]);
        if (i + 1 < n_emit) o.put(",\n");
    }
    o.put("\n    },\n");
    o.put(suffix);
    o.flush();
    bool ok = o.ok;
    if (fclose(f) != 0) ok = false;
    if (!ok) return xs_set_error(XS_ERR_IO, std::string(path) + ": write failed");
    return XS_OK;
}
