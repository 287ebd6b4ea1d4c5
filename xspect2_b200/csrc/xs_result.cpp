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
#include <string>
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
    std::vector<uint32_t> order(n_docs), first_order;
    std::vector<uint64_t> totals(n_docs, 0);
    uint64_t total_kmers = 0;
    ScoreCache cache;
    std::string scratch;
    auto row_order = [&](const uint32_t* row) {
        std::iota(order.begin(), order.end(), 0u);
        std::partial_sort(order.begin(), order.end(), order.end(), [row](uint32_t a, uint32_t b) { return row[a] > row[b]; });
    };
    auto rec_key = [&](uint64_t i, const char** p, size_t* len) {
        uint64_t b = i ? rec_key_end[i - 1] : 0;
        *p = rec_keys + b; *len = (size_t)(rec_key_end[i] - b);
    };
    // one pass per section; mode 0 = hits, 1 = scores
    auto section = [&](int mode) {
        for (uint64_t i = 0; i < n_emit; ++i) {
            const uint32_t* row = counts + rec_index[i] * (uint64_t)n_docs;
            row_order(row);
            if (mode == 0) {
                if (i == 0) { for (uint32_t d : order) if (doc_include[d]) first_order.push_back(d); }
                for (uint32_t d = 0; d < n_docs; ++d) totals[d] += row[d];
                total_kmers += num_kmers[i];
            }
            const char* kp; size_t kl;
            rec_key(i, &kp, &kl);
            o.put("        "); o.put(kp, kl); o.put(": {");
            bool any = false;
            for (uint32_t d : order) {
                if (!doc_include[d]) continue;
                o.put(any ? ",\n            " : "\n            ");
                any = true;
                o.put(dkey[d]); o.put(": ");
                if (mode == 0) o.put_u64(row[d]); else o.put(cache.get(row[d], num_kmers[i], scratch));
            }
            o.put(any ? "\n        }" : "}");
            if (mode == 0 && i + 1 < n_emit) o.put(",\n");
            if (mode == 1) o.put(",\n");
        }
    };
    o.put(prefix);
    o.put("    \"hits\": {\n");
    section(0);
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
        rec_key(i, &kp, &kl);
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
