// Host-side text helpers of the path (no GPU): edit distance, spaCy-compatible string
// ids, whitespace tokeniser.  Exposed through the same C ABI (include/fandom_search.h).
#include <stdint.h>
#include <string.h>

#include <utility>
#include <vector>

#include "../../include/fandom_search.h"

namespace {

// Decode UTF-8 into code points.  Malformed bytes decode as themselves so that the
// function is total; well-formed input gives exactly Python's str code points.
void decode_utf8(const char* s, int64_t n, std::vector<uint32_t>& out) {
    out.clear();
    out.reserve(static_cast<size_t>(n));
    const unsigned char* p = reinterpret_cast<const unsigned char*>(s);
    int64_t i = 0;
    while (i < n) {
        const unsigned char c = p[i];
        int len = 1;
        uint32_t cp = c;
        if (c >= 0xF0 && c < 0xF8) {
            len = 4;
            cp = c & 0x07;
        } else if (c >= 0xE0) {
            len = 3;
            cp = c & 0x0F;
        } else if (c >= 0xC0) {
            len = 2;
            cp = c & 0x1F;
        }
        if (len > 1) {
            bool ok = i + len <= n;
            for (int k = 1; ok && k < len; ++k) ok = (p[i + k] & 0xC0) == 0x80;
            if (ok) {
                for (int k = 1; k < len; ++k) cp = (cp << 6) | (p[i + k] & 0x3F);
            } else {
                len = 1;
                cp = c;
            }
        }
        out.push_back(cp);
        i += len;
    }
}

}  // namespace

extern "C" {

// Unit-cost Levenshtein distance over code points (replaces Levenshtein.distance,
// /root/reference search.py:14,190).  The shorter string is the pattern: up to 64 code points it
// runs as Myers' bit-vector algorithm in Hyyro's formulation for the global distance (one 64-bit
// word carries a whole column of the dynamic programme: ~12 word operations per code point of the
// longer string instead of one cell update per pair of code points); longer patterns fall back to
// the two-row dynamic programme.  Both give the same integer.
int32_t fs_levenshtein_utf8(const char* a, int64_t a_len, const char* b, int64_t b_len) {
    if (a_len < 0 || b_len < 0 || (a_len > 0 && !a) || (b_len > 0 && !b)) return -1;
    thread_local std::vector<uint32_t> ua, ub;
    thread_local std::vector<int32_t> row;
    decode_utf8(a, a_len, ua);
    decode_utf8(b, b_len, ub);
    const std::vector<uint32_t>& x = ua.size() >= ub.size() ? ua : ub;  // longer
    const std::vector<uint32_t>& y = ua.size() >= ub.size() ? ub : ua;  // shorter
    const size_t n = y.size();
    if (n == 0) return static_cast<int32_t>(x.size());
    if (n <= 64) {
        // match masks of the pattern: a direct table for ASCII, a short list for everything else
        thread_local uint64_t ascii_eq[128];
        thread_local std::vector<std::pair<uint32_t, uint64_t>> other_eq;
        other_eq.clear();
        for (size_t j = 0; j < n; ++j)
            if (y[j] < 128) ascii_eq[y[j]] = 0;
        for (size_t j = 0; j < n; ++j) {
            const uint64_t bit = 1ull << j;
            if (y[j] < 128) {
                ascii_eq[y[j]] |= bit;
            } else {
                bool found = false;
                for (auto& e : other_eq)
                    if (e.first == y[j]) {
                        e.second |= bit;
                        found = true;
                        break;
                    }
                if (!found) other_eq.emplace_back(y[j], bit);
            }
        }
        // characters of the text that do not occur in the pattern must read an all-zero mask:
        // the ASCII table is only valid for the pattern's own characters, so test membership
        thread_local uint8_t ascii_in[128];
        for (size_t j = 0; j < n; ++j)
            if (y[j] < 128) ascii_in[y[j]] = 1;
        const uint64_t top = 1ull << (n - 1);
        uint64_t pv = n == 64 ? ~0ull : ((1ull << n) - 1), mv = 0;
        int32_t score = static_cast<int32_t>(n);
        for (size_t i = 0; i < x.size(); ++i) {
            const uint32_t c = x[i];
            uint64_t eq = 0;
            if (c < 128) {
                if (ascii_in[c]) eq = ascii_eq[c];
            } else {
                for (const auto& e : other_eq)
                    if (e.first == c) {
                        eq = e.second;
                        break;
                    }
            }
            const uint64_t xv = eq | mv;
            const uint64_t xh = (((eq & pv) + pv) ^ pv) | eq;
            uint64_t ph = mv | ~(xh | pv);
            uint64_t mh = pv & xh;
            if (ph & top)
                ++score;
            else if (mh & top)
                --score;
            ph = (ph << 1) | 1ull;
            mh <<= 1;
            pv = mh | ~(xv | ph);
            mv = ph & xv;
        }
        for (size_t j = 0; j < n; ++j)
            if (y[j] < 128) ascii_in[y[j]] = 0;
        return score;
    }
    row.resize(n + 1);
    for (size_t j = 0; j <= n; ++j) row[j] = static_cast<int32_t>(j);
    for (size_t i = 1; i <= x.size(); ++i) {
        int32_t diag = row[0];
        row[0] = static_cast<int32_t>(i);
        const uint32_t xc = x[i - 1];
        for (size_t j = 1; j <= n; ++j) {
            const int32_t up = row[j];
            int32_t best = diag + (xc != y[j - 1] ? 1 : 0);
            if (up + 1 < best) best = up + 1;
            if (row[j - 1] + 1 < best) best = row[j - 1] + 1;
            row[j] = best;
            diag = up;
        }
    }
    return row[n];
}

// MurmurHash64A (Austin Appleby, public domain algorithm).  spaCy's string ids
// (Token.orth / Token.lower, search.py:195,327) are this hash of the UTF-8 bytes, seed 1.
uint64_t fs_murmurhash64a(const void* key, int64_t len, uint64_t seed) {
    const uint64_t m = 0xc6a4a7935bd1e995ULL;
    const int r = 47;
    uint64_t h = seed ^ (static_cast<uint64_t>(len) * m);
    const unsigned char* data = static_cast<const unsigned char*>(key);
    const int64_t nblocks = len / 8;
    for (int64_t i = 0; i < nblocks; ++i) {
        uint64_t k;
        memcpy(&k, data + i * 8, 8);
        k *= m;
        k ^= k >> r;
        k *= m;
        h ^= k;
        h *= m;
    }
    const unsigned char* tail = data + nblocks * 8;
    switch (len & 7) {
        case 7: h ^= static_cast<uint64_t>(tail[6]) << 48;  // fallthrough
        case 6: h ^= static_cast<uint64_t>(tail[5]) << 40;  // fallthrough
        case 5: h ^= static_cast<uint64_t>(tail[4]) << 32;  // fallthrough
        case 4: h ^= static_cast<uint64_t>(tail[3]) << 24;  // fallthrough
        case 3: h ^= static_cast<uint64_t>(tail[2]) << 16;  // fallthrough
        case 2: h ^= static_cast<uint64_t>(tail[1]) << 8;   // fallthrough
        case 1:
            h ^= static_cast<uint64_t>(tail[0]);
            h *= m;
    }
    h ^= h >> r;
    h *= m;
    h ^= h >> r;
    return h;
}

// Split on ASCII whitespace (space, \t, \n, \v, \f, \r) -- the tokenisation rule of the
// synthetic corpora and of the oracle's spaCy stand-in (oracle/shims/spacy).
int64_t fs_tokenize_ws(const char* text, int64_t len, int64_t* starts, int64_t* ends, int64_t cap) {
    int64_t n = 0;
    int64_t i = 0;
    auto is_ws = [](unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); };
    while (i < len) {
        while (i < len && is_ws(static_cast<unsigned char>(text[i]))) ++i;
        if (i >= len) break;
        const int64_t s = i;
        while (i < len && !is_ws(static_cast<unsigned char>(text[i]))) ++i;
        if (n < cap) {
            starts[n] = s;
            ends[n] = i;
        }
        ++n;
    }
    return n;
}

}  // extern "C"
