// Native host stages either side of the GPU search (host form):
//
//   fs_vocab_*   lexicon key -> embedding-row id hash map           (spaCy vocab lookup,
//                                                                    search.py:74-75)
//   fs_batch_*   multi-threaded file read + whitespace tokenise + id lookup -> CSR
//                (replaces the per-file Python/spaCy loop of search.py:164-169)
//   fs_records_* top-10 / Levenshtein / per-word argmin over the surviving pairs
//                (search.py:182-226), leaving only string formatting to Python
//
// No GPU code here; everything is exposed through the same C ABI.
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <sys/stat.h>
#include <unistd.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <atomic>
#include <memory>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/fandom_search.h"

namespace fs {
void set_error(const char* fmt, ...);
}

namespace {

inline bool is_ws(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); }

inline uint64_t hash_bytes(const char* p, int64_t n) { return fs_murmurhash64a(p, n, 0x5eedULL); }

}  // namespace

// ---------------------------------------------------------------------------------------
// vocabulary
// ---------------------------------------------------------------------------------------
struct fs_vocab {
    std::string blob;                 // concatenated keys
    std::vector<int64_t> key_off;     // [n+1]
    std::vector<int32_t> key_row;     // [n]
    std::vector<int32_t> slots;       // open addressing: key index or -1
    uint64_t mask = 0;

    int32_t find(const char* p, int64_t n) const {
        uint64_t s = hash_bytes(p, n) & mask;
        while (true) {
            const int32_t k = slots[s];
            if (k < 0) return -1;
            const int64_t a = key_off[k], len = key_off[k + 1] - a;
            if (len == n && memcmp(blob.data() + a, p, static_cast<size_t>(n)) == 0) return key_row[k];
            s = (s + 1) & mask;
        }
    }
};

// ---------------------------------------------------------------------------------------
// encoded batch of files
// ---------------------------------------------------------------------------------------
template <typename T>
struct RawArray {  // uninitialised storage (no zero fill, no page touching before the writers)
    std::unique_ptr<T[]> p;
    size_t n = 0;
    void alloc(size_t count) {
        p.reset(new T[count > 0 ? count : 1]);
        n = count;
    }
    T* data() { return p.get(); }
    const T* data() const { return p.get(); }
    size_t size() const { return n; }
};

struct fs_batch {
    RawArray<char> text;              // all file bytes concatenated
    std::vector<int64_t> file_off;    // [n_files+1] byte offsets into text
    std::vector<int64_t> tok_off;     // [n_files+1] CSR offsets (tokens)
    RawArray<int32_t> tok;            // [T] row id, or -(1+u) for the u-th unique OOV string
    RawArray<int64_t> tok_start;      // [T] byte offsets into text
    RawArray<int64_t> tok_end;        // [T]
    std::vector<int64_t> oov_start;   // [U] one representative span per unique OOV string
    std::vector<int64_t> oov_end;     // [U]
    std::vector<int32_t> file_status; // [n_files] 0 ok, 1 unreadable
};

namespace {

template <typename F>
void parallel_for(int64_t n, int n_threads, F&& body) {
    std::atomic<int64_t> next(0);
    auto work = [&]() {
        int64_t k;
        while ((k = next.fetch_add(1)) < n) body(k);
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
}

inline int64_t count_tokens(const char* text, int64_t len) {
    int64_t n = 0;
    bool in_tok = false;
    for (int64_t i = 0; i < len; ++i) {
        const bool ws = is_ws(static_cast<unsigned char>(text[i]));
        n += (!ws && !in_tok);
        in_tok = !ws;
    }
    return n;
}

}  // namespace

extern "C" {

fs_vocab* fs_vocab_create(const char* keys_blob, const int64_t* key_offsets, const int32_t* rows,
                          int64_t n_keys) {
    if (n_keys < 0 || (n_keys > 0 && (!keys_blob || !key_offsets || !rows))) {
        fs::set_error("fs_vocab_create: invalid argument");
        return nullptr;
    }
    fs_vocab* v = new fs_vocab();
    v->blob.assign(keys_blob, keys_blob + (n_keys ? key_offsets[n_keys] : 0));
    v->key_off.assign(key_offsets, key_offsets + n_keys + 1);
    if (n_keys == 0) v->key_off.assign(1, 0);
    v->key_row.assign(rows, rows + n_keys);
    uint64_t cap = 16;
    while (cap < static_cast<uint64_t>(n_keys) * 2 + 2) cap <<= 1;
    v->mask = cap - 1;
    v->slots.assign(cap, -1);
    for (int64_t k = 0; k < n_keys; ++k) {
        const int64_t a = v->key_off[k], len = v->key_off[k + 1] - a;
        uint64_t s = hash_bytes(v->blob.data() + a, len) & v->mask;
        bool dup = false;
        while (v->slots[s] >= 0) {
            const int32_t o = v->slots[s];
            const int64_t oa = v->key_off[o], olen = v->key_off[o + 1] - oa;
            if (olen == len && memcmp(v->blob.data() + oa, v->blob.data() + a, static_cast<size_t>(len)) == 0) {
                dup = true;  // duplicate key: the LAST entry wins, like building a Python dict
                v->slots[s] = static_cast<int32_t>(k);
                break;
            }
            s = (s + 1) & v->mask;
        }
        if (!dup) v->slots[s] = static_cast<int32_t>(k);
    }
    return v;
}

void fs_vocab_destroy(fs_vocab* v) { delete v; }

int32_t fs_vocab_lookup(const fs_vocab* v, const char* key, int64_t len) {
    if (!v || len < 0) return -1;
    return v->find(key, len);
}

// Read, tokenise and encode `n_files` files with `n_threads` threads.  Two passes over the
// text (count, then encode straight into the final arrays): no per-file vectors, no copies.
fs_batch* fs_batch_encode_files(const fs_vocab* v, const char* const* paths, int64_t n_files,
                                int32_t n_threads) {
    if (!v || n_files < 0 || (n_files > 0 && !paths)) {
        fs::set_error("fs_batch_encode_files: invalid argument");
        return nullptr;
    }
    auto T0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!getenv("FS_PROFILE_HOST")) return;
        auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[fs_batch_encode_files] %s %.1f ms\n", what,
                std::chrono::duration<double, std::milli>(t - T0).count());
        T0 = t;
    };
    fs_batch* b = new fs_batch();
    b->file_off.assign(n_files + 1, 0);
    b->tok_off.assign(n_files + 1, 0);
    b->file_status.assign(n_files, 0);
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_files) n_threads = static_cast<int32_t>(n_files > 0 ? n_files : 1);

    // sizes, then every file is read straight into its slot of the text buffer
    std::vector<int64_t> sizes(n_files, 0);
    parallel_for(n_files, n_threads, [&](int64_t k) {
        struct stat st;
        if (stat(paths[k], &st) != 0 || !S_ISREG(st.st_mode)) {
            b->file_status[k] = 1;
            return;
        }
        sizes[k] = static_cast<int64_t>(st.st_size);
    });
    for (int64_t k = 0; k < n_files; ++k) b->file_off[k + 1] = b->file_off[k] + sizes[k];
    b->text.alloc(static_cast<size_t>(b->file_off[n_files]));
    std::vector<int64_t> counts(n_files, 0);
    parallel_for(n_files, n_threads, [&](int64_t k) {
        if (b->file_status[k] || sizes[k] == 0) return;
        const int fd = open(paths[k], O_RDONLY);
        if (fd < 0) {
            b->file_status[k] = 1;
            return;
        }
        char* dst = b->text.data() + b->file_off[k];
        int64_t got = 0;
        while (got < sizes[k]) {
            const ssize_t r = read(fd, dst + got, static_cast<size_t>(sizes[k] - got));
            if (r <= 0) break;
            got += r;
        }
        close(fd);
        if (got != sizes[k])  // file changed under us: pad with spaces
            memset(dst + got, ' ', static_cast<size_t>(sizes[k] - got));
        counts[k] = count_tokens(dst, sizes[k]);
    });
    lap("read+count");
    for (int64_t k = 0; k < n_files; ++k) b->tok_off[k + 1] = b->tok_off[k] + counts[k];
    const int64_t T = b->tok_off[n_files];
    b->tok.alloc(static_cast<size_t>(T));
    b->tok_start.alloc(static_cast<size_t>(T));
    b->tok_end.alloc(static_cast<size_t>(T));
    parallel_for(n_files, n_threads, [&](int64_t k) {
        const char* text = b->text.data() + b->file_off[k];
        const int64_t len = sizes[k], base = b->file_off[k];
        int64_t o = b->tok_off[k];
        int32_t* tok = b->tok.data();
        int64_t* st = b->tok_start.data();
        int64_t* en = b->tok_end.data();
        int64_t i = 0;
        while (i < len) {
            while (i < len && is_ws(static_cast<unsigned char>(text[i]))) ++i;
            if (i >= len) break;
            const int64_t s = i;
            while (i < len && !is_ws(static_cast<unsigned char>(text[i]))) ++i;
            tok[o] = v->find(text + s, i - s);  // -1 = OOV for now
            st[o] = base + s;
            en[o] = base + i;
            ++o;
        }
    });
    lap("tokenise+lookup");
    // serial pass over the OOV tokens only: number the unique strings in order of appearance
    std::unordered_map<std::string, int32_t> uniq;
    const char* text = b->text.data();
    for (int64_t t = 0; t < T; ++t) {
        if (b->tok.data()[t] >= 0) continue;
        const int64_t s = b->tok_start.data()[t], e = b->tok_end.data()[t];
        std::string key(text + s, static_cast<size_t>(e - s));
        auto it = uniq.find(key);
        int32_t u;
        if (it == uniq.end()) {
            u = static_cast<int32_t>(uniq.size());
            uniq.emplace(std::move(key), u);
            b->oov_start.push_back(s);
            b->oov_end.push_back(e);
        } else {
            u = it->second;
        }
        b->tok.data()[t] = -(1 + u);
    }
    lap("oov");
    return b;
}

void fs_batch_destroy(fs_batch* b) { delete b; }

// what: 0 n_files, 1 n_tokens, 2 n_unique_oov, 3 text bytes
int64_t fs_batch_info(const fs_batch* b, int32_t what) {
    if (!b) return -1;
    switch (what) {
        case 0: return static_cast<int64_t>(b->file_status.size());
        case 1: return static_cast<int64_t>(b->tok.size());
        case 2: return static_cast<int64_t>(b->oov_start.size());
        case 3: return static_cast<int64_t>(b->text.size());
        default: return -1;
    }
}

// which: 0 text(char) 1 file_off(i64) 2 tok_off(i64) 3 tok(i32) 4 tok_start(i64) 5 tok_end(i64)
//        6 oov_start(i64) 7 oov_end(i64) 8 file_status(i32).  Pointers stay valid until destroy.
void* fs_batch_array(fs_batch* b, int32_t which) {
    if (!b) return nullptr;
    switch (which) {
        case 0: return b->text.data();
        case 1: return b->file_off.data();
        case 2: return b->tok_off.data();
        case 3: return b->tok.data();
        case 4: return b->tok_start.data();
        case 5: return b->tok_end.data();
        case 6: return b->oov_start.data();
        case 7: return b->oov_end.data();
        case 8: return b->file_status.data();
        default: return nullptr;
    }
}

// ---------------------------------------------------------------------------------------
// records: top-k per fan window, Levenshtein, per-word argmin (search.py:182-226)
// ---------------------------------------------------------------------------------------
// matches      [n] fs_match, any order (as the GPU emitted them)
// tie          [n] int32 secondary order key for equal distances (LSH first table) or NULL
// fan text:    text + tok_start/tok_end (as in fs_batch); tok_off[n_works+1] CSR offsets
// script text: script_blob + script_word_off[n_script_words+1] (lower-cased words)
// Output (capacity cap_out rows, returns the number of rows, or -(needed) if cap_out is too small):
//   out_work, out_word (fan word index inside its work), out_window_ix (0..w-1),
//   out_match_ix (script window start), out_distance, out_lev
// rows are sorted by (work, word).
int64_t fs_records_best(const fs_match* matches, const int32_t* tie, int64_t n, int32_t window,
                        int32_t topk, const char* text, const int64_t* tok_start,
                        const int64_t* tok_end, const int64_t* tok_off, int64_t n_works,
                        const char* script_blob, const int64_t* script_word_off,
                        int64_t n_script_words, int32_t* out_work, int32_t* out_word,
                        int32_t* out_window_ix, int32_t* out_match_ix, double* out_distance,
                        int32_t* out_lev, int64_t cap_out) {
    if (n < 0 || window < 1 || topk < 1 || (n > 0 && (!matches || !text || !tok_start || !tok_end ||
                                                     !tok_off || !script_blob || !script_word_off))) {
        fs::set_error("fs_records_best: invalid argument");
        return INT64_MIN;
    }
    std::vector<int64_t> order(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i) order[i] = i;
    // neighbours() yields candidates sorted by distance (stable over nearpy's candidate order:
    // first table, then script position); windows are visited in ascending position
    std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
        const fs_match& x = matches[a];
        const fs_match& y = matches[b];
        if (x.fan_pos != y.fan_pos) return x.fan_pos < y.fan_pos;
        if (x.distance != y.distance) return x.distance < y.distance;
        const int32_t tx = tie ? tie[a] : 0, ty = tie ? tie[b] : 0;
        if (tx != ty) return tx < ty;
        return x.script_pos < y.script_pos;
    });
    struct Best {
        double combined;
        double distance;
        int32_t window_ix, match_ix, lev;
        bool set;
    };
    std::vector<int64_t> keys;  // global fan token position of each winner
    std::unordered_map<int64_t, Best> best;
    best.reserve(static_cast<size_t>(n) * 2 + 16);
    std::string fan_ctx, match_str;
    int64_t run_start = 0;
    for (int64_t r = 0; r < n; ++r) {
        const fs_match& m = matches[order[r]];
        if (r > 0 && matches[order[r - 1]].fan_pos != m.fan_pos) run_start = r;
        if (r - run_start >= topk) continue;  // NearestFilter(10)
        if (m.script_pos < 0 || m.script_pos + window > n_script_words) continue;
        // str(list of Token) vs str(Span): "[a, b, c]" vs "a b c"  (search.py:123,189)
        fan_ctx.assign("[");
        match_str.clear();
        for (int k = 0; k < window; ++k) {
            const int64_t t = static_cast<int64_t>(m.fan_pos) + k;
            if (k) {
                fan_ctx.append(", ");
                match_str.push_back(' ');
            }
            fan_ctx.append(text + tok_start[t], static_cast<size_t>(tok_end[t] - tok_start[t]));
            const int64_t a = script_word_off[m.script_pos + k];
            match_str.append(script_blob + a, static_cast<size_t>(script_word_off[m.script_pos + k + 1] - a));
        }
        fan_ctx.push_back(']');
        const int32_t lev = fs_levenshtein_utf8(match_str.data(), static_cast<int64_t>(match_str.size()),
                                                fan_ctx.data(), static_cast<int64_t>(fan_ctx.size()));
        const double combined = m.distance * lev;
        for (int k = 0; k < window; ++k) {
            const int64_t pos = static_cast<int64_t>(m.fan_pos) + k;
            auto it = best.find(pos);
            if (it == best.end()) {
                best.emplace(pos, Best{combined, m.distance, k, m.script_pos, lev, true});
                keys.push_back(pos);
            } else if (combined < it->second.combined) {  // first minimal record wins ties
                it->second = Best{combined, m.distance, k, m.script_pos, lev, true};
            }
        }
    }
    const int64_t rows = static_cast<int64_t>(keys.size());
    if (rows > cap_out) return -rows;
    std::sort(keys.begin(), keys.end());
    int64_t w = 0;
    for (int64_t i = 0; i < rows; ++i) {
        const int64_t pos = keys[i];
        while (w + 1 < n_works && tok_off[w + 1] <= pos) ++w;
        const Best& bb = best[pos];
        out_work[i] = static_cast<int32_t>(w);
        out_word[i] = static_cast<int32_t>(pos - tok_off[w]);
        out_window_ix[i] = bb.window_ix;
        out_match_ix[i] = bb.match_ix;
        out_distance[i] = bb.distance;
        out_lev[i] = bb.lev;
    }
    return rows;
}

}  // extern "C"
