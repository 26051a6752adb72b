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
#include <emmintrin.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <sys/stat.h>
#include <unistd.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <atomic>
#include <charconv>
#include <memory>
#include <mutex>
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
// Open-addressing table whose 16-byte slots hold short keys INLINE: a word of up to 8 bytes (most
// of any vocabulary) is matched with one 64-bit compare against the slot itself -- one cache line
// per probe instead of slot -> key offsets -> key bytes -> row id.  Longer keys keep the 64-bit hash
// in the slot and are verified against the key blob.
struct fs_vocab {
    struct Slot {
        uint64_t key8;  // short key: its bytes, zero padded; long key: its 64-bit hash
        int32_t row;    // short key: embedding row id; long key: key index
        int32_t len;    // 0 = empty; 1..8 = short key length; -1 = long key
    };
    std::string blob;                 // concatenated keys
    std::vector<int64_t> key_off;     // [n+1]
    std::vector<int32_t> key_row;     // [n]
    std::vector<Slot> slots;
    uint64_t mask = 0;

    static uint64_t mix(uint64_t x) {  // splitmix64 finaliser
        x ^= x >> 30;
        x *= 0xbf58476d1ce4e5b9ULL;
        x ^= x >> 27;
        x *= 0x94d049bb133111ebULL;
        x ^= x >> 31;
        return x;
    }
    static uint64_t load8(const char* p, int64_t n) {  // n <= 8 bytes, zero padded
        uint64_t v = 0;
        memcpy(&v, p, static_cast<size_t>(n));
        return v;
    }
    // `padded`: at least 8 bytes are readable at p (the batch text buffer has that slack)
    int32_t find(const char* p, int64_t n, bool padded = false) const {
        if (n <= 0) return -1;
        if (n <= 8) {
            uint64_t k8;
            if (padded) {
                memcpy(&k8, p, 8);
                if (n < 8) k8 &= (1ULL << (8 * n)) - 1;
            } else {
                k8 = load8(p, n);
            }
            uint64_t s = mix(k8 + static_cast<uint64_t>(n)) & mask;
            while (true) {
                const Slot& e = slots[s];
                if (e.len == 0) return -1;
                if (e.len == n && e.key8 == k8) return e.row;
                s = (s + 1) & mask;
            }
        }
        const uint64_t h = hash_bytes(p, n);
        uint64_t s = h & mask;
        while (true) {
            const Slot& e = slots[s];
            if (e.len == 0) return -1;
            if (e.len == -1 && e.key8 == h) {
                const int64_t a = key_off[e.row], len = key_off[e.row + 1] - a;
                if (len == n && memcmp(blob.data() + a, p, static_cast<size_t>(n)) == 0) return key_row[e.row];
            }
            s = (s + 1) & mask;
        }
    }
    // insert or overwrite (duplicate key: the LAST entry wins, like building a Python dict)
    void insert(int64_t k) {
        const int64_t a = key_off[k], n = key_off[k + 1] - a;
        const char* p = blob.data() + a;
        if (n <= 0) return;
        if (n <= 8) {
            const uint64_t k8 = load8(p, n);
            uint64_t s = mix(k8 + static_cast<uint64_t>(n)) & mask;
            while (slots[s].len != 0 && !(slots[s].len == n && slots[s].key8 == k8)) s = (s + 1) & mask;
            slots[s] = Slot{k8, key_row[k], static_cast<int32_t>(n)};
            return;
        }
        const uint64_t h = hash_bytes(p, n);
        uint64_t s = h & mask;
        while (slots[s].len != 0) {
            const Slot& e = slots[s];
            if (e.len == -1 && e.key8 == h) {
                const int64_t oa = key_off[e.row], olen = key_off[e.row + 1] - oa;
                if (olen == n && memcmp(blob.data() + oa, p, static_cast<size_t>(n)) == 0) break;
            }
            s = (s + 1) & mask;
        }
        slots[s] = Slot{h, static_cast<int32_t>(k), -1};
    }
};

// ---------------------------------------------------------------------------------------
// encoded batch of files
// ---------------------------------------------------------------------------------------
// Batch buffers are recycled: a cluster needs ~100 MB of them (text, ids, offsets), a fresh
// allocation of that size comes from mmap and costs a page fault per 4 KB on first touch (a quarter of
// the tokenising time), and a run allocates the same sizes cluster after cluster.
namespace {
struct BlockPool {
    std::mutex mu;
    std::vector<std::pair<void*, size_t>> blocks;  // free blocks
    size_t held = 0;
    static constexpr size_t kMaxHeld = size_t(3) << 30;
    void* take(size_t bytes, size_t* cap) {
        {
            std::lock_guard<std::mutex> lock(mu);
            size_t best = blocks.size();
            for (size_t i = 0; i < blocks.size(); ++i)
                if (blocks[i].second >= bytes && blocks[i].second <= 2 * bytes + (1 << 20) &&
                    (best == blocks.size() || blocks[i].second < blocks[best].second))
                    best = i;
            if (best != blocks.size()) {
                void* p = blocks[best].first;
                *cap = blocks[best].second;
                held -= blocks[best].second;
                blocks.erase(blocks.begin() + static_cast<long>(best));
                return p;
            }
        }
        *cap = bytes + bytes / 8 + 4096;  // headroom: the next cluster is a few percent larger or smaller
        return malloc(*cap);
    }
    void give(void* p, size_t cap) {
        if (!p) return;
        {
            std::lock_guard<std::mutex> lock(mu);
            if (held + cap <= kMaxHeld && blocks.size() < 64) {
                blocks.emplace_back(p, cap);
                held += cap;
                return;
            }
        }
        free(p);
    }
};
BlockPool g_pool;
}  // namespace

template <typename T>
struct RawArray {  // uninitialised, recycled storage (no zero fill, no page touching before the writers)
    T* p = nullptr;
    size_t n = 0, cap_bytes = 0;
    RawArray() = default;
    RawArray(const RawArray&) = delete;
    RawArray& operator=(const RawArray&) = delete;
    ~RawArray() { g_pool.give(p, cap_bytes); }
    void alloc(size_t count) {
        g_pool.give(p, cap_bytes);
        p = static_cast<T*>(g_pool.take((count > 0 ? count : 1) * sizeof(T), &cap_bytes));
        n = count;
    }
    T* data() { return p; }
    const T* data() const { return p; }
    size_t size() const { return n; }
};

struct fs_batch {
    RawArray<char> text;              // all file bytes concatenated
    std::vector<int64_t> file_off;    // [n_files+1] byte offsets into text
    std::vector<int64_t> tok_off;     // [n_files+1] CSR offsets (tokens)
    RawArray<int32_t> tok;            // [T] row id, or -(1+u) for the u-th unique OOV string
    RawArray<int64_t> tok_start;      // [T] byte offsets into text
    RawArray<int64_t> tok_end;        // [T]
    RawArray<uint32_t> tok_start32;   // [T] the same offsets and lengths in the compact form the device
    RawArray<uint16_t> tok_len16;     // [T] takes (fs_search_submit_rows); lengths clamped at 65535
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

// Whitespace classes of 64 bytes at once (SSE2, part of every x86-64): bit i = text[i] is ' ' or
// 9..13.  The caller guarantees 64 readable bytes at p (the batch text buffer carries that slack).
inline uint64_t ws_mask64(const char* p) {
    const __m128i sp = _mm_set1_epi8(' '), nine = _mm_set1_epi8(9), four = _mm_set1_epi8(4);
    uint64_t m = 0;
    for (int k = 0; k < 4; ++k) {
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(p + 16 * k));
        const __m128i t = _mm_sub_epi8(c, nine);
        const __m128i in_range = _mm_cmpeq_epi8(_mm_min_epu8(t, four), t);  // 9 <= c <= 13 (unsigned)
        const __m128i w = _mm_or_si128(in_range, _mm_cmpeq_epi8(c, sp));
        m |= static_cast<uint64_t>(static_cast<uint32_t>(_mm_movemask_epi8(w))) << (16 * k);
    }
    return m;
}

// whitespace mask of the block at text + i, with everything at or beyond `len` counted as whitespace
inline uint64_t ws_block(const char* text, int64_t i, int64_t len) {
    uint64_t ws = ws_mask64(text + i);
    if (len - i < 64) ws |= ~0ull << (len - i);
    return ws;
}

inline int64_t count_tokens(const char* text, int64_t len) {
    int64_t n = 0;
    uint64_t prev_ws = 1;  // the position before the text counts as whitespace
    for (int64_t i = 0; i < len; i += 64) {
        const uint64_t ws = ws_block(text, i, len);
        n += __builtin_popcountll(~ws & ((ws << 1) | prev_ws));  // token starts
        prev_ws = ws >> 63;
    }
    return n;
}

// calls emit(start, end) for every maximal run of non-whitespace bytes of text[0, len)
template <typename F>
inline void for_each_token(const char* text, int64_t len, F&& emit) {
    bool in_tok = false;
    int64_t s = 0;
    for (int64_t blk = 0; blk < len; blk += 64) {
        const uint64_t ws = ws_block(text, blk, len);
        int pos = 0;
        while (pos < 64) {
            const uint64_t m = (in_tok ? ws : ~ws) & (~0ull << pos);
            if (!m) break;
            const int b = __builtin_ctzll(m);
            if (!in_tok) {
                s = blk + b;
                in_tok = true;
            } else {
                emit(s, blk + b);
                in_tok = false;
            }
            pos = b + 1;
        }
    }
    if (in_tok) emit(s, len);
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
    v->slots.assign(cap, fs_vocab::Slot{0, 0, 0});
    for (int64_t k = 0; k < n_keys; ++k) v->insert(k);
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
    b->text.alloc(static_cast<size_t>(b->file_off[n_files]) + 72);  // + slack for 64-byte blocks and 8-byte key loads
    memset(b->text.data() + b->file_off[n_files], ' ', 72);
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
    b->tok_start32.alloc(static_cast<size_t>(T));
    b->tok_len16.alloc(static_cast<size_t>(T));
    // Out-of-vocabulary tokens are collected per file while the text is hot in the tokenising thread's
    // cache: (token index, length, key) with key = the bytes themselves for words of <= 8 bytes, else
    // their 64-bit hash.  The serial pass below then only walks these compact lists.
    struct OovTok {
        int64_t t;
        uint64_t key;
        int32_t len;
    };
    std::vector<std::vector<OovTok>> oov_of_file(static_cast<size_t>(n_files));
    parallel_for(n_files, n_threads, [&](int64_t k) {
        const char* text = b->text.data() + b->file_off[k];
        const int64_t len = sizes[k], base = b->file_off[k];
        int64_t o = b->tok_off[k];
        int32_t* tok = b->tok.data();
        int64_t* st = b->tok_start.data();
        int64_t* en = b->tok_end.data();
        uint32_t* st32 = b->tok_start32.data();
        uint16_t* len16 = b->tok_len16.data();
        std::vector<OovTok>& oov = oov_of_file[static_cast<size_t>(k)];
        for_each_token(text, len, [&](int64_t s, int64_t i) {
            const int64_t n = i - s;
            const int32_t row = v->find(text + s, n, true);  // -1 = OOV for now
            tok[o] = row;
            st[o] = base + s;
            en[o] = base + i;
            st32[o] = static_cast<uint32_t>(base + s);
            len16[o] = static_cast<uint16_t>(n > 65535 ? 65535 : n);
            if (row < 0) {
                uint64_t key;
                if (n <= 8) {
                    memcpy(&key, text + s, 8);  // (the text buffer has slack behind its last byte)
                    if (n < 8) key &= (1ULL << (8 * n)) - 1;
                } else {
                    key = hash_bytes(text + s, n);
                }
                oov.push_back(OovTok{o, key, static_cast<int32_t>(n > INT32_MAX ? INT32_MAX : n)});
            }
            ++o;
        });
    });
    lap("tokenise+lookup");
    // Serial pass over the OOV tokens only: number the unique strings in order of appearance (files in
    // order, tokens in order).  Open addressing on (length, key); long words are verified against the
    // bytes of their first occurrence.
    struct OovSlot {
        uint64_t key;
        int64_t first_start;
        int32_t len;  // 0 = empty
        int32_t id;
    };
    std::vector<OovSlot> table(4096, OovSlot{0, 0, 0, 0});
    uint64_t tmask = table.size() - 1;
    const char* text = b->text.data();
    auto slot_of = [&](const std::vector<OovSlot>& tab, uint64_t mask, uint64_t key, int32_t len, int64_t start) {
        uint64_t sl = fs_vocab::mix(key + static_cast<uint64_t>(len)) & mask;
        while (true) {
            const OovSlot& e = tab[sl];
            if (e.len == 0) return sl;
            if (e.len == len && e.key == key &&
                (len <= 8 || memcmp(text + e.first_start, text + start, static_cast<size_t>(len)) == 0))
                return sl;
            sl = (sl + 1) & mask;
        }
    };
    for (int64_t k = 0; k < n_files; ++k) {
        for (const OovTok& w : oov_of_file[static_cast<size_t>(k)]) {
            const int64_t s = b->tok_start.data()[w.t];
            const int64_t wlen = b->tok_end.data()[w.t] - s;
            const int32_t len = static_cast<int32_t>(wlen > INT32_MAX ? INT32_MAX : wlen);
            uint64_t sl = slot_of(table, tmask, w.key, len, s);
            if (table[sl].len == 0) {
                table[sl] = OovSlot{w.key, s, len, static_cast<int32_t>(b->oov_start.size())};
                b->oov_start.push_back(s);
                b->oov_end.push_back(s + wlen);
                if (b->oov_start.size() * 2 > table.size()) {  // grow and re-insert
                    std::vector<OovSlot> bigger(table.size() * 4, OovSlot{0, 0, 0, 0});
                    const uint64_t bmask = bigger.size() - 1;
                    for (const OovSlot& e : table)
                        if (e.len != 0) bigger[slot_of(bigger, bmask, e.key, e.len, e.first_start)] = e;
                    table.swap(bigger);
                    tmask = bmask;
                    sl = slot_of(table, tmask, w.key, len, s);
                }
            }
            b->tok.data()[w.t] = -(1 + table[sl].id);
        }
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
        case 3: return b->file_off.empty() ? 0 : b->file_off.back();  // (the buffer has 8 bytes of slack behind)
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
        case 9: return b->tok_start32.data();
        case 10: return b->tok_len16.data();
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
    return fs_records_best_mt(matches, tie, n, window, topk, text, tok_start, tok_end, tok_off, n_works,
                              script_blob, script_word_off, n_script_words, out_work, out_word, out_window_ix,
                              out_match_ix, out_distance, out_lev, cap_out, 1);
}

// The same with `n_threads` host threads: works are independent (a window never leaves its work), so
// the position-sorted match list is cut at work boundaries into contiguous ranges of about equal
// length, every thread runs the single-pass algorithm on its range into its own buffers, and the
// buffers are concatenated in range order -- the rows come out sorted by (work, word) as before.
int64_t fs_records_best_mt(const fs_match* matches, const int32_t* tie, int64_t n, int32_t window,
                           int32_t topk, const char* text, const int64_t* tok_start,
                           const int64_t* tok_end, const int64_t* tok_off, int64_t n_works,
                           const char* script_blob, const int64_t* script_word_off,
                           int64_t n_script_words, int32_t* out_work, int32_t* out_word,
                           int32_t* out_window_ix, int32_t* out_match_ix, double* out_distance,
                           int32_t* out_lev, int64_t cap_out, int32_t n_threads) {
    if (n < 0 || window < 1 || topk < 1 || (n > 0 && (!matches || !text || !tok_start || !tok_end ||
                                                     !tok_off || !script_blob || !script_word_off))) {
        fs::set_error("fs_records_best: invalid argument");
        return INT64_MIN;
    }
    constexpr int kRing = 8;
    if (window > kRing) {
        fs::set_error("fs_records_best: window > 8");
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
    struct Row {
        int32_t work, word, window_ix, match_ix, lev;
        double distance;
    };
    // A fan word at global position `pos` receives records from the windows starting at
    // pos-w+1 .. pos, which arrive here in ascending order: a ring of 8 >= w open positions is
    // enough, and every position below the current window start is final and can be emitted --
    // no hash map, no final sort (rows come out ordered by position = by (work, word)).
    auto process = [&](int64_t r_begin, int64_t r_end, std::vector<Row>& rows_out) {
        struct Best {
            int64_t pos;
            double combined;
            double distance;
            int32_t window_ix, match_ix, lev;
        };
        Best ring[kRing];
        for (auto& e : ring) e.pos = -1;
        int64_t w = 0;
        auto emit = [&](const Best& bb) {
            while (w + 1 < n_works && tok_off[w + 1] <= bb.pos) ++w;
            rows_out.push_back(Row{static_cast<int32_t>(w), static_cast<int32_t>(bb.pos - tok_off[w]), bb.window_ix,
                                   bb.match_ix, bb.lev, bb.distance});
        };
        auto flush_below = [&](int64_t limit) {  // emit open positions < limit in ascending order
            int64_t lo = INT64_MAX;
            for (const auto& e : ring)
                if (e.pos >= 0 && e.pos < lo) lo = e.pos;
            if (lo == INT64_MAX) return;
            if (limit - lo > kRing) limit = lo + kRing;  // open positions span less than the ring
            for (int64_t pos = lo; pos < limit; ++pos) {
                Best& e = ring[pos % kRing];
                if (e.pos == pos) {
                    emit(e);
                    e.pos = -1;
                }
            }
        };
        std::string fan_ctx, match_str;
        int64_t run_start = r_begin, cur_p = -1;
        for (int64_t r = r_begin; r < r_end; ++r) {
            const fs_match& m = matches[order[r]];
            if (r > r_begin && matches[order[r - 1]].fan_pos != m.fan_pos) run_start = r;
            if (r - run_start >= topk) continue;  // NearestFilter(10)
            if (m.script_pos < 0 || m.script_pos + window > n_script_words) continue;
            if (m.fan_pos != cur_p) {  // positions before this window start are final
                flush_below(m.fan_pos);
                cur_p = m.fan_pos;
            }
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
                Best& e = ring[pos % kRing];
                if (e.pos != pos) {
                    e = Best{pos, combined, m.distance, k, m.script_pos, lev};
                } else if (combined < e.combined) {  // first minimal record wins ties
                    e = Best{pos, combined, m.distance, k, m.script_pos, lev};
                }
            }
        }
        flush_below(INT64_MAX);
    };
    // ranges of the sorted list, cut where the work changes
    if (n_threads < 1) n_threads = 1;
    if (n < 4096) n_threads = 1;
    std::vector<int64_t> cut(1, 0);
    for (int t = 1; t < n_threads; ++t) {
        int64_t r = n * t / n_threads;
        if (r <= cut.back()) continue;
        while (r < n && matches[order[r]].work == matches[order[r - 1]].work) ++r;
        if (r > cut.back() && r < n) cut.push_back(r);
    }
    cut.push_back(n);
    const int64_t n_ranges = static_cast<int64_t>(cut.size()) - 1;
    std::vector<std::vector<Row>> parts(static_cast<size_t>(n_ranges));
    parallel_for(n_ranges, static_cast<int>(n_ranges), [&](int64_t k) {
        parts[static_cast<size_t>(k)].reserve(static_cast<size_t>((cut[k + 1] - cut[k]) * 2 + 16));
        process(cut[k], cut[k + 1], parts[static_cast<size_t>(k)]);
    });
    int64_t rows = 0;
    for (const auto& part : parts) rows += static_cast<int64_t>(part.size());
    if (rows > cap_out) return -rows;
    int64_t o = 0;
    for (const auto& part : parts)
        for (const Row& r : part) {
            out_work[o] = r.work;
            out_word[o] = r.word;
            out_window_ix[o] = r.window_ix;
            out_match_ix[o] = r.match_ix;
            out_distance[o] = r.distance;
            out_lev[o] = r.lev;
            ++o;
        }
    return rows;
}


// ---------------------------------------------------------------------------------------
// CSV text of the winning records (search.py:206-217 rows through csv.writer, search.py:331-334)
// ---------------------------------------------------------------------------------------
}  // extern "C"

namespace {

// repr(float) of CPython ("short" float repr: shortest round-trip digits; exponent form iff
// decpt <= -4 or decpt > 16; exponent with sign and at least two digits)
void append_py_float(std::string& out, double x) {
    if (x != x) {
        out += "nan";
        return;
    }
    if (x == 1.0 / 0.0 || x == -1.0 / 0.0) {
        out += x > 0 ? "inf" : "-inf";
        return;
    }
    if (x == 0.0) {
        out += std::signbit(x) ? "-0.0" : "0.0";
        return;
    }
    char buf[64];
    auto res = std::to_chars(buf, buf + sizeof(buf), x, std::chars_format::scientific);
    const char* p = buf;
    const char* end = res.ptr;
    if (*p == '-') {
        out.push_back('-');
        ++p;
    }
    char digits[32];
    int nd = 0;
    while (p < end && *p != 'e') {
        if (*p != '.') digits[nd++] = *p;
        ++p;
    }
    int exp10 = 0;
    if (p < end && *p == 'e') {
        ++p;
        bool neg = false;
        if (*p == '+' || *p == '-') neg = *p++ == '-';
        while (p < end) exp10 = exp10 * 10 + (*p++ - '0');
        if (neg) exp10 = -exp10;
    }
    const int decpt = exp10 + 1;
    if (decpt <= -4 || decpt > 16) {
        out.push_back(digits[0]);
        if (nd > 1) {
            out.push_back('.');
            out.append(digits + 1, static_cast<size_t>(nd - 1));
        }
        out.push_back('e');
        const int e = decpt - 1;
        out.push_back(e < 0 ? '-' : '+');
        char eb[8];
        const int ae = e < 0 ? -e : e;
        auto r2 = std::to_chars(eb, eb + sizeof(eb), ae);
        if (r2.ptr - eb < 2) out.push_back('0');
        out.append(eb, static_cast<size_t>(r2.ptr - eb));
    } else if (decpt <= 0) {
        out += "0.";
        out.append(static_cast<size_t>(-decpt), '0');
        out.append(digits, static_cast<size_t>(nd));
    } else if (decpt >= nd) {
        out.append(digits, static_cast<size_t>(nd));
        out.append(static_cast<size_t>(decpt - nd), '0');
        out += ".0";
    } else {
        out.append(digits, static_cast<size_t>(decpt));
        out.push_back('.');
        out.append(digits + decpt, static_cast<size_t>(nd - decpt));
    }
}

template <typename I>
void append_int(std::string& out, I v) {
    char buf[32];
    auto r = std::to_chars(buf, buf + sizeof(buf), v);
    out.append(buf, static_cast<size_t>(r.ptr - buf));
}

// csv.writer, excel dialect, QUOTE_MINIMAL: quote a field that contains the delimiter, the quote
// character or a line-terminator character; double embedded quotes
void append_csv_field(std::string& out, const char* s, size_t n) {
    bool quote = false;
    for (size_t i = 0; i < n; ++i) {
        const char c = s[i];
        if (c == ',' || c == '"' || c == '\r' || c == '\n') {
            quote = true;
            break;
        }
    }
    if (!quote) {
        out.append(s, n);
        return;
    }
    out.push_back('"');
    for (size_t i = 0; i < n; ++i) {
        if (s[i] == '"') out.push_back('"');
        out.push_back(s[i]);
    }
    out.push_back('"');
}

}  // namespace

extern "C" {

// Formats `rows` winning records (arrays as returned by fs_records_best) as CSV text.
//   names_blob/names_off [n_works+1]     FAN_WORK_FILENAME per work
//   text/tok_start/tok_end/tok_off       fan tokens (as in fs_batch)
//   script_blob/script_word_off          ORIGINAL_SCRIPT_WORD (lower-cased), global word index
//   script_orth [n_script_words] u64     ORIGINAL_SCRIPT_ORTH_ID
//   char_blob/char_off [n_script_words+1], char_none [n_script_words] u8 (1 = None -> empty field)
//   scene [n_script_words] i64, scene_none [n_script_words] u8
//   word_base                            first global word index of the script (ORIGINAL_SCRIPT_WORD_INDEX
//                                        is local to its script)
// Returns a malloc'ed buffer in *out_text (free with fs_free) and its length, or a negative status.
int64_t fs_records_format_csv(int64_t rows, const int32_t* work, const int32_t* word,
                              const int32_t* window_ix, const int32_t* match_ix, const double* distance,
                              const int32_t* lev, const char* names_blob, const int64_t* names_off,
                              const char* text, const int64_t* tok_start, const int64_t* tok_end,
                              const int64_t* tok_off, const char* script_blob,
                              const int64_t* script_word_off, const uint64_t* script_orth,
                              const char* char_blob, const int64_t* char_off, const uint8_t* char_none,
                              const int64_t* scene, const uint8_t* scene_none, int64_t word_base,
                              char** out_text) {
    if (rows < 0 || !out_text || (rows > 0 && (!work || !word || !window_ix || !match_ix || !distance ||
                                               !lev || !names_blob || !names_off || !text || !tok_start ||
                                               !tok_end || !tok_off || !script_blob || !script_word_off ||
                                               !script_orth || !char_off || !char_none || !scene ||
                                               !scene_none))) {
        fs::set_error("fs_records_format_csv: invalid argument");
        return FS_E_INVALID;
    }
    std::string out;
    out.reserve(static_cast<size_t>(rows) * 160 + 16);
    for (int64_t i = 0; i < rows; ++i) {
        const int32_t w = work[i];
        const int64_t tpos = tok_off[w] + word[i];
        const int64_t g = static_cast<int64_t>(match_ix[i]) + window_ix[i];  // global script word
        append_csv_field(out, names_blob + names_off[w], static_cast<size_t>(names_off[w + 1] - names_off[w]));
        out.push_back(',');
        append_int(out, word[i]);
        out.push_back(',');
        const char* fw = text + tok_start[tpos];
        const size_t fn = static_cast<size_t>(tok_end[tpos] - tok_start[tpos]);
        append_csv_field(out, fw, fn);
        out.push_back(',');
        append_int(out, fs_murmurhash64a(fw, static_cast<int64_t>(fn), 1));
        out.push_back(',');
        append_int(out, g - word_base);
        out.push_back(',');
        append_csv_field(out, script_blob + script_word_off[g],
                         static_cast<size_t>(script_word_off[g + 1] - script_word_off[g]));
        out.push_back(',');
        append_int(out, script_orth[g]);
        out.push_back(',');
        if (!char_none[g])
            append_csv_field(out, char_blob + char_off[g], static_cast<size_t>(char_off[g + 1] - char_off[g]));
        out.push_back(',');
        if (!scene_none[g]) append_int(out, scene[g]);
        out.push_back(',');
        append_py_float(out, distance[i]);
        out.push_back(',');
        append_int(out, lev[i]);
        out.push_back(',');
        append_py_float(out, distance[i] * static_cast<double>(lev[i]));
        out += "\r\n";
    }
    char* buf = static_cast<char*>(malloc(out.size() + 1));
    if (!buf) {
        fs::set_error("fs_records_format_csv: out of memory");
        return FS_E_NOMEM;
    }
    memcpy(buf, out.data(), out.size());
    buf[out.size()] = 0;
    *out_text = buf;
    return static_cast<int64_t>(out.size());
}

void fs_free(void* p) { free(p); }

// repr(float) as CPython prints it (exposed for tests of the CSV formatter)
int64_t fs_format_py_float(double x, char* buf, int64_t cap) {
    std::string s;
    append_py_float(s, x);
    if (static_cast<int64_t>(s.size()) + 1 > cap) return -static_cast<int64_t>(s.size());
    memcpy(buf, s.data(), s.size());
    buf[s.size()] = 0;
    return static_cast<int64_t>(s.size());
}

}  // extern "C"
