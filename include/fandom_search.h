/*
 * fandom_search.h -- C ABI of the B200-native reuse-search hot path.
 *
 * This is the drop-in boundary for the one step of senderle/fandom-search that
 * the reference runs as
 *
 *     pool.map(multi_search_wrapper, fan_cluster)          (search.py:381-386)
 *       -> AnnIndexSearch.search(filename)                 (search.py:163-226)
 *            -> mk_vectors / rolling 6-gram windows        (search.py:65-84,169-173)
 *            -> engine.neighbours(row) per window          (search.py:176-178)
 *            -> distance < distance_threshold              (search.py:182-184)
 *
 * The reference has no FFI; its seam is that Python call.  The entry points
 * below are what a ctypes binding of that seam calls (see INTEGRATION.md for
 * the reference-side stub).  Plain pointers and sizes only; no exceptions and
 * no C++/torch types cross this boundary.
 *
 * Conventions
 *   - Every function returns FS_OK (0) or a negative fs_status.  A human
 *     readable message for the last failure on the calling thread is available
 *     from fs_last_error().
 *   - "_dev" entry points take DEVICE pointers, enqueue work on `stream`
 *     (a cudaStream_t passed as void*) and do not synchronise.
 *     "_host" entry points take HOST pointers, do the H2D/D2H copies
 *     themselves and return after the stream has drained.
 *   - The caller owns every buffer it passes in.  An fs_index owns only its
 *     device copies and workspace.  An fs_index is bound to one device and is
 *     not thread safe: one per process/GPU, calls issued from one thread.
 *   - Token positions are indices into the CSR token array of the call
 *     (all works of the batch concatenated); `work` is the CSR row.
 */
#ifndef FANDOM_SEARCH_H_
#define FANDOM_SEARCH_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FS_ABI_VERSION 2

typedef enum fs_status {
    FS_OK = 0,
    FS_E_INVALID = -1,   /* bad argument                                          */
    FS_E_CUDA = -2,      /* a CUDA runtime/driver call failed                     */
    FS_E_OVERFLOW = -3,  /* an output/candidate buffer was too small; counters
                            hold the required sizes, call again with more room  */
    FS_E_NOMEM = -4,     /* device or host allocation failed                      */
    FS_E_NODEVICE = -5   /* no sm_100 device / driver entry point unavailable    */
} fs_status;

typedef struct fs_index fs_index; /* opaque */

/* One surviving (fan window, script window) pair.
 * Replaces one element of the list built at search.py:182-184:
 * (match_ix, match_str, distance) for the fan window fan_ix. */
typedef struct fs_match {
    int32_t fan_pos;    /* first token of the fan window, index into the CSR token array */
    int32_t script_pos; /* first token of the script window (= match_ix, search.py:183)  */
    double distance;    /* 1.0 - dot(unit(script window), unit(fan window)), float64     */
    int32_t work;       /* CSR row (fanwork) the window belongs to                       */
    uint32_t flags;     /* FS_MATCH_* bits                                               */
} fs_match;

#define FS_MATCH_EXACT 1u /* the six embedding-row ids are identical (distance-0 case,
                             what ao3.py:353-355 reads as BEST_COMBINED_DISTANCE <= 0) */

/* LSH-emulation mode (fs_index_set_lsh): bits [8,16) of flags hold 1 + the first hash table in
 * which the fan and script windows share a bucket key, 0 when they share none (the reference's
 * nearpy index would never have compared that pair). */
#define FS_MATCH_LSH_SHIFT 8

/* One winning record of the device-side post-processing (fs_search_submit_rows): the fan word
 * `word` of work `work` is best covered by the fan window that starts `window_ix` words before it,
 * matched with the script window starting at `match_ix`.  The CSV row of search.py:206-217 follows
 * from these six values (fs_records_format_csv). */
typedef struct fs_row {
    int32_t work;      /* CSR row (fanwork)                                             */
    int32_t word;      /* FAN_WORK_WORD_INDEX                                           */
    int32_t window_ix; /* position of the word inside the winning window (0 .. w-1)     */
    int32_t match_ix;  /* first script word of the matched script window               */
    double distance;   /* BEST_MATCH_DISTANCE                                           */
    int32_t lev;       /* BEST_LEVENSHTEIN_DISTANCE                                     */
    int32_t reserved;
} fs_row;

/* One exact 6-gram hit of the hash-join kernel. */
typedef struct fs_pair {
    int32_t fan_pos;
    int32_t script_pos;
} fs_pair;

/* Device-side counters written by the search entry points (all int64). */
enum {
    FS_CNT_CANDIDATES = 0, /* pairs that passed the tensor-core pre-filter          */
    FS_CNT_MATCHES = 1,    /* pairs with float64 distance < threshold                */
    FS_CNT_EXACT = 2,      /* pairs emitted by the hash-join                         */
    FS_CNT_WINDOWS = 3,    /* fan windows searched = sum max(T_i - w + 1, 0)         */
    FS_CNT_OVERFLOW = 4,   /* FS_OVERFLOW_* bits, written on the device by the search entry
                              points: how a caller of the stream-ordered "_dev" variants
                              learns that a buffer was too small                     */
    FS_CNT_ROWS = 5,       /* winning rows of the device-side post-processing            */
    FS_CNT_COUNT = 6
};
#define FS_OVERFLOW_CANDIDATES 1 /* the internal candidate buffer overflowed (pairs were lost):
                                    fs_index_reserve more candidates and search again        */
#define FS_OVERFLOW_MATCHES 2    /* more matches than `cap`: counters[FS_CNT_MATCHES] says how many */
#define FS_OVERFLOW_ROWS 4       /* more winning rows than `cap_rows`: counters[FS_CNT_ROWS] says how many */
#define FS_OVERFLOW_TEXT 8       /* a window text too long for the device Levenshtein (a token of 64 KiB,
                                    or both strings above 250 code points): the rows are incomplete, do
                                    this batch's records on the host (fs_records_best_mt)               */

/* Options for fs_index_set_option. */
enum {
    FS_OPT_SHIFTS_PER_STAGE = 1, /* MMA token-row shifts served by one smem stage (0 = all) */
    FS_OPT_GRID_LIMIT = 3,       /* max CTAs of the persistent distance kernel (0 = #SMs)  */
    FS_OPT_CTA_PAIR = 5,         /* 1: two CTAs of a cluster share one tcgen05.mma.cta_group::2
                                    (M = 2 x 128 fan windows, each CTA stages half the script tile) */
    FS_OPT_PACKED_SHUFFLE = 7,   /* 0: fp32 epilogue shuffles; 1: row shifts of E = 3, 6 move fp16x2
                                    pairs; 2: E = 6 sums the diagonal entirely in fp16x2 arithmetic */
    FS_OPT_A_RESIDENT = 6,       /* 1: (CTA pairs, dim <= 320) the fan tile stays resident in shared
                                    memory while the script tiles stream                         */
    FS_OPT_OPERAND_BITS = 9,     /* 16: fp16 operands (fixed pre-filter slack); 8: fp8 e4m3 operands
                                    (tcgen05.mma.kind::f8f6f4; the rounding error of every window is
                                    measured and enters its pre-filter threshold, so the candidates stay
                                    a guaranteed superset; CTA pairs only).  Re-converts the index.   */
    FS_OPT_TILE_GROUP = 10,      /* bit 0 (default on): with the resident fan tile, E = 6 and an embedding
                                    of at most three 128-byte chunks per row, all chunks of a script tile
                                    land on one barrier and are issued as one block of MMAs (off: one
                                    stage, one barrier and one issue block per chunk); bit 1 (default
                                    on): the fp16x2 epilogue hands the accumulator back as soon as its
                                    last column is packed, before the sums; bit 2 (default on, E = 6):
                                    an epilogue warp loads both of its 32-column chunks at once, hands
                                    the accumulator back, takes the row maxima on the fp32 values and
                                    packs to fp16x2 only the chunks that survive the bound; bit 5
                                    (default on, with bit 2): a chunk that survives is bounded again over
                                    four spans of 8 outputs before the diagonal sum is run; bit 6
                                    (default on, with bits 0 and 2, operand rows of <= 256 bytes): script
                                    tiles of 128 columns and two co-resident CTA pairs per TPC -- four
                                    accumulator stages in flight per SM instead of two; bit 4
                                    (default off, with bit 2): the fan row's bound stays in registers
                                    over the sweep of the script and the chunk bounds of the next tile
                                    are prefetched into shared memory by cp.async (no L2 round trip on
                                    the epilogue's critical path); bit 3 (default off, with bit 4): two
                                    sets of 8 epilogue warps drain alternate tiles                    */
    FS_OPT_PREFILTER_DIMS = 11,  /* embedding columns kept in the operand rows of the tensor-core pre-filter:
                                    -1 = automatic (default: whole 128-byte chunks of operand row holding
                                    ~5/6 of the table's energy: 256 of 300, 640 of 768), 0 = all, n = the
                                    n columns of largest energy.  What a
                                    window holds in the dropped columns enters its pre-filter threshold as
                                    |f_drop| |s_drop|, so the candidates stay a guaranteed superset and the
                                    float64 decision is unchanged.  Re-converts the index.             */
    FS_OPT_FUSED_GATHER = 12,    /* 1 (default): the distance kernel fetches the fan rows of its tile straight from
                                    the operand-row table by token id (TMA tile::gather4; search.py:65-84's
                                    mk_vectors never materialises a fan matrix) -- applies to the default
                                    128-column kernel (fs_index_get_info 15) and batches with at most 65536 extra
                                    rows; 0: the fan operand matrix is written by gather_kernel first.
                                    fs_index_get_info(16) tells which one the last search used.  Results are
                                    identical (the same bytes reach the tensor cores).                   */
    FS_OPT_DIAG = 4              /* 1 (dense), 2, 3 or 6: the tensor cores accumulate window/E shifts
                                    and the epilogue adds E diagonal neighbours (same products,
                                    E-fold fewer tensor-core flops)                            */
};

int fs_abi_version(void);
const char* fs_last_error(void);
int fs_device_count(void);

/*
 * Build the script-side index (replaces AnnIndexSearch.__init__ +
 * build_lsh_engine, search.py:131-154, 86-124).
 *
 *   table        host  [n_rows, dim] float32   embedding rows (spaCy vectors table)
 *   extra        host  [n_extra, dim] float32  rows for ids >= n_rows (OOV pseudo
 *                                              vectors of search.py:79-83); may be NULL
 *   script_tok   host  [n_script_tok] int32    embedding-row id per script token
 *   script_off   host  [n_scripts + 1] int64   CSR offsets; windows never straddle scripts
 *   window       6 in the reference (search.py:337)
 *   threshold    0.1 in the reference (search.py:340); matches have distance < threshold
 */
int fs_index_create(fs_index** out, int device,
                    const float* table, int64_t n_rows, int32_t dim,
                    const float* extra, int64_t n_extra,
                    const int32_t* script_tok, int64_t n_script_tok,
                    const int64_t* script_off, int64_t n_scripts,
                    int32_t window, double threshold);
int fs_index_destroy(fs_index* idx);

/* Pre-size the workspace (token capacity of one batch, candidate capacity).
 * Optional: search calls grow the workspace on demand (which synchronises). */
int fs_index_reserve(fs_index* idx, int64_t max_tokens, int64_t max_candidates);
int fs_index_set_option(fs_index* idx, int32_t option, int64_t value);
/* Parity mode: emulate a random-binary-projection LSH index with the given hyperplanes
 * (replaces the nearpy lookup of search.py:112-116,178 for a SEEDED reference run).
 *   normals  host [n_tables * n_bits, window * dim] float64, table-major
 * After this call every match carries the first-shared-table field described at
 * FS_MATCH_LSH_SHIFT.  n_tables = 0 switches the mode off. */
int fs_index_set_lsh(fs_index* idx, const double* normals, int32_t n_tables, int32_t n_bits);
/* what = 0: script windows, 1: dim_pad, 2: SM count, 3: candidate capacity, 4: shifts per stage,
 * 5: diagonal factor, 6: CTA pair, 7: A-resident, 8: pack level, 11: operand bits, 12: tile-group bits,
 * 13: embedding columns kept by the pre-filter, 14: share of the table's energy they hold (ppm),
 * 15: 1 when the 128-column kernel runs the current configuration,
 * 16: 1 when the last search fetched its fan rows by fused gather (FS_OPT_FUSED_GATHER) */
int64_t fs_index_get_info(const fs_index* idx, int32_t what);

/*
 * Search one batch of fanworks (replaces the pool.map over one cluster,
 * search.py:381-386, up to and including the threshold test at :182-184).
 *
 *   tok        [n_tok] int32   embedding-row id per fan token, works concatenated
 *   off        [n_works+1] int64 CSR offsets
 *   extra      [n_extra, dim] float32 rows for ids >= n_rows + (index extras); may be NULL
 *   out        [cap] fs_match  in arbitrary order
 *   counters   [FS_CNT_COUNT] int64
 *
 * Returns FS_E_OVERFLOW (host variant only) when counters[FS_CNT_MATCHES] > cap or
 * the internal candidate buffer overflowed; the first `cap` slots are valid.  The "_dev"
 * variant cannot return what the device has not computed yet: once its stream has drained the
 * caller MUST test counters[FS_CNT_OVERFLOW] (FS_OVERFLOW_* bits, set on the device).
 */
int fs_search_csr_dev(fs_index* idx, void* stream,
                      const int32_t* tok, int64_t n_tok,
                      const int64_t* off, int64_t n_works,
                      const float* extra, int64_t n_extra,
                      fs_match* out, int64_t cap, int64_t* counters);
int fs_search_csr_host(fs_index* idx,
                       const int32_t* tok, int64_t n_tok,
                       const int64_t* off, int64_t n_works,
                       const float* extra, int64_t n_extra,
                       fs_match* out, int64_t cap, int64_t* counters);

/*
 * Asynchronous form of fs_search_csr_host: up to two batches in flight per index.
 * fs_search_submit enqueues the host -> device copies (page-locked `tok`/`off`/`extra` make them
 * true DMA; the buffers must stay valid and unchanged until the matching collect returns), the
 * kernels and the counter read-back, and returns at once with a ticket; fs_search_collect waits for
 * that batch only, copies its matches out (on a stream of its own, under the next batch's kernels)
 * and reports overflow exactly like fs_search_csr_host.  Replaces the blocking pool.map of
 * search.py:381-386 by a pipeline: while batch k is searched, batch k+1 is already queued and the
 * host works on the records of batch k-1.  Tickets are collected in any order, each once.
 */
int fs_search_submit(fs_index* idx,
                     const int32_t* tok, int64_t n_tok,
                     const int64_t* off, int64_t n_works,
                     const float* extra, int64_t n_extra,
                     int64_t cap, int32_t* ticket);
int fs_search_collect(fs_index* idx, int32_t ticket, fs_match* out, int64_t cap, int64_t* counters);

/*
 * Search + post-processing on the device: top-10 per fan window
 * (NearestFilter(10)), Levenshtein of "[t0, ..., t5]" vs "s0 ... s5", six records per pair and the
 * per-word argmin with first-inserted ties (search.py:182-226) run on the GPU right behind the
 * float64 rescoring; only the winning rows come back, sorted by (work, word) -- the output of
 * fs_records_best on the same matches, bit for bit.
 *
 * fs_index_set_script_text registers the lower-cased script words (once per index):
 *   blob, word_off [n_words + 1]   n_words must equal the index's script token count
 * fs_search_submit_rows takes, besides the CSR batch, the verbatim token texts of the batch:
 *   text [text_bytes], tok_start [n_tok] uint32, tok_len [n_tok] uint16 (fs_batch arrays 0, 9, 10)
 *   lsh_filter != 0: LSH-emulation mode -- pairs that share no table are dropped and the first
 *   shared table orders equal distances, as in the reference's candidate order
 * Single-script indexes only.  FS_OVERFLOW_ROWS / FS_OVERFLOW_TEXT are reported like the other
 * overflow bits; on either the ticket can still be collected with fs_search_collect (raw matches).
 */
int fs_index_set_script_text(fs_index* idx, const char* blob, const int64_t* word_off, int64_t n_words);
int fs_search_submit_rows(fs_index* idx,
                          const int32_t* tok, int64_t n_tok,
                          const int64_t* off, int64_t n_works,
                          const float* extra, int64_t n_extra,
                          const char* text, int64_t text_bytes,
                          const uint32_t* tok_start, const uint16_t* tok_len,
                          int32_t lsh_filter, int64_t cap_matches, int64_t cap_rows, int32_t* ticket);
int fs_search_collect_rows(fs_index* idx, int32_t ticket, fs_row* out, int64_t cap, int64_t* counters);
/* Accumulate the per-script-word reuse histogram (see fs_reuse_histogram_dev below) from the winning
 * rows of every fs_search_submit_rows batch, on the device, before the rows are copied out:
 * counts_dev [n_script_tok, n_thr] int64 and thresholds_dev [n_thr] double are DEVICE buffers owned by
 * the caller (zero the counts first); n_thr = 0 switches it off.  A batch that ends with an overflow
 * bit set is not counted (the caller redoes it on the host and counts it there). */
int fs_index_set_reuse_histogram(fs_index* idx, int64_t* counts_dev, const double* thresholds_dev, int32_t n_thr);

/* Exact 6-gram hash-join only (the distance-0 special case). */
int fs_exact_join_dev(fs_index* idx, void* stream,
                      const int32_t* tok, int64_t n_tok,
                      const int64_t* off, int64_t n_works,
                      fs_pair* out, int64_t cap, int64_t* counters);
int fs_exact_join_host(fs_index* idx,
                       const int32_t* tok, int64_t n_tok,
                       const int64_t* off, int64_t n_works,
                       fs_pair* out, int64_t cap, int64_t* counters);

/*
 * Stage-level entry points used by the parity tests and by bench.py to time
 * one kernel at a time.  All pointers are DEVICE pointers.
 */
/* token gather + window norms: emb_out [n_tok, dim_pad] operand rows (e4m3 bytes, or fp16);
 * thr_out [n_tok][4] float = per window start (norm of the scaled window, norm of the rounding
 * error of its operand rows, norm of the columns the operand rows drop, 0), NaN where the window
 * leaves its work */
int fs_stage_embed_dev(fs_index* idx, void* stream,
                       const int32_t* tok, int64_t n_tok,
                       const int64_t* off, int64_t n_works,
                       const float* extra, int64_t n_extra,
                       void* emb_out, float* thr_out);
/* tensor-core window-vs-script contraction, dense dump: dots [n_tok, ld] float
 * (scaled by the index's global scale^2); for small inputs only. */
int fs_stage_dots_dev(fs_index* idx, void* stream,
                      const int32_t* tok, int64_t n_tok,
                      const int64_t* off, int64_t n_works,
                      const float* extra, int64_t n_extra,
                      float* dots, int64_t ld);
/* tensor-core pre-filter only: candidate pairs, no float64 rescoring */
int fs_stage_candidates_dev(fs_index* idx, void* stream,
                            const int32_t* tok, int64_t n_tok,
                            const int64_t* off, int64_t n_works,
                            const float* extra, int64_t n_extra,
                            fs_pair* out, int64_t cap, int64_t* counters);
/* Device time of the distance kernel, measured with CUDA events recorded on the
 * launching stream around every launch since the last reset (ring of 256 launches).
 * fs_timing_read synchronises on the recorded events. */
int fs_timing_reset(fs_index* idx);
int fs_timing_read(fs_index* idx, double* total_ms, int64_t* launches);
float fs_index_scale(const fs_index* idx);

/*
 * Host-side text helpers of the path (native, no GPU).
 */
/* Unit-cost edit distance over Unicode code points of two UTF-8 strings
 * (replaces Levenshtein.distance, search.py:14,190). */
int32_t fs_levenshtein_utf8(const char* a, int64_t a_len, const char* b, int64_t b_len);
/* MurmurHash64A(key, len, seed) -- spaCy's 64-bit string id uses seed 1
 * (FAN_WORK_ORTH_ID / ORIGINAL_SCRIPT_ORTH_ID columns, search.py:195,327). */
uint64_t fs_murmurhash64a(const void* key, int64_t len, uint64_t seed);
/* Split UTF-8 text on ASCII whitespace.  Writes up to cap (start,end) byte
 * offsets; returns the number of tokens found (may exceed cap). */
int64_t fs_tokenize_ws(const char* text, int64_t len, int64_t* starts, int64_t* ends, int64_t cap);

/*
 * Per-script-word reuse histogram (the aggregation ao3.py format_data does on the match CSV,
 * ao3.py:351-363,407-411): counts[word * n_thr + k] += (combined[i] <= thresholds[k]) for
 * every winning record i with ORIGINAL_SCRIPT_WORD_INDEX word_ix[i].  All DEVICE pointers;
 * counts [n_words, n_thr] int64 is accumulated into (zero it before the first cluster).
 */
int fs_reuse_histogram_dev(void* stream, const int32_t* word_ix, const double* combined, int64_t n,
                           const double* thresholds, int32_t n_thr, int64_t n_words, int64_t* counts);

/*
 * Native host stages either side of the GPU search (no GPU needed).
 */
typedef struct fs_vocab fs_vocab; /* lexicon key -> embedding-row id (Token.has_vector/.vector keys,
                                     search.py:74-75) */
typedef struct fs_batch fs_batch; /* one cluster of files read, tokenised and encoded to CSR */

/* keys_blob: the UTF-8 keys concatenated; key_offsets [n_keys+1]; rows [n_keys]. */
fs_vocab* fs_vocab_create(const char* keys_blob, const int64_t* key_offsets, const int32_t* rows,
                          int64_t n_keys);
void fs_vocab_destroy(fs_vocab* v);
int32_t fs_vocab_lookup(const fs_vocab* v, const char* key, int64_t len); /* row id or -1 */

/* Read `n_files` files with `n_threads` threads, split on ASCII whitespace, look every token up
 * (replaces the per-file read + tokenise + mk_vectors lookups of search.py:164-169).
 * Tokens without a vector get the id -(1+u), u numbering the batch's unique OOV strings in order
 * of first appearance. */
fs_batch* fs_batch_encode_files(const fs_vocab* v, const char* const* paths, int64_t n_files,
                                int32_t n_threads);
void fs_batch_destroy(fs_batch* b);
int64_t fs_batch_info(const fs_batch* b, int32_t what);  /* 0 files, 1 tokens, 2 unique OOV, 3 text bytes */
/* 0 text(char) 1 file_off(i64[files+1]) 2 tok_off(i64[files+1]) 3 tok(i32[T]) 4 tok_start(i64[T])
 * 5 tok_end(i64[T]) 6 oov_start(i64[U]) 7 oov_end(i64[U]) 8 file_status(i32[files])
 * 9 tok_start32(u32[T]) 10 tok_len16(u16[T], clamped at 65535: what fs_search_submit_rows takes);
 * valid until destroy */
void* fs_batch_array(fs_batch* b, int32_t which);

/* Top-k per fan window (NearestFilter(10)), Levenshtein of "[t0, ..., t5]" vs "s0 ... s5"
 * (search.py:123,189-190), six records per pair and the per-word argmin with first-inserted ties
 * (search.py:192-225).  Returns the number of winning rows (sorted by work, word), or -(rows needed)
 * when cap_out is too small. */
int64_t fs_records_best(const fs_match* matches, const int32_t* tie, int64_t n, int32_t window,
                        int32_t topk, const char* text, const int64_t* tok_start,
                        const int64_t* tok_end, const int64_t* tok_off, int64_t n_works,
                        const char* script_blob, const int64_t* script_word_off,
                        int64_t n_script_words, int32_t* out_work, int32_t* out_word,
                        int32_t* out_window_ix, int32_t* out_match_ix, double* out_distance,
                        int32_t* out_lev, int64_t cap_out);
/* The same on n_threads host threads (works are independent: the position-sorted match list is cut
 * at work boundaries); identical output. */
int64_t fs_records_best_mt(const fs_match* matches, const int32_t* tie, int64_t n, int32_t window,
                           int32_t topk, const char* text, const int64_t* tok_start,
                           const int64_t* tok_end, const int64_t* tok_off, int64_t n_works,
                           const char* script_blob, const int64_t* script_word_off,
                           int64_t n_script_words, int32_t* out_work, int32_t* out_word,
                           int32_t* out_window_ix, int32_t* out_match_ix, double* out_distance,
                           int32_t* out_lev, int64_t cap_out, int32_t n_threads);

/* CSV text of the winning records (arrays as returned by fs_records_best): the rows of
 * search.py:206-217 exactly as csv.writer(...).writerows writes them (search.py:331-334: excel
 * dialect, minimal quoting, "\r\n", None -> empty, floats as repr()).  names_*: FAN_WORK_FILENAME
 * per work; script_*: lower-cased script words by GLOBAL word index, their orth ids, character
 * (char_none = 1 -> None) and scene (scene_none = 1 -> None); word_base: first global word index
 * of the script the rows belong to.  *out_text is malloc'ed (release with fs_free); returns its
 * length or a negative fs_status. */
int64_t fs_records_format_csv(int64_t rows, const int32_t* work, const int32_t* word,
                              const int32_t* window_ix, const int32_t* match_ix, const double* distance,
                              const int32_t* lev, const char* names_blob, const int64_t* names_off,
                              const char* text, const int64_t* tok_start, const int64_t* tok_end,
                              const int64_t* tok_off, const char* script_blob,
                              const int64_t* script_word_off, const uint64_t* script_orth,
                              const char* char_blob, const int64_t* char_off, const uint8_t* char_none,
                              const int64_t* scene, const uint8_t* scene_none, int64_t word_base,
                              char** out_text);
void fs_free(void* p);
int64_t fs_format_py_float(double x, char* buf, int64_t cap); /* CPython repr(float) */

#ifdef __cplusplus
}
#endif
#endif /* FANDOM_SEARCH_H_ */
