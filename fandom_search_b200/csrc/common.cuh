// Shared device/host helpers for the reuse-search kernels (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/fandom_search.h"

namespace fs {

// ---------------------------------------------------------------------------
// error plumbing (no exceptions cross the C ABI)
// ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define FS_CUDA_CHECK(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            fs::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),        \
                          __FILE__, __LINE__);                                           \
            return FS_E_CUDA;                                                            \
        }                                                                                \
    } while (0)

// ---------------------------------------------------------------------------
// geometry of the distance kernel
// ---------------------------------------------------------------------------
constexpr int kBlockM = 128;            // fan windows per tile  (UMMA M, one TMEM lane each)
constexpr int kBlockN = 256;            // script windows per tile (UMMA N)
constexpr int kChunkK = 64;             // fp16 elements per smem row = 128 B = one swizzle row
constexpr int kUmmaK = 16;              // K of one tcgen05.mma.kind::f16
constexpr int kBoxRows = 136;           // rows per TMA box (128 + 8 halo rows for the shifts)
constexpr int kStageABytes = kBoxRows * 128;       // 17408 (17 swizzle atoms)
constexpr int kStageBBytes = 2 * kBoxRows * 128;   // 34816 (272 rows >= 256 + 5)
constexpr int kStageBytes = kStageABytes + kStageBBytes;  // 52224
constexpr int kPairStageBytes = 2 * kStageABytes;  // CTA-pair mode: A box + this CTA's half of B
constexpr int kAccumStages = 2;         // TMEM double buffer: 2 x 256 columns = all 512
constexpr int kTmemCols = 512;
constexpr int kEpiWarps = 16;           // 4 TMEM lane quarters x 4 column groups of 64
constexpr int kEpiColGroups = kEpiWarps / 4;
constexpr int kEpiCols = kBlockN / kEpiColGroups;      // 64 columns = 2 chunks of 32 per warp
constexpr int kProducerWarp = kEpiWarps;               // single-thread roles get the highest
constexpr int kMmaWarp = kEpiWarps + 1;                // warp ids (scheduler priority)
constexpr int kDistThreads = 32 * (kEpiWarps + 2);     // 576
constexpr int kHaloCols = kBlockN + 8;
constexpr int kNormTileBytes = kAccumStages * kHaloCols * 8;               // 4224: (B, half2(D, H)) per column
// E = 6 (one MMA shift, no row offsets inside a stage): the four 32-lane TMEM quarters of a tile
// hold OVERLAPPING fan rows, quarter q = rows [27 q, 27 q + 32) of the tile (four TMA boxes of 32
// rows), so every quarter sums its own diagonals -- no boundary rows to publish, no boundary
// pass, no barrier among the epilogue warps -- at the price of 108 instead of 123 windows per tile.
constexpr int kQuarterRows6 = 32 - 5;
constexpr int kOverlapBoxRows = 32;
constexpr int kOverlapABytes = 4 * kOverlapBoxRows * 128;  // 16384: fan bytes of one stage at E = 6
constexpr int kOverlapBBytes = 128 * 128;                  // 16384: script bytes of one box at E = 6 (no halo rows)
// fan windows produced per 128-row tile (tiles overlap by E-1 rows; E = 6: per quarter)
__host__ __device__ constexpr int dist_m_step(int diag) { return diag == 6 ? 4 * kQuarterRows6 : kBlockM - (diag - 1); }
__host__ __device__ constexpr int dist_pub_slots(int diag) { return (diag > 1 && diag != 6) ? 2 * (diag - 1) : 1; }
__host__ __device__ constexpr int dist_pub_bytes(int diag) {
    return diag == 6 ? 0 : 4 * dist_pub_slots(diag) * kHaloCols * 4;
}
// A-resident mode (CTA pairs, d_pad <= 320): the fan tile (all kAResChunks 64-column chunks,
// 87 KB) stays in shared memory for the whole sweep over the script tiles; only this CTA's half
// of the script tile streams through the stage ring (17 KB per stage).
constexpr int kAResChunks = 5;
constexpr int kAResBytes = kAResChunks * kStageABytes;  // 87040
// Grouped stages (A-resident, E = 6): a script tile's chunks share one full/empty barrier pair, and the
// ring also takes the part of the resident area that a narrow embedding leaves unused.
constexpr int kGroupMaxChunks = 3;
__host__ __device__ constexpr int dist_stage_bytes(bool pair, bool ares) {
    return ares ? kStageABytes : (pair ? 2 * kStageABytes : kStageBytes);
}
__host__ __device__ constexpr int dist_stages(int diag, bool pair, bool ares) {
    (void)diag;
    return ares ? 7 : (pair ? 6 : 4);
}
// per published boundary row and 32-column chunk, the maximum of the row over the chunk's 40 loaded
// columns (one half; 8 chunks per tile, two buffers): lets the boundary pass reject a chunk at once
__host__ __device__ constexpr int dist_rowmax_bytes(int diag) {
    return diag == 6 ? 0 : 2 * 4 * dist_pub_slots(diag) * 8 * 2;
}
constexpr int kMmStageBytes = kEpiWarps * 8 * 8;  // per epilogue warp: two buffers of four chunk bounds
__host__ __device__ constexpr int dist_smem_bytes(int diag, bool pair, bool ares) {
    return (ares ? kAResBytes : 0) + dist_stages(diag, pair, ares) * dist_stage_bytes(pair, ares) +
           1024 /*align slack*/ + 256 /*barriers*/ + dist_pub_bytes(diag) + kNormTileBytes +
           dist_rowmax_bytes(diag) + (diag == 6 ? kMmStageBytes : 0);
}
static_assert(dist_smem_bytes(1, false, false) <= 232448 && dist_smem_bytes(3, true, false) <= 232448 &&
                  dist_smem_bytes(6, true, false) <= 232448 && dist_smem_bytes(6, false, false) <= 232448 &&
                  dist_smem_bytes(3, true, true) <= 232448 && dist_smem_bytes(6, true, true) <= 232448 &&
                  dist_smem_bytes(1, true, true) <= 232448,
              "distance kernel exceeds the 227 KB shared memory limit");

// bit 4 of FS_OPT_TILE_GROUP with the one-pass epilogue (E = 6, fp16x2 sums, bit 2): the epilogue whose
// per-tile bounds are prefetched (distance.cu, epilogue_onepass_loop); bit 3 on top of it: two sets of
// 8 epilogue warps drain alternate tiles
__host__ __device__ constexpr bool dist_prefetch_epilogue(int diag, int pack, bool dump, int group_bits) {
    return diag == 6 && pack == 2 && !dump && (group_bits & 4) != 0 && (group_bits & 16) != 0;
}
__host__ __device__ constexpr bool dist_alt_sets(int diag, int pack, bool dump, int group_bits) {
    return dist_prefetch_epilogue(diag, pack, dump, group_bits) && (group_bits & 8) != 0;
}

struct DistParams {
    // pre-filter: keep (i, j) iff acc_ij > A_i * B_j - C_i * D_j - G_i * H_j  (window_norm_kernel, embed.cu).
    // The two slack factors of a side travel as ONE register, half2 rounded UP (a larger slack only
    // widens the superset): the epilogue's live state is what it was with two terms.
    const float2* fan_ac;       // [Mpad]  (A_i, half2(C_i, G_i)) = (|fan window|, (|its rounding error|, |its dropped
                                //         elements|)), A = NaN when invalid
    const float2* script_bd;    // [Npad]  (B_j, half2(D_j, H_j)) = ((1-thr-eps)|s| - err, (|s| + err, |dropped|)), B = NaN when invalid
    const float2* script_mm32;  // [Npad]  (min B, half2(max D, max H)) over columns j .. j+31
    int64_t n_fan_tok;        // rows of the fan token matrix (M)
    int64_t n_script_tok;     // rows of the script token matrix (N)
    int32_t chunks;           // ceil(dim_pad / 64) 64-column chunks
    int32_t last_chunk_ksteps;// UMMA K-steps (16 columns) in the last chunk: 1..4
    int32_t window;           // 6
    int32_t diag;             // E: epilogue adds E diagonal neighbours, MMAs do window/E shifts
    int32_t pack;             // 1: fp16x2-packed epilogue shuffles (E = 3, 6)
    int32_t ares;             // 1: A-resident variant (pair mode, chunks <= kAResChunks)
    int32_t f8;               // 1: operands are fp8 e4m3 (tcgen05.mma.kind::f8f6f4, K = 32)
    int32_t pair;             // 1: CTA-pair kernel (cta_group::2, M = 2 x 128 fan tiles)
    int32_t shifts_per_stage; // S: MMA shifts served by one smem stage (divides window/E)
    int32_t group;            // 1: (A-resident, E = 6, chunks <= kGroupMaxChunks) all chunks of a script tile
                              //    land on ONE barrier and are issued as one block of MMAs
    int32_t wait_mode;        // experiment switch of the barrier waits (FS_DEBUG_WAIT; see mbar_wait_mode)
    int32_t tiles_m, tiles_n;
    fs_pair* cand;            // candidate output
    int64_t cand_cap;
    unsigned long long* counters;  // [FS_CNT_COUNT]
    float* dump;              // optional dense dump [n_fan_tok, dump_ld]
    int64_t dump_ld;
    // fused gather (128-column kernel only): the fan tile is fetched by TMA tile::gather4 straight from the
    // operand-row table [table | script extras | this batch's extras] by token id -- the fan operand matrix is
    // never materialised.  Ids outside [0, n_valid_rows) fetch row n_table_rows, which lies outside the
    // tensor map and arrives as zeros (what gather_kernel writes for unknown ids).
    const int32_t* fan_tok;   // [n_fan_tok] row ids, or nullptr: the fan tile comes from the materialised matrix
    int32_t n_valid_rows;
    int32_t n_table_rows;
};


// ---------------------------------------------------------------------------
// embedding-row sources: ids [0, n_base) -> base table, then the script-side
// extras (OOV rows registered with the index), then the per-batch fan extras
// ---------------------------------------------------------------------------
struct RescoreParams {
    const fs_pair* cand;
    const unsigned long long* counters;  // n candidates at FS_CNT_CANDIDATES
    int64_t cand_cap;
    const int32_t* fan_tok;
    int64_t n_fan_tok;
    const int64_t* fan_off;
    int32_t n_works;
    const int32_t* script_tok;
    const float* table;  // [n_base, dim] fp32
    int64_t n_base;
    const float* script_extra;  // [n_script_extra, dim]
    int64_t n_script_extra;
    const float* fan_extra;  // [n_fan_extra, dim]
    int64_t n_fan_extra;
    int32_t dim;
    int32_t window;
    double threshold;
    fs_match* out;
    int64_t out_cap;
    unsigned long long* match_counter;
    unsigned long long* overflow;  // FS_OVERFLOW_* bits (counters + FS_CNT_OVERFLOW)
};

// device-side post-processing of the match list (postprocess.cu)
struct PostParams {
    const fs_match* matches;
    unsigned long long* counters;  // reads FS_CNT_MATCHES, writes FS_CNT_ROWS
    int64_t match_cap;
    int32_t n_tok, window, topk, lsh;
    const int64_t* fan_off;
    int32_t n_works;
    // fan tokens: bytes of token t = fan_text[tok_start[t] .. + tok_len[t])
    const uint8_t* fan_text;
    const uint32_t* tok_start;
    const uint16_t* tok_len;
    // script words (lower-cased): bytes of word j = script_text[script_word_off[j] .. script_word_off[j+1])
    const uint8_t* script_text;
    const int64_t* script_word_off;
    int64_t n_script_words;
    // workspace
    int32_t* head;                 // [n_tok] first match of the window starting here
    int32_t* next;                 // [match_cap]
    int32_t* m_lev;                // [match_cap] -1 = dropped
    int32_t* m_rank;               // [match_cap]
    unsigned long long* best_key;  // [n_tok] minimal combined distance (order-preserving bits)
    unsigned long long* best_tie;  // [n_tok] first inserted among the minimal records
    int32_t* winner;               // [n_tok] winning match, -1 = none
    int32_t* block_count;          // [ceil(n_tok / 1024)]
    int64_t* block_off;
    fs_row* rows;
    int64_t rows_cap;
    unsigned long long* overflow;
};

struct LshParams {
    fs_match* matches;
    const unsigned long long* match_counter;
    int64_t match_cap;
    const double* normals;  // [n_tables * n_bits, window * dim]
    int32_t n_tables, n_bits;
    const int32_t* fan_tok;
    const int32_t* script_tok;
    const float* table;
    int64_t n_base;
    const float* script_extra;
    int64_t n_script_extra;
    const float* fan_extra;
    int64_t n_fan_extra;
    int32_t dim, window;
};

// Operand rows are raw bytes: fp16 (2 B per element) or fp8 e4m3 (1 B per element); the plumbing
// counts a row in 2-byte units ("dim_pad": fp16 elements, or fp8 elements / 2), so the gather, the
// tensor maps (plain byte movers) and the 128-byte chunking are the same for both.
// *_sq: per row (squared norm of the scaled fp32 row, squared norm of the rounding error of its kept
// elements, squared norm of its dropped elements, unused).
struct GatherSources {
    const __half* base16;
    const float4* base_sq;
    int64_t n_base;
    const __half* sx16;  // script extras
    const float4* sx_sq;
    int64_t n_sx;
    const __half* fx16;  // fan extras of this batch
    const float4* fx_sq;
    int64_t n_fx;
};

bool distance_uses_n128(const DistParams& p);  // the 128-column kernel (the only one with the fused gather)
int make_token_map(CUtensorMap* map, const void* base, int64_t rows, int32_t dim_pad, int32_t box_rows);
int launch_distance(const CUtensorMap& map_fan, const CUtensorMap& map_fan32, const CUtensorMap& map_script,
                    const CUtensorMap& map_script128, const CUtensorMap& map_script64, const DistParams& p,
                    int grid_limit, cudaStream_t stream);
int launch_convert_rows(const float* src, int64_t n_rows, int32_t dim, int32_t dim_pad, int32_t kept,
                        const int32_t* perm, float scale, bool f8, float limit_sq, __half* dst, float4* sq,
                        cudaStream_t stream);
int launch_column_energy(const float* src, int64_t n_rows, int32_t dim, double* energy, cudaStream_t stream);
int launch_rownorm_max(const float* src, int64_t n_rows, int32_t dim, unsigned int* out, cudaStream_t stream);
int launch_absmax(const float* src, int64_t n, unsigned int* out, cudaStream_t stream);
int launch_gather(const int32_t* tok, int64_t n_tok, const GatherSources& src, int32_t dim_pad,
                  __half* emb, float4* tok_sq, int sm_count, cudaStream_t stream);
// tok_sq only (fused gather: the operand rows stay in the table)
int launch_gather_sq(const int32_t* tok, int64_t n_tok, const GatherSources& src, float4* tok_sq, int sm_count,
                     cudaStream_t stream);
int launch_window_norm(const float4* tok_sq, int64_t n_tok, const int64_t* off, int32_t n_rows,
                       int32_t window, float coef, bool script_side, float2* out, float4* out_plain,
                       int64_t n_pad, unsigned long long* window_counter, cudaStream_t stream);
int launch_sliding_minmax32(const float2* src, float2* dst, int64_t n, cudaStream_t stream);
int launch_rescore(const RescoreParams& p, int sm_count, cudaStream_t stream);
int launch_lsh(const LshParams& p, int sm_count, cudaStream_t stream);
int launch_postprocess(const PostParams& p, int sm_count, cudaStream_t stream);
int launch_reuse_histogram_rows(const fs_row* rows, const unsigned long long* counters, int64_t rows_cap,
                                const double* thresholds, int32_t n_thr, int64_t n_words,
                                unsigned long long* counts, int sm_count, cudaStream_t stream);
int64_t postprocess_scan_blocks(int64_t n_tok);
uint32_t hash_filter_bits(int64_t n_script_tok);
int launch_hash_build(const int32_t* tok, int64_t n_tok, const int64_t* off, int32_t n_rows,
                      int32_t window, unsigned long long* table, uint32_t slots, uint32_t* filter,
                      uint32_t filter_bits, cudaStream_t stream);
int launch_hash_probe(const int32_t* tok, int64_t n_tok, const int64_t* off, int32_t n_rows,
                      const int32_t* script_tok, int32_t window, const unsigned long long* table,
                      uint32_t slots, const uint32_t* filter, uint32_t filter_bits, fs_pair* out, int64_t cap,
                      unsigned long long* counter, int sm_count, cudaStream_t stream);

// ---------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// Whole-warp wait with ONE polling lane (18 warps x 32 lanes spinning on try_wait is needless
// pressure on the barrier unit): one lane polls (optionally backing off), the warp re-converges,
// and every lane then performs one try_wait that succeeds immediately (its own acquire).
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity, uint32_t backoff_ns) {
    if ((threadIdx.x & 31) == 0) {
        while (!mbar_try_wait(bar, parity)) {
            if (backoff_ns) __nanosleep(backoff_ns);
        }
    }
    __syncwarp();
    while (!mbar_try_wait(bar, parity)) {
    }
}

// try_wait with a suspend-time hint: the hardware parks the thread until the phase completes or `ns`
// have passed -- no issue slots and no polling traffic while parked
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(ns)
        : "memory");
    return ok != 0;
}
// non-blocking probe of a phase (no hardware parking at all)
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// experiment switch of the epilogue's accumulator wait (DistParams::wait_mode & 15, FS_DEBUG_WAIT):
//  0 one polling lane on try_wait, then the warp (the round-1 form)   1 the same, napping 40 ns
//  2 every lane parked on try_wait with a 4 us hint                   3 one lane parked with the hint, then the warp
//  8 one lane spinning on test_wait, then the warp                    9 every lane spinning on test_wait
// 10 one lane on try_wait with a 100 ns hint, then the warp          11 every lane on try_wait with a 100 ns hint
__device__ __forceinline__ void mbar_wait_mode(uint32_t bar, uint32_t parity, int mode) {
    // 5, 6 (default), 7: every lane; a failed first check is followed by ONE sleep of 100 / 200 / 400 ns, then the
    // suspend-hint loop of mode 2.  +2.4 .. 3.5 % over mode 2 in four A/B runs on the same box (the accumulator
    // is never ready sooner than that after a miss, and every re-check is three instructions per warp).
    if (mode >= 5 && mode <= 7) {
        if (mbar_try_wait(bar, parity)) return;
        __nanosleep(mode == 5 ? 100u : mode == 6 ? 200u : 400u);
        while (!mbar_try_wait_hint(bar, parity, 4000u)) {
        }
        return;
    }
    if (mode == 2 || mode == 11) {
        const uint32_t ns = mode == 2 ? 4000u : 100u;
        while (!mbar_try_wait_hint(bar, parity, ns)) {
        }
        return;
    }
    if (mode == 9) {
        while (!mbar_test_wait(bar, parity)) {
        }
        return;
    }
    if ((threadIdx.x & 31) == 0) {
        if (mode == 3 || mode == 10) {
            const uint32_t ns = mode == 3 ? 4000u : 100u;
            while (!mbar_try_wait_hint(bar, parity, ns)) {
            }
        } else if (mode == 8) {
            while (!mbar_test_wait(bar, parity)) {
            }
        } else {
            while (!mbar_try_wait(bar, parity)) {
                if (mode == 1) __nanosleep(40);
            }
        }
    }
    __syncwarp();
    while (!mbar_try_wait(bar, parity)) {
    }
}

// every lane polls: for waits that almost always succeed at once (the MMA issuer's)
__device__ __forceinline__ void mbar_wait_all(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// one elected lane of a converged warp (the same lane every time: lowest active)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .b32 rx;\n\t"
        ".reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, px;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// 2-D tiled TMA load: box of the tensor map at (col, row) -> swizzled smem, completes on mbarrier
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int32_t col, int32_t row) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(col), "r"(row)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, fp16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     bar)
                 : "memory");
}
// ---- CTA-pair (cta_group::2) variants ------------------------------------------------
// In a 2-CTA cluster the shared::cluster address of the odd CTA has bit 24 set; clearing it
// addresses the same offset in the even ("leader") CTA.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same smem offset in the leader CTA (cluster rank 0)
__device__ __forceinline__ uint32_t leader_addr(uint32_t local) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(0u));
    return r;
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {  // arrive on the leader CTA's copy
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(leader_addr(bar)) : "memory");
}
// TMA load issued by either CTA of the pair; bytes are accounted on the LEADER's mbarrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                                 int32_t col, int32_t row) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_addr(bar)), "r"(col), "r"(row)
        : "memory");
}
// Four rows of a 2-D tensor, picked by index, land as four consecutive 128-byte rows at dst (SWIZZLE_128B is
// applied by shared-memory row as for a tiled box; rows outside the tensor arrive as zeros); issued by either
// CTA of the pair, bytes accounted on the LEADER's mbarrier.  The tensor map's box is {64 columns, 1 row}.
__device__ __forceinline__ void tma_gather4_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int32_t col,
                                                 int32_t r0, int32_t r1, int32_t r2, int32_t r3) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        :
        : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_addr(bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2),
          "r"(r3)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols)
                 : "memory");
}
// D[tmem of both CTAs] (+)= A * B^T with M = 256 split over the CTA pair, issued by the leader
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                              uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 8-bit operands (e4m3): K = 32 per instruction, same 32 bytes of every operand row as kind::f16
__device__ __forceinline__ void umma_f8_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                             uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// One tcgen05.mma from the two 32-bit halves of its shared-memory descriptors: `hi` (stride, version,
// swizzle mode) is the same constant for every operand, `lo` = start address >> 4 | LBO << 16, so an
// operand at a byte offset inside the stage is lo + (offset >> 4) -- one 32-bit add with an immediate
// when the offset is a compile-time constant (the 14-bit address field cannot carry: a stage lies
// inside one CTA's 228 KB of shared memory).
template <bool kPair, bool kF8>
__device__ __forceinline__ void umma_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi,
                                          uint32_t idesc, uint32_t accumulate) {
    if (kPair && kF8) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
            "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
            "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], da, db, %4, p;\n\t}\n"
            ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate) : "memory");
    } else if (kPair) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
            "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}\n"
            ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
            "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}\n"
            ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate) : "memory");
    }
}
// the same MMA under a guard predicate (`enable` is warp-uniform): K-steps that a short last chunk
// does not have are skipped without a branch in the issuing warp's instruction stream
template <bool kF8>
__device__ __forceinline__ void umma_lohi_pair_if(uint32_t enable, uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo,
                                                  uint32_t hi, uint32_t idesc, uint32_t accumulate) {
    if (kF8) {
        asm volatile(
            "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\tsetp.ne.b32 q, %6, 0;\n\t"
            "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
            "@q tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], da, db, %4, p;\n\t}\n"
            ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate), "r"(enable) : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\tsetp.ne.b32 q, %6, 0;\n\t"
            "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
            "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}\n"
            ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate), "r"(enable) : "memory");
    }
}
// arrive on the mbarrier at this offset in BOTH CTAs once the leader's MMAs retired
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(bar), "h"(static_cast<uint16_t>(3))
        : "memory");
}

__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t gets lane (base_lane + t), columns c..c+31
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// 32 columns at taddr plus the 8 halo columns at taddr8, in ONE asm statement: with two statements
// ptxas re-uses the first registers for the second load and copies all 32 results away first
__device__ __forceinline__ void tmem_ld_32x40(uint32_t taddr, uint32_t taddr8, uint32_t (&r)[40]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%40];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%32, %33, %34, %35, %36, %37, %38, %39}, [%41];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
          "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]),
          "=r"(r[38]), "=r"(r[39])
        : "r"(taddr), "r"(taddr8)
        : "memory");
}

// 64 columns at taddr plus the 8 halo columns at taddr8 in one asm statement (the two 32-column chunks
// of an epilogue warp and their halo, loaded together)
__device__ __forceinline__ void tmem_ld_32x72(uint32_t taddr, uint32_t taddr8, uint32_t (&r)[72]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%72];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%64, %65, %66, %67, %68, %69, %70, %71}, [%73];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63]), "=r"(r[64]), "=r"(r[65]), "=r"(r[66]), "=r"(r[67]), "=r"(r[68]), "=r"(r[69]), "=r"(r[70]), "=r"(r[71])
        : "r"(taddr), "r"(taddr8)
        : "memory");
}

// K-major, 128-byte-swizzled shared memory operand descriptor (UMMA "matrix descriptor"):
//   [0,14)  start address >> 4          [16,30) leading byte offset >> 4 (unused for SW128 K-major)
//   [32,46) stride byte offset >> 4 = 1024 B between 8-row groups
//   [46,48) descriptor version = 1      [49,52) base offset      [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t base_offset) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(base_offset & 7) << 49;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

// instruction descriptor for kind::f16: D=f32, A=B=f16, both K-major, M x N
__host__ __device__ constexpr uint32_t umma_idesc_f16(int m, int n) {
    return (1u << 4)                               // c_format = F32
           | (0u << 7) | (0u << 10)                // a_format = b_format = F16
           | (0u << 15) | (0u << 16)               // a_major = b_major = K
           | (static_cast<uint32_t>(n >> 3) << 17) // n_dim
           | (static_cast<uint32_t>(m >> 4) << 24);  // m_dim
}

__device__ __forceinline__ const float* lsh_row_ptr(const LshParams& p, int64_t id) {
    if (id >= 0 && id < p.n_base) return p.table + id * p.dim;
    id -= p.n_base;
    if (id >= 0 && id < p.n_script_extra) return p.script_extra + id * p.dim;
    id -= p.n_script_extra;
    if (id >= 0 && id < p.n_fan_extra) return p.fan_extra + id * p.dim;
    return nullptr;
}

// first CSR row whose end is > t  (off has n_rows+1 entries, off[0] = 0)
__device__ __forceinline__ int32_t csr_row_of(const int64_t* __restrict__ off, int32_t n_rows,
                                              int64_t t) {
    int32_t lo = 0, hi = n_rows;  // invariant: off[lo] <= t < off[hi]
    while (hi - lo > 1) {
        int32_t mid = (lo + hi) >> 1;
        if (__ldg(off + mid) <= t)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

// CSR row of token t for a block that walks consecutive tokens: thread 0 binary-searches the
// row of the block's first token once (shared), every thread then advances linearly -- rows are
// thousands of tokens long, so this replaces ~12 dependent loads per token by ~1.
__device__ __forceinline__ int32_t csr_row_from_hint(const int64_t* __restrict__ off, int32_t n_rows,
                                                     int64_t t, int32_t hint) {
    int32_t row = hint;
    while (row + 1 < n_rows && __ldg(off + row + 1) <= t) ++row;
    return row;
}

#endif  // __CUDACC__

}  // namespace fs
