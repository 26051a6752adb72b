// C ABI of the reuse-search hot path (see include/fandom_search.h).
#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <new>
#include <vector>

#include "common.cuh"

namespace fs {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

constexpr int kTimingRing = 256;
// Slack of the tensor-core pre-filter: fp16 operand rounding is bounded by
// 2*2^-11 * sum|f_k s_k| <= 0.00098 |f||s| (Cauchy-Schwarz); the fp16x2-packed epilogue
// shuffles round each shuffled partial sum once more, <= 2^-10 |f||s| in total, and the all-fp16
// epilogue of E = 6 (pack level 2) <= 2^-9 |f||s| (distance.cu); fp32 accumulation and fp32
// norms add ~1e-5.  Every pair with float64 cos > 1-thr passes
// cos_approx > 1-thr-kEps.
// The operand rounding (fp8 or fp16) is measured per window and enters the pre-filter bound; this
// slack covers what is left: fp32 accumulation order in the tensor core and the fp16x2 epilogue
// sums (<= 2^-9 |f||s|).
constexpr double kEpsAccum = 3.0e-3;
// automatic choice of the pre-filter columns: keep this share of the table's energy (choose_kept_columns)
constexpr double kAutoKeepEnergy = 0.83;
// room for one batch's out-of-vocabulary rows behind the table (fused gather); batches with more fall back
// to the materialised fan matrix
constexpr int64_t kFusedExtraRows = 1 << 16;
// largest scaled row norm of an fp8 table: window token dots, |sum| <= window * norm^2, must still
// fit the fp16 range of the packed epilogue (norm = 95.7 for 6-gram windows)
inline float f8_row_norm(int32_t window) { return std::sqrt(55000.0f / static_cast<float>(window)); }

template <typename T>
static int dev_alloc(T** p, int64_t count) {
    *p = nullptr;
    if (count <= 0) count = 1;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), static_cast<size_t>(count) * sizeof(T));
    if (e != cudaSuccess) {
        set_error("cudaMalloc of %lld bytes failed: %s",
                  static_cast<long long>(count * static_cast<int64_t>(sizeof(T))),
                  cudaGetErrorString(e));
        (void)cudaGetLastError();
        return FS_E_NOMEM;
    }
    return FS_OK;
}

template <typename T>
static int dev_grow(T** p, int64_t* cap, int64_t need) {
    if (need <= *cap && *p) return FS_OK;
    // Growing a buffer synchronises the device (a batch may be in flight): leave headroom from the
    // first allocation on, so that clusters a few percent larger than the first do not stall the
    // pipeline; later growth is geometric.  (Requests of at most 64 Ki elements stay exact: tests
    // size candidate buffers tightly on purpose.)
    int64_t ncap = *cap > 0 ? *cap : (need > (1 << 16) ? need + need / 4 : need);
    while (ncap < need) ncap += ncap / 2 + 1024;
    if (*p) {
        cudaDeviceSynchronize();
        cudaFree(*p);
        *p = nullptr;
    }
    int r = dev_alloc(p, ncap);
    if (r != FS_OK) {
        *cap = 0;
        return r;
    }
    *cap = ncap;
    return FS_OK;
}

}  // namespace fs

using namespace fs;

struct fs_index {
    int device = 0;
    int sm_count = 148;
    int32_t dim = 0, window = 6;
    int32_t dim_pad = 0;        // operand row length in 2-byte units (fp16 elements, or fp8 elements / 2)
    int32_t dim_pad_elems = 0;  // operand row length in elements
    bool diag_user = false;     // FS_OPT_DIAG was set by the caller (else chosen per operand type)
    bool ready = false;         // operand tables built (false after a failed re-conversion)
    float row_limit_sq = 0.f;   // squared norm of the longest scaled row of the index
    int32_t operand_bits = 8;   // 16: fp16 operands, 8: fp8 e4m3 operands (default)
    int32_t prefilter_auto = 1; // 1 (default): the kept columns are chosen by energy share; 0: prefilter_dims (0 = all)
    int32_t prefilter_dims = 0; // FS_OPT_PREFILTER_DIMS: embedding columns kept in the operand rows (0 = auto)
    int32_t kept_dims = 0;      // columns actually kept (<= dim), in the order of `perm`
    int32_t* perm = nullptr;    // [dim] source column of operand element c: columns by descending energy
    double* col_energy = nullptr;  // [dim] sum of squares per column over the table and the script extras
    double kept_energy = 1.0;   // share of the table's energy in the kept columns
    int64_t n_extra_rows = 0;
    double threshold = 0.1;
    float scale = 1.f;
    int64_t n_base = 0, n_sx = 0;

    float* table32 = nullptr;
    // operand rows [table | script extras | room for one batch's fan extras]: ONE allocation, so that the
    // distance kernel can fetch fan rows from it by token id (fused gather); table16, sx16 and fx_rows point
    // into it
    __half* rows16 = nullptr;
    CUtensorMap map_rows;          // box {64 columns, 1 row}: TMA tile::gather4
    int64_t n_table_rows = 0;      // n_base + n_sx + kFusedExtraRows
    __half* fx_rows = nullptr;
    int32_t fused_gather = 1;      // FS_OPT_FUSED_GATHER
    int32_t last_fused = 0;        // the last search fetched its fan rows by fused gather
    __half* table16 = nullptr;
    float4* table_sq = nullptr;   // per row (norm^2, kept rounding error^2, dropped^2, -) of the scaled row
    float* sx32 = nullptr;
    __half* sx16 = nullptr;
    float4* sx_sq = nullptr;

    int32_t* script_tok = nullptr;
    int64_t n_script_tok = 0;
    int64_t* script_off = nullptr;
    int32_t n_scripts = 0;
    int64_t n_script_windows = 0;
    __half* script_emb = nullptr;
    float4* script_tok_sq = nullptr;
    float2* script_norm = nullptr;      // (B_j, half2(D_j, H_j)) per script window start
    float2* script_norm_min = nullptr;  // (min B, half2(max D, max H)) over 32 columns
    int32_t tiles_n = 0;
    CUtensorMap map_script;
    CUtensorMap map_script128;  // boxes of 128 rows (E = 6: no halo rows)
    CUtensorMap map_script64;   // boxes of 64 rows (128-column tiles, distance_kernel_n128)

    unsigned long long* hash_table = nullptr;
    uint32_t* hash_filter = nullptr;  // one bit per value of the top bits of the key hash (probe pre-test)
    uint32_t hash_filter_bits = 0;
    uint32_t hash_slots = 0;

    // per-batch workspace
    int64_t tok_cap = 0;  // rows of fan_emb
    int64_t emb_cap = 0;  // elements of fan_emb
    __half* fan_emb = nullptr;
    int64_t sq_cap = 0;
    float4* fan_tok_sq = nullptr;
    int64_t thr_cap = 0;
    float2* fan_thr = nullptr;  // (A_i, half2(C_i, G_i)) per fan window start
    int64_t cand_cap = 0;
    fs_pair* cand = nullptr;
    int64_t fx_cap = 0;  // fan extra rows
    __half* fx16 = nullptr;
    int64_t fxsq_cap = 0;
    float4* fx_sq = nullptr;

    // staging of the _host entry points
    cudaStream_t stream = nullptr;
    int64_t h_tok_cap = 0;
    int32_t* h_tok = nullptr;
    int64_t h_off_cap = 0;
    int64_t* h_off = nullptr;
    int64_t h_extra_cap = 0;
    float* h_extra = nullptr;
    int64_t h_out_cap = 0;
    fs_match* h_out = nullptr;
    int64_t h_pair_cap = 0;
    fs_pair* h_pair = nullptr;
    unsigned long long* h_counters = nullptr;

    // fs_search_submit / fs_search_collect: kSlots batches in flight.  A slot owns the device copies of
    // its inputs and its outputs; the workspace above is shared (the kernels of consecutive batches
    // are ordered on `stream`).  Inputs travel on `stream_in`, match lists on `stream_out`, so the H2D
    // of batch k+1 and the D2H of batch k-1 run under the distance kernel of batch k.
    struct Slot {
        bool busy = false;
        int64_t tok_cap = 0, off_cap = 0, extra_cap = 0, out_cap = 0;
        int32_t* d_tok = nullptr;
        int64_t* d_off = nullptr;
        float* d_extra = nullptr;
        fs_match* d_out = nullptr;
        unsigned long long* d_counters = nullptr;
        long long* h_counters = nullptr;  // page-locked
        int64_t cap = 0;
        // device-side post-processing (fs_search_submit_rows): verbatim token texts and the rows
        bool with_rows = false;
        int64_t text_cap = 0, tstart_cap = 0, tlen_cap = 0, rows_buf_cap = 0, rows_cap = 0;
        uint8_t* d_text = nullptr;
        uint32_t* d_tstart = nullptr;
        uint16_t* d_tlen = nullptr;
        fs_row* d_rows = nullptr;
        cudaEvent_t ev_in = nullptr, ev_done = nullptr;
    };
    static constexpr int kSlots = 2;
    Slot slots[kSlots];
    cudaStream_t stream_in = nullptr, stream_out = nullptr;

    // script words for the device-side Levenshtein, and the post-processing workspace (shared by the
    // slots: the kernels of consecutive batches are ordered on `stream`)
    uint8_t* script_text = nullptr;
    int64_t* script_word_off = nullptr;
    int64_t n_script_words = 0;
    int64_t pp_tok_cap = 0, pp_match_cap = 0, pp_blocks_cap = 0;
    int32_t *pp_head = nullptr, *pp_winner = nullptr, *pp_next = nullptr, *pp_lev = nullptr, *pp_rank = nullptr,
            *pp_block_count = nullptr;
    unsigned long long *pp_best_key = nullptr, *pp_best_tie = nullptr;
    int64_t* pp_block_off = nullptr;

    // per-script-word reuse histogram accumulated from the device rows (caller-owned device buffers)
    unsigned long long* hist_counts = nullptr;
    const double* hist_thresholds = nullptr;
    int32_t hist_n_thr = 0;

    // LSH emulation (parity mode)
    double* lsh_normals = nullptr;
    int32_t lsh_tables = 0, lsh_bits = 0;

    // options
    int32_t diag = 1;              // diagonal-sum factor E of the distance kernel
    int32_t pair = 0;              // CTA-pair (cta_group::2) kernel
    int32_t ares = 1;              // A-resident variant of the pair kernel (used when the row fits: <= 640 B)
    int32_t pack = 2;              // epilogue diagonal sums: 0 fp32 shuffles, 1 fp16x2 shuffles, 2 fp16x2 arithmetic
    int32_t shifts_per_stage = 0;  // 0 = all MMA shifts of a chunk in one stage
    int32_t tile_group = 103;      // FS_OPT_TILE_GROUP: bit 0 grouped stages, bit 1 early TMEM release, bit 2 one-pass epilogue,
                                   // bit 5 second-level rejection, bit 6 128-column tiles with two CTA pairs per TPC (rows of
                                   // <= 256 bytes); bit 4 prefetched bounds and bit 3 alternating epilogue warp sets are off
                                   // (measured slower, profiles/r02_sweep_epilogue_variants.jsonl)
    int32_t grid_limit = 0;

    // timing ring
    cudaEvent_t ev_start[kTimingRing];
    cudaEvent_t ev_stop[kTimingRing];
    bool ev_created = false;
    int64_t ev_count = 0;
};

static int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

extern "C" {

int fs_abi_version(void) { return FS_ABI_VERSION; }

const char* fs_last_error(void) { return g_err; }

int fs_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

int fs_index_destroy(fs_index* idx) {
    if (!idx) return FS_OK;
    cudaSetDevice(idx->device);
    cudaDeviceSynchronize();
    void* ptrs[] = {idx->table32, idx->rows16,    idx->table_sq,   idx->sx32,       nullptr,
                    idx->sx_sq,   idx->script_tok, idx->script_off, idx->script_emb, idx->script_tok_sq,
                    idx->script_norm, idx->hash_table, idx->fan_emb, idx->fan_tok_sq, idx->fan_thr,
                    idx->cand,    idx->fx16,      idx->fx_sq,      idx->h_tok,      idx->h_off,
                    idx->h_extra, idx->h_out,     idx->h_pair,     idx->h_counters,
                    idx->lsh_normals, idx->script_norm_min, idx->perm, idx->col_energy,
                    idx->script_text, idx->script_word_off, idx->pp_head, idx->pp_winner, idx->pp_next,
                    idx->pp_lev, idx->pp_rank, idx->pp_block_count, idx->pp_best_key, idx->pp_best_tie,
                    idx->pp_block_off, idx->hash_filter};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    if (idx->ev_created) {
        for (int i = 0; i < kTimingRing; ++i) {
            cudaEventDestroy(idx->ev_start[i]);
            cudaEventDestroy(idx->ev_stop[i]);
        }
    }
    for (auto& sl : idx->slots) {
        void* q[] = {sl.d_tok, sl.d_off, sl.d_extra, sl.d_out, sl.d_counters, sl.d_text, sl.d_tstart, sl.d_tlen, sl.d_rows};
        for (void* p : q)
            if (p) cudaFree(p);
        if (sl.h_counters) cudaFreeHost(sl.h_counters);
        if (sl.ev_in) cudaEventDestroy(sl.ev_in);
        if (sl.ev_done) cudaEventDestroy(sl.ev_done);
    }
    if (idx->stream) cudaStreamDestroy(idx->stream);
    if (idx->stream_in) cudaStreamDestroy(idx->stream_in);
    if (idx->stream_out) cudaStreamDestroy(idx->stream_out);
    delete idx;
    return FS_OK;
}

// Which embedding columns the operand rows keep (FS_OPT_PREFILTER_DIMS).  The columns are ordered by
// their energy over the index's table and script extras (a permutation: every dot product is
// unchanged), the first `kept_dims` enter the operand rows, and what a window holds in the dropped
// ones is bounded per pair by |f_drop| |s_drop| in the pre-filter threshold (window_norm_kernel), so
// the candidates stay a guaranteed superset whatever is dropped: fewer K-steps per tile for a wider
// threshold.  0 = automatic (see kAutoKeepEnergy).
static int choose_kept_columns(fs_index* idx) {
    cudaStream_t st = idx->stream;
    const int32_t dim = idx->dim;
    int r;
    std::vector<double> energy(static_cast<size_t>(dim), 0.0);
    if (!idx->perm) {
        if ((r = dev_alloc(&idx->perm, dim)) != FS_OK) return r;
        if ((r = dev_alloc(&idx->col_energy, dim)) != FS_OK) return r;
    }
    FS_CUDA_CHECK(cudaMemsetAsync(idx->col_energy, 0, sizeof(double) * dim, st));
    if ((r = launch_column_energy(idx->table32, idx->n_base, dim, idx->col_energy, st)) != FS_OK) return r;
    if ((r = launch_column_energy(idx->sx32, idx->n_sx, dim, idx->col_energy, st)) != FS_OK) return r;
    FS_CUDA_CHECK(cudaMemcpyAsync(energy.data(), idx->col_energy, sizeof(double) * dim, cudaMemcpyDeviceToHost, st));
    FS_CUDA_CHECK(cudaStreamSynchronize(st));
    std::vector<int32_t> perm(static_cast<size_t>(dim));
    for (int32_t c = 0; c < dim; ++c) perm[c] = c;
    std::stable_sort(perm.begin(), perm.end(), [&](int32_t a, int32_t b) { return energy[a] > energy[b]; });
    double total = 0.0;
    for (double e : energy) total += e;
    const int32_t kstep = idx->operand_bits == 8 ? 2 * kUmmaK : kUmmaK;
    int32_t kept = dim;
    if (idx->prefilter_dims > 0) {
        kept = idx->prefilter_dims < dim ? idx->prefilter_dims : dim;
    } else if (idx->prefilter_auto && total > 0.0 && dim > 4 * kstep) {
        // automatic: whole 128-byte chunks of operand row (the unit the stage ring and the L2 -> SM
        // stream move), as few as keep kAutoKeepEnergy of the energy; never more than the embedding has
        const int32_t chunk = idx->operand_bits == 8 ? 2 * kChunkK : kChunkK;
        double run = 0.0;
        int32_t need = dim;
        for (int32_t c = 0; c < dim; ++c) {
            run += energy[perm[c]];
            if (run >= kAutoKeepEnergy * total) {
                need = c + 1;
                break;
            }
        }
        kept = static_cast<int32_t>(round_up(need, chunk));
        if (kept > dim) kept = dim;
    }
    if (kept < 1) kept = 1;
    if (kept >= dim)  // nothing dropped: keep the table's own column order
        for (int32_t c = 0; c < dim; ++c) perm[c] = c;
    idx->kept_dims = kept;
    double kept_e = 0.0;
    for (int32_t c = 0; c < kept; ++c) kept_e += energy[perm[c]];
    idx->kept_energy = total > 0.0 ? kept_e / total : 1.0;
    FS_CUDA_CHECK(cudaMemcpyAsync(idx->perm, perm.data(), sizeof(int32_t) * dim, cudaMemcpyHostToDevice, st));
    FS_CUDA_CHECK(cudaStreamSynchronize(st));
    return FS_OK;
}

// (Re)builds everything that depends on the operand type: the converted table and script
// extras, the script token matrix, its window norms and tensor map.  fp16: one global scale
// 1/max|x|.  fp8 e4m3: scale f8_row_norm/max|row| (so that the window's token dots still fit the
// fp16 range of the packed epilogue).  In both cases the rounding error of every row is MEASURED
// (including underflow of rows that are tiny next to the largest) and enters the pre-filter bound
// per window (window_norm_kernel), which keeps the candidate set a guaranteed superset.
static int prepare_operands(fs_index* idx) {
    cudaStream_t st = idx->stream;
    const bool f8 = idx->operand_bits == 8;
    idx->ready = false;
    FS_CUDA_CHECK(cudaSetDevice(idx->device));
    FS_CUDA_CHECK(cudaDeviceSynchronize());
    int r;
    if ((r = choose_kept_columns(idx)) != FS_OK) return r;
    idx->dim_pad_elems = static_cast<int32_t>(round_up(idx->kept_dims, f8 ? 2 * kUmmaK : kUmmaK));  // K of one tcgen05.mma
    idx->dim_pad = f8 ? idx->dim_pad_elems / 2 : idx->dim_pad_elems;
    if (!idx->diag_user) {
        // Default diagonal factor (all variants are parity-tested and selectable).  E = 6 runs one MMA
        // shift per K-step (a sixth of the dense tensor work) and leaves the rest to the epilogue; since
        // the epilogue rejects almost every chunk from its row maxima before summing a diagonal, that
        // is the faster split for fp8 operands at every width (C2 workload, M windows/s, E=6 vs E=3:
        // 86.5 vs 82.2 at d=300, 72.7 vs 55.4 at d=512, 56.4 vs 35.9 at d=768).  fp16 operands keep
        // E = 3 below 416 elements (their MMAs cost twice as much; measured before the rejection).
        const int32_t wide = f8 ? 0 : 416;
        if (idx->window % 6 == 0 && idx->dim_pad_elems >= wide)
            idx->diag = 6;
        else
            idx->diag = idx->window % 3 == 0 ? 3 : (idx->window % 2 == 0 ? 2 : 1);
        idx->shifts_per_stage = 0;
    }
    void* stale[] = {idx->rows16, idx->script_emb, idx->fan_emb, idx->fx16};
    for (void* q : stale)
        if (q) cudaFree(q);
    idx->rows16 = idx->table16 = idx->sx16 = idx->fx_rows = idx->script_emb = idx->fan_emb = idx->fx16 = nullptr;
    idx->emb_cap = idx->tok_cap = idx->fx_cap = 0;
    idx->n_table_rows = idx->n_base + idx->n_sx + kFusedExtraRows;
    if ((r = dev_alloc(&idx->rows16, idx->n_table_rows * idx->dim_pad)) != FS_OK) return r;
    idx->table16 = idx->rows16;
    idx->sx16 = idx->table16 + idx->n_base * idx->dim_pad;
    idx->fx_rows = idx->sx16 + idx->n_sx * idx->dim_pad;
    if ((r = make_token_map(&idx->map_rows, idx->rows16, idx->n_table_rows, idx->dim_pad, 1)) != FS_OK) return r;
    if ((r = dev_alloc(&idx->script_emb, idx->n_script_tok * idx->dim_pad)) != FS_OK) return r;
    unsigned int* d_max = reinterpret_cast<unsigned int*>(idx->h_counters);
    FS_CUDA_CHECK(cudaMemsetAsync(d_max, 0, sizeof(unsigned long long) * FS_CNT_COUNT, st));
    // slot 0: largest squared row norm; slot 1: largest |element| (fp16 scale)
    if ((r = launch_rownorm_max(idx->table32, idx->n_base, idx->dim, d_max, st)) != FS_OK) return r;
    if ((r = launch_rownorm_max(idx->sx32, idx->n_sx, idx->dim, d_max, st)) != FS_OK) return r;
    if (!f8) {
        if ((r = launch_absmax(idx->table32, idx->n_base * idx->dim, d_max + 1, st)) != FS_OK) return r;
        if ((r = launch_absmax(idx->sx32, idx->n_sx * idx->dim, d_max + 1, st)) != FS_OK) return r;
    }
    unsigned int h_max_bits[2] = {0, 0};
    FS_CUDA_CHECK(cudaMemcpyAsync(h_max_bits, d_max, sizeof(h_max_bits), cudaMemcpyDeviceToHost, st));
    FS_CUDA_CHECK(cudaStreamSynchronize(st));
    float h_norm_sq, h_abs;
    memcpy(&h_norm_sq, &h_max_bits[0], sizeof(float));
    memcpy(&h_abs, &h_max_bits[1], sizeof(float));
    idx->scale = 1.0f;
    if (f8 && h_norm_sq > 0.f && std::isfinite(h_norm_sq))
        idx->scale = f8_row_norm(idx->window) / std::sqrt(h_norm_sq);
    else if (!f8 && h_abs > 0.f && std::isfinite(h_abs))
        idx->scale = 1.0f / h_abs;
    // no per-batch row may be longer than the longest row of the index (see convert_rows_kernel)
    idx->row_limit_sq = (h_norm_sq > 0.f && std::isfinite(h_norm_sq))
                            ? h_norm_sq * idx->scale * idx->scale * 1.0001f
                            : 0.f;
    if ((r = launch_convert_rows(idx->table32, idx->n_base, idx->dim, idx->dim_pad, idx->kept_dims, idx->perm,
                                 idx->scale, f8, 0.f, idx->table16, idx->table_sq, st)) != FS_OK)
        return r;
    if ((r = launch_convert_rows(idx->sx32, idx->n_sx, idx->dim, idx->dim_pad, idx->kept_dims, idx->perm,
                                 idx->scale, f8, 0.f, idx->sx16, idx->sx_sq, st)) != FS_OK)
        return r;
    const int64_t n_pad = static_cast<int64_t>(idx->tiles_n) * kBlockN;
    GatherSources src{idx->table16, idx->table_sq, idx->n_base, idx->sx16, idx->sx_sq,
                      idx->n_sx,    nullptr,       nullptr,     0};
    if ((r = launch_gather(idx->script_tok, idx->n_script_tok, src, idx->dim_pad, idx->script_emb,
                           idx->script_tok_sq, idx->sm_count, st)) != FS_OK)
        return r;
    FS_CUDA_CHECK(cudaMemsetAsync(idx->script_tok_sq + idx->n_script_tok, 0, sizeof(float4) * 8, st));
    unsigned long long* d_cnt = idx->h_counters;
    FS_CUDA_CHECK(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * FS_CNT_COUNT, st));
    const float coef = static_cast<float>(1.0 - idx->threshold - kEpsAccum);
    if ((r = launch_window_norm(idx->script_tok_sq, idx->n_script_tok, idx->script_off, idx->n_scripts,
                                idx->window, coef, true, idx->script_norm, nullptr, n_pad,
                                d_cnt + FS_CNT_WINDOWS, st)) != FS_OK)
        return r;
    if ((r = launch_sliding_minmax32(idx->script_norm, idx->script_norm_min, n_pad, st)) != FS_OK) return r;
    unsigned long long h_cnt[FS_CNT_COUNT];
    FS_CUDA_CHECK(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
    if (idx->n_script_tok > 0)
    {
        if ((r = make_token_map(&idx->map_script, idx->script_emb, idx->n_script_tok, idx->dim_pad, kBoxRows)) != FS_OK)
            return r;
        if ((r = make_token_map(&idx->map_script128, idx->script_emb, idx->n_script_tok, idx->dim_pad, 128)) != FS_OK)
            return r;
        if ((r = make_token_map(&idx->map_script64, idx->script_emb, idx->n_script_tok, idx->dim_pad, 64)) != FS_OK)
            return r;
    }
    FS_CUDA_CHECK(cudaStreamSynchronize(st));
    idx->n_script_windows = static_cast<int64_t>(h_cnt[FS_CNT_WINDOWS]);
    idx->ready = true;
    return FS_OK;
}

int fs_index_create(fs_index** out, int device, const float* table, int64_t n_rows, int32_t dim,
                    const float* extra, int64_t n_extra, const int32_t* script_tok,
                    int64_t n_script_tok, const int64_t* script_off, int64_t n_scripts,
                    int32_t window, double threshold) {
    if (!out || !table || n_rows < 0 || dim <= 0 || n_extra < 0 || (n_extra > 0 && !extra) ||
        n_script_tok < 0 || (n_script_tok > 0 && !script_tok) || !script_off || n_scripts < 1 ||
        window < 1 || window > 8 || !(threshold > 0.0) || !(threshold < 1.0)) {
        set_error("fs_index_create: invalid argument");
        return FS_E_INVALID;
    }
    if (n_script_tok >= (1ll << 31) - 1024 || n_rows + n_extra >= (1ll << 31) - 1) {
        set_error("fs_index_create: sizes exceed int32 positions");
        return FS_E_INVALID;
    }
    if (script_off[0] != 0 || script_off[n_scripts] != n_script_tok) {
        set_error("fs_index_create: script_off must start at 0 and end at n_script_tok");
        return FS_E_INVALID;
    }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        (void)cudaGetLastError();
        set_error("fs_index_create: CUDA device %d not available (%d visible); this path has no CPU fallback",
                  device, ndev);
        return FS_E_NODEVICE;
    }
    FS_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    FS_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("fs_index_create: device %d is sm_%d%d; these kernels are built for sm_100a only",
                  device, prop.major, prop.minor);
        return FS_E_NODEVICE;
    }
    fs_index* idx = new (std::nothrow) fs_index();
    if (!idx) return FS_E_NOMEM;
    idx->device = device;
    idx->sm_count = prop.multiProcessorCount;
    idx->dim = dim;
    idx->dim_pad = static_cast<int32_t>(round_up(dim, kUmmaK));  // K granularity of tcgen05.mma.kind::f16
    idx->window = window;
    idx->threshold = threshold;
    idx->n_base = n_rows;
    idx->n_sx = n_extra;
    idx->n_script_tok = n_script_tok;
    idx->n_scripts = static_cast<int32_t>(n_scripts);
    idx->pair = 1;

#define FS_TRY(expr)                    \
    do {                                \
        int _r = (expr);                \
        if (_r != FS_OK) {              \
            fs_index_destroy(idx);      \
            return _r;                  \
        }                               \
    } while (0)
#define FS_TRY_CUDA(expr)                                                                   \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,     \
                      __LINE__);                                                            \
            fs_index_destroy(idx);                                                          \
            return FS_E_CUDA;                                                               \
        }                                                                                   \
    } while (0)

    FS_TRY_CUDA(cudaStreamCreateWithFlags(&idx->stream, cudaStreamNonBlocking));
    FS_TRY_CUDA(cudaStreamCreateWithFlags(&idx->stream_in, cudaStreamNonBlocking));
    FS_TRY_CUDA(cudaStreamCreateWithFlags(&idx->stream_out, cudaStreamNonBlocking));
    for (auto& sl : idx->slots) {
        FS_TRY_CUDA(cudaEventCreateWithFlags(&sl.ev_in, cudaEventDisableTiming));
        // blocking sync: the thread waiting in fs_search_collect sleeps instead of spinning -- on a
        // node with 8 ranks and 32 cores a spinning waiter per rank is a quarter of the rank's cores
        FS_TRY_CUDA(cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming | cudaEventBlockingSync));
        FS_TRY(dev_alloc(&sl.d_counters, FS_CNT_COUNT));
        FS_TRY_CUDA(cudaMallocHost(reinterpret_cast<void**>(&sl.h_counters), sizeof(long long) * FS_CNT_COUNT));
    }
    cudaStream_t st = idx->stream;
    for (int i = 0; i < kTimingRing; ++i) {
        FS_TRY_CUDA(cudaEventCreate(&idx->ev_start[i]));
        FS_TRY_CUDA(cudaEventCreate(&idx->ev_stop[i]));
    }
    idx->ev_created = true;

    FS_TRY(dev_alloc(&idx->h_counters, FS_CNT_COUNT));
    FS_TRY(dev_alloc(&idx->table32, n_rows * dim));
    FS_TRY(dev_alloc(&idx->sx32, n_extra * dim));
    FS_TRY(dev_alloc(&idx->table_sq, n_rows));
    FS_TRY(dev_alloc(&idx->sx_sq, n_extra));
    if (n_rows)
        FS_TRY_CUDA(cudaMemcpyAsync(idx->table32, table, sizeof(float) * n_rows * dim,
                                    cudaMemcpyHostToDevice, st));
    if (n_extra)
        FS_TRY_CUDA(cudaMemcpyAsync(idx->sx32, extra, sizeof(float) * n_extra * dim,
                                    cudaMemcpyHostToDevice, st));
    // script side: tokens, CSR; padded so that any tile stepping (256 - (E-1) columns) stays in bounds
    idx->tiles_n = static_cast<int32_t>((n_script_tok + kBlockN - 8 - 1) / (kBlockN - 8)) + 1;
    const int64_t n_pad = static_cast<int64_t>(idx->tiles_n) * kBlockN;
    FS_TRY(dev_alloc(&idx->script_tok, n_script_tok + 8));
    FS_TRY(dev_alloc(&idx->script_off, n_scripts + 1));
    FS_TRY(dev_alloc(&idx->script_tok_sq, n_script_tok + 8));
    FS_TRY(dev_alloc(&idx->script_norm, n_pad));
    FS_TRY(dev_alloc(&idx->script_norm_min, n_pad));
    FS_TRY_CUDA(cudaMemsetAsync(idx->script_tok, 0xFF, sizeof(int32_t) * (n_script_tok + 8), st));
    if (n_script_tok)
        FS_TRY_CUDA(cudaMemcpyAsync(idx->script_tok, script_tok, sizeof(int32_t) * n_script_tok,
                                    cudaMemcpyHostToDevice, st));
    FS_TRY_CUDA(cudaMemcpyAsync(idx->script_off, script_off, sizeof(int64_t) * (n_scripts + 1),
                                cudaMemcpyHostToDevice, st));
    // operand rows (table, script extras, script token matrix), window norms, tensor map
    FS_TRY(prepare_operands(idx));

    // load factor <= 1/16 up to 2^22 slots (32 MB of table, L2 resident), never above 1/2
    uint32_t slots = 1024;
    while (slots < 2 * static_cast<uint64_t>(n_script_tok) ||
           (slots < 16 * static_cast<uint64_t>(n_script_tok) && slots < (1u << 22)))
        slots <<= 1;
    idx->hash_slots = slots;
    FS_TRY(dev_alloc(&idx->hash_table, slots));
    idx->hash_filter_bits = hash_filter_bits(n_script_tok);
    FS_TRY(dev_alloc(&idx->hash_filter, idx->hash_filter_bits / 32));
    FS_TRY(launch_hash_build(idx->script_tok, n_script_tok, idx->script_off, idx->n_scripts, window,
                             idx->hash_table, slots, idx->hash_filter, idx->hash_filter_bits, st));
    FS_TRY_CUDA(cudaStreamSynchronize(st));
#undef FS_TRY
#undef FS_TRY_CUDA
    *out = idx;
    return FS_OK;
}

// the configuration distance_kernel_n128 runs (fs_index_get_info 15) with the fused gather switched on
static bool n128_config(const fs_index* idx) {
    return idx->pair && idx->ares && idx->diag == 6 && idx->pack == 2 && (idx->tile_group & 69) == 69 &&
           (idx->dim_pad + kChunkK - 1) / kChunkK <= 2 && idx->window == 6;
}
static bool fused_config(const fs_index* idx) { return idx->fused_gather && n128_config(idx); }

static int grow_fan_matrix(fs_index* idx, int64_t max_tokens) {
    int r;
    if ((r = dev_grow(&idx->fan_emb, &idx->emb_cap, max_tokens * idx->dim_pad)) != FS_OK) return r;
    idx->tok_cap = idx->emb_cap / idx->dim_pad;
    return FS_OK;
}

int fs_index_reserve(fs_index* idx, int64_t max_tokens, int64_t max_candidates) {
    if (!idx || max_tokens < 0 || max_candidates < 0) {
        set_error("fs_index_reserve: invalid argument");
        return FS_E_INVALID;
    }
    FS_CUDA_CHECK(cudaSetDevice(idx->device));
    int r;
    // (the fan operand matrix is only needed when the fused gather does not apply: grown on demand there)
    if (!fused_config(idx) && (r = grow_fan_matrix(idx, max_tokens)) != FS_OK) return r;
    if ((r = dev_grow(&idx->fan_tok_sq, &idx->sq_cap, max_tokens + 8)) != FS_OK) return r;
    // per-window bounds, padded to whole tiles of the SMALLEST row step any variant uses (E = 6: 108)
    if ((r = dev_grow(&idx->fan_thr, &idx->thr_cap, (max_tokens / dist_m_step(6) + 3) * kBlockM)) != FS_OK)
        return r;
    if ((r = dev_grow(&idx->cand, &idx->cand_cap, max_candidates)) != FS_OK) return r;
    return FS_OK;
}

int fs_index_set_option(fs_index* idx, int32_t option, int64_t value) {
    if (!idx) return FS_E_INVALID;
    switch (option) {
        case FS_OPT_SHIFTS_PER_STAGE:
            if (value < 0 || value > 8 || (value > 0 && (idx->window / idx->diag) % value != 0)) {
                set_error("shifts per stage must divide window/diag and be <= 8 (0 = all)");
                return FS_E_INVALID;
            }
            idx->shifts_per_stage = static_cast<int32_t>(value);
            return FS_OK;
        case FS_OPT_PACKED_SHUFFLE:
            idx->pack = value < 0 ? 0 : (value > 2 ? 2 : static_cast<int32_t>(value));
            return FS_OK;
        case FS_OPT_A_RESIDENT:
            idx->ares = value ? 1 : 0;
            return FS_OK;
        case FS_OPT_CTA_PAIR:
            idx->pair = value ? 1 : 0;
            return FS_OK;
        case FS_OPT_DIAG:
            if (!(value == 1 || value == 2 || value == 3 || value == 6) || idx->window % value != 0) {
                set_error("diagonal factor must be 1, 2, 3 or 6 and divide the window");
                return FS_E_INVALID;
            }
            idx->diag = static_cast<int32_t>(value);
            idx->diag_user = true;
            idx->shifts_per_stage = 0;
            return FS_OK;
        case FS_OPT_OPERAND_BITS: {
            if (value != 8 && value != 16) {
                set_error("operand bits must be 16 (fp16) or 8 (fp8 e4m3)");
                return FS_E_INVALID;
            }
            if (idx->operand_bits == value) return FS_OK;
            idx->operand_bits = static_cast<int32_t>(value);
            return prepare_operands(idx);
        }
        case FS_OPT_PREFILTER_DIMS: {
            if (value < -1 || value > idx->dim) {
                set_error("pre-filter dimensions must be -1 (automatic), 0 (all) or 1..dim");
                return FS_E_INVALID;
            }
            const int32_t want_auto = value == -1 ? 1 : 0;
            const int32_t want_dims = value > 0 ? static_cast<int32_t>(value) : 0;
            if (want_auto == idx->prefilter_auto && want_dims == idx->prefilter_dims) return FS_OK;
            idx->prefilter_auto = want_auto;
            idx->prefilter_dims = want_dims;
            return prepare_operands(idx);
        }
        case FS_OPT_FUSED_GATHER:
            idx->fused_gather = value != 0 ? 1 : 0;
            return FS_OK;
        case FS_OPT_GRID_LIMIT:
            idx->grid_limit = static_cast<int32_t>(value < 0 ? 0 : value);
            return FS_OK;
        case FS_OPT_TILE_GROUP:
            idx->tile_group = static_cast<int32_t>(value & 511);  // (bits 7, 8: floor probes of a -DFS_FLOOR_PROBE build)
            return FS_OK;
        default:
            set_error("unknown option %d", option);
            return FS_E_INVALID;
    }
}

int fs_index_set_lsh(fs_index* idx, const double* normals, int32_t n_tables, int32_t n_bits) {
    if (!idx || n_tables < 0 || n_tables > 255 || (n_tables > 0 && (!normals || n_bits < 1 || n_bits > 64))) {
        set_error("fs_index_set_lsh: invalid argument");
        return FS_E_INVALID;
    }
    FS_CUDA_CHECK(cudaSetDevice(idx->device));
    FS_CUDA_CHECK(cudaDeviceSynchronize());
    if (idx->lsh_normals) cudaFree(idx->lsh_normals);
    idx->lsh_normals = nullptr;
    idx->lsh_tables = idx->lsh_bits = 0;
    if (n_tables == 0) return FS_OK;
    const int64_t n = static_cast<int64_t>(n_tables) * n_bits * idx->window * idx->dim;
    int r = dev_alloc(&idx->lsh_normals, n);
    if (r != FS_OK) return r;
    FS_CUDA_CHECK(cudaMemcpy(idx->lsh_normals, normals, sizeof(double) * n, cudaMemcpyHostToDevice));
    idx->lsh_tables = n_tables;
    idx->lsh_bits = n_bits;
    return FS_OK;
}

int64_t fs_index_get_info(const fs_index* idx, int32_t what) {
    if (!idx) return -1;
    switch (what) {
        case 0: return idx->n_script_windows;
        case 1: return idx->dim_pad_elems;
        case 2: return idx->sm_count;
        case 3: return idx->cand_cap;
        case 4: return idx->shifts_per_stage;
        case 5: return idx->diag;
        case 6: return idx->pair;
        case 7: return idx->ares;
        case 8: return idx->pack;
        case 11: return idx->operand_bits;
        case 12: return idx->tile_group;
        case 13: return idx->kept_dims;
        case 14: return static_cast<int64_t>(idx->kept_energy * 1e6);  // share of the energy kept, ppm
        case 15:  // 1: the 128-column kernel (distance_kernel_n128) runs this configuration
            return n128_config(idx) ? 1 : 0;
        case 16:  // 1: the last search fetched its fan rows by fused gather (no fan operand matrix)
            return idx->last_fused;
        default: return -1;
    }
}

float fs_index_scale(const fs_index* idx) { return idx ? idx->scale : 0.f; }

int fs_timing_reset(fs_index* idx) {
    if (!idx) return FS_E_INVALID;
    idx->ev_count = 0;
    return FS_OK;
}

int fs_timing_read(fs_index* idx, double* total_ms, int64_t* launches) {
    if (!idx || !total_ms || !launches) return FS_E_INVALID;
    FS_CUDA_CHECK(cudaSetDevice(idx->device));
    const int64_t n = idx->ev_count < kTimingRing ? idx->ev_count : kTimingRing;
    double tot = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        FS_CUDA_CHECK(cudaEventSynchronize(idx->ev_stop[i]));
        float ms = 0.f;
        FS_CUDA_CHECK(cudaEventElapsedTime(&ms, idx->ev_start[i], idx->ev_stop[i]));
        tot += ms;
    }
    *total_ms = tot;
    *launches = n;
    return FS_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// shared pipeline
// ---------------------------------------------------------------------------
namespace {

enum class Mode { kSearch, kCandidates, kDots, kEmbed };

struct BatchArgs {
    const int32_t* tok;
    int64_t n_tok;
    const int64_t* off;
    int64_t n_works;
    const float* extra;
    int64_t n_extra;
};

int check_batch(const fs_index* idx, const BatchArgs& a, const char* who) {
    if (idx && !idx->ready) {
        set_error("%s: the index is unusable (a re-conversion of its operands failed)", who);
        return FS_E_INVALID;
    }
    if (!idx || a.n_tok < 0 || a.n_works < 0 || (a.n_tok > 0 && !a.tok) || !a.off || a.n_extra < 0 ||
        (a.n_extra > 0 && !a.extra)) {
        set_error("%s: invalid argument", who);
        return FS_E_INVALID;
    }
    if (a.n_tok >= (1ll << 31) - 1024 || a.n_works >= (1ll << 31) - 1) {
        set_error("%s: batch exceeds int32 positions; split it", who);
        return FS_E_INVALID;
    }
    return FS_OK;
}

// gather + window thresholds of one batch into the index workspace
// emb_out == nullptr: fused gather -- the batch's extra rows are converted into the table's tail (a.n_extra <=
// kFusedExtraRows) and only the per-token squares are gathered; the distance kernel fetches the rows itself
int embed_batch(fs_index* idx, cudaStream_t st, const BatchArgs& a, unsigned long long* counters,
                __half* emb_out, float2* thr_out, float4* thr_plain, int64_t thr_pad) {
    int r;
    const bool fused = emb_out == nullptr;
    __half* fx_dst = idx->fx_rows;
    if (a.n_extra > 0) {
        if (!fused) {
            if ((r = dev_grow(&idx->fx16, &idx->fx_cap, a.n_extra * idx->dim_pad)) != FS_OK) return r;
            fx_dst = idx->fx16;
        }
        if ((r = dev_grow(&idx->fx_sq, &idx->fxsq_cap, a.n_extra)) != FS_OK) return r;
        if ((r = launch_convert_rows(a.extra, a.n_extra, idx->dim, idx->dim_pad, idx->kept_dims, idx->perm,
                                     idx->scale, idx->operand_bits == 8, idx->row_limit_sq, fx_dst,
                                     idx->fx_sq, st)) != FS_OK)
            return r;
    }
    GatherSources src{idx->table16, idx->table_sq, idx->n_base, idx->sx16,  idx->sx_sq,
                      idx->n_sx,    fx_dst,        idx->fx_sq,  a.n_extra};
    if (fused)
        r = launch_gather_sq(a.tok, a.n_tok, src, idx->fan_tok_sq, idx->sm_count, st);
    else
        r = launch_gather(a.tok, a.n_tok, src, idx->dim_pad, emb_out, idx->fan_tok_sq, idx->sm_count, st);
    if (r != FS_OK) return r;
    // zero the halo so the window sums never read stale squares
    FS_CUDA_CHECK(cudaMemsetAsync(idx->fan_tok_sq + a.n_tok, 0, sizeof(float4) * 8, st));
    // fan side of the pre-filter bound: (|f|, |f - qf|) per window (window_norm_kernel)
    return launch_window_norm(idx->fan_tok_sq, a.n_tok, a.off, static_cast<int32_t>(a.n_works),
                              idx->window, 0.0f, false, thr_out, thr_plain, thr_pad,
                              counters ? counters + FS_CNT_WINDOWS : nullptr, st);
}

int run_pipeline(fs_index* idx, cudaStream_t st, const BatchArgs& a, Mode mode, fs_match* out,
                 int64_t cap, fs_pair* cand_out, int64_t cand_out_cap, float* dots, int64_t dots_ld,
                 unsigned long long* counters) {
    int r;
    FS_CUDA_CHECK(cudaSetDevice(idx->device));
    if (counters) FS_CUDA_CHECK(cudaMemsetAsync(counters, 0, sizeof(unsigned long long) * FS_CNT_COUNT, st));
    if (a.n_tok == 0 || idx->n_script_tok == 0) return FS_OK;
    const int32_t m_step = dist_m_step(idx->diag), n_step = kBlockN - (idx->diag - 1);
    const int32_t tiles_m = static_cast<int32_t>((a.n_tok + m_step - 1) / m_step);
    const int32_t tiles_n = static_cast<int32_t>((idx->n_script_tok + n_step - 1) / n_step);
    // (+1 tile: in pair mode an odd tile count is rounded up to a full pair)
    const int64_t thr_pad = static_cast<int64_t>(tiles_m + 1) * kBlockM;
    int64_t want_cand = idx->cand_cap > 0 ? idx->cand_cap : (1 << 20);
    if ((r = fs_index_reserve(idx, a.n_tok, want_cand)) != FS_OK) return r;
    if (thr_pad > idx->thr_cap) {  // the workspace must hold whole tiles
        set_error("internal: workspace too small (%lld > %lld bounds)", static_cast<long long>(thr_pad),
                  static_cast<long long>(idx->thr_cap));
        return FS_E_INVALID;
    }
    DistParams p{};
    p.fan_ac = idx->fan_thr;
    p.script_bd = idx->script_norm;
    p.script_mm32 = idx->script_norm_min;
    p.n_fan_tok = a.n_tok;
    p.n_script_tok = idx->n_script_tok;
    p.chunks = (idx->dim_pad + kChunkK - 1) / kChunkK;
    p.last_chunk_ksteps = (idx->dim_pad - (p.chunks - 1) * kChunkK) / kUmmaK;
    p.window = idx->window;
    p.diag = idx->diag;
    p.pair = idx->pair;
    p.f8 = idx->operand_bits == 8;
    p.ares = idx->ares;
    p.pack = idx->pack;
    p.shifts_per_stage = idx->shifts_per_stage > 0 ? idx->shifts_per_stage : idx->window / idx->diag;
    p.group = idx->tile_group;
    {
        // 2: every lane of an epilogue warp waits on the accumulator barrier itself (try_wait with a
        // suspend hint) -- 11 % faster than one polling lane + __syncwarp (profiles/r02_sweep_wait_modes.jsonl);
        // 6 (default): the same after one 200 ns sleep when the first check fails (common.cuh, mbar_wait_mode)
        static const int wait_mode = getenv("FS_DEBUG_WAIT") ? atoi(getenv("FS_DEBUG_WAIT")) : 6;
        p.wait_mode = wait_mode;
    }
    p.tiles_m = tiles_m;
    p.tiles_n = tiles_n;
    p.cand = (mode == Mode::kCandidates) ? cand_out : idx->cand;
    p.cand_cap = (mode == Mode::kCandidates) ? cand_out_cap : idx->cand_cap;
    p.counters = counters;
    p.dump = (mode == Mode::kDots) ? dots : nullptr;
    p.dump_ld = dots_ld;

    // Fused gather (the default kernel): the fan tile is fetched from the operand-row table by token id
    // inside the distance kernel; otherwise the fan operand matrix is materialised first (gather_kernel).
    const bool fused = idx->fused_gather && distance_uses_n128(p) && a.n_extra <= kFusedExtraRows;
    idx->last_fused = fused ? 1 : 0;
    CUtensorMap map_fan, map_fan32;
    if (fused) {
        if ((r = embed_batch(idx, st, a, counters, nullptr, idx->fan_thr, nullptr, thr_pad)) != FS_OK) return r;
        p.fan_tok = a.tok;
        p.n_valid_rows = static_cast<int32_t>(idx->n_base + idx->n_sx + a.n_extra);
        p.n_table_rows = static_cast<int32_t>(idx->n_table_rows);
        map_fan = map_fan32 = idx->map_rows;
    } else {
        if ((r = grow_fan_matrix(idx, a.n_tok)) != FS_OK) return r;
        if ((r = embed_batch(idx, st, a, counters, idx->fan_emb, idx->fan_thr, nullptr, thr_pad)) != FS_OK) return r;
        if ((r = make_token_map(&map_fan, idx->fan_emb, a.n_tok, idx->dim_pad, kBoxRows)) != FS_OK) return r;
        // E = 6 loads the fan tile as four overlapping boxes of 32 rows (common.cuh)
        if ((r = make_token_map(&map_fan32, idx->fan_emb, a.n_tok, idx->dim_pad, kOverlapBoxRows)) != FS_OK) return r;
    }
    const int grid_limit = idx->grid_limit > 0 ? idx->grid_limit : idx->sm_count;
    const int slot = static_cast<int>(idx->ev_count % kTimingRing);
    FS_CUDA_CHECK(cudaEventRecord(idx->ev_start[slot], st));
    if ((r = launch_distance(map_fan, map_fan32, idx->map_script, idx->map_script128, idx->map_script64, p, grid_limit,
                             st)) != FS_OK)
        return r;
    FS_CUDA_CHECK(cudaEventRecord(idx->ev_stop[slot], st));
    idx->ev_count++;
    if (mode != Mode::kSearch) return FS_OK;

    RescoreParams rp{};
    rp.cand = idx->cand;
    rp.counters = counters;
    rp.cand_cap = idx->cand_cap;
    rp.fan_tok = a.tok;
    rp.n_fan_tok = a.n_tok;
    rp.fan_off = a.off;
    rp.n_works = static_cast<int32_t>(a.n_works);
    rp.script_tok = idx->script_tok;
    rp.table = idx->table32;
    rp.n_base = idx->n_base;
    rp.script_extra = idx->sx32;
    rp.n_script_extra = idx->n_sx;
    rp.fan_extra = a.extra;
    rp.n_fan_extra = a.n_extra;
    rp.dim = idx->dim;
    rp.window = idx->window;
    rp.threshold = idx->threshold;
    rp.out = out;
    rp.out_cap = cap;
    rp.match_counter = counters + FS_CNT_MATCHES;
    rp.overflow = counters + FS_CNT_OVERFLOW;
    if ((r = launch_rescore(rp, idx->sm_count, st)) != FS_OK) return r;
    if (idx->lsh_tables > 0) {
        LshParams lp{};
        lp.matches = out;
        lp.match_counter = counters + FS_CNT_MATCHES;
        lp.match_cap = cap;
        lp.normals = idx->lsh_normals;
        lp.n_tables = idx->lsh_tables;
        lp.n_bits = idx->lsh_bits;
        lp.fan_tok = a.tok;
        lp.script_tok = idx->script_tok;
        lp.table = idx->table32;
        lp.n_base = idx->n_base;
        lp.script_extra = idx->sx32;
        lp.n_script_extra = idx->n_sx;
        lp.fan_extra = a.extra;
        lp.n_fan_extra = a.n_extra;
        lp.dim = idx->dim;
        lp.window = idx->window;
        return launch_lsh(lp, idx->sm_count, st);
    }
    return FS_OK;
}

// stage a host batch into the index's device staging buffers
int stage_host_batch(fs_index* idx, const BatchArgs& h, BatchArgs* d) {
    int r;
    cudaStream_t st = idx->stream;
    if ((r = dev_grow(&idx->h_tok, &idx->h_tok_cap, h.n_tok + 8)) != FS_OK) return r;
    if ((r = dev_grow(&idx->h_off, &idx->h_off_cap, h.n_works + 1)) != FS_OK) return r;
    if ((r = dev_grow(&idx->h_extra, &idx->h_extra_cap, h.n_extra * idx->dim)) != FS_OK) return r;
    if (h.n_tok)
        FS_CUDA_CHECK(cudaMemcpyAsync(idx->h_tok, h.tok, sizeof(int32_t) * h.n_tok,
                                      cudaMemcpyHostToDevice, st));
    // ids read past the end of the batch by the last (invalid) windows must stay harmless
    FS_CUDA_CHECK(cudaMemsetAsync(idx->h_tok + h.n_tok, 0xFF, sizeof(int32_t) * 8, st));
    FS_CUDA_CHECK(cudaMemcpyAsync(idx->h_off, h.off, sizeof(int64_t) * (h.n_works + 1),
                                  cudaMemcpyHostToDevice, st));
    if (h.n_extra)
        FS_CUDA_CHECK(cudaMemcpyAsync(idx->h_extra, h.extra, sizeof(float) * h.n_extra * idx->dim,
                                      cudaMemcpyHostToDevice, st));
    *d = BatchArgs{idx->h_tok, h.n_tok, idx->h_off, h.n_works, idx->h_extra, h.n_extra};
    return FS_OK;
}

}  // namespace

extern "C" {

int fs_search_csr_dev(fs_index* idx, void* stream, const int32_t* tok, int64_t n_tok,
                      const int64_t* off, int64_t n_works, const float* extra, int64_t n_extra,
                      fs_match* out, int64_t cap, int64_t* counters) {
    BatchArgs a{tok, n_tok, off, n_works, extra, n_extra};
    int r = check_batch(idx, a, "fs_search_csr_dev");
    if (r != FS_OK) return r;
    if (!counters || cap < 0 || (cap > 0 && !out)) {
        set_error("fs_search_csr_dev: invalid output arguments");
        return FS_E_INVALID;
    }
    return run_pipeline(idx, static_cast<cudaStream_t>(stream), a, Mode::kSearch, out, cap, nullptr, 0,
                        nullptr, 0, reinterpret_cast<unsigned long long*>(counters));
}

// Enqueue one host batch into a free slot: inputs H2D on stream_in, the kernels on `stream` behind
// whatever is already queued there, the counters D2H into the slot's page-locked words.  No
// synchronisation with the device unless a buffer has to grow.
struct TextArgs {  // verbatim token texts of a batch (fs_search_submit_rows); text == nullptr: none
    const char* text = nullptr;
    int64_t text_bytes = 0;
    const uint32_t* tok_start = nullptr;
    const uint16_t* tok_len = nullptr;
    int32_t lsh_filter = 0;
    int64_t cap_rows = 0;
};

static int submit_slot(fs_index* idx, const BatchArgs& h, int64_t cap, int32_t* ticket,
                       const TextArgs& tx = TextArgs()) {
    int r;
    int s = -1;
    for (int k = 0; k < fs_index::kSlots; ++k)
        if (!idx->slots[k].busy) {
            s = k;
            break;
        }
    if (s < 0) {
        set_error("fs_search_submit: %d batches are already in flight; collect one first", fs_index::kSlots);
        return FS_E_INVALID;
    }
    fs_index::Slot& sl = idx->slots[s];
    // the workspace of the shared pipeline must not be re-allocated under a batch in flight:
    // run_pipeline's own reserve (dev_grow) synchronises the device before it frees anything
    if ((r = dev_grow(&sl.d_tok, &sl.tok_cap, h.n_tok + 8)) != FS_OK) return r;
    if ((r = dev_grow(&sl.d_off, &sl.off_cap, h.n_works + 1)) != FS_OK) return r;
    if ((r = dev_grow(&sl.d_extra, &sl.extra_cap, h.n_extra * idx->dim)) != FS_OK) return r;
    if ((r = dev_grow(&sl.d_out, &sl.out_cap, cap)) != FS_OK) return r;
    const bool with_rows = tx.text != nullptr || tx.cap_rows > 0;
    if (with_rows) {
        if ((r = dev_grow(&sl.d_text, &sl.text_cap, tx.text_bytes + 16)) != FS_OK) return r;
        if ((r = dev_grow(&sl.d_tstart, &sl.tstart_cap, h.n_tok + 8)) != FS_OK) return r;
        if ((r = dev_grow(&sl.d_tlen, &sl.tlen_cap, h.n_tok + 8)) != FS_OK) return r;
        if ((r = dev_grow(&sl.d_rows, &sl.rows_buf_cap, tx.cap_rows)) != FS_OK) return r;
        if (idx->pp_tok_cap < h.n_tok + 8) {  // the four per-position arrays grow together
            void* old[] = {idx->pp_head, idx->pp_winner, idx->pp_best_key, idx->pp_best_tie};
            for (void* q : old)
                if (q) {
                    cudaDeviceSynchronize();
                    cudaFree(q);
                }
            idx->pp_head = idx->pp_winner = nullptr;
            idx->pp_best_key = idx->pp_best_tie = nullptr;
            idx->pp_tok_cap = 0;
            const int64_t want = h.n_tok + h.n_tok / 4 + 1024;
            if ((r = dev_alloc(&idx->pp_head, want)) != FS_OK) return r;
            if ((r = dev_alloc(&idx->pp_winner, want)) != FS_OK) return r;
            if ((r = dev_alloc(&idx->pp_best_key, want)) != FS_OK) return r;
            if ((r = dev_alloc(&idx->pp_best_tie, want)) != FS_OK) return r;
            idx->pp_tok_cap = want;
        }
        if (idx->pp_match_cap < cap) {
            void* old[] = {idx->pp_next, idx->pp_lev, idx->pp_rank};
            for (void* q : old)
                if (q) {
                    cudaDeviceSynchronize();
                    cudaFree(q);
                }
            idx->pp_next = idx->pp_lev = idx->pp_rank = nullptr;
            const int64_t want = cap + cap / 4 + 1024;
            if ((r = dev_alloc(&idx->pp_next, want)) != FS_OK) return r;
            if ((r = dev_alloc(&idx->pp_lev, want)) != FS_OK) return r;
            if ((r = dev_alloc(&idx->pp_rank, want)) != FS_OK) return r;
            idx->pp_match_cap = want;
        }
        const int64_t nb = postprocess_scan_blocks(h.n_tok) + 1;
        if (idx->pp_blocks_cap < nb) {
            void* old[] = {idx->pp_block_count, idx->pp_block_off};
            for (void* q : old)
                if (q) {
                    cudaDeviceSynchronize();
                    cudaFree(q);
                }
            idx->pp_block_count = nullptr;
            idx->pp_block_off = nullptr;
            const int64_t want = nb + nb / 4 + 64;
            if ((r = dev_alloc(&idx->pp_block_count, want)) != FS_OK) return r;
            if ((r = dev_alloc(&idx->pp_block_off, want)) != FS_OK) return r;
            idx->pp_blocks_cap = want;
        }
    }
    cudaStream_t in = idx->stream_in, st = idx->stream;
    if (with_rows && h.n_tok) {
        if (tx.text_bytes)
            FS_CUDA_CHECK(cudaMemcpyAsync(sl.d_text, tx.text, static_cast<size_t>(tx.text_bytes), cudaMemcpyHostToDevice, in));
        FS_CUDA_CHECK(cudaMemcpyAsync(sl.d_tstart, tx.tok_start, sizeof(uint32_t) * h.n_tok, cudaMemcpyHostToDevice, in));
        FS_CUDA_CHECK(cudaMemcpyAsync(sl.d_tlen, tx.tok_len, sizeof(uint16_t) * h.n_tok, cudaMemcpyHostToDevice, in));
    }
    if (h.n_tok)
        FS_CUDA_CHECK(cudaMemcpyAsync(sl.d_tok, h.tok, sizeof(int32_t) * h.n_tok, cudaMemcpyHostToDevice, in));
    // ids read past the end of the batch by the last (invalid) windows must stay harmless
    FS_CUDA_CHECK(cudaMemsetAsync(sl.d_tok + h.n_tok, 0xFF, sizeof(int32_t) * 8, in));
    FS_CUDA_CHECK(cudaMemcpyAsync(sl.d_off, h.off, sizeof(int64_t) * (h.n_works + 1), cudaMemcpyHostToDevice, in));
    if (h.n_extra)
        FS_CUDA_CHECK(cudaMemcpyAsync(sl.d_extra, h.extra, sizeof(float) * h.n_extra * idx->dim,
                                      cudaMemcpyHostToDevice, in));
    FS_CUDA_CHECK(cudaEventRecord(sl.ev_in, in));
    FS_CUDA_CHECK(cudaStreamWaitEvent(st, sl.ev_in, 0));
    BatchArgs d{sl.d_tok, h.n_tok, sl.d_off, h.n_works, sl.d_extra, h.n_extra};
    if ((r = run_pipeline(idx, st, d, Mode::kSearch, sl.d_out, cap, nullptr, 0, nullptr, 0, sl.d_counters)) != FS_OK)
        return r;
    if (with_rows && h.n_tok > 0 && idx->n_script_tok > 0) {
        PostParams pp{};
        pp.matches = sl.d_out;
        pp.counters = sl.d_counters;
        pp.match_cap = cap;
        pp.n_tok = static_cast<int32_t>(h.n_tok);
        pp.window = idx->window;
        pp.topk = 10;  // NearestFilter(10), search.py:119-120
        pp.lsh = tx.lsh_filter ? 1 : 0;
        pp.fan_off = sl.d_off;
        pp.n_works = static_cast<int32_t>(h.n_works);
        pp.fan_text = sl.d_text;
        pp.tok_start = sl.d_tstart;
        pp.tok_len = sl.d_tlen;
        pp.script_text = idx->script_text;
        pp.script_word_off = idx->script_word_off;
        pp.n_script_words = idx->n_script_words;
        pp.head = idx->pp_head;
        pp.next = idx->pp_next;
        pp.m_lev = idx->pp_lev;
        pp.m_rank = idx->pp_rank;
        pp.best_key = idx->pp_best_key;
        pp.best_tie = idx->pp_best_tie;
        pp.winner = idx->pp_winner;
        pp.block_count = idx->pp_block_count;
        pp.block_off = idx->pp_block_off;
        pp.rows = sl.d_rows;
        pp.rows_cap = tx.cap_rows;
        pp.overflow = sl.d_counters + FS_CNT_OVERFLOW;
        if ((r = launch_postprocess(pp, idx->sm_count, st)) != FS_OK) return r;
        if (idx->hist_counts &&
            (r = launch_reuse_histogram_rows(sl.d_rows, sl.d_counters, tx.cap_rows, idx->hist_thresholds,
                                             idx->hist_n_thr, idx->n_script_tok, idx->hist_counts, idx->sm_count,
                                             st)) != FS_OK)
            return r;
    }
    FS_CUDA_CHECK(cudaMemcpyAsync(sl.h_counters, sl.d_counters, sizeof(int64_t) * FS_CNT_COUNT,
                                  cudaMemcpyDeviceToHost, st));
    FS_CUDA_CHECK(cudaEventRecord(sl.ev_done, st));
    sl.cap = cap;
    sl.with_rows = with_rows;
    sl.rows_cap = tx.cap_rows;
    sl.busy = true;
    *ticket = s;
    return FS_OK;
}

static int collect_slot_rows(fs_index* idx, int32_t ticket, fs_row* out, int64_t cap, int64_t* counters) {
    fs_index::Slot& sl = idx->slots[ticket];
    FS_CUDA_CHECK(cudaEventSynchronize(sl.ev_done));
    for (int k = 0; k < FS_CNT_COUNT; ++k) counters[k] = sl.h_counters[k];
    const int64_t flags = counters[FS_CNT_OVERFLOW];
    if ((flags & (FS_OVERFLOW_CANDIDATES | FS_OVERFLOW_MATCHES | FS_OVERFLOW_TEXT | FS_OVERFLOW_ROWS)) ||
        counters[FS_CNT_ROWS] > cap) {
        // the ticket stays valid: the caller may still fetch the raw matches (fs_search_collect)
        set_error("device post-processing incomplete (overflow bits %lld, %lld rows for %lld slots)",
                  static_cast<long long>(flags), static_cast<long long>(counters[FS_CNT_ROWS]),
                  static_cast<long long>(cap));
        return FS_E_OVERFLOW;
    }
    sl.busy = false;
    const int64_t n = counters[FS_CNT_ROWS];
    if (n > 0) {
        FS_CUDA_CHECK(cudaMemcpyAsync(out, sl.d_rows, sizeof(fs_row) * n, cudaMemcpyDeviceToHost, idx->stream_out));
        FS_CUDA_CHECK(cudaStreamSynchronize(idx->stream_out));
    }
    return FS_OK;
}

static int collect_slot(fs_index* idx, int32_t ticket, fs_match* out, int64_t cap, int64_t* counters) {
    fs_index::Slot& sl = idx->slots[ticket];
    FS_CUDA_CHECK(cudaEventSynchronize(sl.ev_done));
    sl.busy = false;
    for (int k = 0; k < FS_CNT_COUNT; ++k) counters[k] = sl.h_counters[k];
    const int64_t n_match = counters[FS_CNT_MATCHES];
    int64_t n_copy = n_match < cap ? n_match : cap;
    if (n_copy > sl.cap) n_copy = sl.cap;
    if (n_copy > 0) {
        // on its own stream: the next batch's kernels are already queued on `stream`
        FS_CUDA_CHECK(cudaMemcpyAsync(out, sl.d_out, sizeof(fs_match) * n_copy, cudaMemcpyDeviceToHost,
                                      idx->stream_out));
        FS_CUDA_CHECK(cudaStreamSynchronize(idx->stream_out));
    }
    if (counters[FS_CNT_OVERFLOW] & FS_OVERFLOW_CANDIDATES) {
        set_error("candidate buffer overflow: %lld candidates, capacity %lld; fs_index_reserve more",
                  static_cast<long long>(counters[FS_CNT_CANDIDATES]), static_cast<long long>(idx->cand_cap));
        return FS_E_OVERFLOW;
    }
    if (n_match > cap || n_match > sl.cap) {
        set_error("match buffer overflow: %lld matches, capacity %lld", static_cast<long long>(n_match),
                  static_cast<long long>(cap < sl.cap ? cap : sl.cap));
        return FS_E_OVERFLOW;
    }
    return FS_OK;
}

int fs_search_submit(fs_index* idx, const int32_t* tok, int64_t n_tok, const int64_t* off, int64_t n_works,
                     const float* extra, int64_t n_extra, int64_t cap, int32_t* ticket) {
    BatchArgs h{tok, n_tok, off, n_works, extra, n_extra};
    int r = check_batch(idx, h, "fs_search_submit");
    if (r != FS_OK) return r;
    if (!ticket || cap < 0) {
        set_error("fs_search_submit: invalid output arguments");
        return FS_E_INVALID;
    }
    FS_CUDA_CHECK(cudaSetDevice(idx->device));
    return submit_slot(idx, h, cap, ticket);
}

int fs_index_set_script_text(fs_index* idx, const char* blob, const int64_t* word_off, int64_t n_words) {
    if (!idx || !word_off || n_words != idx->n_script_tok || (n_words > 0 && !blob) || word_off[0] != 0) {
        set_error("fs_index_set_script_text: one word per script token is expected");
        return FS_E_INVALID;
    }
    FS_CUDA_CHECK(cudaSetDevice(idx->device));
    FS_CUDA_CHECK(cudaDeviceSynchronize());
    if (idx->script_text) cudaFree(idx->script_text);
    if (idx->script_word_off) cudaFree(idx->script_word_off);
    idx->script_text = nullptr;
    idx->script_word_off = nullptr;
    idx->n_script_words = 0;
    int r;
    if ((r = dev_alloc(&idx->script_text, word_off[n_words] + 16)) != FS_OK) return r;
    if ((r = dev_alloc(&idx->script_word_off, n_words + 1)) != FS_OK) return r;
    if (word_off[n_words] > 0)
        FS_CUDA_CHECK(cudaMemcpy(idx->script_text, blob, static_cast<size_t>(word_off[n_words]), cudaMemcpyHostToDevice));
    FS_CUDA_CHECK(cudaMemcpy(idx->script_word_off, word_off, sizeof(int64_t) * (n_words + 1), cudaMemcpyHostToDevice));
    idx->n_script_words = n_words;
    return FS_OK;
}

int fs_index_set_reuse_histogram(fs_index* idx, int64_t* counts_dev, const double* thresholds_dev, int32_t n_thr) {
    if (!idx || n_thr < 0 || (n_thr > 0 && (!counts_dev || !thresholds_dev))) {
        set_error("fs_index_set_reuse_histogram: invalid argument");
        return FS_E_INVALID;
    }
    FS_CUDA_CHECK(cudaSetDevice(idx->device));
    FS_CUDA_CHECK(cudaStreamSynchronize(idx->stream));
    idx->hist_counts = n_thr > 0 ? reinterpret_cast<unsigned long long*>(counts_dev) : nullptr;
    idx->hist_thresholds = n_thr > 0 ? thresholds_dev : nullptr;
    idx->hist_n_thr = n_thr;
    return FS_OK;
}

int fs_search_submit_rows(fs_index* idx, const int32_t* tok, int64_t n_tok, const int64_t* off, int64_t n_works,
                          const float* extra, int64_t n_extra, const char* text, int64_t text_bytes,
                          const uint32_t* tok_start, const uint16_t* tok_len, int32_t lsh_filter,
                          int64_t cap_matches, int64_t cap_rows, int32_t* ticket) {
    BatchArgs h{tok, n_tok, off, n_works, extra, n_extra};
    int r = check_batch(idx, h, "fs_search_submit_rows");
    if (r != FS_OK) return r;
    if (!ticket || cap_matches < 0 || cap_rows < 0 || text_bytes < 0 || text_bytes >= (1ll << 32) ||
        (n_tok > 0 && (!tok_start || !tok_len)) || (text_bytes > 0 && !text)) {
        set_error("fs_search_submit_rows: invalid argument (batch text must stay below 4 GiB)");
        return FS_E_INVALID;
    }
    if (idx->n_scripts != 1 || idx->n_script_words != idx->n_script_tok) {
        set_error("fs_search_submit_rows: a single-script index with fs_index_set_script_text is required");
        return FS_E_INVALID;
    }
    FS_CUDA_CHECK(cudaSetDevice(idx->device));
    TextArgs tx;
    tx.text = text ? text : "";
    tx.text_bytes = text_bytes;
    tx.tok_start = tok_start;
    tx.tok_len = tok_len;
    tx.lsh_filter = lsh_filter;
    tx.cap_rows = cap_rows > 0 ? cap_rows : 1;
    return submit_slot(idx, h, cap_matches, ticket, tx);
}

int fs_search_collect_rows(fs_index* idx, int32_t ticket, fs_row* out, int64_t cap, int64_t* counters) {
    if (!idx || ticket < 0 || ticket >= fs_index::kSlots || !idx->slots[ticket].busy ||
        !idx->slots[ticket].with_rows || !counters || cap < 0 || (cap > 0 && !out)) {
        set_error("fs_search_collect_rows: invalid argument (not a ticket of fs_search_submit_rows?)");
        return FS_E_INVALID;
    }
    FS_CUDA_CHECK(cudaSetDevice(idx->device));
    return collect_slot_rows(idx, ticket, out, cap, counters);
}

int fs_search_collect(fs_index* idx, int32_t ticket, fs_match* out, int64_t cap, int64_t* counters) {
    if (!idx || ticket < 0 || ticket >= fs_index::kSlots || !idx->slots[ticket].busy || !counters || cap < 0 ||
        (cap > 0 && !out)) {
        set_error("fs_search_collect: invalid argument (unknown or already collected ticket?)");
        return FS_E_INVALID;
    }
    FS_CUDA_CHECK(cudaSetDevice(idx->device));
    return collect_slot(idx, ticket, out, cap, counters);
}

int fs_search_csr_host(fs_index* idx, const int32_t* tok, int64_t n_tok, const int64_t* off,
                       int64_t n_works, const float* extra, int64_t n_extra, fs_match* out,
                       int64_t cap, int64_t* counters) {
    BatchArgs h{tok, n_tok, off, n_works, extra, n_extra};
    int r = check_batch(idx, h, "fs_search_csr_host");
    if (r != FS_OK) return r;
    if (!counters || cap < 0 || (cap > 0 && !out)) {
        set_error("fs_search_csr_host: invalid output arguments");
        return FS_E_INVALID;
    }
    FS_CUDA_CHECK(cudaSetDevice(idx->device));
    int32_t ticket = -1;
    if ((r = submit_slot(idx, h, cap, &ticket)) != FS_OK) return r;
    return collect_slot(idx, ticket, out, cap, counters);
}

int fs_exact_join_dev(fs_index* idx, void* stream, const int32_t* tok, int64_t n_tok,
                      const int64_t* off, int64_t n_works, fs_pair* out, int64_t cap,
                      int64_t* counters) {
    BatchArgs a{tok, n_tok, off, n_works, nullptr, 0};
    int r = check_batch(idx, a, "fs_exact_join_dev");
    if (r != FS_OK) return r;
    if (!counters || cap < 0 || (cap > 0 && !out)) {
        set_error("fs_exact_join_dev: invalid output arguments");
        return FS_E_INVALID;
    }
    FS_CUDA_CHECK(cudaSetDevice(idx->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long* cnt = reinterpret_cast<unsigned long long*>(counters);
    FS_CUDA_CHECK(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long) * FS_CNT_COUNT, st));
    return launch_hash_probe(tok, n_tok, off, static_cast<int32_t>(n_works), idx->script_tok,
                             idx->window, idx->hash_table, idx->hash_slots, idx->hash_filter,
                             idx->hash_filter_bits, out, cap,
                             cnt + FS_CNT_EXACT, idx->sm_count, st);
}

int fs_exact_join_host(fs_index* idx, const int32_t* tok, int64_t n_tok, const int64_t* off,
                       int64_t n_works, fs_pair* out, int64_t cap, int64_t* counters) {
    BatchArgs h{tok, n_tok, off, n_works, nullptr, 0};
    int r = check_batch(idx, h, "fs_exact_join_host");
    if (r != FS_OK) return r;
    if (!counters || cap < 0 || (cap > 0 && !out)) {
        set_error("fs_exact_join_host: invalid output arguments");
        return FS_E_INVALID;
    }
    FS_CUDA_CHECK(cudaSetDevice(idx->device));
    BatchArgs d;
    if ((r = stage_host_batch(idx, h, &d)) != FS_OK) return r;
    if ((r = dev_grow(&idx->h_pair, &idx->h_pair_cap, cap)) != FS_OK) return r;
    cudaStream_t st = idx->stream;
    if ((r = fs_exact_join_dev(idx, st, d.tok, d.n_tok, d.off, d.n_works, idx->h_pair, cap,
                               reinterpret_cast<int64_t*>(idx->h_counters))) != FS_OK)
        return r;
    FS_CUDA_CHECK(cudaMemcpyAsync(counters, idx->h_counters, sizeof(int64_t) * FS_CNT_COUNT,
                                  cudaMemcpyDeviceToHost, st));
    FS_CUDA_CHECK(cudaStreamSynchronize(st));
    const int64_t n = counters[FS_CNT_EXACT];
    const int64_t n_copy = n < cap ? n : cap;
    if (n_copy > 0) {
        FS_CUDA_CHECK(cudaMemcpyAsync(out, idx->h_pair, sizeof(fs_pair) * n_copy,
                                      cudaMemcpyDeviceToHost, st));
        FS_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    if (n > cap) {
        set_error("pair buffer overflow: %lld pairs, capacity %lld", static_cast<long long>(n),
                  static_cast<long long>(cap));
        return FS_E_OVERFLOW;
    }
    return FS_OK;
}

int fs_stage_embed_dev(fs_index* idx, void* stream, const int32_t* tok, int64_t n_tok,
                       const int64_t* off, int64_t n_works, const float* extra, int64_t n_extra,
                       void* emb_out, float* thr_out) {
    BatchArgs a{tok, n_tok, off, n_works, extra, n_extra};
    int r = check_batch(idx, a, "fs_stage_embed_dev");
    if (r != FS_OK) return r;
    if (!emb_out || !thr_out) {
        set_error("fs_stage_embed_dev: invalid output arguments");
        return FS_E_INVALID;
    }
    FS_CUDA_CHECK(cudaSetDevice(idx->device));
    if ((r = fs_index_reserve(idx, n_tok, idx->cand_cap > 0 ? idx->cand_cap : 1024)) != FS_OK) return r;
    // (the packed bounds go to the workspace the search would use; the caller gets the plain values)
    if (n_tok > idx->thr_cap) {
        set_error("fs_stage_embed_dev: internal workspace too small");
        return FS_E_INVALID;
    }
    return embed_batch(idx, static_cast<cudaStream_t>(stream), a, nullptr, static_cast<__half*>(emb_out),
                       idx->fan_thr, reinterpret_cast<float4*>(thr_out), n_tok);
}

int fs_stage_dots_dev(fs_index* idx, void* stream, const int32_t* tok, int64_t n_tok,
                      const int64_t* off, int64_t n_works, const float* extra, int64_t n_extra,
                      float* dots, int64_t ld) {
    BatchArgs a{tok, n_tok, off, n_works, extra, n_extra};
    int r = check_batch(idx, a, "fs_stage_dots_dev");
    if (r != FS_OK) return r;
    if (!dots || ld <= 0) {
        set_error("fs_stage_dots_dev: invalid output arguments");
        return FS_E_INVALID;
    }
    return run_pipeline(idx, static_cast<cudaStream_t>(stream), a, Mode::kDots, nullptr, 0, nullptr, 0,
                        dots, ld, nullptr);
}

int fs_stage_candidates_dev(fs_index* idx, void* stream, const int32_t* tok, int64_t n_tok,
                            const int64_t* off, int64_t n_works, const float* extra, int64_t n_extra,
                            fs_pair* out, int64_t cap, int64_t* counters) {
    BatchArgs a{tok, n_tok, off, n_works, extra, n_extra};
    int r = check_batch(idx, a, "fs_stage_candidates_dev");
    if (r != FS_OK) return r;
    if (!counters || cap < 0 || (cap > 0 && !out)) {
        set_error("fs_stage_candidates_dev: invalid output arguments");
        return FS_E_INVALID;
    }
    return run_pipeline(idx, static_cast<cudaStream_t>(stream), a, Mode::kCandidates, nullptr, 0, out,
                        cap, nullptr, 0, reinterpret_cast<unsigned long long*>(counters));
}

}  // extern "C"
