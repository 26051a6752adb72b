// Per-script-word reuse histogram: the step right after the search.
//
// ao3.py format_data (ao3.py:351-363, 407-411) reads the match CSV back and, for thresholds
// t in {0, .05, ..., .5}, counts per ORIGINAL_SCRIPT_WORD_INDEX the rows with
// BEST_COMBINED_DISTANCE <= t (boolean columns summed by a pandas group-by).  This kernel
// computes the same table straight from the winning records of a cluster, before they are
// ever formatted as text: counts[word][k] += (combined <= thresholds[k]).
#include "common.cuh"

namespace fs {

__global__ void reuse_histogram_kernel(const int32_t* __restrict__ word_ix,
                                       const double* __restrict__ combined, int64_t n,
                                       const double* __restrict__ thresholds, int32_t n_thr,
                                       int64_t n_words, unsigned long long* __restrict__ counts) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t w = word_ix[i];
        if (w < 0 || w >= n_words) continue;
        const double c = combined[i];
        for (int k = 0; k < n_thr; ++k)
            if (c <= thresholds[k]) atomicAdd(counts + w * n_thr + k, 1ull);
    }
}

// The same straight from the device's winning rows (fs_search_submit_rows): the rows never leave the
// GPU before they are counted.  ORIGINAL_SCRIPT_WORD_INDEX = match_ix + window_ix (search.py:200),
// BEST_COMBINED_DISTANCE = distance * lev (search.py:217).
__global__ void reuse_histogram_rows_kernel(const fs_row* __restrict__ rows,
                                            const unsigned long long* __restrict__ counters, int64_t rows_cap,
                                            const double* __restrict__ thresholds, int32_t n_thr,
                                            int64_t n_words, unsigned long long* __restrict__ counts) {
    unsigned long long n = counters[FS_CNT_ROWS];
    if (n > static_cast<unsigned long long>(rows_cap)) n = rows_cap;
    if (counters[FS_CNT_OVERFLOW] != 0) return;  // incomplete batch: the host redoes it and counts it itself
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < static_cast<int64_t>(n);
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const fs_row r = rows[i];
        const int64_t w = static_cast<int64_t>(r.match_ix) + r.window_ix;
        if (w < 0 || w >= n_words) continue;
        const double c = r.distance * static_cast<double>(r.lev);
        for (int k = 0; k < n_thr; ++k)
            if (c <= thresholds[k]) atomicAdd(counts + w * n_thr + k, 1ull);
    }
}

int launch_reuse_histogram_rows(const fs_row* rows, const unsigned long long* counters, int64_t rows_cap,
                                const double* thresholds, int32_t n_thr, int64_t n_words,
                                unsigned long long* counts, int sm_count, cudaStream_t stream) {
    reuse_histogram_rows_kernel<<<sm_count * 4, 256, 0, stream>>>(rows, counters, rows_cap, thresholds, n_thr,
                                                                 n_words, counts);
    FS_CUDA_CHECK(cudaGetLastError());
    return FS_OK;
}

}  // namespace fs

extern "C" int fs_reuse_histogram_dev(void* stream, const int32_t* word_ix, const double* combined,
                                      int64_t n, const double* thresholds, int32_t n_thr,
                                      int64_t n_words, int64_t* counts) {
    if (n < 0 || n_thr < 1 || n_words < 0 || !thresholds || !counts || (n > 0 && (!word_ix || !combined))) {
        fs::set_error("fs_reuse_histogram_dev: invalid argument");
        return FS_E_INVALID;
    }
    if (n == 0) return FS_OK;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    fs::reuse_histogram_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        word_ix, combined, n, thresholds, n_thr, n_words, reinterpret_cast<unsigned long long*>(counts));
    FS_CUDA_CHECK(cudaGetLastError());
    return FS_OK;
}
