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
