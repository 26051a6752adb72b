// LSH-emulation post-filter (parity mode only).
//
// The reference does not compare a fan window with every script window: nearpy looks the
// window up in 15 hash tables keyed by the sign pattern of 14 random projections
// (search.py:112-116 RandomBinaryProjections('rbp{i}', 14); key = '1' if normal . v > 0.0)
// and only scores the union of those buckets (search.py:178).  A pair is therefore seen by
// the reference iff the RAW (un-normalised) fan and script window vectors have the same 14
// sign bits in at least one table.  Given the hyperplanes of a (seeded) reference run, this
// kernel evaluates exactly that predicate in float64 for every surviving match and records
// the FIRST table in which the keys agree (nearpy's candidate order), so an exhaustive GPU
// search can be reduced to precisely what that LSH index would have returned.
#include "common.cuh"

namespace fs {

constexpr int kLshThreads = 256;

__global__ void __launch_bounds__(kLshThreads)
lsh_first_table_kernel(const LshParams p) {
    extern __shared__ double lsh_smem[];
    const int wd = p.window * p.dim;
    double* f = lsh_smem;            // [wd]
    double* s = lsh_smem + wd;       // [wd]
    unsigned char* bits = reinterpret_cast<unsigned char*>(lsh_smem + 2 * wd);  // [2][planes]
    const int planes = p.n_tables * p.n_bits;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long n = *p.match_counter;
    if (n > static_cast<unsigned long long>(p.match_cap)) n = p.match_cap;
    for (int64_t mi = blockIdx.x; mi < static_cast<int64_t>(n); mi += gridDim.x) {
        const fs_match m = p.matches[mi];
        for (int k = 0; k < p.window; ++k) {
            const float* fr = lsh_row_ptr(p, __ldg(p.fan_tok + m.fan_pos + k));
            const float* sr = lsh_row_ptr(p, __ldg(p.script_tok + m.script_pos + k));
            for (int e = threadIdx.x; e < p.dim; e += blockDim.x) {
                f[k * p.dim + e] = fr ? static_cast<double>(fr[e]) : 0.0;
                s[k * p.dim + e] = sr ? static_cast<double>(sr[e]) : 0.0;
            }
        }
        __syncthreads();
        for (int h = warp; h < planes; h += kLshThreads / 32) {
            const double* nrm = p.normals + static_cast<int64_t>(h) * wd;
            double pf = 0.0, ps = 0.0;
            for (int e = lane; e < wd; e += 32) {
                const double nv = __ldg(nrm + e);
                pf = fma(nv, f[e], pf);
                ps = fma(nv, s[e], ps);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                pf += __shfl_xor_sync(0xffffffffu, pf, o);
                ps += __shfl_xor_sync(0xffffffffu, ps, o);
            }
            if (lane == 0) {
                bits[h] = pf > 0.0;
                bits[planes + h] = ps > 0.0;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int first = -1;
            for (int t = 0; t < p.n_tables && first < 0; ++t) {
                bool same = true;
                for (int b = 0; b < p.n_bits; ++b)
                    same = same && bits[t * p.n_bits + b] == bits[planes + t * p.n_bits + b];
                if (same) first = t;
            }
            const uint32_t keep = p.matches[mi].flags & ~(0xFFu << FS_MATCH_LSH_SHIFT);
            p.matches[mi].flags = keep | (static_cast<uint32_t>(first + 1) << FS_MATCH_LSH_SHIFT);
        }
        __syncthreads();
    }
}

int launch_lsh(const LshParams& p, int sm_count, cudaStream_t stream) {
    const int wd = p.window * p.dim;
    const size_t smem = sizeof(double) * 2 * wd + 2 * static_cast<size_t>(p.n_tables) * p.n_bits + 16;
    // the attribute belongs to the (function, device) pair and the call is cheap: set it on the
    // current device before every launch that needs more than the default 48 KB
    if (smem > 48 * 1024)
        FS_CUDA_CHECK(cudaFuncSetAttribute(lsh_first_table_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem)));
    lsh_first_table_kernel<<<sm_count * 4, kLshThreads, smem, stream>>>(p);
    FS_CUDA_CHECK(cudaGetLastError());
    return FS_OK;
}

}  // namespace fs
