// Float64 re-scoring of the tensor-core candidates.
//
// Restates, per candidate pair, exactly what the reference computes per candidate
// (search.py:123 store_vector, :178 neighbours -> nearpy CosineDistance):
//     nv_s = s / ||s||_2      (unit-normalised when stored)
//     nv_f = f / ||f||_2      (unit-normalised query)
//     distance = 1.0 - dot(nv_s, nv_f)            all float64
// from the ORIGINAL float32 embedding rows (the reference assigns float32 spaCy
// vectors into a float64 array, search.py:72-75), then applies
// `distance < distance_threshold` (search.py:184).  One warp per candidate.
#include "common.cuh"

namespace fs {

__device__ __forceinline__ const float* row_ptr(const RescoreParams& p, int64_t id) {
    if (id >= 0 && id < p.n_base) return p.table + id * p.dim;
    id -= p.n_base;
    if (id >= 0 && id < p.n_script_extra) return p.script_extra + id * p.dim;
    id -= p.n_script_extra;
    if (id >= 0 && id < p.n_fan_extra) return p.fan_extra + id * p.dim;
    return nullptr;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void rescore_kernel(const RescoreParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    unsigned long long n = p.counters[FS_CNT_CANDIDATES];
    if (n > static_cast<unsigned long long>(p.cand_cap)) {
        // pairs were lost in the pre-filter's compaction: tell the caller (the host entry points turn
        // this into FS_E_OVERFLOW, a device-pointer caller reads the bit once its stream has drained)
        if (warp_global == 0 && lane == 0) atomicOr(p.overflow, static_cast<unsigned long long>(FS_OVERFLOW_CANDIDATES));
        n = p.cand_cap;
    }
    for (int64_t c = warp_global; c < static_cast<int64_t>(n); c += n_warps) {
        const fs_pair pr = p.cand[c];
        double sf = 0.0, ss = 0.0;
        bool same = true;
        for (int k = 0; k < p.window; ++k) {
            const int32_t idf = __ldg(p.fan_tok + pr.fan_pos + k);
            const int32_t ids = __ldg(p.script_tok + pr.script_pos + k);
            same = same && (idf == ids);
            const float* f = row_ptr(p, idf);
            const float* s = row_ptr(p, ids);
            for (int e = lane; e < p.dim; e += 32) {
                const double fv = f ? static_cast<double>(f[e]) : 0.0;
                const double sv = s ? static_cast<double>(s[e]) : 0.0;
                sf = fma(fv, fv, sf);
                ss = fma(sv, sv, ss);
            }
        }
        sf = warp_sum(sf);
        ss = warp_sum(ss);
        const double nf = sqrt(sf), ns = sqrt(ss);
        // unitvec leaves an all-zero vector unchanged -> dot = 0 -> distance = 1
        const double inv_f_den = nf > 0.0 ? nf : 1.0;
        const double inv_s_den = ns > 0.0 ? ns : 1.0;
        double dot = 0.0;
        for (int k = 0; k < p.window; ++k) {
            const float* f = row_ptr(p, __ldg(p.fan_tok + pr.fan_pos + k));
            const float* s = row_ptr(p, __ldg(p.script_tok + pr.script_pos + k));
            if (!f || !s) continue;
            for (int e = lane; e < p.dim; e += 32) {
                const double fv = static_cast<double>(f[e]) / inv_f_den;
                const double sv = static_cast<double>(s[e]) / inv_s_den;
                dot = fma(sv, fv, dot);
            }
        }
        dot = warp_sum(dot);
        const double dist = 1.0 - dot;
        if (lane == 0 && dist < p.threshold) {
            const unsigned long long slot = atomicAdd(p.match_counter, 1ull);
            if (slot < static_cast<unsigned long long>(p.out_cap)) {
                fs_match m;
                m.fan_pos = pr.fan_pos;
                m.script_pos = pr.script_pos;
                m.distance = dist;
                m.work = csr_row_of(p.fan_off, p.n_works, pr.fan_pos);
                m.flags = same ? FS_MATCH_EXACT : 0u;
                p.out[slot] = m;
            } else {
                atomicOr(p.overflow, static_cast<unsigned long long>(FS_OVERFLOW_MATCHES));
            }
        }
    }
}

int launch_rescore(const RescoreParams& p, int sm_count, cudaStream_t stream) {
    rescore_kernel<<<sm_count * 4, 256, 0, stream>>>(p);
    FS_CUDA_CHECK(cudaGetLastError());
    return FS_OK;
}

}  // namespace fs
