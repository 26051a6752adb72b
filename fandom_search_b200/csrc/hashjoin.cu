// Exact 6-gram hash-join with stream compaction.
//
// The reference has no separate exact path: verbatim reuse is the distance-0 case of
// the vector search (search.py:182-184) that ao3.py:353-355 later reads as
// BEST_COMBINED_DISTANCE <= 0.  Two windows have identical vectors exactly when their
// `w` embedding-row ids are identical, so the exact set is an integer join:
//   build : open-addressing table keyed by a 64-bit mix of the w row ids of every
//           script window (slot = tag:32 | script_pos:32), linear probing
//   probe : one thread per fan window; ids compared exactly on a tag hit, every
//           matching script position emitted through an atomic cursor.
// HBM-bound: 4 B of token id read per fan window, table (<= 8 MB) L2 resident.
#include "common.cuh"

namespace fs {

constexpr unsigned long long kEmptySlot = 0xFFFFFFFFFFFFFFFFull;

__device__ __forceinline__ uint64_t mix64(uint64_t h) {
    h ^= h >> 33;
    h *= 0xff51afd7ed558ccdULL;
    h ^= h >> 33;
    h *= 0xc4ceb9fe1a85ec53ULL;
    h ^= h >> 33;
    return h;
}

__device__ __forceinline__ uint64_t window_hash(const int32_t* __restrict__ tok, int32_t window) {
    uint64_t h = 0x9E3779B97F4A7C15ULL;
    for (int k = 0; k < window; ++k) {
        h = (h ^ static_cast<uint64_t>(static_cast<uint32_t>(__ldg(tok + k)))) * 0xff51afd7ed558ccdULL;
        h ^= h >> 29;
    }
    return mix64(h);
}

__global__ void hash_build_kernel(const int32_t* __restrict__ tok, int64_t n_tok,
                                  const int64_t* __restrict__ off, int32_t n_rows, int32_t window,
                                  unsigned long long* __restrict__ table, uint32_t mask) {
    const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= n_tok) return;
    const int32_t row = csr_row_of(off, n_rows, t);
    if (t + window > __ldg(off + row + 1)) return;
    const uint64_t h = window_hash(tok + t, window);
    const unsigned long long entry = ((h >> 32) << 32) | static_cast<uint32_t>(t);
    uint32_t slot = static_cast<uint32_t>(h) & mask;
    while (true) {
        const unsigned long long prev = atomicCAS(table + slot, kEmptySlot, entry);
        if (prev == kEmptySlot) break;
        slot = (slot + 1) & mask;
    }
}

// One block walks segments of kProbeSeg consecutive tokens; the CSR row of the segment start
// is searched once, every thread then handles kProbeSeg/256 windows with independent loads in
// flight (ids -> hash -> table slot), so the L2 latencies of different windows overlap.
constexpr int kProbeSeg = 2048;
constexpr int kProbePerThread = kProbeSeg / 256;

__global__ void __launch_bounds__(256)
hash_probe_kernel(const int32_t* __restrict__ tok, int64_t n_tok, const int64_t* __restrict__ off,
                  int32_t n_rows, const int32_t* __restrict__ script_tok, int32_t window,
                  const unsigned long long* __restrict__ table, uint32_t mask,
                  fs_pair* __restrict__ out, int64_t cap, unsigned long long* counter) {
    __shared__ int32_t row_hint;
    for (int64_t seg = static_cast<int64_t>(blockIdx.x) * kProbeSeg; seg < n_tok;
         seg += static_cast<int64_t>(gridDim.x) * kProbeSeg) {
        __syncthreads();
        if (threadIdx.x == 0) row_hint = csr_row_of(off, n_rows, seg);
        __syncthreads();
        uint64_t h[kProbePerThread];
        bool valid[kProbePerThread];
        int32_t row = row_hint;
#pragma unroll
        for (int u = 0; u < kProbePerThread; ++u) {
            const int64_t t = seg + u * 256 + threadIdx.x;
            valid[u] = false;
            h[u] = 0;
            if (t < n_tok) {
                row = csr_row_from_hint(off, n_rows, t, row);
                if (t + window <= __ldg(off + row + 1)) {
                    valid[u] = true;
                    h[u] = window_hash(tok + t, window);
                }
            }
        }
        unsigned long long e[kProbePerThread];
#pragma unroll
        for (int u = 0; u < kProbePerThread; ++u)
            e[u] = valid[u] ? __ldg(table + (static_cast<uint32_t>(h[u]) & mask)) : kEmptySlot;
#pragma unroll
        for (int u = 0; u < kProbePerThread; ++u) {
            if (e[u] == kEmptySlot) continue;  // the common case: no script window hashes here
            const int64_t t = seg + u * 256 + threadIdx.x;
            const uint32_t tag = static_cast<uint32_t>(h[u] >> 32);
            uint32_t slot = static_cast<uint32_t>(h[u]) & mask;
            unsigned long long cur = e[u];
            while (cur != kEmptySlot) {
                if (static_cast<uint32_t>(cur >> 32) == tag) {
                    const int32_t j = static_cast<int32_t>(static_cast<uint32_t>(cur));
                    bool same = true;
                    for (int k = 0; k < window; ++k)
                        same = same && (__ldg(tok + t + k) == __ldg(script_tok + j + k));
                    if (same) {
                        const unsigned long long s = atomicAdd(counter, 1ull);
                        if (s < static_cast<unsigned long long>(cap)) {
                            out[s].fan_pos = static_cast<int32_t>(t);
                            out[s].script_pos = j;
                        }
                    }
                }
                slot = (slot + 1) & mask;
                cur = __ldg(table + slot);
            }
        }
    }
}

int launch_hash_build(const int32_t* tok, int64_t n_tok, const int64_t* off, int32_t n_rows,
                      int32_t window, unsigned long long* table, uint32_t slots,
                      cudaStream_t stream) {
    FS_CUDA_CHECK(cudaMemsetAsync(table, 0xFF, static_cast<size_t>(slots) * 8, stream));
    if (n_tok <= 0) return FS_OK;
    const int threads = 256;
    const int64_t blocks = (n_tok + threads - 1) / threads;
    hash_build_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(tok, n_tok, off, n_rows,
                                                                            window, table, slots - 1);
    FS_CUDA_CHECK(cudaGetLastError());
    return FS_OK;
}

int launch_hash_probe(const int32_t* tok, int64_t n_tok, const int64_t* off, int32_t n_rows,
                      const int32_t* script_tok, int32_t window, const unsigned long long* table,
                      uint32_t slots, fs_pair* out, int64_t cap, unsigned long long* counter,
                      int sm_count, cudaStream_t stream) {
    if (n_tok <= 0) return FS_OK;
    const int threads = 256;
    int64_t blocks = (n_tok + kProbeSeg - 1) / kProbeSeg;
    const int64_t max_blocks = static_cast<int64_t>(sm_count) * 8;
    if (blocks > max_blocks) blocks = max_blocks;
    hash_probe_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(
        tok, n_tok, off, n_rows, script_tok, window, table, slots - 1, out, cap, counter);
    FS_CUDA_CHECK(cudaGetLastError());
    return FS_OK;
}

}  // namespace fs
