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

// Two independent 32-bit polynomial hashes of the w row ids (one IMAD per id each), finished
// with the murmur3 32-bit finaliser: `lo` picks the slot, `hi` is the tag stored beside the
// position.  Ids are compared exactly on a tag hit, so hash quality only affects speed.
__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}

struct WinHash {
    uint32_t lo, hi;
};

template <typename Load>
__device__ __forceinline__ WinHash window_hash(Load load, int32_t window) {
    uint32_t a = 0x811C9DC5u, b = 0x9E3779B9u;
    for (int k = 0; k < window; ++k) {
        const uint32_t id = static_cast<uint32_t>(load(k));
        a = a * 0x01000193u + id;
        b = b * 0x9E3779B1u + (id ^ 0x5bd1e995u);
    }
    WinHash h;
    h.lo = fmix32(a);
    h.hi = fmix32(b ^ (a >> 7));
    return h;
}

__global__ void hash_build_kernel(const int32_t* __restrict__ tok, int64_t n_tok,
                                  const int64_t* __restrict__ off, int32_t n_rows, int32_t window,
                                  unsigned long long* __restrict__ table, uint32_t mask) {
    const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= n_tok) return;
    const int32_t row = csr_row_of(off, n_rows, t);
    if (t + window > __ldg(off + row + 1)) return;
    const int32_t* p = tok + t;
    const WinHash h = window_hash([&](int k) { return __ldg(p + k); }, window);
    const unsigned long long entry =
        (static_cast<unsigned long long>(h.hi) << 32) | static_cast<uint32_t>(t);
    uint32_t slot = h.lo & mask;
    while (true) {
        const unsigned long long prev = atomicCAS(table + slot, kEmptySlot, entry);
        if (prev == kEmptySlot) break;
        slot = (slot + 1) & mask;
    }
}

// One block walks segments of kProbeSeg consecutive tokens: the ids of the segment (+ halo) are
// staged in shared memory with coalesced loads (4 B of HBM traffic per window), the CSR row of
// the segment start is searched once, and every thread handles kProbeSeg/256 windows with their
// table loads issued back to back so the L2 latencies overlap.
constexpr int kProbeSeg = 2048;
constexpr int kProbePerThread = kProbeSeg / 256;
constexpr int kProbeHalo = 8;  // >= window - 1

__global__ void __launch_bounds__(256)
hash_probe_kernel(const int32_t* __restrict__ tok, int64_t n_tok, const int64_t* __restrict__ off,
                  int32_t n_rows, const int32_t* __restrict__ script_tok, int32_t window,
                  const unsigned long long* __restrict__ table, uint32_t mask,
                  fs_pair* __restrict__ out, int64_t cap, unsigned long long* counter) {
    __shared__ int32_t ids[kProbeSeg + kProbeHalo];
    __shared__ int32_t row_hint;
    for (int64_t seg = static_cast<int64_t>(blockIdx.x) * kProbeSeg; seg < n_tok;
         seg += static_cast<int64_t>(gridDim.x) * kProbeSeg) {
        __syncthreads();
        if (threadIdx.x == 0) row_hint = csr_row_of(off, n_rows, seg);
        for (int i = threadIdx.x; i < kProbeSeg + kProbeHalo; i += 256)
            ids[i] = (seg + i < n_tok) ? __ldg(tok + seg + i) : -1;
        __syncthreads();
        uint32_t lo[kProbePerThread], hi[kProbePerThread];
        unsigned long long e[kProbePerThread];
        int32_t row = row_hint;
        int64_t row_end = __ldg(off + row + 1);
#pragma unroll
        for (int u = 0; u < kProbePerThread; ++u) {
            const int32_t i = u * 256 + threadIdx.x;
            const int64_t t = seg + i;
            e[u] = kEmptySlot;
            if (t < n_tok) {
                while (row_end <= t && row + 1 < n_rows) {
                    ++row;
                    row_end = __ldg(off + row + 1);
                }
                if (t + window <= row_end) {
                    const WinHash h = window_hash([&](int k) { return ids[i + k]; }, window);
                    lo[u] = h.lo;
                    hi[u] = h.hi;
                    e[u] = __ldg(table + (h.lo & mask));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kProbePerThread; ++u) {
            if (e[u] == kEmptySlot) continue;  // the common case: no script window hashes here
            const int32_t i = u * 256 + threadIdx.x;
            uint32_t slot = lo[u] & mask;
            unsigned long long cur = e[u];
            while (cur != kEmptySlot) {
                if (static_cast<uint32_t>(cur >> 32) == hi[u]) {
                    const int32_t j = static_cast<int32_t>(static_cast<uint32_t>(cur));
                    bool same = true;
                    for (int k = 0; k < window; ++k) same = same && (ids[i + k] == __ldg(script_tok + j + k));
                    if (same) {
                        const unsigned long long s = atomicAdd(counter, 1ull);
                        if (s < static_cast<unsigned long long>(cap)) {
                            out[s].fan_pos = static_cast<int32_t>(seg + i);
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
