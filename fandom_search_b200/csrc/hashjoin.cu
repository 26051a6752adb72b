// Exact 6-gram hash-join with stream compaction.
//
// The reference has no separate exact path: verbatim reuse is the distance-0 case of
// the vector search (search.py:182-184) that ao3.py:353-355 later reads as
// BEST_COMBINED_DISTANCE <= 0.  Two windows have identical vectors exactly when their
// `w` embedding-row ids are identical, so the exact set is an integer join:
//   build : open-addressing table keyed by a hash of the w row ids of every script window
//           (slot = tag:32 | script_pos:32), linear probing, load factor <= 1/16; plus a filter of
//           ~8 bits per script window, one bit set per window (4 .. 128 KB, 32 KB at the headline
//           script)
//   probe : the filter, copied to shared memory, answers "no script window has this hash" for ~9
//           of 10 fan windows without touching the table; the others compare ids exactly on a tag
//           hit; every matching script position is emitted through an atomic cursor.
// 4 B of token id read from HBM per fan window; table (<= 32 MB) L2 resident.
// History (10 M tokens vs a 25 k-token script, profiles/r02_hash_probe_*): round 1 probed a table
// loaded to 0.38 directly -- a third of the windows walked a chain of dependent L2 loads and each
// warp waited for its slowest lane (157 us, 0.04 of the HBM roof, issue active 36 %); a "slot in use"
// bitmap in L1 gave 60 us, then issue-bound at ~120 instructions per window, half of them in the
// rarely-taken table branch that nearly every warp still entered for some lane; rolling the hash
// over ids held in registers and queueing the rare path gave 43 us, bounded by L1 tag look-ups of
// the 32 scattered bitmap words per load (l1tex 69 %); the filter now sits in shared memory.
#include <cstdlib>

#include "common.cuh"

namespace fs {

constexpr unsigned long long kEmptySlot = 0xFFFFFFFFFFFFFFFFull;

// Key hash of a window: the polynomial  a = S*P^w + sum_k id[k]*P^(w-1-k)  (mod 2^32), which rolls
// from one window to the next with two multiply-adds; the slot is the top bits of a*K (Fibonacci
// hashing: they depend on every bit of a).  A second polynomial, finished with the murmur3 32-bit
// finaliser, is the tag stored beside the position.  Ids are compared exactly on a tag hit, so hash
// quality only affects speed.
constexpr uint32_t kMulA = 0x01000193u, kSeedA = 0x811C9DC5u;
constexpr uint32_t kMulB = 0x9E3779B1u, kSeedB = 0x9E3779B9u;
constexpr uint32_t kSlotMul = 0x9E3779B1u;

__host__ __device__ constexpr uint32_t upow(uint32_t b, int e) {
    uint32_t r = 1;
    for (int i = 0; i < e; ++i) r *= b;
    return r;
}

__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}

struct WinHash {
    uint32_t a, tag;
};

template <typename Load>
__device__ __forceinline__ WinHash window_hash(Load load, int32_t window) {
    uint32_t a = kSeedA, b = kSeedB;
    for (int k = 0; k < window; ++k) {
        const uint32_t id = static_cast<uint32_t>(load(k));
        a = a * kMulA + id;
        b = b * kMulB + (id ^ 0x5bd1e995u);
    }
    WinHash h;
    h.a = a;
    h.tag = fmix32(b ^ (a >> 7));
    return h;
}

__device__ __forceinline__ void prefetch_l1(const void* p) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

// slot of the table and bit of the filter both come from the top bits of h = a * kSlotMul
__device__ __forceinline__ uint32_t mix_of(uint32_t a) { return a * kSlotMul; }

__global__ void hash_build_kernel(const int32_t* __restrict__ tok, int64_t n_tok,
                                  const int64_t* __restrict__ off, int32_t n_rows, int32_t window,
                                  unsigned long long* __restrict__ table, uint32_t mask, uint32_t shift,
                                  uint32_t* __restrict__ filter, uint32_t filter_shift) {
    const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= n_tok) return;
    const int32_t row = csr_row_of(off, n_rows, t);
    if (t + window > __ldg(off + row + 1)) return;
    const int32_t* p = tok + t;
    const WinHash h = window_hash([&](int k) { return __ldg(p + k); }, window);
    const unsigned long long entry =
        (static_cast<unsigned long long>(h.tag) << 32) | static_cast<uint32_t>(t);
    const uint32_t m = mix_of(h.a);
    const uint32_t fbit = m >> filter_shift;
    atomicOr(filter + (fbit >> 5), 1u << (fbit & 31));
    uint32_t slot = m >> shift;
    while (true) {
        const unsigned long long prev = atomicCAS(table + slot, kEmptySlot, entry);
        if (prev == kEmptySlot) break;
        slot = (slot + 1) & mask;
    }
}

// Probe.  The filter (one bit per value of the top bits of the key hash, set for every script
// window; 32 KB at the headline script) is copied into SHARED memory by every block -- 32 lanes
// testing 32 random bits cost ~3 bank-conflict wavefronts there, against ~31 tag look-ups when the
// same words sit in L1, which is what bounded the previous version.  A thread takes kProbePer
// CONSECUTIVE windows: their kProbePer + w - 1 ids come straight from global memory as 128-bit
// loads into registers (a warp reads 1 KB contiguous; the few ids shared with the neighbouring
// thread hit L1; the next piece is prefetched), the key hash rolls from window to window.  The
// windows that pass the filter (~1 in 10, plus the real hits) are pushed on a per-warp queue in
// shared memory and then handled one per LANE -- tag, table walk, exact id compare -- so that the
// rare path does not serialise the warp once per window as a divergent branch would.  Warps own
// contiguous runs of 256-token pieces (balanced to within one piece), so the CSR cursor only ever
// walks forward.
constexpr int kProbePer = 8;
constexpr int kProbeWarpSeg = 32 * kProbePer;
constexpr int kProbeWarps = 16;
static_assert(kProbeWarpSeg <= 256, "queue entries are bytes");

// CSR row of token t, searched by a whole warp: every round the 32 lanes test 32 spread positions
// of the remaining range (off is non-decreasing, so the answers are 1..1 0..0 and one ballot narrows
// the range ~32 x): 2-3 dependent loads instead of the ~11 of a binary search.
__device__ __forceinline__ int32_t csr_row_of_warp(const int64_t* __restrict__ off, int32_t n_rows,
                                                   int64_t t, int lane) {
    int32_t lo = 0, hi = n_rows;  // invariant: off[lo] <= t < off[hi]
    while (hi - lo > 1) {
        const int32_t span = hi - lo - 1;  // candidates lo+1 .. hi-1
        const int32_t mine = lo + 1 + static_cast<int32_t>((static_cast<int64_t>(span - 1) * lane) / 31);
        const unsigned le = __ballot_sync(0xFFFFFFFFu, __ldg(off + mine) <= t);
        const int n_le = __popc(le);
        const int32_t below = __shfl_sync(0xFFFFFFFFu, mine, n_le > 0 ? n_le - 1 : 0);
        const int32_t above = __shfl_sync(0xFFFFFFFFu, mine, n_le < 32 ? n_le : 31);
        if (n_le > 0) lo = below;
        if (n_le < 32) hi = above;
    }
    return lo;
}

template <int kW, bool kVec>
__global__ void __launch_bounds__(kProbeWarps * 32, 3)
hash_probe_kernel(const int32_t* __restrict__ tok, int64_t n_tok, const int64_t* __restrict__ off,
                  int32_t n_rows, const int32_t* __restrict__ script_tok,
                  const unsigned long long* __restrict__ table, uint32_t mask, uint32_t shift,
                  const uint32_t* __restrict__ filter, uint32_t filter_shift, fs_pair* __restrict__ out,
                  int64_t cap, unsigned long long* counter) {
    constexpr int kIds = kProbePer + kW - 1;       // ids a thread needs
    constexpr int kIdVec = (kIds + 3) / 4;         // ... as 128-bit loads
    constexpr uint32_t kPw = upow(kMulA, kW);
    constexpr uint32_t kRollC = kSeedA * kPw * (1u - kMulA);  // a' = a*P + id_in - id_out*P^w + kRollC
    extern __shared__ __align__(16) uint32_t probe_smem[];
    const uint32_t filter_words = 1u << (32 - filter_shift - 5);
    uint32_t* bits = probe_smem;
    // [kProbeWarps][2][kProbeWarpSeg] bytes: the windows of a piece that passed the filter; two buffers,
    // because they are looked up in the table one piece LATER (their table lines, prefetched when the
    // filter bit was seen, have arrived by then).  Every lane writes its own hits at an offset from a
    // ballot prefix sum: the count is a warp-uniform register, no shared-memory atomics.
    uint8_t* queue_pos = reinterpret_cast<uint8_t*>(probe_smem + filter_words);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_piece = (n_tok + kProbeWarpSeg - 1) / kProbeWarpSeg;
    const int64_t gw = static_cast<int64_t>(blockIdx.x) * kProbeWarps + warp;
    const int64_t nw = static_cast<int64_t>(gridDim.x) * kProbeWarps;
    int64_t piece = n_piece * gw / nw;
    const int64_t piece_end = n_piece * (gw + 1) / nw;
    // ids are prefetched two pieces ahead (1 KB + the halo line each)
    auto prefetch_piece = [&](int64_t pc) {
        const int64_t at = pc * kProbeWarpSeg + lane * 32;
        if (pc < piece_end && lane <= 8 && at < n_tok) prefetch_l1(tok + at);
    };
    prefetch_piece(piece);
    prefetch_piece(piece + 1);
    for (uint32_t i = threadIdx.x; i < filter_words / 4; i += kProbeWarps * 32)
        reinterpret_cast<uint4*>(bits)[i] = __ldg(reinterpret_cast<const uint4*>(filter) + i);
    __syncthreads();
    uint8_t* my_pos = queue_pos + warp * 2 * kProbeWarpSeg;
    const uint32_t lanes_below = (1u << lane) - 1;

    auto lookup_queued = [&](int buf, int64_t base, int32_t n_queued) {
        for (int32_t q = lane; q < n_queued; q += 32) {
            const int64_t t = base + my_pos[buf * kProbeWarpSeg + q];
            const int32_t* p = tok + t;
            const WinHash h = window_hash([&](int k) { return __ldg(p + k); }, kW);
            uint32_t slot = mix_of(h.a) >> shift;
            unsigned long long cur = __ldg(table + slot);
            while (cur != kEmptySlot) {  // 15 of 16 of the filter's false positives end at once
                if (static_cast<uint32_t>(cur >> 32) == h.tag) {
                    const int32_t j = static_cast<int32_t>(static_cast<uint32_t>(cur));
                    bool same = true;
#pragma unroll
                    for (int k = 0; k < kW; ++k) same = same && (__ldg(p + k) == __ldg(script_tok + j + k));
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
    };

    int32_t row = -1;
    int64_t row_end = 0, row_end_next = 0;  // off[row + 1], off[row + 2] (loaded one work ahead)
    int buf = 0;
    int32_t n_prev = 0;  // windows queued by the previous piece (warp-uniform)
    for (; piece < piece_end; ++piece, buf ^= 1) {
        const int64_t base = piece * kProbeWarpSeg;
        const int64_t t0 = base + lane * kProbePer;
        prefetch_piece(piece + 2);
        if (row < 0) {  // first piece of the warp
            row = csr_row_of_warp(off, n_rows, base, lane);
            row_end = __ldg(off + row + 1);
            row_end_next = __ldg(off + min(row + 2, n_rows));
        }
        uint32_t hit = 0;
        if (t0 + kW <= n_tok) {
            while (row_end <= t0 && row + 1 < n_rows) {
                ++row;
                row_end = row_end_next;
                row_end_next = __ldg(off + min(row + 2, n_rows));
            }
            // windows that lie inside one work
            uint32_t valid = 0;
            if (t0 + (kProbePer - 1) + kW <= row_end) {
                valid = (1u << kProbePer) - 1;
            } else {
                int32_t r = row;
                int64_t e = row_end;
                for (int u = 0; u < kProbePer; ++u) {
                    const int64_t t = t0 + u;
                    while (e <= t && r + 1 < n_rows) e = __ldg(off + (++r) + 1);
                    if (t + kW <= e) valid |= 1u << u;
                }
            }
            if (valid) {
                uint32_t id[4 * kIdVec];
                if (kVec && t0 + 4 * kIdVec <= n_tok) {
                    const int4* src = reinterpret_cast<const int4*>(tok + t0);
#pragma unroll
                    for (int q = 0; q < kIdVec; ++q) {
                        const int4 v = __ldg(src + q);
                        id[4 * q] = v.x, id[4 * q + 1] = v.y, id[4 * q + 2] = v.z, id[4 * q + 3] = v.w;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < kIds; ++k)
                        id[k] = (t0 + k < n_tok) ? static_cast<uint32_t>(__ldg(tok + t0 + k)) : 0xFFFFFFFFu;
                }
                uint32_t a = kSeedA;
#pragma unroll
                for (int k = 0; k < kW; ++k) a = a * kMulA + id[k];
#pragma unroll
                for (int u = 0; u < kProbePer; ++u) {
                    if (u > 0) a = a * kMulA + id[u + kW - 1] - id[u - 1] * kPw + kRollC;
                    const uint32_t m = mix_of(a);
                    const uint32_t fbit = m >> filter_shift;
                    const uint32_t pass = (bits[fbit >> 5] >> (fbit & 31)) & 1u;
                    if (pass) prefetch_l1(table + (m >> shift));  // rare
                    hit |= pass << u;
                }
                hit &= valid;
            }
        }
        // queue offset of this lane = hits of the lanes below: prefix sum of popc(hit) (0..8, four bits)
        // from one ballot per bit
        const uint32_t mine = __popc(hit);
        int32_t at = 0, n_now = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const unsigned who = __ballot_sync(0xFFFFFFFFu, (mine >> b) & 1u);
            at += __popc(who & lanes_below) << b;
            n_now += __popc(who) << b;
        }
        at += buf * kProbeWarpSeg;
        while (hit) {
            my_pos[at++] = static_cast<uint8_t>(lane * kProbePer + __ffs(hit) - 1);
            hit &= hit - 1;
        }
        lookup_queued(buf ^ 1, base - kProbeWarpSeg, n_prev);  // the previous piece's
        n_prev = n_now;
        __syncwarp();
    }
    lookup_queued(buf ^ 1, piece * kProbeWarpSeg - kProbeWarpSeg, n_prev);  // the last piece's
}

static uint32_t log2_of(uint32_t pow2) {
    uint32_t bits = 0;
    while ((1u << bits) < pow2) ++bits;
    return bits;
}

uint32_t hash_filter_bits(int64_t n_script_tok) {
    // ~8 bits per script window (about one fan window in nine passes), 4 KB .. 128 KB of shared memory
    uint32_t bits = 1u << 15;
    uint64_t per_window = 8;
    if (const char* e = getenv("FS_DEBUG_FILTER_BITS_PER_WINDOW")) per_window = static_cast<uint64_t>(atoi(e));
    while (bits < per_window * static_cast<uint64_t>(n_script_tok) && bits < (1u << 20)) bits <<= 1;
    return bits;
}

int launch_hash_build(const int32_t* tok, int64_t n_tok, const int64_t* off, int32_t n_rows,
                      int32_t window, unsigned long long* table, uint32_t slots, uint32_t* filter,
                      uint32_t filter_bits, cudaStream_t stream) {
    FS_CUDA_CHECK(cudaMemsetAsync(table, 0xFF, static_cast<size_t>(slots) * 8, stream));
    FS_CUDA_CHECK(cudaMemsetAsync(filter, 0, static_cast<size_t>(filter_bits) / 8, stream));
    if (n_tok <= 0) return FS_OK;
    const int threads = 256;
    const int64_t blocks = (n_tok + threads - 1) / threads;
    hash_build_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(
        tok, n_tok, off, n_rows, window, table, slots - 1, 32 - log2_of(slots), filter, 32 - log2_of(filter_bits));
    FS_CUDA_CHECK(cudaGetLastError());
    return FS_OK;
}

namespace {
struct ProbeArgs {
    const int32_t* tok;
    int64_t n_tok;
    const int64_t* off;
    int32_t n_rows;
    const int32_t* script_tok;
    const unsigned long long* table;
    uint32_t slots;
    const uint32_t* filter;
    uint32_t filter_bits;
    fs_pair* out;
    int64_t cap;
    unsigned long long* counter;
};

template <int kW, bool kVec>
int launch_probe_wv(const ProbeArgs& a, int sm_count, cudaStream_t stream) {
    const size_t smem = a.filter_bits / 8 + static_cast<size_t>(kProbeWarps) * 2 * kProbeWarpSeg;
    auto kernel = hash_probe_kernel<kW, kVec>;
    // per device and instantiation: the opt-in only matters above 48 KB, where it is one driver call
    if (smem > 48 * 1024)
        FS_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    // resident blocks per SM: 3 by registers (__launch_bounds__), fewer when the filter is large
    int per_sm = static_cast<int>((227 * 1024) / (smem + 2048));
    per_sm = per_sm < 1 ? 1 : per_sm > 3 ? 3 : per_sm;
    const int64_t pieces = (a.n_tok + kProbeWarpSeg - 1) / kProbeWarpSeg;
    int64_t blocks = (pieces + kProbeWarps - 1) / kProbeWarps;
    const int64_t max_blocks = static_cast<int64_t>(sm_count) * per_sm;
    if (blocks > max_blocks) blocks = max_blocks;
    kernel<<<static_cast<unsigned>(blocks), kProbeWarps * 32, smem, stream>>>(
        a.tok, a.n_tok, a.off, a.n_rows, a.script_tok, a.table, a.slots - 1, 32 - log2_of(a.slots), a.filter,
        32 - log2_of(a.filter_bits), a.out, a.cap, a.counter);
    FS_CUDA_CHECK(cudaGetLastError());
    return FS_OK;
}

template <int kW>
int launch_probe_w(const ProbeArgs& a, int sm_count, cudaStream_t stream) {
    // 128-bit id loads need the alignment
    return (reinterpret_cast<uintptr_t>(a.tok) & 15u) == 0 ? launch_probe_wv<kW, true>(a, sm_count, stream)
                                                           : launch_probe_wv<kW, false>(a, sm_count, stream);
}
}  // namespace

int launch_hash_probe(const int32_t* tok, int64_t n_tok, const int64_t* off, int32_t n_rows,
                      const int32_t* script_tok, int32_t window, const unsigned long long* table,
                      uint32_t slots, const uint32_t* filter, uint32_t filter_bits, fs_pair* out, int64_t cap,
                      unsigned long long* counter, int sm_count, cudaStream_t stream) {
    if (n_tok <= 0) return FS_OK;
    const ProbeArgs a{tok, n_tok, off, n_rows, script_tok, table, slots, filter, filter_bits, out, cap, counter};
    switch (window) {
        case 1: return launch_probe_w<1>(a, sm_count, stream);
        case 2: return launch_probe_w<2>(a, sm_count, stream);
        case 3: return launch_probe_w<3>(a, sm_count, stream);
        case 4: return launch_probe_w<4>(a, sm_count, stream);
        case 5: return launch_probe_w<5>(a, sm_count, stream);
        case 6: return launch_probe_w<6>(a, sm_count, stream);
        case 7: return launch_probe_w<7>(a, sm_count, stream);
        case 8: return launch_probe_w<8>(a, sm_count, stream);
        default:
            set_error("hash join: window must be 1..8");
            return FS_E_INVALID;
    }
}

}  // namespace fs
