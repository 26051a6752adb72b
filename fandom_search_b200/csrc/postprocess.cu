// Device-side post-processing of the float64 match list: what the reference does
// per fan window after `engine.neighbours` (search.py:178-226), for a whole cluster at once and
// without the match list ever leaving the GPU:
//
//   NearestFilter(10)      top-10 of a window's candidates by (distance, first LSH table, script
//                          position) -- nearpy's stable sort over its candidate order
//   search.py:189-190      Levenshtein("s0 ... s5", "[t0, ..., t5]") over Unicode code points
//   search.py:192-218      six records per surviving pair, keyed by fan word
//   search.py:224-225      per fan word the record of minimal BEST_COMBINED_DISTANCE, the FIRST
//                          inserted one among equals (windows ascend, candidates in sorted order)
//   search.py:226          rows sorted by word index
//
// Everything is an order-independent reduction once a match knows its RANK inside its window:
//   rank(m)   = number of the window's matches that sort before m            (list walk, early exit)
//   best(pos) = min over records of (combined, window start, rank)           (two 64-bit atomicMin
//               phases: the exact float64 `combined` first, the insertion order among equals second)
// and the winning records are compacted in position order by a prefix sum over the token positions.
// Only the winning rows (32 bytes each) cross PCIe.
#include "common.cuh"

namespace fs {

namespace {

constexpr int kLevThreads = 32;
constexpr int kLevMaxPattern = 250;  // code points of the shorter string handled in shared memory (48 KB per block)
// per thread: uint32 pattern [250] first (4-byte aligned), then the uint16 row [251], padded to a multiple of 4
constexpr int kLevSmemPerThread = (kLevMaxPattern * 4 + (kLevMaxPattern + 1) * 2 + 3) / 4 * 4;

__device__ __forceinline__ unsigned long long orderable(double x) {
    x += 0.0;  // -0.0 -> +0.0: `a < b` does not tell them apart either
    const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(x));
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// candidate order inside a window: distance, then first LSH table, then script position
__device__ __forceinline__ bool sorts_before(const fs_match& a, const fs_match& b, int lsh) {
    if (a.distance != b.distance) return a.distance < b.distance;
    if (lsh) {
        const uint32_t ta = (a.flags >> FS_MATCH_LSH_SHIFT) & 0xFFu, tb = (b.flags >> FS_MATCH_LSH_SHIFT) & 0xFFu;
        if (ta != tb) return ta < tb;
    }
    return a.script_pos < b.script_pos;
}

__device__ __forceinline__ bool kept_by_lsh(const fs_match& m, int lsh) {
    return !lsh || ((m.flags >> FS_MATCH_LSH_SHIFT) & 0xFFu) != 0;
}

__device__ __forceinline__ int64_t n_matches(const PostParams& p) {
    unsigned long long n = p.counters[FS_CNT_MATCHES];
    if (n > static_cast<unsigned long long>(p.match_cap)) n = p.match_cap;
    return static_cast<int64_t>(n);
}

// sequential UTF-8 decoder over a virtual string made of `n_parts` byte ranges with separators
// (script window: words joined by " "; fan window: "[" + tokens joined by ", " + "]")
struct WindowText {
    const uint8_t* base;
    int64_t start[8];
    int32_t len[8];
    int32_t n_parts;
    bool brackets;  // fan form "[a, b]" (separator ", "), else script form "a b"

    // byte at virtual offset i, or -1 past the end; `part`/`off` are the cursor
    int32_t part, off, phase;  // phase: 0 opening bracket, 1 inside part, 2 separator byte 0, 3 separator byte 1, 4 closing, 5 end
    __device__ void rewind() {
        part = 0;
        off = 0;
        phase = brackets ? 0 : 1;
        settle();
    }
    __device__ void settle() {  // skip empty parts
        while (phase == 1 && off >= len[part]) {
            if (part + 1 < n_parts) {
                phase = 2;
                return;
            }
            phase = brackets ? 4 : 5;
            return;
        }
    }
    __device__ int32_t next_byte() {
        switch (phase) {
            case 0:
                phase = 1;
                settle();
                return '[';
            case 1: {
                const int32_t c = base[start[part] + off];
                ++off;
                settle();
                return c;
            }
            case 2:
                if (brackets) {
                    phase = 3;
                    return ',';
                }
                ++part;
                off = 0;
                phase = 1;
                settle();
                return ' ';
            case 3:
                ++part;
                off = 0;
                phase = 1;
                settle();
                return ' ';
            case 4:
                phase = 5;
                return ']';
            default:
                return -1;
        }
    }
    __device__ int32_t total_bytes() const {
        int32_t n = brackets ? 2 : 0;
        for (int k = 0; k < n_parts; ++k) n += len[k];
        return n + (n_parts > 1 ? (n_parts - 1) * (brackets ? 2 : 1) : 0);
    }
};

// Next code point of a byte stream, decoded as host_text.cpp does (malformed bytes decode as
// themselves); -1 at the end.  `peeked` buffers the bytes read ahead of a rejected sequence.
struct Utf8Cursor {
    WindowText* t;
    int32_t pending[4];
    int32_t n_pending;
    __device__ void init(WindowText* text) {
        t = text;
        t->rewind();
        n_pending = 0;
    }
    __device__ int32_t byte() {
        if (n_pending > 0) {
            const int32_t b = pending[0];
            for (int k = 1; k < n_pending; ++k) pending[k - 1] = pending[k];
            --n_pending;
            return b;
        }
        return t->next_byte();
    }
    __device__ int32_t next() {
        const int32_t c = byte();
        if (c < 0) return -1;
        int len = 1;
        uint32_t cp = static_cast<uint32_t>(c);
        if (c >= 0xF0 && c < 0xF8) {
            len = 4;
            cp = c & 0x07;
        } else if (c >= 0xE0) {
            len = 3;
            cp = c & 0x0F;
        } else if (c >= 0xC0) {
            len = 2;
            cp = c & 0x1F;
        }
        if (len == 1) return c;
        int32_t cont[3];
        int got = 0;
        bool ok = true;
        for (; got < len - 1; ++got) {
            cont[got] = byte();
            if (cont[got] < 0 || (cont[got] & 0xC0) != 0x80) {
                ok = false;
                ++got;
                break;
            }
        }
        if (ok) {
            for (int k = 0; k < len - 1; ++k) cp = (cp << 6) | (cont[k] & 0x3F);
            return static_cast<int32_t>(cp);
        }
        // not a well-formed sequence: the lead byte stands for itself, the bytes read ahead go back
        int n_back = 0;
        int32_t back[4];
        for (int k = 0; k < got; ++k)
            if (cont[k] >= 0) back[n_back++] = cont[k];
        for (int k = 0; k < n_pending; ++k) back[n_back++] = pending[k];
        for (int k = 0; k < n_back; ++k) pending[k] = back[k];
        n_pending = n_back;
        return c;
    }
};

}  // namespace

// ---- 1. link every (kept) match into the list of its fan window -------------------------------
__global__ void post_link_kernel(const PostParams p) {
    const int64_t n = n_matches(p);
    for (int64_t m = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; m < n;
         m += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const fs_match mt = p.matches[m];
        if (!kept_by_lsh(mt, p.lsh)) {
            p.next[m] = -2;
            continue;
        }
        p.next[m] = atomicExch(p.head + mt.fan_pos, static_cast<int32_t>(m));
    }
}

// ---- 2. rank inside the window, top-k cut, Levenshtein, minimum of `combined` per fan word -------
__global__ void __launch_bounds__(kLevThreads) post_rank_lev_kernel(const PostParams p) {
    extern __shared__ __align__(16) uint8_t lev_smem[];
    uint32_t* pat = reinterpret_cast<uint32_t*>(lev_smem + threadIdx.x * kLevSmemPerThread);
    uint16_t* row = reinterpret_cast<uint16_t*>(lev_smem + threadIdx.x * kLevSmemPerThread + kLevMaxPattern * 4);
    const int64_t n = n_matches(p);
    for (int64_t m = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; m < n;
         m += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        p.m_lev[m] = -1;
        if (p.next[m] == -2) continue;  // dropped by the LSH filter
        const fs_match mt = p.matches[m];
        // rank = matches of the same window that sort before this one (stop at topk: dropped anyway)
        int32_t rank = 0;
        for (int32_t e = p.head[mt.fan_pos]; e >= 0 && rank < p.topk; e = p.next[e])
            if (e != m && sorts_before(p.matches[e], mt, p.lsh)) ++rank;
        if (rank >= p.topk) continue;  // NearestFilter(10)
        if (mt.script_pos < 0 || mt.script_pos + p.window > p.n_script_words) continue;
        WindowText fan, scr;
        fan.base = p.fan_text;
        fan.n_parts = p.window;
        fan.brackets = true;
        scr.base = p.script_text;
        scr.n_parts = p.window;
        scr.brackets = false;
        for (int k = 0; k < p.window; ++k) {
            fan.start[k] = p.tok_start[mt.fan_pos + k];
            fan.len[k] = p.tok_len[mt.fan_pos + k];
            scr.start[k] = p.script_word_off[mt.script_pos + k];
            scr.len[k] = static_cast<int32_t>(p.script_word_off[mt.script_pos + k + 1] - scr.start[k]);
        }
        // a token of 65535+ bytes was clamped by the host encoder; strings whose SHORTER side does not
        // fit the shared-memory row are left to the host (the caller sees FS_OVERFLOW_TEXT)
        bool too_long = false;
        for (int k = 0; k < p.window; ++k) too_long = too_long || fan.len[k] >= 65535;
        // the shorter string (in bytes, an upper bound of its code points) becomes the pattern
        WindowText* ps = scr.total_bytes() <= fan.total_bytes() ? &scr : &fan;
        WindowText* ts = ps == &scr ? &fan : &scr;
        Utf8Cursor cur;
        cur.init(ps);
        int32_t np = 0;
        for (int32_t c; (c = cur.next()) >= 0;) {
            if (np < kLevMaxPattern) pat[np] = static_cast<uint32_t>(c);
            ++np;
        }
        if (np > kLevMaxPattern) too_long = true;
        if (too_long) {
            atomicOr(p.overflow, static_cast<unsigned long long>(FS_OVERFLOW_TEXT));
            continue;
        }
        for (int j = 0; j <= np; ++j) row[j] = static_cast<uint16_t>(j);
        cur.init(ts);
        int32_t i = 0;
        bool text_overflow = false;
        for (int32_t c; (c = cur.next()) >= 0;) {
            ++i;
            if (i >= 65535) {
                text_overflow = true;
                break;
            }
            uint32_t diag = row[0];
            row[0] = static_cast<uint16_t>(i);
            for (int j = 1; j <= np; ++j) {
                const uint32_t up = row[j];
                uint32_t best = diag + (static_cast<uint32_t>(c) != pat[j - 1] ? 1u : 0u);
                best = min(best, up + 1u);
                best = min(best, static_cast<uint32_t>(row[j - 1]) + 1u);
                row[j] = static_cast<uint16_t>(best);
                diag = up;
            }
        }
        if (text_overflow) {
            atomicOr(p.overflow, static_cast<unsigned long long>(FS_OVERFLOW_TEXT));
            continue;
        }
        const int32_t lev = i == 0 ? np : static_cast<int32_t>(row[np]);
        p.m_lev[m] = lev;
        p.m_rank[m] = rank;
        const unsigned long long key = orderable(mt.distance * static_cast<double>(lev));
        for (int k = 0; k < p.window; ++k) atomicMin(p.best_key + mt.fan_pos + k, key);
    }
}

// ---- 3. among the records of minimal `combined`: the first inserted ------------------------------
__global__ void post_tie_kernel(const PostParams p) {
    const int64_t n = n_matches(p);
    for (int64_t m = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; m < n;
         m += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int32_t lev = p.m_lev[m];
        if (lev < 0) continue;
        const fs_match mt = p.matches[m];
        const unsigned long long key = orderable(mt.distance * static_cast<double>(lev));
        // insertion order of the reference: windows in ascending position, candidates in sorted order
        const unsigned long long tie = (static_cast<unsigned long long>(static_cast<uint32_t>(mt.fan_pos)) << 32) |
                                       (static_cast<unsigned long long>(p.m_rank[m]) << 8);
        for (int k = 0; k < p.window; ++k)
            if (p.best_key[mt.fan_pos + k] == key) atomicMin(p.best_tie + mt.fan_pos + k, tie | static_cast<unsigned>(k));
    }
}

__global__ void post_winner_kernel(const PostParams p) {
    const int64_t n = n_matches(p);
    for (int64_t m = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; m < n;
         m += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int32_t lev = p.m_lev[m];
        if (lev < 0) continue;
        const fs_match mt = p.matches[m];
        const unsigned long long key = orderable(mt.distance * static_cast<double>(lev));
        const unsigned long long tie = (static_cast<unsigned long long>(static_cast<uint32_t>(mt.fan_pos)) << 32) |
                                       (static_cast<unsigned long long>(p.m_rank[m]) << 8);
        for (int k = 0; k < p.window; ++k)
            if (p.best_key[mt.fan_pos + k] == key && p.best_tie[mt.fan_pos + k] == (tie | static_cast<unsigned>(k)))
                p.winner[mt.fan_pos + k] = static_cast<int32_t>(m);
    }
}

// ---- 4. rows in position order: block counts, scan of the counts, ordered write -----------------
constexpr int kScanBlock = 1024;

__global__ void __launch_bounds__(kScanBlock) post_count_kernel(const PostParams p) {
    const int64_t pos = static_cast<int64_t>(blockIdx.x) * kScanBlock + threadIdx.x;
    const int has = pos < p.n_tok && p.winner[pos] >= 0;
    const int total = __syncthreads_count(has);
    if (threadIdx.x == 0) p.block_count[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) post_scan_kernel(const PostParams p, int32_t n_blocks) {
    __shared__ int64_t warp_tot[32];
    __shared__ int64_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int32_t base = 0; base < n_blocks; base += 1024) {
        const int32_t b = base + threadIdx.x;
        const int64_t v = b < n_blocks ? p.block_count[b] : 0;
        int64_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
        __syncthreads();
        int64_t before = carry_s;
        for (int w = 0; w < (threadIdx.x >> 5); ++w) before += warp_tot[w];
        if (b < n_blocks) p.block_off[b] = before + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) p.counters[FS_CNT_ROWS] = static_cast<unsigned long long>(carry_s);
}

__global__ void __launch_bounds__(kScanBlock) post_emit_kernel(const PostParams p) {
    __shared__ int32_t warp_tot[kScanBlock / 32];
    const int64_t pos = static_cast<int64_t>(blockIdx.x) * kScanBlock + threadIdx.x;
    const int32_t m = pos < p.n_tok ? p.winner[pos] : -1;
    const unsigned ballot = __ballot_sync(0xffffffffu, m >= 0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_tot[warp] = __popc(ballot);
    __syncthreads();
    int64_t o = p.block_off[blockIdx.x];
    for (int w = 0; w < warp; ++w) o += warp_tot[w];
    o += __popc(ballot & ((1u << lane) - 1u));
    if (m >= 0) {
        if (o < p.rows_cap) {
            const fs_match mt = p.matches[m];
            fs_row r;
            r.work = mt.work;
            r.word = static_cast<int32_t>(pos - __ldg(p.fan_off + mt.work));
            r.window_ix = static_cast<int32_t>(pos - mt.fan_pos);
            r.match_ix = mt.script_pos;
            r.distance = mt.distance;
            r.lev = p.m_lev[m];
            r.reserved = 0;
            p.rows[o] = r;
        } else {
            atomicOr(p.overflow, static_cast<unsigned long long>(FS_OVERFLOW_ROWS));
        }
    }
}

int launch_postprocess(const PostParams& p, int sm_count, cudaStream_t st) {
    if (p.n_tok <= 0) return FS_OK;
    FS_CUDA_CHECK(cudaMemsetAsync(p.head, 0xFF, sizeof(int32_t) * p.n_tok, st));
    FS_CUDA_CHECK(cudaMemsetAsync(p.winner, 0xFF, sizeof(int32_t) * p.n_tok, st));
    FS_CUDA_CHECK(cudaMemsetAsync(p.best_key, 0xFF, sizeof(unsigned long long) * p.n_tok, st));
    FS_CUDA_CHECK(cudaMemsetAsync(p.best_tie, 0xFF, sizeof(unsigned long long) * p.n_tok, st));
    const int grid = sm_count * 8;
    post_link_kernel<<<grid, 256, 0, st>>>(p);
    post_rank_lev_kernel<<<sm_count * 16, kLevThreads, kLevThreads * kLevSmemPerThread, st>>>(p);
    post_tie_kernel<<<grid, 256, 0, st>>>(p);
    post_winner_kernel<<<grid, 256, 0, st>>>(p);
    const int32_t n_blocks = static_cast<int32_t>((p.n_tok + kScanBlock - 1) / kScanBlock);
    post_count_kernel<<<n_blocks, kScanBlock, 0, st>>>(p);
    post_scan_kernel<<<1, 1024, 0, st>>>(p, n_blocks);
    post_emit_kernel<<<n_blocks, kScanBlock, 0, st>>>(p);
    FS_CUDA_CHECK(cudaGetLastError());
    return FS_OK;
}

int64_t postprocess_scan_blocks(int64_t n_tok) { return (n_tok + kScanBlock - 1) / kScanBlock; }

}  // namespace fs
