// Window-vs-script distance search on tcgen05 tensor cores.
//
// Replaces the reference's hot loop (search.py:176-184): for every fan window one
// engine.neighbours(row) call = cosine distance against script windows, then
// `distance < distance_threshold`.
//
// Formulation.  A window vector is the concatenation of `w` consecutive token rows
// (search.py:94-95, 170-173), so with E_f [T_f, d_pad] and E_s [T_s, d_pad] (fp16,
// row-major, zero padded) the window matrices are overlapping strided views and
//
//     dot(fanwin_i, scriptwin_j) = sum_{s<w} sum_c  E_f[i+s, c-chunk] . E_s[j+s, c-chunk]
//
// One smem stage holds a 64-column chunk of rows [m0+s0, m0+s0+136) of E_f and rows
// [n0+s0, n0+s0+272) of E_s.  The S shifts served by that stage are plain row offsets
// of the UMMA shared-memory descriptors (row pitch 128 B inside the 128B-swizzled
// tile), so every token row is fetched from L2 once per tile instead of `w` times.
//
// Roles (576 threads, 1 CTA/SM, persistent over a contiguous range of tiles):
//   warps 0..15    : epilogue           (tcgen05.ld, diagonal sum, norm/threshold compare, compaction)
//   warp 16 lane 0 : TMA producer       (cp.async.bulk.tensor, mbarrier stage ring)
//   warp 17 lane 0 : tcgen05.mma issuer (128x256x16, fp32 accumulators in TMEM, 2 buffers)
//
// Epilogue.  A pair survives iff  acc[i][j] > A_i * B_j - C_i * D_j  with (A, C) = (|fanwin_i|, norm of
// its operand rounding error) and (B, D) = ((1-thr-eps)|scriptwin_j| - error, |scriptwin_j| + error),
// NaN for windows that straddle a work/script boundary (window_norm_kernel, embed.cu): a guaranteed
// superset of {cos > 1 - thr}.  Survivors are appended to a global candidate list through an atomic
// cursor; they are re-scored in float64 afterwards.
#include "common.cuh"

#include <mutex>
#include <type_traits>

#ifdef FS_TIMELINE
// Debug build only (FS_NVCC_EXTRA=-DFS_TIMELINE): clock64 stamps of CTA 0, read back with
// fs_debug_timeline; [role][tile][slot], roles 0 producer, 1 MMA issuer, 2 + w epilogue warp w.
constexpr int kTlRoles = 2 + 16, kTlTiles = 64, kTlSlots = 12;
__device__ long long g_timeline[kTlRoles][kTlTiles][kTlSlots];
#define FS_TL(role, tile, slot)                                                                   \
    do {                                                                                          \
        if (blockIdx.x == 0 && (tile) < kTlTiles && (threadIdx.x & 31) == 0)                      \
            g_timeline[role][tile][slot] = clock64();                                             \
    } while (0)
extern "C" int fs_debug_timeline(long long* out) {
    return cudaMemcpyFromSymbol(out, g_timeline, sizeof(g_timeline)) == cudaSuccess ? 0 : -1;
}
#else
#define FS_TL(role, tile, slot) do { } while (0)
#endif

namespace fs {

// ---------------------------------------------------------------------------------------------
// Diagonal sum of one 32-column chunk, in place: on return r[x] (x < 32) holds
//   out[lane][x] = sum_{d<E} acc[lane+d][x+d]
// for lanes < 32-(E-1) (the other lanes' values are unused; their rows are finished from shared
// memory).  r[0..39] holds acc[lane][c0 .. c0+39] on entry.  Row shifts are warp shuffles, the
// scarce resource of this epilogue (~0.5 warp-shuffles/clk/SM measured): with kPack two
// neighbouring partial sums travel as one fp16x2 word, halving the shuffles.  A shuffled partial
// sum is then rounded to fp16 once: |error| <= 2^-11 * (sum of the magnitudes of the shuffled
// terms) <= 2^-10 |f||s| in the worst case, which the pre-filter slack (kEpsPacked) covers --
// the decision itself is always re-made in float64.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float h2_lo(uint32_t u) {
    return __low2float(*reinterpret_cast<const __half2*>(&u));
}
__device__ __forceinline__ float h2_hi(uint32_t u) {
    return __high2float(*reinterpret_cast<const __half2*>(&u));
}

template <int kDiag, int kPack>
__device__ __forceinline__ void diag_sum_inplace(uint32_t (&r)[40]) {
    constexpr uint32_t kFull = 0xffffffffu;
    auto f = [&](int x) { return __uint_as_float(r[x]); };
    if (kDiag == 1) return;
    if (kDiag == 6) {
        if (kPack) {
            // stage 1: D2[x] = a[x] + a[x+1]@(lane+1) for x < 36
#pragma unroll
            for (int k = 0; k < 18; ++k) {
                const uint32_t s1 = __shfl_down_sync(kFull, pack_h2(f(2 * k + 1), f(2 * k + 2)), 1);
                r[2 * k] = __float_as_uint(f(2 * k) + h2_lo(s1));
                r[2 * k + 1] = __float_as_uint(f(2 * k + 1) + h2_hi(s1));
            }
            // stage 2: out[x] = D2[x] + D2[x+2]@(lane+2) + D2[x+4]@(lane+4)
            uint32_t t2[18], t4[18];
#pragma unroll
            for (int k = 1; k < 18; ++k) {
                const uint32_t q = pack_h2(f(2 * k), f(2 * k + 1));
                if (k <= 16) t2[k] = __shfl_down_sync(kFull, q, 2);
                if (k >= 2) t4[k] = __shfl_down_sync(kFull, q, 4);
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                r[2 * k] = __float_as_uint(f(2 * k) + h2_lo(t2[k + 1]) + h2_lo(t4[k + 2]));
                r[2 * k + 1] = __float_as_uint(f(2 * k + 1) + h2_hi(t2[k + 1]) + h2_hi(t4[k + 2]));
            }
        } else {
#pragma unroll
            for (int x = 0; x < 36; ++x)
                r[x] = __float_as_uint(f(x) + __shfl_down_sync(kFull, f(x + 1), 1));
#pragma unroll
            for (int x = 0; x < 32; ++x)
                r[x] = __float_as_uint(f(x) + __shfl_down_sync(kFull, f(x + 2), 2) +
                                       __shfl_down_sync(kFull, f(x + 4), 4));
        }
        return;
    }
    if (kDiag == 3 && kPack) {
        uint32_t s1[16], s2[17];
#pragma unroll
        for (int k = 0; k < 17; ++k) {
            const uint32_t q = pack_h2(f(2 * k + 1), f(2 * k + 2));
            if (k < 16) s1[k] = __shfl_down_sync(kFull, q, 1);
            s2[k] = __shfl_down_sync(kFull, q, 2);
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            // out[2k]   = a[2k]   + a[2k+1]@+1 + a[2k+2]@+2 ; out[2k+1] = a[2k+1] + a[2k+2]@+1 + a[2k+3]@+2
            r[2 * k] = __float_as_uint(f(2 * k) + h2_lo(s1[k]) + h2_hi(s2[k]));
            r[2 * k + 1] = __float_as_uint(f(2 * k + 1) + h2_hi(s1[k]) + h2_lo(s2[k + 1]));
        }
        return;
    }
    // generic: E-1 full-precision shuffles per output
#pragma unroll
    for (int x = 0; x < 32; ++x) {
        float acc = f(x);
#pragma unroll
        for (int d = 1; d < kDiag; ++d) acc += __shfl_down_sync(kFull, f(x + d), d);
        r[x] = __float_as_uint(acc);
    }
}

// E = 6 with the whole diagonal sum in fp16x2 arithmetic (kPack == 2): the accumulators are packed
// to half2 pairs once, all row shifts and additions run on pairs (HADD2), and only the chunk
// maximum is widened again: ~4.7 instead of ~17 instructions per output.  Every packed value
// and every HADD2 result is rounded to fp16: |error| <= 4 * 2^-11 * sum_k |G_k| <= 2^-9 |f||s|
// (G_k = the six partial dots of the window, sum |G_k| <= |f||s|), covered by the pre-filter
// slack; the decision is re-made in float64.  On return o[k] = half2(out[2k], out[2k+1]).
// pk[k] = half2(a[2k], a[2k+1]) for the 40 loaded columns; odd(k) = half2(a[2k+1], a[2k+2]) is a byte
// permute of two neighbours, so the fp32 accumulators are dead once pk is built (the next chunk's
// TMEM load is issued into the same registers while this chunk is summed).
__device__ __forceinline__ uint32_t h2_odd(const uint32_t (&pk)[20], int k) {
    return __byte_perm(pk[k], pk[k + 1], 0x5432);
}
__device__ __forceinline__ uint32_t h2_add(uint32_t a, uint32_t b) {
    const __half2 s = __hadd2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&s);
}

__device__ __forceinline__ float diag6_half(const uint32_t (&pk)[20], uint32_t (&o)[16]) {
    constexpr uint32_t kFull = 0xffffffffu;
    // D2 pair k = (a[2k], a[2k+1]) + (a[2k+1], a[2k+2])@(lane+1),  k = 0..17
    uint32_t d2[18];
#pragma unroll
    for (int k = 0; k < 18; ++k) d2[k] = h2_add(pk[k], __shfl_down_sync(kFull, h2_odd(pk, k), 1));
    // out pair k = D2[k] + D2[k+1]@(lane+2) + D2[k+2]@(lane+4),  k = 0..15
    __half2 mx = __float2half2_rn(-60000.f);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        o[k] = h2_add(h2_add(d2[k], __shfl_down_sync(kFull, d2[k + 1], 2)), __shfl_down_sync(kFull, d2[k + 2], 4));
        mx = __hmax2(mx, *reinterpret_cast<const __half2*>(&o[k]));
    }
    return fmaxf(__low2float(mx), __high2float(mx));
}

// E = 3 in fp16x2 arithmetic (kPack == 2): out pair k = (a[2k], a[2k+1]) + (a[2k+1], a[2k+2])@(lane+1)
// + (a[2k+2], a[2k+3])@(lane+2).  Three roundings per output: |error| <= 3 * 2^-11 * sum_d |a_d|
// <= 1.5e-3 |f||s|, inside the pre-filter slack.
__device__ __forceinline__ float diag3_half(const uint32_t (&pk)[20], uint32_t (&o)[16]) {
    constexpr uint32_t kFull = 0xffffffffu;
    __half2 mx = __float2half2_rn(-60000.f);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const uint32_t q1 = __shfl_down_sync(kFull, h2_odd(pk, k), 1);
        const uint32_t q2 = __shfl_down_sync(kFull, pk[k + 1], 2);
        o[k] = h2_add(h2_add(pk[k], q1), q2);
        mx = __hmax2(mx, *reinterpret_cast<const __half2*>(&o[k]));
    }
    return fmaxf(__low2float(mx), __high2float(mx));
}

// E = 2 in fp16x2 arithmetic: out pair k = (a[2k], a[2k+1]) + (a[2k+1], a[2k+2])@(lane+1): one shuffle
// per two outputs, two roundings per output.
__device__ __forceinline__ float diag2_half(const uint32_t (&pk)[20], uint32_t (&o)[16]) {
    constexpr uint32_t kFull = 0xffffffffu;
    __half2 mx = __float2half2_rn(-60000.f);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        o[k] = h2_add(pk[k], __shfl_down_sync(kFull, h2_odd(pk, k), 1));
        mx = __hmax2(mx, *reinterpret_cast<const __half2*>(&o[k]));
    }
    return fmaxf(__low2float(mx), __high2float(mx));
}

// ---------------------------------------------------------------------------------------------
// Tile schedule of one worker (a CTA, or a CTA pair): a contiguous range of linearised (m, n)
// tiles, n fastest; in pair mode the unit is (pair of consecutive m tiles, n).  All three roles
// walk the same sequence; (m unit, n tile) is advanced incrementally (one 64-bit division in the
// constructor, none per tile).
// ---------------------------------------------------------------------------------------------
struct Tile {
    int32_t m0, n0;
    bool fan_first;  // first / last script tile of the current fan tile (A-resident mode)
    bool fan_last;
};

template <int kDiag, bool kPair, int kBN = kBlockN>
struct TileWalk {
    int64_t left;
    int32_t um, un, tn;
    uint32_t cta_rank;
    bool first;

    __device__ __forceinline__ TileWalk(const DistParams& p, int64_t worker, int64_t n_workers, uint32_t rank)
        : tn(p.tiles_n), cta_rank(rank), first(true) {
        const int64_t units_m = kPair ? (p.tiles_m + 1) / 2 : p.tiles_m;
        const int64_t total = units_m * tn;
        const int64_t per = (total + n_workers - 1) / n_workers;
        const int64_t begin = per * worker;
        const int64_t end = min(total, begin + per);
        left = end > begin ? end - begin : 0;
        um = tn > 0 ? static_cast<int32_t>(begin / tn) : 0;
        un = tn > 0 ? static_cast<int32_t>(begin - static_cast<int64_t>(um) * tn) : 0;
    }
    // first script column of the tile `skip` (0 or 1) positions after the one next() would return, or -1
    __device__ __forceinline__ int32_t peek_n0(int skip) const {
        if (left <= skip) return -1;
        int32_t u = un + skip;
        if (u >= tn) u -= tn;
        return u * (kBN - (kDiag - 1));
    }
    __device__ __forceinline__ bool next(Tile& out) {
        if (left <= 0) return false;
        out.m0 = static_cast<int32_t>(kPair ? 2 * um + cta_rank : um) * dist_m_step(kDiag);
        out.n0 = un * (kBN - (kDiag - 1));
        out.fan_first = first || un == 0;
        --left;
        out.fan_last = left == 0 || un + 1 == tn;
        if (++un == tn) {
            un = 0;
            ++um;
        }
        first = false;
        return true;
    }
};

// ---------------------------------------------------------------------------------------------
// Epilogue of one accumulator tile, run by the 16 epilogue warps:
// warp -> TMEM lane quarter (warp & 3, a hardware restriction) x group of 64 columns.
//
// E > 1: out[i][j] = sum_{d<E} acc[i+d][j+d].  The column shift is a register index; the row
// shift is a warp shuffle for lanes < 32-(E-1).  The last E-1 rows of a quarter need rows of the
// NEXT quarter (another warp): every warp publishes its first and last E-1 rows to shared memory
// while it streams its chunks, and after one barrier per tile those few boundary rows (3(E-1) of
// 128) are summed from shared memory, one column per lane.
// ---------------------------------------------------------------------------------------------
//
// The accumulator is handed back to the MMA issuer (tempty) as soon as its last column has been read
// into registers, BEFORE the boundary pass.  kPack == 2 publishes the boundary rows as halves into
// one of two buffers (`halo` is already the buffer of this accumulator stage), so a single barrier
// per tile is enough: the buffer is next written two tiles later, after every warp has passed the
// following tile's barrier.
template <int kDiag, bool kDump, int kPack, bool kPair>
__device__ __forceinline__ void epilogue_tile(const DistParams& p, const int32_t m0, const int32_t n0,
                                              const int as, const uint32_t tfull_addr, const uint32_t tempty_addr,
                                              const uint32_t aphase, const uint32_t tmem_base, float* halo,
                                              float2* norm_tile, __half* rowmax, const int warp, const int lane,
                                              const int tl_tile = 0) {
    (void)tl_tile;
    constexpr bool kOverlap = kDiag == 6;  // lane quarters hold overlapping fan rows: nothing crosses quarters
    constexpr int kMStep = dist_m_step(kDiag);
    constexpr int kNStep = kBlockN - (kDiag - 1);
    constexpr int kPubSlots = dist_pub_slots(kDiag);
    constexpr int kEdge = kDiag - 1;    // boundary rows per side
    constexpr int kTail0 = 32 - kEdge;  // first lane of the tail rows
    const int quarter = warp & 3;
    const int group = warp >> 2;
    const int row = quarter * (kOverlap ? kQuarterRows6 : 32) + lane;  // fan row inside the tile
    const int epi_tid = warp * 32 + lane;  // 0..511
    // publish slot of this lane: head rows 0..E-2 -> slots 0..E-2, tail rows -> E-1..2E-3
    const int pub_slot = kOverlap ? -1 : (lane < kEdge ? lane : (lane >= kTail0 ? kEdge + lane - kTail0 : -1));
    auto pub_at = [&](int q, int slot) -> float* { return halo + (q * kPubSlots + slot) * kHaloCols; };
    // kPack == 2: the whole diagonal sum runs on fp16x2 pairs, and the boundary rows are published as halves
    constexpr bool kHalf = kPack == 2 && (kDiag == 6 || kDiag == 3 || kDiag == 2);
    const bool early_release = kHalf && (p.group & 2) != 0;
    __half* halo_h = reinterpret_cast<__half*>(halo);
    auto pub_half_at = [&](int q, int slot) -> __half* { return halo_h + (q * kPubSlots + slot) * kHaloCols; };

    const int32_t gi = m0 + row;
    const bool row_ok = kOverlap ? lane < kQuarterRows6 : row < kMStep;
    const float kNaN = __int_as_float(0x7fc00000);
    // (A_i, half2(C_i, G_i)) of this lane's fan window (padded to a tile multiple); NaN = never a candidate
    const float2 ac = row_ok ? __ldg(p.fan_ac + gi) : make_float2(kNaN, kNaN);
    // pre-filter bound of a fan row (a, half2(c, g)) against script-side (B, half2(D, H)): a B - c D - g H
    auto bound_row = [](float a, float cg, const float2& bd) {
        const uint32_t u = __float_as_uint(cg), v = __float_as_uint(bd.y);
        const float2 x = __half22float2(*reinterpret_cast<const __half2*>(&u));
        const float2 y = __half22float2(*reinterpret_cast<const __half2*>(&v));
        return fmaf(-x.y, y.y, fmaf(-x.x, y.x, a * bd.x));
    };
    auto bound_of = [&](float a, const float2& bd) { return bound_row(a, ac.y, bd); };
    // rows whose sum needs another warp's rows are finished in the boundary pass
    const float a_main = (kDiag > 1 && lane >= kTail0) ? kNaN : ac.x;
    // E > 1: (B_j, D_j) of this tile staged once in smem, NaN baked in for the E-1 columns that
    // belong to the next tile
    float2* ns_tile = norm_tile + as * kHaloCols;
    if (kDiag > 1 && !kOverlap && epi_tid < kHaloCols)
        ns_tile[epi_tid] = epi_tid < kNStep ? __ldg(p.script_bd + n0 + epi_tid) : make_float2(kNaN, kNaN);
    mbar_wait_mode(tfull_addr, aphase, p.wait_mode & 15);
    tc_fence_after();
    FS_TL(2 + warp, tl_tile, 1);
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                           static_cast<uint32_t>(as * kBlockN + group * kEpiCols);
    if (kHalf && kDiag == 6 && !kDump && (p.group & 4) != 0) {
        // Both 32-column chunks of this warp in one pass.  All 72 columns are loaded at once and the
        // accumulator stage is handed back immediately; the row maxima are taken on the fp32 values
        // (FMNMX3) and only the two maxima are rounded -- rounding is monotone, so
        // half(max(a)) == max(half(a)) and the bound is bit for bit the one of the per-chunk pass --
        // and the 20 fp32 -> fp16x2 packs of a chunk are spent only on a chunk that survives.
        uint32_t q[72];
        {
            const int halo_off = (group * kEpiCols + 64 < kBlockN) ? 64 : 56;  // (see load_chunk below)
            tmem_ld_32x72(taddr, taddr + halo_off, q);
        }
        float2 mm2[2];
        mm2[0] = __ldg(p.script_mm32 + n0 + group * kEpiCols);
        mm2[1] = __ldg(p.script_mm32 + n0 + group * kEpiCols + 32);
        FS_TL(2 + warp, tl_tile, 4);
        tmem_ld_wait();
        FS_TL(2 + warp, tl_tile, 5);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
            if (kPair)
                mbar_arrive_leader(tempty_addr);
            else
                mbar_arrive(tempty_addr);
        }
        auto f = [&](int i) { return __uint_as_float(q[i]); };
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            const int b = 32 * ch;
            float ma = f(b), mb = f(b + 16), mc = f(b + 24);
#pragma unroll
            for (int k = 1; k < 16; ++k) ma = fmaxf(ma, f(b + k));        // columns 0..15
#pragma unroll
            for (int k = 17; k < 24; ++k) mb = fmaxf(mb, f(b + k));       // columns 16..23
#pragma unroll
            for (int k = 25; k < 40; ++k) mc = fmaxf(mc, f(b + k));       // columns 24..39
            const uint32_t m = pack_h2(fmaxf(ma, mb), fmaxf(mb, mc));
            const uint32_t b01 = h2_add(m, __shfl_down_sync(0xffffffffu, m, 1));
            const uint32_t bsum =
                h2_add(h2_add(b01, __shfl_down_sync(0xffffffffu, b01, 2)), __shfl_down_sync(0xffffffffu, b01, 4));
            const float thr_chunk = bound_of(a_main, mm2[ch]);
            if (!__any_sync(0xffffffffu, h2_lo(bsum) > thr_chunk || h2_hi(bsum) > thr_chunk)) continue;
            if (p.group & 32) {
                // Second level (FS_OPT_TILE_GROUP bit 5), for the few chunks the 24-column bound lets through:
                // the same bound over four spans of 8 outputs -- outputs [8 s, 8 s + 8) only read columns
                // [8 s, 8 s + 13) -- summed with the same association.  On text that shares its frequent words
                // with the script a row maximum over 13 columns is far less often a full token match than one
                // over 24, so most of these chunks stop here (~45 instructions) instead of running the diagonal
                // sum (~300) -- and a warp that does not run it does not fall a tile behind the others.
                float n0 = f(b), n1 = f(b + 8), n2 = f(b + 16), n3 = f(b + 24);
#pragma unroll
                for (int k = 1; k < 13; ++k) {
                    n0 = fmaxf(n0, f(b + k));
                    n1 = fmaxf(n1, f(b + 8 + k));
                    n2 = fmaxf(n2, f(b + 16 + k));
                    n3 = fmaxf(n3, f(b + 24 + k));
                }
                const uint32_t u01 = pack_h2(n0, n1), u23 = pack_h2(n2, n3);
                const uint32_t v01 = h2_add(u01, __shfl_down_sync(0xffffffffu, u01, 1));
                const uint32_t v23 = h2_add(u23, __shfl_down_sync(0xffffffffu, u23, 1));
                const uint32_t s01 =
                    h2_add(h2_add(v01, __shfl_down_sync(0xffffffffu, v01, 2)), __shfl_down_sync(0xffffffffu, v01, 4));
                const uint32_t s23 =
                    h2_add(h2_add(v23, __shfl_down_sync(0xffffffffu, v23, 2)), __shfl_down_sync(0xffffffffu, v23, 4));
                if (!__any_sync(0xffffffffu, h2_lo(s01) > thr_chunk || h2_hi(s01) > thr_chunk ||
                                                 h2_lo(s23) > thr_chunk || h2_hi(s23) > thr_chunk))
                    continue;
            }
            uint32_t pk[20], o16[16];
#pragma unroll
            for (int k = 0; k < 20; ++k) pk[k] = pack_h2(f(b + 2 * k), f(b + 2 * k + 1));
            const float mx = diag6_half(pk, o16);
            if (mx > thr_chunk) {
                const int c0 = group * kEpiCols + b;
                const int32_t gj0 = n0 + c0;
#pragma unroll
                for (int x = 0; x < 32; ++x) {
                    const float2 bd = (c0 + x < kNStep) ? __ldg(p.script_bd + gj0 + x) : make_float2(kNaN, kNaN);
                    const float v = (x & 1) ? h2_hi(o16[x >> 1]) : h2_lo(o16[x >> 1]);
                    if (v > bound_of(a_main, bd)) {
                        const unsigned long long slot = atomicAdd(p.counters + FS_CNT_CANDIDATES, 1ull);
                        if (slot < static_cast<unsigned long long>(p.cand_cap)) {
                            p.cand[slot].fan_pos = gi;
                            p.cand[slot].script_pos = gj0 + x;
                        }
                    }
                }
            }
        }
        FS_TL(2 + warp, tl_tile, 2);
        return;
    }
    uint32_t r[40];
    auto load_chunk = [&](int ch) {
        if (kDiag > 1) {
            // the 8 halo columns of the LAST chunk lie outside the tile: any readable columns do
            // (they only enter outputs >= kNStep, which carry +inf norms)
            const int halo_off = (group * kEpiCols + ch * 32 + 32 < kBlockN) ? 32 : 24;
            tmem_ld_32x40(taddr + ch * 32, taddr + ch * 32 + halo_off, r);
        } else {
            tmem_ld_32x32(taddr + ch * 32, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
        }
    };
    // kHalf: the accumulators are dead once packed, so the load of chunk ch+1 is issued before chunk
    // ch is summed and its latency hides behind the shuffles
    if (kHalf) load_chunk(0);
#pragma unroll 1
    for (int ch = 0; ch < kEpiCols / 32; ++ch) {
        const int c0 = group * kEpiCols + ch * 32;  // first column inside the tile
        __syncwarp();
        // prefetch the chunk's (min B, max D): the latency hides behind the TMEM load.  (Fetching it
        // one tile ahead in the caller changed nothing and cost 8 live registers -> spills.)
        const float2 mm = kDump ? make_float2(0.f, 0.f) : __ldg(p.script_mm32 + n0 + c0);
        if (!kHalf) load_chunk(ch);
        tmem_ld_wait();
        uint32_t pk[20];
        if (kHalf) {
#pragma unroll
            for (int k = 0; k < 20; ++k) pk[k] = pack_h2(__uint_as_float(r[2 * k]), __uint_as_float(r[2 * k + 1]));
            if (ch + 1 < kEpiCols / 32) {
                load_chunk(ch + 1);
            } else if (early_release) {
                // the last accumulator column of this warp is packed: hand the TMEM stage back NOW, so
                // that the next MMAs into it are issued under the sums and tests of this chunk
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (kPair)
                        mbar_arrive_leader(tempty_addr);
                    else
                        mbar_arrive(tempty_addr);
                }
            }
        }
        if (kDiag > 1 && pub_slot >= 0) {
            if (kHalf) {
                uint4* dst = reinterpret_cast<uint4*>(pub_half_at(quarter, pub_slot) + c0);
#pragma unroll
                for (int q = 0; q < 5; ++q) dst[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
            } else {
                uint4* dst = reinterpret_cast<uint4*>(pub_at(quarter, pub_slot) + c0);
#pragma unroll
                for (int q = 0; q < 10; ++q)
                    dst[q] = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
            }
        }
        const int32_t gj0 = n0 + c0;
        // smallest pre-filter bound of the chunk for this lane's row (NaN: row not decided here)
        const float thr_chunk = bound_of(a_main, mm);
        if (kHalf && !kDump) {
            // Cheap rejection before any diagonal sum.  With m(l) = max over the 40 loaded columns of
            // row l,  out[lane][x] = sum_d a[lane+d][x+d] <= sum_d m(lane+d): one HMNMX2 tree and E-1
            // shuffles of a single value instead of 16 (E-1) pair shuffles.  The sum is associated
            // exactly as in diagE_half and fp16 addition is monotone, so bound >= every out[lane][x]
            // bit for bit: a chunk skipped here would also have failed the `mx > thr_chunk` test
            // below -- the candidate list is unchanged.  On unrelated text nearly every chunk of
            // every warp stops here (C2 workload: 99.8 % at E = 3, 95 % at E = 6).
            auto hmax = [](uint32_t a, uint32_t b) {
                const __half2 m = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
                return *reinterpret_cast<const uint32_t*>(&m);
            };
            // Two maxima per row: outputs x < 16 only read columns [0, 21) and outputs x >= 16 only
            // columns [16, 37), so m = half2(max of columns 0..23, max of columns 16..39) bounds the two
            // halves of the chunk separately, in the SAME three shuffles -- on text that shares its
            // frequent words with the script (large single dots scattered over the chunk) far fewer
            // chunks reach the full diagonal sum.
            uint32_t ma = pk[0], mb = pk[8], mc = pk[12];
#pragma unroll
            for (int k = 1; k < 8; ++k) ma = hmax(ma, pk[k]);        // columns 0..15
#pragma unroll
            for (int k = 9; k < 12; ++k) mb = hmax(mb, pk[k]);       // columns 16..23
#pragma unroll
            for (int k = 13; k < 20; ++k) mc = hmax(mc, pk[k]);      // columns 24..39
            const uint32_t m_lo = hmax(ma, mb), m_hi = hmax(mb, mc);
            const uint32_t m = hmax(__byte_perm(m_lo, m_hi, 0x5410), __byte_perm(m_lo, m_hi, 0x7632));
            // boundary rows publish the maximum of the whole row: the boundary pass rejects its chunks
            // the same way
            if (kDiag > 1 && !kOverlap && pub_slot >= 0) {
                const uint32_t m_all = hmax(m, __byte_perm(m, m, 0x1032));
                reinterpret_cast<uint16_t*>(rowmax)[(quarter * kPubSlots + pub_slot) * 8 + (c0 >> 5)] =
                    static_cast<uint16_t>(m_all & 0xffffu);
            }
            uint32_t bsum;
            if (kDiag == 6) {
                const uint32_t b01 = h2_add(m, __shfl_down_sync(0xffffffffu, m, 1));
                bsum = h2_add(h2_add(b01, __shfl_down_sync(0xffffffffu, b01, 2)), __shfl_down_sync(0xffffffffu, b01, 4));
            } else if (kDiag == 3) {
                bsum = h2_add(h2_add(m, __shfl_down_sync(0xffffffffu, m, 1)), __shfl_down_sync(0xffffffffu, m, 2));
            } else {
                bsum = h2_add(m, __shfl_down_sync(0xffffffffu, m, 1));
            }
            if (!__any_sync(0xffffffffu, h2_lo(bsum) > thr_chunk || h2_hi(bsum) > thr_chunk)) continue;
        }
        float mx;
        uint32_t o16[16];
        if (kHalf) {
            // o16[k] = half2(out[2k], out[2k+1])
            mx = kDiag == 6 ? diag6_half(pk, o16) : (kDiag == 3 ? diag3_half(pk, o16) : diag2_half(pk, o16));
        } else {
            diag_sum_inplace<kDiag, kPack>(r);  // r[x] <- out[lane][c0 + x]
            mx = -INFINITY;
            if (!kDump) {
#pragma unroll
                for (int x = 0; x < 32; ++x) mx = fmaxf(mx, __uint_as_float(r[x]));
            }
        }
        auto out_val = [&](int x) {
            if (kHalf) return (x & 1) ? h2_hi(o16[x >> 1]) : h2_lo(o16[x >> 1]);
            return __uint_as_float(r[x]);
        };
        if (kDump) {
#pragma unroll
            for (int x = 0; x < 32; ++x) {
                if (row_ok && (kDiag == 1 || lane < kTail0) && gi < p.n_fan_tok && c0 + x < kNStep &&
                    gj0 + x < p.dump_ld)
                    p.dump[static_cast<int64_t>(gi) * p.dump_ld + gj0 + x] = out_val(x);
            }
        } else {
            // one max over the chunk against the smallest bound of the chunk rejects the chunk;
            // the exact per-element test runs only on the rare survivor
            if (mx > thr_chunk) {
#pragma unroll
                for (int x = 0; x < 32; ++x) {
                    const float2 bd = (c0 + x < kNStep) ? __ldg(p.script_bd + gj0 + x) : make_float2(kNaN, kNaN);
                    if (out_val(x) > bound_of(a_main, bd)) {
                        const unsigned long long slot = atomicAdd(p.counters + FS_CNT_CANDIDATES, 1ull);
                        if (slot < static_cast<unsigned long long>(p.cand_cap)) {
                            p.cand[slot].fan_pos = gi;
                            p.cand[slot].script_pos = gj0 + x;
                        }
                    }
                }
            }
        }
    }
    // every accumulator column of this warp is in registers: release the TMEM stage
    FS_TL(2 + warp, tl_tile, 2);
    if (!early_release) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
            if (kPair)
                mbar_arrive_leader(tempty_addr);  // the leader's MMA waits for both CTAs
            else
                mbar_arrive(tempty_addr);
        }
    }
    if (kDiag > 1 && !kOverlap) {
        // boundary rows: tail rows of quarters 0..2 (quarter 3's belong to the next tile)
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
        float a_l[kEdge > 0 ? kEdge : 1], c_l[kEdge > 0 ? kEdge : 1];  // (kDiag == 1 never gets here)
#pragma unroll
        for (int tr = 0; tr < kEdge; ++tr) {
            a_l[tr] = __shfl_sync(0xffffffffu, ac.x, kTail0 + tr);
            c_l[tr] = __shfl_sync(0xffffffffu, ac.y, kTail0 + tr);  // half2(C, G) of that row
        }
        if (quarter < 3) {
#pragma unroll
            for (int it = 0; it < kEpiCols / 32; ++it) {
                const int cc = group * (kEpiCols / 32) + it;  // 32-column chunk of the tile
                const int c = cc * 32 + lane;
                // kHalf: upper bound of each boundary row's sums over this chunk from the published
                // row maxima (summed in the order of `v` below: fp32 addition is monotone), against
                // the chunk's smallest pre-filter bound -- the same rejection as in the main pass
                float rmx[2 * (kEdge > 0 ? kEdge : 1)];  // row maxima: tail rows of this quarter, head rows of the next
                float2 mm = make_float2(0.f, 0.f);
                if (kHalf && !kDump) {
#pragma unroll
                    for (int s2 = 0; s2 < kEdge; ++s2) {
                        rmx[s2] = __half2float(rowmax[(quarter * kPubSlots + kEdge + s2) * 8 + cc]);
                        rmx[kEdge + s2] = __half2float(rowmax[((quarter + 1) * kPubSlots + s2) * 8 + cc]);
                    }
                    mm = __ldg(p.script_mm32 + n0 + cc * 32);
                }
#pragma unroll
                for (int tr = 0; tr < kEdge; ++tr) {
                    const int L = kTail0 + tr;
                    if (kHalf && !kDump) {
                        // rows L .. L+E-1 = entries tr .. tr+E-1 of rmx, summed in the order of `v`
                        float bsum = 0.f;
#pragma unroll
                        for (int d = 0; d < kDiag; ++d) bsum += rmx[tr + d];
                        if (!(bsum > bound_row(a_l[tr], c_l[tr], mm))) continue;
                    }
                    float v = 0.f;
#pragma unroll
                    for (int d = 0; d < kDiag; ++d) {
                        const int lp = L + d;  // row inside this quarter, or lp-32 of the next
                        if (kHalf) {
                            const __half* src = lp < 32 ? pub_half_at(quarter, kEdge + lp - kTail0)
                                                        : pub_half_at(quarter + 1, lp - 32);
                            v += __half2float(src[c + d]);
                        } else {
                            const float* src = lp < 32 ? pub_at(quarter, kEdge + lp - kTail0)
                                                       : pub_at(quarter + 1, lp - 32);
                            v += src[c + d];
                        }
                    }
                    const int32_t gr = m0 + quarter * 32 + L;
                    if (kDump) {
                        if (gr < p.n_fan_tok && c < kNStep && n0 + c < p.dump_ld)
                            p.dump[static_cast<int64_t>(gr) * p.dump_ld + n0 + c] = v;
                    } else if (v > bound_row(a_l[tr], c_l[tr], ns_tile[c])) {
                        const unsigned long long slot = atomicAdd(p.counters + FS_CNT_CANDIDATES, 1ull);
                        if (slot < static_cast<unsigned long long>(p.cand_cap)) {
                            p.cand[slot].fan_pos = gr;
                            p.cand[slot].script_pos = n0 + c;
                        }
                    }
                }
            }
        }
        // single-buffered rows (fp32 layout): the next tile's chunks overwrite them
        if (!kHalf) asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// One-pass epilogue of E = 6 (fp16x2 sums) with nothing but the accumulator on a warp's critical
// path (FS_OPT_TILE_GROUP bit 4).  A clock64 timeline of the 16-warp epilogue
// (profiles/r02_timeline_*.txt) showed what a tile cost each warp: ~250 clk waiting for (A, C, G) of its
// fan row, ~450 clk for the two chunk bounds (min B, max D|H) -- both L2 round trips, the L1 is all
// shared memory here -- ~30 clk for the TMEM load itself, ~500 clk of arithmetic, and that the MMA
// issuer idled 1000 clk per tile waiting for the accumulator to come back: the epilogue warps, not
// the tensor pipe and not the TMEM port, paced the kernel.  Here
//   * (A, C, G) stay in registers for the whole sweep over the script (the fan tile changes every
//     tiles_n tiles),
//   * the chunk bounds of the NEXT tile are copied into a private shared-memory slot by cp.async while
//     this tile is processed (no register is held across the 72 accumulator registers),
//   * kAlt (bit 3): warps 0..7 drain the even tiles of the CTA and warps 8..15 the odd ones, two
//     64-column passes each, so that one set's arithmetic runs under the other set's loads.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_8(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

template <bool kPair, bool kAlt>
__device__ __forceinline__ void epilogue_onepass_loop(const DistParams& p, TileWalk<6, kPair>& walk,
                                                      const uint32_t tfull0, const uint32_t tempty0,
                                                      const uint32_t tmem_base, float2* mm_stage, const int warp,
                                                      const int lane) {
    constexpr int kNStep = kBlockN - 5;
    constexpr int kPasses = kAlt ? 2 : 1;
    constexpr int kVals = 2 * kPasses;  // chunk bounds of this warp per tile
    const int quarter = warp & 3;
    const int set = kAlt ? warp >> 3 : 0;
    const int group0 = kAlt ? ((warp >> 2) & 1) * 2 : warp >> 2;
    const int row = quarter * kQuarterRows6 + lane;
    const bool lane_ok = lane < kQuarterRows6;
    const float kNaN = __int_as_float(0x7fc00000);
    float2* my = mm_stage + warp * 8;  // two buffers of (up to) four chunk bounds
    const uint32_t my_s = smem_u32(my);
    auto bound_row = [](float a, float cg, const float2& bd) {
        const uint32_t u = __float_as_uint(cg), v = __float_as_uint(bd.y);
        const float2 x = __half22float2(*reinterpret_cast<const __half2*>(&u));
        const float2 y = __half22float2(*reinterpret_cast<const __half2*>(&v));
        return fmaf(-x.y, y.y, fmaf(-x.x, y.x, a * bd.x));
    };
    auto prefetch = [&](int32_t n0, int buf) {
        if (lane < kVals)
            cp_async_8(my_s + static_cast<uint32_t>((buf * 4 + lane) * 8),
                       p.script_mm32 + n0 + (group0 + (lane >> 1)) * kEpiCols + (lane & 1) * 32);
    };
    int buf = 0;
    {
        const int32_t n_first = walk.peek_n0(kAlt ? set : 0);
        if (n_first >= 0) prefetch(n_first, 0);
    }
    uint32_t aphase = 0;
    int as = kAlt ? set : 0;
    int32_t ac_m0 = -1;
    float2 ac = make_float2(kNaN, kNaN);
    Tile tile;
    int tl_tile = -1;
    while (walk.next(tile)) {
        ++tl_tile;
        if (kAlt && (tl_tile & 1) != set) continue;
        FS_TL(2 + warp, tl_tile, 0);
        cp_async_wait_all();  // this tile's bounds have landed (they were requested a tile ago)
        __syncwarp();
        {
            const int32_t n_next = walk.peek_n0(kAlt ? 1 : 0);
            if (n_next >= 0) prefetch(n_next, buf ^ 1);
        }
        const int32_t gi = tile.m0 + row;
        if (tile.m0 != ac_m0) {  // new fan tile: once per sweep over the script
            ac = lane_ok ? __ldg(p.fan_ac + gi) : make_float2(kNaN, kNaN);
            ac_m0 = tile.m0;
        }
        const uint32_t tfull_addr = tfull0 + 8u * as, tempty_addr = tempty0 + 8u * as;
        mbar_wait_mode(tfull_addr, aphase, p.wait_mode & 15);
        tc_fence_after();
        FS_TL(2 + warp, tl_tile, 1);
#pragma unroll
        for (int pass = 0; pass < kPasses; ++pass) {
            const int group = group0 + pass;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                                   static_cast<uint32_t>(as * kBlockN + group * kEpiCols);
            uint32_t q[72];
#ifdef FS_FLOOR_PROBE
            if (!(p.group & 256))  // (probe build: bit 8 = no TMEM load at all, the accumulator is handed straight back)
#endif
            {
                // the 8 halo columns of the LAST group lie outside the tile: any readable columns do
                // (they only enter outputs >= kNStep, which carry NaN bounds)
                const int halo_off = (group * kEpiCols + 64 < kBlockN) ? 64 : 56;
                tmem_ld_32x72(taddr, taddr + halo_off, q);
            }
            float thr_ch[2];
            thr_ch[0] = bound_row(ac.x, ac.y, my[buf * 4 + 2 * pass]);
            thr_ch[1] = bound_row(ac.x, ac.y, my[buf * 4 + 2 * pass + 1]);
            FS_TL(2 + warp, tl_tile, 4 + 3 * pass);
            tmem_ld_wait();
            FS_TL(2 + warp, tl_tile, 5 + 3 * pass);
            if (pass == kPasses - 1) {  // the warp's last accumulator column is in registers
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (kPair)
                        mbar_arrive_leader(tempty_addr);
                    else
                        mbar_arrive(tempty_addr);
                }
            }
            auto f = [&](int i) { return __uint_as_float(q[i]); };
#ifdef FS_FLOOR_PROBE
            if (p.group & 384) continue;  // (probe build: bit 7 = load and release only, no arithmetic)
#endif
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                const int b = 32 * ch;
                float ma = f(b), mb = f(b + 16), mc = f(b + 24);
#pragma unroll
                for (int k = 1; k < 16; ++k) ma = fmaxf(ma, f(b + k));   // columns 0..15
#pragma unroll
                for (int k = 17; k < 24; ++k) mb = fmaxf(mb, f(b + k));  // columns 16..23
#pragma unroll
                for (int k = 25; k < 40; ++k) mc = fmaxf(mc, f(b + k));  // columns 24..39
                const uint32_t m = pack_h2(fmaxf(ma, mb), fmaxf(mb, mc));
                const uint32_t b01 = h2_add(m, __shfl_down_sync(0xffffffffu, m, 1));
                const uint32_t bsum =
                    h2_add(h2_add(b01, __shfl_down_sync(0xffffffffu, b01, 2)), __shfl_down_sync(0xffffffffu, b01, 4));
                const float thr_chunk = thr_ch[ch];
                if (!__any_sync(0xffffffffu, h2_lo(bsum) > thr_chunk || h2_hi(bsum) > thr_chunk)) continue;
                uint32_t pk[20], o16[16];
#pragma unroll
                for (int k = 0; k < 20; ++k) pk[k] = pack_h2(f(b + 2 * k), f(b + 2 * k + 1));
                const float mx = diag6_half(pk, o16);
                if (mx > thr_chunk) {
                    const int c0 = group * kEpiCols + b;
                    const int32_t gj0 = tile.n0 + c0;
#pragma unroll
                    for (int x = 0; x < 32; ++x) {
                        const float2 bd = (c0 + x < kNStep) ? __ldg(p.script_bd + gj0 + x) : make_float2(kNaN, kNaN);
                        const float v = (x & 1) ? h2_hi(o16[x >> 1]) : h2_lo(o16[x >> 1]);
                        if (v > bound_row(ac.x, ac.y, bd)) {
                            const unsigned long long slot = atomicAdd(p.counters + FS_CNT_CANDIDATES, 1ull);
                            if (slot < static_cast<unsigned long long>(p.cand_cap)) {
                                p.cand[slot].fan_pos = gi;
                                p.cand[slot].script_pos = gj0 + x;
                            }
                        }
                    }
                }
            }
        }
        FS_TL(2 + warp, tl_tile, 2);
        buf ^= 1;
        if (kAlt) {
            aphase ^= 1u;  // the set always drains accumulator stage `set`
        } else if (++as == kAccumStages) {
            as = 0;
            aphase ^= 1u;
        }
        FS_TL(2 + warp, tl_tile, 3);
    }
    cp_async_wait_all();
}

// kDiag = E: the MMAs accumulate only the shifts {0, E, 2E, ...} (w/E of them) and the epilogue
// adds E diagonal neighbours, out[i][j] = sum_{d<E} acc[i+d][j+d].  E = 1 is the plain dense
// contraction.  E > 1 re-uses every partial sum for E windows (w/E times fewer tensor-core
// flops for bit-for-bit the same set of products, summed in fp32); tiles then overlap by E-1
// rows/columns (step 128-(E-1) x 256-(E-1)).
//
// kPair: two CTAs of a cluster (one TPC) run ONE tcgen05.mma.cta_group::2 of M = 256: each CTA
// owns its own 128-window fan tile (and the TMEM accumulator for it) but stages only HALF of
// the script tile, so shared-memory operand reads drop from 96 to 64 B/clk/SM and the smem fill
// from 52 to 35 KB per stage -- the shared-memory pipe, not the tensor pipe, is what saturates
// first once E > 1.  The leader CTA (cluster rank 0) issues the MMAs; TMA bytes of both CTAs are
// accounted on the leader's `full` barrier; tcgen05.commit multicasts to both CTAs' barriers.
//
// kARes (pair mode, d_pad <= 320): the fan tile stays RESIDENT in shared memory while the CTA
// sweeps the script tiles, so only the script half-tile streams (87 instead of 174 KB per
// tile).  With E >= 3 the L2 -> SM traffic (~6 TB/s chip-wide), not the tensor pipe, is the
// wall otherwise.
//
// kF8: the operand rows hold fp8 e4m3 (128 elements per 128-byte chunk row, K = 32 per MMA) and the
// MMAs are tcgen05.mma.kind::f8f6f4 -- byte for byte the same tiles, stages and descriptors as
// fp16; half the operand traffic and half the tensor-pipe work per element of the embedding.
template <int kDiag, bool kDump, bool kPair, bool kARes, int kPack, bool kF8>
__global__ void __launch_bounds__(kDistThreads, 1)
distance_kernel(const __grid_constant__ CUtensorMap map_fan, const __grid_constant__ CUtensorMap map_fan32,
                const __grid_constant__ CUtensorMap map_script, const __grid_constant__ CUtensorMap map_script128,
                const DistParams p) {
    static_assert(!kARes || kPair, "the A-resident variant exists for CTA pairs only");
    static_assert(!kF8 || kPair, "fp8 operands exist for CTA pairs only");
    constexpr bool kOverlap = kDiag == 6;  // overlapping lane quarters (common.cuh)
    extern __shared__ uint8_t smem_raw[];
    constexpr int kNumStages = dist_stages(kDiag, kPair, kARes);
    constexpr int kStageSz = dist_stage_bytes(kPair, kARes);
    const uint32_t smem_a_res = (smem_u32(smem_raw) + 1023u) & ~1023u;  // resident fan tile (kARes)
    const uint32_t smem_base = smem_a_res + (kARes ? kAResBytes : 0);
    // layout: [resident A] [stages x (A | B)] [barriers] [boundary rows] [norm tile]
    const uint32_t bar_base = smem_base + kNumStages * kStageSz;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kNumStages + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kNumStages + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kNumStages + kAccumStages + s); };
    const uint32_t afull_bar = bar_base + 8u * (2 * kNumStages + 2 * kAccumStages);
    const uint32_t aempty_bar = afull_bar + 8u;
    const uint32_t tmem_slot = aempty_bar + 8u;
    uint32_t* tmem_slot_ptr =
        reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    float* halo = reinterpret_cast<float*>(smem_raw + (bar_base + 256u - smem_u32(smem_raw)));
    float2* norm_tile = reinterpret_cast<float2*>(halo + dist_pub_bytes(kDiag) / 4);
    __half* rowmax_base = reinterpret_cast<__half*>(norm_tile + kAccumStages * kHaloCols);
    float2* mm_stage = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(rowmax_base) + dist_rowmax_bytes(kDiag));

    // Warp roles.  The warp scheduler prefers the HIGHEST warp id among eligible warps, so the
    // two single-thread roles that everything else waits for get the two highest ids: with
    // the MMA issuer below the epilogue warps of its scheduler it is starved exactly while the
    // epilogue is busy, and tile t+1's MMAs no longer overlap tile t's epilogue.
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = kPair ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;

    if (warp == kProducerWarp && lane == 0) {
        tma_prefetch_desc(kOverlap ? &map_fan32 : &map_fan);
        tma_prefetch_desc(kOverlap ? &map_script128 : &map_script);
        for (int s = 0; s < kNumStages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < kAccumStages; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), (kPair ? 2 : 1) * (dist_alt_sets(kDiag, kPack, kDump, p.group) ? kEpiWarps / 2 : kEpiWarps));
        }
        mbar_init(afull_bar, 1);
        mbar_init(aempty_bar, 1);
        mbar_fence_init();
    }
    if (warp == kMmaWarp) {
        if (kPair)
            tmem_alloc_pair(tmem_slot, kTmemCols);
        else
            tmem_alloc(tmem_slot, kTmemCols);
    }
    tc_fence_before();
    if (kPair)
        cluster_sync_all();
    else
        __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int64_t n_workers = kPair ? gridDim.x / 2 : gridDim.x;
    const int64_t worker = kPair ? blockIdx.x / 2 : blockIdx.x;
    TileWalk<kDiag, kPair> walk(p, worker, n_workers, cta_rank);
    Tile tile;
    constexpr int kShiftRows = kDiag;                   // token rows between two MMA shifts
    const int S = p.shifts_per_stage;                   // MMA shifts served by one smem stage
    const int shift_groups = (p.window / kDiag) / S;    // stages per 64-column chunk
    // Grouped stages (A-resident, E = 6, narrow embeddings): the chunks of a script tile are one unit of
    // the ring -- one full/empty barrier pair per TILE, so the MMA issuer makes one wait and one elected
    // block of MMAs per tile instead of one per chunk (its instruction stream, ~85 instructions per
    // chunk, was what paced the kernel: profiles/r01_timeline_s2.txt).  The ring starts right behind the
    // resident chunks actually used and holds n_groups tiles.
    const bool grouped = kARes && kOverlap && (p.group & 1) != 0 && p.chunks <= kGroupMaxChunks;
    const uint32_t group_ring = smem_a_res + static_cast<uint32_t>(p.chunks) * kStageABytes;
    const int n_groups_fit = (kAResBytes - p.chunks * kStageABytes + kNumStages * kStageSz) / (p.chunks * kStageSz);
    const int n_groups = n_groups_fit < kNumStages ? n_groups_fit : kNumStages;
    int grp = 0;
    uint32_t gphase = 0;

    // The two control roles run as WHOLE warps (all lanes converged, one elected lane issues):
    // addresses and descriptors then stay warp-uniform and live in uniform registers.  Run by a
    // single diverged lane, every tcgen05.mma costs a ~25-instruction R2UR/ELECT "waterfall",
    // and the issue thread -- not the tensor pipe -- bounds the kernel once E > 1.
    if (warp == kProducerWarp) {
        // ------------------------------------------------------------ TMA producer
        int stage = 0;
        uint32_t phase = 0;
        uint32_t a_phase = 0;
        int tl_tile = -1;
        while (walk.next(tile)) {
            ++tl_tile;
            FS_TL(0, tl_tile, 0);
            const int32_t m0 = tile.m0;
            const int32_t n0 = tile.n0;
            if (kARes && tile.fan_first) {
                // new fan tile: wait until every MMA that read the previous one has retired
                mbar_wait_warp(aempty_bar, a_phase ^ 1u, 64);
                if (elect_one()) {
                    if (leader) mbar_expect_tx(afull_bar, 2 * p.chunks * (kOverlap ? kOverlapABytes : kStageABytes));
                    for (int c = 0; c < p.chunks; ++c) {
                        if (kOverlap) {
                            for (int q = 0; q < 4; ++q)
                                tma_load_2d_pair(smem_a_res + c * kStageABytes + q * (kOverlapBoxRows * 128),
                                                 &map_fan32, afull_bar, c * kChunkK, m0 + q * kQuarterRows6);
                        } else {
                            tma_load_2d_pair(smem_a_res + c * kStageABytes, &map_fan, afull_bar, c * kChunkK, m0);
                        }
                    }
                }
                __syncwarp();
                a_phase ^= 1u;
            }
            if (grouped) {
                // one barrier pair per script tile: all its chunks land on full_bar(grp)
                mbar_wait_warp(empty_bar(grp), gphase ^ 1u, 32);
                FS_TL(0, tl_tile, 1);
                if (elect_one()) {
                    if (leader) mbar_expect_tx(full_bar(grp), 2 * p.chunks * kOverlapBBytes);
                    const uint32_t dst = group_ring + static_cast<uint32_t>(grp * p.chunks) * kStageSz;
                    for (int c = 0; c < p.chunks; ++c)
                        tma_load_2d_pair(dst + c * kStageSz, &map_script128, full_bar(grp), c * kChunkK,
                                         n0 + static_cast<int32_t>(cta_rank) * (kBlockN / 2));
                }
                __syncwarp();
                FS_TL(0, tl_tile, 2);
                if (++grp == n_groups) {
                    grp = 0;
                    gphase ^= 1u;
                }
                continue;
            }
            for (int c = 0; c < p.chunks; ++c) {
                for (int g = 0; g < shift_groups; ++g) {
                    const int32_t s0 = g * S * kShiftRows;  // first token-row shift of this stage
                    mbar_wait_warp(empty_bar(stage), phase ^ 1u, 32);
                    FS_TL(0, tl_tile, 1 + 2 * (c < 5 ? c : 4));
                    const uint32_t a_dst = smem_base + stage * kStageSz;
                    const uint32_t b_dst = kARes ? a_dst : a_dst + kStageABytes;
                    if (elect_one()) {
                        if (kARes) {
                            // only this CTA's half of the script tile streams
                            // (E = 6 reads no row beyond its 128: boxes without the 8 halo rows)
                            if (leader) mbar_expect_tx(full_bar(stage), 2 * (kOverlap ? kOverlapBBytes : kStageABytes));
                            tma_load_2d_pair(b_dst, kOverlap ? &map_script128 : &map_script, full_bar(stage),
                                             c * kChunkK, n0 + s0 + static_cast<int32_t>(cta_rank) * (kBlockN / 2));
                        } else if (kPair) {
                            // this CTA stages its own fan rows and script rows [128 r, 128 r + 136)
                            if (leader)
                                mbar_expect_tx(full_bar(stage),
                                               2 * (kOverlap ? kOverlapABytes + kOverlapBBytes : kPairStageBytes));
                            if (kOverlap) {
                                for (int q = 0; q < 4; ++q)
                                    tma_load_2d_pair(a_dst + q * (kOverlapBoxRows * 128), &map_fan32, full_bar(stage),
                                                     c * kChunkK, m0 + q * kQuarterRows6);
                            } else {
                                tma_load_2d_pair(a_dst, &map_fan, full_bar(stage), c * kChunkK, m0 + s0);
                            }
                            tma_load_2d_pair(b_dst, kOverlap ? &map_script128 : &map_script, full_bar(stage),
                                             c * kChunkK, n0 + s0 + static_cast<int32_t>(cta_rank) * (kBlockN / 2));
                        } else {
                            mbar_expect_tx(full_bar(stage), kOverlap ? kOverlapABytes + 2 * kOverlapBBytes : kStageBytes);
                            if (kOverlap) {
                                for (int q = 0; q < 4; ++q)
                                    tma_load_2d(a_dst + q * (kOverlapBoxRows * 128), &map_fan32, full_bar(stage),
                                                c * kChunkK, m0 + q * kQuarterRows6);
                            } else {
                                tma_load_2d(a_dst, &map_fan, full_bar(stage), c * kChunkK, m0 + s0);
                            }
                            tma_load_2d(b_dst, kOverlap ? &map_script128 : &map_script, full_bar(stage), c * kChunkK,
                                        n0 + s0);
                            tma_load_2d(b_dst + (kOverlap ? kOverlapBBytes : kStageABytes),
                                        kOverlap ? &map_script128 : &map_script, full_bar(stage), c * kChunkK,
                                        n0 + s0 + (kOverlap ? 128 : kBoxRows));
                        }
                    }
                    __syncwarp();
                    FS_TL(0, tl_tile, 2 + 2 * (c < 5 ? c : 4));
                    if (++stage == kNumStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == kMmaWarp && leader) {
        // ------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = umma_idesc_f16(kPair ? 2 * kBlockM : kBlockM, kBlockN);
        int stage = 0;
        uint32_t phase = 0;
        int as = 0;
        uint32_t aphase = 0;
        uint32_t a_phase = 0;
        uint32_t st_src = smem_base;  // shared-memory address of the current stage
        // grouped stages: bit 4c+k set = K-step k of chunk c exists
        uint32_t group_en = 0;
        for (int c = 0; c < p.chunks && c < kGroupMaxChunks; ++c)
            group_en |= ((1u << (c == p.chunks - 1 ? p.last_chunk_ksteps : kChunkK / kUmmaK)) - 1u) << (4 * c);
        int tl_tile = -1;
        while (walk.next(tile)) {
            ++tl_tile;
            FS_TL(1, tl_tile, 0);
            if (kARes && tile.fan_first) {
                mbar_wait_warp(afull_bar, a_phase, 0);  // the resident fan tile has landed (both CTAs)
                a_phase ^= 1u;
            }
            // (every lane polls, no naps: this wait is on the accumulator round trip -- one polling
            // lane napping 20 ns cost 10 %, profiles/r01_sweep_early.jsonl)
            if (p.wait_mode & 16) {
                while (!mbar_try_wait_hint(tempty_bar(as), aphase ^ 1u, 100u)) {
                }
            } else if (p.wait_mode & 32) {
                while (!mbar_test_wait(tempty_bar(as), aphase ^ 1u)) {
                }
            } else {
                mbar_wait_all(tempty_bar(as), aphase ^ 1u);
            }
            tc_fence_after();
            FS_TL(1, tl_tile, 1);
            const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * kBlockN);
            if (kARes && kOverlap && grouped) {
                // whole tile: one wait, one elected block; every operand offset is an immediate
                // relative to the resident tile (A) and to the group's first stage (B)
                constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);
                mbar_wait_all(full_bar(grp), gphase);
                tc_fence_after();
                FS_TL(1, tl_tile, 2);
                const uint32_t a_lo = ((smem_a_res & 0x3FFFFu) >> 4) | (1u << 16);
                const uint32_t b_src = group_ring + static_cast<uint32_t>(grp * p.chunks) * kStageSz;
                const uint32_t b_lo = ((b_src & 0x3FFFFu) >> 4) | (1u << 16);
                if (elect_one()) {
#pragma unroll
                    for (int c = 0; c < kGroupMaxChunks; ++c) {
#pragma unroll
                        for (int k = 0; k < kChunkK / kUmmaK; ++k) {
                            const uint32_t a_off = static_cast<uint32_t>((c * kStageABytes + k * kUmmaK * 2) >> 4);
                            const uint32_t b_off = static_cast<uint32_t>((c * kStageSz + k * kUmmaK * 2) >> 4);
                            if (c == 0 && k == 0)
                                umma_lohi<kPair, kF8>(tmem_d, a_lo, b_lo, kHi, idesc, 0u);
                            else
                                umma_lohi_pair_if<kF8>((group_en >> (4 * c + k)) & 1u, tmem_d, a_lo + a_off,
                                                       b_lo + b_off, kHi, idesc, 1u);
                        }
                    }
                    umma_commit_pair(empty_bar(grp));   // the group's stages are free (both CTAs) ...
                    umma_commit_pair(tfull_bar(as));    // ... and the accumulator tile is complete
                    if (tile.fan_last) umma_commit_pair(aempty_bar);
                }
                __syncwarp();
                FS_TL(1, tl_tile, 3);
                if (++grp == n_groups) {
                    grp = 0;
                    gphase ^= 1u;
                }
                if (++as == kAccumStages) {
                    as = 0;
                    aphase ^= 1u;
                }
                continue;
            }
            const int n_shift = kDiag == 6 ? 1 : S;  // E = 6: the window is one shift
            uint32_t accumulate = 0;
            // This warp is the only issuer of the CTA pair and runs one dependent instruction stream:
            // every instruction it spends per MMA is time the tensor pipe waits (ncu: the warp was
            // busy 100 % of the kernel at ~29 instructions per MMA, the tensor pipe 67 % active).
            // Descriptors are therefore {lo + constant, hi}: a full 128-byte chunk of S shifts is
            // issued as S x 4 MMAs whose operand offsets are immediates.
            constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
            auto desc_lo = [](uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); };
            auto issue_chunk = [&](auto shifts, auto ksteps_c, uint32_t a_lo, uint32_t b_lo) {
                constexpr int kS = decltype(shifts)::value;
                constexpr int kK = decltype(ksteps_c)::value;
#pragma unroll
                for (int s = 0; s < kS; ++s) {
#pragma unroll
                    for (int k = 0; k < kK; ++k) {
                        // row shift s * E rows of 128 B, K-step k * 32 B (the 128B swizzle is a
                        // function of the absolute smem address, so base_offset stays 0)
                        const uint32_t off = static_cast<uint32_t>((s * kShiftRows * 128 + k * kUmmaK * 2) >> 4);
                        umma_lohi<kPair, kF8>(tmem_d, a_lo + off, b_lo + off, kDescHi, idesc, accumulate);
                        accumulate = 1;
                    }
                }
            };
            auto issue_shifts = [&](auto shifts, int ksteps, uint32_t a_lo, uint32_t b_lo) {
                switch (ksteps) {  // the last chunk of a row may hold 1..3 K-steps only
                    case 4: issue_chunk(shifts, std::integral_constant<int, 4>{}, a_lo, b_lo); break;
                    case 3: issue_chunk(shifts, std::integral_constant<int, 3>{}, a_lo, b_lo); break;
                    case 2: issue_chunk(shifts, std::integral_constant<int, 2>{}, a_lo, b_lo); break;
                    default: issue_chunk(shifts, std::integral_constant<int, 1>{}, a_lo, b_lo); break;
                }
            };
            for (int c = 0; c < p.chunks; ++c) {
                // the last chunk may hold fewer than 64 real columns (d_pad is a multiple of
                // 16, not 64): its trailing K-steps are all-zero TMA fill and are skipped
                const bool last_chunk = c == p.chunks - 1;
                const int ksteps = last_chunk ? p.last_chunk_ksteps : kChunkK / kUmmaK;
                for (int g = 0; g < shift_groups; ++g) {
                    mbar_wait_all(full_bar(stage), phase);
                    tc_fence_after();
                    FS_TL(1, tl_tile, 2 + 2 * (c < 4 ? c : 4));
                    // resident mode: the fan chunk sits in the resident tile and the stage holds
                    // only script rows; the stage's first shift is then an offset into the tile
                    const uint32_t a_src = kARes ? smem_a_res + c * kStageABytes +
                                                       static_cast<uint32_t>(g * S * kShiftRows * 128)
                                                 : st_src;
                    const uint32_t b_src = kARes ? st_src : st_src + kStageABytes;
                    const uint32_t a_lo = desc_lo(a_src), b_lo = desc_lo(b_src);
                    if (elect_one()) {
                        if (n_shift == 1) {
                            issue_shifts(std::integral_constant<int, 1>{}, ksteps, a_lo, b_lo);
                        } else if (n_shift == 2) {
                            issue_shifts(std::integral_constant<int, 2>{}, ksteps, a_lo, b_lo);
                        } else if (n_shift == 3) {
                            issue_shifts(std::integral_constant<int, 3>{}, ksteps, a_lo, b_lo);
                        } else {
                            for (int s = 0; s < n_shift; ++s) {
                                for (int k = 0; k < ksteps; ++k) {
                                    const uint32_t off =
                                        static_cast<uint32_t>((s * kShiftRows * 128 + k * kUmmaK * 2) >> 4);
                                    umma_lohi<kPair, kF8>(tmem_d, a_lo + off, b_lo + off, kDescHi, idesc, accumulate);
                                    accumulate = 1;
                                }
                            }
                        }
                        // frees the smem stage (in both CTAs) when these MMAs retire
                        if (kPair)
                            umma_commit_pair(empty_bar(stage));
                        else
                            umma_commit(empty_bar(stage));
                        if (last_chunk && g + 1 == shift_groups) {
                            // accumulator tile complete (in both CTAs)
                            if (kPair)
                                umma_commit_pair(tfull_bar(as));
                            else
                                umma_commit(tfull_bar(as));
                            // last script tile of this fan tile: the resident tile may be replaced
                            // once these MMAs have retired
                            if (kARes && tile.fan_last) umma_commit_pair(aempty_bar);
                        }
                    }
                    __syncwarp();
                    FS_TL(1, tl_tile, 3 + 2 * (c < 4 ? c : 4));
                    accumulate = 1;
                    st_src += kStageSz;
                    if (++stage == kNumStages) {
                        stage = 0;
                        st_src = smem_base;
                        phase ^= 1u;
                    }
                }
            }
            if (++as == kAccumStages) {
                as = 0;
                aphase ^= 1u;
            }
        }
    } else if (warp < kEpiWarps) {
        // ------------------------------------------------------------ epilogue (16 warps)
        int as = 0;
        uint32_t aphase = 0;
        // kPack == 2: two half-precision boundary-row buffers, one per accumulator stage
        constexpr bool kHalfRows = kPack == 2 && (kDiag == 6 || kDiag == 3 || kDiag == 2);
        int tl_tile = -1;
        if (kDiag == 6 && dist_prefetch_epilogue(kDiag, kPack, kDump, p.group)) {
            if (kDiag == 6) {  // (the walk's type)
                auto& walk6 = reinterpret_cast<TileWalk<6, kPair>&>(walk);
                if (dist_alt_sets(kDiag, kPack, kDump, p.group))
                    epilogue_onepass_loop<kPair, true>(p, walk6, tfull_bar(0), tempty_bar(0), tmem_base, mm_stage, warp, lane);
                else
                    epilogue_onepass_loop<kPair, false>(p, walk6, tfull_bar(0), tempty_bar(0), tmem_base, mm_stage, warp, lane);
            }
        } else
        while (walk.next(tile)) {
            ++tl_tile;
            FS_TL(2 + warp, tl_tile, 0);
            float* halo_t = halo + (kHalfRows ? as * (dist_pub_bytes(kDiag) / 8) : 0);
            __half* rowmax_t = rowmax_base + as * (dist_rowmax_bytes(kDiag) / 4);
            epilogue_tile<kDiag, kDump, kPack, kPair>(p, tile.m0, tile.n0, as, tfull_bar(as), tempty_bar(as),
                                                      aphase, tmem_base, halo_t, norm_tile, rowmax_t, warp, lane, tl_tile);
            FS_TL(2 + warp, tl_tile, 3);
            if (++as == kAccumStages) {
                as = 0;
                aphase ^= 1u;
            }
        }
    }

    tc_fence_before();
    if (kPair)
        cluster_sync_all();  // the peer may still arrive on / multicast into this CTA
    else
        __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        if (kPair)
            tmem_dealloc_pair(tmem_base, kTmemCols);
        else
            tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (fn) return fn;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !ptr) return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    return fn;
}

// tensor map over a row-major matrix [rows, dim_pad] of 2-byte units; box = 64 columns (128 B) x box_rows, SW128
int make_token_map(CUtensorMap* map, const void* base, int64_t rows, int32_t dim_pad, int32_t box_rows) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled driver entry point unavailable");
        return FS_E_NODEVICE;
    }
    if (rows < 1) rows = 1;
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(dim_pad), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(dim_pad) * sizeof(__half)};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(kChunkK), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld dim_pad=%d)",
                  static_cast<int>(r), static_cast<long long>(rows), dim_pad);
        return FS_E_CUDA;
    }
    return FS_OK;
}

template <int kDiag, bool kPair, bool kARes, int kPack, bool kF8 = false>
static int launch_distance_t(const CUtensorMap& map_fan, const CUtensorMap& map_fan32,
                             const CUtensorMap& map_script, const CUtensorMap& map_script128, const DistParams& p,
                             int grid, cudaStream_t stream) {
    // The shared-memory attribute belongs to the (function, DEVICE) pair: one bit per device ordinal,
    // under a lock (indexes on several GPUs may live in one process and be driven from several threads).
    {
        static std::mutex mu;
        static uint64_t configured = 0;
        int dev = 0;
        FS_CUDA_CHECK(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lock(mu);
        if (dev >= 64 || !((configured >> dev) & 1ull)) {
            FS_CUDA_CHECK(cudaFuncSetAttribute(distance_kernel<kDiag, false, kPair, kARes, kPack, kF8>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               dist_smem_bytes(kDiag, kPair, kARes)));
            FS_CUDA_CHECK(cudaFuncSetAttribute(distance_kernel<kDiag, true, kPair, kARes, kPack, kF8>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               dist_smem_bytes(kDiag, kPair, kARes)));
            if (dev < 64) configured |= 1ull << dev;
        }
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(kDistThreads);
    cfg.dynamicSmemBytes = dist_smem_bytes(kDiag, kPair, kARes);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kPair ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (p.dump)
        FS_CUDA_CHECK(cudaLaunchKernelEx(&cfg, distance_kernel<kDiag, true, kPair, kARes, kPack, kF8>, map_fan, map_fan32, map_script, map_script128, p));
    else
        FS_CUDA_CHECK(cudaLaunchKernelEx(&cfg, distance_kernel<kDiag, false, kPair, kARes, kPack, kF8>, map_fan, map_fan32, map_script, map_script128, p));
    return FS_OK;
}


// =============================================================================================
// distance_kernel_n128: the default configuration (E = 6, CTA pairs, resident fan tile, grouped
// stages, one-pass fp16x2 epilogue, operand rows of at most two 128-byte chunks) re-cut for TWO
// co-resident CTA pairs per TPC: script tiles of 128 columns, 8 epilogue warps, 256 TMEM columns
// (two accumulator stages of 128) and ~100 KB of shared memory per CTA.
//
// Why.  Timelines and floor probes of the 256-column kernel (profiles/r02_timeline_*.txt,
// r02_sweep_floor_probes.jsonl) showed a pipeline bound by LATENCY, not by any unit: from the
// hand-back of an accumulator stage to its next hand-back pass ~3100 clk (MMA issue, ~1030 clk of
// MMAs, the drain/commit latency of the tensor pipe, barrier wake-ups, the TMEM load, the remote
// arrives), and with two stages a tile takes half of that -- the tile time follows
// (T_mma + 2070) / 2 for 6, 8 and 10 MMAs per tile, the tensor pipe idles a third of the time and
// an epilogue that does nothing is no faster.  All 512 TMEM columns are two stages of 256, so the
// depth can only grow by narrowing the tile: two independent pipelines of 128-column tiles per SM
// keep four accumulator stages in flight on the same tensor pipe.
// =============================================================================================
static bool n128_applies(const DistParams& p);

namespace n128 {
constexpr int kBN = 128;                      // script windows per tile (UMMA N)
constexpr int kEpi = 8;                       // epilogue warps: 4 TMEM lane quarters x 2 groups of 64 columns
constexpr int kThreads = 32 * (kEpi + 2);     // + TMA producer + MMA issuer
constexpr int kStages = 2;                    // accumulator stages of 128 columns
constexpr int kTmem = 256;
constexpr int kMaxChunks = 2;                 // operand rows of <= 256 bytes
constexpr int kABytes = 4 * kOverlapBoxRows * 128;  // 16 KB: one 128-byte chunk of the 128-row fan tile
constexpr int kBBytes = (kBN / 2) * 128;            // 8 KB: this CTA's 64 script rows of one chunk
constexpr int kGroups = 4;                    // script tiles in flight in the ring
__host__ __device__ constexpr int smem_bytes() {
    return 1024 + kMaxChunks * kABytes + kGroups * kMaxChunks * kBBytes + 256 /*barriers*/;
}
static_assert(2 * (smem_bytes() + 1024) <= 233472, "two CTAs per SM must fit");
}  // namespace n128

template <bool kF8>
__global__ void __launch_bounds__(n128::kThreads, 2)
distance_kernel_n128(const __grid_constant__ CUtensorMap map_fan32, const __grid_constant__ CUtensorMap map_script64,
                     const DistParams p) {
    using namespace n128;
    constexpr int kNStep = kBN - 5;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_a = (smem_u32(smem_raw) + 1023u) & ~1023u;   // resident fan tile
    const uint32_t smem_b = smem_a + kMaxChunks * kABytes;           // ring of script tiles
    const uint32_t bar_base = smem_b + kGroups * kMaxChunks * kBBytes;
    auto full_bar = [&](int g) { return bar_base + 8u * g; };
    auto empty_bar = [&](int g) { return bar_base + 8u * (kGroups + g); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kGroups + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kGroups + kStages + s); };
    const uint32_t afull_bar = bar_base + 8u * (2 * kGroups + 2 * kStages);
    const uint32_t aempty_bar = afull_bar + 8u;
    const uint32_t tmem_slot = aempty_bar + 8u;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const bool leader = cta_rank == 0;
    constexpr int kProducerW = kEpi, kMmaW = kEpi + 1;

    if (warp == kProducerW && lane == 0) {
        tma_prefetch_desc(&map_fan32);
        tma_prefetch_desc(&map_script64);
        for (int g = 0; g < kGroups; ++g) {
            mbar_init(full_bar(g), 1);
            mbar_init(empty_bar(g), 1);
        }
        for (int s = 0; s < kStages; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 2 * kEpi);
        }
        mbar_init(afull_bar, 1);
        mbar_init(aempty_bar, 1);
        mbar_fence_init();
    }
    if (warp == kMmaW) tmem_alloc_pair(tmem_slot, kTmem);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    TileWalk<6, true, kBN> walk(p, blockIdx.x / 2, gridDim.x / 2, cta_rank);
    Tile tile;
    const int chunks = p.chunks;

    if (warp == kProducerW) {
        // ------------------------------------------------------------ TMA producer
        int grp = 0;
        uint32_t gphase = 0, a_phase = 0;
        while (walk.next(tile)) {
            if (tile.fan_first) {
                mbar_wait_warp(aempty_bar, a_phase ^ 1u, 64);
                if (p.fan_tok != nullptr) {
                    // fused gather: map_fan32 is the operand-row TABLE; every lane fetches four rows of the
                    // tile by token id (lane -> quarter lane / 8, rows 4 (lane % 8) .. + 3 of its 32-row box)
                    const int q = lane >> 3, g = lane & 7;
                    const int64_t t0 = static_cast<int64_t>(tile.m0) + q * kQuarterRows6 + 4 * g;
                    int32_t id[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int32_t v = (t0 + k < p.n_fan_tok) ? __ldg(p.fan_tok + t0 + k) : -1;
                        id[k] = (v >= 0 && v < p.n_valid_rows) ? v : p.n_table_rows;  // outside the map: zeros
                    }
                    if (lane == 0 && leader) mbar_expect_tx(afull_bar, 2 * chunks * kABytes);
                    __syncwarp();
                    for (int c = 0; c < chunks; ++c)
                        tma_gather4_pair(smem_a + c * kABytes + q * (kOverlapBoxRows * 128) + g * (4 * 128), &map_fan32,
                                         afull_bar, c * kChunkK, id[0], id[1], id[2], id[3]);
                } else if (elect_one()) {
                    if (leader) mbar_expect_tx(afull_bar, 2 * chunks * kABytes);
                    for (int c = 0; c < chunks; ++c)
                        for (int q = 0; q < 4; ++q)
                            tma_load_2d_pair(smem_a + c * kABytes + q * (kOverlapBoxRows * 128), &map_fan32, afull_bar,
                                             c * kChunkK, tile.m0 + q * kQuarterRows6);
                }
                __syncwarp();
                a_phase ^= 1u;
            }
            mbar_wait_warp(empty_bar(grp), gphase ^ 1u, 32);
            if (elect_one()) {
                if (leader) mbar_expect_tx(full_bar(grp), 2 * chunks * kBBytes);
                for (int c = 0; c < chunks; ++c)
                    tma_load_2d_pair(smem_b + (grp * kMaxChunks + c) * kBBytes, &map_script64, full_bar(grp), c * kChunkK,
                                     tile.n0 + static_cast<int32_t>(cta_rank) * (kBN / 2));
            }
            __syncwarp();
            if (++grp == kGroups) {
                grp = 0;
                gphase ^= 1u;
            }
        }
    } else if (warp == kMmaW && leader) {
        // ------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = umma_idesc_f16(2 * kBlockM, kBN);
        constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);
        int grp = 0, as = 0;
        uint32_t gphase = 0, aphase = 0, a_phase = 0;
        uint32_t group_en = 0;  // bit 4c+k: K-step k of chunk c exists
        for (int c = 0; c < chunks; ++c)
            group_en |= ((1u << (c == chunks - 1 ? p.last_chunk_ksteps : kChunkK / kUmmaK)) - 1u) << (4 * c);
        const uint32_t a_lo = ((smem_a & 0x3FFFFu) >> 4) | (1u << 16);
        while (walk.next(tile)) {
            if (tile.fan_first) {
                mbar_wait_warp(afull_bar, a_phase, 0);
                a_phase ^= 1u;
            }
            mbar_wait_all(tempty_bar(as), aphase ^ 1u);
            mbar_wait_all(full_bar(grp), gphase);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * kBN);
            const uint32_t b_src = smem_b + static_cast<uint32_t>(grp * kMaxChunks) * kBBytes;
            const uint32_t b_lo = ((b_src & 0x3FFFFu) >> 4) | (1u << 16);
            if (elect_one()) {
#pragma unroll
                for (int c = 0; c < kMaxChunks; ++c) {
#pragma unroll
                    for (int k = 0; k < kChunkK / kUmmaK; ++k) {
                        const uint32_t a_off = static_cast<uint32_t>((c * kABytes + k * kUmmaK * 2) >> 4);
                        const uint32_t b_off = static_cast<uint32_t>((c * kBBytes + k * kUmmaK * 2) >> 4);
                        if (c == 0 && k == 0)
                            umma_lohi<true, kF8>(tmem_d, a_lo, b_lo, kHi, idesc, 0u);
                        else
                            umma_lohi_pair_if<kF8>((group_en >> (4 * c + k)) & 1u, tmem_d, a_lo + a_off, b_lo + b_off,
                                                   kHi, idesc, 1u);
                    }
                }
                umma_commit_pair(empty_bar(grp));
                umma_commit_pair(tfull_bar(as));
                if (tile.fan_last) umma_commit_pair(aempty_bar);
            }
            __syncwarp();
            if (++grp == kGroups) {
                grp = 0;
                gphase ^= 1u;
            }
            if (++as == kStages) {
                as = 0;
                aphase ^= 1u;
            }
        }
    } else if (warp < kEpi) {
        // ------------------------------------------------------------ epilogue (8 warps)
        const int quarter = warp & 3;
        const int group = warp >> 2;
        const int row = quarter * kQuarterRows6 + lane;
        const bool lane_ok = lane < kQuarterRows6;
        const float kNaN = __int_as_float(0x7fc00000);
        auto bound_row = [](float a, float cg, const float2& bd) {
            const uint32_t u = __float_as_uint(cg), v = __float_as_uint(bd.y);
            const float2 x = __half22float2(*reinterpret_cast<const __half2*>(&u));
            const float2 y = __half22float2(*reinterpret_cast<const __half2*>(&v));
            return fmaf(-x.y, y.y, fmaf(-x.x, y.x, a * bd.x));
        };
        int as = 0;
        uint32_t aphase = 0;
        // (Keeping the fan row's bounds in registers over the sweep and prefetching the chunk bounds with
        // cp.async -- the two L2 round trips of a warp's tile -- was measured here too: 134 M instead of 146 M
        // windows/s, as in the 256-column kernel; profiles/r02_sweep_n128.jsonl.)
        int32_t ac_m0 = -1;
        float2 ac = make_float2(kNaN, kNaN);
        while (walk.next(tile)) {
            const int32_t gi = tile.m0 + row;
            const int32_t n0 = tile.n0;
            if (tile.m0 != ac_m0) {  // new fan tile: the row's bounds stay in registers over the sweep of the script
                ac = lane_ok ? __ldg(p.fan_ac + gi) : make_float2(kNaN, kNaN);
                ac_m0 = tile.m0;
            }
            mbar_wait_mode(tfull_bar(as), aphase, p.wait_mode & 15);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                                   static_cast<uint32_t>(as * kBN + group * kEpiCols);
            uint32_t q[72];
            // the 8 halo columns of the second group lie outside the tile: any readable columns do
            // (they only enter outputs >= kNStep, which carry NaN bounds)
            tmem_ld_32x72(taddr, taddr + (group == 0 ? 64 : 56), q);
            float2 mm2[2];
            mm2[0] = __ldg(p.script_mm32 + n0 + group * kEpiCols);
            mm2[1] = __ldg(p.script_mm32 + n0 + group * kEpiCols + 32);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(tempty_bar(as));
            auto f = [&](int i) { return __uint_as_float(q[i]); };
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                const int b = 32 * ch;
                float ma = f(b), mb = f(b + 16), mc = f(b + 24);
#pragma unroll
                for (int k = 1; k < 16; ++k) ma = fmaxf(ma, f(b + k));   // columns 0..15
#pragma unroll
                for (int k = 17; k < 24; ++k) mb = fmaxf(mb, f(b + k));  // columns 16..23
#pragma unroll
                for (int k = 25; k < 40; ++k) mc = fmaxf(mc, f(b + k));  // columns 24..39
                const uint32_t m = pack_h2(fmaxf(ma, mb), fmaxf(mb, mc));
                const uint32_t b01 = h2_add(m, __shfl_down_sync(0xffffffffu, m, 1));
                const uint32_t bsum =
                    h2_add(h2_add(b01, __shfl_down_sync(0xffffffffu, b01, 2)), __shfl_down_sync(0xffffffffu, b01, 4));
                const float thr_chunk = bound_row(ac.x, ac.y, mm2[ch]);
                if (!__any_sync(0xffffffffu, h2_lo(bsum) > thr_chunk || h2_hi(bsum) > thr_chunk)) continue;
                uint32_t pk[20], o16[16];
#pragma unroll
                for (int k = 0; k < 20; ++k) pk[k] = pack_h2(f(b + 2 * k), f(b + 2 * k + 1));
                const float mx = diag6_half(pk, o16);
                if (mx > thr_chunk) {
                    const int c0 = group * kEpiCols + b;
                    const int32_t gj0 = n0 + c0;
#pragma unroll
                    for (int x = 0; x < 32; ++x) {
                        const float2 bd = (c0 + x < kNStep) ? __ldg(p.script_bd + gj0 + x) : make_float2(kNaN, kNaN);
                        const float v = (x & 1) ? h2_hi(o16[x >> 1]) : h2_lo(o16[x >> 1]);
                        if (v > bound_row(ac.x, ac.y, bd)) {
                            const unsigned long long slot = atomicAdd(p.counters + FS_CNT_CANDIDATES, 1ull);
                            if (slot < static_cast<unsigned long long>(p.cand_cap)) {
                                p.cand[slot].fan_pos = gi;
                                p.cand[slot].script_pos = gj0 + x;
                            }
                        }
                    }
                }
            }
            if (++as == kStages) {
                as = 0;
                aphase ^= 1u;
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == kMmaW) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, kTmem);
    }
}

// usable when: fp8/fp16 operands in CTA pairs, E = 6 with the fp16x2 one-pass epilogue, resident fan
// tile, operand rows of at most two 128-byte chunks, candidate search (no dense dump)
bool distance_uses_n128(const DistParams& p) { return n128_applies(p); }

static bool n128_applies(const DistParams& p) {
    return p.pair && p.ares && p.diag == 6 && p.pack == 2 && !p.dump && (p.group & 5) == 5 && (p.group & 64) != 0 &&
           p.chunks <= n128::kMaxChunks && p.window == 6;
}

template <bool kF8>
static int launch_n128(const CUtensorMap& map_fan32, const CUtensorMap& map_script64, DistParams p, int grid_limit,
                       cudaStream_t stream) {
    p.tiles_n = static_cast<int32_t>((p.n_script_tok + (n128::kBN - 5) - 1) / (n128::kBN - 5));
    const int64_t units_m = (p.tiles_m + 1) / 2;
    const int64_t total = units_m * p.tiles_n;
    if (total <= 0) return FS_OK;
    {
        static std::mutex mu;
        static uint64_t configured = 0;
        int dev = 0;
        FS_CUDA_CHECK(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lock(mu);
        if (dev >= 64 || !((configured >> dev) & 1ull)) {
            FS_CUDA_CHECK(cudaFuncSetAttribute(distance_kernel_n128<kF8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               n128::smem_bytes()));
            if (dev < 64) configured |= 1ull << dev;
        }
    }
    // two co-resident CTA pairs per TPC: twice as many clusters as the 256-column kernel
    const int64_t clusters = grid_limit > 0 ? grid_limit : 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(2 * (total < clusters ? total : clusters)));
    cfg.blockDim = dim3(n128::kThreads);
    cfg.dynamicSmemBytes = n128::smem_bytes();
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    FS_CUDA_CHECK(cudaLaunchKernelEx(&cfg, distance_kernel_n128<kF8>, map_fan32, map_script64, p));
    return FS_OK;
}

int launch_distance(const CUtensorMap& map_fan, const CUtensorMap& map_fan32, const CUtensorMap& map_script,
                    const CUtensorMap& map_script128, const CUtensorMap& map_script64, const DistParams& p,
                    int grid_limit, cudaStream_t stream) {
    if (n128_applies(p)) {
        if (p.f8) return launch_n128<true>(map_fan32, map_script64, p, grid_limit, stream);
        return launch_n128<false>(map_fan32, map_script64, p, grid_limit, stream);
    }
    const int64_t units_m = p.pair ? (p.tiles_m + 1) / 2 : p.tiles_m;
    const int64_t total = units_m * p.tiles_n;
    if (total <= 0) return FS_OK;
    int grid;
    if (p.pair) {
        const int64_t clusters = grid_limit / 2 > 0 ? grid_limit / 2 : 1;
        grid = 2 * static_cast<int>(total < clusters ? total : clusters);
    } else {
        grid = static_cast<int>(total < grid_limit ? total : grid_limit);
    }
    const bool ares = p.pair && p.ares && p.chunks <= kAResChunks;
    const int pack = (p.diag == 6 || p.diag == 3) ? p.pack : (p.diag == 2 && p.pack == 2 ? 2 : 0);
#define FS_LAUNCH(E, PAIR, ARES, PACK) \
    return launch_distance_t<E, PAIR, ARES, PACK>(map_fan, map_fan32, map_script, map_script128, p, grid, stream)
#define FS_LAUNCH_PACK(E, PAIR, ARES)              \
    do {                                           \
        if (pack == 2) FS_LAUNCH(E, PAIR, ARES, 2); \
        if (pack >= 1) FS_LAUNCH(E, PAIR, ARES, 1); \
        FS_LAUNCH(E, PAIR, ARES, 0);               \
    } while (0)
#define FS_LAUNCH_MODE(E)                          \
    do {                                           \
        if (ares) FS_LAUNCH_PACK(E, true, true);   \
        if (p.pair) FS_LAUNCH_PACK(E, true, false); \
        FS_LAUNCH_PACK(E, false, false);           \
    } while (0)
    if (p.f8) {
        // fp8 operands: CTA pairs; the epilogue variants that are the defaults
        if (!p.pair) {
            set_error("fp8 operands need CTA pairs");
            return FS_E_INVALID;
        }
        if (ares && pack == 2 && (p.diag == 3 || p.diag == 6)) {
            // resident fan tile: built for the two default diagonal factors with the fp16x2 epilogue;
            // every other combination streams the fan tile
            if (p.diag == 3) return launch_distance_t<3, true, true, 2, true>(map_fan, map_fan32, map_script, map_script128, p, grid, stream);
            return launch_distance_t<6, true, true, 2, true>(map_fan, map_fan32, map_script, map_script128, p, grid, stream);
        }
        switch (p.diag) {
            case 1: return launch_distance_t<1, true, false, 0, true>(map_fan, map_fan32, map_script, map_script128, p, grid, stream);
            case 2:
                if (pack == 2) return launch_distance_t<2, true, false, 2, true>(map_fan, map_fan32, map_script, map_script128, p, grid, stream);
                return launch_distance_t<2, true, false, 0, true>(map_fan, map_fan32, map_script, map_script128, p, grid, stream);
            case 3:
                if (pack == 2) return launch_distance_t<3, true, false, 2, true>(map_fan, map_fan32, map_script, map_script128, p, grid, stream);
                return launch_distance_t<3, true, false, 1, true>(map_fan, map_fan32, map_script, map_script128, p, grid, stream);
            case 6:
                if (pack == 2) return launch_distance_t<6, true, false, 2, true>(map_fan, map_fan32, map_script, map_script128, p, grid, stream);
                return launch_distance_t<6, true, false, 1, true>(map_fan, map_fan32, map_script, map_script128, p, grid, stream);
            default:
                set_error("unsupported diagonal factor %d", p.diag);
                return FS_E_INVALID;
        }
    }
    switch (p.diag) {
        case 1:
            if (ares) FS_LAUNCH(1, true, true, 0);
            if (p.pair) FS_LAUNCH(1, true, false, 0);
            FS_LAUNCH(1, false, false, 0);
        case 2:
            if (pack == 2) {
                if (ares) FS_LAUNCH(2, true, true, 2);
                if (p.pair) FS_LAUNCH(2, true, false, 2);
                FS_LAUNCH(2, false, false, 2);
            }
            if (ares) FS_LAUNCH(2, true, true, 0);
            if (p.pair) FS_LAUNCH(2, true, false, 0);
            FS_LAUNCH(2, false, false, 0);
        case 3: FS_LAUNCH_MODE(3);
        case 6: FS_LAUNCH_MODE(6);
        default:
            set_error("unsupported diagonal factor %d", p.diag);
            return FS_E_INVALID;
    }
#undef FS_LAUNCH
#undef FS_LAUNCH_PACK
#undef FS_LAUNCH_MODE
}

}  // namespace fs
