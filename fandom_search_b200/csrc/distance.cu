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
// Roles (320 threads, 1 CTA/SM, persistent over a contiguous range of tiles):
//   warps 0..7    : epilogue           (tcgen05.ld, diagonal sum, norm/threshold compare, compaction)
//   warp 8 lane 0 : TMA producer       (cp.async.bulk.tensor, mbarrier stage ring)
//   warp 9 lane 0 : tcgen05.mma issuer (128x256x16, fp32 accumulators in TMEM, 2 buffers)
//
// Epilogue.  acc[i][j] > thr_fan[i] * norm_script[j]  <=>  cos > 1 - thr - eps, with
// thr_fan = (1-thr-eps)*|fanwin_i| (+inf for windows that straddle a work boundary) and
// norm_script = |scriptwin_j| (+inf for invalid).  Survivors are appended to a global
// candidate list through an atomic cursor; they are re-scored in float64 afterwards.
#include "common.cuh"

namespace fs {

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
template <int kDiag, bool kDump, bool kPair>
__global__ void __launch_bounds__(kDistThreads, 1)
distance_kernel(const __grid_constant__ CUtensorMap map_fan,
                const __grid_constant__ CUtensorMap map_script, const DistParams p) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int kNumStages = kPair ? kPairStages : kStages;
    constexpr int kStageSz = kPair ? kPairStageBytes : kStageBytes;
    static_assert(kNumStages * kStageSz == kStages * kStageBytes, "stage ring must fill the same smem");
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    // layout: [stages x (A | B)] [barriers] [halo rows] [zero row] [norm tile]
    const uint32_t bar_base = smem_base + kNumStages * kStageSz;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kNumStages + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kNumStages + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kNumStages + kAccumStages + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * kNumStages + 2 * kAccumStages);
    uint32_t* tmem_slot_ptr =
        reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    float* halo = reinterpret_cast<float*>(smem_raw + (bar_base + 256u - smem_u32(smem_raw)));
    float* zero_row = halo + kHaloBytes / 4;
    float* norm_tile = zero_row + kZeroRowBytes / 4;
    for (int i = threadIdx.x; i < kHaloCols; i += blockDim.x) zero_row[i] = 0.f;

    // Warp roles.  The warp scheduler prefers the HIGHEST warp id among eligible warps, so the
    // two single-thread roles that everything else waits for get the two highest ids: with
    // the MMA issuer below the epilogue warps of its scheduler it is starved exactly while the
    // epilogue is busy, and tile t+1's MMAs no longer overlap tile t's epilogue.
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr int kProducerWarp = kEpiWarps;      // 8
    constexpr int kMmaWarp = kEpiWarps + 1;       // 9
    const uint32_t cta_rank = kPair ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;

    if (warp == kProducerWarp && lane == 0) {
        tma_prefetch_desc(&map_fan);
        tma_prefetch_desc(&map_script);
        for (int s = 0; s < kNumStages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < kAccumStages; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), kPair ? 2 * kEpiWarps : kEpiWarps);
        }
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

    // Work units: a contiguous range of linearised (m, n) tiles (n fastest) per CTA; in pair
    // mode per cluster, the unit being (pair of consecutive m tiles, n).
    constexpr int kMStep = kBlockM - (kDiag - 1);
    constexpr int kNStep = kBlockN - (kDiag - 1);
    const int64_t units_m = kPair ? (p.tiles_m + 1) / 2 : p.tiles_m;
    const int64_t total_tiles = units_m * p.tiles_n;
    const int64_t n_workers = kPair ? gridDim.x / 2 : gridDim.x;
    const int64_t worker = kPair ? blockIdx.x / 2 : blockIdx.x;
    const int64_t per_cta = (total_tiles + n_workers - 1) / n_workers;
    const int64_t tile_begin = per_cta * worker;
    const int64_t tile_end = min(total_tiles, tile_begin + per_cta);
    auto tile_m0 = [&](int64_t t) {
        const int64_t um = t / p.tiles_n;
        return static_cast<int32_t>((kPair ? 2 * um + cta_rank : um) * kMStep);
    };
    auto tile_n0 = [&](int64_t t) { return static_cast<int32_t>(t % p.tiles_n) * kNStep; };
    const int S = p.shifts_per_stage;                 // MMA shifts served by one smem stage
    const int shift_groups = (p.window / kDiag) / S;  // stages per 64-column chunk
    const int stages_per_tile = p.chunks * shift_groups;

    if (warp == kProducerWarp && lane == 0) {
        // ------------------------------------------------------------ TMA producer
        int stage = 0;
        uint32_t phase = 0;
        for (int64_t t = tile_begin; t < tile_end; ++t) {
            const int32_t m0 = tile_m0(t);
            const int32_t n0 = tile_n0(t);
            for (int c = 0; c < p.chunks; ++c) {
                for (int g = 0; g < shift_groups; ++g) {
                    const int32_t s0 = g * S * kDiag;  // first token-row shift of this stage
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t a_dst = smem_base + stage * kStageSz;
                    const uint32_t b_dst = a_dst + kStageABytes;
                    if (kPair) {
                        // this CTA stages its own fan rows and script rows [128 r, 128 r + 136)
                        if (leader) mbar_expect_tx(full_bar(stage), 2 * kPairStageBytes);
                        tma_load_2d_pair(a_dst, &map_fan, full_bar(stage), c * kChunkK, m0 + s0);
                        tma_load_2d_pair(b_dst, &map_script, full_bar(stage), c * kChunkK,
                                         n0 + s0 + static_cast<int32_t>(cta_rank) * (kBlockN / 2));
                    } else {
                        mbar_expect_tx(full_bar(stage), kStageBytes);
                        tma_load_2d(a_dst, &map_fan, full_bar(stage), c * kChunkK, m0 + s0);
                        tma_load_2d(b_dst, &map_script, full_bar(stage), c * kChunkK, n0 + s0);
                        tma_load_2d(b_dst + kStageABytes, &map_script, full_bar(stage), c * kChunkK,
                                    n0 + s0 + kBoxRows);
                    }
                    if (++stage == kNumStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == kMmaWarp && lane == 0 && leader) {
        // ------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = umma_idesc_f16(kPair ? 2 * kBlockM : kBlockM, kBlockN);
        int stage = 0;
        uint32_t phase = 0;
        int as = 0;
        uint32_t aphase = 0;
        for (int64_t t = tile_begin; t < tile_end; ++t) {
            mbar_wait(tempty_bar(as), aphase ^ 1u);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * kBlockN);
            uint32_t accumulate = 0;
            for (int it = 0; it < stages_per_tile; ++it) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const uint32_t a_src = smem_base + stage * kStageSz;
                const uint32_t b_src = a_src + kStageABytes;
                // the last chunk may hold fewer than 64 real columns (d_pad is a multiple of
                // 16, not 64): its trailing K-steps are all-zero TMA fill and are skipped
                const int chunk = it / shift_groups;
                const int ksteps = (chunk == p.chunks - 1) ? p.last_chunk_ksteps : kChunkK / kUmmaK;
                for (int s = 0; s < S; ++s) {
                    // row shift inside the stage = s * kDiag rows of 128 B: a plain offset of
                    // the descriptor start address (swizzle is a function of the absolute
                    // address, so base_offset stays 0)
                    const uint32_t row_off = static_cast<uint32_t>(s * kDiag * 128);
                    const uint32_t bo = p.base_offset_mode ? static_cast<uint32_t>(s * kDiag) & 7u : 0u;
                    for (int k = 0; k < ksteps; ++k) {
                        const uint32_t off = row_off + static_cast<uint32_t>(k * kUmmaK * 2);
                        if (kPair)
                            umma_f16_pair(tmem_d, umma_smem_desc(a_src + off, bo),
                                          umma_smem_desc(b_src + off, bo), idesc, accumulate);
                        else
                            umma_f16(tmem_d, umma_smem_desc(a_src + off, bo),
                                     umma_smem_desc(b_src + off, bo), idesc, accumulate);
                        accumulate = 1;
                    }
                }
                // frees the smem stage (in both CTAs) when these MMAs retire
                if (kPair)
                    umma_commit_pair(empty_bar(stage));
                else
                    umma_commit(empty_bar(stage));
                if (++stage == kNumStages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
            // accumulator tile complete (in both CTAs)
            if (kPair)
                umma_commit_pair(tfull_bar(as));
            else
                umma_commit(tfull_bar(as));
            if (++as == kAccumStages) {
                as = 0;
                aphase ^= 1u;
            }
        }
    } else if (warp < kEpiWarps) {
        // ------------------------------------------------------------ epilogue (8 warps)
        // warp -> TMEM lane quarter (warp & 3, a hardware restriction) x column half
        const int quarter = warp & 3;
        const int half = warp >> 2;
        const int row = quarter * 32 + lane;
        const int epi_tid = warp * 32 + lane;  // 0..255
        int as = 0;
        uint32_t aphase = 0;
        for (int64_t t = tile_begin; t < tile_end; ++t) {
            const int32_t m0 = tile_m0(t);
            const int32_t n0 = tile_n0(t);
            const int32_t gi = m0 + row;
            const bool row_ok = row < kMStep;
            const float thr = row_ok ? __ldg(p.thr_fan + gi) : INFINITY;  // padded to a tile multiple
            // E > 1: script-window norms of this tile staged once in smem, +inf baked in for the
            // E-1 columns that belong to the next tile (published by the first chunk barrier)
            float* ns_tile = norm_tile + as * kHaloCols;
            if (kDiag > 1) {
                ns_tile[epi_tid] = epi_tid < kNStep ? __ldg(p.norm_script + n0 + epi_tid) : INFINITY;
                if (epi_tid < kHaloCols - kBlockN) ns_tile[kBlockN + epi_tid] = INFINITY;
            }
            // rows i+d of the last lanes live in the NEXT lane quarter: its warps publish their
            // first E-1 rows chunk by chunk (quarter 3 has no successor inside the tile; those
            // outputs belong to the next tile and read a row of zeros)
            float* pub_row = halo + (((as * 4 + quarter) * kHaloRows + lane) * kHaloCols);
            const float* edge_row[kDiag];  // [d]: row (lane + d - 32) of the next quarter
            bool edge[kDiag];
#pragma unroll
            for (int d = 1; d < kDiag; ++d) {
                edge[d] = lane + d >= 32;
                edge_row[d] = (quarter < 3 && edge[d])
                                  ? halo + (((as * 4 + quarter + 1) * kHaloRows + (lane + d - 32)) * kHaloCols)
                                  : zero_row;
            }
            mbar_wait(tfull_bar(as), aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                                   static_cast<uint32_t>(as * kBlockN + half * (kBlockN / 2));
#pragma unroll 1
            for (int ch = 0; ch < 4; ++ch) {
                const int c0 = half * (kBlockN / 2) + ch * 32;  // first column inside the tile
                uint32_t r[40];
                __syncwarp();
                tmem_ld_32x32(taddr + ch * 32, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
                if (kDiag > 1) {
                    if (c0 + 32 < kBlockN) {
                        tmem_ld_32x8(taddr + ch * 32 + 32, *reinterpret_cast<uint32_t(*)[8]>(&r[32]));
                    } else {
#pragma unroll
                        for (int x = 32; x < 40; ++x) r[x] = 0u;
                    }
                }
                tmem_ld_wait();
                float v[32];
                if (kDiag > 1) {
                    if (lane < kDiag - 1) {
                        uint4* dst = reinterpret_cast<uint4*>(pub_row + c0);
#pragma unroll
                        for (int q = 0; q < 10; ++q)
                            dst[q] = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
                    }
                    asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
                    float h[kDiag][40];  // [d]: the 40 columns of the successor row, edge lanes only
#pragma unroll
                    for (int d = 1; d < kDiag; ++d) {
                        if (edge[d]) {
                            const float4* src = reinterpret_cast<const float4*>(edge_row[d] + c0);
#pragma unroll
                            for (int q = 0; q < 10; ++q) {
                                const float4 f = src[q];
                                h[d][4 * q] = f.x;
                                h[d][4 * q + 1] = f.y;
                                h[d][4 * q + 2] = f.z;
                                h[d][4 * q + 3] = f.w;
                            }
                        }
                    }
#pragma unroll
                    for (int x = 0; x < 32; ++x) {
                        float acc = __uint_as_float(r[x]);
#pragma unroll
                        for (int d = 1; d < kDiag; ++d) {
                            float o = __shfl_down_sync(0xffffffffu, __uint_as_float(r[x + d]), d);
                            if (edge[d]) o = h[d][x + d];
                            acc += o;
                        }
                        v[x] = acc;
                    }
                } else {
#pragma unroll
                    for (int x = 0; x < 32; ++x) v[x] = __uint_as_float(r[x]);
                }
                const int32_t gj0 = n0 + c0;
                if (kDump) {
                    if (row_ok && gi < p.n_fan_tok) {
#pragma unroll
                        for (int x = 0; x < 32; ++x) {
                            if (c0 + x < kNStep && gj0 + x < p.dump_ld)
                                p.dump[static_cast<int64_t>(gi) * p.dump_ld + gj0 + x] = v[x];
                        }
                    }
                } else {
                    bool any = false;
                    const float4* ns4 = (kDiag == 1)
                                            ? reinterpret_cast<const float4*>(p.norm_script + gj0)
                                            : reinterpret_cast<const float4*>(ns_tile + c0);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 n4 = (kDiag == 1) ? __ldg(ns4 + q) : ns4[q];
                        any |= v[4 * q + 0] > thr * n4.x;
                        any |= v[4 * q + 1] > thr * n4.y;
                        any |= v[4 * q + 2] > thr * n4.z;
                        any |= v[4 * q + 3] > thr * n4.w;
                    }
                    if (any) {  // rare: hits are sparse
#pragma unroll
                        for (int x = 0; x < 32; ++x) {
                            const float nsv = (kDiag == 1) ? __ldg(p.norm_script + gj0 + x) : ns_tile[c0 + x];
                            if (v[x] > thr * nsv) {
                                const unsigned long long slot =
                                    atomicAdd(p.counters + FS_CNT_CANDIDATES, 1ull);
                                if (slot < static_cast<unsigned long long>(p.cand_cap)) {
                                    p.cand[slot].fan_pos = gi;
                                    p.cand[slot].script_pos = gj0 + x;
                                }
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (kPair)
                    mbar_arrive_leader(tempty_bar(as));  // the leader's MMA waits for both CTAs
                else
                    mbar_arrive(tempty_bar(as));
            }
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

// tensor map over a row-major fp16 matrix [rows, dim_pad]; box = 64 columns x 136 rows, SW128
int make_token_map(CUtensorMap* map, const void* base, int64_t rows, int32_t dim_pad) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled driver entry point unavailable");
        return FS_E_NODEVICE;
    }
    if (rows < 1) rows = 1;
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(dim_pad), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(dim_pad) * sizeof(__half)};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(kChunkK), static_cast<cuuint32_t>(kBoxRows)};
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

template <int kDiag, bool kPair>
static int launch_distance_t(const CUtensorMap& map_fan, const CUtensorMap& map_script,
                             const DistParams& p, int grid, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        FS_CUDA_CHECK(cudaFuncSetAttribute(distance_kernel<kDiag, false, kPair>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kDistSmemBytes));
        FS_CUDA_CHECK(cudaFuncSetAttribute(distance_kernel<kDiag, true, kPair>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kDistSmemBytes));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(kDistThreads);
    cfg.dynamicSmemBytes = kDistSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kPair ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (p.dump)
        FS_CUDA_CHECK(cudaLaunchKernelEx(&cfg, distance_kernel<kDiag, true, kPair>, map_fan, map_script, p));
    else
        FS_CUDA_CHECK(cudaLaunchKernelEx(&cfg, distance_kernel<kDiag, false, kPair>, map_fan, map_script, p));
    return FS_OK;
}

int launch_distance(const CUtensorMap& map_fan, const CUtensorMap& map_script, const DistParams& p,
                    int grid_limit, cudaStream_t stream) {
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
    const int key = p.diag * 2 + (p.pair ? 1 : 0);
    switch (key) {
        case 2: return launch_distance_t<1, false>(map_fan, map_script, p, grid, stream);
        case 3: return launch_distance_t<1, true>(map_fan, map_script, p, grid, stream);
        case 4: return launch_distance_t<2, false>(map_fan, map_script, p, grid, stream);
        case 5: return launch_distance_t<2, true>(map_fan, map_script, p, grid, stream);
        case 6: return launch_distance_t<3, false>(map_fan, map_script, p, grid, stream);
        case 7: return launch_distance_t<3, true>(map_fan, map_script, p, grid, stream);
        default:
            set_error("unsupported diagonal factor %d", p.diag);
            return FS_E_INVALID;
    }
}

}  // namespace fs
