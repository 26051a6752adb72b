// Token gather and window norms.
//
// Replaces mk_vectors (search.py:65-84) and the rolling-window build
// (search.py:94-95, 169-173).  The reference materialises a float64 [T-5, 1800]
// window matrix per work; here only the per-token matrix E [T, dim_pad] (fp16) is
// written and a window is the strided view E[i : i+w, :] that the distance kernel's
// TMA loads address directly.  HBM-bound: 4 B read + dim_pad*2 B written per token
// (table rows are L2 hits).
#include "common.cuh"
#include <cuda_fp8.h>

namespace fs {

// fp32 rows -> scaled operand rows (fp16, or fp8 e4m3 when kF8) padded to the row length, plus per
// row (squared norm of the scaled fp32 row, squared norm of the rounding error of its KEPT elements,
// squared norm of its DROPPED elements); one warp per row.  Used once for the base table and per
// batch for OOV extras.
//
// Kept / dropped elements (pre-filter dimensions): operand element c is source column perm[c]; only
// the first `kept` columns of that order enter the operand row, the remaining dim - kept are left
// to the bound  |f_drop . s_drop| <= |f_drop| |s_drop|  of the distance epilogue (window_norm_kernel).
// perm orders the columns by their energy over the index's table, so the dropped ones weigh least.
template <bool kF8>
__global__ void convert_rows_kernel(const float* __restrict__ src, int64_t n_rows, int32_t dim,
                                    int32_t n_elems, int32_t kept, const int32_t* __restrict__ perm,
                                    float scale, float limit_sq, void* __restrict__ dst,
                                    float4* __restrict__ sq) {
    const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    const float* s = src + row * dim;
    float acc = 0.f, drop = 0.f;
    for (int c = lane; c < dim; c += 32) {
        const float v = s[__ldg(perm + c)] * scale;
        acc = fmaf(v, v, acc);
        if (c >= kept) drop = fmaf(v, v, drop);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        drop += __shfl_xor_sync(0xffffffffu, drop, o);
    }
    // A per-batch row longer than the longest row of the index would overflow the fp16 range of
    // the epilogue sums: its operand is shrunk to the limit and its error set to +inf, which turns
    // every window it belongs to into an unconditional candidate (decided by the float64 rescoring
    // from the untouched fp32 row).
    const bool clamp = limit_sq > 0.f && acc > limit_sq;
    const float shrink = clamp ? sqrtf(limit_sq / acc) : 1.f;
    float err = 0.f;
    for (int c = lane; c < n_elems; c += 32) {
        const float v = c < kept ? s[__ldg(perm + c)] * scale : 0.f;
        float back;
        if (kF8) {
            const __nv_fp8_storage_t q = __nv_cvt_float_to_fp8(v * shrink, __NV_SATFINITE, __NV_E4M3);
            static_cast<uint8_t*>(dst)[row * n_elems + c] = q;
            back = __half2float(__half(__nv_cvt_fp8_to_halfraw(q, __NV_E4M3)));
        } else {
            const __half h = __float2half_rn(v * shrink);
            static_cast<__half*>(dst)[row * n_elems + c] = h;
            back = __half2float(h);
        }
        err = fmaf(v - back, v - back, err);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) err += __shfl_xor_sync(0xffffffffu, err, o);
    if (lane == 0) sq[row] = make_float4(acc, clamp ? INFINITY : err, drop, 0.f);
}

// energy[c] += sum over rows of src[row][c]^2 (which columns the pre-filter may drop)
__global__ void column_energy_kernel(const float* __restrict__ src, int64_t n_rows, int32_t dim,
                                     double* __restrict__ energy) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= dim) return;
    double acc = 0.0;
    for (int64_t r = blockIdx.y; r < n_rows; r += gridDim.y) {
        const float v = src[r * dim + c];
        if (isfinite(v)) acc += static_cast<double>(v) * v;
    }
    atomicAdd(energy + c, acc);
}

// max over rows of the squared row norm (global fp8 scale); atomicMax on the bit pattern
__global__ void rownorm_max_kernel(const float* __restrict__ src, int64_t n_rows, int32_t dim,
                                   unsigned int* out) {
    const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    float acc = 0.f;
    for (int c = lane; c < dim; c += 32) {
        const float v = src[row * dim + c];
        if (isfinite(v)) acc = fmaf(v, v, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) atomicMax(out, __float_as_uint(acc));
}

// max |x| over a float array (for the global fp16 scale); result via atomicMax on the bit pattern
__global__ void absmax_kernel(const float* __restrict__ src, int64_t n, unsigned int* out) {
    float m = 0.f;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float v = fabsf(src[i]);
        if (isfinite(v)) m = fmaxf(m, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}

// E[t, :] = rows16[tok[t]].  One 16-byte chunk per thread and per unrolled step, fully
// coalesced stores; kGatherUnroll independent row loads are in flight per thread so that the
// kernel is bound by HBM write bandwidth rather than by L2 read latency.
constexpr int kGatherUnroll = 4;

__device__ __forceinline__ int4 gather_chunk(const GatherSources& src, const int4* base16,
                                             const int4* sx16, const int4* fx16, int64_t id,
                                             int32_t chunks16, int32_t c, float4* sq) {
    int4 v = make_int4(0, 0, 0, 0);  // unknown ids embed as the zero vector
    *sq = make_float4(0.f, 0.f, 0.f, 0.f);
    if (id >= 0 && id < src.n_base) {
        v = __ldg(base16 + id * chunks16 + c);
        if (c == 0) *sq = __ldg(src.base_sq + id);
    } else if ((id -= src.n_base) >= 0 && id < src.n_sx) {
        v = __ldg(sx16 + id * chunks16 + c);
        if (c == 0) *sq = __ldg(src.sx_sq + id);
    } else if ((id -= src.n_sx) >= 0 && id < src.n_fx) {
        v = __ldg(fx16 + id * chunks16 + c);
        if (c == 0) *sq = __ldg(src.fx_sq + id);
    }
    return v;
}

__global__ void __launch_bounds__(256)
gather_kernel(const int32_t* __restrict__ tok, int64_t n_tok, const GatherSources src,
              int32_t chunks16 /* dim_pad*2/16 */, int4* __restrict__ emb,
              float4* __restrict__ tok_sq) {
    const int64_t total = n_tok * chunks16;
    const int4* base16 = reinterpret_cast<const int4*>(src.base16);
    const int4* sx16 = reinterpret_cast<const int4*>(src.sx16);
    const int4* fx16 = reinterpret_cast<const int4*>(src.fx16);
    // (token, chunk) of a flat chunk index advances by a constant step per unrolled slot; the
    // 64-bit division is done once per thread, then replaced by add-with-carry
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t stride_t = stride / chunks16;
    const int32_t stride_c = static_cast<int32_t>(stride - stride_t * chunks16);
    int64_t g0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    int64_t t0 = g0 / chunks16;
    int32_t c0 = static_cast<int32_t>(g0 - t0 * chunks16);
    while (g0 < total) {
        int4 v[kGatherUnroll];
        float4 sq[kGatherUnroll];
        int64_t t[kGatherUnroll];
        int32_t c[kGatherUnroll];
        int64_t id[kGatherUnroll];
#pragma unroll
        for (int u = 0; u < kGatherUnroll; ++u) {
            t[u] = t0;
            c[u] = c0;
            id[u] = (g0 + u * stride) < total ? __ldg(tok + t0) : -1;
            t0 += stride_t;
            c0 += stride_c;
            if (c0 >= chunks16) {
                c0 -= chunks16;
                ++t0;
            }
        }
#pragma unroll
        for (int u = 0; u < kGatherUnroll; ++u)
            v[u] = gather_chunk(src, base16, sx16, fx16, id[u], chunks16, c[u], &sq[u]);
#pragma unroll
        for (int u = 0; u < kGatherUnroll; ++u) {
            const int64_t g = g0 + u * stride;
            if (g < total) {
                emb[g] = v[u];
                if (c[u] == 0) tok_sq[t[u]] = sq[u];
            }
        }
        g0 += stride * kGatherUnroll;
    }
}

// Fused-gather path: only the per-token squares are gathered (16 B per token from L2-resident tables); the
// operand rows are fetched from the table by the distance kernel itself (TMA tile::gather4).
__global__ void __launch_bounds__(256)
gather_sq_kernel(const int32_t* __restrict__ tok, int64_t n_tok, const GatherSources src,
                 float4* __restrict__ tok_sq) {
    for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < n_tok;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        int64_t id = __ldg(tok + t);
        float4 sq = make_float4(0.f, 0.f, 0.f, 0.f);  // unknown ids embed as the zero vector
        if (id >= 0 && id < src.n_base)
            sq = __ldg(src.base_sq + id);
        else if ((id -= src.n_base) >= 0 && id < src.n_sx)
            sq = __ldg(src.sx_sq + id);
        else if ((id -= src.n_sx) >= 0 && id < src.n_fx)
            sq = __ldg(src.fx_sq + id);
        tok_sq[t] = sq;
    }
}

// Per window start t, with n = sqrt(sum_{k<w} tok_sq[t+k].x) (norm of the scaled fp32 window),
// e = sqrt(sum tok_sq[t+k].y) (norm of the rounding error of its operand rows) and g = sqrt(sum
// tok_sq[t+k].z) (norm of the elements the operand rows drop):
//   fan side    out[t] = (n, e, g)                     = (A_i, C_i, G_i)
//   script side out[t] = (coef * n - e, n + e, g)      = (B_j, D_j, H_j),  coef = 1 - thr - eps
// when the window lies inside its CSR row, else NaN -- every comparison with NaN is false,
// whatever the sign of the other factor; stored as (x, half2(y, z)), the halves rounded UP
// (pack_bound).  The distance epilogue keeps a pair iff
//   acc_ij > A_i * B_j - C_i * D_j - G_i * H_j.
// With qf, qs the rounded kept parts of the windows:
//   f.s = qf.qs + (f_kept - qf).qs + f_kept.(s_kept - qs) + f_drop.s_drop
//   => qf.qs >= f.s - |f_kept - qf| |qs| - |f| |s_kept - qs| - |f_drop| |s_drop|,  |qs| <= n_s + e_s,
// so every pair with f.s > (1 - thr) |f||s| passes: each window carries its own measured slack and a
// badly represented window (rows that underflow the operand format, or whose weight sits in the
// dropped elements) only widens its own row/column.  out is padded with NaN up to n_pad.
// (x, half2(y, z)) with the halves rounded towards +inf (a slack factor may only grow)
__device__ __forceinline__ float2 pack_bound(float x, float y, float z) {
    const __half2 h = __halves2half2(__float2half_ru(y), __float2half_ru(z));
    return make_float2(x, __uint_as_float(*reinterpret_cast<const uint32_t*>(&h)));
}

__global__ void window_norm_kernel(const float4* __restrict__ tok_sq, int64_t n_tok,
                                   const int64_t* __restrict__ off, int32_t n_rows, int32_t window,
                                   float coef, int script_side, float2* __restrict__ out,
                                   float4* __restrict__ out_plain, int64_t n_pad,
                                   unsigned long long* window_counter) {
    __shared__ int32_t row_hint;
    const int64_t t0 = static_cast<int64_t>(blockIdx.x) * blockDim.x;
    if (threadIdx.x == 0) row_hint = t0 < n_tok ? csr_row_of(off, n_rows, t0) : 0;
    __syncthreads();
    const int64_t t = t0 + threadIdx.x;
    unsigned int valid = 0;
    if (t < n_pad) {
        const float nan = __int_as_float(0x7fc00000);
        float4 r = make_float4(nan, nan, nan, 0.f);
        if (t < n_tok) {
            const int32_t row = csr_row_from_hint(off, n_rows, t, row_hint);
            if (t + window <= __ldg(off + row + 1)) {
                float s = 0.f, e = 0.f, g = 0.f;
                for (int k = 0; k < window; ++k) {
                    const float4 q = tok_sq[t + k];
                    s += q.x;
                    e += q.y;
                    g += q.z;
                }
                const float n = sqrtf(s), en = sqrtf(e), gn = sqrtf(g);
                r = script_side ? make_float4(coef * n - en, n + en, gn, 0.f) : make_float4(n, en, gn, 0.f);
                valid = 1;
            }
        }
        out[t] = pack_bound(r.x, r.y, r.z);
        if (out_plain) out_plain[t] = r;  // (tests: the unrounded values)
    }
    if (window_counter) {
        const unsigned int n = __reduce_add_sync(0xffffffffu, valid);
        if ((threadIdx.x & 31) == 0 && n) atomicAdd(window_counter, static_cast<unsigned long long>(n));
    }
}

// dst[j] = (min B, half2(max D, max H)) over the valid (B not NaN) entries j .. j+31, clamped at n: lets
// the distance epilogue reject a whole 32-column chunk with one compare of its largest accumulator
// (A minB - C maxD - G maxH <= A B_j - C D_j - G H_j for A, C, D, G, H >= 0).  No valid entry: (+inf, (0, 0)).
__global__ void sliding_minmax32_kernel(const float2* __restrict__ src, float2* __restrict__ dst, int64_t n) {
    const int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= n) return;
    float mn = INFINITY;
    __half2 mx = __float2half2_rn(0.f);
    for (int k = 0; k < 32 && j + k < n; ++k) {
        const float2 v = src[j + k];
        if (v.x != v.x) continue;  // invalid window: its halves are NaN too
        mn = fminf(mn, v.x);
        const uint32_t u = __float_as_uint(v.y);
        mx = __hmax2(mx, *reinterpret_cast<const __half2*>(&u));
    }
    dst[j] = make_float2(mn, __uint_as_float(*reinterpret_cast<const uint32_t*>(&mx)));
}

int launch_sliding_minmax32(const float2* src, float2* dst, int64_t n, cudaStream_t stream) {
    if (n <= 0) return FS_OK;
    sliding_minmax32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(src, dst, n);
    FS_CUDA_CHECK(cudaGetLastError());
    return FS_OK;
}

int launch_convert_rows(const float* src, int64_t n_rows, int32_t dim, int32_t dim_pad, int32_t kept,
                        const int32_t* perm, float scale, bool f8, float limit_sq, __half* dst, float4* sq,
                        cudaStream_t stream) {
    if (n_rows <= 0) return FS_OK;
    const int threads = 256;
    const int64_t blocks = (n_rows * 32 + threads - 1) / threads;
    if (f8)
        convert_rows_kernel<true><<<static_cast<unsigned>(blocks), threads, 0, stream>>>(
            src, n_rows, dim, 2 * dim_pad, kept, perm, scale, limit_sq, dst, sq);
    else
        convert_rows_kernel<false><<<static_cast<unsigned>(blocks), threads, 0, stream>>>(
            src, n_rows, dim, dim_pad, kept, perm, scale, limit_sq, dst, sq);
    FS_CUDA_CHECK(cudaGetLastError());
    return FS_OK;
}

int launch_column_energy(const float* src, int64_t n_rows, int32_t dim, double* energy, cudaStream_t stream) {
    if (n_rows <= 0) return FS_OK;
    const dim3 grid(static_cast<unsigned>((dim + 127) / 128), static_cast<unsigned>(n_rows < 256 ? n_rows : 256));
    column_energy_kernel<<<grid, 128, 0, stream>>>(src, n_rows, dim, energy);
    FS_CUDA_CHECK(cudaGetLastError());
    return FS_OK;
}

int launch_rownorm_max(const float* src, int64_t n_rows, int32_t dim, unsigned int* out, cudaStream_t stream) {
    if (n_rows <= 0) return FS_OK;
    const int threads = 256;
    const int64_t blocks = (n_rows * 32 + threads - 1) / threads;
    rownorm_max_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(src, n_rows, dim, out);
    FS_CUDA_CHECK(cudaGetLastError());
    return FS_OK;
}

int launch_absmax(const float* src, int64_t n, unsigned int* out, cudaStream_t stream) {
    if (n <= 0) return FS_OK;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    absmax_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(src, n, out);
    FS_CUDA_CHECK(cudaGetLastError());
    return FS_OK;
}

int launch_gather(const int32_t* tok, int64_t n_tok, const GatherSources& src, int32_t dim_pad,
                  __half* emb, float4* tok_sq, int sm_count, cudaStream_t stream) {
    if (n_tok <= 0) return FS_OK;
    const int32_t chunks16 = dim_pad * 2 / 16;
    const int64_t total = n_tok * chunks16;
    const int threads = 256;
    int64_t blocks = (total + threads - 1) / threads;
    const int64_t max_blocks = static_cast<int64_t>(sm_count) * 8;  // one resident wave, grid-stride
    if (blocks > max_blocks) blocks = max_blocks;
    gather_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(
        tok, n_tok, src, chunks16, reinterpret_cast<int4*>(emb), tok_sq);
    FS_CUDA_CHECK(cudaGetLastError());
    return FS_OK;
}

int launch_gather_sq(const int32_t* tok, int64_t n_tok, const GatherSources& src, float4* tok_sq, int sm_count,
                     cudaStream_t stream) {
    if (n_tok <= 0) return FS_OK;
    const int threads = 256;
    int64_t blocks = (n_tok + threads - 1) / threads;
    const int64_t max_blocks = static_cast<int64_t>(sm_count) * 8;
    if (blocks > max_blocks) blocks = max_blocks;
    gather_sq_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(tok, n_tok, src, tok_sq);
    FS_CUDA_CHECK(cudaGetLastError());
    return FS_OK;
}

int launch_window_norm(const float4* tok_sq, int64_t n_tok, const int64_t* off, int32_t n_rows,
                       int32_t window, float coef, bool script_side, float2* out, float4* out_plain,
                       int64_t n_pad, unsigned long long* window_counter, cudaStream_t stream) {
    if (n_pad <= 0) return FS_OK;
    const int threads = 256;
    const int64_t blocks = (n_pad + threads - 1) / threads;
    window_norm_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(
        tok_sq, n_tok, off, n_rows, window, coef, script_side ? 1 : 0, out, out_plain, n_pad, window_counter);
    FS_CUDA_CHECK(cudaGetLastError());
    return FS_OK;
}

}  // namespace fs
