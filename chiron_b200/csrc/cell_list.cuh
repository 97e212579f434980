// Cell-list build kernels shared by the reference-shaped list builder (nlist.cu) and the barostat loop
// (mc.cu): counting sort of the particles into cells of edge >= cutoff + skin and the 27-cell bitmap
// sweep that emits NeighborListNsqrd's (N, M) arrays (chiron/neighbors.py:548-729) without a sort.
#pragma once
#include <math.h>
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// Row finalisation shared by both builders: padding value, pad mask, count.
// `first` is the smallest listed neighbour id (valid when count > 0).
// ---------------------------------------------------------------------------------------------
static __device__ __forceinline__ void finish_row(int i, int count, uint32_t first, int M,
                                           uint32_t* __restrict__ list, int32_t* __restrict__ mask,
                                           int32_t* __restrict__ nn, int lane) {
    // neighbors.py:606-609: fill = argmax(mask) (first True, 0 if none); if fill == i: fill += 1
    uint32_t fill = count > 0 ? first : 0u;
    if (fill == (uint32_t)i) fill += 1u;
    const int stored = count < M ? count : M;
    for (int k = stored + lane; k < M; k += 32) list[(size_t)i * M + k] = fill;
    if (mask != nullptr)   // the barostat loop's superset build keeps ids and counts only
        for (int k = lane; k < M; k += 32) mask[(size_t)i * M + k] = (k < count) ? 1 : 0;
    if (lane == 0) nn[i] = count;
}


// Optional double buffering: when `flip` is given and *flip != 0 a kernel works on the alternate
// pointers (the barostat loop builds into whichever list set is NOT current; the host path passes
// flip = NULL).
template <class T>
static __device__ __forceinline__ T* cl_pick(T* a, T* b, const int* flip) {
    return (flip != nullptr && *flip != 0) ? b : a;
}

// max_i n_i and the number of rows with n_i == M (the reference's growth trigger, neighbors.py:709)
static __global__ void k_count_stats(const int32_t* nn, int n, int M, int* __restrict__ out,
                                     const int32_t* nn_alt = nullptr, const int* flip = nullptr,
                                     const int* skip = nullptr) {
    if (skip != nullptr && *skip != 0) return;
    nn = cl_pick(nn, nn_alt, flip);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int v = i < n ? nn[i] : 0;
    int eq = (i < n && v == M) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
        eq += __shfl_xor_sync(0xffffffffu, eq, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (v > 0) atomicMax(&out[0], v);
        if (eq > 0) atomicAdd(&out[1], eq);
    }
}


struct CellGrid {
    int nx, ny, nz;
    float inv_cx, inv_cy, inv_cz;  // 1 / cell edge
    int S;                         // stencil radius: cells of edge >= (cutoff+skin) / S, (2S+1)^3 cells swept
};

struct SweepConst {
    float c;               // cutoff + skin (exact predicate)
    float c2_lo, c2_hi;    // r2 < c2_lo: inside, r2 >= c2_hi: outside, in between: exact predicate
    float inv_lx, inv_ly, inv_lz;
};

// Geometry of one list build, read from DEVICE memory by every kernel of the build: the host path
// uploads it (chx_nlist_build_cell), the barostat loop (mc.cu) computes it on the device because the
// box -- and with it the cell grid -- changes with every volume proposal.
struct CellParams {
    Box box;
    CellGrid g;
    SweepConst sc;
    int ncell;
    int valid;             // 0: the build is skipped (grid unusable: the caller falls back)
};

// Cells of edge >= (cutoff+skin)(1+1e-5) / S (binning uses a rounded wrapped coordinate, the margin keeps
// every pair inside the predicate within the (2S+1)^3 stencil), at least 2S+1 and at most 256 per edge.
// S = 2 (half-size cells, 125 of them: 2.2x less volume to search than 27 full-size ones) when the box and
// the cell capacity allow it, else S = 1.
__host__ __device__ inline bool make_cell_params(float lx, float ly, float lz, float cutoff_plus_skin,
                                                 int ncell_cap, CellParams& P) {
    P.valid = 0;
    int nx = 0, ny = 0, nz = 0, S = 0;
    for (int s = 2; s >= 1 && S == 0; --s) {
        const double rc = (double)cutoff_plus_skin * (1.0 + 1e-5) / s;
        nx = (int)floor((double)lx / rc); ny = (int)floor((double)ly / rc); nz = (int)floor((double)lz / rc);
        nx = nx > 256 ? 256 : nx; ny = ny > 256 ? 256 : ny; nz = nz > 256 ? 256 : nz;
        const int need = 2 * s + 1;
        if (nx >= need && ny >= need && nz >= need && (ncell_cap <= 0 || nx * ny * nz <= ncell_cap)) S = s;
    }
    if (S == 0) return false;
    P.box = make_box(lx, ly, lz);
    P.g.nx = nx; P.g.ny = ny; P.g.nz = nz; P.g.S = S;
    P.g.inv_cx = (float)(nx / (double)lx); P.g.inv_cy = (float)(ny / (double)ly); P.g.inv_cz = (float)(nz / (double)lz);
    P.ncell = nx * ny * nz;
    P.sc.c = cutoff_plus_skin;
    const double c2 = (double)cutoff_plus_skin * (double)cutoff_plus_skin;
    const double lmax = lx > ly ? (lx > lz ? lx : lz) : (ly > lz ? ly : lz);
    const double band = c2 * 4e-6 + 8.0 * 1.2e-7 * lmax * cutoff_plus_skin;
    P.sc.c2_lo = (float)(c2 - band); P.sc.c2_hi = (float)(c2 + band);
    P.sc.inv_lx = 1.0f / lx; P.sc.inv_ly = 1.0f / ly; P.sc.inv_lz = 1.0f / lz;
    P.valid = 1;
    return true;
}

__device__ __forceinline__ int cell_coord(float x, float L, float inv_c, int nc) {
    float w = ref_wrap(x, L);
    int c = (int)(w * inv_c);
    c = c < 0 ? 0 : c;
    return c >= nc ? nc - 1 : c;
}

static __global__ void k_cell_count(const float* x, int n, const CellParams* __restrict__ P,
                                    int* __restrict__ cell_of, int* __restrict__ cell_count,
                                    const float* x_alt = nullptr, const int* flip = nullptr) {
    if (!P->valid) return;
    x = cl_pick(x, x_alt, flip);
    const Box box = P->box;
    const CellGrid g = P->g;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cx = cell_coord(x[3 * i], box.lx, g.inv_cx, g.nx);
    int cy = cell_coord(x[3 * i + 1], box.ly, g.inv_cy, g.ny);
    int cz = cell_coord(x[3 * i + 2], box.lz, g.inv_cz, g.nz);
    int c = (cx * g.ny + cy) * g.nz + cz;
    cell_of[i] = c;
    atomicAdd(&cell_count[c], 1);
}

// single-block exclusive scan: start[c] = sum_{c'<c} count[c'], start[ncell] = n; count is reset
// to 0 so it can serve as the fill cursor.
static __global__ void k_cell_scan(int* __restrict__ count, int* __restrict__ start,
                                   const CellParams* __restrict__ P) {
    if (!P->valid) return;
    const int ncell = P->ncell;
    __shared__ int part[1024];
    const int t = threadIdx.x;
    const int per = (ncell + blockDim.x - 1) / blockDim.x;
    const int lo = t * per, hi = min(ncell, lo + per);
    int s = 0;
    for (int c = lo; c < hi; ++c) s += count[c];
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < blockDim.x; o <<= 1) {
        int v = t >= o ? part[t - o] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    int run = part[t] - s;
    for (int c = lo; c < hi; ++c) {
        int k = count[c];
        start[c] = run;
        count[c] = 0;
        run += k;
    }
    if (t == blockDim.x - 1) start[ncell] = part[t];
}

// slot of particle i in cell order; xs4[slot] = (x, y, z, id) so that the sweep reads one coalesced
// float4 per candidate instead of an index and three scattered floats
static __global__ void k_cell_fill(const float* x, const int* __restrict__ cell_of, int n,
                                   const CellParams* __restrict__ P, const int* __restrict__ start,
                                   int* __restrict__ cursor, int* __restrict__ order, float4* __restrict__ xs4,
                                   const float* x_alt = nullptr, const int* flip = nullptr) {
    if (!P->valid) return;
    x = cl_pick(x, x_alt, flip);
    const Box box = P->box;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = cell_of[i];
    const int slot = start[c] + atomicAdd(&cursor[c], 1);
    order[slot] = i;
    // wrapped into [0, L): the sweep's fast test resolves periodic images per CELL (a shift of the
    // row particle), not per pair; the exact predicate reads the caller's coordinates
    xs4[slot] = make_float4(ref_wrap(x[3 * i], box.lx), ref_wrap(x[3 * i + 1], box.ly),
                            ref_wrap(x[3 * i + 2], box.lz), __int_as_float(i));
}


// 27-cell sweep, bitmap variant (n <= 32 * 32 * W particles).  One warp per row i.  Candidates come as
// coalesced float4 (cell order); a fast FMA test decides the clear cases and only pairs within a few
// ulps of the cutoff go through the reference's exact predicate (orientation i < j, like the half
// list); hits set bit j of a per-warp bitmap in shared memory, and reading the bitmap back in word
// order yields the row in ascending id order -- no sort.  Lane l owns bitmap words [l W, (l+1) W),
// stored with a stride of W + 1 words so that the lanes hit different banks.
template <int WARPS>
static __global__ void __launch_bounds__(WARPS * 32)
k_build_cell_bm(const float* x, const float4* __restrict__ xs4, int n,
                const CellParams* __restrict__ P, int M, int wshift, const int* __restrict__ cell_of,
                const int* __restrict__ start, uint32_t* list, int32_t* mask, int32_t* nn,
                const float* x_alt = nullptr, uint32_t* list_alt = nullptr, int32_t* mask_alt = nullptr,
                int32_t* nn_alt = nullptr, const int* flip = nullptr) {
    extern __shared__ uint32_t smem[];
    if (!P->valid) return;
    x = cl_pick(x, x_alt, flip);
    list = cl_pick(list, list_alt, flip);
    mask = cl_pick(mask, mask_alt, flip);
    nn = cl_pick(nn, nn_alt, flip);
    const Box box = P->box;
    const CellGrid g = P->g;
    const SweepConst sc = P->sc;
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const int i = blockIdx.x * WARPS + w;
    if (i >= n) return;
    const int W = 1 << wshift;
    uint32_t* bm = smem + (size_t)w * 32 * (W + 1);
    for (int k = lane; k < 32 * (W + 1); k += 32) bm[k] = 0u;
    __syncwarp();
    const float xi = x[3 * i], yi = x[3 * i + 1], zi = x[3 * i + 2];
    const float xw = ref_wrap(xi, box.lx), yw = ref_wrap(yi, box.ly), zw = ref_wrap(zi, box.lz);
    const int ci = cell_of[i];
    const int cz = ci % g.nz, cy = (ci / g.nz) % g.ny, cx = ci / (g.nz * g.ny);
    const int S = g.S;
    for (int dx = -S; dx <= S; ++dx) {
        int ax = cx + dx;
        float xs = xw;                       // row particle shifted into the neighbour cell's image
        if (ax < 0) { ax += g.nx; xs = xw + box.lx; } else if (ax >= g.nx) { ax -= g.nx; xs = xw - box.lx; }
        for (int dy = -S; dy <= S; ++dy) {
            int ay = cy + dy;
            float ys = yw;
            if (ay < 0) { ay += g.ny; ys = yw + box.ly; } else if (ay >= g.ny) { ay -= g.ny; ys = yw - box.ly; }
            // the cells cz - S .. cz + S of this column are contiguous in cell order: one run, or two
            // when the column wraps around the box in z
            const int col = (ax * g.ny + ay) * g.nz;
#define CL_SWEEP_RUN(A0, A1, ZS)                                                                          \
            do {                                                                                          \
                const float zs = (ZS);                                                                    \
                const int s = start[col + (A0)], e = start[col + (A1) + 1];                               \
                for (int t = s + lane; t < e; t += 32) {                                                  \
                    const float4 p = xs4[t];                                                              \
                    const int j = __float_as_int(p.w);                                                    \
                    if (j <= i) continue;                                                                 \
                    const float ddx = xs - p.x, ddy = ys - p.y, ddz = zs - p.z;                           \
                    const float r2 = fmaf(ddz, ddz, fmaf(ddy, ddy, ddx * ddx));                           \
                    if (r2 >= sc.c2_hi) continue;                                                         \
                    if (r2 >= sc.c2_lo) {                                                                 \
                        float rx, ry, rz, d;                                                              \
                        ref_displacement<true>(xi, yi, zi, x[3 * j], x[3 * j + 1], x[3 * j + 2], box, rx, ry, rz, d); \
                        if (!(d < sc.c)) continue;                                                        \
                    }                                                                                     \
                    const int q = j >> 5;                                                                 \
                    atomicOr(&bm[q + (q >> wshift)], 1u << (j & 31));                                     \
                }                                                                                         \
            } while (0)
            CL_SWEEP_RUN(max(cz - S, 0), min(cz + S, g.nz - 1), zw);
            if (cz - S < 0) CL_SWEEP_RUN(cz - S + g.nz, g.nz - 1, zw + box.lz);
            if (cz + S >= g.nz) CL_SWEEP_RUN(0, cz + S - g.nz, zw - box.lz);
#undef CL_SWEEP_RUN
        }
    }
    __syncwarp();
    // read back: lane l scans its W words in order
    const uint32_t* mine = bm + (size_t)lane * (W + 1);
    int cnt = 0;
    uint32_t first = 0xffffffffu;
    for (int k = 0; k < W; ++k) {
        const uint32_t v = mine[k];
        if (v && first == 0xffffffffu) first = (uint32_t)(((lane << wshift) + k) << 5) + (uint32_t)(__ffs(v) - 1);
        cnt += __popc(v);
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int count = __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    int pos = incl - cnt;
    if (cnt > 0 && pos < M) {
        for (int k = 0; k < W && pos < M; ++k) {
            uint32_t v = mine[k];
            const uint32_t base = (uint32_t)(((lane << wshift) + k) << 5);
            while (v && pos < M) {
                const int b = __ffs(v) - 1;
                v &= v - 1u;
                list[(size_t)i * M + pos] = base + (uint32_t)b;
                ++pos;
            }
        }
    }
    __syncwarp();
    finish_row(i, count, first, M, list, mask, nn, lane);
}


// shared-memory configuration of the bitmap sweep for n particles; false when a bitmap does not fit
static inline bool cell_bm_config(int n, int& wshift, int& warps, size_t& smem) {
    const int nwords = chx_div_up(n, 32);
    wshift = 0;
    while ((32 << wshift) < nwords) ++wshift;
    const size_t per_warp = (size_t)32 * ((1u << wshift) + 1) * sizeof(uint32_t);
    if (per_warp * 2 > 200 * 1024) return false;
    warps = 8;
    while (warps > 2 && per_warp * warps > 200 * 1024) warps >>= 1;
    smem = per_warp * warps;
    return true;
}

static inline int cell_bm_launch(chx_ctx* ctx, const float* x, const float4* xs4, int n, const CellParams* P_dev,
                                 int M, int wshift, int warps, size_t smem, const int* cell_of, const int* start,
                                 uint32_t* list, int32_t* mask, int32_t* nn, const float* x_alt = nullptr,
                                 uint32_t* list_alt = nullptr, int32_t* mask_alt = nullptr,
                                 int32_t* nn_alt = nullptr, const int* flip = nullptr) {
#define BM_LAUNCH(WN)                                                                                       \
    do {                                                                                                    \
        CHX_CUDA(cudaFuncSetAttribute(k_build_cell_bm<WN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_build_cell_bm<WN><<<chx_div_up(n, WN), WN * 32, smem, ctx->stream>>>(x, xs4, n, P_dev, M, wshift, cell_of, \
                                                                                start, list, mask, nn, x_alt,   \
                                                                                list_alt, mask_alt, nn_alt, flip); \
    } while (0)
    if (warps == 8) BM_LAUNCH(8); else if (warps == 4) BM_LAUNCH(4); else BM_LAUNCH(2);
#undef BM_LAUNCH
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}
