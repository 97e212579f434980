// Reference-shaped pair finding: NeighborListNsqrd (O(N^2) restatement and O(N) cell list that
// emits the identical arrays), calculate, check, PairListNsqrd.
//   chiron/neighbors.py:513-907 (neighbor list), :1018-1289 (pair list)
#include <stdlib.h>
#include <math.h>
#include "common.cuh"

#define WARPS_PER_BLOCK 8

// ---------------------------------------------------------------------------------------------
// Row finalisation shared by both builders: padding value, pad mask, count.
// `first` is the smallest listed neighbour id (valid when count > 0).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void finish_row(int i, int count, uint32_t first, int M,
                                           uint32_t* __restrict__ list, int32_t* __restrict__ mask,
                                           int32_t* __restrict__ nn, int lane) {
    // neighbors.py:606-609: fill = argmax(mask) (first True, 0 if none); if fill == i: fill += 1
    uint32_t fill = count > 0 ? first : 0u;
    if (fill == (uint32_t)i) fill += 1u;
    const int stored = count < M ? count : M;
    for (int k = stored + lane; k < M; k += 32) list[(size_t)i * M + k] = fill;
    for (int k = lane; k < M; k += 32) mask[(size_t)i * M + k] = (k < count) ? 1 : 0;
    if (lane == 0) nn[i] = count;
}

// ---------------------------------------------------------------------------------------------
// O(N^2) builder: one warp per row, j ascending, ballot compaction keeps the order.
// ---------------------------------------------------------------------------------------------
template <bool PERIODIC>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
k_build_nsq(const float* __restrict__ x, int n, Box box, float c, int M,
            uint32_t* __restrict__ list, int32_t* __restrict__ mask, int32_t* __restrict__ nn) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (i >= n) return;
    const float xi = x[3 * i], yi = x[3 * i + 1], zi = x[3 * i + 2];
    int count = 0;
    uint32_t first = 0u;
    for (int j0 = (i + 1) & ~31; j0 < n; j0 += 32) {
        const int j = j0 + lane;
        bool hit = false;
        if (j > i && j < n) {
            float rx, ry, rz, d;
            ref_displacement<PERIODIC>(xi, yi, zi, x[3 * j], x[3 * j + 1], x[3 * j + 2], box, rx,
                                       ry, rz, d);
            hit = d < c;
        }
        const unsigned b = __ballot_sync(0xffffffffu, hit);
        if (b) {
            if (count == 0) first = (uint32_t)(j0 + __ffs(b) - 1);
            if (hit) {
                const int pos = count + __popc(b & ((1u << lane) - 1u));
                if (pos < M) list[(size_t)i * M + pos] = (uint32_t)j;
            }
            count += __popc(b);
        }
    }
    finish_row(i, count, first, M, list, mask, nn, lane);
}

// max_i n_i and #{i : n_i == M}
__global__ void k_count_stats(const int32_t* __restrict__ nn, int n, int M, int* __restrict__ out) {
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

static int finish_build(chx_ctx* ctx, const int32_t* nn, int n, int M, int* max_count_host,
                        int* count_eq_M_host, int* extra_flag_dev, int* extra_flag_host) {
    int* stats = (int*)chx_scratch(ctx, 256);
    if (!stats) return CHX_CUDA_ERROR;
    // NOTE: stats lives at the start of the scratch area; builders place their own workspace after it
    CHX_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(int), ctx->stream));
    k_count_stats<<<chx_div_up(n, 256), 256, 0, ctx->stream>>>(nn, n, M, stats);
    CHX_LAUNCHED(ctx);
    CHX_CUDA(cudaMemcpyAsync(ctx->host_pinned, stats, 2 * sizeof(int), cudaMemcpyDeviceToHost,
                             ctx->stream));
    if (extra_flag_dev)
        CHX_CUDA(cudaMemcpyAsync(ctx->host_pinned + 2, extra_flag_dev, sizeof(int),
                                 cudaMemcpyDeviceToHost, ctx->stream));
    CHX_CUDA(cudaStreamSynchronize(ctx->stream));
    if (max_count_host) *max_count_host = ctx->host_pinned[0];
    if (count_eq_M_host) *count_eq_M_host = ctx->host_pinned[1];
    if (extra_flag_host) *extra_flag_host = ctx->host_pinned[2];
    return CHX_OK;
}

// ---------------------------------------------------------------------------------------------
// Cell list: counting sort of particles into cells of edge >= cutoff+skin, 27-cell sweep with the
// exact predicate, per-row bitonic sort so rows come out in ascending id order.
// ---------------------------------------------------------------------------------------------
struct CellGrid {
    int nx, ny, nz;
    float inv_cx, inv_cy, inv_cz;  // 1 / cell edge
};

__device__ __forceinline__ int cell_coord(float x, float L, float inv_c, int nc) {
    float w = ref_wrap(x, L);
    int c = (int)(w * inv_c);
    c = c < 0 ? 0 : c;
    return c >= nc ? nc - 1 : c;
}

__global__ void k_cell_count(const float* __restrict__ x, int n, Box box, CellGrid g,
                             int* __restrict__ cell_of, int* __restrict__ cell_count) {
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
__global__ void k_cell_scan(int* __restrict__ count, int* __restrict__ start, int ncell) {
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
__global__ void k_cell_fill(const float* __restrict__ x, const int* __restrict__ cell_of, int n, Box box,
                            const int* __restrict__ start, int* __restrict__ cursor,
                            int* __restrict__ order, float4* __restrict__ xs4) {
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

// warp-level bitonic sort of `len` (power of two) uint32 in shared memory, ascending
__device__ __forceinline__ void warp_bitonic_sort(uint32_t* buf, int len, int lane) {
    for (int k = 2; k <= len; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < len; t += 32) {
                int p = t ^ j;
                if (p > t) {
                    uint32_t a = buf[t], b = buf[p];
                    bool up = (t & k) == 0;
                    if ((a > b) == up) { buf[t] = b; buf[p] = a; }
                }
            }
            __syncwarp();
        }
    }
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_build_cell(const float* __restrict__ x, int n, Box box, CellGrid g, float c, int M, int cap,
             const int* __restrict__ cell_of, const int* __restrict__ start,
             const int* __restrict__ order, uint32_t* __restrict__ list,
             int32_t* __restrict__ mask, int32_t* __restrict__ nn, int* __restrict__ overflow) {
    extern __shared__ uint32_t smem[];
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const int i = blockIdx.x * WARPS + w;
    if (i >= n) return;
    uint32_t* buf = smem + (size_t)w * cap;
    const float xi = x[3 * i], yi = x[3 * i + 1], zi = x[3 * i + 2];
    const int ci = cell_of[i];
    const int cz = ci % g.nz, cy = (ci / g.nz) % g.ny, cx = ci / (g.nz * g.ny);
    int count = 0;
    for (int dx = -1; dx <= 1; ++dx) {
        int ax = cx + dx; ax = ax < 0 ? ax + g.nx : (ax >= g.nx ? ax - g.nx : ax);
        for (int dy = -1; dy <= 1; ++dy) {
            int ay = cy + dy; ay = ay < 0 ? ay + g.ny : (ay >= g.ny ? ay - g.ny : ay);
            for (int dz = -1; dz <= 1; ++dz) {
                int az = cz + dz; az = az < 0 ? az + g.nz : (az >= g.nz ? az - g.nz : az);
                const int cc = (ax * g.ny + ay) * g.nz + az;
                const int s = start[cc], e = start[cc + 1];
                for (int t0 = s; t0 < e; t0 += 32) {
                    const int t = t0 + lane;
                    bool hit = false;
                    int j = -1;
                    if (t < e) {
                        j = order[t];
                        if (j > i) {
                            float rx, ry, rz, d;
                            ref_displacement<true>(xi, yi, zi, x[3 * j], x[3 * j + 1], x[3 * j + 2],
                                                   box, rx, ry, rz, d);
                            hit = d < c;
                        }
                    }
                    const unsigned b = __ballot_sync(0xffffffffu, hit);
                    if (hit) {
                        const int pos = count + __popc(b & ((1u << lane) - 1u));
                        if (pos < cap) buf[pos] = (uint32_t)j;
                    }
                    count += __popc(b);
                }
            }
        }
    }
    __syncwarp();
    if (count > cap) {
        if (lane == 0) atomicExch(overflow, 1);
        if (lane == 0) nn[i] = count;
        return;
    }
    int len = 32;
    while (len < count) len <<= 1;
    for (int t = count + lane; t < len; t += 32) buf[t] = 0xffffffffu;
    __syncwarp();
    warp_bitonic_sort(buf, len, lane);
    const int stored = count < M ? count : M;
    for (int k = lane; k < stored; k += 32) list[(size_t)i * M + k] = buf[k];
    const uint32_t first = count > 0 ? buf[0] : 0u;
    finish_row(i, count, first, M, list, mask, nn, lane);
}

// 27-cell sweep, bitmap variant (n <= 32 * 32 * W particles).  One warp per row i.  Candidates come as
// coalesced float4 (cell order); a fast FMA test decides the clear cases and only pairs within a few
// ulps of the cutoff go through the reference's exact predicate (orientation i < j, like the half
// list); hits set bit j of a per-warp bitmap in shared memory, and reading the bitmap back in word
// order yields the row in ascending id order -- no sort.  Lane l owns bitmap words [l W, (l+1) W),
// stored with a stride of W + 1 words so that the lanes hit different banks.
struct SweepConst {
    float c;               // cutoff + skin (exact predicate)
    float c2_lo, c2_hi;    // r2 < c2_lo: inside, r2 >= c2_hi: outside, in between: exact predicate
    float inv_lx, inv_ly, inv_lz;
};

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_build_cell_bm(const float* __restrict__ x, const float4* __restrict__ xs4, int n, Box box, CellGrid g,
                SweepConst sc, int M, int wshift, const int* __restrict__ cell_of,
                const int* __restrict__ start, uint32_t* __restrict__ list, int32_t* __restrict__ mask,
                int32_t* __restrict__ nn) {
    extern __shared__ uint32_t smem[];
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
    for (int dx = -1; dx <= 1; ++dx) {
        int ax = cx + dx;
        float xs = xw;                       // row particle shifted into the neighbour cell's image
        if (ax < 0) { ax += g.nx; xs = xw + box.lx; } else if (ax >= g.nx) { ax -= g.nx; xs = xw - box.lx; }
        for (int dy = -1; dy <= 1; ++dy) {
            int ay = cy + dy;
            float ys = yw;
            if (ay < 0) { ay += g.ny; ys = yw + box.ly; } else if (ay >= g.ny) { ay -= g.ny; ys = yw - box.ly; }
            for (int dz = -1; dz <= 1; ++dz) {
                int az = cz + dz;
                float zs = zw;
                if (az < 0) { az += g.nz; zs = zw + box.lz; } else if (az >= g.nz) { az -= g.nz; zs = zw - box.lz; }
                const int cc = (ax * g.ny + ay) * g.nz + az;
                const int s = start[cc], e = start[cc + 1];
                for (int t = s + lane; t < e; t += 32) {
                    const float4 p = xs4[t];
                    const int j = __float_as_int(p.w);
                    if (j <= i) continue;
                    const float ddx = xs - p.x, ddy = ys - p.y, ddz = zs - p.z;
                    const float r2 = fmaf(ddz, ddz, fmaf(ddy, ddy, ddx * ddx));
                    if (r2 >= sc.c2_hi) continue;
                    if (r2 >= sc.c2_lo) {
                        float rx, ry, rz, d;
                        ref_displacement<true>(xi, yi, zi, x[3 * j], x[3 * j + 1], x[3 * j + 2], box, rx, ry, rz, d);
                        if (!(d < sc.c)) continue;
                    }
                    const int q = j >> 5;
                    atomicOr(&bm[q + (q >> wshift)], 1u << (j & 31));
                }
            }
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

// ---------------------------------------------------------------------------------------------
// calculate / check
// ---------------------------------------------------------------------------------------------
template <bool PERIODIC>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
k_nlist_calculate(const float* __restrict__ x, int n, Box box, float cutoff, int M,
                  const uint32_t* __restrict__ list, const int32_t* __restrict__ pad,
                  int32_t* __restrict__ n_out, int32_t* __restrict__ mask_out,
                  float* __restrict__ dist, float* __restrict__ rij) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (i >= n) return;
    const float xi = x[3 * i], yi = x[3 * i + 1], zi = x[3 * i + 2];
    int cnt = 0;
    for (int k = lane; k < M; k += 32) {
        const size_t o = (size_t)i * M + k;
        const uint32_t j = list[o];
        float rx, ry, rz, d;
        ref_displacement<PERIODIC>(xi, yi, zi, x[3 * j], x[3 * j + 1], x[3 * j + 2], box, rx, ry,
                                   rz, d);
        const int m = (d < cutoff) && (pad[o] != 0);
        mask_out[o] = m;
        dist[o] = d;
        rij[3 * o] = rx; rij[3 * o + 1] = ry; rij[3 * o + 2] = rz;
        cnt += m;
    }
    cnt = warp_sum(cnt);
    if (lane == 0) n_out[i] = cnt;
}

template <bool PERIODIC>
__global__ void k_nlist_check(const float* __restrict__ x, const float* __restrict__ ref, int n,
                              Box box, float half_skin, int32_t* __restrict__ flag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool moved = false;
    if (i < n) {
        float rx, ry, rz, d;
        ref_displacement<PERIODIC>(x[3 * i], x[3 * i + 1], x[3 * i + 2], ref[3 * i], ref[3 * i + 1],
                                   ref[3 * i + 2], box, rx, ry, rz, d);
        moved = d >= half_skin;
    }
    if (__syncthreads_or(moved) && threadIdx.x == 0) atomicOr(flag, 1);
}

// ---------------------------------------------------------------------------------------------
// PairListNsqrd
// ---------------------------------------------------------------------------------------------
__global__ void k_pairlist_build(int n, uint32_t* __restrict__ all_pairs,
                                 uint8_t* __restrict__ red) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long total = (long long)n * (n - 1);
    if (t >= total) return;
    const int i = (int)(t / (n - 1)), k = (int)(t % (n - 1));
    const uint32_t j = k < i ? k : k + 1;
    all_pairs[t] = j;
    red[t] = (uint32_t)i < j;
}

template <bool PERIODIC>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
k_pairlist_calculate(const float* __restrict__ x, int n, Box box, float cutoff, bool use_cutoff,
                     int32_t* __restrict__ n_out, int32_t* __restrict__ mask_out,
                     float* __restrict__ dist, float* __restrict__ rij) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (i >= n) return;
    const float xi = x[3 * i], yi = x[3 * i + 1], zi = x[3 * i + 2];
    int cnt = 0;
    for (int k = lane; k < n - 1; k += 32) {
        const int j = k < i ? k : k + 1;
        const size_t o = (size_t)i * (n - 1) + k;
        float rx, ry, rz, d;
        ref_displacement<PERIODIC>(xi, yi, zi, x[3 * j], x[3 * j + 1], x[3 * j + 2], box, rx, ry,
                                   rz, d);
        const int m = (i < j) && (!use_cutoff || d < cutoff);
        mask_out[o] = m;
        dist[o] = d;
        rij[3 * o] = rx; rij[3 * o + 1] = ry; rij[3 * o + 2] = rz;
        cnt += m;
    }
    cnt = warp_sum(cnt);
    if (lane == 0) n_out[i] = cnt;
}

extern "C" {

int chx_nlist_build_nsq(chx_ctx* ctx, const float* x, int n, float lx, float ly, float lz,
                        int periodic, float cutoff_plus_skin, int M, uint32_t* neighbor_list,
                        int32_t* neighbor_mask, int32_t* n_neighbors, int* max_count_host,
                        int* count_eq_M_host) {
    CHX_REQUIRE(ctx && x && neighbor_list && neighbor_mask && n_neighbors, "NULL argument");
    CHX_REQUIRE(n > 0 && M > 0, "n and M must be positive");
    Box box = make_box(lx, ly, lz);
    const int blocks = chx_div_up(n, WARPS_PER_BLOCK);
    if (periodic)
        k_build_nsq<true><<<blocks, WARPS_PER_BLOCK * 32, 0, ctx->stream>>>(
            x, n, box, cutoff_plus_skin, M, neighbor_list, neighbor_mask, n_neighbors);
    else
        k_build_nsq<false><<<blocks, WARPS_PER_BLOCK * 32, 0, ctx->stream>>>(
            x, n, box, cutoff_plus_skin, M, neighbor_list, neighbor_mask, n_neighbors);
    CHX_LAUNCHED(ctx);
    return finish_build(ctx, n_neighbors, n, M, max_count_host, count_eq_M_host, nullptr, nullptr);
}

int chx_nlist_build_cell(chx_ctx* ctx, const float* x, int n, float lx, float ly, float lz,
                         int periodic, float cutoff_plus_skin, int M, uint32_t* neighbor_list,
                         int32_t* neighbor_mask, int32_t* n_neighbors, int* max_count_host,
                         int* count_eq_M_host) {
    CHX_REQUIRE(ctx && x && neighbor_list && neighbor_mask && n_neighbors, "NULL argument");
    CHX_REQUIRE(n > 0 && M > 0, "n and M must be positive");
    // cell edge >= (cutoff+skin)(1+1e-5): binning uses a rounded wrapped coordinate, the margin
    // keeps every pair inside the predicate within the 27-cell stencil.
    const double rc = (double)cutoff_plus_skin * (1.0 + 1e-5);
    int nx = (int)floor((double)lx / rc), ny = (int)floor((double)ly / rc),
        nz = (int)floor((double)lz / rc);
    if (!periodic || nx < 3 || ny < 3 || nz < 3)
        return chx_nlist_build_nsq(ctx, x, n, lx, ly, lz, periodic, cutoff_plus_skin, M,
                                   neighbor_list, neighbor_mask, n_neighbors, max_count_host,
                                   count_eq_M_host);
    nx = nx > 256 ? 256 : nx; ny = ny > 256 ? 256 : ny; nz = nz > 256 ? 256 : nz;
    CellGrid g;
    g.nx = nx; g.ny = ny; g.nz = nz;
    g.inv_cx = (float)(nx / (double)lx); g.inv_cy = (float)(ny / (double)ly);
    g.inv_cz = (float)(nz / (double)lz);
    const int ncell = nx * ny * nz;
    Box box = make_box(lx, ly, lz);
    // scratch layout: [stats 256B][overflow flag][xs4 n float4][cell_of n][count ncell+1][start ncell+1][order n]
    const size_t ints = 64 + 64 + 4 * (size_t)n + (size_t)n + 2 * ((size_t)ncell + 1) + (size_t)n;
    int* base = (int*)chx_scratch(ctx, ints * sizeof(int));
    if (!base) return CHX_CUDA_ERROR;
    int* overflow = base + 64;
    float4* xs4 = reinterpret_cast<float4*>(base + 128);     // 512 B into a 256 B aligned block
    int* cell_of = base + 128 + 4 * (size_t)n;
    int* count = cell_of + n;
    int* start = count + ncell + 1;
    int* order = start + ncell + 1;
    CHX_CUDA(cudaMemsetAsync(overflow, 0, sizeof(int), ctx->stream));
    CHX_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * (ncell + 1), ctx->stream));
    k_cell_count<<<chx_div_up(n, 256), 256, 0, ctx->stream>>>(x, n, box, g, cell_of, count);
    CHX_LAUNCHED(ctx);
    k_cell_scan<<<1, 1024, 0, ctx->stream>>>(count, start, ncell);
    CHX_LAUNCHED(ctx);
    k_cell_fill<<<chx_div_up(n, 256), 256, 0, ctx->stream>>>(x, cell_of, n, box, start, count, order, xs4);
    CHX_LAUNCHED(ctx);
    // bitmap sweep while one warp's bitmap (n bits, padded) fits next to seven others in shared memory
    {
        const int nwords = chx_div_up(n, 32);
        int wshift = 0;
        while ((32 << wshift) < nwords) ++wshift;
        const size_t per_warp = (size_t)32 * ((1u << wshift) + 1) * sizeof(uint32_t);
        static int use_bm = -1;
        if (use_bm < 0) { const char* e = getenv("CHX_NLIST_SORT_SWEEP"); use_bm = (e && e[0] == '1') ? 0 : 1; }
        if (use_bm && per_warp * 2 <= 200 * 1024) {
            SweepConst sc;
            sc.c = cutoff_plus_skin;
            const double c2 = (double)cutoff_plus_skin * (double)cutoff_plus_skin;
            const double lmax = fmax(lx, fmax(ly, lz));
            const double band = c2 * 4e-6 + 8.0 * 1.2e-7 * lmax * cutoff_plus_skin;
            sc.c2_lo = (float)(c2 - band); sc.c2_hi = (float)(c2 + band);
            sc.inv_lx = 1.0f / lx; sc.inv_ly = 1.0f / ly; sc.inv_lz = 1.0f / lz;
            int warps = 8;
            while (warps > 2 && per_warp * warps > 200 * 1024) warps >>= 1;
            const size_t smem = per_warp * warps;
#define BM_LAUNCH(WN)                                                                                   \
            do {                                                                                        \
                CHX_CUDA(cudaFuncSetAttribute(k_build_cell_bm<WN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                k_build_cell_bm<WN><<<chx_div_up(n, WN), WN * 32, smem, ctx->stream>>>(                   \
                    x, xs4, n, box, g, sc, M, wshift, cell_of, start, neighbor_list, neighbor_mask, n_neighbors); \
            } while (0)
            if (warps == 8) BM_LAUNCH(8); else if (warps == 4) BM_LAUNCH(4); else BM_LAUNCH(2);
#undef BM_LAUNCH
            CHX_LAUNCHED(ctx);
            return finish_build(ctx, n_neighbors, n, M, max_count_host, count_eq_M_host, nullptr, nullptr);
        }
    }
    int cap = 64;
    while (cap < 2 * M && cap < 2048) cap <<= 1;
    int ovf = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        const int warps = cap > 1024 ? 4 : 8;
        const size_t smem = (size_t)warps * cap * sizeof(uint32_t);
        if (warps == 8)
            k_build_cell<8><<<chx_div_up(n, 8), 256, smem, ctx->stream>>>(
                x, n, box, g, cutoff_plus_skin, M, cap, cell_of, start, order, neighbor_list,
                neighbor_mask, n_neighbors, overflow);
        else
            k_build_cell<4><<<chx_div_up(n, 4), 128, smem, ctx->stream>>>(
                x, n, box, g, cutoff_plus_skin, M, cap, cell_of, start, order, neighbor_list,
                neighbor_mask, n_neighbors, overflow);
        CHX_LAUNCHED(ctx);
        int rc2 = finish_build(ctx, n_neighbors, n, M, max_count_host, count_eq_M_host, overflow, &ovf);
        if (rc2 != CHX_OK) return rc2;
        if (!ovf) return CHX_OK;
        if (cap >= 2048) break;
        cap = 2048;
        CHX_CUDA(cudaMemsetAsync(overflow, 0, sizeof(int), ctx->stream));
    }
    // a row holds more than 2048 neighbours: use the O(N^2) kernel (streams rows, no staging)
    return chx_nlist_build_nsq(ctx, x, n, lx, ly, lz, periodic, cutoff_plus_skin, M, neighbor_list,
                               neighbor_mask, n_neighbors, max_count_host, count_eq_M_host);
}

int chx_nlist_calculate(chx_ctx* ctx, const float* x, int n, float lx, float ly, float lz,
                        int periodic, float cutoff, int M, const uint32_t* neighbor_list,
                        const int32_t* neighbor_mask, int32_t* n_out, int32_t* mask_out,
                        float* dist_out, float* rij_out) {
    CHX_REQUIRE(ctx && x && neighbor_list && neighbor_mask && n_out && mask_out && dist_out && rij_out,
                "NULL argument");
    CHX_REQUIRE(n > 0 && M > 0, "n and M must be positive");
    Box box = make_box(lx, ly, lz);
    const int blocks = chx_div_up(n, WARPS_PER_BLOCK);
    if (periodic)
        k_nlist_calculate<true><<<blocks, WARPS_PER_BLOCK * 32, 0, ctx->stream>>>(
            x, n, box, cutoff, M, neighbor_list, neighbor_mask, n_out, mask_out, dist_out, rij_out);
    else
        k_nlist_calculate<false><<<blocks, WARPS_PER_BLOCK * 32, 0, ctx->stream>>>(
            x, n, box, cutoff, M, neighbor_list, neighbor_mask, n_out, mask_out, dist_out, rij_out);
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

int chx_nlist_check(chx_ctx* ctx, const float* x, const float* ref_x, int n, float lx, float ly,
                    float lz, int periodic, float half_skin, int32_t* flag_dev) {
    CHX_REQUIRE(ctx && x && ref_x && flag_dev, "NULL argument");
    CHX_REQUIRE(n > 0, "n must be positive");
    CHX_CUDA(cudaMemsetAsync(flag_dev, 0, sizeof(int32_t), ctx->stream));
    Box box = make_box(lx, ly, lz);
    if (periodic)
        k_nlist_check<true><<<chx_div_up(n, 256), 256, 0, ctx->stream>>>(x, ref_x, n, box, half_skin, flag_dev);
    else
        k_nlist_check<false><<<chx_div_up(n, 256), 256, 0, ctx->stream>>>(x, ref_x, n, box, half_skin, flag_dev);
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

int chx_pairlist_build(chx_ctx* ctx, int n, uint32_t* all_pairs, uint8_t* reduction_mask) {
    CHX_REQUIRE(ctx && all_pairs && reduction_mask, "NULL argument");
    CHX_REQUIRE(n > 1, "n must be > 1");
    const long long total = (long long)n * (n - 1);
    k_pairlist_build<<<chx_div_up(total, 256), 256, 0, ctx->stream>>>(n, all_pairs, reduction_mask);
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

int chx_pairlist_calculate(chx_ctx* ctx, const float* x, int n, float lx, float ly, float lz,
                           int periodic, float cutoff, int32_t* n_out, int32_t* mask_out,
                           float* dist_out, float* rij_out) {
    CHX_REQUIRE(ctx && x && n_out && mask_out && dist_out && rij_out, "NULL argument");
    CHX_REQUIRE(n > 1, "n must be > 1");
    Box box = make_box(lx, ly, lz);
    const bool use_cutoff = cutoff >= 0.0f;
    const int blocks = chx_div_up(n, WARPS_PER_BLOCK);
    if (periodic)
        k_pairlist_calculate<true><<<blocks, WARPS_PER_BLOCK * 32, 0, ctx->stream>>>(
            x, n, box, cutoff, use_cutoff, n_out, mask_out, dist_out, rij_out);
    else
        k_pairlist_calculate<false><<<blocks, WARPS_PER_BLOCK * 32, 0, ctx->stream>>>(
            x, n, box, cutoff, use_cutoff, n_out, mask_out, dist_out, rij_out);
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

}  // extern "C"
