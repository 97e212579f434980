// Reference-shaped pair finding: NeighborListNsqrd (O(N^2) restatement and O(N) cell list that
// emits the identical arrays), calculate, check, PairListNsqrd.
//   chiron/neighbors.py:513-907 (neighbor list), :1018-1289 (pair list)
#include <stdlib.h>
#include <math.h>
#include <string.h>
#include "common.cuh"
#include "cell_list.cuh"

#define WARPS_PER_BLOCK 8

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
    for (int dx = -g.S; dx <= g.S; ++dx) {
        int ax = cx + dx; ax = ax < 0 ? ax + g.nx : (ax >= g.nx ? ax - g.nx : ax);
        for (int dy = -g.S; dy <= g.S; ++dy) {
            int ay = cy + dy; ay = ay < 0 ? ay + g.ny : (ay >= g.ny ? ay - g.ny : ay);
            for (int dz = -g.S; dz <= g.S; ++dz) {
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
        uint32_t j = list[o];
        // padding ids can point one past the end (a row without neighbours pads with 0, bumped to 1 when
        // i == 0: neighbors.py:606-609 on a single particle); JAX clamps the gather index
        j = j < (uint32_t)n ? j : (uint32_t)(n - 1);
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
    CellParams P;
    if (!periodic || !make_cell_params(lx, ly, lz, cutoff_plus_skin, 0, P))
        return chx_nlist_build_nsq(ctx, x, n, lx, ly, lz, periodic, cutoff_plus_skin, M,
                                   neighbor_list, neighbor_mask, n_neighbors, max_count_host,
                                   count_eq_M_host);
    const CellGrid g = P.g;
    const int ncell = P.ncell;
    const Box box = P.box;
    // scratch layout: [stats 256B][overflow flag 256B][CellParams 256B][xs4 n float4][cell_of n][count ncell+1][start ncell+1][order n]
    const size_t ints = 192 + 4 * (size_t)n + (size_t)n + 2 * ((size_t)ncell + 1) + (size_t)n;
    int* base = (int*)chx_scratch(ctx, ints * sizeof(int));
    if (!base) return CHX_CUDA_ERROR;
    int* overflow = base + 64;
    CellParams* P_dev = reinterpret_cast<CellParams*>(base + 128);
    float4* xs4 = reinterpret_cast<float4*>(base + 192);     // 768 B into a 256 B aligned block
    int* cell_of = base + 192 + 4 * (size_t)n;
    int* count = cell_of + n;
    int* start = count + ncell + 1;
    int* order = start + ncell + 1;
    // the build kernels read their geometry from device memory (shared with the barostat loop, mc.cu)
    static_assert(sizeof(CellParams) <= 256, "CellParams must fit its scratch slot");
    memcpy(ctx->host_pinned + 8, &P, sizeof(P));
    CHX_CUDA(cudaMemcpyAsync(P_dev, ctx->host_pinned + 8, sizeof(P), cudaMemcpyHostToDevice, ctx->stream));
    CHX_CUDA(cudaMemsetAsync(overflow, 0, sizeof(int), ctx->stream));
    CHX_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * (ncell + 1), ctx->stream));
    k_cell_count<<<chx_div_up(n, 256), 256, 0, ctx->stream>>>(x, n, P_dev, cell_of, count);
    CHX_LAUNCHED(ctx);
    k_cell_scan<<<1, 1024, 0, ctx->stream>>>(count, start, P_dev);
    CHX_LAUNCHED(ctx);
    k_cell_fill<<<chx_div_up(n, 256), 256, 0, ctx->stream>>>(x, cell_of, n, P_dev, start, count, order, xs4);
    CHX_LAUNCHED(ctx);
    // bitmap sweep while one warp's bitmap (n bits, padded) fits next to another one in shared memory
    {
        int wshift, warps;
        size_t smem;
        static int use_bm = -1;
        if (use_bm < 0) { const char* e = getenv("CHX_NLIST_SORT_SWEEP"); use_bm = (e && e[0] == '1') ? 0 : 1; }
        if (use_bm && cell_bm_config(n, wshift, warps, smem)) {
            int rc = cell_bm_launch(ctx, x, xs4, n, P_dev, M, wshift, warps, smem, cell_of, start, neighbor_list,
                                    neighbor_mask, n_neighbors);
            if (rc != CHX_OK) return rc;
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
