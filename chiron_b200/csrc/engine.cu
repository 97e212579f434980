// Fused LJ Langevin engine: the throughput path behind LangevinIntegrator.run for
// LJPotential + NeighborListNsqrd under periodic boundaries
//   chiron/integrators.py:110-218, chiron/neighbors.py:548-907, chiron/potential.py:193-300.
//
// Data layout in HBM (per replica r, all arrays padded to NP = 32*ceil(N/32)):
//   xs   float4  position, w = original particle id (int bits; -1 marks padding)
//   vs   float4  velocity, w = mass
//   fs   float4  force,    w = per-particle potential energy (half of each pair)
//   refi float4  positions at the last internal rebuild     (drives the engine's own rebuild)
//   refu float4  positions at the last REFERENCE rebuild    (what nbr_list.ref_positions holds)
// Particles are kept sorted by the Morton code of a fine cell grid (~4 particles per cell), so a
// "block" of 32 consecutive particles is a compact blob handled by one warp.
//
// Neighbour structure ("tiles"): for every block a table of candidate j particles (all particles
// within cutoff+skin of any particle of the block), grouped 32 to a tile.  A tile is 256 bytes:
// 32 x {j id | periodic image code << 24} followed by 32 x 32-bit masks, one mask per i lane, bit k
// set iff pair (i, j_k) was within cutoff+skin at build time with the reference's exact fp32
// predicate.  The force kernel loads a tile with two coalesced 128 B reads, gathers the 32 j
// positions once, and every lane walks only ITS set bits, fetching j coordinates with warp shuffles
// -- no per-pair memory traffic, no atomics, no padding work.  Both (i,j) and (j,i) are listed, so
// f_i is complete in registers and is written with one coalesced store.
#include <vector>
#include "common.cuh"

#define FULL 0xffffffffu
#define IMG_NONE 13u  // image code of "no shift": (1,1,1) in base 3
#define HALT_NONE 0x7f7f7f7f  // cudaMemset-able "no halt" value of MdCtrl::halt_step

struct MdCtrl {
    int halt_step;       // first step whose BAOAB update invalidated the internal tables
                         // (HALT_NONE = none); kernels of later steps turn into no-ops
    int pad1;
    int overflow;        // table capacity exceeded during a build
    int pad0;
    unsigned long long cand_pairs2;  // sum of mask popcounts at the last build (= 2 * P_cand)
    unsigned long long int_pairs2;   // directed interacting pairs seen by the last energy kernel
};

struct MdRep {            // per replica control block
    uint32_t key[2][2];   // loop key, double buffered by step parity
    int user_step;        // last step at which the reference rebuild condition fired
    int user_rebuilds;
    float kT;
    int pad;
};

struct MdGeom {
    Box box;
    float inv_lx, inv_ly, inv_lz;
    int ncx, ncy, ncz, bits;      // fine cell grid, Morton bits per dimension
    float inv_cx, inv_cy, inv_cz; // 1 / cell edge
    float cx, cy, cz;             // cell edge
    int n, np, nblk, ncm;         // particles, padded, blocks, Morton cells
};

struct chx_ljmd {
    chx_ctx* ctx;
    chx_ljmd_params p;
    MdGeom g;
    int R;
    int cur;                 // which half of the double buffers is live
    float4 *xs[2], *vs[2], *refu[2];
    float4 *fs, *refi;
    int *cell_count, *cell_start, *cell_of, *order;
    uint32_t* tiles;
    int* ntiles;
    uint8_t* generic;
    float4* bcenter;          // per block: centre of its bounding box at the last build
    int tcap;
    MdCtrl* ctrl;
    MdRep* rep;
    MdCtrl* ctrl_host;       // pinned
    MdRep* rep_host;         // pinned (R entries)
    double* e_scratch;       // R doubles
    float internal_skin;
    long long rebuilds, steps, launches0;
    bool have_state;
};

// ---------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t spread3(uint32_t v) {
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__host__ __device__ __forceinline__ uint32_t morton3(int x, int y, int z) {
    return spread3((uint32_t)x) | (spread3((uint32_t)y) << 1) | (spread3((uint32_t)z) << 2);
}

__device__ __forceinline__ int fine_cell(float x, float inv_c, int nc) {
    int c = (int)floorf(x * inv_c);
    c = c < 0 ? 0 : c;
    return c >= nc ? nc - 1 : c;
}

__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

// ---------------------------------------------------------------------------------------------
// import / export between the caller's (N,3) arrays and the sorted float4 state
// ---------------------------------------------------------------------------------------------
__global__ void k_md_import(const float* __restrict__ x, const float* __restrict__ v,
                            const float* __restrict__ mass, MdGeom g, float4* __restrict__ xs,
                            float4* __restrict__ vs, float4* __restrict__ refu) {
    const int r = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.np) return;
    const size_t o = (size_t)r * g.np + i;
    if (i >= g.n) {
        xs[o] = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
        vs[o] = make_float4(0.f, 0.f, 0.f, 1.f);
        refu[o] = xs[o];
        return;
    }
    const size_t s = ((size_t)r * g.n + i) * 3;
    float px = x[s], py = x[s + 1], pz = x[s + 2];
    refu[o] = make_float4(px, py, pz, __int_as_float(i));
    // positions outside [0, L) are wrapped on import (the sort and the image codes assume it)
    if (px < 0.f || px >= g.box.lx) px = ref_wrap(px, g.box.lx);
    if (py < 0.f || py >= g.box.ly) py = ref_wrap(py, g.box.ly);
    if (pz < 0.f || pz >= g.box.lz) pz = ref_wrap(pz, g.box.lz);
    xs[o] = make_float4(px, py, pz, __int_as_float(i));
    vs[o] = make_float4(v[s], v[s + 1], v[s + 2], mass[i]);
}

__global__ void k_md_export_ids(const float4* __restrict__ src, const float4* __restrict__ xs,
                                MdGeom g, float* __restrict__ dst) {
    const int r = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.np) return;
    const size_t o = (size_t)r * g.np + i;
    const int id = __float_as_int(xs[o].w);
    if (id < 0) return;
    const float4 a = src[o];
    const size_t d = ((size_t)r * g.n + id) * 3;
    dst[d] = a.x; dst[d + 1] = a.y; dst[d + 2] = a.z;
}

// ---------------------------------------------------------------------------------------------
// sort: Morton cell id -> counting sort -> per-cell order by original id (deterministic)
// ---------------------------------------------------------------------------------------------
__global__ void k_md_cellcount(const float4* __restrict__ xs, MdGeom g, int* __restrict__ cell_of,
                               int* __restrict__ count) {
    const int r = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.np) return;
    const size_t o = (size_t)r * g.np + i;
    const float4 p = xs[o];
    if (__float_as_int(p.w) < 0) { cell_of[o] = -1; return; }
    const int m = (int)morton3(fine_cell(p.x, g.inv_cx, g.ncx), fine_cell(p.y, g.inv_cy, g.ncy),
                               fine_cell(p.z, g.inv_cz, g.ncz));
    cell_of[o] = m;
    atomicAdd(&count[(size_t)r * (g.ncm + 1) + m], 1);
}

__global__ void __launch_bounds__(1024)
k_md_scan(int* __restrict__ count, int* __restrict__ start, int ncm) {
    __shared__ int part[1024];
    const int r = blockIdx.x;
    count += (size_t)r * (ncm + 1);
    start += (size_t)r * (ncm + 1);
    const int t = threadIdx.x;
    const int per = (ncm + 1023) / 1024;
    const int lo = t * per, hi = min(ncm, lo + per);
    int s = 0;
    for (int c = lo; c < hi; ++c) s += count[c];
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        int v = t >= o ? part[t - o] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    int run = part[t] - s;
    for (int c = lo; c < hi; ++c) {
        const int k = count[c];
        start[c] = run;
        count[c] = 0;
        run += k;
    }
    if (t == 1023) start[ncm] = part[t];
}

__global__ void k_md_place(const int* __restrict__ cell_of, MdGeom g, const int* __restrict__ start,
                           int* __restrict__ cursor, int* __restrict__ order) {
    const int r = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.np) return;
    const size_t o = (size_t)r * g.np + i;
    const int m = cell_of[o];
    if (m < 0) return;
    const size_t cb = (size_t)r * (g.ncm + 1);
    order[(size_t)r * g.np + start[cb + m] + atomicAdd(&cursor[cb + m], 1)] = i;
}

// one thread per Morton cell: insertion sort of the cell's slots by original particle id
__global__ void k_md_cellsort(const float4* __restrict__ xs, MdGeom g, const int* __restrict__ start,
                              int* __restrict__ order) {
    const int r = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g.ncm) return;
    const size_t cb = (size_t)r * (g.ncm + 1);
    const int s = start[cb + c], e = start[cb + c + 1];
    if (e - s < 2) return;
    int* ord = order + (size_t)r * g.np;
    const float4* x = xs + (size_t)r * g.np;
    for (int a = s + 1; a < e; ++a) {
        const int oa = ord[a];
        const int ka = __float_as_int(x[oa].w);
        int b = a - 1;
        while (b >= s && __float_as_int(x[ord[b]].w) > ka) { ord[b + 1] = ord[b]; --b; }
        ord[b + 1] = oa;
    }
}

__global__ void k_md_gather(const int* __restrict__ order, MdGeom g, const float4* __restrict__ xs0,
                            const float4* __restrict__ vs0, const float4* __restrict__ ru0,
                            float4* __restrict__ xs1, float4* __restrict__ vs1,
                            float4* __restrict__ ru1, float4* __restrict__ refi) {
    const int r = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.np) return;
    const size_t o = (size_t)r * g.np + p;
    if (p >= g.n) {
        const float4 pad = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
        xs1[o] = pad; vs1[o] = make_float4(0.f, 0.f, 0.f, 1.f); ru1[o] = pad; refi[o] = pad;
        return;
    }
    const size_t q = (size_t)r * g.np + order[o];
    const float4 a = xs0[q];
    xs1[o] = a; vs1[o] = vs0[q]; ru1[o] = ru0[q]; refi[o] = a;
}

// ---------------------------------------------------------------------------------------------
// table build: one warp per block of 32 particles
// ---------------------------------------------------------------------------------------------
#define BW 4          // warps per CTA in the build kernel
#define QCAP 2048     // candidate queue entries per warp

__device__ __forceinline__ float axis_gap(float v, float lo, float hi) {
    return fmaxf(0.f, fmaxf(lo - v, v - hi));
}

__global__ void __launch_bounds__(BW * 32)
k_md_build(const float4* __restrict__ xs_all, const int* __restrict__ start_all, MdGeom g, float R,
           int tcap, uint32_t* __restrict__ tiles_all, int* __restrict__ ntiles_all,
           uint8_t* __restrict__ generic_all, float4* __restrict__ bcenter_all, float drift,
           MdCtrl* __restrict__ ctrl) {
    __shared__ uint32_t queue[BW][QCAP];
    __shared__ uint32_t idxbuf[BW][32];
    const int r = blockIdx.y;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * BW + w;
    if (b >= g.nblk) return;
    const float4* xs = xs_all + (size_t)r * g.np;
    const int* start = start_all + (size_t)r * (g.ncm + 1);
    uint32_t* tiles = tiles_all + ((size_t)r * g.nblk + b) * tcap * 64;
    const int i = b * 32 + lane;
    const float4 xi = xs[i];
    const bool valid = __float_as_int(xi.w) >= 0;
    const float INF = __int_as_float(0x7f800000);
    const float lox = warp_min(valid ? xi.x : INF), hix = warp_max(valid ? xi.x : -INF);
    const float loy = warp_min(valid ? xi.y : INF), hiy = warp_max(valid ? xi.y : -INF);
    const float loz = warp_min(valid ? xi.z : INF), hiz = warp_max(valid ? xi.z : -INF);
    const float Rm = R * (1.0f + 2e-5f) + 1e-6f;
    int c0x = (int)floorf((lox - Rm) * g.inv_cx), c1x = (int)floorf((hix + Rm) * g.inv_cx);
    int c0y = (int)floorf((loy - Rm) * g.inv_cy), c1y = (int)floorf((hiy + Rm) * g.inv_cy);
    int c0z = (int)floorf((loz - Rm) * g.inv_cz), c1z = (int)floorf((hiz + Rm) * g.inv_cz);
    int nx = c1x - c0x + 1, ny = c1y - c0y + 1, nz = c1z - c0z + 1;
    bool gen = false;
    if (nx >= g.ncx) { nx = g.ncx; c0x = 0; gen = true; }
    if (ny >= g.ncy) { ny = g.ncy; c0y = 0; gen = true; }
    if (nz >= g.ncz) { nz = g.ncz; c0z = 0; gen = true; }
    // non-generic blocks resolve periodic images per TILE (force kernel): everything the block can
    // interact with until the next rebuild must stay within half a box of the block centre
    if (0.5f * (hix - lox) + Rm + drift >= g.box.hx || 0.5f * (hiy - loy) + Rm + drift >= g.box.hy ||
        0.5f * (hiz - loz) + Rm + drift >= g.box.hz)
        gen = true;
    const int total = nx * ny * nz;
    const float Rm2 = Rm * Rm;

    int qn = 0;        // queue fill (warp uniform)
    int nslots = 0;    // table slots used (warp uniform)
    uint32_t cur = 0;  // mask bits of the tile being filled
    bool ovf = false;
    unsigned long long pairs = 0;

    // drains the queue: every candidate is tested against the 32 particles of the block
    auto drain = [&]() {
        __syncwarp();
        for (int q = 0; q < qn; ++q) {
            const uint32_t code = queue[w][q];
            const int p = (int)(code & 0xffffffu);
            const float4 xj = xs[p];
            bool hit = false;
            if (valid && p != i) {
                float rx, ry, rz, d;
                ref_displacement<true>(xi.x, xi.y, xi.z, xj.x, xj.y, xj.z, g.box, rx, ry, rz, d);
                hit = d < R;
            }
            const unsigned bal = __ballot_sync(FULL, hit);
            if (bal) {
                const int k = nslots & 31;
                if (lane == 0) idxbuf[w][k] = code;
                if (hit) { cur |= 1u << k; ++pairs; }
                ++nslots;
                if (k == 31) {
                    __syncwarp();
                    const int t = (nslots >> 5) - 1;
                    if (t < tcap) {
                        tiles[(size_t)t * 64 + lane] = idxbuf[w][lane];
                        tiles[(size_t)t * 64 + 32 + lane] = cur;
                    } else {
                        ovf = true;
                    }
                    cur = 0;
                    __syncwarp();
                }
            }
        }
        qn = 0;
        __syncwarp();
    };

    for (int cb = 0; cb < total; cb += 32) {
        const int c = cb + lane;
        int s = 0, e = 0;
        uint32_t img = IMG_NONE;
        float shx = 0.f, shy = 0.f, shz = 0.f;
        if (c < total) {
            const int iz = c % nz, iy = (c / nz) % ny, ix = c / (nz * ny);
            int ux = c0x + ix, uy = c0y + iy, uz = c0z + iz;  // unwrapped cell coordinates
            int wx = ux, wy = uy, wz = uz, sx = 1, sy = 1, sz = 1;
            if (wx < 0) { wx += g.ncx; sx = 0; } else if (wx >= g.ncx) { wx -= g.ncx; sx = 2; }
            if (wy < 0) { wy += g.ncy; sy = 0; } else if (wy >= g.ncy) { wy -= g.ncy; sy = 2; }
            if (wz < 0) { wz += g.ncz; sz = 0; } else if (wz >= g.ncz) { wz -= g.ncz; sz = 2; }
            // cell-level prefilter in unwrapped coordinates
            const float gx = axis_gap((ux + 0.5f) * g.cx, lox - 0.5f * g.cx, hix + 0.5f * g.cx);
            const float gy = axis_gap((uy + 0.5f) * g.cy, loy - 0.5f * g.cy, hiy + 0.5f * g.cy);
            const float gz = axis_gap((uz + 0.5f) * g.cz, loz - 0.5f * g.cz, hiz + 0.5f * g.cz);
            if (gen || gx * gx + gy * gy + gz * gz < Rm2 * 1.0001f) {
                const int m = (int)morton3(wx, wy, wz);
                s = start[m]; e = start[m + 1];
                img = (uint32_t)(sx * 9 + sy * 3 + sz);
                shx = (sx - 1) * g.box.lx; shy = (sy - 1) * g.box.ly; shz = (sz - 1) * g.box.lz;
            }
        }
        // lock-step over the cells' particles so queue positions are deterministic
        int maxlen = e - s;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(FULL, maxlen, o));
        for (int k = 0; k < maxlen; ++k) {
            bool push = false;
            const int p = s + k;
            if (p < e) {
                if (gen) {
                    push = true;  // small box: no image bookkeeping, exact test decides
                } else {
                    const float4 xj = xs[p];
                    const float dx = axis_gap(xj.x + shx, lox, hix);
                    const float dy = axis_gap(xj.y + shy, loy, hiy);
                    const float dz = axis_gap(xj.z + shz, loz, hiz);
                    push = dx * dx + dy * dy + dz * dz < Rm2;
                }
            }
            const unsigned bal = __ballot_sync(FULL, push);
            if (push) queue[w][qn + __popc(bal & ((1u << lane) - 1u))] = (uint32_t)p | (img << 24);
            qn += __popc(bal);
            if (qn + 32 > QCAP) drain();
        }
    }
    drain();
    // flush the partially filled tile, padding with the lane's own particle and empty masks
    const int k = nslots & 31;
    if (k != 0) {
        __syncwarp();
        const int t = nslots >> 5;
        if (t < tcap) {
            tiles[(size_t)t * 64 + lane] = lane < k ? idxbuf[w][lane] : ((uint32_t)(b * 32) | (IMG_NONE << 24));
            tiles[(size_t)t * 64 + 32 + lane] = cur;
        } else {
            ovf = true;
        }
    }
    if (lane == 0) {
        ntiles_all[(size_t)r * g.nblk + b] = min((nslots + 31) >> 5, tcap);
        generic_all[(size_t)r * g.nblk + b] = gen ? 1 : 0;
        bcenter_all[(size_t)r * g.nblk + b] = make_float4(0.5f * (lox + hix), 0.5f * (loy + hiy), 0.5f * (loz + hiz), 0.f);
        if (ovf) atomicExch(&ctrl->overflow, 1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pairs += __shfl_xor_sync(FULL, pairs, o);
    if (lane == 0 && pairs) atomicAdd(&ctrl->cand_pairs2, pairs);
}

// ---------------------------------------------------------------------------------------------
// force / energy over the tiles
// ---------------------------------------------------------------------------------------------
struct LjConst {
    float sig2;      // sigma^2
    float eps24;     // 24 eps
    float eps4;      // 4 eps
    float rc;        // cutoff (exact predicate)
    float rc2_lo, rc2_hi;  // guard band around cutoff^2 for the fast predicate
};

#define FW 8  // warps per CTA in the force kernel

template <bool ENERGY>
__global__ void __launch_bounds__(FW * 32)
k_md_force(const float4* __restrict__ xs_all, float4* __restrict__ fs_all,
           float4* __restrict__ refu_all, const uint32_t* __restrict__ tiles_all,
           const int* __restrict__ ntiles_all, const uint8_t* __restrict__ generic_all,
           const float4* __restrict__ bcenter_all, MdGeom g, LjConst lj, int tcap, MdCtrl* __restrict__ ctrl, MdRep* __restrict__ rep, int step,
           int honor_halt, double* __restrict__ energy_out) {
    __shared__ uint32_t sj[FW][32];
    __shared__ double red[FW];
    __shared__ unsigned long long redn[FW];
    // the tables are stale from step halt_step on: that step's forces are evaluated after the rebuild
    if (honor_halt && *((volatile int*)&ctrl->halt_step) <= step) return;
    const int r = blockIdx.y;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * FW + w;
    float e_acc = 0.f;
    unsigned npair = 0;
    if (b < g.nblk) {
        const float4* xs = xs_all + (size_t)r * g.np;
        const int i = b * 32 + lane;
        const float4 xi0 = xs[i];
        float4 xi = xi0;
        if (rep[r].user_step == step) {  // reference rebuild (neighbors.py:903-905 -> build)
            refu_all[(size_t)r * g.np + i] = xi0;
            if (b == 0 && lane == 0) rep[r].user_rebuilds++;
        }
        const uint32_t* tp = tiles_all + ((size_t)r * g.nblk + b) * tcap * 64;
        const int nt = ntiles_all[(size_t)r * g.nblk + b];
        const bool gen = generic_all[(size_t)r * g.nblk + b] != 0;
        // positions are wrapped every step (integrators.py:239), so a particle may have jumped by a
        // box length since the build: images are resolved against the block centre, once per
        // particle per tile instead of once per pair
        const float4 bc = bcenter_all[(size_t)r * g.nblk + b];
        if (!gen) {
            xi.x -= g.box.lx * rintf((xi.x - bc.x) * g.inv_lx);
            xi.y -= g.box.ly * rintf((xi.y - bc.y) * g.inv_ly);
            xi.z -= g.box.lz * rintf((xi.z - bc.z) * g.inv_lz);
        }
        float fx = 0.f, fy = 0.f, fz = 0.f;
        for (int t = 0; t < nt; ++t, tp += 64) {
            const uint32_t code = tp[lane];
            uint32_t m = tp[32 + lane];
            const int j = (int)(code & 0xffffffu);
            float4 xj = xs[j];
            __syncwarp();
            sj[w][lane] = (uint32_t)j;
            if (!gen) {
                xj.x -= g.box.lx * rintf((xj.x - bc.x) * g.inv_lx);
                xj.y -= g.box.ly * rintf((xj.y - bc.y) * g.inv_ly);
                xj.z -= g.box.lz * rintf((xj.z - bc.z) * g.inv_lz);
            }
            __syncwarp();
            while (__any_sync(FULL, m != 0u)) {
                const bool act = m != 0u;
                const int bit = act ? (__ffs(m) - 1) : 0;
                m &= m - 1u;
                const float sx = __shfl_sync(FULL, xj.x, bit);
                const float sy = __shfl_sync(FULL, xj.y, bit);
                const float sz = __shfl_sync(FULL, xj.z, bit);
                float dx = xi.x - sx, dy = xi.y - sy, dz = xi.z - sz;
                if (gen) {
                    dx -= g.box.lx * rintf(dx * g.inv_lx);
                    dy -= g.box.ly * rintf(dy * g.inv_ly);
                    dz -= g.box.lz * rintf(dz * g.inv_lz);
                }
                const float r2 = dx * dx + dy * dy + dz * dz;
                bool in = act && r2 < lj.rc2_hi;
                if (in && r2 > lj.rc2_lo) {
                    // within a few ulps of the cutoff: decide with the reference's exact predicate
                    const float4 xo = xs[sj[w][bit]];
                    float rx, ry, rz, d;
                    ref_displacement<true>(xi0.x, xi0.y, xi0.z, xo.x, xo.y, xo.z, g.box, rx, ry, rz, d);
                    in = d < lj.rc;
                }
                if (in) {
                    const float inv = rcp_approx(r2);
                    const float s2 = lj.sig2 * inv;
                    const float s6 = s2 * s2 * s2;
                    const float f = (lj.eps24 * inv) * (s6 * (2.0f * s6 - 1.0f));
                    fx += f * dx; fy += f * dy; fz += f * dz;
                    if (ENERGY) { e_acc += lj.eps4 * (s6 * (s6 - 1.0f)); ++npair; }
                }
            }
        }
        fs_all[(size_t)r * g.np + i] = make_float4(fx, fy, fz, 0.5f * e_acc);
    }
    if (ENERGY) {
        double e = warp_sum((double)e_acc * 0.5);
        int np = warp_sum((int)npair);
        if (lane == 0) { red[w] = e; redn[w] = (unsigned long long)np; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            unsigned long long c = 0;
            for (int k = 0; k < FW; ++k) { s += red[k]; c += redn[k]; }
            if (s != 0.0) atomicAdd(&energy_out[r], s);
            if (c) atomicAdd(&ctrl->int_pairs2, c);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// fused BAOAB (+ trailing B of the previous step) + wrap + both rebuild checks
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_md_baoab(float4* __restrict__ xs_all, float4* __restrict__ vs_all, const float4* __restrict__ fs_all,
           const float4* __restrict__ refi_all, const float4* __restrict__ refu_all, MdGeom g,
           float h, float a, float b, float half_skin_user, float half_skin_int2, int step,
           int trailing, MdCtrl* __restrict__ ctrl, MdRep* __restrict__ rep) {
    __shared__ uint32_t sk[2];
    // a halt raised by ANOTHER block of this same launch (halt_step == step) must not stop us
    if (*((volatile int*)&ctrl->halt_step) < step) return;
    const int r = blockIdx.y;
    if (threadIdx.x == 0) {
        uint32_t c0, c1, s0, s1;
        threefry_split(rep[r].key[step & 1][0], rep[r].key[step & 1][1], c0, c1, s0, s1);
        sk[0] = s0; sk[1] = s1;
        if (blockIdx.x == 0) { rep[r].key[(step + 1) & 1][0] = c0; rep[r].key[(step + 1) & 1][1] = c1; }
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool moved_int = false, moved_user = false;
    if (i < g.np) {
        const size_t o = (size_t)r * g.np + i;
        float4 x = xs_all[o];
        const int id = __float_as_int(x.w);
        if (id >= 0) {
            float4 v = vs_all[o];
            const float4 f = fs_all[o];
            const float m = v.w;
            const float kT = rep[r].kT;
            const float bs = __fmul_rn(b, __fsqrt_rn(__fdiv_rn(kT, m)));
            const unsigned long long total = 3ull * (unsigned long long)g.n;
            const uint32_t k0 = sk[0], k1 = sk[1];
            float xc[3] = {x.x, x.y, x.z}, vc[3] = {v.x, v.y, v.z};
            const float fc[3] = {f.x, f.y, f.z};
            const float L[3] = {g.box.lx, g.box.ly, g.box.lz};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float kick = __fdiv_rn(__fmul_rn(h, fc[c]), m);
                if (trailing) vc[c] = __fadd_rn(vc[c], kick);
                vc[c] = __fadd_rn(vc[c], kick);
                xc[c] = __fadd_rn(xc[c], __fmul_rn(h, vc[c]));
                const float xi = normal_from_bits(random_bits_elem(k0, k1, 3ull * id + c, total));
                vc[c] = __fadd_rn(__fmul_rn(a, vc[c]), __fmul_rn(bs, xi));
                xc[c] = __fadd_rn(xc[c], __fmul_rn(h, vc[c]));
                xc[c] = ref_wrap(xc[c], L[c]);
            }
            xs_all[o] = make_float4(xc[0], xc[1], xc[2], x.w);
            vs_all[o] = make_float4(vc[0], vc[1], vc[2], m);
            // reference rebuild condition (exact, neighbors.py:864-868)
            const float4 ru = refu_all[o];
            float rx, ry, rz, d;
            ref_displacement<true>(xc[0], xc[1], xc[2], ru.x, ru.y, ru.z, g.box, rx, ry, rz, d);
            moved_user = d >= half_skin_user;
            // engine's own list validity (fast min-image)
            const float4 ri = refi_all[o];
            float dx = xc[0] - ri.x, dy = xc[1] - ri.y, dz = xc[2] - ri.z;
            dx -= g.box.lx * rintf(dx * g.inv_lx);
            dy -= g.box.ly * rintf(dy * g.inv_ly);
            dz -= g.box.lz * rintf(dz * g.inv_lz);
            moved_int = dx * dx + dy * dy + dz * dz >= half_skin_int2;
        }
    }
    const int any_int = __syncthreads_or(moved_int);
    const int any_user = __syncthreads_or(moved_user);
    if (threadIdx.x == 0) {
        if (any_int) atomicMin(&ctrl->halt_step, step);
        if (any_user) rep[r].user_step = step;
    }
}

__global__ void k_md_kick(float4* __restrict__ vs_all, const float4* __restrict__ fs_all, MdGeom g,
                          float h, int R) {
    const size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (o >= (size_t)R * g.np) return;
    float4 v = vs_all[o];
    const float4 f = fs_all[o];
    v.x = __fadd_rn(v.x, __fdiv_rn(__fmul_rn(h, f.x), v.w));
    v.y = __fadd_rn(v.y, __fdiv_rn(__fmul_rn(h, f.y), v.w));
    v.z = __fadd_rn(v.z, __fdiv_rn(__fmul_rn(h, f.z), v.w));
    vs_all[o] = v;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int md_alloc(chx_ljmd* md) {
    const size_t np = (size_t)md->R * md->g.np;
    for (int k = 0; k < 2; ++k) {
        CHX_CUDA(cudaMalloc(&md->xs[k], np * sizeof(float4)));
        CHX_CUDA(cudaMalloc(&md->vs[k], np * sizeof(float4)));
        CHX_CUDA(cudaMalloc(&md->refu[k], np * sizeof(float4)));
    }
    CHX_CUDA(cudaMalloc(&md->fs, np * sizeof(float4)));
    CHX_CUDA(cudaMalloc(&md->refi, np * sizeof(float4)));
    const size_t nc = (size_t)md->R * (md->g.ncm + 1);
    CHX_CUDA(cudaMalloc(&md->cell_count, nc * sizeof(int)));
    CHX_CUDA(cudaMalloc(&md->cell_start, nc * sizeof(int)));
    CHX_CUDA(cudaMalloc(&md->cell_of, np * sizeof(int)));
    CHX_CUDA(cudaMalloc(&md->order, np * sizeof(int)));
    const size_t nb = (size_t)md->R * md->g.nblk;
    CHX_CUDA(cudaMalloc(&md->tiles, nb * md->tcap * 64 * sizeof(uint32_t)));
    CHX_CUDA(cudaMalloc(&md->ntiles, nb * sizeof(int)));
    CHX_CUDA(cudaMalloc(&md->generic, nb));
    CHX_CUDA(cudaMalloc(&md->bcenter, nb * sizeof(float4)));
    CHX_CUDA(cudaMalloc(&md->ctrl, sizeof(MdCtrl)));
    CHX_CUDA(cudaMalloc(&md->rep, md->R * sizeof(MdRep)));
    CHX_CUDA(cudaMalloc(&md->e_scratch, md->R * sizeof(double)));
    CHX_CUDA(cudaMallocHost(&md->ctrl_host, sizeof(MdCtrl)));
    CHX_CUDA(cudaMallocHost(&md->rep_host, md->R * sizeof(MdRep)));
    CHX_CUDA(cudaMemset(md->ctrl, 0, sizeof(MdCtrl)));
    CHX_CUDA(cudaMemset(md->rep, 0, md->R * sizeof(MdRep)));
    CHX_CUDA(cudaMemset(md->fs, 0, np * sizeof(float4)));
    return CHX_OK;
}

static LjConst md_lj(const chx_ljmd* md) {
    LjConst c;
    c.sig2 = md->p.sigma * md->p.sigma;
    c.eps24 = 24.0f * md->p.epsilon;
    c.eps4 = 4.0f * md->p.epsilon;
    c.rc = md->p.cutoff;
    const double rc2 = (double)md->p.cutoff * (double)md->p.cutoff;
    c.rc2_lo = (float)(rc2 * (1.0 - 2e-6));
    c.rc2_hi = (float)(rc2 * (1.0 + 2e-6));
    return c;
}

// sort + table build from the live buffers; grows the table capacity on overflow
static int md_rebuild(chx_ljmd* md) {
    chx_ctx* ctx = md->ctx;
    const MdGeom& g = md->g;
    cudaStream_t st = ctx->stream;
    const int R = md->R;
    const dim3 gp(chx_div_up(g.np, 256), R);
    const int a = md->cur, b = 1 - md->cur;
    CHX_CUDA(cudaMemsetAsync(md->cell_count, 0, (size_t)R * (g.ncm + 1) * sizeof(int), st));
    k_md_cellcount<<<gp, 256, 0, st>>>(md->xs[a], g, md->cell_of, md->cell_count);
    CHX_LAUNCHED(ctx);
    k_md_scan<<<R, 1024, 0, st>>>(md->cell_count, md->cell_start, g.ncm);
    CHX_LAUNCHED(ctx);
    k_md_place<<<gp, 256, 0, st>>>(md->cell_of, g, md->cell_start, md->cell_count, md->order);
    CHX_LAUNCHED(ctx);
    k_md_cellsort<<<dim3(chx_div_up(g.ncm, 128), R), 128, 0, st>>>(md->xs[a], g, md->cell_start, md->order);
    CHX_LAUNCHED(ctx);
    k_md_gather<<<gp, 256, 0, st>>>(md->order, g, md->xs[a], md->vs[a], md->refu[a], md->xs[b],
                                    md->vs[b], md->refu[b], md->refi);
    CHX_LAUNCHED(ctx);
    md->cur = b;
    const float R_list = md->p.cutoff + md->internal_skin;
    for (int attempt = 0; attempt < 6; ++attempt) {
        CHX_CUDA(cudaMemsetAsync(&md->ctrl->overflow, 0, sizeof(int), st));
        CHX_CUDA(cudaMemsetAsync(&md->ctrl->cand_pairs2, 0, sizeof(unsigned long long), st));
        k_md_build<<<dim3(chx_div_up(g.nblk, BW), R), BW * 32, 0, st>>>(
            md->xs[b], md->cell_start, g, R_list, md->tcap, md->tiles, md->ntiles, md->generic, md->bcenter,
            md->internal_skin, md->ctrl);
        CHX_LAUNCHED(ctx);
        CHX_CUDA(cudaMemcpyAsync(md->ctrl_host, md->ctrl, sizeof(MdCtrl), cudaMemcpyDeviceToHost, st));
        CHX_CUDA(cudaStreamSynchronize(st));
        if (!md->ctrl_host->overflow) {
            md->rebuilds++;
            return CHX_OK;
        }
        md->tcap *= 2;
        CHX_CUDA(cudaFree(md->tiles));
        CHX_CUDA(cudaMalloc(&md->tiles, (size_t)R * g.nblk * md->tcap * 64 * sizeof(uint32_t)));
    }
    chx_set_error("neighbour table overflow: more than %d candidates per block", md->tcap * 32);
    return CHX_NEIGHBOR_OVERFLOW;
}

static int md_force(chx_ljmd* md, int step, bool energy, int honor_halt, double* e_dev) {
    const MdGeom& g = md->g;
    const dim3 gf(chx_div_up(g.nblk, FW), md->R);
    const int c = md->cur;
    if (energy)
        k_md_force<true><<<gf, FW * 32, 0, md->ctx->stream>>>(
            md->xs[c], md->fs, md->refu[c], md->tiles, md->ntiles, md->generic, md->bcenter, g, md_lj(md), md->tcap,
            md->ctrl, md->rep, step, honor_halt, e_dev);
    else
        k_md_force<false><<<gf, FW * 32, 0, md->ctx->stream>>>(
            md->xs[c], md->fs, md->refu[c], md->tiles, md->ntiles, md->generic, md->bcenter, g, md_lj(md), md->tcap,
            md->ctrl, md->rep, step, honor_halt, e_dev);
    CHX_LAUNCHED(md->ctx);
    return CHX_OK;
}

extern "C" {

int chx_ljmd_create(chx_ctx* ctx, const chx_ljmd_params* p, chx_ljmd** out) {
    CHX_REQUIRE(ctx && p && out, "NULL argument");
    CHX_REQUIRE(p->n > 0 && p->n < (1 << 24), "n must be in (0, 2^24)");
    CHX_REQUIRE(p->n_replicas >= 1, "n_replicas must be >= 1");
    CHX_REQUIRE(p->lx > 0 && p->ly > 0 && p->lz > 0, "box must be positive");
    CHX_REQUIRE(p->cutoff > 0 && p->skin >= 0, "cutoff must be positive, skin non-negative");
    chx_ljmd* md = new chx_ljmd();
    md->ctx = ctx;
    md->p = *p;
    md->R = p->n_replicas;
    md->cur = 0;
    md->internal_skin = (p->internal_skin > 0.f && p->internal_skin <= p->skin) ? p->internal_skin : p->skin;
    MdGeom& g = md->g;
    g.box = make_box(p->lx, p->ly, p->lz);
    g.inv_lx = 1.0f / p->lx; g.inv_ly = 1.0f / p->ly; g.inv_lz = 1.0f / p->lz;
    g.n = p->n;
    g.nblk = chx_div_up(p->n, 32);
    g.np = g.nblk * 32;
    // ~4 particles per fine cell, at most 64 cells per edge (6 Morton bits)
    const double vol = (double)p->lx * p->ly * p->lz;
    const double edge = cbrt(4.0 * vol / (double)p->n);
    auto pick = [&](double L) { int c = (int)floor(L / edge); return c < 1 ? 1 : (c > 64 ? 64 : c); };
    g.ncx = pick(p->lx); g.ncy = pick(p->ly); g.ncz = pick(p->lz);
    int mx = g.ncx > g.ncy ? g.ncx : g.ncy; mx = mx > g.ncz ? mx : g.ncz;
    g.bits = 1;
    while ((1 << g.bits) < mx) ++g.bits;
    g.ncm = 1 << (3 * g.bits);
    g.cx = p->lx / g.ncx; g.cy = p->ly / g.ncy; g.cz = p->lz / g.ncz;
    g.inv_cx = g.ncx / p->lx; g.inv_cy = g.ncy / p->ly; g.inv_cz = g.ncz / p->lz;
    // initial table capacity: 2x the mean number of candidates of a compact block
    const double rho = (double)p->n / vol;
    const double Rl = p->cutoff + md->internal_skin;
    const double a = cbrt(32.0 / rho);
    const double mink = a * a * a + 6 * a * a * Rl + 3 * M_PI * a * Rl * Rl + 4.0 / 3.0 * M_PI * Rl * Rl * Rl;
    double cand = mink * rho;
    if (cand > p->n) cand = p->n;
    md->tcap = (int)(2.0 * cand / 32.0) + 4;
    md->rebuilds = 0; md->steps = 0; md->have_state = false;
    md->launches0 = ctx->launches;
    int rc = md_alloc(md);
    if (rc != CHX_OK) { delete md; return rc; }
    *out = md;
    return CHX_OK;
}

int chx_ljmd_destroy(chx_ljmd* md) {
    if (!md) return CHX_OK;
    cudaStreamSynchronize(md->ctx->stream);
    for (int k = 0; k < 2; ++k) { cudaFree(md->xs[k]); cudaFree(md->vs[k]); cudaFree(md->refu[k]); }
    cudaFree(md->fs); cudaFree(md->refi); cudaFree(md->cell_count); cudaFree(md->cell_start);
    cudaFree(md->cell_of); cudaFree(md->order); cudaFree(md->tiles); cudaFree(md->ntiles);
    cudaFree(md->generic); cudaFree(md->bcenter); cudaFree(md->ctrl); cudaFree(md->rep); cudaFree(md->e_scratch);
    cudaFreeHost(md->ctrl_host); cudaFreeHost(md->rep_host);
    delete md;
    return CHX_OK;
}

int chx_ljmd_set_state(chx_ljmd* md, const float* x, const float* v, const float* mass,
                       const float* kT_per_replica_host) {
    CHX_REQUIRE(md && x && v && mass, "NULL argument");
    chx_ctx* ctx = md->ctx;
    const MdGeom& g = md->g;
    cudaStream_t st = ctx->stream;
    for (int r = 0; r < md->R; ++r) {
        md->rep_host[r] = MdRep();
        md->rep_host[r].kT = kT_per_replica_host ? kT_per_replica_host[r] : md->p.kT;
        md->rep_host[r].user_step = -1;
    }
    CHX_CUDA(cudaMemcpyAsync(md->rep, md->rep_host, md->R * sizeof(MdRep), cudaMemcpyHostToDevice, st));
    CHX_CUDA(cudaMemsetAsync(md->ctrl, 0, sizeof(MdCtrl), st));
    const dim3 gp(chx_div_up(g.np, 256), md->R);
    k_md_import<<<gp, 256, 0, st>>>(x, v, mass, g, md->xs[md->cur], md->vs[md->cur], md->refu[md->cur]);
    CHX_LAUNCHED(ctx);
    int rc = md_rebuild(md);
    if (rc != CHX_OK) return rc;
    rc = md_force(md, -2, false, 0, nullptr);
    if (rc != CHX_OK) return rc;
    md->have_state = true;
    return CHX_OK;
}

int chx_ljmd_get_state(chx_ljmd* md, float* x, float* v, float* force, float* ref_x) {
    CHX_REQUIRE(md && md->have_state, "engine has no state");
    const MdGeom& g = md->g;
    const dim3 gp(chx_div_up(g.np, 256), md->R);
    cudaStream_t st = md->ctx->stream;
    const int c = md->cur;
    if (x) { k_md_export_ids<<<gp, 256, 0, st>>>(md->xs[c], md->xs[c], g, x); CHX_LAUNCHED(md->ctx); }
    if (v) { k_md_export_ids<<<gp, 256, 0, st>>>(md->vs[c], md->xs[c], g, v); CHX_LAUNCHED(md->ctx); }
    if (force) { k_md_export_ids<<<gp, 256, 0, st>>>(md->fs, md->xs[c], g, force); CHX_LAUNCHED(md->ctx); }
    if (ref_x) { k_md_export_ids<<<gp, 256, 0, st>>>(md->refu[c], md->xs[c], g, ref_x); CHX_LAUNCHED(md->ctx); }
    return CHX_OK;
}

int chx_ljmd_run(chx_ljmd* md, int nsteps, uint32_t* keys_host, int report_interval,
                 double* energies_dev, int n_reports_capacity) {
    CHX_REQUIRE(md && md->have_state && keys_host, "engine has no state or keys are NULL");
    CHX_REQUIRE(nsteps >= 0, "nsteps must be >= 0");
    if (nsteps == 0) return CHX_OK;
    chx_ctx* ctx = md->ctx;
    const MdGeom& g = md->g;
    cudaStream_t st = ctx->stream;
    const int R = md->R;
    const bool report = energies_dev != nullptr && report_interval > 0;
    const int n_reports = report ? (nsteps + report_interval - 1) / report_interval : 0;
    CHX_REQUIRE(!report || n_reports <= n_reports_capacity, "energy buffer too small");
    // upload loop keys (parity 0), reset per-run control
    CHX_CUDA(cudaMemcpyAsync(md->rep_host, md->rep, R * sizeof(MdRep), cudaMemcpyDeviceToHost, st));
    CHX_CUDA(cudaStreamSynchronize(st));
    for (int r = 0; r < R; ++r) {
        md->rep_host[r].key[0][0] = keys_host[2 * r];
        md->rep_host[r].key[0][1] = keys_host[2 * r + 1];
        md->rep_host[r].user_step = -1;
    }
    CHX_CUDA(cudaMemcpyAsync(md->rep, md->rep_host, R * sizeof(MdRep), cudaMemcpyHostToDevice, st));
    CHX_CUDA(cudaMemsetAsync(&md->ctrl->halt_step, 0x7f, sizeof(int), st));
    if (report) CHX_CUDA(cudaMemsetAsync(energies_dev, 0, sizeof(double) * R * n_reports, st));

    const float dt = md->p.dt, gamma = md->p.gamma;
    const float h = dt * 0.5f;
    const float a = (float)exp((double)(float)(-gamma * dt));
    const float e2 = (float)exp((double)(float)(-2.0f * gamma * dt));
    const float bcoef = sqrtf(1.0f - e2);
    const float hs_user = (float)((double)md->p.skin / 2.0);
    const float hs_int = 0.5f * md->internal_skin;
    const float hs_int2 = hs_int * hs_int;
    const dim3 gb(chx_div_up(g.np, 256), R);
    const int CH = 32;

    auto force_at = [&](int s, int honor) -> int {
        const bool en = report && (s % report_interval == 0);
        // energies are stored [report][replica]; the kernel indexes energy_out[r]
        double* e = en ? energies_dev + (size_t)(s / report_interval) * R : nullptr;
        return md_force(md, s, en, honor, e);
    };

    int t = 0;
    while (t < nsteps) {
        const int chunk = nsteps - t < CH ? nsteps - t : CH;
        for (int s = t; s < t + chunk; ++s) {
            k_md_baoab<<<gb, 256, 0, st>>>(md->xs[md->cur], md->vs[md->cur], md->fs, md->refi,
                                           md->refu[md->cur], g, h, a, bcoef, hs_user, hs_int2, s,
                                           s > 0 ? 1 : 0, md->ctrl, md->rep);
            CHX_LAUNCHED(ctx);
            int rc = force_at(s, 1);
            if (rc != CHX_OK) return rc;
        }
        CHX_CUDA(cudaMemcpyAsync(md->ctrl_host, md->ctrl, sizeof(MdCtrl), cudaMemcpyDeviceToHost, st));
        CHX_CUDA(cudaStreamSynchronize(st));
        const int hs = md->ctrl_host->halt_step;
        if (hs == HALT_NONE) { t += chunk; continue; }
        int rc = md_rebuild(md);
        if (rc != CHX_OK) return rc;
        CHX_CUDA(cudaMemsetAsync(&md->ctrl->halt_step, 0x7f, sizeof(int), st));
        rc = force_at(hs, 0);
        if (rc != CHX_OK) return rc;
        t = hs + 1;
    }
    // trailing B of the last step (integrators.py:195)
    k_md_kick<<<chx_div_up((long long)R * g.np, 256), 256, 0, st>>>(md->vs[md->cur], md->fs, g, h, R);
    CHX_LAUNCHED(ctx);
    CHX_CUDA(cudaMemcpyAsync(md->rep_host, md->rep, R * sizeof(MdRep), cudaMemcpyDeviceToHost, st));
    CHX_CUDA(cudaStreamSynchronize(st));
    for (int r = 0; r < R; ++r) {
        keys_host[2 * r] = md->rep_host[r].key[nsteps & 1][0];
        keys_host[2 * r + 1] = md->rep_host[r].key[nsteps & 1][1];
    }
    md->steps += nsteps;
    return CHX_OK;
}

int chx_ljmd_energy(chx_ljmd* md, double* energy_dev) {
    CHX_REQUIRE(md && md->have_state && energy_dev, "engine has no state or energy_dev is NULL");
    cudaStream_t st = md->ctx->stream;
    CHX_CUDA(cudaMemsetAsync(energy_dev, 0, sizeof(double) * md->R, st));
    CHX_CUDA(cudaMemsetAsync(&md->ctrl->int_pairs2, 0, sizeof(unsigned long long), st));
    return md_force(md, -2, true, 0, energy_dev);
}

int chx_ljmd_stats(chx_ljmd* md, long long* stats_host) {
    CHX_REQUIRE(md && stats_host, "NULL argument");
    cudaStream_t st = md->ctx->stream;
    CHX_CUDA(cudaMemcpyAsync(md->ctrl_host, md->ctrl, sizeof(MdCtrl), cudaMemcpyDeviceToHost, st));
    CHX_CUDA(cudaMemcpyAsync(md->rep_host, md->rep, md->R * sizeof(MdRep), cudaMemcpyDeviceToHost, st));
    CHX_CUDA(cudaStreamSynchronize(st));
    long long user = 0;
    for (int r = 0; r < md->R; ++r) user += md->rep_host[r].user_rebuilds;
    stats_host[0] = md->rebuilds;
    stats_host[1] = (long long)(md->ctrl_host->cand_pairs2 / 2);
    stats_host[2] = (long long)(md->ctrl_host->int_pairs2 / 2);
    stats_host[3] = md->steps;
    stats_host[4] = md->ctx->launches - md->launches0;
    stats_host[5] = user;
    stats_host[6] = md->tcap;
    stats_host[7] = md->g.nblk;
    return CHX_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// measurement hooks
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fma_peak(int iters, float* __restrict__ sink) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
    float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float b = 0.999f, c = 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
            a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
        }
    }
    const float s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 12345.678f) sink[0] = s;
}

extern "C" {

int chx_ljmd_force_only(chx_ljmd* md, int repeats) {
    CHX_REQUIRE(md && md->have_state, "engine has no state");
    for (int k = 0; k < repeats; ++k) {
        int rc = md_force(md, -2, false, 0, nullptr);
        if (rc != CHX_OK) return rc;
    }
    return CHX_OK;
}

int chx_fma_peak(chx_ctx* ctx, int iters, double* flops_host) {
    CHX_REQUIRE(ctx && iters > 0, "bad argument");
    float* sink = (float*)chx_scratch(ctx, 256);
    if (!sink) return CHX_CUDA_ERROR;
    const int blocks = ctx->sm_count * 8;
    k_fma_peak<<<blocks, 256, 0, ctx->stream>>>(iters, sink);
    CHX_LAUNCHED(ctx);
    if (flops_host) *flops_host = 2.0 * 64.0 * (double)iters * 256.0 * (double)blocks;
    return CHX_OK;
}

}  // extern "C"
