// Potentials over reference-shaped inputs (API-parity path):
//   LJ over a NeighborListNsqrd         chiron/potential.py:193-213, 263-300
//   LJ over all pairs (pair list / none) chiron/potential.py:26-63, 235-258; neighbors.py:1106-1216
//   harmonic oscillator                 chiron/potential.py:413-418
//   subset delta energy                 new fast path for mcmc.py:733-777 with atom_subset
// The fused engine (engine.cu) is the throughput path; these kernels mirror the reference's data
// layout one-to-one and use the exact fp32 predicate everywhere.
#include "common.cuh"

// pair energy / force scalar from the exactly-rounded distance d (d > 0)
__device__ __forceinline__ void lj_pair(float d, float sigma, float eps, float& e, float& f) {
    const float q = sigma / d;
    const float q2 = q * q;
    const float q6 = q2 * q2 * q2;
    const float q12 = q6 * q6;
    e = 4.0f * eps * (q12 - q6);
    f = 24.0f * (eps / (d * d)) * (2.0f * q12 - q6);
}

__device__ __forceinline__ void block_add_double(double v, double* target) {
    __shared__ double sh[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        double t = lane < nw ? sh[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0 && t != 0.0) atomicAdd(target, t);
    }
}

// one warp per row of the half list; f_i reduced in the warp, f_j scattered with atomics.  Pairs are
// classified with the fast test (exact predicate only within a few ulps of the cutoff); energies and
// forces come from r2 without sqrt / division (relative error ~1e-7).
template <bool PERIODIC, bool WANT_E, bool WANT_F>
__global__ void __launch_bounds__(256, WANT_F ? 1 : 8)
k_lj_nlist(const float* __restrict__ x, int n, Box box, FastCut fc, const uint32_t* __restrict__ list,
           const int32_t* __restrict__ nn, int M, float sigma, float eps,
           double* __restrict__ energy, float* __restrict__ force) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    double e_acc = 0.0;
    if (i < n) {
        const float xi = x[3 * i], yi = x[3 * i + 1], zi = x[3 * i + 2];
        const float sigma2 = sigma * sigma;
        int cnt = nn[i];
        cnt = cnt < M ? cnt : M;
        float fx = 0.f, fy = 0.f, fz = 0.f, e_row = 0.f;
        if (!WANT_F) {
            e_row = lj_nlist_row_energy<PERIODIC>(Pos3{x}, i, lane, box, fc, list + (size_t)i * M, cnt, sigma2, eps);
        } else
        for (int k = lane; k < cnt; k += 32) {
            const uint32_t j = list[(size_t)i * M + k];
            float r2, dx, dy, dz;
            if (fast_within<PERIODIC>(xi, yi, zi, x[3 * j], x[3 * j + 1], x[3 * j + 2], box, fc, r2, dx, dy, dz)) {
                float e, f;
                lj_pair_r2(r2, sigma2, eps, e, f);
                if (WANT_E) e_row += e;
                if (WANT_F) {
                    const float px = f * dx, py = f * dy, pz = f * dz;
                    fx += px; fy += py; fz += pz;
                    atomicAdd(&force[3 * j], -px);
                    atomicAdd(&force[3 * j + 1], -py);
                    atomicAdd(&force[3 * j + 2], -pz);
                }
            }
        }
        e_acc = (double)e_row;      // at most M / 32 terms per lane in fp32, rows summed in fp64
        if (WANT_F) {
            fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
            if (lane == 0) {
                atomicAdd(&force[3 * i], fx);
                atomicAdd(&force[3 * i + 1], fy);
                atomicAdd(&force[3 * i + 2], fz);
            }
        }
    }
    if (WANT_E) block_add_double(e_acc, energy);
}

// all pairs: blockIdx.x tiles i (128 per block), blockIdx.y splits the j range; j tiles staged in
// shared memory; every ordered pair is visited (energy halved), so no j-side scatter is needed.
#define AP_TI 128
template <bool PERIODIC, bool WANT_E, bool WANT_F>
__global__ void __launch_bounds__(AP_TI)
k_lj_allpairs(const float* __restrict__ x, int n, Box box, float sigma, float eps, float cutoff,
              bool use_cutoff, int j_chunk, double* __restrict__ energy, float* __restrict__ force) {
    __shared__ float sx[AP_TI], sy[AP_TI], sz[AP_TI];
    const int i = blockIdx.x * AP_TI + threadIdx.x;
    const int j_lo = blockIdx.y * j_chunk;
    const int j_hi = min(n, j_lo + j_chunk);
    float xi = 0.f, yi = 0.f, zi = 0.f;
    if (i < n) { xi = x[3 * i]; yi = x[3 * i + 1]; zi = x[3 * i + 2]; }
    float fx = 0.f, fy = 0.f, fz = 0.f;
    double e_acc = 0.0;
    for (int j0 = j_lo; j0 < j_hi; j0 += AP_TI) {
        const int jl = j0 + threadIdx.x;
        if (jl < j_hi) { sx[threadIdx.x] = x[3 * jl]; sy[threadIdx.x] = x[3 * jl + 1]; sz[threadIdx.x] = x[3 * jl + 2]; }
        __syncthreads();
        const int cnt = min(AP_TI, j_hi - j0);
        if (i < n) {
            for (int t = 0; t < cnt; ++t) {
                const int j = j0 + t;
                if (j == i) continue;
                float rx, ry, rz, d;
                // the reference reduces over i < j only (reduction_mask, neighbors.py:1094-1104): the
                // pair is evaluated in that orientation from both sides (fp32 minimum images are not
                // exactly antisymmetric)
                if (i < j) {
                    ref_displacement<PERIODIC>(xi, yi, zi, sx[t], sy[t], sz[t], box, rx, ry, rz, d);
                } else {
                    ref_displacement<PERIODIC>(sx[t], sy[t], sz[t], xi, yi, zi, box, rx, ry, rz, d);
                    rx = -rx; ry = -ry; rz = -rz;
                }
                if (!use_cutoff || d < cutoff) {
                    float e, f;
                    lj_pair(d, sigma, eps, e, f);
                    if (WANT_E) e_acc += (double)e;
                    if (WANT_F) { fx += f * rx; fy += f * ry; fz += f * rz; }
                }
            }
        }
        __syncthreads();
    }
    if (WANT_F && i < n) {
        atomicAdd(&force[3 * i], fx);
        atomicAdd(&force[3 * i + 1], fy);
        atomicAdd(&force[3 * i + 2], fz);
    }
    if (WANT_E) block_add_double(0.5 * e_acc, energy);
}

template <bool WANT_F>
__global__ void k_ho(const float* __restrict__ x, int n, const float* __restrict__ x0, int n0,
                     float k, double* __restrict__ energy, float* __restrict__ force) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0.0;
    if (i < n) {
        const int r = n0 == 1 ? 0 : i;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float dx = __fsub_rn(x[3 * i + c], x0[3 * r + c]);
            e += (double)__fmul_rn(dx, dx);
            if (WANT_F) force[3 * i + c] = -k * dx;
        }
    }
    if (energy) block_add_double(e, energy);
}

__global__ void k_ho_finish(double* energy, float k, float U0) {
    // 0.5 * k * sum + U0 in fp32 like the reference
    const float s = (float)(*energy);
    *energy = (double)__fadd_rn(__fmul_rn(__fmul_rn(0.5f, k), s), U0);
}

// one block per moved particle: sum_j w_j [u(xn_m, xn_j) - u(xo_m, xo_j)], w_j = 1/2 if j moved too
template <bool PERIODIC>
__global__ void __launch_bounds__(256)
k_lj_subset_delta(const float* __restrict__ xo, const float* __restrict__ xn, int n,
                  const uint32_t* __restrict__ moved, const uint8_t* __restrict__ is_moved, Box box,
                  float sigma, float eps, float cutoff, double* __restrict__ delta) {
    const int m = (int)moved[blockIdx.x];
    const float ox = xo[3 * m], oy = xo[3 * m + 1], oz = xo[3 * m + 2];
    const float nx = xn[3 * m], ny = xn[3 * m + 1], nz = xn[3 * m + 2];
    double acc = 0.0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        if (j == m) continue;
        const double w = is_moved[j] ? 0.5 : 1.0;
        float rx, ry, rz, d, e, f;
        // orientation of the reference's half list: displacement(x_lower id, x_higher id)
        if (m < j) ref_displacement<PERIODIC>(nx, ny, nz, xn[3 * j], xn[3 * j + 1], xn[3 * j + 2], box, rx, ry, rz, d);
        else ref_displacement<PERIODIC>(xn[3 * j], xn[3 * j + 1], xn[3 * j + 2], nx, ny, nz, box, rx, ry, rz, d);
        if (d < cutoff) { lj_pair(d, sigma, eps, e, f); acc += w * (double)e; }
        if (m < j) ref_displacement<PERIODIC>(ox, oy, oz, xo[3 * j], xo[3 * j + 1], xo[3 * j + 2], box, rx, ry, rz, d);
        else ref_displacement<PERIODIC>(xo[3 * j], xo[3 * j + 1], xo[3 * j + 2], ox, oy, oz, box, rx, ry, rz, d);
        if (d < cutoff) { lj_pair(d, sigma, eps, e, f); acc -= w * (double)e; }
    }
    block_add_double(acc, delta);
}

__global__ void k_mark_moved(const uint32_t* __restrict__ moved, int n_moved, uint8_t* __restrict__ is_moved) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_moved) is_moved[moved[t]] = 1;
}

// Generalisation the reference's API hints at (potential.py:131-137 takes one sigma /
// epsilon): per-particle parameters with Lorentz-Berthelot mixing, sigma_ij = (sigma_i + sigma_j) / 2,
// eps_ij = sqrt(eps_i eps_j), and an optional energy shift that makes every pair energy zero at the cutoff.
// Same half list, same exact cutoff predicate, same pair formula (potential.py:208-212) per pair.
template <bool PERIODIC, bool WANT_F>
__global__ void __launch_bounds__(256)
k_lj_nlist_mixed(const float* __restrict__ x, int n, Box box, FastCut fc, const uint32_t* __restrict__ list,
                 const int32_t* __restrict__ nn, int M, const float* __restrict__ sigma_i,
                 const float* __restrict__ eps_i, int shift, float r_switch, double* __restrict__ energy,
                 float* __restrict__ force) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    double e_acc = 0.0;
    if (i < n) {
        const float xi = x[3 * i], yi = x[3 * i + 1], zi = x[3 * i + 2];
        const float si = sigma_i[i], ei = eps_i[i];
        int cnt = nn[i];
        cnt = cnt < M ? cnt : M;
        float fx = 0.f, fy = 0.f, fz = 0.f, e_row = 0.f;
        for (int k = lane; k < cnt; k += 32) {
            const uint32_t j = list[(size_t)i * M + k];
            float r2, dx, dy, dz;
            if (fast_within<PERIODIC>(xi, yi, zi, x[3 * j], x[3 * j + 1], x[3 * j + 2], box, fc, r2, dx, dy, dz)) {
                const float sij = 0.5f * (si + sigma_i[j]);
                const float eij = sqrtf(ei * eps_i[j]);
                float e, f;
                lj_pair_r2(r2, sij * sij, eij, e, f);
                if (shift) {
                    float ec, fc_unused;
                    lj_pair_r2(fc.c * fc.c, sij * sij, eij, ec, fc_unused);
                    e -= ec;
                }
                if (r_switch > 0.f && r2 > r_switch * r_switch) {
                    // switching function of OpenMM's NonbondedForce: S = 1 - 6 t^5 + 15 t^4 - 10 t^3,
                    // t = (r - r_switch) / (r_cut - r_switch); E -> E S, (F/r) -> (F/r) S - E S' / r
                    const float rr = sqrtf(r2);
                    const float w = 1.0f / (fc.c - r_switch);
                    const float t = (rr - r_switch) * w;
                    const float S = 1.0f + t * t * t * (-10.0f + t * (15.0f - 6.0f * t));
                    const float dS = -30.0f * t * t * (1.0f - t) * (1.0f - t) * w;
                    f = f * S - e * dS / rr;
                    e *= S;
                }
                e_row += e;
                if (WANT_F) {
                    const float px = f * dx, py = f * dy, pz = f * dz;
                    fx += px; fy += py; fz += pz;
                    atomicAdd(&force[3 * j], -px);
                    atomicAdd(&force[3 * j + 1], -py);
                    atomicAdd(&force[3 * j + 2], -pz);
                }
            }
        }
        e_acc = (double)e_row;
        if (WANT_F) {
            fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
            if (lane == 0) {
                atomicAdd(&force[3 * i], fx);
                atomicAdd(&force[3 * i + 1], fy);
                atomicAdd(&force[3 * i + 2], fz);
            }
        }
    }
    if (energy) block_add_double(e_acc, energy);
}

extern "C" {

int chx_lj_nlist_energy_force_mixed(chx_ctx* ctx, const float* x, int n, float lx, float ly, float lz, int periodic,
                                    const uint32_t* neighbor_list, const int32_t* n_neighbors, int M,
                                    const float* sigma_per_particle, const float* epsilon_per_particle, float cutoff,
                                    int shift, float switch_distance, double* energy_dev, float* force) {
    CHX_REQUIRE(ctx && x && neighbor_list && n_neighbors && sigma_per_particle && epsilon_per_particle, "NULL argument");
    CHX_REQUIRE(n > 0 && M > 0, "n and M must be positive");
    CHX_REQUIRE(energy_dev || force, "nothing to compute");
    CHX_REQUIRE(switch_distance >= 0.f && switch_distance < cutoff, "switch_distance must lie in [0, cutoff)");
    Box box = make_box(lx, ly, lz);
    if (energy_dev) CHX_CUDA(cudaMemsetAsync(energy_dev, 0, sizeof(double), ctx->stream));
    if (force) CHX_CUDA(cudaMemsetAsync(force, 0, sizeof(float) * 3 * (size_t)n, ctx->stream));
    const FastCut fc = make_fast_cut(cutoff, lx, ly, lz, periodic != 0);
    const int blocks = chx_div_up(n, 8);
#define LAUNCH(P, F)                                                                                       \
    k_lj_nlist_mixed<P, F><<<blocks, 256, 0, ctx->stream>>>(x, n, box, fc, neighbor_list, n_neighbors, M,    \
                                                            sigma_per_particle, epsilon_per_particle, shift, \
                                                            switch_distance, energy_dev, force)
    if (periodic) { if (force) LAUNCH(true, true); else LAUNCH(true, false); }
    else { if (force) LAUNCH(false, true); else LAUNCH(false, false); }
#undef LAUNCH
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

int chx_lj_nlist_energy_force(chx_ctx* ctx, const float* x, int n, float lx, float ly, float lz,
                              int periodic, const uint32_t* neighbor_list,
                              const int32_t* n_neighbors, int M, float sigma, float epsilon,
                              float cutoff, double* energy_dev, float* force) {
    CHX_REQUIRE(ctx && x && neighbor_list && n_neighbors, "NULL argument");
    CHX_REQUIRE(n > 0 && M > 0, "n and M must be positive");
    CHX_REQUIRE(energy_dev || force, "nothing to compute");
    Box box = make_box(lx, ly, lz);
    if (energy_dev) CHX_CUDA(cudaMemsetAsync(energy_dev, 0, sizeof(double), ctx->stream));
    if (force) CHX_CUDA(cudaMemsetAsync(force, 0, sizeof(float) * 3 * (size_t)n, ctx->stream));
    const int blocks = chx_div_up(n, 8);
    const FastCut fc = make_fast_cut(cutoff, lx, ly, lz, periodic != 0);
#define LAUNCH(P, E, F)                                                                              \
    k_lj_nlist<P, E, F><<<blocks, 256, 0, ctx->stream>>>(x, n, box, fc, neighbor_list, n_neighbors, M, \
                                                         sigma, epsilon, energy_dev, force)
    if (periodic) {
        if (energy_dev && force) LAUNCH(true, true, true);
        else if (energy_dev) LAUNCH(true, true, false);
        else LAUNCH(true, false, true);
    } else {
        if (energy_dev && force) LAUNCH(false, true, true);
        else if (energy_dev) LAUNCH(false, true, false);
        else LAUNCH(false, false, true);
    }
#undef LAUNCH
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

int chx_lj_allpairs_energy_force(chx_ctx* ctx, const float* x, int n, float lx, float ly, float lz,
                                 int periodic, float sigma, float epsilon, float cutoff,
                                 double* energy_dev, float* force) {
    CHX_REQUIRE(ctx && x, "NULL argument");
    CHX_REQUIRE(n > 0, "n must be positive");
    CHX_REQUIRE(energy_dev || force, "nothing to compute");
    Box box = make_box(lx, ly, lz);
    if (energy_dev) CHX_CUDA(cudaMemsetAsync(energy_dev, 0, sizeof(double), ctx->stream));
    if (force) CHX_CUDA(cudaMemsetAsync(force, 0, sizeof(float) * 3 * (size_t)n, ctx->stream));
    const int bi = chx_div_up(n, AP_TI);
    int split = (2 * ctx->sm_count + bi - 1) / bi;
    const int max_split = chx_div_up(n, AP_TI);
    split = split < 1 ? 1 : (split > max_split ? max_split : split);
    int j_chunk = chx_div_up(n, split);
    j_chunk = chx_div_up(j_chunk, AP_TI) * AP_TI;
    split = chx_div_up(n, j_chunk);
    dim3 grid(bi, split);
    const bool uc = cutoff >= 0.0f;
#define LAUNCH(P, E, F)                                                                        \
    k_lj_allpairs<P, E, F><<<grid, AP_TI, 0, ctx->stream>>>(x, n, box, sigma, epsilon, cutoff, uc, \
                                                            j_chunk, energy_dev, force)
    if (periodic) {
        if (energy_dev && force) LAUNCH(true, true, true);
        else if (energy_dev) LAUNCH(true, true, false);
        else LAUNCH(true, false, true);
    } else {
        if (energy_dev && force) LAUNCH(false, true, true);
        else if (energy_dev) LAUNCH(false, true, false);
        else LAUNCH(false, false, true);
    }
#undef LAUNCH
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

int chx_ho_energy_force(chx_ctx* ctx, const float* x, int n, const float* x0, int n0, float k,
                        float U0, double* energy_dev, float* force) {
    CHX_REQUIRE(ctx && x && x0, "NULL argument");
    CHX_REQUIRE(n > 0 && (n0 == 1 || n0 == n), "x0 must have 1 or n rows");
    CHX_REQUIRE(energy_dev || force, "nothing to compute");
    if (energy_dev) CHX_CUDA(cudaMemsetAsync(energy_dev, 0, sizeof(double), ctx->stream));
    if (force) k_ho<true><<<chx_div_up(n, 256), 256, 0, ctx->stream>>>(x, n, x0, n0, k, energy_dev, force);
    else k_ho<false><<<chx_div_up(n, 256), 256, 0, ctx->stream>>>(x, n, x0, n0, k, energy_dev, force);
    CHX_LAUNCHED(ctx);
    if (energy_dev) {
        k_ho_finish<<<1, 1, 0, ctx->stream>>>(energy_dev, k, U0);
        CHX_LAUNCHED(ctx);
    }
    return CHX_OK;
}

int chx_lj_subset_delta_energy(chx_ctx* ctx, const float* x_old, const float* x_new, int n,
                               const uint32_t* moved, int n_moved, float lx, float ly, float lz,
                               int periodic, float sigma, float epsilon, float cutoff,
                               double* delta_dev) {
    CHX_REQUIRE(ctx && x_old && x_new && moved && delta_dev, "NULL argument");
    CHX_REQUIRE(n > 0 && n_moved > 0, "n and n_moved must be positive");
    uint8_t* is_moved = (uint8_t*)chx_scratch(ctx, (size_t)n);
    if (!is_moved) return CHX_CUDA_ERROR;
    CHX_CUDA(cudaMemsetAsync(is_moved, 0, (size_t)n, ctx->stream));
    CHX_CUDA(cudaMemsetAsync(delta_dev, 0, sizeof(double), ctx->stream));
    k_mark_moved<<<chx_div_up(n_moved, 256), 256, 0, ctx->stream>>>(moved, n_moved, is_moved);
    CHX_LAUNCHED(ctx);
    Box box = make_box(lx, ly, lz);
    if (periodic)
        k_lj_subset_delta<true><<<n_moved, 256, 0, ctx->stream>>>(x_old, x_new, n, moved, is_moved, box, sigma, epsilon, cutoff, delta_dev);
    else
        k_lj_subset_delta<false><<<n_moved, 256, 0, ctx->stream>>>(x_old, x_new, n, moved, is_moved, box, sigma, epsilon, cutoff, delta_dev);
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

}  // extern "C"
