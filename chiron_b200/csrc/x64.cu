// x64 variants of the parity-path kernels (the reference with `jax_enable_x64`: positions, displacements, energies
// and forces in float64).  BASELINE.json north_star: lists bit-exact as pair sets, energies / forces within rel 1e-10.
//   Space.displacement / wrap          chiron/neighbors.py:45-112, 116-175
//   NeighborListNsqrd build/calc/check chiron/neighbors.py:595-626, 671-729, 773-787, 864-907
//   LJPotential energy / force         chiron/potential.py:193-300 (force = -grad, potential.py:322-326)
//   BAOAB update                       chiron/integrators.py:174-195 (noise handed in: the float64 jax.random
//                                      stream draws 64-bit words, only the fp32 stream is generated in-kernel)
// One IEEE rounding per reference operation: __dadd_rn / __dmul_rn / __ddiv_rn / __dsqrt_rn are never contracted.
// These kernels are the API-parity path in double precision (FP64 throughput of the part is 1/64 of FP32); the
// throughput engine stays fp32 like the reference's default.
#include <math.h>
#include "common.cuh"
#include "cell_list.cuh"

struct BoxD { double lx, ly, lz, hx, hy, hz; };
static inline BoxD make_box_d(double lx, double ly, double lz) {
    BoxD b; b.lx = lx; b.ly = ly; b.lz = lz; b.hx = lx * 0.5; b.hy = ly * 0.5; b.hz = lz * 0.5; return b;
}

// jnp.mod(t, L) for L > 0: fmod (exact) + one rounded add when the remainder is negative
__device__ __forceinline__ double ref_mod_d(double t, double L) {
    double rem = (t >= 0.0 && t < L) ? t : fmod(t, L);
    if (rem != 0.0 && rem < 0.0) rem = __dadd_rn(rem, L);
    return rem;
}
__device__ __forceinline__ double ref_minimg_d(double a, double b, double L, double h) {
    return __dsub_rn(ref_mod_d(__dadd_rn(__dsub_rn(a, b), h), L), h);
}
template <bool PERIODIC>
__device__ __forceinline__ void ref_displacement_d(const double* a, const double* b, const BoxD& box, double& rx,
                                                   double& ry, double& rz, double& d) {
    if (PERIODIC) {
        rx = ref_minimg_d(a[0], b[0], box.lx, box.hx);
        ry = ref_minimg_d(a[1], b[1], box.ly, box.hy);
        rz = ref_minimg_d(a[2], b[2], box.lz, box.hz);
    } else {
        rx = __dsub_rn(a[0], b[0]); ry = __dsub_rn(a[1], b[1]); rz = __dsub_rn(a[2], b[2]);
    }
    d = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry)), __dmul_rn(rz, rz)));
}

template <bool PERIODIC>
__global__ void k_displacement_d(const double* __restrict__ x1, const double* __restrict__ x2, long long n, BoxD box,
                                 double* __restrict__ r, double* __restrict__ dist) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double rx, ry, rz, d;
    ref_displacement_d<PERIODIC>(x1 + 3 * i, x2 + 3 * i, box, rx, ry, rz, d);
    r[3 * i] = rx; r[3 * i + 1] = ry; r[3 * i + 2] = rz;
    dist[i] = d;
}

__global__ void k_wrap_d(const double* __restrict__ x, long long n, BoxD box, double* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double L[3] = {box.lx, box.ly, box.lz};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double v = x[3 * i + c];
        out[3 * i + c] = __dsub_rn(v, __dmul_rn(floor(__ddiv_rn(v, L[c])), L[c]));   // x - floor(x / L) * L
    }
}

// O(N^2) builder, one warp per row, j ascending (ballot compaction keeps the order); finish_row pads like the fp32 path
template <bool PERIODIC>
__global__ void __launch_bounds__(256)
k_build_nsq_d(const double* __restrict__ x, int n, BoxD box, double c, int M, uint32_t* __restrict__ list,
              int32_t* __restrict__ mask, int32_t* __restrict__ nn) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    int count = 0;
    uint32_t first = 0u;
    for (int j0 = (i + 1) & ~31; j0 < n; j0 += 32) {
        const int j = j0 + lane;
        bool hit = false;
        if (j > i && j < n) {
            double rx, ry, rz, d;
            ref_displacement_d<PERIODIC>(x + 3 * (size_t)i, x + 3 * (size_t)j, box, rx, ry, rz, d);
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

// calculate(): r_ik = minimg(x_i - x_list[i,k]), mask = (d < cutoff) & pad, n = sum(mask); optional LJ energy and
// forces over the masked entries in the same pass
template <bool PERIODIC>
__global__ void __launch_bounds__(256)
k_nlist_calculate_d(const double* __restrict__ x, int n, BoxD box, double cutoff, int M,
                    const uint32_t* __restrict__ list, const int32_t* __restrict__ pad, int32_t* __restrict__ n_out,
                    int32_t* __restrict__ mask_out, double* __restrict__ dist, double* __restrict__ rij, double sigma,
                    double epsilon, double* __restrict__ energy, double* __restrict__ force) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    int cnt = 0;
    double e_row = 0.0, fx = 0.0, fy = 0.0, fz = 0.0;
    for (int k = lane; k < M; k += 32) {
        const size_t o = (size_t)i * M + k;
        uint32_t j = list[o];
        j = j < (uint32_t)n ? j : (uint32_t)(n - 1);
        double rx, ry, rz, d;
        ref_displacement_d<PERIODIC>(x + 3 * (size_t)i, x + 3 * (size_t)j, box, rx, ry, rz, d);
        const int m = (d < cutoff) && (pad[o] != 0);
        if (mask_out) mask_out[o] = m;
        if (dist) dist[o] = d;
        if (rij) { rij[3 * o] = rx; rij[3 * o + 1] = ry; rij[3 * o + 2] = rz; }
        cnt += m;
        if (m && (energy || force)) {
            // potential.py:208-212: q = sigma / d, integer powers by repeated multiplication
            const double q = __ddiv_rn(sigma, d);
            const double q2 = __dmul_rn(q, q);
            const double q6 = __dmul_rn(__dmul_rn(q2, q2), q2);
            const double q12 = __dmul_rn(q6, q6);
            e_row += __dmul_rn(__dmul_rn(4.0, epsilon), __dsub_rn(q12, q6));
            if (force) {
                const double f = 24.0 * (epsilon / (d * d)) * (2.0 * q12 - q6);
                fx += f * rx; fy += f * ry; fz += f * rz;
                atomicAdd(&force[3 * (size_t)j], -f * rx);
                atomicAdd(&force[3 * (size_t)j + 1], -f * ry);
                atomicAdd(&force[3 * (size_t)j + 2], -f * rz);
            }
        }
    }
    cnt = warp_sum(cnt);
    if (lane == 0 && n_out) n_out[i] = cnt;
    if (energy) {
        e_row = warp_sum(e_row);
        if (lane == 0 && e_row != 0.0) atomicAdd(energy, e_row);
    }
    if (force) {
        fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
        if (lane == 0) {
            atomicAdd(&force[3 * (size_t)i], fx);
            atomicAdd(&force[3 * (size_t)i + 1], fy);
            atomicAdd(&force[3 * (size_t)i + 2], fz);
        }
    }
}

template <bool PERIODIC>
__global__ void k_nlist_check_d(const double* __restrict__ x, const double* __restrict__ ref, int n, BoxD box,
                                double half_skin, int32_t* __restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool moved = false;
    if (i < n) {
        double rx, ry, rz, d;
        ref_displacement_d<PERIODIC>(x + 3 * (size_t)i, ref + 3 * (size_t)i, box, rx, ry, rz, d);
        moved = d >= half_skin;
    }
    if (__syncthreads_or(moved) && threadIdx.x == 0) atomicOr(flag, 1);
}

// BAOAB sub-steps of one Langevin step with the noise handed in (integrators.py:181-189 + the wrap of :239):
//   v += (dt/2 F)/m ; x += dt/2 v ; v = a v + (b sqrt(kT/m)) xi ; x += dt/2 v ; wrap
__global__ void k_baoab_d(double* __restrict__ x, double* __restrict__ v, const double* __restrict__ F,
                          const double* __restrict__ mass, const double* __restrict__ xi, int n, double h, double a,
                          double b, double kT, BoxD box, int wrap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double m = mass[i];
    const double bs = __dmul_rn(b, __dsqrt_rn(__ddiv_rn(kT, m)));
    const double L[3] = {box.lx, box.ly, box.lz};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const size_t o = 3 * (size_t)i + c;
        double vv = __dadd_rn(v[o], __ddiv_rn(__dmul_rn(h, F[o]), m));
        double xx = __dadd_rn(x[o], __dmul_rn(h, vv));
        vv = __dadd_rn(__dmul_rn(a, vv), __dmul_rn(bs, xi[o]));
        xx = __dadd_rn(xx, __dmul_rn(h, vv));
        if (wrap) xx = __dsub_rn(xx, __dmul_rn(floor(__ddiv_rn(xx, L[c])), L[c]));
        x[o] = xx; v[o] = vv;
    }
}

__global__ void k_kick_d(double* __restrict__ v, const double* __restrict__ F, const double* __restrict__ mass, int n,
                         double h) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const size_t o = 3 * (size_t)i + c;
        v[o] = __dadd_rn(v[o], __ddiv_rn(__dmul_rn(h, F[o]), mass[i]));
    }
}

extern "C" {

int chx_displacement_f64(chx_ctx* ctx, const double* x1, const double* x2, long long n, double lx, double ly, double lz,
                         int periodic, double* r_out, double* dist_out) {
    CHX_REQUIRE(ctx && x1 && x2 && r_out && dist_out, "NULL argument");
    if (n <= 0) return CHX_OK;
    CHX_REQUIRE(!periodic || (lx > 0 && ly > 0 && lz > 0), "periodic displacement needs a positive box");
    const BoxD box = make_box_d(lx, ly, lz);
    const int g = chx_div_up(n, 256);
    if (periodic) k_displacement_d<true><<<g, 256, 0, ctx->stream>>>(x1, x2, n, box, r_out, dist_out);
    else k_displacement_d<false><<<g, 256, 0, ctx->stream>>>(x1, x2, n, box, r_out, dist_out);
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

int chx_wrap_f64(chx_ctx* ctx, const double* x, long long n, double lx, double ly, double lz, double* out) {
    CHX_REQUIRE(ctx && x && out, "NULL argument");
    CHX_REQUIRE(lx > 0 && ly > 0 && lz > 0, "wrap needs a positive box");
    if (n <= 0) return CHX_OK;
    k_wrap_d<<<chx_div_up(n, 256), 256, 0, ctx->stream>>>(x, n, make_box_d(lx, ly, lz), out);
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

int chx_nlist_build_nsq_f64(chx_ctx* ctx, const double* x, int n, double lx, double ly, double lz, int periodic,
                            double cutoff_plus_skin, int M, uint32_t* neighbor_list, int32_t* neighbor_mask,
                            int32_t* n_neighbors, int* max_count_host, int* count_eq_M_host) {
    CHX_REQUIRE(ctx && x && neighbor_list && neighbor_mask && n_neighbors, "NULL argument");
    CHX_REQUIRE(n > 0 && M > 0, "n and M must be positive");
    const BoxD box = make_box_d(lx, ly, lz);
    if (periodic) k_build_nsq_d<true><<<chx_div_up(n, 8), 256, 0, ctx->stream>>>(x, n, box, cutoff_plus_skin, M, neighbor_list, neighbor_mask, n_neighbors);
    else k_build_nsq_d<false><<<chx_div_up(n, 8), 256, 0, ctx->stream>>>(x, n, box, cutoff_plus_skin, M, neighbor_list, neighbor_mask, n_neighbors);
    CHX_LAUNCHED(ctx);
    int* stats = (int*)chx_scratch(ctx, 256);
    if (!stats) return CHX_CUDA_ERROR;
    CHX_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(int), ctx->stream));
    k_count_stats<<<chx_div_up(n, 256), 256, 0, ctx->stream>>>(n_neighbors, n, M, stats);
    CHX_LAUNCHED(ctx);
    CHX_CUDA(cudaMemcpyAsync(ctx->host_pinned, stats, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CHX_CUDA(cudaStreamSynchronize(ctx->stream));
    if (max_count_host) *max_count_host = ctx->host_pinned[0];
    if (count_eq_M_host) *count_eq_M_host = ctx->host_pinned[1];
    return CHX_OK;
}

int chx_nlist_calculate_f64(chx_ctx* ctx, const double* x, int n, double lx, double ly, double lz, int periodic,
                            double cutoff, int M, const uint32_t* neighbor_list, const int32_t* neighbor_mask,
                            int32_t* n_out, int32_t* mask_out, double* dist_out, double* rij_out) {
    CHX_REQUIRE(ctx && x && neighbor_list && neighbor_mask, "NULL argument");
    CHX_REQUIRE(n > 0 && M > 0, "n and M must be positive");
    const BoxD box = make_box_d(lx, ly, lz);
    if (periodic) k_nlist_calculate_d<true><<<chx_div_up(n, 8), 256, 0, ctx->stream>>>(x, n, box, cutoff, M, neighbor_list, neighbor_mask, n_out, mask_out, dist_out, rij_out, 0.0, 0.0, nullptr, nullptr);
    else k_nlist_calculate_d<false><<<chx_div_up(n, 8), 256, 0, ctx->stream>>>(x, n, box, cutoff, M, neighbor_list, neighbor_mask, n_out, mask_out, dist_out, rij_out, 0.0, 0.0, nullptr, nullptr);
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

int chx_nlist_check_f64(chx_ctx* ctx, const double* x, const double* ref_x, int n, double lx, double ly, double lz,
                        int periodic, double half_skin, int32_t* flag_dev) {
    CHX_REQUIRE(ctx && x && ref_x && flag_dev, "NULL argument");
    CHX_CUDA(cudaMemsetAsync(flag_dev, 0, sizeof(int32_t), ctx->stream));
    if (n <= 0) return CHX_OK;
    const BoxD box = make_box_d(lx, ly, lz);
    if (periodic) k_nlist_check_d<true><<<chx_div_up(n, 256), 256, 0, ctx->stream>>>(x, ref_x, n, box, half_skin, flag_dev);
    else k_nlist_check_d<false><<<chx_div_up(n, 256), 256, 0, ctx->stream>>>(x, ref_x, n, box, half_skin, flag_dev);
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

int chx_lj_nlist_energy_force_f64(chx_ctx* ctx, const double* x, int n, double lx, double ly, double lz, int periodic,
                                  double sigma, double epsilon, double cutoff, int M, const uint32_t* neighbor_list,
                                  const int32_t* neighbor_mask, double* energy_dev, double* force_dev) {
    CHX_REQUIRE(ctx && x && neighbor_list && neighbor_mask, "NULL argument");
    CHX_REQUIRE(energy_dev || force_dev, "energy_dev and force_dev are both NULL");
    CHX_REQUIRE(n > 0 && M > 0, "n and M must be positive");
    if (energy_dev) CHX_CUDA(cudaMemsetAsync(energy_dev, 0, sizeof(double), ctx->stream));
    if (force_dev) CHX_CUDA(cudaMemsetAsync(force_dev, 0, sizeof(double) * 3 * (size_t)n, ctx->stream));
    const BoxD box = make_box_d(lx, ly, lz);
    if (periodic) k_nlist_calculate_d<true><<<chx_div_up(n, 8), 256, 0, ctx->stream>>>(x, n, box, cutoff, M, neighbor_list, neighbor_mask, nullptr, nullptr, nullptr, nullptr, sigma, epsilon, energy_dev, force_dev);
    else k_nlist_calculate_d<false><<<chx_div_up(n, 8), 256, 0, ctx->stream>>>(x, n, box, cutoff, M, neighbor_list, neighbor_mask, nullptr, nullptr, nullptr, nullptr, sigma, epsilon, energy_dev, force_dev);
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

int chx_baoab_update_f64(chx_ctx* ctx, double* x, double* v, const double* F, const double* mass, const double* noise,
                         int n, double half_dt, double a, double b, double kT, double lx, double ly, double lz, int wrap) {
    CHX_REQUIRE(ctx && x && v && F && mass && noise, "NULL argument");
    if (n <= 0) return CHX_OK;
    k_baoab_d<<<chx_div_up(n, 256), 256, 0, ctx->stream>>>(x, v, F, mass, noise, n, half_dt, a, b, kT,
                                                          make_box_d(lx, ly, lz), wrap);
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

int chx_kick_f64(chx_ctx* ctx, double* v, const double* F, const double* mass, int n, double half_dt) {
    CHX_REQUIRE(ctx && v && F && mass, "NULL argument");
    if (n <= 0) return CHX_OK;
    k_kick_d<<<chx_div_up(n, 256), 256, 0, ctx->stream>>>(v, F, mass, n, half_dt);
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

}  // extern "C"
