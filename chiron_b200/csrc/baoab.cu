// BAOAB building blocks for arbitrary potentials (chiron/integrators.py:174-195), Maxwell-Boltzmann
// initialisation (chiron/utils.py:116-144) and the Monte Carlo displacement proposal
// (chiron/mcmc.py:733-752).  All arithmetic follows the reference op by op (one fp32 rounding per
// jnp operation, no FMA contraction) so a step reproduces the JAX path to the last few ulps; the
// noise is generated in-kernel on the reference's legacy-threefry stream.
#include "common.cuh"

template <bool WRAP, bool CHECK>
__global__ void __launch_bounds__(256)
k_baoab(float* __restrict__ x, float* __restrict__ v, const float* __restrict__ force,
        const float* __restrict__ mass, int n, float h, float a, float b, float kT, uint32_t k0,
        uint32_t k1, int trailing_kick, Box box, const float* __restrict__ ref, float half_skin,
        int32_t* __restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool moved = false;
    if (i < n) {
        const float m = mass[i];
        const float sv = __fsqrt_rn(__fdiv_rn(kT, m));
        const float bs = __fmul_rn(b, sv);
        const unsigned long long total = 3ull * (unsigned long long)n;
        float xn[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float xc = x[3 * i + c], vc = v[3 * i + c];
            const float kick = __fdiv_rn(__fmul_rn(h, force[3 * i + c]), m);
            if (trailing_kick) vc = __fadd_rn(vc, kick);            // B of the previous step
            vc = __fadd_rn(vc, kick);                               // B
            xc = __fadd_rn(xc, __fmul_rn(h, vc));                   // A
            const float xi = normal_from_bits(random_bits_elem(k0, k1, 3ull * i + c, total));
            vc = __fadd_rn(__fmul_rn(a, vc), __fmul_rn(bs, xi));    // O
            xc = __fadd_rn(xc, __fmul_rn(h, vc));                   // A
            if (WRAP) xc = ref_wrap(xc, c == 0 ? box.lx : (c == 1 ? box.ly : box.lz));
            x[3 * i + c] = xc;
            v[3 * i + c] = vc;
            xn[c] = xc;
        }
        if (CHECK) {
            float rx, ry, rz, d;
            if (WRAP)
                ref_displacement<true>(xn[0], xn[1], xn[2], ref[3 * i], ref[3 * i + 1], ref[3 * i + 2], box, rx, ry, rz, d);
            else
                ref_displacement<false>(xn[0], xn[1], xn[2], ref[3 * i], ref[3 * i + 1], ref[3 * i + 2], box, rx, ry, rz, d);
            moved = d >= half_skin;
        }
    }
    if (CHECK) {
        if (__syncthreads_or(moved) && threadIdx.x == 0) atomicOr(flag, 1);
    }
}

__global__ void k_kick(float* __restrict__ v, const float* __restrict__ force,
                       const float* __restrict__ mass, int n, float h) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * n) return;
    v[t] = __fadd_rn(v[t], __fdiv_rn(__fmul_rn(h, force[t]), mass[t / 3]));
}

__global__ void k_init_velocities(float* __restrict__ v, const float* __restrict__ mass, int n,
                                  float kT, uint32_t k0, uint32_t k1) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * n) return;
    const float sv = __fsqrt_rn(__fdiv_rn(kT, mass[t / 3]));
    v[t] = __fmul_rn(sv, normal_from_bits(random_bits_elem(k0, k1, (unsigned long long)t, 3ull * n)));
}

template <bool WRAP>
__global__ void k_mc_displace(const float* __restrict__ x, int n, uint32_t k0, uint32_t k1,
                              float sigma, const float* __restrict__ subset, Box box,
                              float* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * n) return;
    const int c = t % 3;
    float d = __fmul_rn(normal_from_bits(random_bits_elem(k0, k1, (unsigned long long)t, 3ull * n)), sigma);
    if (subset) d = __fmul_rn(d, subset[t / 3]);
    float xc = __fadd_rn(x[t], d);
    if (WRAP) xc = ref_wrap(xc, c == 0 ? box.lx : (c == 1 ? box.ly : box.lz));
    out[t] = xc;
}

extern "C" {

int chx_baoab_update(chx_ctx* ctx, float* x, float* v, const float* force, const float* mass, int n,
                     float half_dt, float a, float b, float kT, uint32_t subkey0, uint32_t subkey1,
                     int trailing_kick, float lx, float ly, float lz, int wrap_periodic,
                     const float* ref_x, float half_skin, int32_t* flag_dev) {
    CHX_REQUIRE(ctx && x && v && force && mass, "NULL argument");
    CHX_REQUIRE(n > 0, "n must be positive");
    const bool check = ref_x != nullptr && flag_dev != nullptr;
    Box box = make_box(lx, ly, lz);
    const int blocks = chx_div_up(n, 256);
#define LAUNCH(W, C)                                                                               \
    k_baoab<W, C><<<blocks, 256, 0, ctx->stream>>>(x, v, force, mass, n, half_dt, a, b, kT, subkey0, \
                                                   subkey1, trailing_kick, box, ref_x, half_skin,  \
                                                   flag_dev)
    if (wrap_periodic) { if (check) LAUNCH(true, true); else LAUNCH(true, false); }
    else { if (check) LAUNCH(false, true); else LAUNCH(false, false); }
#undef LAUNCH
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

int chx_kick(chx_ctx* ctx, float* v, const float* force, const float* mass, int n, float half_dt) {
    CHX_REQUIRE(ctx && v && force && mass, "NULL argument");
    CHX_REQUIRE(n > 0, "n must be positive");
    k_kick<<<chx_div_up(3LL * n, 256), 256, 0, ctx->stream>>>(v, force, mass, n, half_dt);
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

int chx_init_velocities(chx_ctx* ctx, float* v, const float* mass, int n, float kT, uint32_t key0,
                        uint32_t key1) {
    CHX_REQUIRE(ctx && v && mass, "NULL argument");
    CHX_REQUIRE(n > 0, "n must be positive");
    k_init_velocities<<<chx_div_up(3LL * n, 256), 256, 0, ctx->stream>>>(v, mass, n, kT, key0, key1);
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

int chx_mc_displace(chx_ctx* ctx, const float* x, int n, uint32_t key0, uint32_t key1, float sigma,
                    const float* subset_mask, float lx, float ly, float lz, int wrap_periodic,
                    float* x_out) {
    CHX_REQUIRE(ctx && x && x_out, "NULL argument");
    CHX_REQUIRE(n > 0, "n must be positive");
    Box box = make_box(lx, ly, lz);
    if (wrap_periodic)
        k_mc_displace<true><<<chx_div_up(3LL * n, 256), 256, 0, ctx->stream>>>(x, n, key0, key1, sigma, subset_mask, box, x_out);
    else
        k_mc_displace<false><<<chx_div_up(3LL * n, 256), 256, 0, ctx->stream>>>(x, n, key0, key1, sigma, subset_mask, box, x_out);
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

}  // extern "C"
