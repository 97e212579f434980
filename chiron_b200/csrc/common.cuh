// Shared host/device helpers for libchiron_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "../../include/chiron_b200.h"

// ---------------------------------------------------------------------------------------------
// Context and error plumbing
// ---------------------------------------------------------------------------------------------
struct chx_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    long long launches = 0;
    int sm_count = 148;
    // scratch owned by the context (grown on demand, never shrunk)
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    int* host_pinned = nullptr;  // small pinned staging area for host scalars (64 ints)
};

void chx_set_error(const char* fmt, ...);
void chx_mc_forget_context(chx_ctx* ctx);   // mc.cu: drop the Metropolis-loop graphs cached for this context
void* chx_scratch(chx_ctx* ctx, size_t bytes);  // device scratch of at least `bytes`, 256B aligned

#define CHX_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t err__ = (call);                                                         \
        if (err__ != cudaSuccess) {                                                         \
            chx_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                     \
                          cudaGetErrorString(err__));                                       \
            return CHX_CUDA_ERROR;                                                          \
        }                                                                                   \
    } while (0)

#define CHX_REQUIRE(cond, msg)                                           \
    do {                                                                 \
        if (!(cond)) {                                                   \
            chx_set_error("%s:%d: %s", __FILE__, __LINE__, msg);         \
            return CHX_BAD_ARG;                                          \
        }                                                                \
    } while (0)

#define CHX_LAUNCHED(ctx)                 \
    do {                                  \
        (ctx)->launches++;                \
        CHX_CUDA(cudaGetLastError());     \
    } while (0)

static inline int chx_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// Exact reference arithmetic (SURVEY.md App. A.1 / chiron/neighbors.py:69-81).
// One IEEE fp32 rounding per operation; the __f*_rn intrinsics are never contracted into FMAs.
// ---------------------------------------------------------------------------------------------
struct Box {
    float lx, ly, lz;
    float hx, hy, hz;  // L * 0.5 (exact)
};

__host__ __device__ inline Box make_box(float lx, float ly, float lz) {
    Box b;
    b.lx = lx; b.ly = ly; b.lz = lz;
    b.hx = lx * 0.5f; b.hy = ly * 0.5f; b.hz = lz * 0.5f;
    return b;
}

// jnp.mod(t, L) for L > 0: fmod (exact) + sign fix-up (one rounded add when the remainder is < 0).
__device__ __forceinline__ float ref_mod(float t, float L) {
    float rem;
    if (t >= 0.0f && t < L) {
        rem = t;                       // fmod is the identity
    } else if (t >= L && t < __fadd_rn(L, L)) {
        rem = __fsub_rn(t, L);         // exact (Sterbenz)
    } else if (t < 0.0f && t > -L) {
        rem = t;                       // |t| < L: fmod keeps t, fix-up below adds L
    } else {
        rem = fmodf(t, L);             // general case, exact
    }
    if (rem != 0.0f && rem < 0.0f) rem = __fadd_rn(rem, L);
    return rem;
}

// One component of the minimum-image displacement: mod(r + L/2, L) - L/2.
__device__ __forceinline__ float ref_minimg(float a, float b, float L, float h) {
    float r = __fsub_rn(a, b);
    return __fsub_rn(ref_mod(__fadd_rn(r, h), L), h);
}

template <bool PERIODIC>
__device__ __forceinline__ void ref_displacement(float ax, float ay, float az, float bx, float by,
                                                 float bz, const Box& box, float& rx, float& ry,
                                                 float& rz, float& d) {
    if (PERIODIC) {
        rx = ref_minimg(ax, bx, box.lx, box.hx);
        ry = ref_minimg(ay, by, box.ly, box.hy);
        rz = ref_minimg(az, bz, box.lz, box.hz);
    } else {
        rx = __fsub_rn(ax, bx);
        ry = __fsub_rn(ay, by);
        rz = __fsub_rn(az, bz);
    }
    float s = __fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz));
    d = __fsqrt_rn(s);
}

// Fast classification of a pair against a cutoff c: the squared distance from FMA arithmetic decides
// every pair that is not within a few ulps of c; only those go through the exact predicate above.
// Kernels that use it evaluate exactly the reference's pair set {d < c} at a fraction of its cost.
struct FastCut {
    float c;             // the cutoff (exact predicate)
    float c2_lo, c2_hi;  // r2 < c2_lo: inside, r2 >= c2_hi: outside
    float inv_lx, inv_ly, inv_lz;
};

__host__ __device__ inline FastCut make_fast_cut(float c, float lx, float ly, float lz, bool periodic) {
    FastCut f;
    f.c = c;
    const double c2 = (double)c * (double)c;
    double lmax = periodic ? (lx > ly ? (lx > lz ? lx : lz) : (ly > lz ? ly : lz)) : 0.0;
    // |fast r2 - d_ref^2| <= ~2 c (a few ulps of the largest coordinate DIFFERENCE; a - b itself is
    // correctly rounded): the band is generous enough for positions spread over +-4 box lengths
    const double band = c2 * 4e-6 + 64.0 * 1.2e-7 * (lmax + 4.0 * (double)c) * (double)c;
    f.c2_lo = (float)(c2 - band);
    f.c2_hi = (float)(c2 + band);
    f.inv_lx = 1.0f / lx; f.inv_ly = 1.0f / ly; f.inv_lz = 1.0f / lz;
    return f;
}

// true iff the reference's predicate d(a, b) < c holds; r2 and (dx, dy, dz) = minimum image of a - b
// from the fast arithmetic (relative error ~1e-7, far inside the 1e-5 energy / force tolerance).
template <bool PERIODIC>
__device__ __forceinline__ bool fast_within(float ax, float ay, float az, float bx, float by, float bz,
                                            const Box& box, const FastCut& fc, float& r2, float& dx,
                                            float& dy, float& dz) {
    dx = ax - bx; dy = ay - by; dz = az - bz;
    if (PERIODIC) {
        dx -= box.lx * rintf(dx * fc.inv_lx);
        dy -= box.ly * rintf(dy * fc.inv_ly);
        dz -= box.lz * rintf(dz * fc.inv_lz);
    }
    r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    if (r2 >= fc.c2_hi) return false;
    if (r2 >= fc.c2_lo) {
        float rx, ry, rz, d;
        ref_displacement<PERIODIC>(ax, ay, az, bx, by, bz, box, rx, ry, rz, d);
        return d < fc.c;
    }
    return true;
}

// LJ pair energy and force scalar (f * (dx,dy,dz) = force on a) from r2, without sqrt or division
__device__ __forceinline__ void lj_pair_r2(float r2, float sigma2, float eps, float& e, float& f_over_r) {
    float inv;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(r2));
    inv = inv * (2.0f - r2 * inv);                 // one Newton step: ~1 ulp
    const float s2 = sigma2 * inv;
    const float s6 = s2 * s2 * s2;
    e = 4.0f * eps * s6 * (s6 - 1.0f);
    f_over_r = 24.0f * eps * inv * s6 * (2.0f * s6 - 1.0f);
}

// Energy of one row of a NeighborListNsqrd half list, this lane's share (entries lane, lane + 32, ...).
// Four entries in flight per lane: the index loads and the position gathers of a batch are issued
// before any arithmetic, so the two dependent memory latencies per entry overlap.  Shared by
// LJPotential.compute_energy (lj.cu) and the Metropolis loop (mc.cu): identical summation order.
// position accessors: the API's (N,3) float array, or a float4 copy (one 16-byte gather per neighbour)
struct Pos3 {
    const float* __restrict__ x;
    __device__ __forceinline__ void get(uint32_t j, float& a, float& b, float& c) const {
        a = x[3 * j]; b = x[3 * j + 1]; c = x[3 * j + 2];
    }
};
struct Pos4 {
    const float4* __restrict__ x;
    __device__ __forceinline__ void get(uint32_t j, float& a, float& b, float& c) const {
        const float4 p = x[j];
        a = p.x; b = p.y; c = p.z;
    }
};

template <bool PERIODIC, class POS>
__device__ __forceinline__ float lj_nlist_row_energy(const POS pos, int i, int lane,
                                                     const Box& box, const FastCut& fc,
                                                     const uint32_t* __restrict__ row, int cnt,
                                                     float sigma2, float eps) {
    float xi, yi, zi;
    pos.get((uint32_t)i, xi, yi, zi);
    float e_row = 0.f;
    for (int k0 = lane; k0 < cnt; k0 += 128) {
        uint32_t j[4];
        float px[4], py[4], pz[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) j[u] = (k0 + 32 * u < cnt) ? row[k0 + 32 * u] : 0xffffffffu;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t jj = j[u] != 0xffffffffu ? j[u] : (uint32_t)i;
            pos.get(jj, px[u], py[u], pz[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float r2, dx, dy, dz;
            if (j[u] != 0xffffffffu &&
                fast_within<PERIODIC>(xi, yi, zi, px[u], py[u], pz[u], box, fc, r2, dx, dy, dz)) {
                float e, f;
                lj_pair_r2(r2, sigma2, eps, e, f);
                e_row += e;
            }
        }
    }
    return e_row;
}

// Space.wrap, one component: x - floor(x / L) * L.
__device__ __forceinline__ float ref_wrap(float x, float L) {
    return __fsub_rn(x, __fmul_rn(floorf(__fdiv_rn(x, L)), L));
}

// ---------------------------------------------------------------------------------------------
// jax.random legacy threefry
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t chx_rotl(uint32_t x, int r) {
    return (x << r) | (x >> (32 - r));
}

__host__ __device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t& x0,
                                                      uint32_t& x1) {
    const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
    x0 += k0; x1 += k1;
#define CHX_TF_ROUND(r) { x0 += x1; x1 = chx_rotl(x1, r); x1 ^= x0; }
    CHX_TF_ROUND(13) CHX_TF_ROUND(15) CHX_TF_ROUND(26) CHX_TF_ROUND(6)
    x0 += k1; x1 += k2 + 1u;
    CHX_TF_ROUND(17) CHX_TF_ROUND(29) CHX_TF_ROUND(16) CHX_TF_ROUND(24)
    x0 += k2; x1 += k0 + 2u;
    CHX_TF_ROUND(13) CHX_TF_ROUND(15) CHX_TF_ROUND(26) CHX_TF_ROUND(6)
    x0 += k0; x1 += k1 + 3u;
    CHX_TF_ROUND(17) CHX_TF_ROUND(29) CHX_TF_ROUND(16) CHX_TF_ROUND(24)
    x0 += k1; x1 += k2 + 4u;
    CHX_TF_ROUND(13) CHX_TF_ROUND(15) CHX_TF_ROUND(26) CHX_TF_ROUND(6)
    x0 += k2; x1 += k0 + 5u;
#undef CHX_TF_ROUND
}

// Element e of random_bits(key, n): counters 0..n-1 (odd n padded with one 0), first half feeds
// the x0 lanes, second half the x1 lanes, outputs concatenated.
__host__ __device__ __forceinline__ uint32_t random_bits_elem(uint32_t k0, uint32_t k1,
                                                              unsigned long long e,
                                                              unsigned long long n) {
    const unsigned long long half = (n + 1ull) >> 1;
    const bool first = e < half;
    const unsigned long long b = first ? e : e - half;
    uint32_t x0 = (uint32_t)b;
    const unsigned long long c1 = b + half;
    uint32_t x1 = (c1 < n) ? (uint32_t)c1 : 0u;
    threefry2x32(k0, k1, x0, x1);
    return first ? x0 : x1;
}

// random.split(key): new key = (bits[0], bits[1]), subkey = (bits[2], bits[3]) of random_bits(key,4).
__host__ __device__ __forceinline__ void threefry_split(uint32_t k0, uint32_t k1, uint32_t& c0,
                                                        uint32_t& c1, uint32_t& s0, uint32_t& s1) {
    uint32_t a0 = 0u, a1 = 2u, b0 = 1u, b1 = 3u;
    threefry2x32(k0, k1, a0, a1);
    threefry2x32(k0, k1, b0, b1);
    c0 = a0; c1 = b0; s0 = a1; s1 = b1;
}

__device__ __forceinline__ float bits_to_unit_float(uint32_t bits) {
    return __fsub_rn(__uint_as_float((bits >> 9) | 0x3F800000u), 1.0f);
}

__device__ __forceinline__ float uniform_from_bits(uint32_t bits, float lo, float hi) {
    float f = bits_to_unit_float(bits);
    return fmaxf(lo, __fadd_rn(__fmul_rn(f, __fsub_rn(hi, lo)), lo));
}

// XLA fp32 ErfInv (Giles).  Separate mul/add like the un-fused HLO evaluation.
__device__ __forceinline__ float erfinv_xla(float x) {
    float w = -log1pf(-__fmul_rn(x, x));
    float p;
    if (w < 5.0f) {
        w = __fsub_rn(w, 2.5f);
        p = 2.81022636e-08f;
        p = __fadd_rn(3.43273939e-07f, __fmul_rn(p, w));
        p = __fadd_rn(-3.5233877e-06f, __fmul_rn(p, w));
        p = __fadd_rn(-4.39150654e-06f, __fmul_rn(p, w));
        p = __fadd_rn(0.00021858087f, __fmul_rn(p, w));
        p = __fadd_rn(-0.00125372503f, __fmul_rn(p, w));
        p = __fadd_rn(-0.00417768164f, __fmul_rn(p, w));
        p = __fadd_rn(0.246640727f, __fmul_rn(p, w));
        p = __fadd_rn(1.50140941f, __fmul_rn(p, w));
    } else {
        w = __fsub_rn(__fsqrt_rn(w), 3.0f);
        p = -0.000200214257f;
        p = __fadd_rn(0.000100950558f, __fmul_rn(p, w));
        p = __fadd_rn(0.00134934322f, __fmul_rn(p, w));
        p = __fadd_rn(-0.00367342844f, __fmul_rn(p, w));
        p = __fadd_rn(0.00573950773f, __fmul_rn(p, w));
        p = __fadd_rn(-0.0076224613f, __fmul_rn(p, w));
        p = __fadd_rn(0.00943887047f, __fmul_rn(p, w));
        p = __fadd_rn(1.00167406f, __fmul_rn(p, w));
        p = __fadd_rn(2.83297682f, __fmul_rn(p, w));
    }
    if (fabsf(x) == 1.0f) return x * __int_as_float(0x7f800000);
    return __fmul_rn(p, x);
}

// random.normal element: sqrt(2) * erfinv(uniform(lo = nextafter(-1, 0), hi = 1)).
__device__ __forceinline__ float normal_from_bits(uint32_t bits) {
    const float lo = -0.99999994f;  // nextafter(-1, 0)
    float u = uniform_from_bits(bits, lo, 1.0f);
    return __fmul_rn(1.41421356f, erfinv_xla(u));
}

// ---------------------------------------------------------------------------------------------
// Reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
