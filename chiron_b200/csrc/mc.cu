// Device-resident Metropolis loop for MonteCarloDisplacementMove (chiron/mcmc.py:243-306 update,
// :357-463 _step, :531-548 _accept_or_reject, :733-787 _propose).
//
// The reference runs every Monte Carlo step from Python: two PRNG splits, a jitted proposal, a full
// energy evaluation and a host-side accept/reject.  Here a move is three launches that never leave
// the device -- propose (+wrap +list check), energy of the proposal, decide (+key advance, counters) --
// and `n_moves` of them are enqueued (as replays of one cached CUDA graph) before the host looks at
// the state again.  The PRNG stream, the proposal arithmetic and the acceptance rule are the
// reference's, so the trajectory is the one the host-driven path (chiron_b200/mcmc.py) produces.
//
// A proposal that would make NeighborListNsqrd.check() true (mcmc.py:754-759: rebuild on the
// proposed positions) stops the loop BEFORE that move (state.halt = 1, key not advanced); the caller
// performs that one move through the building-block path, which rebuilds the list, and re-enters.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "common.cuh"
#include "cell_list.cuh"

struct McArgs {
    chx_mc_displace_args a;
    float* x0;
    float* x1;
    float4* q0;          // float4 copies of x0 / x1 (context scratch): one gather per neighbour
    float4* q1;
    chx_mc_state* st;
    double* acc;
};

__device__ __forceinline__ void lj_pair_e(float d, float sigma, float eps, float& e) {
    const float q = sigma / d;
    const float q2 = q * q;
    const float q6 = q2 * q2 * q2;
    const float q12 = q6 * q6;
    e = 4.0f * eps * (q12 - q6);
}

__device__ __forceinline__ void mc_block_add(double v, double* target) {
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

// ---- propose: x' = wrap(x + sigma * normal(subkey, (n,3)) * mask), list check --------------------
template <bool WRAP, bool CHECK>
__global__ void __launch_bounds__(256)
k_mcl_propose(int n, const float* __restrict__ mask, Box box, const float* __restrict__ ref,
              float half_skin, float* __restrict__ x0, float* __restrict__ x1,
              float4* __restrict__ q0, float4* __restrict__ q1, chx_mc_state* __restrict__ st) {
    if (*((volatile int*)&st->halt)) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool moved = false;
    if (i < n) {
        uint32_t c0, c1, s0, s1;
        threefry_split(st->key[0], st->key[1], c0, c1, s0, s1);   // new_PRNG_key (states.py:150-154)
        const float* xc = st->sel ? x1 : x0;
        float* xp = st->sel ? x0 : x1;
        const float sigma = st->sigma_disp;
        const float m = mask ? mask[i] : 1.0f;
        const unsigned long long total = 3ull * (unsigned long long)n;
        float xn[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            // a zero mask multiplies the (finite) noise to +-0, and x + (+-0) == x: no need to draw it
            float d = 0.0f;
            if (m != 0.0f) {
                d = __fmul_rn(normal_from_bits(random_bits_elem(s0, s1, 3ull * i + c, total)), sigma);
                if (mask) d = __fmul_rn(d, m);
            }
            float v = __fadd_rn(xc[3 * i + c], d);
            if (WRAP) v = ref_wrap(v, c == 0 ? box.lx : (c == 1 ? box.ly : box.lz));
            xp[3 * i + c] = v;
            xn[c] = v;
        }
        (st->sel ? q0 : q1)[i] = make_float4(xn[0], xn[1], xn[2], 0.f);
        if (CHECK) {
            float rx, ry, rz, d;
            ref_displacement<WRAP>(xn[0], xn[1], xn[2], ref[3 * i], ref[3 * i + 1], ref[3 * i + 2], box, rx, ry, rz, d);
            moved = d >= half_skin;
        }
    }
    if (CHECK) {
        if (__syncthreads_or(moved) && threadIdx.x == 0) atomicExch(&st->halt, 1);
    }
}

// ---- energies of the proposal (which = 1) or of the current state (which = 0) --------------------
__device__ __forceinline__ const float* mc_buf(const chx_mc_state* st, const float* x0, const float* x1, int which) {
    return ((st->sel ^ which) & 1) ? x1 : x0;
}

// 8 CTAs of 256 threads per SM (32 registers): the kernel waits on dependent gathers, more resident
// warps buy 12 % (21.7 k -> 24.3 k displacement moves/s at N = 32,768)
#ifndef CHX_MCL_MINB
#define CHX_MCL_MINB 8
#endif
template <bool PERIODIC>
__global__ void __launch_bounds__(256, CHX_MCL_MINB)
k_mcl_lj_nlist(int n, Box box, FastCut fc, const uint32_t* __restrict__ list, const int32_t* __restrict__ nn, int M,
               float sigma, float eps, const float4* __restrict__ q0,
               const float4* __restrict__ q1, const chx_mc_state* __restrict__ st, int which,
               double* __restrict__ acc) {
    if (st->halt) return;
    const float4* x = ((st->sel ^ which) & 1) ? q1 : q0;
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    double e_acc = 0.0;
    if (i < n) {
        // same code as LJPotential.compute_energy (k_lj_nlist, lj.cu): identical reduced potentials
        int cnt = nn[i];
        cnt = cnt < M ? cnt : M;
        const float e_row = lj_nlist_row_energy<PERIODIC>(Pos4{x}, i, lane, box, fc, list + (size_t)i * M, cnt,
                                                          sigma * sigma, eps);
        e_acc = (double)e_row;
    }
    mc_block_add(e_acc, acc);
}

// all pairs i < j: one block per row, threads over j (PairListNsqrd / nbr_list=None, small N)
template <bool PERIODIC>
__global__ void __launch_bounds__(128)
k_mcl_lj_allpairs(int n, Box box, float sigma, float eps, float cutoff, const float* __restrict__ x0,
                  const float* __restrict__ x1, const chx_mc_state* __restrict__ st, int which,
                  double* __restrict__ acc) {
    if (st->halt) return;
    const float* x = mc_buf(st, x0, x1, which);
    const int i = blockIdx.x;
    const float xi = x[3 * i], yi = x[3 * i + 1], zi = x[3 * i + 2];
    double e_acc = 0.0;
    for (int j = i + 1 + threadIdx.x; j < n; j += blockDim.x) {
        float rx, ry, rz, d;
        ref_displacement<PERIODIC>(xi, yi, zi, x[3 * j], x[3 * j + 1], x[3 * j + 2], box, rx, ry, rz, d);
        if (cutoff < 0.0f || d < cutoff) {
            float e;
            lj_pair_e(d, sigma, eps, e);
            e_acc += (double)e;
        }
    }
    mc_block_add(e_acc, acc);
}

__global__ void __launch_bounds__(256)
k_mcl_ho(int n, const float* __restrict__ xref, int n0, const float* __restrict__ x0,
         const float* __restrict__ x1, const chx_mc_state* __restrict__ st, int which,
         double* __restrict__ acc) {
    if (st->halt) return;
    const float* x = mc_buf(st, x0, x1, which);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0.0;
    if (i < n) {
        const int r = n0 == 1 ? 0 : i;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float dx = __fsub_rn(x[3 * i + c], xref[3 * r + c]);
            e += (double)__fmul_rn(dx, dx);
        }
    }
    mc_block_add(e, acc);
}

// delta energy of the moved subset: one block per moved particle (see k_lj_subset_delta, lj.cu)
template <bool PERIODIC>
__global__ void __launch_bounds__(256)
k_mcl_subset_delta(int n, const uint32_t* __restrict__ moved, const float* __restrict__ mask, Box box,
                   float sigma, float eps, float cutoff, const float* __restrict__ x0,
                   const float* __restrict__ x1, const chx_mc_state* __restrict__ st,
                   double* __restrict__ acc) {
    if (st->halt) return;
    const float* xo = mc_buf(st, x0, x1, 0);
    const float* xn = mc_buf(st, x0, x1, 1);
    const int m = (int)moved[blockIdx.x];
    const float ox = xo[3 * m], oy = xo[3 * m + 1], oz = xo[3 * m + 2];
    const float nx = xn[3 * m], ny = xn[3 * m + 1], nz = xn[3 * m + 2];
    double a = 0.0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        if (j == m) continue;
        const double w = mask[j] != 0.0f ? 0.5 : 1.0;
        float rx, ry, rz, d, e;
        // orientation of the reference's half list: displacement(x_lower id, x_higher id)
        if (m < j) ref_displacement<PERIODIC>(nx, ny, nz, xn[3 * j], xn[3 * j + 1], xn[3 * j + 2], box, rx, ry, rz, d);
        else ref_displacement<PERIODIC>(xn[3 * j], xn[3 * j + 1], xn[3 * j + 2], nx, ny, nz, box, rx, ry, rz, d);
        if (d < cutoff) { lj_pair_e(d, sigma, eps, e); a += w * (double)e; }
        if (m < j) ref_displacement<PERIODIC>(ox, oy, oz, xo[3 * j], xo[3 * j + 1], xo[3 * j + 2], box, rx, ry, rz, d);
        else ref_displacement<PERIODIC>(xo[3 * j], xo[3 * j + 1], xo[3 * j + 2], ox, oy, oz, box, rx, ry, rz, d);
        if (d < cutoff) { lj_pair_e(d, sigma, eps, e); a -= w * (double)e; }
    }
    mc_block_add(a, acc);
}

// float4 copy of the current configuration at the start of a call
__global__ void k_mcl_pack(int n, const float* __restrict__ x0, const float* __restrict__ x1,
                           float4* __restrict__ q0, float4* __restrict__ q1,
                           const chx_mc_state* __restrict__ st) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* x = st->sel ? x1 : x0;
    (st->sel ? q1 : q0)[i] = make_float4(x[3 * i], x[3 * i + 1], x[3 * i + 2], 0.f);
}

// ---- reduced potential and the Metropolis decision ---------------------------------------------------
struct McThermo {
    int potential;
    float k, U0;        // harmonic oscillator
    double beta, pv;    // u = beta * (U + pv)
};

__device__ __forceinline__ float mc_reduced(const McThermo& t, double acc) {
    float U = (float)acc;
    if (t.potential == CHX_MC_HO) U = __fadd_rn(__fmul_rn(__fmul_rn(0.5f, t.k), U), t.U0);   // potential.py:413-418
    if (t.potential == CHX_MC_IDEAL) U = 0.0f;
    return (float)(t.beta * ((double)U + t.pv));
}

// u of the current state (first entry of a loop, or after the caller changed the state)
__global__ void k_mcl_init(McThermo t, chx_mc_state* __restrict__ st, double* __restrict__ acc) {
    const double a = *acc;
    *acc = 0.0;
    st->u_current = mc_reduced(t, a);
    st->have_u = 1;
}

__global__ void k_mcl_decide(McThermo t, chx_mc_state* __restrict__ st, double* __restrict__ acc) {
    const double a = *acc;
    *acc = 0.0;
    if (st->halt) return;
    const float u_cur = st->u_current;
    float u_new;
    if (t.potential == CHX_MC_LJ_SUBSET_DELTA) u_new = __fadd_rn(u_cur, (float)(a * t.beta));
    else u_new = mc_reduced(t, a);
    const float lr = __fadd_rn(-u_new, u_cur);                    // mcmc.py:777
    uint32_t c0, c1, s0, s1;
    threefry_split(st->key[0], st->key[1], c0, c1, s0, s1);       // the split _propose made
    bool accept = false;
    if (u_new != u_new) {
        // NaN energy: rejected without drawing the uniform (mcmc.py:417-430)
        st->nan_seen += 1;
    } else {
        uint32_t d0, d1, t0, t1;
        threefry_split(c0, c1, d0, d1, t0, t1);                   // proposed_sampler_state.new_PRNG_key
        c0 = d0; c1 = d1;
        const float uni = uniform_from_bits(random_bits_elem(t0, t1, 0ull, 1ull), 0.0f, 1.0f);
        accept = (-lr <= 0.0f) || (uni < expf(lr));                // mcmc.py:541-548
    }
    st->key[0] = c0; st->key[1] = c1;
    if (accept) {
        st->sel ^= 1;
        st->u_current = u_new;
        st->n_accepted += 1;
    }
    st->n_proposed += 1;
    st->moves_done += 1;
}

// ---- host side ---------------------------------------------------------------------------------------
static int mc_launch_energy(chx_ctx* ctx, const McArgs& m, int which) {
    const chx_mc_displace_args& a = m.a;
    const Box box = make_box(a.lx, a.ly, a.lz);
    cudaStream_t st = ctx->stream;
    switch (a.potential) {
    case CHX_MC_LJ_NLIST: {
        const FastCut fc = make_fast_cut(a.cutoff, a.lx, a.ly, a.lz, a.periodic != 0);
        if (a.periodic)
            k_mcl_lj_nlist<true><<<chx_div_up(a.n, 8), 256, 0, st>>>(a.n, box, fc, a.neighbor_list, a.n_neighbors, a.M, a.sigma, a.epsilon, m.q0, m.q1, m.st, which, m.acc);
        else
            k_mcl_lj_nlist<false><<<chx_div_up(a.n, 8), 256, 0, st>>>(a.n, box, fc, a.neighbor_list, a.n_neighbors, a.M, a.sigma, a.epsilon, m.q0, m.q1, m.st, which, m.acc);
        break;
    }
    case CHX_MC_LJ_SUBSET_DELTA:
        if (which == 0) {
            // the loop only needs differences; u of the current state is whatever the caller set
            return CHX_OK;
        }
        if (a.periodic)
            k_mcl_subset_delta<true><<<a.n_subset, 256, 0, st>>>(a.n, a.subset_ids, a.subset_mask, box, a.sigma, a.epsilon, a.cutoff, m.x0, m.x1, m.st, m.acc);
        else
            k_mcl_subset_delta<false><<<a.n_subset, 256, 0, st>>>(a.n, a.subset_ids, a.subset_mask, box, a.sigma, a.epsilon, a.cutoff, m.x0, m.x1, m.st, m.acc);
        break;
    case CHX_MC_LJ_ALLPAIRS:
        if (a.periodic)
            k_mcl_lj_allpairs<true><<<a.n, 128, 0, st>>>(a.n, box, a.sigma, a.epsilon, a.cutoff, m.x0, m.x1, m.st, which, m.acc);
        else
            k_mcl_lj_allpairs<false><<<a.n, 128, 0, st>>>(a.n, box, a.sigma, a.epsilon, a.cutoff, m.x0, m.x1, m.st, which, m.acc);
        break;
    case CHX_MC_HO:
        k_mcl_ho<<<chx_div_up(a.n, 256), 256, 0, st>>>(a.n, a.x0, a.n0, m.x0, m.x1, m.st, which, m.acc);
        break;
    case CHX_MC_IDEAL:
        return CHX_OK;   // U = 0 (potential.py:93-127): nothing to launch
    default:
        chx_set_error("unknown potential kind %d", a.potential);
        return CHX_BAD_ARG;
    }
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

static McThermo mc_thermo(const chx_mc_displace_args& a) {
    McThermo t;
    t.potential = a.potential; t.k = a.k; t.U0 = a.U0; t.beta = a.beta; t.pv = a.pv;
    return t;
}

static int mc_launch_move(chx_ctx* ctx, const McArgs& m) {
    const chx_mc_displace_args& a = m.a;
    const Box box = make_box(a.lx, a.ly, a.lz);
    const bool check = a.ref_positions != nullptr;
    const int blocks = chx_div_up(a.n, 256);
    const float hs = 0.5f * a.skin;
#define PROPOSE(W, C)                                                                                   \
    k_mcl_propose<W, C><<<blocks, 256, 0, ctx->stream>>>(a.n, a.subset_mask, box, a.ref_positions, hs, \
                                                         m.x0, m.x1, m.q0, m.q1, m.st)
    if (a.periodic) { if (check) PROPOSE(true, true); else PROPOSE(true, false); }
    else { if (check) PROPOSE(false, true); else PROPOSE(false, false); }
#undef PROPOSE
    CHX_LAUNCHED(ctx);
    int rc = mc_launch_energy(ctx, m, 1);
    if (rc != CHX_OK) return rc;
    k_mcl_decide<<<1, 1, 0, ctx->stream>>>(mc_thermo(a), m.st, m.acc);
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

// launches per move, for the context's launch counter when a graph is replayed
static int mc_launches_per_move(const chx_mc_displace_args& a) { return a.potential == CHX_MC_IDEAL ? 2 : 3; }

// one cached graph of MC_GRAPH_MOVES moves per distinct argument set (a handful at most)
#define MC_GRAPH_MOVES 10
struct McGraph {
    chx_ctx* ctx;
    McArgs key;
    cudaGraphExec_t exec;
    unsigned long long stamp;
};
static std::vector<McGraph> g_mc_graphs;
static unsigned long long g_mc_stamp = 0;

static int mc_graph_get(chx_ctx* ctx, const McArgs& m, cudaGraphExec_t* out) {
    for (auto& g : g_mc_graphs)
        if (g.ctx == ctx && memcmp(&g.key, &m, sizeof(McArgs)) == 0) {
            g.stamp = ++g_mc_stamp;
            *out = g.exec;
            return CHX_OK;
        }
    cudaStream_t user = ctx->stream, cap = nullptr;
    CHX_CUDA(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
    ctx->stream = cap;
    const long long launches_before = ctx->launches;
    cudaGraph_t graph = nullptr;
    cudaError_t err = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
    int rc = CHX_OK;
    if (err == cudaSuccess) {
        for (int k = 0; k < MC_GRAPH_MOVES && rc == CHX_OK; ++k) rc = mc_launch_move(ctx, m);
        err = cudaStreamEndCapture(cap, &graph);
    }
    ctx->stream = user;
    ctx->launches = launches_before;
    cudaGraphExec_t exec = nullptr;
    if (err == cudaSuccess && rc == CHX_OK) err = cudaGraphInstantiate(&exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    cudaStreamDestroy(cap);
    if (rc != CHX_OK) return rc;
    if (err != cudaSuccess) {
        chx_set_error("Monte Carlo graph capture failed: %s", cudaGetErrorString(err));
        return CHX_CUDA_ERROR;
    }
    if (g_mc_graphs.size() >= 16) {   // drop the least recently used
        size_t lru = 0;
        for (size_t k = 1; k < g_mc_graphs.size(); ++k)
            if (g_mc_graphs[k].stamp < g_mc_graphs[lru].stamp) lru = k;
        cudaGraphExecDestroy(g_mc_graphs[lru].exec);
        g_mc_graphs.erase(g_mc_graphs.begin() + lru);
    }
    McGraph g;
    g.ctx = ctx; g.exec = exec; g.stamp = ++g_mc_stamp;
    memcpy(&g.key, &m, sizeof(McArgs));
    g_mc_graphs.push_back(g);
    *out = exec;
    return CHX_OK;
}

// ---- small systems: the whole loop in ONE launch --------------------------------------------------------
// For n <= MC_SMALL_MAX_N a move is far too little work for three launches (the loop above runs at the
// graph-node latency floor, ~2 us per node).  One CTA keeps both position buffers in shared memory and
// runs all n_moves moves back to back: propose (+wrap, +check), energy, decision by thread 0, swap of the
// buffer roles.  Same PRNG stream, same arithmetic per term, so the trajectory equals the multi-launch
// path's (sums differ in the order of fp64 additions only).
#define MC_SMALL_MAX_N 2048
#define MC_SMALL_THREADS 512

__device__ __forceinline__ double mc_cta_sum(double v, double* sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();                 // sh may still be read from the previous call
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int k = 0; k < (MC_SMALL_THREADS >> 5); ++k) t += sh[k];   // same order in every thread
    return t;
}

template <int POT, bool PERIODIC>
__device__ __forceinline__ double mc_small_energy(const chx_mc_displace_args& a, const Box& box, const FastCut& fc,
                                                  const float* xa, const float* xb, double* sh) {
    // U of xb (HO, list LJ) or U(xb) - U(xa) (subset delta); xa / xb live in shared memory
    double e = 0.0;
    const int n = a.n;
    if (POT == CHX_MC_HO) {
        for (int i = threadIdx.x; i < n; i += MC_SMALL_THREADS) {
            const int r = a.n0 == 1 ? 0 : i;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float dx = __fsub_rn(xb[3 * i + c], a.x0[3 * r + c]);
                e += (double)__fmul_rn(dx, dx);
            }
        }
    } else if (POT == CHX_MC_LJ_NLIST) {
        const int lane = threadIdx.x & 31;
        for (int i = threadIdx.x >> 5; i < n; i += MC_SMALL_THREADS >> 5) {
            int cnt = a.n_neighbors[i];
            cnt = cnt < a.M ? cnt : a.M;
            e += (double)lj_nlist_row_energy<PERIODIC>(Pos3{xb}, i, lane, box, fc, a.neighbor_list + (size_t)i * a.M,
                                                       cnt, a.sigma * a.sigma, a.epsilon);
        }
    } else if (POT == CHX_MC_LJ_SUBSET_DELTA) {
        for (int s = 0; s < a.n_subset; ++s) {
            const int m = (int)a.subset_ids[s];
            const float ox = xa[3 * m], oy = xa[3 * m + 1], oz = xa[3 * m + 2];
            const float nx = xb[3 * m], ny = xb[3 * m + 1], nz = xb[3 * m + 2];
            for (int j = threadIdx.x; j < n; j += MC_SMALL_THREADS) {
                if (j == m) continue;
                const double w = a.subset_mask[j] != 0.0f ? 0.5 : 1.0;
                float rx, ry, rz, d, ee;
                if (m < j) ref_displacement<PERIODIC>(nx, ny, nz, xb[3 * j], xb[3 * j + 1], xb[3 * j + 2], box, rx, ry, rz, d);
                else ref_displacement<PERIODIC>(xb[3 * j], xb[3 * j + 1], xb[3 * j + 2], nx, ny, nz, box, rx, ry, rz, d);
                if (d < a.cutoff) { lj_pair_e(d, a.sigma, a.epsilon, ee); e += w * (double)ee; }
                if (m < j) ref_displacement<PERIODIC>(ox, oy, oz, xa[3 * j], xa[3 * j + 1], xa[3 * j + 2], box, rx, ry, rz, d);
                else ref_displacement<PERIODIC>(xa[3 * j], xa[3 * j + 1], xa[3 * j + 2], ox, oy, oz, box, rx, ry, rz, d);
                if (d < a.cutoff) { lj_pair_e(d, a.sigma, a.epsilon, ee); e -= w * (double)ee; }
            }
        }
    }
    return mc_cta_sum(e, sh);
}

template <int POT, bool PERIODIC, bool CHECK>
__global__ void __launch_bounds__(MC_SMALL_THREADS)
k_mcl_small(chx_mc_displace_args a, McThermo th, FastCut fc, float* __restrict__ x0, float* __restrict__ x1,
            chx_mc_state* __restrict__ st, int n_moves) {
    extern __shared__ float sm_pos[];
    __shared__ double sh[MC_SMALL_THREADS >> 5];
    __shared__ chx_mc_state s_st;
    __shared__ int s_accept;
    const int n = a.n;
    float* xa = sm_pos;
    float* xb = sm_pos + 3 * n;
    const Box box = make_box(a.lx, a.ly, a.lz);
    if (threadIdx.x == 0) s_st = *st;
    __syncthreads();
    float* xg = s_st.sel ? x1 : x0;
    for (int t = threadIdx.x; t < 3 * n; t += MC_SMALL_THREADS) { xa[t] = xg[t]; xb[t] = xg[t]; }
    __syncthreads();
    if (!s_st.have_u) {
        double u0 = 0.0;
        if (POT != CHX_MC_LJ_SUBSET_DELTA && POT != CHX_MC_IDEAL) u0 = mc_small_energy<POT, PERIODIC>(a, box, fc, xa, xa, sh);
        if (threadIdx.x == 0) { s_st.u_current = mc_reduced(th, u0); s_st.have_u = 1; }
        __syncthreads();
    }
    const float half_skin = 0.5f * a.skin;
    for (int mv = 0; mv < n_moves; ++mv) {
        uint32_t c0, c1, s0, s1;
        threefry_split(s_st.key[0], s_st.key[1], c0, c1, s0, s1);
        const float sigma = s_st.sigma_disp;
        const unsigned long long total = 3ull * (unsigned long long)n;
        bool moved = false;
        for (int i = threadIdx.x; i < n; i += MC_SMALL_THREADS) {
            const float msk = a.subset_mask ? a.subset_mask[i] : 1.0f;
            float xn[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                // a zero mask multiplies the (finite) noise to +-0, and x + (+-0) == x: no need to draw it
                float d = 0.0f;
                if (msk != 0.0f) {
                    d = __fmul_rn(normal_from_bits(random_bits_elem(s0, s1, 3ull * i + c, total)), sigma);
                    if (a.subset_mask) d = __fmul_rn(d, msk);
                }
                float v = __fadd_rn(xa[3 * i + c], d);
                if (PERIODIC) v = ref_wrap(v, c == 0 ? box.lx : (c == 1 ? box.ly : box.lz));
                xb[3 * i + c] = v;
                xn[c] = v;
            }
            if (CHECK) {
                float rx, ry, rz, d;
                ref_displacement<PERIODIC>(xn[0], xn[1], xn[2], a.ref_positions[3 * i], a.ref_positions[3 * i + 1],
                                           a.ref_positions[3 * i + 2], box, rx, ry, rz, d);
                moved = moved || d >= half_skin;
            }
        }
        const int any_moved = __syncthreads_or(CHECK && moved);     // also publishes xb
        if (any_moved) {
            if (threadIdx.x == 0) s_st.halt = 1;
            break;
        }
        double acc = 0.0;
        if (POT != CHX_MC_IDEAL) acc = mc_small_energy<POT, PERIODIC>(a, box, fc, xa, xb, sh);
        if (threadIdx.x == 0) {
            const float u_cur = s_st.u_current;
            float u_new;
            if (POT == CHX_MC_LJ_SUBSET_DELTA) u_new = __fadd_rn(u_cur, (float)(acc * th.beta));
            else u_new = mc_reduced(th, acc);
            const float lr = __fadd_rn(-u_new, u_cur);
            bool accept = false;
            if (u_new != u_new) {
                s_st.nan_seen += 1;
            } else {
                uint32_t d0, d1, t0, t1;
                threefry_split(c0, c1, d0, d1, t0, t1);
                c0 = d0; c1 = d1;
                const float uni = uniform_from_bits(random_bits_elem(t0, t1, 0ull, 1ull), 0.0f, 1.0f);
                accept = (-lr <= 0.0f) || (uni < expf(lr));
            }
            s_st.key[0] = c0; s_st.key[1] = c1;
            if (accept) { s_st.u_current = u_new; s_st.n_accepted += 1; }
            s_st.n_proposed += 1;
            s_st.moves_done += 1;
            s_accept = accept ? 1 : 0;
        }
        __syncthreads();
        if (s_accept) { float* t = xa; xa = xb; xb = t; }
        __syncthreads();
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 3 * n; t += MC_SMALL_THREADS) xg[t] = xa[t];
    if (threadIdx.x == 0) *st = s_st;
}

static int mc_launch_small(chx_ctx* ctx, const McArgs& m, int n_moves) {
    const chx_mc_displace_args& a = m.a;
    const McThermo th = mc_thermo(a);
    const FastCut fc = make_fast_cut(a.cutoff > 0.f ? a.cutoff : 1.f, a.lx, a.ly, a.lz, a.periodic != 0);
    const size_t smem = (size_t)6 * a.n * sizeof(float);
    const bool chk = a.ref_positions != nullptr;
#define SMALL_LAUNCH(POT, P, C)                                                                              \
    do {                                                                                                     \
        CHX_CUDA(cudaFuncSetAttribute(k_mcl_small<POT, P, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_mcl_small<POT, P, C><<<1, MC_SMALL_THREADS, smem, ctx->stream>>>(a, th, fc, m.x0, m.x1, m.st, n_moves); \
    } while (0)
#define SMALL_POT(POT)                                                                                       \
    do {                                                                                                     \
        if (a.periodic) { if (chk) SMALL_LAUNCH(POT, true, true); else SMALL_LAUNCH(POT, true, false); }      \
        else { if (chk) SMALL_LAUNCH(POT, false, true); else SMALL_LAUNCH(POT, false, false); }               \
    } while (0)
    switch (a.potential) {
    case CHX_MC_HO: SMALL_POT(CHX_MC_HO); break;
    case CHX_MC_IDEAL: SMALL_POT(CHX_MC_IDEAL); break;
    case CHX_MC_LJ_NLIST: SMALL_POT(CHX_MC_LJ_NLIST); break;
    case CHX_MC_LJ_SUBSET_DELTA: SMALL_POT(CHX_MC_LJ_SUBSET_DELTA); break;
    default: chx_set_error("potential kind %d has no single-CTA loop", a.potential); return CHX_BAD_ARG;
    }
#undef SMALL_POT
#undef SMALL_LAUNCH
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

// chx_context_destroy: the graphs captured for this context go with it (a later context may reuse
// the address)
void chx_mc_forget_context(chx_ctx* ctx) {
    for (size_t k = 0; k < g_mc_graphs.size();) {
        if (g_mc_graphs[k].ctx == ctx) {
            cudaGraphExecDestroy(g_mc_graphs[k].exec);
            g_mc_graphs.erase(g_mc_graphs.begin() + k);
        } else {
            ++k;
        }
    }
}

extern "C" {

int chx_mc_displace_run(chx_ctx* ctx, const chx_mc_displace_args* args, float* x0, float* x1,
                        chx_mc_state* state_dev, chx_mc_state* state_host, int n_moves) {
    CHX_REQUIRE(ctx && args && x0 && x1 && state_dev && state_host, "NULL argument");
    CHX_REQUIRE(args->n > 0 && n_moves >= 0, "n must be positive, n_moves non-negative");
    const chx_mc_displace_args& a = *args;
    if (a.potential == CHX_MC_LJ_NLIST)
        CHX_REQUIRE(a.neighbor_list && a.n_neighbors && a.ref_positions && a.M > 0, "neighbour list arrays missing");
    if (a.potential == CHX_MC_LJ_SUBSET_DELTA)
        CHX_REQUIRE(a.subset_ids && a.subset_mask && a.n_subset > 0, "subset ids / mask missing");
    if (a.potential == CHX_MC_HO) CHX_REQUIRE(a.x0 && (a.n0 == 1 || a.n0 == a.n), "x0 must have 1 or n rows");
    McArgs m;
    memset(&m, 0, sizeof(m));   // padding bytes too: the struct is the graph cache key
    m.a = a; m.x0 = x0; m.x1 = x1; m.st = state_dev;
    // context scratch: [acc 256 B][q0 n float4][q1 n float4]
    unsigned char* scratch = (unsigned char*)chx_scratch(ctx, 256 + 2 * (size_t)a.n * sizeof(float4));
    if (!scratch) return CHX_CUDA_ERROR;
    m.acc = (double*)scratch;
    m.q0 = (float4*)(scratch + 256);
    m.q1 = m.q0 + a.n;
    cudaStream_t st = ctx->stream;
    state_host->halt = 0;
    state_host->moves_done = 0;
    CHX_CUDA(cudaMemcpyAsync(state_dev, state_host, sizeof(chx_mc_state), cudaMemcpyHostToDevice, st));
    {
        // small systems: the whole loop in one single-CTA launch (CHX_MC_NO_SMALL=1 keeps the launches per move)
        static int use_small = -1;
        if (use_small < 0) { const char* e = getenv("CHX_MC_NO_SMALL"); use_small = (e && e[0] == '1') ? 0 : 1; }
        const bool small_pot = a.potential == CHX_MC_HO || a.potential == CHX_MC_IDEAL ||
                               a.potential == CHX_MC_LJ_SUBSET_DELTA ||
                               (a.potential == CHX_MC_LJ_NLIST && (long long)a.n * a.M <= 16384);   // one CTA: tiny lists only
        if (use_small && small_pot && a.n <= MC_SMALL_MAX_N && n_moves > 0) {
            int rc = mc_launch_small(ctx, m, n_moves);
            if (rc != CHX_OK) return rc;
            CHX_CUDA(cudaMemcpyAsync(state_host, state_dev, sizeof(chx_mc_state), cudaMemcpyDeviceToHost, st));
            CHX_CUDA(cudaStreamSynchronize(st));
            return CHX_OK;
        }
    }
    CHX_CUDA(cudaMemsetAsync(m.acc, 0, sizeof(double), st));
    k_mcl_pack<<<chx_div_up(a.n, 256), 256, 0, st>>>(a.n, x0, x1, m.q0, m.q1, state_dev);
    CHX_LAUNCHED(ctx);
    int rc = CHX_OK;
    if (!state_host->have_u) {
        rc = mc_launch_energy(ctx, m, 0);
        if (rc != CHX_OK) return rc;
        k_mcl_init<<<1, 1, 0, st>>>(mc_thermo(a), state_dev, m.acc);
        CHX_LAUNCHED(ctx);
    }
    int left = n_moves;
    static int use_graph = -1;
    if (use_graph < 0) { const char* e = getenv("CHX_MC_NOGRAPH"); use_graph = (e && e[0] == '1') ? 0 : 1; }
    if (use_graph && left >= MC_GRAPH_MOVES) {
        cudaGraphExec_t exec = nullptr;
        rc = mc_graph_get(ctx, m, &exec);
        if (rc != CHX_OK) return rc;
        while (left >= MC_GRAPH_MOVES) {
            CHX_CUDA(cudaGraphLaunch(exec, st));
            ctx->launches += MC_GRAPH_MOVES * mc_launches_per_move(a);
            left -= MC_GRAPH_MOVES;
        }
    }
    for (; left > 0 && rc == CHX_OK; --left) rc = mc_launch_move(ctx, m);
    if (rc != CHX_OK) return rc;
    CHX_CUDA(cudaMemcpyAsync(state_host, state_dev, sizeof(chx_mc_state), cudaMemcpyDeviceToHost, st));
    CHX_CUDA(cudaStreamSynchronize(st));
    return CHX_OK;
}

}  // extern "C"

// =====================================================================================================
// Barostat loop (chiron/mcmc.py:913-1009): propose a volume, rebuild the list on the scaled system,
// energy, decide -- all on the device.  The geometry of a proposal (cell grid, sweep and energy
// cutoffs of the proposed box) lives in device memory because it changes with every move.
// =====================================================================================================
struct McbMove {
    CellParams P;          // list build on the proposed box
    FastCut fc;            // energy cutoff with the proposed box (or the current one: initial energy)
    float box_new[3];
    float volume_box;      // lx' ly' lz' (the volume get_reduced_potential sees, states.py:313-323)
    float log_volume;      // N log(V1 / V0) (mcmc.py:1000-1003)
};

struct McbArgs {
    chx_mc_barostat_args a;
    float* x0;
    float* x1;
    float4* q0;
    float4* q1;
    chx_mc_baro_state* st;
    double* acc;
    int* stats;            // [0] max row count, [1] rows with count == M of the list just built
    McbMove* mv;
    float4* xs4;
    int* cell_of;
    int* count;
    int* start;
    int* order;
    int ncell_cap;
};

__global__ void __launch_bounds__(256)
k_mcb_propose(int n, float cutoff, float list_radius, int ncell_cap, float* __restrict__ x0,
              float* __restrict__ x1, float4* __restrict__ q0, float4* __restrict__ q1,
              chx_mc_baro_state* __restrict__ st, McbMove* __restrict__ mv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (*((volatile int*)&st->halt)) {
        if (i == 0) mv->P.valid = 0;      // nothing of this move runs
        return;
    }
    // scalar chain of mcmc.py:956-974 in fp32 (every thread, identical)
    uint32_t c0, c1, s0, s1;
    threefry_split(st->key[0], st->key[1], c0, c1, s0, s1);
    const float lx = st->box[0], ly = st->box[1], lz = st->box[2];
    const float v0 = __fmul_rn(__fmul_rn(lx, ly), lz);
    const float dvmax = __fmul_rn(st->volume_max_scale, v0);
    const float dv = __fmul_rn(uniform_from_bits(random_bits_elem(s0, s1, 0ull, 1ull), -1.0f, 1.0f), dvmax);
    const float v1 = __fadd_rn(v0, dv);
    const float ratio = __fdiv_rn(v1, v0);
    const float sc = powf(ratio, 0.33333334f);
    if (i < n) {
        const float* xc = st->sel ? x1 : x0;
        float* xp = st->sel ? x0 : x1;
        const float a = __fmul_rn(xc[3 * i], sc), b = __fmul_rn(xc[3 * i + 1], sc), c = __fmul_rn(xc[3 * i + 2], sc);
        xp[3 * i] = a; xp[3 * i + 1] = b; xp[3 * i + 2] = c;
        (st->sel ? q0 : q1)[i] = make_float4(a, b, c, 0.f);
    }
    if (i == 0) {
        const float nlx = __fmul_rn(lx, sc), nly = __fmul_rn(ly, sc), nlz = __fmul_rn(lz, sc);
        mv->box_new[0] = nlx; mv->box_new[1] = nly; mv->box_new[2] = nlz;
        mv->volume_box = __fmul_rn(__fmul_rn(nlx, nly), nlz);
        mv->log_volume = __fmul_rn((float)n, logf(ratio));
        mv->fc = make_fast_cut(cutoff, nlx, nly, nlz, true);
        CellParams P;
        const bool ok = v1 > 0.0f && make_cell_params(nlx, nly, nlz, list_radius, ncell_cap, P);
        if (ok) mv->P = P;
        else { mv->P.valid = 0; atomicExch(&st->halt, 1); }
    }
}

// geometry of the CURRENT box (initial energy of a call)
__global__ void k_mcb_current(float cutoff, const chx_mc_baro_state* __restrict__ st, McbMove* __restrict__ mv) {
    mv->fc = make_fast_cut(cutoff, st->box[0], st->box[1], st->box[2], true);
    mv->volume_box = __fmul_rn(__fmul_rn(st->box[0], st->box[1]), st->box[2]);
    mv->P.valid = 0;
}

__global__ void k_mcb_pack(int n, const float* __restrict__ x0, const float* __restrict__ x1,
                           float4* __restrict__ q0, float4* __restrict__ q1,
                           const chx_mc_baro_state* __restrict__ st) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* x = st->sel ? x1 : x0;
    (st->sel ? q1 : q0)[i] = make_float4(x[3 * i], x[3 * i + 1], x[3 * i + 2], 0.f);
}

// energy over list set (sel ^ which) at positions q(sel ^ which); box and cutoffs from mv
__global__ void __launch_bounds__(256, 6)
k_mcb_lj_nlist(int n, int M, float sigma, float eps, const uint32_t* __restrict__ list0,
               const uint32_t* __restrict__ list1, const int32_t* __restrict__ nn0,
               const int32_t* __restrict__ nn1, const float4* __restrict__ q0, const float4* __restrict__ q1,
               const chx_mc_baro_state* __restrict__ st, const McbMove* __restrict__ mv, int which,
               double* __restrict__ acc) {
    if (st->halt) return;
    const int set = (st->sel ^ which) & 1;
    const float4* x = set ? q1 : q0;
    const uint32_t* list = set ? list1 : list0;
    const int32_t* nn = set ? nn1 : nn0;
    const FastCut fc = mv->fc;
    const Box box = which ? make_box(mv->box_new[0], mv->box_new[1], mv->box_new[2])
                          : make_box(st->box[0], st->box[1], st->box[2]);
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    double e_acc = 0.0;
    if (i < n) {
        int cnt = nn[i];
        cnt = cnt < M ? cnt : M;
        e_acc = (double)lj_nlist_row_energy<true>(Pos4{x}, i, lane, box, fc, list + (size_t)i * M, cnt,
                                                  sigma * sigma, eps);
    }
    mc_block_add(e_acc, acc);
}

__global__ void k_mcb_init(double beta, double pressure, chx_mc_baro_state* __restrict__ st,
                           const McbMove* __restrict__ mv, double* __restrict__ acc) {
    const double a = *acc;
    *acc = 0.0;
    st->u_current = (float)(beta * ((double)(float)a + pressure * (double)mv->volume_box));
    st->have_u = 1;
}

__global__ void k_mcb_decide(int M, double beta, double pressure, chx_mc_baro_state* __restrict__ st,
                             const McbMove* __restrict__ mv, double* __restrict__ acc, int* __restrict__ stats) {
    const double a = *acc;
    *acc = 0.0;
    const int max_count = stats[0], eq_m = stats[1];
    stats[0] = 0; stats[1] = 0;
    if (st->halt) return;
    (void)eq_m;
    if (max_count >= M) {                 // n_max_neighbors must grow (neighbors.py:700-729): caller's job
        st->halt = 1;
        return;
    }
    const float u_cur = st->u_current;
    const float u_new = (float)(beta * ((double)(float)a + pressure * (double)mv->volume_box));
    const float lr = __fadd_rn(-__fsub_rn(u_new, u_cur), mv->log_volume);      // mcmc.py:1004-1006
    uint32_t c0, c1, s0, s1;
    threefry_split(st->key[0], st->key[1], c0, c1, s0, s1);
    bool accept = false;
    if (u_new != u_new) {
        st->nan_seen += 1;
    } else {
        uint32_t d0, d1, t0, t1;
        threefry_split(c0, c1, d0, d1, t0, t1);
        c0 = d0; c1 = d1;
        const float uni = uniform_from_bits(random_bits_elem(t0, t1, 0ull, 1ull), 0.0f, 1.0f);
        accept = (-lr <= 0.0f) || (uni < expf(lr));
    }
    st->key[0] = c0; st->key[1] = c1;
    st->last_volume = mv->volume_box;
    if (accept) {
        st->sel ^= 1;
        st->u_current = u_new;
        st->box[0] = mv->box_new[0]; st->box[1] = mv->box_new[1]; st->box[2] = mv->box_new[2];
        st->n_accepted += 1;
    }
    st->n_proposed += 1;
    st->moves_done += 1;
}

static int mcb_launch_energy(chx_ctx* ctx, const McbArgs& m, int which) {
    const chx_mc_barostat_args& a = m.a;
    k_mcb_lj_nlist<<<chx_div_up(a.n, 8), 256, 0, ctx->stream>>>(
        a.n, a.M, a.sigma, a.epsilon, a.neighbor_list[0], a.neighbor_list[1], a.n_neighbors[0], a.n_neighbors[1],
        m.q0, m.q1, m.st, m.mv, which, m.acc);
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

// List of the proposal by filtering the superset list (built once per call at radius (cutoff+skin)(1+delta)):
// one warp per row, exact predicate on the scaled positions, ballot compaction keeps the id order, so the
// arrays equal the ones the 125-cell sweep writes.
__global__ void __launch_bounds__(256, 6)
k_mcb_filter(int n, int M, const uint32_t* __restrict__ sup_list, const int32_t* __restrict__ sup_nn,
             const float4* __restrict__ q0, const float4* __restrict__ q1,
             const chx_mc_baro_state* __restrict__ st, const McbMove* __restrict__ mv,
             uint32_t* list0, uint32_t* list1, int32_t* mask0, int32_t* mask1, int32_t* nn0, int32_t* nn1) {
    if (st->halt) return;
    const int tgt = (st->sel ^ 1) & 1;
    const float4* x = tgt ? q1 : q0;
    uint32_t* list = tgt ? list1 : list0;
    int32_t* mask = tgt ? mask1 : mask0;
    int32_t* nn = tgt ? nn1 : nn0;
    const Box box = make_box(mv->box_new[0], mv->box_new[1], mv->box_new[2]);
    FastCut fc;
    fc.c = mv->P.sc.c; fc.c2_lo = mv->P.sc.c2_lo; fc.c2_hi = mv->P.sc.c2_hi;
    fc.inv_lx = mv->P.sc.inv_lx; fc.inv_ly = mv->P.sc.inv_ly; fc.inv_lz = mv->P.sc.inv_lz;
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    const float4 xi = x[i];
    int cnt = sup_nn[i];
    cnt = cnt < M ? cnt : M;
    int count = 0;
    uint32_t first = 0u;
    const uint32_t* row = sup_list + (size_t)i * M;
    // four chunks of 32 entries in flight: the index loads and the position gathers are issued before
    // any test, so the two dependent memory latencies per entry overlap
    for (int k0 = 0; k0 < cnt; k0 += 128) {
        uint32_t j[4];
        float4 p[4];
        bool hit[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = k0 + 32 * u + lane;
            j[u] = k < cnt ? row[k] : 0xffffffffu;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) p[u] = x[j[u] != 0xffffffffu ? j[u] : (uint32_t)i];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float r2, dx, dy, dz;
            hit[u] = j[u] != 0xffffffffu &&
                     fast_within<true>(xi.x, xi.y, xi.z, p[u].x, p[u].y, p[u].z, box, fc, r2, dx, dy, dz);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned b = __ballot_sync(0xffffffffu, hit[u]);
            if (b) {
                if (count == 0) first = __shfl_sync(0xffffffffu, j[u], __ffs(b) - 1);
                if (hit[u]) {
                    const int pos = count + __popc(b & ((1u << lane) - 1u));
                    if (pos < M) list[(size_t)i * M + pos] = j[u];
                }
                count += __popc(b);
            }
        }
    }
    // Incremental row finalisation (neighbors.py:606-609 semantics as in finish_row): the target set holds
    // a complete, valid row from an earlier build (both sets start as copies of the current list), so
    // only what differs is written -- the pad mask between the old and the new count, and the padding ids
    // where entries turned into padding or when the fill value changed.  Saves the 2 x N x M words a full
    // rewrite moves per proposal.
    const int old_count = nn[i];
    uint32_t fill = count > 0 ? first : 0u;
    if (fill == (uint32_t)i) fill += 1u;
    const int stored = count < M ? count : M;
    const int old_stored = old_count < M ? old_count : M;
    const bool had_pad = old_stored < M;
    const uint32_t old_fill = had_pad ? list[(size_t)i * M + M - 1] : fill;
    __syncwarp();
    const int pad_to = (had_pad && old_fill == fill) ? (old_stored > stored ? old_stored : stored) : M;
    for (int k = stored + lane; k < pad_to; k += 32) list[(size_t)i * M + k] = fill;
    const int lo = count < old_count ? count : old_count, hi = count < old_count ? old_count : count;
    for (int k = lo + lane; k < (hi < M ? hi : M); k += 32) mask[(size_t)i * M + k] = (k < count) ? 1 : 0;
    if (lane == 0) nn[i] = count;
}

static int mcb_launch_move_filtered(chx_ctx* ctx, const McbArgs& m) {
    const chx_mc_barostat_args& a = m.a;
    cudaStream_t st = ctx->stream;
    const int* flip = &m.st->sel;
    k_mcb_propose<<<chx_div_up(a.n, 256), 256, 0, st>>>(a.n, a.cutoff, a.cutoff_plus_skin, m.ncell_cap, m.x0, m.x1,
                                                        m.q0, m.q1, m.st, m.mv);
    CHX_LAUNCHED(ctx);
    k_mcb_filter<<<chx_div_up(a.n, 8), 256, 0, st>>>(a.n, a.M, a.superset_list, a.superset_nn, m.q0, m.q1, m.st, m.mv,
                                                     a.neighbor_list[0], a.neighbor_list[1], a.neighbor_mask[0],
                                                     a.neighbor_mask[1], a.n_neighbors[0], a.n_neighbors[1]);
    CHX_LAUNCHED(ctx);
    k_count_stats<<<chx_div_up(a.n, 256), 256, 0, st>>>(a.n_neighbors[1], a.n, a.M, m.stats, a.n_neighbors[0], flip,
                                                        &m.st->halt);
    CHX_LAUNCHED(ctx);
    int rc = mcb_launch_energy(ctx, m, 1);
    if (rc != CHX_OK) return rc;
    k_mcb_decide<<<1, 1, 0, st>>>(a.M, a.beta, a.pressure, m.st, m.mv, m.acc, m.stats);
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

static int mcb_launch_move(chx_ctx* ctx, const McbArgs& m, int wshift, int warps, size_t smem) {
    const chx_mc_barostat_args& a = m.a;
    cudaStream_t st = ctx->stream;
    const int* flip = &m.st->sel;
    k_mcb_propose<<<chx_div_up(a.n, 256), 256, 0, st>>>(a.n, a.cutoff, a.cutoff_plus_skin, m.ncell_cap, m.x0, m.x1,
                                                        m.q0, m.q1, m.st, m.mv);
    CHX_LAUNCHED(ctx);
    // list build on the proposal: positions x[1 - sel] -> list set 1 - sel
    CHX_CUDA(cudaMemsetAsync(m.count, 0, sizeof(int) * ((size_t)m.ncell_cap + 1), st));
    k_cell_count<<<chx_div_up(a.n, 256), 256, 0, st>>>(m.x1, a.n, &m.mv->P, m.cell_of, m.count, m.x0, flip);
    CHX_LAUNCHED(ctx);
    k_cell_scan<<<1, 1024, 0, st>>>(m.count, m.start, &m.mv->P);
    CHX_LAUNCHED(ctx);
    k_cell_fill<<<chx_div_up(a.n, 256), 256, 0, st>>>(m.x1, m.cell_of, a.n, &m.mv->P, m.start, m.count, m.order,
                                                      m.xs4, m.x0, flip);
    CHX_LAUNCHED(ctx);
    int rc = cell_bm_launch(ctx, m.x1, m.xs4, a.n, &m.mv->P, a.M, wshift, warps, smem, m.cell_of, m.start,
                            a.neighbor_list[1], a.neighbor_mask[1], a.n_neighbors[1], m.x0, a.neighbor_list[0],
                            a.neighbor_mask[0], a.n_neighbors[0], flip);
    if (rc != CHX_OK) return rc;
    k_count_stats<<<chx_div_up(a.n, 256), 256, 0, st>>>(a.n_neighbors[1], a.n, a.M, m.stats, a.n_neighbors[0], flip,
                                                        &m.st->halt);
    CHX_LAUNCHED(ctx);
    rc = mcb_launch_energy(ctx, m, 1);
    if (rc != CHX_OK) return rc;
    k_mcb_decide<<<1, 1, 0, st>>>(a.M, a.beta, a.pressure, m.st, m.mv, m.acc, m.stats);
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

extern "C" {

int chx_mc_barostat_run(chx_ctx* ctx, const chx_mc_barostat_args* args, float* x0, float* x1,
                        chx_mc_baro_state* state_dev, chx_mc_baro_state* state_host, int n_moves) {
    CHX_REQUIRE(ctx && args && x0 && x1 && state_dev && state_host, "NULL argument");
    const chx_mc_barostat_args& a = *args;
    CHX_REQUIRE(a.n > 0 && a.M > 0 && n_moves >= 0, "n and M must be positive, n_moves non-negative");
    for (int k = 0; k < 2; ++k)
        CHX_REQUIRE(a.neighbor_list[k] && a.neighbor_mask[k] && a.n_neighbors[k], "two complete list sets are required");
    int wshift, warps;
    size_t smem;
    CHX_REQUIRE(cell_bm_config(a.n, wshift, warps, smem), "too many particles for the bitmap list builder");
    CellParams P0;
    CHX_REQUIRE(make_cell_params(state_host->box[0], state_host->box[1], state_host->box[2], a.cutoff_plus_skin, 0, P0),
                "the box must hold at least 3 cells of edge cutoff + skin per dimension");
    McbArgs m;
    memset(&m, 0, sizeof(m));
    m.a = a; m.x0 = x0; m.x1 = x1; m.st = state_dev;
    m.ncell_cap = a.ncell_capacity > 0 ? a.ncell_capacity : 2 * P0.ncell;
    // context scratch: [acc 64 B][stats 64 B][pad][McbMove 512 B at +256][q0 | q1 | xs4: n float4 each]
    //                  [cell_of n][count cap+1][start cap+1][order n]
    static_assert(sizeof(McbMove) <= 512, "McbMove must fit its scratch slot");
    const size_t n4 = (size_t)a.n * sizeof(float4);
    const size_t bytes = 768 + 3 * n4 + sizeof(int) * (2 * (size_t)a.n + 2 * ((size_t)m.ncell_cap + 1));
    unsigned char* scratch = (unsigned char*)chx_scratch(ctx, bytes);
    if (!scratch) return CHX_CUDA_ERROR;
    m.acc = (double*)scratch;
    m.stats = (int*)(scratch + 64);
    m.mv = (McbMove*)(scratch + 256);
    m.q0 = (float4*)(scratch + 768);
    m.q1 = m.q0 + a.n;
    m.xs4 = m.q1 + a.n;
    m.cell_of = (int*)(m.xs4 + a.n);
    m.count = m.cell_of + a.n;
    m.start = m.count + m.ncell_cap + 1;
    m.order = m.start + m.ncell_cap + 1;
    cudaStream_t st = ctx->stream;
    state_host->halt = 0;
    state_host->moves_done = 0;
    CHX_CUDA(cudaMemcpyAsync(state_dev, state_host, sizeof(chx_mc_baro_state), cudaMemcpyHostToDevice, st));
    CHX_CUDA(cudaMemsetAsync(scratch, 0, 768, st));
    int rc = CHX_OK;
    if (!state_host->have_u) {
        k_mcb_current<<<1, 1, 0, st>>>(a.cutoff, state_dev, m.mv);
        CHX_LAUNCHED(ctx);
        k_mcb_pack<<<chx_div_up(a.n, 256), 256, 0, st>>>(a.n, x0, x1, m.q0, m.q1, state_dev);
        CHX_LAUNCHED(ctx);
        rc = mcb_launch_energy(ctx, m, 0);
        if (rc != CHX_OK) return rc;
        k_mcb_init<<<1, 1, 0, st>>>(a.beta, a.pressure, state_dev, m.mv, m.acc);
        CHX_LAUNCHED(ctx);
    }
    // superset list: one sweep at a slightly larger radius now, a filter pass per move
    bool filtered = false;
    if (a.superset_list && a.superset_nn && n_moves > 0 && state_host->volume_max_scale < 0.5f) {
        static int use_sup = -1;
        if (use_sup < 0) { const char* e = getenv("CHX_MC_NO_SUPERSET"); use_sup = (e && e[0] == '1') ? 0 : 1; }
        const double delta = pow(1.0 - (double)state_host->volume_max_scale, -(double)n_moves / 3.0) - 1.0 + 1e-4;
        CellParams Ps;
        const float r_sup = (float)((double)a.cutoff_plus_skin * (1.0 + delta));
        if (use_sup && delta <= 0.03 &&
            make_cell_params(state_host->box[0], state_host->box[1], state_host->box[2], r_sup, m.ncell_cap, Ps)) {
            // the superset build reuses the geometry slot of the moves (k_mcb_propose overwrites it later)
            CellParams* Ps_dev = &m.mv->P;
            memcpy(ctx->host_pinned + 8, &Ps, sizeof(Ps));
            CHX_CUDA(cudaMemcpyAsync(Ps_dev, ctx->host_pinned + 8, sizeof(Ps), cudaMemcpyHostToDevice, st));
            const int cur = state_host->sel & 1;
            const float* xc = cur ? x1 : x0;
            CHX_CUDA(cudaMemsetAsync(m.count, 0, sizeof(int) * ((size_t)m.ncell_cap + 1), st));
            k_cell_count<<<chx_div_up(a.n, 256), 256, 0, st>>>(xc, a.n, Ps_dev, m.cell_of, m.count);
            CHX_LAUNCHED(ctx);
            k_cell_scan<<<1, 1024, 0, st>>>(m.count, m.start, Ps_dev);
            CHX_LAUNCHED(ctx);
            k_cell_fill<<<chx_div_up(a.n, 256), 256, 0, st>>>(xc, m.cell_of, a.n, Ps_dev, m.start, m.count, m.order, m.xs4);
            CHX_LAUNCHED(ctx);
            // ids and counts go to the superset arrays; no pad mask is written (both list sets stay valid
            // rows of an earlier build, which k_mcb_filter's incremental row update relies on)
            rc = cell_bm_launch(ctx, xc, m.xs4, a.n, Ps_dev, a.M, wshift, warps, smem, m.cell_of, m.start,
                                a.superset_list, nullptr, a.superset_nn);
            if (rc != CHX_OK) return rc;
            k_count_stats<<<chx_div_up(a.n, 256), 256, 0, st>>>(a.superset_nn, a.n, a.M, m.stats);
            CHX_LAUNCHED(ctx);
            CHX_CUDA(cudaMemcpyAsync(ctx->host_pinned, m.stats, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
            CHX_CUDA(cudaMemsetAsync(m.stats, 0, 2 * sizeof(int), st));
            CHX_CUDA(cudaStreamSynchronize(st));
            filtered = ctx->host_pinned[0] < a.M;      // every superset row fits
        }
    }
    // a move is ~0.1 ms (filtered) / ~0.3 ms (sweep) of device time: plain launches, no graph
    for (int k = 0; k < n_moves && rc == CHX_OK; ++k)
        rc = filtered ? mcb_launch_move_filtered(ctx, m) : mcb_launch_move(ctx, m, wshift, warps, smem);
    if (rc != CHX_OK) return rc;
    CHX_CUDA(cudaMemcpyAsync(state_host, state_dev, sizeof(chx_mc_baro_state), cudaMemcpyDeviceToHost, st));
    CHX_CUDA(cudaStreamSynchronize(st));
    return CHX_OK;
}

}  // extern "C"
