// Context management, error strings, Space kernels and jax.random kernels.
#include <stdarg.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

void chx_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void* chx_scratch(chx_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->scratch_bytes) return ctx->scratch;
    if (ctx->scratch) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(ctx->scratch);
        ctx->scratch = nullptr;
        ctx->scratch_bytes = 0;
    }
    size_t want = bytes + (bytes >> 2) + 4096;
    if (cudaMalloc(&ctx->scratch, want) != cudaSuccess) {
        chx_set_error("scratch allocation of %zu bytes failed", want);
        return nullptr;
    }
    ctx->scratch_bytes = want;
    return ctx->scratch;
}

extern "C" {

int chx_version(void) { return 100; }

const char* chx_last_error_string(void) { return g_err; }

int chx_context_create(int device, void* cuda_stream, chx_ctx** out) {
    CHX_REQUIRE(out != nullptr, "out is NULL");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        chx_set_error("no CUDA device available (%s); libchiron_b200 has no CPU fallback",
                      cudaGetErrorString(e));
        return CHX_CUDA_ERROR;
    }
    CHX_REQUIRE(device >= 0 && device < count, "device index out of range");
    CHX_CUDA(cudaSetDevice(device));
    chx_ctx* ctx = new chx_ctx();
    ctx->device = device;
    ctx->stream = (cudaStream_t)cuda_stream;
    cudaDeviceProp prop;
    CHX_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    CHX_CUDA(cudaMallocHost((void**)&ctx->host_pinned, 64 * sizeof(int)));
    *out = ctx;
    return CHX_OK;
}

int chx_context_set_stream(chx_ctx* ctx, void* cuda_stream) {
    CHX_REQUIRE(ctx != nullptr, "ctx is NULL");
    ctx->stream = (cudaStream_t)cuda_stream;
    return CHX_OK;
}

int chx_context_destroy(chx_ctx* ctx) {
    if (!ctx) return CHX_OK;
    cudaSetDevice(ctx->device);
    chx_mc_forget_context(ctx);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->host_pinned) cudaFreeHost(ctx->host_pinned);
    delete ctx;
    return CHX_OK;
}

int chx_synchronize(chx_ctx* ctx) {
    CHX_REQUIRE(ctx != nullptr, "ctx is NULL");
    CHX_CUDA(cudaStreamSynchronize(ctx->stream));
    return CHX_OK;
}

long long chx_launch_count(chx_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Space
// ---------------------------------------------------------------------------------------------
template <bool PERIODIC>
__global__ void k_displacement(const float* __restrict__ x1, const float* __restrict__ x2,
                               long long n, Box box, float* __restrict__ r_out,
                               float* __restrict__ d_out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float rx, ry, rz, d;
    ref_displacement<PERIODIC>(x1[3 * i], x1[3 * i + 1], x1[3 * i + 2], x2[3 * i], x2[3 * i + 1],
                               x2[3 * i + 2], box, rx, ry, rz, d);
    r_out[3 * i] = rx; r_out[3 * i + 1] = ry; r_out[3 * i + 2] = rz;
    d_out[i] = d;
}

__global__ void k_wrap(const float* __restrict__ x, long long n, Box box, float* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[3 * i] = ref_wrap(x[3 * i], box.lx);
    out[3 * i + 1] = ref_wrap(x[3 * i + 1], box.ly);
    out[3 * i + 2] = ref_wrap(x[3 * i + 2], box.lz);
}

// ---------------------------------------------------------------------------------------------
// jax.random
// ---------------------------------------------------------------------------------------------
// One thread per threefry block: both outputs are used (elements b and b + half).
template <int MODE>  // 0 = normal, 1 = uniform
__global__ void k_random(uint32_t k0, uint32_t k1, unsigned long long n, float lo, float hi,
                         float* __restrict__ out) {
    const unsigned long long half = (n + 1ull) >> 1;
    unsigned long long b = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (b >= half) return;
    uint32_t x0 = (uint32_t)b;
    unsigned long long c1 = b + half;
    uint32_t x1 = c1 < n ? (uint32_t)c1 : 0u;
    threefry2x32(k0, k1, x0, x1);
    if (MODE == 0) {
        out[b] = normal_from_bits(x0);
        if (c1 < n) out[c1] = normal_from_bits(x1);
    } else {
        out[b] = uniform_from_bits(x0, lo, hi);
        if (c1 < n) out[c1] = uniform_from_bits(x1, lo, hi);
    }
}

__global__ void k_scale(const float* __restrict__ x, long long n, float s, float* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = __fmul_rn(x[i], s);
}

extern "C" {

int chx_displacement(chx_ctx* ctx, const float* x1, const float* x2, long long n, float lx,
                     float ly, float lz, int periodic, float* r_out, float* d_out) {
    CHX_REQUIRE(ctx && x1 && x2 && r_out && d_out, "NULL argument");
    if (n <= 0) return CHX_OK;
    Box box = make_box(lx, ly, lz);
    int blocks = chx_div_up(n, 256);
    if (periodic) k_displacement<true><<<blocks, 256, 0, ctx->stream>>>(x1, x2, n, box, r_out, d_out);
    else k_displacement<false><<<blocks, 256, 0, ctx->stream>>>(x1, x2, n, box, r_out, d_out);
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

int chx_wrap(chx_ctx* ctx, const float* x, long long n, float lx, float ly, float lz, int periodic,
             float* out) {
    CHX_REQUIRE(ctx && x && out, "NULL argument");
    if (n <= 0) return CHX_OK;
    if (!periodic) {
        if (out != x)
            CHX_CUDA(cudaMemcpyAsync(out, x, sizeof(float) * 3 * n, cudaMemcpyDeviceToDevice,
                                     ctx->stream));
        return CHX_OK;
    }
    k_wrap<<<chx_div_up(n, 256), 256, 0, ctx->stream>>>(x, n, make_box(lx, ly, lz), out);
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

int chx_threefry_split_host(const uint32_t key_host[2], uint32_t out_host[4]) {
    CHX_REQUIRE(key_host && out_host, "NULL argument");
    threefry_split(key_host[0], key_host[1], out_host[0], out_host[1], out_host[2], out_host[3]);
    return CHX_OK;
}

int chx_threefry_split_host_n(const uint32_t* keys_host, int n, uint32_t* out_host) {
    CHX_REQUIRE(keys_host && out_host && n >= 0, "bad argument");
    for (int k = 0; k < n; ++k)
        threefry_split(keys_host[2 * k], keys_host[2 * k + 1], out_host[4 * k], out_host[4 * k + 1], out_host[4 * k + 2],
                       out_host[4 * k + 3]);
    return CHX_OK;
}

int chx_random_bits_host(const uint32_t key_host[2], long long n, uint32_t* out_host) {
    CHX_REQUIRE(key_host && out_host && n >= 0, "bad argument");
    for (long long e = 0; e < n; ++e)
        out_host[e] = random_bits_elem(key_host[0], key_host[1], (unsigned long long)e,
                                       (unsigned long long)n);
    return CHX_OK;
}

int chx_random_normal(chx_ctx* ctx, uint32_t key0, uint32_t key1, long long n, float* out) {
    CHX_REQUIRE(ctx && out, "NULL argument");
    if (n <= 0) return CHX_OK;
    long long half = (n + 1) / 2;
    k_random<0><<<chx_div_up(half, 256), 256, 0, ctx->stream>>>(key0, key1, (unsigned long long)n,
                                                               0.f, 1.f, out);
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

int chx_random_uniform(chx_ctx* ctx, uint32_t key0, uint32_t key1, long long n, float lo, float hi,
                       float* out) {
    CHX_REQUIRE(ctx && out, "NULL argument");
    if (n <= 0) return CHX_OK;
    long long half = (n + 1) / 2;
    k_random<1><<<chx_div_up(half, 256), 256, 0, ctx->stream>>>(key0, key1, (unsigned long long)n,
                                                               lo, hi, out);
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

int chx_scale(chx_ctx* ctx, const float* x, long long n_elems, float s, float* x_out) {
    CHX_REQUIRE(ctx && x && x_out, "NULL argument");
    if (n_elems <= 0) return CHX_OK;
    k_scale<<<chx_div_up(n_elems, 256), 256, 0, ctx->stream>>>(x, n_elems, s, x_out);
    CHX_LAUNCHED(ctx);
    return CHX_OK;
}

}  // extern "C"
