/*
 * chiron_oracle.c -- plain-C restatement of chiron's particle hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
 * load this library; nothing under chiron_b200/ does.  It restates, operation by operation in
 * IEEE fp32 (compile with -ffp-contract=off, no fast-math), what XLA:CPU evaluates for the jitted
 * functions of the reference (choderalab/chiron; file:line relative to that repository):
 *
 *   orc_displacement        chiron/neighbors.py:45-83  (periodic), :116-152 (non periodic)
 *   orc_wrap                chiron/neighbors.py:85-112
 *   orc_nlist_build_rows    chiron/neighbors.py:548-626, 671-729  (O(N^2) half list, padded to M)
 *   orc_nlist_check         chiron/neighbors.py:828-907
 *   orc_lj_nlist            chiron/neighbors.py:731-826 (calculate over the PADDED (N,M) list) +
 *                           chiron/potential.py:193-213,272-279 (masked energy) + the force reverse-mode
 *                           AD produces for it (potential.py:21-24; analytic form :322-326)
 *   orc_random_bits/normal  jax.random legacy threefry2x32 + XLA fp32 ErfInv (third-party, restated
 *                           from the published algorithm; see oracle/jax_random.py for the citation)
 *   orc_langevin_lj         chiron/integrators.py:110-218 (BAOAB loop, wrap, check -> rebuild)
 *
 * Parity pin: checked against the NumPy oracle (itself pinned to the reference's golden vectors,
 * tests/test_oracle_golden.py) in tests/test_oracle_c.py.  Threads: OpenMP over list rows.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- Space ------------------------------------------------------------------------------------ */
/* jnp.mod(t, L): C fmod (exact) + sign fix-up */
static inline float jnp_modf(float t, float L) {
    float rem = fmodf(t, L);
    if (rem != 0.0f && ((rem < 0.0f) != (L < 0.0f))) rem = rem + L;
    return rem;
}

static inline float minimg(float a, float b, float L, float h) {
    float r = a - b;
    float t = r + h;
    /* fast exits that are bit-identical to fmod + fix-up */
    float m;
    if (t >= 0.0f && t < L) m = t;
    else m = jnp_modf(t, L);
    return m - h;
}

static inline float disp(const float* a, const float* b, const float* L, int periodic, float* r) {
    if (periodic) {
        r[0] = minimg(a[0], b[0], L[0], L[0] * 0.5f);
        r[1] = minimg(a[1], b[1], L[1], L[1] * 0.5f);
        r[2] = minimg(a[2], b[2], L[2], L[2] * 0.5f);
    } else {
        r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2];
    }
    float s = (r[0] * r[0] + r[1] * r[1]) + r[2] * r[2];
    return sqrtf(s);
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers: the baseline legs set the thread count explicitly */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_displacement(const float* x1, const float* x2, long n, const float* box, int periodic,
                      float* r_out, float* d_out) {
    for (long i = 0; i < n; ++i) d_out[i] = disp(x1 + 3 * i, x2 + 3 * i, box, periodic, r_out + 3 * i);
}

void orc_wrap(float* x, long n, const float* box) {
    for (long i = 0; i < n; ++i)
        for (int c = 0; c < 3; ++c) {
            float q = x[3 * i + c] / box[c];
            x[3 * i + c] = x[3 * i + c] - floorf(q) * box[c];
        }
}

/* ---- NeighborListNsqrd -------------------------------------------------------------------------- */
/* rows [row0,row1): nn[i-row0] = untruncated count of j>i with d<c; list/mask rows of width M (first M
 * neighbours ascending, padded with the first neighbour, +1 if that equals i; mask[k] = k < n_i).
 * list/mask may be NULL (count only).  Returns the max count over the rows. */
int orc_nlist_build_rows(const float* x, int n, const float* box, int periodic, float cutoff_plus_skin,
                         int M, int row0, int row1, uint32_t* list, int32_t* mask, int32_t* nn) {
    int maxc = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(max : maxc)
    for (int i = row0; i < row1; ++i) {
        float r[3];
        int cnt = 0;
        uint32_t first = 0;
        uint32_t* row = list ? list + (size_t)(i - row0) * M : NULL;
        for (int j = i + 1; j < n; ++j) {
            float d = disp(x + 3 * (size_t)i, x + 3 * (size_t)j, box, periodic, r);
            if (d < cutoff_plus_skin) {
                if (cnt == 0) first = (uint32_t)j;
                if (row && cnt < M) row[cnt] = (uint32_t)j;
                ++cnt;
            }
        }
        uint32_t fill = cnt ? first : 0u;
        if (fill == (uint32_t)i) fill += 1u;
        if (row) {
            for (int k = cnt < M ? cnt : M; k < M; ++k) row[k] = fill;
            int32_t* mrow = mask + (size_t)(i - row0) * M;
            for (int k = 0; k < M; ++k) mrow[k] = k < cnt ? 1 : 0;
        }
        nn[i - row0] = cnt;
        if (cnt > maxc) maxc = cnt;
    }
    return maxc;
}

/* Same arrays as orc_nlist_build_rows(row0=0,row1=n) through a cell grid (edge >= cutoff+skin, 27-cell
 * sweep, exact predicate, rows sorted ascending).  NOT part of the reference algorithm: a set-up
 * accelerator so that bench.py can obtain the reference's list at N = 262,144 without waiting for the
 * O(N^2) build (which is timed separately on a row sample), and a cross-check of the N^2 builder. */
static int cmp_u32(const void* a, const void* b) {
    uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

int orc_nlist_build_cells(const float* x, int n, const float* box, float cutoff_plus_skin, int M,
                          uint32_t* list, int32_t* mask, int32_t* nn) {
    int nc[3];
    for (int c = 0; c < 3; ++c) {
        nc[c] = (int)floorf(box[c] / (cutoff_plus_skin * 1.0001f));
        if (nc[c] < 3) return orc_nlist_build_rows(x, n, box, 1, cutoff_plus_skin, M, 0, n, list, mask, nn);
        if (nc[c] > 256) nc[c] = 256;
    }
    const int ncell = nc[0] * nc[1] * nc[2];
    int* cell_of = (int*)malloc(sizeof(int) * (size_t)n);
    int* start = (int*)calloc((size_t)ncell + 1, sizeof(int));
    int* order = (int*)malloc(sizeof(int) * (size_t)n);
    for (int i = 0; i < n; ++i) {
        int cc[3];
        for (int c = 0; c < 3; ++c) {
            float w = x[3 * (size_t)i + c] - floorf(x[3 * (size_t)i + c] / box[c]) * box[c];
            int k = (int)floorf(w / box[c] * nc[c]);
            cc[c] = k < 0 ? 0 : (k >= nc[c] ? nc[c] - 1 : k);
        }
        cell_of[i] = (cc[0] * nc[1] + cc[1]) * nc[2] + cc[2];
        start[cell_of[i] + 1]++;
    }
    for (int c = 0; c < ncell; ++c) start[c + 1] += start[c];
    int* cur = (int*)malloc(sizeof(int) * (size_t)ncell);
    memcpy(cur, start, sizeof(int) * (size_t)ncell);
    for (int i = 0; i < n; ++i) order[cur[cell_of[i]]++] = i;
    free(cur);
    int maxc = 0;
#pragma omp parallel reduction(max : maxc)
    {
        int capn = 1024;
        uint32_t* buf = (uint32_t*)malloc(sizeof(uint32_t) * capn);
#pragma omp for schedule(dynamic, 64)
        for (int i = 0; i < n; ++i) {
            int cnt = 0;
            const int ci = cell_of[i];
            const int cz = ci % nc[2], cy = (ci / nc[2]) % nc[1], cx = ci / (nc[2] * nc[1]);
            for (int dx = -1; dx <= 1; ++dx)
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dz = -1; dz <= 1; ++dz) {
                        const int ox = (cx + dx + nc[0]) % nc[0], oy = (cy + dy + nc[1]) % nc[1],
                                  oz = (cz + dz + nc[2]) % nc[2];
                        const int c = (ox * nc[1] + oy) * nc[2] + oz;
                        for (int k = start[c]; k < start[c + 1]; ++k) {
                            const int j = order[k];
                            if (j <= i) continue;
                            float r[3];
                            if (disp(x + 3 * (size_t)i, x + 3 * (size_t)j, box, 1, r) < cutoff_plus_skin) {
                                if (cnt == capn) { capn *= 2; buf = (uint32_t*)realloc(buf, sizeof(uint32_t) * capn); }
                                buf[cnt++] = (uint32_t)j;
                            }
                        }
                    }
            qsort(buf, cnt, sizeof(uint32_t), cmp_u32);
            uint32_t fill = cnt ? buf[0] : 0u;
            if (fill == (uint32_t)i) fill += 1u;
            if (list) {
                uint32_t* row = list + (size_t)i * M;
                int32_t* mrow = mask + (size_t)i * M;
                for (int k = 0; k < M; ++k) { row[k] = k < cnt ? buf[k] : fill; mrow[k] = k < cnt ? 1 : 0; }
            }
            nn[i] = cnt;
            if (cnt > maxc) maxc = cnt;
        }
        free(buf);
    }
    free(cell_of); free(start); free(order);
    return maxc;
}

/* wall time of the reference O(N^2) row loop on the rows row0, row0+stride, ... (counts only) */
double orc_nlist_time_rows(const float* x, int n, const float* box, float cutoff_plus_skin, int row0,
                           int stride, long long* pair_tests) {
    double t0 = 0.0;
#ifdef _OPENMP
    t0 = omp_get_wtime();
#endif
    long long tests = 0, hits = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : tests, hits)
    for (int i = row0; i < n; i += stride) {
        float r[3];
        for (int j = i + 1; j < n; ++j)
            hits += disp(x + 3 * (size_t)i, x + 3 * (size_t)j, box, 1, r) < cutoff_plus_skin;
        tests += n - i - 1;
    }
    if (pair_tests) *pair_tests = tests + (hits & 0);   /* keep `hits` live */
#ifdef _OPENMP
    return omp_get_wtime() - t0;
#else
    return 0.0;
#endif
}

int orc_nlist_check(const float* x, const float* ref, int n, const float* box, int periodic, float half_skin) {
    int any = 0;
#pragma omp parallel for reduction(| : any)
    for (int i = 0; i < n; ++i) {
        float r[3];
        if (disp(x + 3 * (size_t)i, ref + 3 * (size_t)i, box, periodic, r) >= half_skin) any |= 1;
    }
    return any;
}

/* ---- LJ over the padded list ---------------------------------------------------------------------- */
/* Energy (double sum of fp32 pair terms) over ALL M padded slots of rows [row0,row1) -- pads are
 * evaluated and masked like the reference does -- and, if F != NULL, the force scattered to i and j.
 * F is (n,3) fp32 and is OVERWRITTEN.  n_int (nullable) = number of masked-in pairs. */
double orc_lj_nlist(const float* x, int n, const float* box, int periodic, float sigma, float epsilon,
                    float cutoff, int M, int row0, int row1, const uint32_t* list, const int32_t* mask,
                    float* F, long long* n_int) {
    double e_tot = 0.0;
    long long pairs = 0;
    int nt = orc_num_threads();
    float* Fp = NULL;
    if (F) Fp = (float*)calloc((size_t)nt * n * 3, sizeof(float));
#pragma omp parallel reduction(+ : e_tot, pairs)
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num();
#else
        int tid = 0;
#endif
        float* Ft = Fp ? Fp + (size_t)tid * n * 3 : NULL;
#pragma omp for schedule(dynamic, 64)
        for (int i = row0; i < row1; ++i) {
            const uint32_t* row = list + (size_t)(i - row0) * M;
            const int32_t* mrow = mask + (size_t)(i - row0) * M;
            float fi[3] = {0.f, 0.f, 0.f};
            for (int k = 0; k < M; ++k) {
                float r[3];
                uint32_t j = row[k];
                float d = disp(x + 3 * (size_t)i, x + 3 * (size_t)j, box, periodic, r);
                int in = (d < cutoff) && mrow[k];
                if (!in) continue;     /* reference multiplies by the 0/1 mask */
                float q = sigma / d;
                float q2 = q * q;
                float q6 = q2 * q2 * q2;
                float q12 = q6 * q6;
                e_tot += (double)((4.0f * epsilon) * (q12 - q6));
                ++pairs;
                if (Ft) {
                    float f = 24.0f * (epsilon / (d * d)) * (2.0f * q12 - q6);
                    for (int c = 0; c < 3; ++c) {
                        float fc = f * r[c];
                        fi[c] += fc;
                        Ft[3 * (size_t)j + c] -= fc;
                    }
                }
            }
            if (Ft) for (int c = 0; c < 3; ++c) Ft[3 * (size_t)i + c] += fi[c];
        }
    }
    if (F) {
#pragma omp parallel for
        for (long k = 0; k < (long)n * 3; ++k) {
            float s = 0.f;
            for (int t = 0; t < nt; ++t) s += Fp[(size_t)t * n * 3 + k];
            F[k] = s;
        }
        free(Fp);
    }
    if (n_int) *n_int = pairs;
    return e_tot;
}

/* ---- jax.random, legacy threefry ------------------------------------------------------------------- */
static inline uint32_t rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

static void threefry2x32(uint32_t k0, uint32_t k1, uint32_t* px0, uint32_t* px1) {
    static const int ROT[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
    uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
    uint32_t x0 = *px0 + ks[0], x1 = *px1 + ks[1];
    for (int g = 1; g <= 5; ++g) {
        for (int q = 0; q < 4; ++q) {
            x0 += x1;
            x1 = rotl(x1, ROT[(g - 1) & 1][q]);
            x1 ^= x0;
        }
        x0 += ks[g % 3];
        x1 += ks[(g + 1) % 3] + (uint32_t)g;
    }
    *px0 = x0; *px1 = x1;
}

void orc_random_bits(const uint32_t* key, long n, uint32_t* out) {
    long half = (n + 1) / 2;
#pragma omp parallel for if (n > 4096)
    for (long b = 0; b < half; ++b) {
        uint32_t x0 = (uint32_t)b;
        long c1 = b + half;
        uint32_t x1 = c1 < n ? (uint32_t)c1 : 0u;
        threefry2x32(key[0], key[1], &x0, &x1);
        out[b] = x0;
        if (c1 < n) out[c1] = x1;
    }
}

void orc_split(const uint32_t* key, uint32_t* out4) { orc_random_bits(key, 4, out4); }

static inline float erfinv_xla(float x) {
    float w = -log1pf(-(x * x));
    float p;
    if (w < 5.0f) {
        static const float c[9] = {2.81022636e-08f, 3.43273939e-07f, -3.5233877e-06f, -4.39150654e-06f,
                                   0.00021858087f, -0.00125372503f, -0.00417768164f, 0.246640727f, 1.50140941f};
        w = w - 2.5f;
        p = c[0];
        for (int k = 1; k < 9; ++k) p = c[k] + p * w;
    } else {
        static const float c[9] = {-0.000200214257f, 0.000100950558f, 0.00134934322f, -0.00367342844f,
                                   0.00573950773f, -0.0076224613f, 0.00943887047f, 1.00167406f, 2.83297682f};
        w = sqrtf(w) - 3.0f;
        p = c[0];
        for (int k = 1; k < 9; ++k) p = c[k] + p * w;
    }
    if (fabsf(x) == 1.0f) return x * INFINITY;
    return p * x;
}

static inline float bits_to_normal(uint32_t bits) {
    union { uint32_t u; float f; } cv;
    cv.u = (bits >> 9) | 0x3F800000u;
    float f = cv.f - 1.0f;
    const float lo = -0.99999994f; /* nextafter(-1, 0) */
    float u = f * (1.0f - lo) + lo;
    if (u < lo) u = lo;
    return 1.41421356f * erfinv_xla(u);
}

void orc_normal(const uint32_t* key, long n, float* out) {
    uint32_t* bits = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n > 0 ? n : 1));
    orc_random_bits(key, n, bits);
#pragma omp parallel for if (n > 4096)
    for (long i = 0; i < n; ++i) out[i] = bits_to_normal(bits[i]);
    free(bits);
}

/* ---- LangevinIntegrator.run for LJPotential + NeighborListNsqrd --------------------------------------- */
typedef struct {
    int n_builds;          /* list builds including the initial one */
    int M;                 /* final n_max_neighbors */
    double t_build_s;      /* wall time spent in builds */
    double t_steps_s;      /* wall time spent in the step loop excluding builds */
    double energy;         /* potential energy after the last step */
    long long p_cand, p_int;
} orc_langevin_stats;

static double now_s(void) {
#ifdef _OPENMP
    return omp_get_wtime();
#else
    return 0.0;
#endif
}

typedef struct { uint32_t* list; int32_t* mask; int32_t* nn; int M; float* ref; } nlist_t;

static int g_build_mode = 0;   /* 0 = reference O(N^2) build, 1 = cell-grid set-up accelerator */
void orc_set_build_mode(int mode) { g_build_mode = mode; }

static void nlist_build(nlist_t* nl, const float* x, int n, const float* box, float cps) {
    /* growth loop of neighbors.py:709-729: while any(n == M): M = max(n) + 10 ; rebuild */
    for (;;) {
        nl->list = (uint32_t*)realloc(nl->list, sizeof(uint32_t) * (size_t)n * nl->M);
        nl->mask = (int32_t*)realloc(nl->mask, sizeof(int32_t) * (size_t)n * nl->M);
        int maxc = g_build_mode ? orc_nlist_build_cells(x, n, box, cps, nl->M, nl->list, nl->mask, nl->nn)
                                : orc_nlist_build_rows(x, n, box, 1, cps, nl->M, 0, n, nl->list, nl->mask, nl->nn);
        int hit = 0;
        for (int i = 0; i < n; ++i) hit |= (nl->nn[i] == nl->M);
        if (!hit) break;
        nl->M = maxc + 10;
    }
    memcpy(nl->ref, x, sizeof(float) * 3 * (size_t)n);
}

/* x, v (n,3) updated in place; key[2] = SamplerState loop key, updated to the key after the loop.
 * Returns 0. */
int orc_langevin_lj(float* x, float* v, const float* mass, int n, const float* box, float sigma,
                    float epsilon, float cutoff, float skin, int n_max_neighbors, float kT, float dt,
                    float gamma, uint32_t* key, int nsteps, orc_langevin_stats* st) {
    nlist_t nl;
    nl.M = n_max_neighbors; nl.list = NULL; nl.mask = NULL;
    nl.nn = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    nl.ref = (float*)malloc(sizeof(float) * 3 * (size_t)n);
    float* F = (float*)malloc(sizeof(float) * 3 * (size_t)n);
    float* xi = (float*)malloc(sizeof(float) * 3 * (size_t)n);
    const float cps = (float)((double)cutoff + (double)skin);
    const float half_skin = (float)((double)skin / 2.0);
    const float a = (float)exp((double)(float)(-gamma * dt));
    const float b = sqrtf(1.0f - (float)exp((double)(float)(-2.0f * gamma * dt)));
    const float h = dt * 0.5f;
    memset(st, 0, sizeof(*st));
    double t0 = now_s();
    nlist_build(&nl, x, n, box, cps);
    st->n_builds = 1;
    st->t_build_s += now_s() - t0;
    double t_loop = now_s();
    double t_b = 0.0;
    long long n_int = 0;
    st->energy = orc_lj_nlist(x, n, box, 1, sigma, epsilon, cutoff, nl.M, 0, n, nl.list, nl.mask, F, &n_int);
    for (int s = 0; s < nsteps; ++s) {
        uint32_t k4[4];
        orc_split(key, k4);
        key[0] = k4[0]; key[1] = k4[1];
        orc_normal(k4 + 2, 3L * n, xi);
#pragma omp parallel for
        for (int i = 0; i < n; ++i) {
            const float m = mass[i];
            const float bs = b * sqrtf(kT / m);
            for (int c = 0; c < 3; ++c) {
                size_t o = 3 * (size_t)i + c;
                float vv = v[o] + (h * F[o]) / m;
                float xx = x[o] + h * vv;
                vv = a * vv + bs * xi[o];
                xx = xx + h * vv;
                xx = xx - floorf(xx / box[c]) * box[c];
                v[o] = vv; x[o] = xx;
            }
        }
        if (orc_nlist_check(x, nl.ref, n, box, 1, half_skin)) {
            double tb = now_s();
            nlist_build(&nl, x, n, box, cps);
            st->n_builds++;
            t_b += now_s() - tb;
        }
        st->energy = orc_lj_nlist(x, n, box, 1, sigma, epsilon, cutoff, nl.M, 0, n, nl.list, nl.mask, F, &n_int);
#pragma omp parallel for
        for (int i = 0; i < n; ++i)
            for (int c = 0; c < 3; ++c) {
                size_t o = 3 * (size_t)i + c;
                v[o] = v[o] + (h * F[o]) / mass[i];
            }
    }
    st->t_steps_s = now_s() - t_loop - t_b;
    st->t_build_s += t_b;
    st->M = nl.M;
    st->p_int = n_int;
    long long pc = 0;
    for (int i = 0; i < n; ++i) pc += nl.nn[i];
    st->p_cand = pc;
    free(nl.list); free(nl.mask); free(nl.nn); free(nl.ref); free(F); free(xi);
    return 0;
}
