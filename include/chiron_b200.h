/*
 * chiron_b200.h -- C ABI of libchiron_b200.so, the sm_100a implementation of chiron's particle
 * hot path (pair finding, LJ energy/force, BAOAB Langevin update, Metropolis move energies).
 *
 * chiron (choderalab/chiron) is pure Python on JAX and has no FFI of its own; the boundary this
 * library replaces is the set of jitted methods of its duck-typed classes.  Every entry point
 * below names the reference method it stands in for (file:line relative to the chiron repo).
 * The Python classes in chiron_b200/ (same names/signatures as chiron's) are the only callers;
 * INTEGRATION.md shows the ctypes stub a chiron maintainer would add.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes.  All array arguments are DEVICE pointers unless the
 *    name ends in _host.  Positions / velocities / forces are float32 (N,3) row-major.
 *    Particle ids are uint32 like the reference (neighbors.py:619,681).
 *  - Orthorhombic boxes only, like the reference (neighbors.py:74-76 uses the diagonal):
 *    box is passed as three floats lx, ly, lz; `periodic` = 0 selects
 *    OrthogonalNonPeriodicSpace semantics (neighbors.py:115-175).
 *  - A chx_ctx binds one device and one CUDA stream.  Every call enqueues work on that stream
 *    and returns without synchronising, except where a host scalar is returned (documented
 *    per function).  A context is not thread-safe; distinct contexts are independent.
 *  - Return value: 0 = CHX_OK, negative = error (see enum); chx_last_error_string() has detail.
 *    No C++ exception crosses the boundary.
 *  - Predicates (d < cutoff, d < cutoff+skin, d >= skin/2) are evaluated with the reference's
 *    exact fp32 operation order, so pair sets are bit-identical.
 */
#ifndef CHIRON_B200_H
#define CHIRON_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct chx_ctx chx_ctx;
typedef struct chx_ljmd chx_ljmd;   /* fused LJ Langevin engine, see below */

enum {
    CHX_OK = 0,
    CHX_BAD_ARG = -1,
    CHX_NEIGHBOR_OVERFLOW = -2,
    CHX_CELL_OVERFLOW = -3,
    CHX_NAN_ENERGY = -4,
    CHX_CUDA_ERROR = -5
};

int chx_version(void);
const char* chx_last_error_string(void);

/* One context per (device, stream).  `cuda_stream` is a cudaStream_t (NULL = legacy default). */
int chx_context_create(int device, void* cuda_stream, chx_ctx** out);
int chx_context_set_stream(chx_ctx* ctx, void* cuda_stream);
int chx_context_destroy(chx_ctx* ctx);
int chx_synchronize(chx_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches). */
long long chx_launch_count(chx_ctx* ctx);

/* ---- Space (chiron/neighbors.py:45-112 periodic, :116-175 non-periodic) ------------------------ */
/* Space.displacement for n independent point pairs: r_out (n,3), d_out (n). */
int chx_displacement(chx_ctx* ctx, const float* x1, const float* x2, long long n,
                     float lx, float ly, float lz, int periodic, float* r_out, float* d_out);
/* Space.wrap: out = x - floor(x/L)*L (identity when periodic == 0); out may alias x. */
int chx_wrap(chx_ctx* ctx, const float* x, long long n, float lx, float ly, float lz,
             int periodic, float* out);

/* ---- NeighborListNsqrd (chiron/neighbors.py:446-907) -------------------------------------------- */
/* build (neighbors.py:548-729): half Verlet list i<j, d < cutoff_plus_skin, ascending ids, first M
 * entries, padded with the first neighbour (0 if none, +1 if that equals i); mask[k] = k < n_i.
 * n_neighbors holds the UNTRUNCATED counts.  Synchronises: *max_count_host = max_i n_i and
 * *count_eq_M_host = number of rows with n_i == M (the reference's growth trigger, :709).
 * _nsq is the O(N^2) restatement; _cell is the O(N) counting-sort cell list producing the
 * identical arrays (periodic only; falls back to _nsq when the box has < 3 cells per edge). */
int chx_nlist_build_nsq(chx_ctx* ctx, const float* x, int n, float lx, float ly, float lz,
                        int periodic, float cutoff_plus_skin, int M, uint32_t* neighbor_list,
                        int32_t* neighbor_mask, int32_t* n_neighbors, int* max_count_host,
                        int* count_eq_M_host);
int chx_nlist_build_cell(chx_ctx* ctx, const float* x, int n, float lx, float ly, float lz,
                         int periodic, float cutoff_plus_skin, int M, uint32_t* neighbor_list,
                         int32_t* neighbor_mask, int32_t* n_neighbors, int* max_count_host,
                         int* count_eq_M_host);
/* calculate (neighbors.py:731-826): n_out (N), mask_out (N,M) = (d<cutoff)&pad, dist (N,M), r_ij (N,M,3). */
int chx_nlist_calculate(chx_ctx* ctx, const float* x, int n, float lx, float ly, float lz,
                        int periodic, float cutoff, int M, const uint32_t* neighbor_list,
                        const int32_t* neighbor_mask, int32_t* n_out, int32_t* mask_out,
                        float* dist_out, float* rij_out);
/* check (neighbors.py:828-907): *flag_dev = any_i ||minimg(x_i - ref_i)|| >= skin/2.  Asynchronous;
 * flag_dev is a device int the caller reads (or hands to the next kernel). */
int chx_nlist_check(chx_ctx* ctx, const float* x, const float* ref_x, int n, float lx, float ly,
                    float lz, int periodic, float half_skin, int32_t* flag_dev);

/* ---- PairListNsqrd (chiron/neighbors.py:1018-1289) ---------------------------------------------- */
/* build: all_pairs (N,N-1) = ids j != i ascending; reduction_mask (N,N-1) uint8 = i < j. */
int chx_pairlist_build(chx_ctx* ctx, int n, uint32_t* all_pairs, uint8_t* reduction_mask);
/* calculate: cutoff < 0 means "no cutoff" (_calc_distance_per_particle_no_cutoff, :1163-1216). */
int chx_pairlist_calculate(chx_ctx* ctx, const float* x, int n, float lx, float ly, float lz,
                           int periodic, float cutoff, int32_t* n_out, int32_t* mask_out,
                           float* dist_out, float* rij_out);

/* ---- Potentials (chiron/potential.py) ------------------------------------------------------------- */
/* LJPotential.compute_energy / compute_force over a built NeighborListNsqrd
 * (potential.py:193-213,263-300; force = -grad = potential.py:322-326 scattered to i and j).
 * energy_dev (double, device, nullable) receives sum over listed pairs with d < cutoff of
 * 4 eps ((s/d)^12 - (s/d)^6); force (N,3, nullable) is overwritten. */
int chx_lj_nlist_energy_force(chx_ctx* ctx, const float* x, int n, float lx, float ly, float lz,
                              int periodic, const uint32_t* neighbor_list,
                              const int32_t* n_neighbors, int M, float sigma, float epsilon,
                              float cutoff, double* energy_dev, float* force);
/* All pairs i<j (PairListNsqrd path neighbors.py:1106-1216 + potential.py:272-279, and the
 * nbr_list=None non-periodic path potential.py:26-63,235-258).  cutoff < 0 = no cutoff. */
int chx_lj_allpairs_energy_force(chx_ctx* ctx, const float* x, int n, float lx, float ly, float lz,
                                 int periodic, float sigma, float epsilon, float cutoff,
                                 double* energy_dev, float* force);
/* HarmonicOscillatorPotential (potential.py:413-418): 0.5 k sum (x - x0)^2 + U0; x0 is (n0,3) with
 * n0 == n or n0 == 1 (broadcast); force = -k (x - x0). */
int chx_ho_energy_force(chx_ctx* ctx, const float* x, int n, const float* x0, int n0, float k,
                        float U0, double* energy_dev, float* force);
/* Single-pass Metropolis delta energy for a displaced subset (new fast path; semantics of
 * mcmc.py:733-777 restricted to atom_subset): for each moved particle m in `moved` (n_moved ids),
 * sum over all j of u(x_new_m, x_j') - u(x_old_m, x_j), pairs inside the subset counted once.
 * delta_dev (double) receives U(new) - U(old). */
int chx_lj_subset_delta_energy(chx_ctx* ctx, const float* x_old, const float* x_new, int n,
                               const uint32_t* moved, int n_moved, float lx, float ly, float lz,
                               int periodic, float sigma, float epsilon, float cutoff,
                               double* delta_dev);

/* ---- jax.random, legacy threefry stream -------------------------------------- */
/* random.split(key, 2) on the host: out_host[0..1] = carried key, out_host[2..3] = subkey. */
int chx_threefry_split_host(const uint32_t key_host[2], uint32_t out_host[4]);
/* the same for n keys: keys_host (n,2) -> out_host (n,4) = [carried key, subkey] per row */
int chx_threefry_split_host_n(const uint32_t* keys_host, int n, uint32_t* out_host);
/* random_bits(key, n) on the host (small n: scalar uniforms for the Metropolis test, mcmc.py:544). */
int chx_random_bits_host(const uint32_t key_host[2], long long n, uint32_t* out_host);
/* random.normal(key, (n,)) / random.uniform(key, (n,), lo, hi), bit-compatible layout. */
int chx_random_normal(chx_ctx* ctx, uint32_t key0, uint32_t key1, long long n, float* out);
int chx_random_uniform(chx_ctx* ctx, uint32_t key0, uint32_t key1, long long n, float lo,
                       float hi, float* out);

/* ---- BAOAB building blocks (chiron/integrators.py:174-195) for arbitrary potentials --------------- */
/* One fused kernel for everything between two force evaluations of step t:
 *   v += h F/m ; x += h v ; v = a v + b sqrt(kT/m) xi ; x += h v ; [wrap] ; [check vs ref_x]
 * with h = dt/2 and xi = random.normal(subkey,(n,3)) generated in-kernel on the reference's
 * stream.  `first_half_kick` = 0 skips the leading B (used to fuse the trailing B of step t-1:
 * pass `trailing_kick` = 1 to apply v += h F/m before anything else).  flag_dev (nullable) is
 * OR-ed with the rebuild condition of neighbors.py:864-868. */
int chx_baoab_update(chx_ctx* ctx, float* x, float* v, const float* force, const float* mass,
                     int n, float half_dt, float a, float b, float kT, uint32_t subkey0,
                     uint32_t subkey1, int trailing_kick, float lx, float ly, float lz,
                     int wrap_periodic, const float* ref_x, float half_skin, int32_t* flag_dev);
/* v += h F/m (the trailing B of the last step). */
int chx_kick(chx_ctx* ctx, float* v, const float* force, const float* mass, int n, float half_dt);
/* Maxwell-Boltzmann velocities (utils.py:116-144): v = sqrt(kT/m) * normal(key,(n,3)). */
int chx_init_velocities(chx_ctx* ctx, float* v, const float* mass, int n, float kT, uint32_t key0,
                        uint32_t key1);

/* ---- Monte Carlo proposals (chiron/mcmc.py:733-752, 956-983) -------------------------------------- */
/* x_out = wrap(x + sigma * normal(key,(n,3)) * subset_mask[:,None]) ; subset_mask nullable (all). */
int chx_mc_displace(chx_ctx* ctx, const float* x, int n, uint32_t key0, uint32_t key1,
                    float sigma, const float* subset_mask, float lx, float ly, float lz,
                    int wrap_periodic, float* x_out);
/* x_out = x * s (barostat coordinate scaling). */
int chx_scale(chx_ctx* ctx, const float* x, long long n_elems, float s, float* x_out);

/* ---- Device-resident Metropolis loop (chiron/mcmc.py:243-306 MCMove.update, :357-463 _step,
 *      :531-548 _accept_or_reject, :733-787 MonteCarloDisplacementMove._propose) --------------------- */
/* n_moves Monte Carlo displacement steps without a host round trip per step.  Per move: split the
 * state's key (states.py:150-154), x' = wrap(x + sigma * normal(subkey,(n,3)) * subset_mask[:,None]),
 * NeighborListNsqrd.check(x') (neighbors.py:828-907), reduced potential u' = beta (U(x') + pv), split
 * again, accept iff -(u - u') <= 0 or uniform(subkey) < exp(u - u') (fp32).  The two position
 * buffers alternate roles: x[state.sel] is the current configuration.
 * Potential kinds: what `U` is evaluated with. */
enum {
    CHX_MC_LJ_NLIST = 0,        /* LJPotential over a built NeighborListNsqrd (potential.py:193-279) */
    CHX_MC_LJ_ALLPAIRS = 1,     /* LJPotential over all pairs i<j (PairListNsqrd / nbr_list=None); cutoff<0: none */
    CHX_MC_HO = 2,              /* HarmonicOscillatorPotential (potential.py:413-418) */
    CHX_MC_IDEAL = 3,           /* IdealGasPotential: U = 0 (potential.py:93-127) */
    CHX_MC_LJ_SUBSET_DELTA = 4  /* LJ, only the rows of the moved subset: u' = u + beta dU (new fast path) */
};
typedef struct {
    int n;                         /* particles */
    int potential;                 /* CHX_MC_* */
    int periodic;                  /* 1: wrap proposals and use minimum images (OrthogonalPeriodicSpace) */
    float lx, ly, lz;
    float sigma, epsilon, cutoff;  /* LJ, md units */
    const float* subset_mask;      /* device (n) 0/1 floats, NULL = all particles move (mcmc.py:742-745) */
    const uint32_t* subset_ids;    /* device ids of the moved particles (CHX_MC_LJ_SUBSET_DELTA) */
    int n_subset;
    const uint32_t* neighbor_list; /* device (n,M), CHX_MC_LJ_NLIST */
    const int32_t* n_neighbors;    /* device (n) */
    int M;
    const float* ref_positions;    /* device (n,3): positions at the last list build; NULL = no check */
    float skin;                    /* the loop halts before a proposal with a displacement >= skin/2 */
    const float* x0;               /* device (n0,3) harmonic-oscillator centres, n0 = 1 or n */
    int n0;
    float k, U0;
    double beta;                   /* mol/kJ */
    double pv;                     /* P V N_A in kJ/mol (0 without a pressure), states.py:313-323 */
} chx_mc_displace_args;
typedef struct {
    uint32_t key[2];               /* SamplerState.current_PRNG_key, advanced by the loop */
    int32_t sel;                   /* which position buffer holds the current configuration */
    int32_t have_u;                /* 0: evaluate u of the current configuration first */
    float u_current;               /* reduced potential of the current configuration */
    float sigma_disp;              /* displacement_sigma, nm */
    int32_t n_accepted, n_proposed;/* running counters (MCMove.statistics) */
    int32_t moves_done;            /* moves completed by this call */
    int32_t halt;                  /* 1: move number `moves_done` needs a neighbour-list rebuild and was not started */
    int32_t nan_seen;              /* proposals rejected because u' was NaN (mcmc.py:417-430) */
    int32_t reserved[5];
} chx_mc_state;
/* x0/x1: device (n,3) buffers; state_dev: device scratch of sizeof(chx_mc_state); state_host is
 * uploaded first and holds the final state on return.  Synchronises once, at the end.
 * Large systems: three launches per move (propose, energy, decide) replayed as a cached CUDA graph of
 * 10 moves.  n <= 2048 with a cheap energy (harmonic oscillator, ideal gas, subset delta, tiny lists):
 * the whole call is ONE single-CTA launch with both position buffers in shared memory. */
int chx_mc_displace_run(chx_ctx* ctx, const chx_mc_displace_args* args, float* x0, float* x1,
                        chx_mc_state* state_dev, chx_mc_state* state_host, int n_moves);

/* n_moves Monte Carlo BAROSTAT steps (chiron/mcmc.py:913-1009 MonteCarloBarostatMove._propose + :357-463
 * _step) for LJPotential over a periodic NeighborListNsqrd, without a host round trip per step.  Per
 * move: split the key; V1 = V0 + uniform(subkey,-1,1) * volume_max_scale * V0 and s = (V1/V0)^(1/3) in
 * fp32 like mcmc.py:956-974; x' = x * s, box' = box * s; the neighbour list is REBUILT on (x', box') into
 * the list set that is not current (cell grid of the proposed box, computed on the device; arrays
 * identical to chx_nlist_build_cell); u' = beta (U(x') + P V'); accept iff the reference's test on
 * -(u' - u) + N log(V1/V0) passes.  x[state.sel] / list set state.sel / state.box are the current
 * configuration.  The loop stops BEFORE a move (state.halt = 1, key not advanced) when the proposed box
 * has fewer than 3 cells per edge or more cells than `ncell_capacity`, or when a row reaches M
 * neighbours (the reference's growth trigger, neighbors.py:709): the caller makes that move through
 * the building blocks and re-enters. */
typedef struct {
    int n;                          /* particles */
    float sigma, epsilon, cutoff;   /* LJ, md units */
    float cutoff_plus_skin;         /* list radius, summed in double and rounded once like neighbors.py:674 */
    int M;                          /* n_max_neighbors of both list sets */
    uint32_t* neighbor_list[2];     /* device (n,M) each; set state.sel holds the current list on entry and the
                                       other set a complete valid list too (a copy): rows are updated in place */
    int32_t* neighbor_mask[2];      /* device (n,M) */
    int32_t* n_neighbors[2];        /* device (n) */
    double beta;                    /* mol/kJ */
    double pressure;                /* P N_A in kJ/mol/nm^3 */
    int ncell_capacity;             /* 0 = 2 x the cells of the entry box */
    /* Optional superset list (device (n,M) ids and (n) counts, caller-allocated scratch): when given, one
     * list of radius (cutoff+skin)(1+delta) is built at the start of the call, delta bounding the
     * compression n_moves moves can produce, and every proposal's list is obtained by FILTERING it with the
     * exact predicate on the scaled positions (rows stay in id order, so the arrays are the ones a fresh
     * build gives) instead of a 125-cell sweep.  Falls back to the sweep when delta > 3 % or a superset row
     * does not fit M. */
    uint32_t* superset_list;
    int32_t* superset_nn;
} chx_mc_barostat_args;
typedef struct {
    uint32_t key[2];
    int32_t sel;                    /* position buffer AND list set of the current configuration */
    int32_t have_u;
    float u_current;
    float volume_max_scale;
    int32_t n_accepted, n_proposed;
    int32_t moves_done;
    int32_t halt;
    int32_t nan_seen;
    float box[3];                   /* current box lengths (in/out) */
    float last_volume;              /* volume of the last evaluated proposal (ThermodynamicState.volume) */
    int32_t reserved;
} chx_mc_baro_state;
int chx_mc_barostat_run(chx_ctx* ctx, const chx_mc_barostat_args* args, float* x0, float* x1,
                        chx_mc_baro_state* state_dev, chx_mc_baro_state* state_host, int n_moves);

/* ---- Fused LJ Langevin engine (integrators.py:110-218 + neighbors.py + potential.py) -------------- */
/* Runs whole trajectories on the device: cell-sorted particles, counting-sort cell list,
 * tiled neighbour structure, ONE kernel per Langevin step (forces over the tiles, then the BAOAB
 * update + wrap + rebuild checks of the warp's own particles), device-side rebuild decision.
 * Produces the same positions / velocities / energies as the building blocks above. */
typedef struct {
    int n;                 /* particles */
    float lx, ly, lz;      /* box */
    float sigma, epsilon;  /* LJ parameters, md units */
    float cutoff, skin;    /* nm; rebuild when any displacement >= skin/2 (neighbors.py:864) */
    float dt, gamma, kT;   /* ps, 1/ps, kJ/mol */
    int n_replicas;        /* >= 1: independent replicas batched in one launch (blockIdx.y) */
    float internal_skin;   /* skin of the engine's own neighbour tables, capped at `skin`; 0 = pick
                              automatically (0.35 sigma; 0.53 sigma for batched replicas).  A tuning knob: the pair set evaluated
                              each step (d < cutoff, exact predicate) does not depend on it */
} chx_ljmd_params;

int chx_ljmd_create(chx_ctx* ctx, const chx_ljmd_params* p, chx_ljmd** out);
int chx_ljmd_destroy(chx_ljmd* md);
/* Upload state (device pointers, original particle order): x (R,N,3); v (R,N,3); mass (N);
 * kT_per_replica_host (R) nullable = params.kT for all.  Sorts, builds the neighbour tables
 * (reference positions := x, like nbr_list.build_from_state, integrators.py:169) and evaluates
 * the forces (integrators.py:171).  Synchronises. */
int chx_ljmd_set_state(chx_ljmd* md, const float* x, const float* v, const float* mass,
                       const float* kT_per_replica_host);
/* Download state into original particle order; any pointer may be NULL. */
int chx_ljmd_get_state(chx_ljmd* md, float* x, float* v, float* force, float* ref_x);
/* Advance all replicas by nsteps BAOAB steps.  keys_host: (R,2) uint32 loop keys, updated in place
 * to the keys after the loop (integrators.py:179).  energies_dev (double, (R, n_reports)) nullable:
 * potential energy after every step with step % report_interval == 0 (integrators.py:197-205).
 * Synchronises at the end. */
int chx_ljmd_run(chx_ljmd* md, int nsteps, uint32_t* keys_host, int report_interval,
                 double* energies_dev, int n_reports_capacity);
/* Batched replicas (R > 1) rebuild their tables together every CH steps, counted from the last set_state across
 * runs.  A phase num/den in [0, 1) makes the first interval after set_state CH * (1 - num/den) steps long, so that
 * two engines sharing one GPU (separate contexts and streams, one host thread each) rebuild at different times
 * and each one's latency-bound rebuild overlaps the other's step kernels.  Takes effect at the next set_state /
 * run.  No reference counterpart (the reference propagates replicas one after the other, multistate.py:497-510). */
int chx_ljmd_set_chunk_phase(chx_ljmd* md, int num, int den);
/* Tell an engine that n_engines engines (separate contexts / streams / host threads) step concurrently on its GPU:
 * the number of warps a block is split over is then chosen for 1/n_engines of the machine.  No reference
 * counterpart. */
int chx_ljmd_set_gpu_share(chx_ljmd* md, int n_engines);
/* Pre-build (mode 0 = off, the default).  When on, a run of batched replicas (R > 1) that ends with tables due for their
 * rebuild (a) evaluates the potential energies of the final positions on the tables that are still valid and keeps
 * them for the next chx_ljmd_energy, and (b) enqueues the rebuild on a side stream before it returns, so that it
 * overlaps what the caller does between two runs (a replica-exchange sweep: energies -> all-gather -> swap
 * decisions -> chx_ljmd_set_kt).  chx_ljmd_scale_velocities / get_state / set_state wait for it on the device, the
 * next chx_ljmd_run checks its overflow bits.  Same tables, same trajectories as without.  No reference counterpart. */
int chx_ljmd_set_prebuild(chx_ljmd* md, int mode);
/* mode 2: the run only caches the energies; the rebuild is enqueued by chx_ljmd_prebuild_now -- e.g. right behind the
 * caller's collective, so that the collective's kernel is not queued behind the rebuild's grids -- or, if that call
 * never comes, by the next run.  mode 1: enqueued by the run itself.  mode 0: off. */
int chx_ljmd_prebuild_now(chx_ljmd* md);
/* Replica exchange support (new; the reference's _perform_swap_proposals is a stub, multistate.py:447-460):
 * change the temperature each replica is thermostatted at (kT, kJ/mol, (R) host floats) and rescale
 * its velocities (v *= scale[r], e.g. sqrt(T_new / T_old)).  Coordinates never move between replicas. */
int chx_ljmd_set_kt(chx_ljmd* md, const float* kT_per_replica_host);
int chx_ljmd_scale_velocities(chx_ljmd* md, const float* scale_per_replica_host);
/* Potential energy of the current positions per replica (double, device, (R)). Asynchronous. */
int chx_ljmd_energy(chx_ljmd* md, double* energy_dev);
/* Measurement hooks (bench.py roofline): launch the force kernel `repeats` times on the current
 * positions (asynchronous); run an FFMA-chain microbenchmark and return the fp32 FLOP count it
 * executed in *flops_host (asynchronous, time it with events on the context's stream); a negative
 * `iters` runs |iters| iterations of the packed FFMA2 variant. */
int chx_ljmd_force_only(chx_ljmd* md, int repeats);
int chx_fma_peak(chx_ctx* ctx, int iters, double* flops_host);
/* Host statistics (8 values): [0]=table rebuilds, [1]=candidate pairs i<j (d<cutoff+internal skin)
 * at the last build, [2]=interacting pairs i<j (d<cutoff) at the last energy evaluation,
 * [3]=steps run, [4]=kernel launches, [5]=reference rebuild events (neighbors.py:903-905) summed
 * over replicas, [6]=table capacity (tiles per block), [7]=blocks per replica.  Synchronises. */
int chx_ljmd_stats(chx_ljmd* md, long long* stats_host8);
/* Shape of the engine's neighbour tables (4 values): [0]=lane slots the force kernel spends on the
 * tiles of the last build (32 lanes x 2 partners x packed trips; [1] of chx_ljmd_stats x 2 / this =
 * lane utilisation), [1]=list words per tile, [2]=bytes per tile, [3]=candidate capacity per block.
 * Synchronises. */
int chx_ljmd_table_stats(chx_ljmd* md, long long* out4);
/* Device time of the step kernel, measured with CUDA events on the launch stream around every CUDA-graph
 * replay of a chunk of steps in which no replica stopped for a table rebuild (so every launch did its
 * full work): *total_ms_host over *steps_host launches since creation or the last reset. */
int chx_ljmd_step_timing(chx_ljmd* md, double* total_ms_host, long long* steps_host, int reset);

/* ---- generalisation of LJPotential the reference's API hints at (potential.py:131-137 takes
 * one sigma / epsilon): per-particle parameters (N floats each) with Lorentz-Berthelot mixing, sigma_ij = (sigma_i +
 * sigma_j)/2, eps_ij = sqrt(eps_i eps_j), over a NeighborListNsqrd; shift != 0 subtracts every pair's energy at the
 * cutoff (continuous potential; forces unchanged); switch_distance > 0 multiplies every pair energy by OpenMM's
 * switching function S = 1 - 6t^5 + 15t^4 - 10t^3, t = (r - switch_distance)/(cutoff - switch_distance), beyond
 * switch_distance (energy and force continuous at the cutoff).  With uniform parameters, shift = 0 and
 * switch_distance = 0 it equals chx_lj_nlist_energy_force. */
int chx_lj_nlist_energy_force_mixed(chx_ctx* ctx, const float* x, int n, float lx, float ly, float lz, int periodic,
                                    const uint32_t* neighbor_list, const int32_t* n_neighbors, int M,
                                    const float* sigma_per_particle, const float* epsilon_per_particle, float cutoff,
                                    int shift, float switch_distance, double* energy_dev, float* force);

/* ---- x64 variants (the reference with jax_enable_x64: float64 positions, displacements, energies, forces) -------
 * Same semantics, argument order and padding rules as the fp32 entry points they mirror, one IEEE float64 rounding
 * per reference operation (BASELINE.json north_star: pair sets bit-exact, energies / forces within rel 1e-10).
 *   chx_displacement_f64 / chx_wrap_f64      chiron/neighbors.py:45-112, 116-175
 *   chx_nlist_build_nsq_f64                  chiron/neighbors.py:595-626, 671-729 (O(N^2) restatement)
 *   chx_nlist_calculate_f64 / _check_f64     chiron/neighbors.py:773-787, 864-907
 *   chx_lj_nlist_energy_force_f64            chiron/potential.py:193-300 over the masked half list; energy_dev
 *                                            (1 double) and force_dev (N,3) are zeroed first, either may be NULL
 *   chx_baoab_update_f64 / chx_kick_f64      chiron/integrators.py:181-189, 195 with the (N,3) noise handed in
 *                                            (the float64 jax.random stream is not generated in-kernel) */
int chx_displacement_f64(chx_ctx* ctx, const double* x1, const double* x2, long long n, double lx, double ly,
                         double lz, int periodic, double* r_out, double* dist_out);
int chx_wrap_f64(chx_ctx* ctx, const double* x, long long n, double lx, double ly, double lz, double* out);
int chx_nlist_build_nsq_f64(chx_ctx* ctx, const double* x, int n, double lx, double ly, double lz, int periodic,
                            double cutoff_plus_skin, int M, uint32_t* neighbor_list, int32_t* neighbor_mask,
                            int32_t* n_neighbors, int* max_count_host, int* count_eq_M_host);
int chx_nlist_calculate_f64(chx_ctx* ctx, const double* x, int n, double lx, double ly, double lz, int periodic,
                            double cutoff, int M, const uint32_t* neighbor_list, const int32_t* neighbor_mask,
                            int32_t* n_out, int32_t* mask_out, double* dist_out, double* rij_out);
int chx_nlist_check_f64(chx_ctx* ctx, const double* x, const double* ref_x, int n, double lx, double ly, double lz,
                        int periodic, double half_skin, int32_t* flag_dev);
int chx_lj_nlist_energy_force_f64(chx_ctx* ctx, const double* x, int n, double lx, double ly, double lz, int periodic,
                                  double sigma, double epsilon, double cutoff, int M, const uint32_t* neighbor_list,
                                  const int32_t* neighbor_mask, double* energy_dev, double* force_dev);
int chx_baoab_update_f64(chx_ctx* ctx, double* x, double* v, const double* F, const double* mass, const double* noise,
                         int n, double half_dt, double a, double b, double kT, double lx, double ly, double lz, int wrap);
int chx_kick_f64(chx_ctx* ctx, double* v, const double* F, const double* mass, int n, double half_dt);

#ifdef __cplusplus
}
#endif
#endif /* CHIRON_B200_H */
