/*
 * particulator_b200.h — C ABI of the B200-native particle-advance library.
 *
 * This is the drop-in boundary for the hot path of aluque/Particulator.jl
 * (advance! -> advance1! -> push + null-collision sampling + apply + droplow!/repack!).
 * The reference has no FFI of its own (pure Julia, multiple dispatch); every entry
 * point below cites the Julia generic function / constructor it replaces, as
 * `file:line` relative to the reference tree.  A Julia host binds these with plain
 * `ccall`; nothing here uses torch, C++ or CUDA types.
 *
 * Conventions
 *   - every function returns int32 status: 0 = ok, <0 = usage error, >0 = sticky
 *     device condition bit-set (PTL_ERR_*), unless documented as returning an id/count.
 *   - handles are small non-negative integers scoped to a context.
 *   - host arrays are borrowed for the duration of the call only.
 *   - x / p are xyz-interleaved (3 doubles per particle), exactly the memory of the
 *     reference's StructArray columns `particles.x`, `particles.p`
 *     (Vector{SVector{3,Float64}}, src/population.jl:14,37).
 *   - one host thread per context, one context per GPU.
 */
#ifndef PARTICULATOR_B200_H
#define PARTICULATOR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTL_ABI_VERSION 1

/* ---- species: src/particledefs.jl:18-22 (ParticleType{:electron|:photon|:positron}),
 *      src/slow-electron.jl:4 (ParticleType{:slow_electron}) ------------------------- */
enum {
    PTL_ELECTRON = 0,
    PTL_PHOTON = 1,
    PTL_POSITRON = 2,
    PTL_SLOW_ELECTRON = 3,
    PTL_NSPECIES = 4
};

/* ---- collision process kinds (one per `collide` method) ---------------------------- */
enum {
    PTL_PROC_NULL = 0,          /* NullCollision           src/collisions.jl:3,58            */
    PTL_PROC_COULOMB = 1,       /* RelativisticCoulomb     src/relativistic_coulomb.jl:5-23  par: Z */
    PTL_PROC_RBEB = 2,          /* RBEB                    src/rbeb.jl:5-80                  par: B,U,N */
    PTL_PROC_MOLLER = 3,        /* Moller                  src/moller.jl:8-37                par: Z,tcut */
    PTL_PROC_BHABA = 4,         /* Bhaba                   src/bhaba.jl:4-33                 par: Z,tcut */
    PTL_PROC_SELTZER = 5,       /* SeltzerBerger           src/seltzer.jl:9-90               aux: SB table id */
    PTL_PROC_COMPTON = 6,       /* Compton / KleinNishinaCompton src/compton.jl:1-28         par: Z */
    PTL_PROC_PHOTOELECTRIC = 7, /* PhotoElectric           src/photo_electric.jl:7-52        par: Z,nbind,bind[0..3] */
    PTL_PROC_BETHE_HEITLER = 8, /* BetheHeitler            src/bethe_heitler.jl:1-25         par: Z */
    PTL_PROC_ANIHILATION = 9,   /* PositronAnihilation     src/anihilation.jl:1-23           par: Z */
    PTL_PROC_LX_EXCITATION = 10,/* Excitation              src/slow-electron.jl:70,108       par: threshold */
    PTL_PROC_LX_IONIZATION = 11,/* Ionization              src/slow-electron.jl:74,116       par: threshold */
    PTL_PROC_LX_ATTACHMENT = 12,/* Attachment              src/slow-electron.jl:78,129       par: threshold */
    PTL_PROC_LX_ELASTIC = 13,   /* Elastic                 src/slow-electron.jl:82,133       par: mass_ratio */
    PTL_NPROC_KINDS = 14
};

#define PTL_MAX_PROCS 128       /* processes per table (reference: tuple length L, collisions.jl:145) */
#define PTL_PROC_NPAR 6

typedef struct {
    int32_t kind;               /* PTL_PROC_*                                               */
    int32_t aux;                /* Seltzer-Berger table id (ptl_sb_table_create) or -1      */
    double  par[PTL_PROC_NPAR]; /* kind-specific parameters, SI units (J, m)                */
} ptl_process_desc;

/* ---- sticky device conditions (bit-set) -------------------------------------------- */
enum {
    PTL_ERR_CAPACITY_OVERFLOW   = 1,  /* @assert n < length(particles)   src/population.jl:107   */
    PTL_ERR_RATE_BOUND_VIOLATED = 2,  /* @assert xi >= 0                 src/collisions.jl:186   */
    PTL_ERR_ENERGY_OUT_OF_TABLE = 4,  /* rate[k,j,i+1] out of bounds     src/collision_table.jl:91 */
    PTL_ERR_NAN_STATE           = 8,
    PTL_ERR_SAMPLER_INVARIANT   = 16  /* src/rbeb.jl:63, src/seltzer.jl:73, src/photo_electric.jl:71 */
};

/* ---- usage errors (negative return values) ------------------------------------------ */
enum {
    PTL_EINVAL = -1,
    PTL_ENODEVICE = -2,   /* no sm_100 device: there is NO CPU fallback */
    PTL_ECUDA = -3,
    PTL_ENOMEM = -4,
    PTL_EHANDLE = -5,
    PTL_ECOMM = -6,       /* NCCL runtime missing or a collective failed (ptl_last_error has the text) */
    PTL_ECAPACITY = -7    /* ptl_population_append on a full population (@assert population.jl:107); the sticky bit is set too */
};

/* uid space: uids key the per-particle RNG streams.  Bit 63 clear = sequential uids handed out by the host side
 * (context counter, ptl_set_uid_counter; rank r of a communicator starts at r * 2^40 + 1); bit 63 set = uids of
 * particles born in collisions, hashed from (parent uid, parent draw counter). */
#define PTL_UID_HASHED_BIT 0x8000000000000000ull

/* ---- fields: src/field.jl:4-52 ----------------------------------------------------- */
enum {
    PTL_FIELD_ZERO = 0,
    PTL_FIELD_HOMOGENEOUS = 1,   /* HomogeneousField(v)            par = v[3]                 field.jl:4-8   */
    PTL_FIELD_DOUBLE_LAYER = 2,  /* DoubleLayerField(z1,z2,v)      par = z1,z2,v[3]           field.jl:10-16 */
    PTL_FIELD_STEP = 3,          /* StepField(z,v1,v2)             par = z,v1[3],v2[3]        field.jl:22-28 */
    PTL_FIELD_CONFINED_DL = 4    /* ConfinedDoubleLayerField       par = sx,sy,sz,ez0         field.jl:35-52 */
};

typedef struct {
    int32_t kind;
    int32_t _pad;
    double  par[7];
} ptl_field_desc;

/* ---- forcings: src/pusher.jl:8-34, src/field.jl:54-70, src/continuum.jl:6-57 ------- */
enum {
    PTL_FORCE_NONE = 0,            /* NullForcing                pusher.jl:11                      */
    PTL_FORCE_EM = 1,              /* ElectromagneticField(e,b)  field.jl:54-70                    */
    PTL_FORCE_CONTINUUM = 2,       /* ContinuumLoss(nel,I,Tcut)  continuum.jl:6-22,63-96           */
    PTL_FORCE_CHEB_CONTINUUM = 3   /* ChebContinuumLoss{N}       continuum.jl:25-57  (cheb_id)     */
};

#define PTL_MAX_FORCINGS 4

typedef struct {
    int32_t  kind;
    uint32_t species_mask;   /* RestrictedForcing{T}: bit s set => acts on species s (pusher.jl:28-34); 0 => all */
    ptl_field_desc e, b;     /* PTL_FORCE_EM */
    double   nel, I, Tcut;   /* PTL_FORCE_CONTINUUM */
    int32_t  cheb_id;        /* PTL_FORCE_CHEB_CONTINUUM: id from ptl_cheb_loss_create */
    int32_t  _pad;
} ptl_forcing_desc;

/* ---- pushers: src/pusher.jl:37-76 --------------------------------------------------- */
enum {
    PTL_PUSHER_NULL = 0,   /* NullPusher: t += dt                     pusher.jl:75-76 */
    PTL_PUSHER_RK2 = 1     /* RK2Pusher(CombinedForcing(...))         pusher.jl:37-63 */
};

typedef struct {
    int32_t  kind;
    uint32_t restrict_mask;  /* RestrictedPusher{T}: bit s set => species s is pushed, others only t += dt
                                (pusher.jl:67-73); 0 => all species pushed */
    int32_t  nforcings;      /* CombinedForcing tuple length (pusher.jl:14-23) */
    int32_t  _pad;
    ptl_forcing_desc forcing[PTL_MAX_FORCINGS];
} ptl_pusher_desc;

/* ---- in-loop callbacks: src/callback.jl:41-184 -------------------------------------- */
#define PTL_MAX_WALLS 4

typedef struct {
    int32_t species;   /* WallCallback{P}: state type recorded       callback.jl:146-164 */
    int32_t coord;     /* 0,1,2  (reference: 1-based coord)                              */
    double  v;         /* wall position                                                   */
    int32_t drop;      /* deactivate after recording                                      */
    int32_t _pad;
} ptl_wall_desc;

typedef struct {
    int32_t nwalls;            /* CombinedCallback of WallCallbacks  callback.jl:41-108   */
    int32_t count_collisions;  /* CollisionCounter                   callback.jl:118-131  */
    ptl_wall_desc wall[PTL_MAX_WALLS];
} ptl_callback_desc;           /* NULL pointer == VoidCallback       callback.jl:9        */

/* ---- diagnostics: src/population.jl:78-223 ------------------------------------------ */
typedef struct {
    int64_t n;          /* nparticles   population.jl:78                                  */
    int64_t nactive;    /* nactives     population.jl:89-97                               */
    double  weight;     /* weight       population.jl:130-140  (sum of w over actives)    */
    double  wenergy;    /* sum w*E  -> meanenergy = wenergy/weight  population.jl:152-166 */
    double  maxenergy;  /* maxenergy    population.jl:172-174  (all rows < n, active or not) */
    double  wx[3];      /* sum w*x      -> spread/posvar  population.jl:180-223           */
    double  wx2[3];     /* sum w*x.^2                                                      */
    double  wr2;        /* sum w*dot(x,x)                                                  */
} ptl_diag_out;

/* ===================================================================================== */
/* context                                                                               */
/* ===================================================================================== */
typedef struct ptl_context ptl_context;

/* Create a context on CUDA device `device`. `stream` is a cudaStream_t passed as void*
 * (NULL => the library creates its own non-blocking stream).  Fails with PTL_ENODEVICE
 * when no compute-capability-10.x device is present: there is no CPU fallback. */
int32_t ptl_context_create(int32_t device, void* stream, ptl_context** out);
int32_t ptl_context_destroy(ptl_context* ctx);
int32_t ptl_abi_version(void);
const char* ptl_last_error(ptl_context* ctx);
/* Return and (optionally) clear the sticky PTL_ERR_* bit-set (synchronises the stream). */
int32_t ptl_error_flags(ptl_context* ctx, int32_t clear);
int32_t ptl_synchronize(ptl_context* ctx);
/* Counter-based RNG state.  Replaces the task-local `rand()` of the reference
 * (src/util.jl:17 and every `collide`): stream = Philox4x32-10 keyed by particle uid,
 * counter = (draw index, advance-call index `step`, seed). */
int32_t ptl_set_rng(ptl_context* ctx, uint64_t seed, uint32_t step);
int32_t ptl_get_rng(ptl_context* ctx, uint64_t* seed, uint32_t* step);
/* Counter behind default uids (uploads / appends without explicit uids).  Restart state: a restored run
 * must not reissue a live uid.  ptl_population_upload with explicit uids raises it past max(uid). */
int32_t  ptl_set_uid_counter(ptl_context* ctx, uint64_t next_uid);
uint64_t ptl_get_uid_counter(ptl_context* ctx);
/* Tuning knobs outside the reference's surface: "kernel" = lepton advance kernel variant (0 default,
 * 3 list-scheduled, 4 re-sorting, 5 warp-private pools; all give identical results), "stream" = 0/1
 * streaming fast path for low-kappa species, "stream_tma" = 0/1 TMA-staged streaming kernel for leptons, "small_pass_rows" = lepton passes with fewer rows than this run
 * on the one-particle-per-lane kernel (latency regime; 0 = always the wavefront kernel), "overlap" = 0/1 run the
 * species of one pass of advance1! concurrently on forked streams (they are independent of each other). */
int32_t ptl_set_option(ptl_context* ctx, const char* name, int64_t value);

/* ===================================================================================== */
/* tables (host-built flat arrays; builders stay on the host, src/collision_table.jl:115-167) */
/* ===================================================================================== */

/* SeltzerBerger sampling table (src/seltzer.jl:9-49): data[ncum, nE] column-major
 * (ncum fastest), log_energy[nE]; pcum = LinRange(0,1,ncum).  Returns id >= 0. */
int32_t ptl_sb_table_create(ptl_context* ctx, int32_t ncum, int32_t nE,
                            const double* log_energy, const double* data);

/* ChebyshevCollisionTable (src/collision_table.jl:63-106): rate[order, nprocs, k+1]
 * column-major (order fastest), ratebound[order, k+1]; BinaryIntervals(k, xmax)
 * (src/cheby.jl:15-21).  Returns table id >= 0. */
int32_t ptl_table_create_cheb(ptl_context* ctx, int32_t order, int32_t nprocs, int32_t k,
                              double xmax, const double* rate, const double* ratebound,
                              const ptl_process_desc* procs);

/* CollisionTable (src/collision_table.jl:15-57): rate[nprocs, nE] column-major (process
 * fastest) on a LinRange (grid_kind 0, L1=first, L2=last) or LogLinRange (grid_kind 1,
 * x = exp(L) - exp(L1), src/util.jl:60-127); constant rate bound `maxrate`
 * (collision_table.jl:33).  The explicit NullCollision row of load_lxcat
 * (src/lxcat.jl:114-121) is just a process of kind PTL_PROC_NULL. Returns table id. */
int32_t ptl_table_create_linear(ptl_context* ctx, int32_t grid_kind, double L1, double L2,
                                int32_t nE, int32_t nprocs, const double* rate,
                                double maxrate, const ptl_process_desc* procs);

/* The same with a VECTOR rate bound on the energy grid (collision_table.jl:35-43):
 * ratebound(E) = w * v[k] + (1 - w) * v[k + 1] with the (k, w) of the rate lookup; ratebound_vec[nE]. */
int32_t ptl_table_create_linear_vb(ptl_context* ctx, int32_t grid_kind, double L1, double L2,
                                   int32_t nE, int32_t nprocs, const double* rate,
                                   const double* ratebound_vec, const ptl_process_desc* procs);

/* ChebContinuumLoss coefficient matrices (src/continuum.jl:25-43): ec, pc [order, k+1]. */
int32_t ptl_cheb_loss_create(ptl_context* ctx, int32_t order, int32_t k, double xmax,
                             const double* ec, const double* pc);

/* Bit-exact tier entry point: evaluate presample + rate(j) for every process and the
 * rate bound at `n` host energies through the SAME device functions the advance kernel
 * uses (src/collision_table.jl:50-57,82-106; src/cheby.jl:127-143).
 * rates_out[nprocs, n] (process fastest), bound_out[n]. */
int32_t ptl_table_eval(ptl_context* ctx, int32_t table, int64_t n, const double* energy,
                       double* rates_out, double* bound_out);

/* ===================================================================================== */
/* populations: src/population.jl                                                        */
/* ===================================================================================== */

/* Population(max_particles, init, collisions, energy_cut)  population.jl:34-44 */
int32_t ptl_population_create(ptl_context* ctx, int32_t species, int64_t capacity,
                              double energy_cut, int32_t table);
int32_t ptl_population_destroy(ptl_context* ctx, int32_t pop);

/* Replace rows [0,n) with host data; sets popl.n = n.  `uid` may be NULL (uids are then
 * assigned from a context counter).  uid keys the RNG stream of each particle. */
int32_t ptl_population_upload(ptl_context* ctx, int32_t pop, int64_t n,
                              const double* x3, const double* p3, const double* w,
                              const double* t, const double* s, const double* r,
                              const uint8_t* active, const uint64_t* uid);
/* Copy rows [0, min(n, max_n)) to host; any pointer may be NULL. Returns rows copied. */
int64_t ptl_population_download(ptl_context* ctx, int32_t pop, int64_t max_n,
                                double* x3, double* p3, double* w, double* t, double* s,
                                double* r, uint8_t* active, uint64_t* uid);

int64_t ptl_population_n(ptl_context* ctx, int32_t pop);          /* nparticles   population.jl:78  */
int64_t ptl_population_capacity(ptl_context* ctx, int32_t pop);
int32_t ptl_population_clear(ptl_context* ctx, int32_t pop);      /* empty!       population.jl:69  */
/* add_particle!(popl, state) population.jl:103-113: returns new row index (0-based), -1 if below the
 * energy cut, PTL_ECAPACITY when the population is full.  Slow path (one particle, host-synchronous). */
int64_t ptl_population_append(ptl_context* ctx, int32_t pop, const double* x3, const double* p3,
                              double w, double t, double s, double r, uint64_t uid);
int32_t ptl_population_deactivate(ptl_context* ctx, int32_t pop, int64_t i); /* remove_particle! :120 */

/* droplow!(popl, thres) population.jl:273-284: flag E < thres (thres==0 => energy_cut),
 * then repack!.  Returns the new n (>= 0) or a negative usage error. */
int64_t ptl_droplow(ptl_context* ctx, int32_t pop, double thres);
/* repack!(popl) population.jl:229-259: tail-fill compaction, same permutation. */
int64_t ptl_repack(ptl_context* ctx, int32_t pop);

/* nparticles/nactives/weight/meanenergy/maxenergy/spread/posvar in one fused reduction. */
int32_t ptl_diag(ptl_context* ctx, int32_t pop, ptl_diag_out* out);
/* Weighted histogram of kinetic energy (quantity 0) or cos(theta_z) = p_z/|p| (quantity 1)
 * over active particles; nbins uniform bins on [lo,hi) (log10-spaced in the quantity if
 * logscale).  Replaces the scripts' StatsBase histogram (scripts/beam.jl:137-146). */
int32_t ptl_histogram(ptl_context* ctx, int32_t pop, int32_t quantity, double lo, double hi,
                      int32_t nbins, int32_t logscale, double* out);

/* roulette!(p, popl) population.jl:291-309 with constant retain probability p.
 * Draws come from the particle's Philox stream (domain-separated from collision draws). */
int32_t ptl_roulette(ptl_context* ctx, int32_t pop, double p);
/* split!(p, popl) population.jl:316-335 with constant mean number of copies p. */
int32_t ptl_split(ptl_context* ctx, int32_t pop, double p);

/* roulette!(f, popl) / split!(f, popl) with an energy-dependent law (population.jl:291-309, 316-335).
 * A closure cannot cross the ABI: the host samples f on `n` nodes, uniform in E [J] (logscale 0) or in
 * log10(E) (logscale 1) between lo and hi, and the library interpolates linearly (flat outside the
 * range).  n == 1 is the constant law. */
int32_t ptl_roulette_law(ptl_context* ctx, int32_t pop, double lo, double hi, int32_t n,
                         int32_t logscale, const double* p);
int32_t ptl_split_law(ptl_context* ctx, int32_t pop, double lo, double hi, int32_t n,
                      int32_t logscale, const double* p);
/* shuffle!(popl) population.jl:266-271: a uniformly distributed permutation of rows [0, n), drawn
 * from the counter-based RNG (row index, seed, step) and applied to every column. */
int32_t ptl_shuffle(ptl_context* ctx, int32_t pop);

/* Raw device pointer of a column for zero-copy interop (NCCL send/recv of column tails,
 * device-side synthetic fills).  col: 0..2 = x0,x1,x2; 3..5 = p0,p1,p2; 6=w 7=t 8=s 9=r
 * (double, planar, `capacity` long); 10 = active (uint8); 11 = uid (uint64). */
void*   ptl_population_column_ptr(ptl_context* ctx, int32_t pop, int32_t col);
/* Set popl.n after an external device-side write into the columns. */
int32_t ptl_population_set_n(ptl_context* ctx, int32_t pop, int64_t n);

/* ===================================================================================== */
/* multi-population + the hot call: src/mixed_population.jl                              */
/* ===================================================================================== */

/* MultiPopulation(:electron => ..., :photon => ..., ...) mixed_population.jl:4-12.
 * Order of `pops` is the processing order of advance1! (:56-93). Returns id. */
int32_t ptl_multipop_create(ptl_context* ctx, const int32_t* pops, int32_t count);

/* init!(mpopl) mixed_population.jl:20-35: setr! on all actives. */
int32_t ptl_init(ptl_context* ctx, int32_t mp);

/* advance!(mpopl, pusher, tfinal, callback) mixed_population.jl:38-47.
 * Mutates all populations until every active particle has t == tfinal; appends births.
 * `cb` may be NULL (VoidCallback).  Increments the context's RNG step index. */
int32_t ptl_advance(ptl_context* ctx, int32_t mp, const ptl_pusher_desc* pusher, double tfinal,
                    const ptl_callback_desc* cb);

/* Statistics of the last ptl_advance call: passes of advance1!, total sub-steps (iterations of
 * mixed_population.jl:66), particle rows visited, births appended. */
typedef struct {
    int64_t passes;
    int64_t substeps;
    int64_t rows;
    int64_t births;
    int64_t launches;   /* kernels launched by the call */
    int64_t main_rows;  /* rows of the largest advance-kernel launch of the call */
    double  main_ms;    /* its duration, measured with CUDA events on the context's stream (0 unless profiling is on) */
} ptl_advance_stats;
int32_t ptl_last_advance_stats(ptl_context* ctx, ptl_advance_stats* out);

/* Measurement aids: cudaEvent timing of the dominant advance kernel (reported in ptl_advance_stats.main_ms)
 * and the number of kernels this context has launched since the last reset. */
int32_t ptl_set_profiling(ptl_context* ctx, int32_t on);
int64_t ptl_launch_count(ptl_context* ctx, int32_t reset);

/* CollisionCounter read-out (callback.jl:118-141): counts[nprocs+1] per table (last = null). */
int32_t ptl_collision_counts(ptl_context* ctx, int32_t table, int64_t* counts, int32_t clear);

/* WallCallback.accum read-out (callback.jl:146-184): recorded mid-states of wall `iwall`.
 * Returns number of records available; copies up to max_n. */
int64_t ptl_wall_records(ptl_context* ctx, int32_t iwall, int64_t max_n, double* x3, double* p3,
                         double* w, double* t, int32_t clear);


/* ===================================================================================== */
/* multi-GPU: one context per GPU + a communicator (SURVEY.md section 8b "threading", 8e) */
/* ===================================================================================== */
/* The reference is single-process; what replaces it at scale is one process (or host thread) per GPU, each
 * with its own context and an independent shard of every species.  The advance path needs NO collective.
 * NCCL (bound at run time: libnccl.so.2 or $PTL_NCCL_LIB) serves the two exchange steps the path has:
 * reducing what run! prints and the population-control callbacks decide on (src/run.jl:31-40,
 * src/callback.jl:203,217,239,263), and periodic population rebalancing. */
#define PTL_COMM_ID_BYTES 128
/* ncclGetUniqueId: call on ONE rank, ship the 128 bytes to the others by any host-side means. */
int32_t ptl_comm_unique_id(uint8_t* id_out);
/* ncclCommInitRank on the context's device.  Also moves the context's default-uid counter to
 * rank * 2^40 + 1 so that uids assigned by different ranks never collide (uids key the RNG streams). */
int32_t ptl_comm_init(ptl_context* ctx, const uint8_t* id, int32_t rank, int32_t nranks);
int32_t ptl_comm_destroy(ptl_context* ctx);
/* Returns 1 when a communicator is attached, 0 otherwise; rank / nranks may be NULL. */
int32_t ptl_comm_info(ptl_context* ctx, int32_t* rank, int32_t* nranks);
/* ptl_diag with GLOBAL sums / max over all ranks (identity without a communicator). */
int32_t ptl_diag_allreduce(ptl_context* ctx, int32_t pop, ptl_diag_out* out);
/* ptl_histogram summed over all ranks. */
int32_t ptl_histogram_allreduce(ptl_context* ctx, int32_t pop, int32_t quantity, double lo, double hi,
                                int32_t nbins, int32_t logscale, double* out);
/* In-place all-reduce of a small host vector (n <= 4096); op 0 = sum, 1 = max, 2 = min. */
int32_t ptl_comm_allreduce_f64(ptl_context* ctx, double* inout, int32_t n, int32_t op);
/* Deterministic rebalancing plan (pure host code, no GPU needed): counts[nranks] ->
 * moves_out[3*m] = (src, dst, k) "move the last k rows of src to the end of dst".  Nothing moves
 * while every rank is within `tolerance` (relative) of the mean.  Returns m >= 0. */
int32_t ptl_rebalance_plan(const int64_t* counts, int32_t nranks, double tolerance,
                           int64_t* moves_out, int32_t max_moves);
/* Rebalance one species over the communicator: all-gather of counts, the plan above, ONE grouped
 * ncclSend/ncclRecv over the 12 column tails (device to device).  Collective: every rank calls it.
 * Returns the local n afterwards; moved_out (may be NULL) = rows sent (+) or received (-). */
int64_t ptl_rebalance(ptl_context* ctx, int32_t pop, double tolerance, int64_t* moved_out);

/* ===================================================================================== */
/* test / diagnostic entry points (deterministic replay of single events)                */
/* ===================================================================================== */

/* Run collide() (A.6 draw order) of process `j` of `table` once per input momentum through the
 * SAME device functions the advance kernel uses; row i uses the Philox stream of uid0+i from
 * draw index 0 with the context's (seed, step).  No population is touched.
 * out[i*24 + ..]: [0]=outcome kind (0 null,1 state change,2 new particle,3 remove,4 replace,
 * 5 replace pair), [1]=species of state2, [2]=species of state3, [3]=draws consumed,
 * [4..6]=p of state1, [7]=s of state1, [8..10]=p2, [11]=s2, [12..14]=p3, [15]=s3. */
int32_t ptl_collide_test(ptl_context* ctx, int32_t species, int32_t table, int32_t j, int64_t n,
                         const double* p3, uint64_t uid0, double* out);
/* First n uniforms of the stream (uid, seed, step). */
int32_t ptl_rng_test(ptl_context* ctx, uint64_t uid, uint64_t seed, uint32_t step, int32_t n, double* out);

#ifdef __cplusplus
}
#endif
#endif /* PARTICULATOR_B200_H */
