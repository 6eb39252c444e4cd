/* lqcd_b200.h -- C ABI of the B200-native pure-gauge SU(3) update path.
 *
 * Drop-in boundary for lattice-qcd-rs v0.2.1 (a pure-Rust crate with no FFI of its own: README.md:32 lists a
 * "C friendly API / ABI" as not implemented).  Each entry point below names the reference loop it replaces
 * (file:line under /root/reference/src).  A Rust shim crate (rust/lattice-qcd-b200, authored)
 * binds exactly these symbols and implements the crate's traits on top of them.
 *
 * Conventions
 *   - every function returns 0 (LQ_OK) or a negative LQ_E_* code; nothing unwinds across the ABI.
 *   - host pointers are borrowed for the duration of the call; the context owns all device memory.
 *   - a context is bound to one CUDA device and one stream; calls on one context are not re-entrant
 *     (matches the reference, where a state value is moved through `next_element` by one thread).
 *   - array layouts at the boundary are the reference's own AoS layouts:
 *       links  : n_links * 18 f64, link = site*D + dir, 3x3 complex column-major (re,im)  field.rs:584-586
 *       efield : n_links *  8 f64, (site*D + dir)*8 + a                                   field.rs:1025-1027
 *       site   : sum_k x_k * prod_{l<k} extent_l   (x_0 fastest)                          lattice.rs:909-916
 *     On a decomposed (multi-rank) context "site" enumerates the rank-local block in the same order.
 *   - all arithmetic is f64 (lib.rs:69 `Real = f64`).
 */
#ifndef LQCD_B200_H
#define LQCD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lq_ctx lq_ctx;

enum {
  LQ_OK = 0,
  LQ_E_BADARG = -1,         /* null pointer, D out of range, extent < 2 (lattice.rs:190-201)            */
  LQ_E_SIZE = -2,           /* StateInitializationError::IncompatibleSize (state.rs:784-786)             */
  LQ_E_CUDA = -3,           /* any CUDA runtime failure (lq_last_cuda_error has the text)                */
  LQ_E_COMM = -4,           /* halo transport failure                                                    */
  LQ_E_ODD_EXTENT = -5,     /* sweeps on a DECOMPOSED context need even extents (single-rank contexts sweep odd
                               lattices in colour classes, see the sweep section)                        */
  LQ_E_GAUSS_DIVERGED = -6, /* StateInitializationError::GaussProjectionError (field.rs:1279-1281)       */
  LQ_E_ZERO_STEPS = -7,     /* MultiIntegrationError::ZeroIntegration (state.rs:331-333, 480-482)        */
  LQ_E_NOSNAPSHOT = -8,
  LQ_E_NODEVICE = -9        /* no CUDA device: there is NO CPU fallback                                  */
};

/* integrator compositions, symplectic_euler_rayon.rs:120-252 */
enum { LQ_SYNC_SYNC = 0, LQ_LEAP_LEAP = 1, LQ_SYNC_LEAP = 2, LQ_LEAP_SYNC = 3, LQ_SYMPLECTIC = 4 };
/* over-relaxation flavours, overrelaxation.rs:86-98 / 158-171 */
/* 0, 1: OverrelaxationSweepRotation / Reverse (overrelaxation.rs:86-98, 158-171; U(3)-valued, as the crate);
 * 2: option beyond the crate -- Brown-Woch reflections in the three SU(2) sub-groups of the heat bath (stays in SU(3)) */
enum { LQ_OR_ROTATION = 0, LQ_OR_REVERSE = 1, LQ_OR_SU2_SUBGROUPS = 2 };
/* behaviour flags (lq_set_flags) */
enum {
  LQ_FLAG_PAULI3_FIXED = 1,   /* use sigma_3 = diag(1,-1); default restates su2.rs:39-45 as coded (diag(1,1)) */
  LQ_FLAG_NO_KICK_MERGE = 2,  /* lq_symplectic_n: do not merge the two adjacent dt/2 E-kicks of consecutive steps
                                 (results are bit-identical either way; this only changes the launch count)    */
  LQ_FLAG_GENERIC_KERNELS = 8, /* use the dimension-generic functor kernels even where a tuned D = 4 kernel exists
                                 (the parity tests run both; results agree to rounding of the staple order)   */
  LQ_FLAG_GAUSS_FUSED = 4,    /* lq_gauss_project(_step): one fused kernel per iteration (projection step + Gauss
                                 field of the projected E, backward neighbours recomputed: 1376 instead of 2208
                                 B/site but 32 instead of 20 matrix products/site) instead of two passes.  Same
                                 results; measured slower on B200 (0.49 vs 0.43 ms at 32^4), so off by default */
  LQ_FLAG_GAUSS_TWO_PASS = 32, /* lq_gauss_project: iterate with the two-pass kernels (Gauss field, then projection step:
                                 2208 B/site) instead of the default D = 4 loop on the transported field U^+ E U (one
                                 kernel per iteration, links read once, 1728 B/site); same results to 1e-15        */
  LQ_FLAG_FOLD_HALO_SYNC = 512, /* decomposed contexts with the peer-memory transport: fold the halo synchronisation of
                                 the Gauss projection loop into its compute kernel (the last boundary block of a launch
                                 releases the epoch, the boundary blocks of the next acquire it) instead of one barrier
                                 kernel per ghost refresh; | 2048: the MD chain too; | 1024: its z faces scheduled
                                 first.  Same results; measured SLOWER on 8 B200s (profiles/r02r-u), so off by default */
  LQ_FLAG_UNIFORM_DIRECTION = 16 /* heat bath: draw the direction of the SU(2) vector uniformly on the sphere; default
                                 restates distribution.rs:199-219 as coded (a normalised Uniform(-1,1)^3 sample, which
                                 over-weights the cube diagonals).  With LQ_FLAG_PAULI3_FIXED and coupling_scale = 1/CA
                                 the sweep is the textbook Cabibbo-Marinari / Kennedy-Pendleton heat bath of the Wilson
                                 action (tests/test_physics.py: <P>/3 = 0.5937 at beta = 6)                          */
};

const char* lq_strerror(int code);
const char* lq_last_cuda_error(void);
int lq_version(void);
int lq_device_count(int* n);

/* ---- context ------------------------------------------------------------------------------------------------
 * LatticeCyclic::new (lattice.rs:190-201) + LatticeStateDefault/LatticeStateEFSyncDefault storage
 * (state.rs:655-659, 1048-1062).  `extent[D]` may differ per direction (the reference has a single `dim`). */
int lq_ctx_create(lq_ctx** out, int device, int D, const int64_t* extent, double lattice_spacing_a, double beta,
                  double CA);
/* Decomposed context: `global_extent[D]` split over `proc_grid[D]` ranks (only the last two directions may be
 * split); `rank_coord[D]` is this rank's position.  One-site-deep ghost layers are kept in the split directions
 * and filled through lq_halo_* (host-staged or peer-mapped transport supplied by the caller's plumbing). */
int lq_ctx_create_dist(lq_ctx** out, int device, int D, const int64_t* global_extent, const int* proc_grid,
                       const int* rank_coord, double lattice_spacing_a, double beta, double CA);
/* `Clone` of a device-resident state (SimulationStateSynchronous: Clone, state.rs:292-295; every integrator call
 * returns a NEW state, symplectic_euler_rayon.rs:245-251): same lattice / beta / flags / t, links and E-field
 * copied device to device.  Snapshots and profiling state are not copied. */
int lq_ctx_clone(const lq_ctx* src, lq_ctx** out);
int lq_ctx_destroy(lq_ctx*);
int lq_set_flags(lq_ctx*, int flags);
int lq_get_flags(lq_ctx*, int* flags);
int lq_set_beta(lq_ctx*, double beta);
int lq_sync(lq_ctx*);                              /* cudaStreamSynchronize on the context stream */
int lq_stream(lq_ctx*, void** cuda_stream_out);    /* the context's cudaStream_t (for CUDA-event timing by the caller) */
int64_t lq_num_sites(const lq_ctx*);               /* rank-local interior sites  (lattice.rs number_of_points)    */
int64_t lq_num_links(const lq_ctx*);               /* rank-local interior links  (number_of_canonical_links_space) */
int64_t lq_t(const lq_ctx*);                       /* LatticeStateWithEField::t  (state.rs:1060)                   */
int lq_set_t(lq_ctx*, int64_t t);
int64_t lq_kernel_launches(const lq_ctx*);         /* number of kernels this context has launched so far         */

/* ---- boundary marshalling (LatticeStateNew::new state.rs:779-792; set_link_matrix :808-815) ----------------- */
int lq_links_upload(lq_ctx*, const double* aos, int64_t n_links);   /* LQ_E_SIZE if n_links mismatches */
int lq_links_download(lq_ctx*, double* aos, int64_t n_links);
/* Pipelined marshalling for a host that streams a batch of independent configurations through one context (same
 * layouts and reference calls as lq_links_upload / _download: LatticeStateNew::new, link_matrix(), state.rs:779-815).
 * upload_begin starts copying `aos` into a device staging buffer on a copy stream and returns at once (pinned host
 * memory makes the copy asynchronous); upload_commit makes the staged links the state's links, ordered after the copy.
 * download_begin snapshots the links into a second staging buffer and starts copying them to `aos` on another copy
 * stream.  lq_copies_wait blocks until every begun copy has finished: only then may the host arrays be reused / read.
 * One upload and one download in flight at a time (LQ_E_BADARG otherwise).  PCIe is idle while a trajectory computes and
 * full duplex, so the next input and the previous result travel behind the kernels (bench.py e2e leg). */
int lq_links_upload_begin(lq_ctx*, const double* aos, int64_t n_links);
int lq_links_upload_commit(lq_ctx*);
int lq_links_download_begin(lq_ctx*, double* aos, int64_t n_links);
int lq_copies_wait(lq_ctx*);
int lq_efield_upload(lq_ctx*, const double* aos, int64_t n_links);
int lq_efield_download(lq_ctx*, double* aos, int64_t n_links);
/* same, from / to DEVICE memory already in the reference AoS layout (no PCIe copy) */
int lq_links_upload_device(lq_ctx*, const double* d_aos, int64_t n_links);
int lq_links_download_device(lq_ctx*, double* d_aos, int64_t n_links);
int lq_efield_upload_device(lq_ctx*, const double* d_aos, int64_t n_links);
int lq_efield_download_device(lq_ctx*, double* d_aos, int64_t n_links);
int lq_links_set_cold(lq_ctx*);                    /* LatticeStateDefault::new_cold, state.rs:671-679   */
int lq_efield_set_zero(lq_ctx*);                   /* EField::new_cold, field.rs:1102-1106              */
/* LinkMatrix::new_determinist (field.rs:646-659): random_su3 per link from Philox stream (seed, counter, link) */
int lq_links_set_random(lq_ctx*, uint64_t seed, uint64_t counter);

/* ---- observables ------------------------------------------------------------------------------------------- */
/* sum_x sum_{i<j} Tr P_ij(x): numerator of average_trace_plaquette (field.rs:775-804; state.rs:115-117) */
int lq_plaquette_sum(lq_ctx*, double out_re_im[2]);
int lq_average_trace_plaquette(lq_ctx*, double out_re_im[2]);
int lq_hamiltonian_links(lq_ctx*, double* h);      /* state.rs:821-849  */
int lq_hamiltonian_efield(lq_ctx*, double* h);     /* state.rs:1370-1385 */
int lq_hamiltonian_total(lq_ctx*, double* h);      /* state.rs:229-231  */

/* field-strength observables on every site (n_sites * 18, reference AoS).  Signed directions: +(d+1) / -(d+1).
 * LinkMatrix::clover (field.rs:807-820), f_mu_nu (:825-835), magnetic_field (:851-874). */
int lq_clover(lq_ctx*, int sdir_i, int sdir_j, double* aos_out, int64_t n_sites);
int lq_f_mu_nu(lq_ctx*, int dir_i, int dir_j, double* aos_out, int64_t n_sites);
int lq_magnetic_field(lq_ctx*, int dir, double* aos_out, int64_t n_sites);

/* ---- molecular dynamics ------------------------------------------------------------------------------------ */
int lq_staples(lq_ctx*, double* aos_out, int64_t n_links);   /* staple(), monte_carlo/mod.rs:339-362 (parity/debug) */
int lq_force(lq_ctx*, double* aos_out, int64_t n_links);     /* derivative_e, state.rs:1420-1448 (dE/dt, no update) */
int lq_efield_step(lq_ctx*, double dt);            /* integrate_efield over the lattice, integrator/mod.rs:240-254 */
int lq_link_step(lq_ctx*, double dt, int use_exp); /* integrate_link, integrator/mod.rs:216-233 (use_exp=0: Euler)  */
int lq_integrate(lq_ctx*, int kind, double dt);    /* one composition; t += 1 except LQ_SYNC_LEAP                  */
/* simulate_symplectic_n, state.rs:470-492: n x (E dt/2, U dt, E dt/2) */
int lq_symplectic_n(lq_ctx*, double dt, int64_t n_steps);
/* simulate_using_leapfrog_n, state.rs:321-358: sync_leap, (n-1) x leap_leap, leap_sync */
int lq_leapfrog_n(lq_ctx*, double dt, int64_t n_steps);
int lq_reunitarize(lq_ctx*);                       /* normalize_link_matrices, state.rs:754-756; su3.rs:279-303 */
/* Integrator options beyond the crate's: compositions of the same two updates -- integrate_efield
 * (integrator/mod.rs:240-254) and integrate_link (:216-233) or its exponential form (use_exp: U <- exp(i dt E) U,
 * su3.rs:832-855, keeps the links in SU(3) and makes the step time-reversible).  The selection is what lq_md_n and
 * lq_hmc_trajectory run; (LQ_INTEGRATOR_SYMPLECTIC_EULER, -, 0) -- the default -- is lq_symplectic_n, the reference's
 * SymplecticEulerRayon::integrate_symplectic.  LQ_INTEGRATOR_OMELYAN: E(l dt) U(dt/2) E((1-2l) dt) U(dt/2) E(l dt),
 * 0 < l < 1/2 (second-order minimum norm: l = 0.1931833275037836). */
enum { LQ_INTEGRATOR_SYMPLECTIC_EULER = 0, LQ_INTEGRATOR_OMELYAN = 1 };
int lq_set_integrator(lq_ctx*, int kind, double lambda, int use_exp);
int lq_md_n(lq_ctx*, double dt, int64_t n_steps);  /* n steps of the selected integrator, t += n */

/* ---- momenta + Gauss law ----------------------------------------------------------------------------------- */
/* EField::new_determinist with Normal(0, sigma) (field.rs:1086-1099; state.rs:1097 sigma = 0.5/beta) */
int lq_momenta_refresh(lq_ctx*, uint64_t seed, uint64_t counter, double sigma);
int lq_gauss_field(lq_ctx*, double* aos_out, int64_t n_sites);  /* EField::gauss, field.rs:1174-1195 (n_sites*18) */
int lq_gauss_sum_div(lq_ctx*, double* out);        /* field.rs:1199-1220 */
int lq_gauss_project_step(lq_ctx*);                /* field.rs:1301-1337 */
int lq_gauss_project(lq_ctx*, int64_t max_steps, int64_t* steps_out);  /* field.rs:1265-1294 */

/* ---- local-update sweeps (even/odd checkerboard; visit order: for dir, for parity) --------------------------
 * Lattices with odd extents (accepted by the reference's sequential sweeps, lattice.rs:190-201) are swept in the colour
 * classes (boundary mask, parity): for dir, for mask, for parity -- bit d of the mask set iff ext[d] is odd and
 * x_d = ext[d] - 1 -- because two colours do not decouple a periodic ring of odd length. */
int lq_sweep_heatbath(lq_ctx*, uint64_t seed, uint64_t counter, double coupling_scale); /* heat_bath.rs:73-123 */
int lq_sweep_overrelax(lq_ctx*, int kind);                                              /* overrelaxation.rs    */
int lq_sweep_metropolis(lq_ctx*, uint64_t seed, uint64_t counter, double spread, int n_update, int64_t* n_accept,
                        double* sum_prob);                              /* metropolis_hastings_sweep.rs:126-174 */

/* MetropolisHastingsDeltaDiagnostic::next_element (metropolis_hastings.rs:374-417: ONE uniformly random link per call,
 * proposal orthonormalize(random_su3_close_to_unity(spread)) * U, accept w.p. min(1, exp(-dS))), batched: the n_hits
 * hits of a call sit on links of one (direction, colour) class drawn from the call's stream and each draws its site
 * from its own Philox stream (seed, counter, hit index), so they do not enter each other's staples; of two hits on the
 * same link the lower index is performed and the other dropped (n_performed <= n_hits).  n_hits = 1 is the reference's
 * call.  force_accept = 1 applies every proposal without the accept step (MetropolisHastings::potential_next_element,
 * metropolis_hastings.rs:96-118).  Single-rank contexts with even extents. */
int lq_metropolis_hits(lq_ctx*, uint64_t seed, uint64_t counter, double spread, int64_t n_hits, int force_accept,
                       int64_t* n_performed, int64_t* n_accept, double* sum_prob);

/* ---- HMC (hybrid_monte_carlo.rs:465-471, 573-613) ----------------------------------------------------------- */
int lq_snapshot(lq_ctx*);                          /* keep a copy of (U, E, t) on the device ...                 */
int lq_restore(lq_ctx*);                           /* ... and go back to it (LQ_E_NOSNAPSHOT if there is none)   */
/* refresh (unless use_current_e) -> optional Gauss projection -> H_old -> n symplectic steps -> H_new ->
 * Bernoulli(clamp(exp(H_old-H_new),0,1)) from Philox stream (seed, counter, 0xFFFFFFFFFE); on reject the links
 * (and t) are restored from the trajectory's own device copy (not the lq_snapshot buffers).  On a decomposed context
 * with lq_set_comm registered the energies are global sums and every rank takes the same decision; without callbacks
 * they are rank-local partial sums and the decision is rank-local -- do not use that configuration for production. */
int lq_hmc_trajectory(lq_ctx*, double dt, int64_t n_steps, uint64_t seed, uint64_t counter, double sigma,
                      int use_current_e, int do_project, double* h_old, double* h_new, double* prob, int* accepted,
                      int64_t* gauss_steps);

/* ---- decomposed contexts: ghost layers and global sums ------------------------------------------------------
 * The library sequences every kernel itself; the caller's plumbing (torch.distributed over NCCL/NVLink, one
 * process per GPU) only moves bytes.  It registers two callbacks:
 *   halo_exchange(user, ctx, which): refresh the ghost layers of field `which` (0 links, 1 efield, 2 gauss field):
 *       for every decomposed direction in ascending order: lq_halo_pack both faces into caller-owned device
 *       buffers, send/recv them to the two neighbours, lq_halo_unpack into the ghost layers.
 *   allreduce_sum(user, vals, n): in-place sum of n host doubles over all ranks.
 * With both set, every reduction (plaquette, Hamiltonians, Gauss residual, Metropolis statistics) returns the
 * GLOBAL value on every rank and lq_hmc_trajectory takes the same accept decision everywhere. */
typedef struct lq_comm {
  void* user;
  int (*halo_exchange)(void* user, lq_ctx* ctx, int which);
  int (*allreduce_sum)(void* user, double* vals, int n);
  /* optional (may be NULL): in-place sum of n doubles in DEVICE memory, enqueued on the context stream (e.g. one
   * ncclAllReduce); when present every reduction costs a single host synchronisation */
  int (*allreduce_sum_device)(void* user, double* d_vals, int n);
} lq_comm;
int lq_set_comm(lq_ctx*, const lq_comm* comm);
/* run all context work on a caller-owned cudaStream_t (e.g. torch's current stream) instead of the private one */
int lq_set_stream(lq_ctx*, void* cuda_stream);
int lq_is_decomposed(const lq_ctx*, int dir);     /* 1 if `dir` carries ghost layers */
int lq_halo_bytes(lq_ctx*, int which, int dir, int64_t* bytes);
/* side: 0 = low face, 1 = high face.  pack reads the interior boundary slice, unpack writes the ghost slice. */
int lq_halo_pack(lq_ctx*, int which, int dir, int side, void* d_buf, int64_t bytes);
int lq_halo_unpack(lq_ctx*, int which, int dir, int side, const void* d_buf, int64_t bytes);
int lq_halo_invalidate(lq_ctx*, int which);
/* Peer-to-peer transport (preferred on one NVLink node): each rank exports CUDA-IPC handles of its field buffers
 * (9 x 64 bytes: U U2 E E2 G G2 T T2 flags), the caller all-gathers them and attaches the neighbours' handles; from then
 * on every ghost refresh is done by the library's own kernels writing the boundary slices straight into the
 * neighbours' ghost layers over NVLink, ordered by release/acquire flags in peer memory (no pack buffers, no NCCL,
 * no host round trip).  `offsets` = n_neighbors x D entries in {-1,0,+1}: the list must be identical on every rank
 * and closed under negation; peer_index[k] selects which of the n_peers opened handle sets neighbour k lives in
 * (two neighbours may be the same rank).  allreduce_sum of lq_set_comm is still used for the global sums. */
int lq_p2p_export(lq_ctx*, void* handles_out, int64_t bytes /* 9 * 64 */);
int lq_p2p_attach(lq_ctx*, int n_peers, const void* peer_handles, int n_neighbors, const int* offsets,
                  const int* peer_index);
int lq_p2p_enabled(const lq_ctx*);
int64_t lq_p2p_exchanges(const lq_ctx*);

/* ---- measurement ---------------------------------------------------------------------------------------------
 * Optional CUDA-event timing of kernel classes on the context stream (used by bench.py for the roofline line:
 * one event pair around every launch of the class, summed on query).  Off by default. */
enum {
  LQ_PROF_EFIELD_LINK_STEP = 0, /* fused force + E kick + link step (the dominant kernel of an MD trajectory) */
  LQ_PROF_EFIELD_STEP = 1,
  LQ_PROF_LINK_STEP = 2,
  LQ_PROF_PLAQUETTE = 3,
  LQ_PROF_GAUSS_FIELD = 4,
  LQ_PROF_GAUSS_STEP = 5,
  LQ_PROF_HEATBATH = 6,
  LQ_PROF_OVERRELAX = 7,
  LQ_PROF_METROPOLIS = 8,
  LQ_PROF_REUNITARIZE = 9,
  LQ_PROF_MOMENTA = 10,
  LQ_PROF_EFIELD_ENERGY = 11,
  LQ_PROF_GAUSS_DIV = 12,
  LQ_PROF_COPY = 13             /* device-to-device copies of the HMC reject path */
};
int lq_profile_enable(lq_ctx*, int on);
int lq_profile_reset(lq_ctx*);
int lq_profile_get(lq_ctx*, int kernel_class, int64_t* launches, double* total_ms);
/* The two ceilings the rooflines are quoted against, measured on this device in the caller's own run: f64 FMA issue
 * rate (TFLOP/s; eight independent chains per thread on every SM, ~20 ms) and a streaming copy of the link buffer
 * (GB/s read + write).  Either pointer may be NULL. */
int lq_measure_peaks(lq_ctx*, double* fp64_tflops, double* copy_gbs);

#ifdef __cplusplus
}
#endif
#endif /* LQCD_B200_H */
