//! Raw bindings of `include/lqcd_b200.h` (AUTHORED, NOT COMPILED here: no rustc/cargo in the image).
//! One line per C entry point; the doc comment names the reference loop it replaces.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_double, c_int, c_void};

#[repr(C)]
pub struct lq_ctx {
    _private: [u8; 0],
}

pub const LQ_OK: c_int = 0;
pub const LQ_E_BADARG: c_int = -1;
pub const LQ_E_SIZE: c_int = -2;
pub const LQ_E_CUDA: c_int = -3;
pub const LQ_E_COMM: c_int = -4;
pub const LQ_E_ODD_EXTENT: c_int = -5;
pub const LQ_E_GAUSS_DIVERGED: c_int = -6;
pub const LQ_E_ZERO_STEPS: c_int = -7;
pub const LQ_E_NOSNAPSHOT: c_int = -8;
pub const LQ_E_NODEVICE: c_int = -9;

pub const LQ_SYNC_SYNC: c_int = 0;
pub const LQ_LEAP_LEAP: c_int = 1;
pub const LQ_SYNC_LEAP: c_int = 2;
pub const LQ_LEAP_SYNC: c_int = 3;
pub const LQ_SYMPLECTIC: c_int = 4;
pub const LQ_OR_ROTATION: c_int = 0;
pub const LQ_OR_REVERSE: c_int = 1;

extern "C" {
    pub fn lq_strerror(code: c_int) -> *const c_char;
    pub fn lq_last_cuda_error() -> *const c_char;
    /// LatticeCyclic::new + state storage (lattice.rs:190-201; state.rs:655-659, 1048-1062)
    pub fn lq_ctx_create(out: *mut *mut lq_ctx, device: c_int, d: c_int, extent: *const i64, a: c_double,
                         beta: c_double, ca: c_double) -> c_int;
    /// Clone (state.rs:292-295)
    pub fn lq_ctx_clone(src: *const lq_ctx, out: *mut *mut lq_ctx) -> c_int;
    pub fn lq_ctx_destroy(c: *mut lq_ctx) -> c_int;
    pub fn lq_num_links(c: *const lq_ctx) -> i64;
    pub fn lq_t(c: *const lq_ctx) -> i64;
    pub fn lq_set_t(c: *mut lq_ctx, t: i64) -> c_int;
    /// LatticeStateNew::new / set_link_matrix (state.rs:779-815)
    pub fn lq_links_upload(c: *mut lq_ctx, aos: *const c_double, n_links: i64) -> c_int;
    pub fn lq_links_download(c: *mut lq_ctx, aos: *mut c_double, n_links: i64) -> c_int;
    // pipelined marshalling (a batch of configurations streamed through one context)
    pub fn lq_links_upload_begin(c: *mut lq_ctx, aos: *const c_double, n_links: i64) -> c_int;
    pub fn lq_links_upload_commit(c: *mut lq_ctx) -> c_int;
    pub fn lq_links_download_begin(c: *mut lq_ctx, aos: *mut c_double, n_links: i64) -> c_int;
    pub fn lq_copies_wait(c: *mut lq_ctx) -> c_int;
    pub fn lq_efield_upload(c: *mut lq_ctx, aos: *const c_double, n_links: i64) -> c_int;
    pub fn lq_efield_download(c: *mut lq_ctx, aos: *mut c_double, n_links: i64) -> c_int;
    pub fn lq_links_set_cold(c: *mut lq_ctx) -> c_int;
    pub fn lq_efield_set_zero(c: *mut lq_ctx) -> c_int;
    pub fn lq_links_set_random(c: *mut lq_ctx, seed: u64, counter: u64) -> c_int;
    /// average_trace_plaquette (field.rs:775-804)
    pub fn lq_average_trace_plaquette(c: *mut lq_ctx, out_re_im: *mut c_double) -> c_int;
    /// hamiltonian_links / _efield / _total (state.rs:821-849, 1370-1385, 229-231)
    pub fn lq_hamiltonian_links(c: *mut lq_ctx, h: *mut c_double) -> c_int;
    pub fn lq_hamiltonian_efield(c: *mut lq_ctx, h: *mut c_double) -> c_int;
    pub fn lq_hamiltonian_total(c: *mut lq_ctx, h: *mut c_double) -> c_int;
    /// SymplecticEulerRayon compositions (symplectic_euler_rayon.rs:120-252)
    pub fn lq_integrate(c: *mut lq_ctx, kind: c_int, dt: c_double) -> c_int;
    /// simulate_symplectic_n (state.rs:470-492)
    pub fn lq_symplectic_n(c: *mut lq_ctx, dt: c_double, n: i64) -> c_int;
    /// integrator options beyond the crate's (Omelyan, exponential link update): what lq_md_n / lq_hmc_trajectory run
    pub fn lq_set_integrator(c: *mut lq_ctx, kind: c_int, lambda: c_double, use_exp: c_int) -> c_int;
    pub fn lq_md_n(c: *mut lq_ctx, dt: c_double, n: i64) -> c_int;
    /// normalize_link_matrices (state.rs:754-756)
    pub fn lq_reunitarize(c: *mut lq_ctx) -> c_int;
    /// EField::new_determinist + project_to_gauss (field.rs:1086-1099, 1265-1294)
    pub fn lq_momenta_refresh(c: *mut lq_ctx, seed: u64, counter: u64, sigma: c_double) -> c_int;
    pub fn lq_gauss_project(c: *mut lq_ctx, max_steps: i64, steps_out: *mut i64) -> c_int;
    /// HeatBathSweep / OverrelaxationSweep* / MetropolisHastingsSweep (monte_carlo/*.rs)
    pub fn lq_sweep_heatbath(c: *mut lq_ctx, seed: u64, counter: u64, coupling_scale: c_double) -> c_int;
    pub fn lq_sweep_overrelax(c: *mut lq_ctx, kind: c_int) -> c_int;
    pub fn lq_sweep_metropolis(c: *mut lq_ctx, seed: u64, counter: u64, spread: c_double, n_update: c_int,
                               n_accept: *mut i64, sum_prob: *mut c_double) -> c_int;
    /// MetropolisHastingsDeltaDiagnostic::next_element (metropolis_hastings.rs:374-417), `n_hits` hits per call;
    /// `force_accept` = 1: MetropolisHastings::potential_next_element (metropolis_hastings.rs:96-118)
    pub fn lq_metropolis_hits(c: *mut lq_ctx, seed: u64, counter: u64, spread: c_double, n_hits: i64,
                              force_accept: c_int, n_performed: *mut i64, n_accept: *mut i64,
                              sum_prob: *mut c_double) -> c_int;
    /// HybridMonteCarloDiagnostic::next_element (hybrid_monte_carlo.rs:465-471, 573-613)
    pub fn lq_hmc_trajectory(c: *mut lq_ctx, dt: c_double, n_steps: i64, seed: u64, counter: u64, sigma: c_double,
                             use_current_e: c_int, do_project: c_int, h_old: *mut c_double, h_new: *mut c_double,
                             prob: *mut c_double, accepted: *mut c_int, gauss_steps: *mut i64) -> c_int;
    pub fn lq_sync(c: *mut lq_ctx) -> c_int;
}

#[allow(unused)]
pub(crate) fn _unused(_: *mut c_void) {}
