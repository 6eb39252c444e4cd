//! `lattice-qcd-b200`: GPU-resident states for `lattice_qcd_rs` 0.2.1.
//!
//! AUTHORED, NOT COMPILED in this environment (no rustc/cargo).  The tested twin of this file is
//! `lattice_qcd_rs_b200/state.py`; both sit on the same C ABI (`include/lqcd_b200.h`).
//!
//! Why a separate crate: `lattice_qcd_rs` is `#![forbid(unsafe_code)]` (lib.rs:31) and FFI needs `unsafe`; and the
//! orphan rule forbids `impl LatticeStateWithEField for LatticeStateEFSyncDefault<GpuState>` from outside the crate
//! (README.md:92-94 suggests it; E0117 rejects it).  So this crate defines its own synchronous state type that
//! implements the reference's traits, re-uses the public generic `SimulationStateLeap<State, D>` (state.rs:856-861)
//! for leap-frog states, and ships `MonteCarlo` impls with the constructor arguments and getters of the reference's
//! `HybridMonteCarloDiagnostic` / `HeatBathSweep` / `OverrelaxationSweep*` / `MetropolisHastingsSweep` /
//! `MetropolisHastingsDeltaDiagnostic`, plus a `MonteCarloDefault` impl (`MetropolisHastingsCuda`).  The reference's
//! generic combinators need nothing from this crate: `McWrapper<MCD, State, Rng, D>` (monte_carlo/mod.rs:210-293),
//! `HybridMethodVec` and `HybridMethodCouple/Triple/...` (hybrid.rs:248-446) are generic over any
//! `State: LatticeState<D>` and any `MonteCarlo<State, D>` / `MonteCarloDefault<State, D>`, so they compose the types
//! below as they are (the Python twin re-implements them only because it cannot import the crate).
//!
//! `LatticeState::link_matrix(&self) -> &LinkMatrix` hands out a HOST borrow (state.rs:74): the state keeps a lazily
//! filled host mirror in a `OnceLock`; every mutation goes through `&mut self` or consumes `self`, where the mirror
//! is dropped, so no interior mutability beyond `OnceLock` is needed.
mod ffi;

use std::sync::OnceLock;

use lattice_qcd_rs::{
    error::{MultiIntegrationError, StateInitializationError},
    field::{EField, LinkMatrix, Su3Adjoint},
    integrator::SymplecticIntegrator,
    lattice::{LatticeCyclic, LatticeLinkCanonical, LatticePoint},
    simulation::{
        monte_carlo::MonteCarlo, LatticeState, LatticeStateDefault, LatticeStateEFSyncDefault, LatticeStateNew,
        LatticeStateWithEField, LatticeStateWithEFieldNew, SimulationStateLeap, SimulationStateSynchronous,
    },
    CMatrix3, Complex, Real,
};
use nalgebra::SVector;

/// C error code -> the reference's error enums (error.rs:93-133).  Nothing unwinds across the ABI.
#[derive(Debug, Clone, PartialEq, Eq)]
pub enum CudaError {
    BadArgument,
    IncompatibleSize,
    Cuda(String),
    OddExtent,
    GaussProjection,
    ZeroSteps,
    NoDevice,
    Other(i32),
}

fn check(rc: i32) -> Result<(), CudaError> {
    match rc {
        ffi::LQ_OK => Ok(()),
        ffi::LQ_E_BADARG => Err(CudaError::BadArgument),
        ffi::LQ_E_SIZE => Err(CudaError::IncompatibleSize),
        ffi::LQ_E_CUDA => Err(CudaError::Cuda(unsafe {
            std::ffi::CStr::from_ptr(ffi::lq_last_cuda_error()).to_string_lossy().into_owned()
        })),
        ffi::LQ_E_ODD_EXTENT => Err(CudaError::OddExtent),
        ffi::LQ_E_GAUSS_DIVERGED => Err(CudaError::GaussProjection),
        ffi::LQ_E_ZERO_STEPS => Err(CudaError::ZeroSteps),
        ffi::LQ_E_NODEVICE => Err(CudaError::NoDevice),
        e => Err(CudaError::Other(e)),
    }
}

impl From<CudaError> for StateInitializationError {
    fn from(e: CudaError) -> Self {
        match e {
            CudaError::GaussProjection => StateInitializationError::GaussProjectionError,
            _ => StateInitializationError::IncompatibleSize,
        }
    }
}

/// Owner of one `lq_ctx` (device buffers of links, E-field, scratch).
struct Ctx(*mut ffi::lq_ctx);
// one context = one stream, used from one thread at a time (the reference moves states through `next_element`)
unsafe impl Send for Ctx {}
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { ffi::lq_ctx_destroy(self.0) };
    }
}
impl Ctx {
    fn new<const D: usize>(l: &LatticeCyclic<D>, beta: Real) -> Result<Self, CudaError> {
        let ext = [l.dim() as i64; D];
        let mut p = std::ptr::null_mut();
        check(unsafe { ffi::lq_ctx_create(&mut p, 0, D as i32, ext.as_ptr(), l.size(), beta, 3.0) })?;
        Ok(Self(p))
    }
    fn try_clone(&self) -> Result<Self, CudaError> {
        let mut p = std::ptr::null_mut();
        check(unsafe { ffi::lq_ctx_clone(self.0, &mut p) })?;
        Ok(Self(p))
    }
    fn n_links(&self) -> usize {
        unsafe { ffi::lq_num_links(self.0) as usize }
    }
    /// `Vec<Matrix3<Complex<f64>>>` is 18 contiguous f64 per link, column-major (nalgebra ArrayStorage): the ABI's
    /// AoS layout, so the Vec's buffer is passed as is.
    fn upload_links(&self, m: &LinkMatrix) -> Result<(), CudaError> {
        check(unsafe { ffi::lq_links_upload(self.0, m.as_slice().as_ptr() as *const f64, m.len() as i64) })
    }
    fn download_links(&self) -> LinkMatrix {
        let mut v = vec![CMatrix3::zeros(); self.n_links()];
        check(unsafe { ffi::lq_links_download(self.0, v.as_mut_ptr() as *mut f64, v.len() as i64) })
            .expect("device -> host copy of the links failed");
        LinkMatrix::new(v)
    }
    fn upload_e<const D: usize>(&self, e: &EField<D>) -> Result<(), CudaError> {
        check(unsafe { ffi::lq_efield_upload(self.0, e.as_slice().as_ptr() as *const f64, (e.len() * D) as i64) })
    }
    fn download_e<const D: usize>(&self) -> EField<D> {
        let mut v = vec![SVector::<Su3Adjoint, D>::from_element(Su3Adjoint::default()); self.n_links() / D];
        check(unsafe { ffi::lq_efield_download(self.0, v.as_mut_ptr() as *mut f64, (v.len() * D) as i64) })
            .expect("device -> host copy of the E field failed");
        EField::new(v)
    }
    fn scalar(&self, f: unsafe extern "C" fn(*mut ffi::lq_ctx, *mut f64) -> i32) -> Real {
        let mut h = 0.0;
        check(unsafe { f(self.0, &mut h) }).expect("device reduction failed");
        h
    }
}

/// Pure-gauge state with device-resident links: the GPU twin of `LatticeStateDefault<D>` (state.rs:655-659).
pub struct LatticeStateCuda<const D: usize> {
    lattice: LatticeCyclic<D>,
    beta: Real,
    ctx: Ctx,
    host_links: OnceLock<LinkMatrix>,
}

impl<const D: usize> LatticeStateCuda<D> {
    /// state.rs:671-679
    pub fn new_cold(size: Real, beta: Real, number_of_points: usize) -> Result<Self, StateInitializationError> {
        let lattice = LatticeCyclic::new(size, number_of_points)?;
        let ctx = Ctx::new(&lattice, beta)?;
        check(unsafe { ffi::lq_links_set_cold(ctx.0) })?;
        Ok(Self { lattice, beta, ctx, host_links: OnceLock::new() })
    }
    /// state.rs:706-715; the host rng only seeds the device Philox streams
    pub fn new_determinist(size: Real, beta: Real, number_of_points: usize, rng: &mut impl rand::Rng)
                           -> Result<Self, StateInitializationError> {
        let lattice = LatticeCyclic::new(size, number_of_points)?;
        let ctx = Ctx::new(&lattice, beta)?;
        check(unsafe { ffi::lq_links_set_random(ctx.0, rng.next_u64(), rng.next_u64() >> 8) })?;
        Ok(Self { lattice, beta, ctx, host_links: OnceLock::new() })
    }
    /// state.rs:754-756
    pub fn normalize_link_matrices(&mut self) {
        check(unsafe { ffi::lq_reunitarize(self.ctx.0) }).expect("reunitarize");
        self.host_links = OnceLock::new();
    }
    /// Hand the configuration back to the CPU types of the reference (serde, observables, ...).
    pub fn to_default(&self) -> Result<LatticeStateDefault<D>, StateInitializationError> {
        LatticeStateDefault::new(self.lattice.clone(), self.beta, self.ctx.download_links())
    }
    /// Streaming a batch of configurations through one device state (include/lqcd_b200.h, "pipelined marshalling"):
    /// start copying the NEXT configuration to the device while the current one is still being worked on.  `next` must
    /// stay alive and unmodified until `commit_links` has been followed by `wait_copies` (the borrow enforces the first
    /// half; keep the owner around for the second).  Pinned host memory makes the copy run behind the kernels.
    pub fn begin_set_link_matrix(&mut self, next: &LinkMatrix) -> Result<(), CudaError> {
        assert_eq!(next.len(), self.ctx.n_links(), "link matrix of the wrong size (state.rs:808-815 panics too)");
        check(unsafe { ffi::lq_links_upload_begin(self.ctx.0, next.as_slice().as_ptr() as *const f64, next.len() as i64) })
    }
    /// The links begun with `begin_set_link_matrix` become the state's links (LatticeState::set_link_matrix, state.rs:808-815).
    pub fn commit_links(&mut self) -> Result<(), CudaError> {
        self.host_links = OnceLock::new();
        check(unsafe { ffi::lq_links_upload_commit(self.ctx.0) })
    }
    /// Start copying the current links into `out` (link_matrix() without blocking); valid after `wait_copies`.
    pub fn begin_download_links(&self, out: &mut [CMatrix3]) -> Result<(), CudaError> {
        assert_eq!(out.len(), self.ctx.n_links());
        check(unsafe { ffi::lq_links_download_begin(self.ctx.0, out.as_mut_ptr() as *mut f64, out.len() as i64) })
    }
    /// Block until every begun copy has finished.
    pub fn wait_copies(&self) -> Result<(), CudaError> {
        check(unsafe { ffi::lq_copies_wait(self.ctx.0) })
    }
}

impl<const D: usize> Clone for LatticeStateCuda<D> {
    fn clone(&self) -> Self {
        Self { lattice: self.lattice.clone(), beta: self.beta, ctx: self.ctx.try_clone().expect("device clone"),
               host_links: OnceLock::new() }
    }
}

impl<const D: usize> LatticeState<D> for LatticeStateCuda<D> {
    const CA: Real = 3_f64;
    fn link_matrix(&self) -> &LinkMatrix {
        self.host_links.get_or_init(|| self.ctx.download_links())
    }
    fn set_link_matrix(&mut self, link_matrix: LinkMatrix) {
        if self.lattice.number_of_canonical_links_space() != link_matrix.len() {
            panic!("Link matrices are not of the correct size"); // state.rs:808-815
        }
        self.ctx.upload_links(&link_matrix).expect("upload");
        self.host_links = OnceLock::new();
    }
    fn lattice(&self) -> &LatticeCyclic<D> {
        &self.lattice
    }
    fn beta(&self) -> Real {
        self.beta
    }
    fn hamiltonian_links(&self) -> Real {
        self.ctx.scalar(ffi::lq_hamiltonian_links)
    }
    fn average_trace_plaquette(&self) -> Option<Complex> {
        let mut v = [0_f64; 2];
        check(unsafe { ffi::lq_average_trace_plaquette(self.ctx.0, v.as_mut_ptr()) }).ok()?;
        Some(Complex::new(v[0], v[1]))
    }
}

impl<const D: usize> LatticeStateNew<D> for LatticeStateCuda<D> {
    type Error = StateInitializationError;
    fn new(lattice: LatticeCyclic<D>, beta: Real, link_matrix: LinkMatrix) -> Result<Self, Self::Error> {
        if !lattice.has_compatible_length_links(&link_matrix) {
            return Err(StateInitializationError::IncompatibleSize); // state.rs:784-786
        }
        let ctx = Ctx::new(&lattice, beta)?;
        ctx.upload_links(&link_matrix)?;
        Ok(Self { lattice, beta, ctx, host_links: OnceLock::new() })
    }
}

/// Links + E-field + step counter on the device: the GPU twin of
/// `LatticeStateEFSyncDefault<LatticeStateDefault<D>, D>` (state.rs:1048-1062).
pub struct LatticeStateEFSyncCuda<const D: usize> {
    inner: LatticeStateCuda<D>,
    host_e: OnceLock<EField<D>>,
}

impl<const D: usize> Clone for LatticeStateEFSyncCuda<D> {
    fn clone(&self) -> Self {
        Self { inner: self.inner.clone(), host_e: OnceLock::new() }
    }
}

impl<const D: usize> LatticeStateEFSyncCuda<D> {
    /// state.rs:1093-1108: Normal(0, 0.5/beta) momenta, Gauss-projected
    pub fn new_random_e_state(lattice_state: LatticeStateCuda<D>, rng: &mut impl rand::Rng) -> Self {
        let mut s = Self { inner: lattice_state, host_e: OnceLock::new() };
        s.reset_e_field(rng).expect("Projection to gauss failed");
        s
    }
    /// state.rs:1111-1121
    pub fn new_e_cold(lattice_state: LatticeStateCuda<D>) -> Self {
        check(unsafe { ffi::lq_efield_set_zero(lattice_state.ctx.0) }).expect("memset");
        Self { inner: lattice_state, host_e: OnceLock::new() }
    }
    /// state.rs:1071-1076
    pub fn state_owned(self) -> LatticeStateCuda<D> {
        self.inner
    }
    fn ctx(&self) -> *mut ffi::lq_ctx {
        self.inner.ctx.0
    }
    fn invalidate(&mut self) {
        self.inner.host_links = OnceLock::new();
        self.host_e = OnceLock::new();
    }
}

impl<const D: usize> LatticeState<D> for LatticeStateEFSyncCuda<D> {
    const CA: Real = 3_f64;
    fn link_matrix(&self) -> &LinkMatrix {
        self.inner.link_matrix()
    }
    fn set_link_matrix(&mut self, link_matrix: LinkMatrix) {
        self.inner.set_link_matrix(link_matrix)
    }
    fn lattice(&self) -> &LatticeCyclic<D> {
        self.inner.lattice()
    }
    fn beta(&self) -> Real {
        self.inner.beta()
    }
    fn hamiltonian_links(&self) -> Real {
        self.inner.hamiltonian_links()
    }
    fn average_trace_plaquette(&self) -> Option<Complex> {
        self.inner.average_trace_plaquette()
    }
}

impl<const D: usize> LatticeStateWithEField<D> for LatticeStateEFSyncCuda<D> {
    /// state.rs:174-189, on the device
    fn reset_e_field<Rng>(&mut self, rng: &mut Rng) -> Result<(), StateInitializationError>
    where
        Rng: rand::Rng + ?Sized,
    {
        rand_distr::Normal::new(0_f64, 0.5_f64 / self.beta())?; // same parameter validation as the reference
        check(unsafe { ffi::lq_momenta_refresh(self.ctx(), rng.next_u64(), rng.next_u64() >> 8, 0.5 / self.beta()) })?;
        let mut steps = 0_i64;
        check(unsafe { ffi::lq_gauss_project(self.ctx(), 0, &mut steps) })?;
        self.host_e = OnceLock::new();
        Ok(())
    }
    fn e_field(&self) -> &EField<D> {
        self.host_e.get_or_init(|| self.inner.ctx.download_e::<D>())
    }
    fn set_e_field(&mut self, e_field: EField<D>) {
        if self.lattice().number_of_points() != e_field.len() {
            panic!("e_field is not of the correct size"); // state.rs:1394-1399
        }
        self.inner.ctx.upload_e(&e_field).expect("upload");
        self.host_e = OnceLock::new();
    }
    fn t(&self) -> usize {
        unsafe { ffi::lq_t(self.ctx()) as usize }
    }
    /// Host per-element callbacks used only by the CPU integrators; the GPU integrator never calls them.  They
    /// delegate to the reference's own formulas so that CPU integrators still work on a downloaded state.
    fn derivative_u(link: &LatticeLinkCanonical<D>, link_matrix: &LinkMatrix, e_field: &EField<D>,
                    lattice: &LatticeCyclic<D>) -> Option<CMatrix3> {
        <LatticeStateEFSyncDefault<LatticeStateDefault<D>, D> as LatticeStateWithEField<D>>::derivative_u(
            link, link_matrix, e_field, lattice)
    }
    fn derivative_e(point: &LatticePoint<D>, link_matrix: &LinkMatrix, e_field: &EField<D>,
                    lattice: &LatticeCyclic<D>) -> Option<SVector<Su3Adjoint, D>> {
        <LatticeStateEFSyncDefault<LatticeStateDefault<D>, D> as LatticeStateWithEField<D>>::derivative_e(
            point, link_matrix, e_field, lattice)
    }
    fn hamiltonian_efield(&self) -> Real {
        self.inner.ctx.scalar(ffi::lq_hamiltonian_efield)
    }
    fn hamiltonian_total(&self) -> Real {
        self.inner.ctx.scalar(ffi::lq_hamiltonian_total)
    }
}

impl<const D: usize> LatticeStateWithEFieldNew<D> for LatticeStateEFSyncCuda<D> {
    type Error = StateInitializationError;
    fn new(lattice: LatticeCyclic<D>, beta: Real, e_field: EField<D>, link_matrix: LinkMatrix, t: usize)
           -> Result<Self, Self::Error> {
        if !lattice.has_compatible_length(&link_matrix, &e_field) {
            return Err(StateInitializationError::IncompatibleSize);
        }
        let inner = LatticeStateCuda::new(lattice, beta, link_matrix)?;
        inner.ctx.upload_e(&e_field)?;
        check(unsafe { ffi::lq_set_t(inner.ctx.0, t as i64) })?;
        Ok(Self { inner, host_e: OnceLock::new() })
    }
}

impl<const D: usize> SimulationStateSynchronous<D> for LatticeStateEFSyncCuda<D> {}

/// `SymplecticIntegrator` (integrator/mod.rs:93-208) with the arithmetic of `SymplecticEulerRayon`
/// (symplectic_euler_rayon.rs:120-252) on the device.  `&self` in, fresh state out, as in the reference.
#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]
pub struct SymplecticEulerCuda;

impl SymplecticEulerCuda {
    pub const fn new() -> Self {
        Self
    }
    fn step<const D: usize>(s: &LatticeStateEFSyncCuda<D>, kind: i32, dt: Real)
                            -> Result<LatticeStateEFSyncCuda<D>, CudaError> {
        let mut n = s.clone();
        check(unsafe { ffi::lq_integrate(n.ctx(), kind, dt) })?;
        n.invalidate();
        Ok(n)
    }
}

type Leap<const D: usize> = SimulationStateLeap<LatticeStateEFSyncCuda<D>, D>;

impl<const D: usize> SymplecticIntegrator<LatticeStateEFSyncCuda<D>, Leap<D>, D> for SymplecticEulerCuda {
    type Error = CudaError;
    fn integrate_sync_sync(&self, l: &LatticeStateEFSyncCuda<D>, dt: Real) -> Result<LatticeStateEFSyncCuda<D>, CudaError> {
        Self::step(l, ffi::LQ_SYNC_SYNC, dt)
    }
    fn integrate_leap_leap(&self, l: &Leap<D>, dt: Real) -> Result<Leap<D>, CudaError> {
        Ok(Leap::new_from_state(Self::step(l.as_ref(), ffi::LQ_LEAP_LEAP, dt)?))
    }
    fn integrate_sync_leap(&self, l: &LatticeStateEFSyncCuda<D>, dt: Real) -> Result<Leap<D>, CudaError> {
        Ok(Leap::new_from_state(Self::step(l, ffi::LQ_SYNC_LEAP, dt)?))
    }
    fn integrate_leap_sync(&self, l: &Leap<D>, dt: Real) -> Result<LatticeStateEFSyncCuda<D>, CudaError> {
        Self::step(l.as_ref(), ffi::LQ_LEAP_SYNC, dt)
    }
    fn integrate_symplectic(&self, l: &LatticeStateEFSyncCuda<D>, dt: Real) -> Result<LatticeStateEFSyncCuda<D>, CudaError> {
        Self::step(l, ffi::LQ_SYMPLECTIC, dt)
    }
}

/// An integrator the crate does not have, offered through the same `SymplecticIntegrator` trait: second-order
/// minimum-norm (Omelyan) steps `E(l dt) U(dt/2) E((1-2l) dt) U(dt/2) E(l dt)` built from the crate's own two updates
/// (`integrate_efield`, integrator/mod.rs:240-254; `integrate_link`, :216-233), with `use_exp` replacing the Euler link
/// update by `U <- exp(i dt E) U` (su3.rs:832-855).  The leap-frog half-step compositions are those of
/// `SymplecticEulerCuda`.
#[derive(Clone, Copy, Debug, PartialEq)]
pub struct OmelyanCuda {
    pub lambda: Real,
    pub use_exp: bool,
}

impl Default for OmelyanCuda {
    fn default() -> Self {
        Self { lambda: 0.193_183_327_503_783_6, use_exp: true }
    }
}

impl<const D: usize> SymplecticIntegrator<LatticeStateEFSyncCuda<D>, Leap<D>, D> for OmelyanCuda {
    type Error = CudaError;
    fn integrate_sync_sync(&self, l: &LatticeStateEFSyncCuda<D>, dt: Real) -> Result<LatticeStateEFSyncCuda<D>, CudaError> {
        SymplecticEulerCuda.integrate_sync_sync(l, dt)
    }
    fn integrate_leap_leap(&self, l: &Leap<D>, dt: Real) -> Result<Leap<D>, CudaError> {
        SymplecticEulerCuda.integrate_leap_leap(l, dt)
    }
    fn integrate_sync_leap(&self, l: &LatticeStateEFSyncCuda<D>, dt: Real) -> Result<Leap<D>, CudaError> {
        SymplecticEulerCuda.integrate_sync_leap(l, dt)
    }
    fn integrate_leap_sync(&self, l: &Leap<D>, dt: Real) -> Result<LatticeStateEFSyncCuda<D>, CudaError> {
        SymplecticEulerCuda.integrate_leap_sync(l, dt)
    }
    fn integrate_symplectic(&self, l: &LatticeStateEFSyncCuda<D>, dt: Real) -> Result<LatticeStateEFSyncCuda<D>, CudaError> {
        let mut n = l.clone();
        check(unsafe { ffi::lq_set_integrator(n.ctx(), 1, self.lambda, self.use_exp as i32) })?;
        let rc = unsafe { ffi::lq_md_n(n.ctx(), dt, 1) };
        check(unsafe { ffi::lq_set_integrator(n.ctx(), 0, self.lambda, 0) })?;
        check(rc)?;
        n.invalidate();
        Ok(n)
    }
}

/// GPU twin of `HybridMonteCarloDiagnostic` (hybrid_monte_carlo.rs:316-471): same constructor arguments and getters.
pub struct HybridMonteCarloCuda<Rng: rand::Rng> {
    delta_t: Real,
    number_of_steps: usize,
    rng: Rng,
    prob_replace_last: Real,
    has_replace_last: bool,
}

impl<Rng: rand::Rng> HybridMonteCarloCuda<Rng> {
    pub const fn new(delta_t: Real, number_of_steps: usize, _integrator: SymplecticEulerCuda, rng: Rng) -> Self {
        Self { delta_t, number_of_steps, rng, prob_replace_last: 0.0, has_replace_last: false }
    }
    pub const fn prob_replace_last(&self) -> Real {
        self.prob_replace_last
    }
    pub const fn has_replace_last(&self) -> bool {
        self.has_replace_last
    }
    pub fn rng_owned(self) -> Rng {
        self.rng
    }
}

impl<Rng: rand::Rng, const D: usize> MonteCarlo<LatticeStateCuda<D>, D> for HybridMonteCarloCuda<Rng> {
    type Error = MultiIntegrationError<CudaError>;
    /// refresh + Gauss-project, n symplectic steps, accept with clamp(exp(H_old - H_new), 0, 1); the old links stay
    /// on the device for the reject path.  The state is moved in and out, so the update is in place.
    fn next_element(&mut self, mut state: LatticeStateCuda<D>) -> Result<LatticeStateCuda<D>, Self::Error> {
        if self.number_of_steps == 0 {
            return Err(MultiIntegrationError::ZeroIntegration);
        }
        let (mut h0, mut h1, mut p, mut acc, mut gs) = (0.0, 0.0, 0.0, 0, 0_i64);
        check(unsafe {
            ffi::lq_hmc_trajectory(state.ctx.0, self.delta_t, self.number_of_steps as i64, self.rng.next_u64(),
                                   self.rng.next_u64() >> 8, 0.5 / state.beta, 0, 1, &mut h0, &mut h1, &mut p,
                                   &mut acc, &mut gs)
        })
        .map_err(|e| MultiIntegrationError::IntegrationError(0, e))?;
        self.prob_replace_last = p;
        self.has_replace_last = acc != 0;
        state.host_links = OnceLock::new();
        Ok(state)
    }
}

macro_rules! sweep {
    ($(#[$doc:meta])* $name:ident, |$s:ident, $st:ident| $call:expr) => {
        $(#[$doc])*
        impl<Rng: rand::Rng, const D: usize> MonteCarlo<LatticeStateCuda<D>, D> for $name<Rng> {
            type Error = CudaError;
            fn next_element(&mut self, mut $st: LatticeStateCuda<D>) -> Result<LatticeStateCuda<D>, CudaError> {
                let $s = self;
                check(unsafe { $call })?;
                $st.host_links = OnceLock::new();
                Ok($st)
            }
        }
    };
}

/// heat_bath.rs:40-157 (even/odd checkerboard order; `coupling_scale = 1` restates heat_bath.rs:77)
pub struct HeatBathSweepCuda<Rng: rand::Rng> {
    pub rng: Rng,
    pub coupling_scale: Real,
}
impl<Rng: rand::Rng> HeatBathSweepCuda<Rng> {
    pub const fn new(rng: Rng) -> Self {
        Self { rng, coupling_scale: 1.0 }
    }
}
sweep!(HeatBathSweepCuda, |s, st| ffi::lq_sweep_heatbath(st.ctx.0, s.rng.next_u64(), s.rng.next_u64() >> 8,
                                                         s.coupling_scale));

/// overrelaxation.rs:58-184; `kind` = LQ_OR_ROTATION | LQ_OR_REVERSE
pub struct OverrelaxationSweepCuda<Rng: rand::Rng> {
    pub kind: i32,
    _rng: std::marker::PhantomData<Rng>,
}
impl OverrelaxationSweepCuda<rand::rngs::ThreadRng> {
    pub const fn rotation() -> Self {
        Self { kind: ffi::LQ_OR_ROTATION, _rng: std::marker::PhantomData }
    }
    pub const fn reverse() -> Self {
        Self { kind: ffi::LQ_OR_REVERSE, _rng: std::marker::PhantomData }
    }
}
sweep!(OverrelaxationSweepCuda, |s, st| ffi::lq_sweep_overrelax(st.ctx.0, s.kind));

/// metropolis_hastings_sweep.rs:41-174
pub struct MetropolisHastingsSweepCuda<Rng: rand::Rng> {
    number_of_update: usize,
    spread: Real,
    number_replace_last: usize,
    prob_replace_mean: Real,
    rng: Rng,
}
impl<Rng: rand::Rng> MetropolisHastingsSweepCuda<Rng> {
    /// `None` for invalid parameters (metropolis_hastings_sweep.rs:73-80)
    pub fn new(number_of_update: usize, spread: Real, rng: Rng) -> Option<Self> {
        if number_of_update == 0 || spread <= 0_f64 || spread >= 1_f64 {
            return None;
        }
        Some(Self { number_of_update, spread, number_replace_last: 0, prob_replace_mean: 0.0, rng })
    }
    pub const fn prob_replace_mean(&self) -> Real {
        self.prob_replace_mean
    }
    pub const fn number_replace_last(&self) -> usize {
        self.number_replace_last
    }
}
impl<Rng: rand::Rng, const D: usize> MonteCarlo<LatticeStateCuda<D>, D> for MetropolisHastingsSweepCuda<Rng> {
    type Error = CudaError;
    fn next_element(&mut self, mut state: LatticeStateCuda<D>) -> Result<LatticeStateCuda<D>, CudaError> {
        let (mut n_acc, mut sum_p) = (0_i64, 0_f64);
        check(unsafe {
            ffi::lq_sweep_metropolis(state.ctx.0, self.rng.next_u64(), self.rng.next_u64() >> 8, self.spread,
                                     self.number_of_update as i32, &mut n_acc, &mut sum_p)
        })?;
        self.number_replace_last = n_acc as usize;
        self.prob_replace_mean = sum_p / state.lattice.number_of_canonical_links_space() as f64;
        state.host_links = OnceLock::new();
        Ok(state)
    }
}

/// metropolis_hastings.rs:300-417 (the README's method): one uniformly random link per call, accepted on the local
/// action difference.  `hits_per_call` > 1 batches independent hits in one pair of launches (all on links of one random
/// (direction, colour) class; hits colliding on a link are dropped); 1 restates the reference call for call.
pub struct MetropolisHastingsDeltaDiagnosticCuda<Rng: rand::Rng> {
    spread: Real,
    hits_per_call: usize,
    has_replace_last: bool,
    prob_replace_last: Real,
    rng: Rng,
}
impl<Rng: rand::Rng> MetropolisHastingsDeltaDiagnosticCuda<Rng> {
    /// `None` for an invalid spread (metropolis_hastings.rs:343-353)
    pub fn new(spread: Real, rng: Rng) -> Option<Self> {
        Self::with_hits_per_call(spread, rng, 1)
    }
    pub fn with_hits_per_call(spread: Real, rng: Rng, hits_per_call: usize) -> Option<Self> {
        if spread <= 0_f64 || spread >= 1_f64 || hits_per_call == 0 {
            return None;
        }
        Some(Self { spread, hits_per_call, has_replace_last: false, prob_replace_last: 0.0, rng })
    }
    pub const fn prob_replace_last(&self) -> Real {
        self.prob_replace_last
    }
    pub const fn has_replace_last(&self) -> bool {
        self.has_replace_last
    }
    pub const fn spread(&self) -> Real {
        self.spread
    }
    pub const fn rng(&self) -> &Rng {
        &self.rng
    }
    pub fn rng_mut(&mut self) -> &mut Rng {
        &mut self.rng
    }
    pub fn rng_owned(self) -> Rng {
        self.rng
    }
}
impl<Rng: rand::Rng, const D: usize> MonteCarlo<LatticeStateCuda<D>, D> for MetropolisHastingsDeltaDiagnosticCuda<Rng> {
    type Error = CudaError;
    fn next_element(&mut self, mut state: LatticeStateCuda<D>) -> Result<LatticeStateCuda<D>, CudaError> {
        let (mut n_perf, mut n_acc, mut sum_p) = (0_i64, 0_i64, 0_f64);
        check(unsafe {
            ffi::lq_metropolis_hits(state.ctx.0, self.rng.next_u64(), self.rng.next_u64() >> 8, self.spread,
                                    self.hits_per_call as i64, 0, &mut n_perf, &mut n_acc, &mut sum_p)
        })?;
        self.prob_replace_last = sum_p / (n_perf.max(1) as f64);
        self.has_replace_last = n_acc > 0;
        state.host_links = OnceLock::new();
        Ok(state)
    }
}

/// metropolis_hastings.rs:40-118 as a `MonteCarloDefault`: the proposal is a device clone with `number_of_update`
/// random links multiplied by a matrix close to one; `McWrapper::new(MetropolisHastingsCuda::new(..)?, rng)` then
/// accepts it on `probability_of_replacement` = exp(H_links(old) - H_links(new)) (monte_carlo/mod.rs:137-142), both
/// energies being device reductions through `LatticeState::hamiltonian_links`.
pub struct MetropolisHastingsCuda {
    number_of_update: usize,
    spread: Real,
}
impl MetropolisHastingsCuda {
    pub fn new(number_of_update: usize, spread: Real) -> Option<Self> {
        if number_of_update == 0 || spread <= 0_f64 || spread >= 1_f64 {
            return None;
        }
        Some(Self { number_of_update, spread })
    }
}
impl<const D: usize> lattice_qcd_rs::simulation::MonteCarloDefault<LatticeStateCuda<D>, D> for MetropolisHastingsCuda {
    type Error = CudaError;
    fn potential_next_element<Rng>(&mut self, state: &LatticeStateCuda<D>, rng: &mut Rng)
                                   -> Result<LatticeStateCuda<D>, CudaError>
    where
        Rng: rand::Rng + ?Sized,
    {
        let mut new = state.clone();
        let (mut n_perf, mut n_acc, mut sum_p) = (0_i64, 0_i64, 0_f64);
        check(unsafe {
            ffi::lq_metropolis_hits(new.ctx.0, rng.next_u64(), rng.next_u64() >> 8, self.spread,
                                    self.number_of_update as i64, 1, &mut n_perf, &mut n_acc, &mut sum_p)
        })?;
        new.host_links = OnceLock::new();
        Ok(new)
    }
}
