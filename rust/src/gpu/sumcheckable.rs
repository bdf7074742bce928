//! `impl Sumcheckable<F>` (src/cleanup/protocols/sumchecks/vecvec_eq.rs:218-225) over device-resident objects.
//!
//! One type serves the three reference objects -- the C ABI gives them the same five entries:
//!   DenseSumcheckObjectSO       src/cleanup/protocols/sumcheck.rs:241-347       `GpuSo::dense_plain / dense_eq_gamma`
//!   DenseDeg2SumcheckObjectSO   src/cleanup/protocols/sumchecks/dense_eq.rs:62-173     `GpuSo::dense_deg2`
//!   VecVecDeg2SumcheckObjectSO  src/cleanup/protocols/sumchecks/vecvec_eq.rs:74-398    `GpuSo::vecvec_deg2`
//! Call protocol is the reference's: `unipoly()` then `bind()`; bind without unipoly, or unipoly twice on the Deg2 objects,
//! returns GKR_ERR_PROTOCOL from the device side and panics here exactly where the reference panics
//! (sumcheck.rs:271-274, dense_eq.rs:105,109-111).
use super::ctx::*;
use super::ffi::*;
use super::gates::GateStack;
use crate::cleanup::protocols::sumchecks::vecvec_eq::Sumcheckable;
use ark_bls12_381::Fr;
use ark_ff::PrimeField;
use liblasso::poly::unipoly::UniPoly;
use std::any::TypeId;
use std::marker::PhantomData;
use std::ptr;
use std::rc::Rc;

pub struct GpuSo<F: PrimeField> {
    ctx: Rc<GpuCtx>,
    raw: *mut gkr_so,
    challenges: Vec<F>,
    // tables the object was created from stay alive as long as it does
    _tables: Vec<DeviceTable>,
    _vecvecs: Vec<DeviceVecVec>,
    _pd: PhantomData<F>,
}

/// F must be ark_bls12_381::Fr: checked once at construction (`dispatch.rs`), then slices are reinterpreted in place.
#[inline]
fn as_fr<F: PrimeField>(v: &[F]) -> &[Fr] {
    assert!(TypeId::of::<F>() == TypeId::of::<Fr>());
    unsafe { std::slice::from_raw_parts(v.as_ptr() as *const Fr, v.len()) }
}
#[inline]
fn from_fr<F: PrimeField>(v: Vec<Fr>) -> Vec<F> {
    assert!(TypeId::of::<F>() == TypeId::of::<Fr>());
    let mut v = std::mem::ManuallyDrop::new(v);
    unsafe { Vec::from_raw_parts(v.as_mut_ptr() as *mut F, v.len(), v.capacity()) }
}

impl<F: PrimeField> GpuSo<F> {
    fn wrap(ctx: &Rc<GpuCtx>, raw: *mut gkr_so, tables: Vec<DeviceTable>, vecvecs: Vec<DeviceVecVec>) -> Self {
        Self { ctx: ctx.clone(), raw, challenges: vec![], _tables: tables, _vecvecs: vecvecs, _pd: PhantomData }
    }

    /// Latency of small rounds (`gkr_so_set_prelaunch`): only for an object that ONE loop drives alone as unipoly, transcript,
    /// bind, ... with no other device work in between -- GenericSumcheckProtocol::prove (sumcheck.rs:101-123), which
    /// `dispatch.rs` wraps.  NOT for the combined loop of pushforward.rs:781-806: there the first kernel of the second object
    /// would queue up behind the pre-launched kernel of the first one and the device-side watchdog (2 s) would have to
    /// break the wait.
    pub fn set_prelaunch(&mut self, on: bool) {
        unsafe { gkr_so_set_prelaunch(self.raw, on as i32) };
    }

    /// DenseSumcheckObjectSO::new(polys, f, num_vars, claim_hint) with a single-output gate (PROD3; FOLDED_PROD with
    /// `gate_consts = make_gamma_pows(gamma, nargs)`)
    pub fn dense_plain(ctx: &Rc<GpuCtx>, gate: i32, gate_param: u32, gate_consts: &[F], polys: Vec<DeviceTable>, num_vars: usize, claim: F) -> Self {
        Self::dense(ctx, GKR_SO_PLAIN, gate, gate_param, gate_consts, polys, num_vars, claim)
    }
    /// the same over EqWrapper(GammaWrapper(f, gamma)) (sumcheck.rs:706-741, 802-829): `polys` ends with the eq table and
    /// `gamma_pows[i] = gamma^i` for i < f.n_outs()
    pub fn dense_eq_gamma(ctx: &Rc<GpuCtx>, gate: i32, gamma_pows: &[F], polys: Vec<DeviceTable>, num_vars: usize, claim: F) -> Self {
        Self::dense(ctx, GKR_SO_EQ_GAMMA, gate, 0, gamma_pows, polys, num_vars, claim)
    }
    fn dense(ctx: &Rc<GpuCtx>, kind: i32, gate: i32, gate_param: u32, consts: &[F], polys: Vec<DeviceTable>, num_vars: usize, claim: F) -> Self {
        let ptrs: Vec<*mut gkr_table> = polys.iter().map(|t| t.raw).collect();
        let mut raw = ptr::null_mut();
        ctx.check(unsafe {
            gkr_so_create_dense(ctx.raw, kind, gate, gate_param, limbs(as_fr(consts)), consts.len() as u32, ptrs.as_ptr(), ptrs.len() as u32,
                                num_vars as u32, limbs(as_fr(&[claim])), &mut raw)
        });
        Self::wrap(ctx, raw, polys, vec![])
    }
    /// DenseDeg2SumcheckObjectSO::new(polys, func, gamma_pows, claim, point)  dense_eq.rs:75-95
    pub fn dense_deg2(ctx: &Rc<GpuCtx>, stack: &GateStack, polys: Vec<DeviceTable>, gamma_pows: &[F], claim: F, point: &[F]) -> Self {
        let ptrs: Vec<*mut gkr_table> = polys.iter().map(|t| t.raw).collect();
        let mut raw = ptr::null_mut();
        ctx.check(unsafe {
            gkr_so_create_deg2_dense(ctx.raw, stack.gate.as_ptr(), stack.repeat.as_ptr(), stack.gate.len() as u32, ptrs.as_ptr(), ptrs.len() as u32,
                                     limbs(as_fr(gamma_pows)), limbs(as_fr(&[claim])), limbs(as_fr(point)), point.len() as u32, &mut raw)
        });
        Self::wrap(ctx, raw, polys, vec![])
    }
    /// VecVecDeg2SumcheckObjectSO::new(polys, func, gamma_pows, claim, point, col_logsize)  vecvec_eq.rs:94-118
    pub fn vecvec_deg2(ctx: &Rc<GpuCtx>, gate: i32, polys: Vec<DeviceVecVec>, gamma_pows: &[F], claim: F, point: &[F], col_logsize: usize) -> Self {
        let ptrs: Vec<*mut gkr_vecvec> = polys.iter().map(|t| t.raw).collect();
        let mut raw = ptr::null_mut();
        ctx.check(unsafe {
            gkr_so_create_deg2_vecvec(ctx.raw, gate, ptrs.as_ptr(), ptrs.len() as u32, limbs(as_fr(gamma_pows)), limbs(as_fr(&[claim])),
                                      limbs(as_fr(point)), point.len() as u32, col_logsize as u32, &mut raw)
        });
        Self::wrap(ctx, raw, vec![], polys)
    }
    /// running claim (`DenseSumcheckObjectSO::claim`, `VecVecDeg2SumcheckObjectSO::claim()`)
    pub fn claim(&self) -> F {
        let mut out = [0u64; 4];
        self.ctx.check(unsafe { gkr_so_claim(self.raw, out.as_mut_ptr()) });
        from_fr::<F>(vec![fr_from_limbs(out)])[0]
    }
}

impl<F: PrimeField> Sumcheckable<F> for GpuSo<F> {
    fn bind(&mut self, t: F) {
        self.ctx.check(unsafe { gkr_so_bind(self.raw, limbs(as_fr(&[t]))) });
        self.challenges.push(t);
    }

    /// the device returns the evaluations at 0..=deg (eval(0) = claim - eval(1) for the dense object, `from12` for the Deg2
    /// objects); `UniPoly::from_evals` is the same interpolation the reference applies (sumcheck.rs:327, vecvec_eq.rs:210-215)
    fn unipoly(&mut self) -> UniPoly<F> {
        let mut evals = vec![Fr::from(0u64); 5];
        let mut n = 0u32;
        self.ctx.check(unsafe { gkr_so_unipoly(self.raw, limbs_mut(&mut evals), &mut n) });
        evals.truncate(n as usize);
        UniPoly::from_evals(&from_fr::<F>(evals))
    }

    fn final_evals(&self) -> Vec<F> {
        let n = unsafe { gkr_so_num_polys(self.raw) } as usize;
        let mut out = vec![Fr::from(0u64); n];
        self.ctx.check(unsafe { gkr_so_final_evals(self.raw, limbs_mut(&mut out)) });
        from_fr::<F>(out)
    }

    fn challenges(&self) -> &[F] {
        &self.challenges
    }
}

impl<F: PrimeField> Drop for GpuSo<F> {
    fn drop(&mut self) {
        unsafe { gkr_so_destroy(self.raw) }
    }
}
