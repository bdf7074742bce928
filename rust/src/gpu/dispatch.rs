//! The three constructor call sites patched into the reference (rust/reference.patch): each returns `Some(device object)`
//! when the field is `ark_bls12_381::Fr` and the gate type has a device functor, otherwise `None` (CPU object as before).
use super::ctx::*;
use super::gates::{single_gate, stack_of};
use super::sumcheckable::GpuSo;
use crate::cleanup::polys::vecvec::VecVecPolynomial;
use crate::cleanup::utils::algfn::AlgFn;
use crate::utils::make_gamma_pows;
use ark_bls12_381::Fr;
use ark_ff::PrimeField;
use std::any::TypeId;

fn is_fr<F: PrimeField>() -> bool {
    TypeId::of::<F>() == TypeId::of::<Fr>()
}
fn fr_slice<F: PrimeField>(v: &[F]) -> &[Fr] {
    unsafe { std::slice::from_raw_parts(v.as_ptr() as *const Fr, v.len()) }
}
fn upload_all<F: PrimeField>(ctx: &std::rc::Rc<GpuCtx>, polys: &[Vec<F>]) -> Vec<DeviceTable> {
    polys.iter().map(|p| DeviceTable::upload(ctx, fr_slice(p))).collect()
}

/// DenseEqSumcheckObject::new(..).rlc(gamma)  (sumcheck.rs:849-872): polys ++ [eq(point)], EqWrapper(GammaWrapper(f, gamma)),
/// claim = gamma_rlc(gamma, evs)
pub fn try_dense_eq<F: PrimeField, Fun: AlgFn<F>>(advice: &[Vec<F>], f: &Fun, point: &[F], evs: &[F], gamma: F) -> Option<GpuSo<F>> {
    if !is_fr::<F>() {
        return None;
    }
    let gate = single_gate(&stack_of::<F, Fun>(f)?)?;
    let ctx = GpuCtx::current();
    let mut tables = upload_all(&ctx, advice);
    tables.push(DeviceTable::eq(&ctx, fr_slice(point), Fr::from(1u64)));
    let gamma_pows = make_gamma_pows(gamma, f.n_outs());
    let claim = crate::cleanup::protocols::sumcheck::gamma_rlc(gamma, evs);
    Some(GpuSo::dense_eq_gamma(&ctx, gate, &gamma_pows, tables, point.len(), claim))
}

/// DenseDeg2SumcheckObject::new(..).rlc(gamma)  (dense_eq.rs:43-60, 199-214)
pub fn try_dense_deg2<F: PrimeField, Fun: AlgFn<F>>(advice: &[Vec<F>], f: &Fun, point: &[F], evs: &[F], gamma: F) -> Option<GpuSo<F>> {
    if !is_fr::<F>() {
        return None;
    }
    let stack = stack_of::<F, Fun>(f)?;
    let ctx = GpuCtx::current();
    let gamma_pows = make_gamma_pows(gamma, f.n_outs());
    let mut claim = evs[0];
    for i in 1..evs.len() {
        claim += gamma_pows[i] * evs[i];
    }
    Some(GpuSo::dense_deg2(&ctx, &stack, upload_all(&ctx, advice), &gamma_pows, claim, point))
}

/// VecVecDeg2SumcheckObject::new(..).rlc(gamma)  (vecvec_eq.rs:53-71, 425-443)
pub fn try_vecvec_deg2<F: PrimeField, Fun: AlgFn<F>>(advice: &[VecVecPolynomial<F>], f: &Fun, point: &[F], evs: &[F], num_vertical_vars: usize,
                                                     gamma: F) -> Option<GpuSo<F>> {
    if !is_fr::<F>() {
        return None;
    }
    let gate = single_gate(&stack_of::<F, Fun>(f)?)?;
    let ctx = GpuCtx::current();
    let polys: Vec<DeviceVecVec> = advice
        .iter()
        .map(|p| DeviceVecVec::upload(&ctx, unsafe { &*(p as *const VecVecPolynomial<F> as *const VecVecPolynomial<Fr>) }))
        .collect();
    let gamma_pows = make_gamma_pows(gamma, f.n_outs());
    let mut claim = evs[0];
    for i in 1..evs.len() {
        claim += gamma_pows[i] * evs[i];
    }
    Some(GpuSo::vecvec_deg2(&ctx, gate, polys, &gamma_pows, claim, point, num_vertical_vars))
}
