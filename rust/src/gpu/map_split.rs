//! `impl MapSplit<Fr>` (src/cleanup/polys/common.rs:23-35) for device-resident tables: witness generation of the GKR layers
//! (bintree_add.rs:173-239, triangle_add.rs:126-157, logup_mainphase.rs:109-119) without the tables leaving HBM.
//!   Vec::algfn_map / algfn_map_split                 src/cleanup/polys/dense.rs:114-185      -> gkr_map_dense
//!   vecvec_map / vecvec_map_split (LO(0))            src/cleanup/polys/vecvec.rs:480-606     -> gkr_map_vecvec (mode 0 / 1)
//!   vecvec_map_split_to_dense                        src/cleanup/polys/vecvec.rs:608-654     -> `map_split_to_dense` (mode 2)
use super::ctx::*;
use super::ffi::*;
use super::gates::stack_of;
use crate::cleanup::polys::common::MapSplit;
use crate::cleanup::protocols::splits::SplitIdx;
use crate::cleanup::utils::algfn::AlgFn;
use ark_bls12_381::Fr;
use std::os::raw::c_void;
use std::ptr;

const MAX_OUT: usize = 256;

fn dense_call<Fnc: AlgFn<Fr>>(polys: &[DeviceTable], func: &Fnc, split: Option<(SplitIdx, usize)>) -> Vec<DeviceTable> {
    let ctx = polys[0].ctx.clone();
    let stack = stack_of::<Fr, Fnc>(func).expect("gate type has no device functor");
    let ptrs: Vec<*mut gkr_table> = polys[..func.n_ins()].iter().map(|t| t.raw).collect();
    let (kind, var, bundle) = match split {
        None => (-1, 0u32, 1u32),
        Some((SplitIdx::LO(k), b)) => (0, k as u32, b as u32),
        Some((SplitIdx::HI(k), b)) => (1, k as u32, b as u32),
    };
    let mut out = vec![ptr::null_mut::<gkr_table>(); MAX_OUT];
    let mut n_out = 0u32;
    ctx.check(unsafe {
        gkr_map_dense(ctx.raw, stack.gate.as_ptr(), stack.repeat.as_ptr(), stack.gate.len() as u32, ptrs.as_ptr(), ptrs.len() as u32, kind, var, bundle,
                      out.as_mut_ptr(), &mut n_out)
    });
    out[..n_out as usize].iter().map(|&r| DeviceTable::from_raw(&ctx, r)).collect()
}

impl MapSplit<Fr> for DeviceTable {
    fn algfn_map_split<Fnc: AlgFn<Fr>>(polys: &[Self], func: Fnc, var_idx: SplitIdx, bundle_size: usize) -> Vec<Self> {
        dense_call(polys, &func, Some((var_idx, bundle_size)))
    }
    fn algfn_map<Fnc: AlgFn<Fr>>(polys: &[Self], func: Fnc) -> Vec<Self> {
        dense_call(polys, &func, None)
    }
}

fn vecvec_call<Fnc: AlgFn<Fr>>(polys: &[DeviceVecVec], func: &Fnc, mode: i32, bundle: usize) -> (std::rc::Rc<GpuCtx>, Vec<*mut c_void>) {
    let ctx = polys[0].ctx.clone();
    let stack = stack_of::<Fr, Fnc>(func).expect("gate type has no device functor");
    let ptrs: Vec<*mut gkr_vecvec> = polys[..func.n_ins()].iter().map(|t| t.raw).collect();
    let mut out = vec![ptr::null_mut::<c_void>(); MAX_OUT];
    let mut n_out = 0u32;
    ctx.check(unsafe {
        gkr_map_vecvec(ctx.raw, stack.gate.as_ptr(), stack.repeat.as_ptr(), stack.gate.len() as u32, ptrs.as_ptr(), ptrs.len() as u32, mode, bundle as u32,
                       out.as_mut_ptr(), &mut n_out)
    });
    out.truncate(n_out as usize);
    (ctx, out)
}

impl MapSplit<Fr> for DeviceVecVec {
    /// the reference only ever splits ragged matrices at SplitIdx::LO(0) (bintree_add.rs:149-170, splits.rs:172-176)
    fn algfn_map_split<Fnc: AlgFn<Fr>>(polys: &[Self], func: Fnc, var_idx: SplitIdx, bundle_size: usize) -> Vec<Self> {
        assert!(matches!(var_idx, SplitIdx::LO(0)), "device vecvec_map_split: LO(0) only");
        let (ctx, out) = vecvec_call(polys, &func, 1, bundle_size);
        out.into_iter().map(|r| DeviceVecVec::from_raw(&ctx, r as *mut gkr_vecvec)).collect()
    }
    fn algfn_map<Fnc: AlgFn<Fr>>(polys: &[Self], func: Fnc) -> Vec<Self> {
        let (ctx, out) = vecvec_call(polys, &func, 0, 1);
        out.into_iter().map(|r| DeviceVecVec::from_raw(&ctx, r as *mut gkr_vecvec)).collect()
    }
}

/// vecvec_map_split_to_dense (vecvec.rs:608-654): the layer where every bucket row has shrunk to one pair
pub fn map_split_to_dense<Fnc: AlgFn<Fr>>(polys: &[DeviceVecVec], func: Fnc, bundle_size: usize) -> Vec<DeviceTable> {
    let (ctx, out) = vecvec_call(polys, &func, 2, bundle_size);
    out.into_iter().map(|r| DeviceTable::from_raw(&ctx, r as *mut gkr_table)).collect()
}
