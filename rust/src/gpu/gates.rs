//! `AlgFn` is an open Rust generic (src/cleanup/utils/algfn.rs:20-34); the device has a closed set of gate functors
//! (`enum gkr_gate_id`).  This module maps the gate TYPES the reference instantiates on the hot path onto
//! "stacks" `Stacked(Repeated(g_0, r_0), Repeated(g_1, r_1))` of device gates:
//!   twisted-Edwards gates        src/cleanup/utils/twisted_edwards_ops.rs:141-147 (one struct per gate, via make_algfn!)
//!   BitCheckFn, IdAlgFn, RepeatedAlgFn, StackedAlgFn   src/cleanup/utils/algfn.rs:130-291
//!   LogupLayerFn, AddInversesFn  pushforward/logup_mainphase.rs:32-61, pushforward/pushforward.rs:255-281
//! The combinators keep their inner gates in private fields, so the stack is recovered from the public API only: the type
//! name gives the leaf gates in order and `n_ins()` gives the repeat counts (every first operand of a `StackedAlgFn` in the
//! reference is a plain gate).  Unknown gate types yield `None` and the caller keeps the CPU object.
use super::ffi::*;
use crate::cleanup::utils::algfn::AlgFn;
use ark_ff::PrimeField;
use std::os::raw::c_int;

/// (type-name fragment, gate id, n_ins); longer names first so `affine_twisted_edwards_add_l1` is not taken for `twisted_edwards_add_l1`
const LEAVES: &[(&str, c_int, usize)] = &[
    ("triangle_twisted_edwards_add_l1", GATE_TRI_L1, 12),
    ("affine_twisted_edwards_add_l1", GATE_AFF_L1, 4),
    ("affine_twisted_edwards_add_l2", GATE_AFF_L2, 3),
    ("affine_twisted_edwards_add_l3", GATE_AFF_L3, 3),
    ("twisted_edwards_add_l1", GATE_PRJ_L1, 6),
    ("twisted_edwards_add_l2", GATE_PRJ_L2, 4),
    ("twisted_edwards_add_l3", GATE_PRJ_L3, 4),
    ("BitCheckFn", GATE_BITCHECK, 1),
    ("LogupLayerFn", GATE_LOGUP_LAYER, 4),
    ("AddInversesFn", GATE_ADD_INVERSES, 2),
    ("IdAlgFn", GATE_ID, 1),
];

#[derive(Clone, Debug, PartialEq, Eq)]
pub struct GateStack {
    pub gate: Vec<c_int>,
    pub repeat: Vec<u32>,
}

/// leaves of `Fun` in order of appearance in its type name
fn leaves_of(name: &str) -> Vec<(c_int, usize)> {
    let mut out = vec![];
    let mut rest = name;
    'scan: while !rest.is_empty() {
        for &(frag, id, n_ins) in LEAVES {
            if rest.starts_with(frag) {
                // must be a whole path segment: preceded by "::" or '<' / ' ' (or start), followed by '<' or end
                out.push((id, n_ins));
                rest = &rest[frag.len()..];
                continue 'scan;
            }
        }
        let mut it = rest.char_indices();
        it.next();
        rest = it.next().map_or("", |(i, _)| &rest[i..]);
    }
    out
}

pub fn stack_of<F: PrimeField, Fun: AlgFn<F>>(f: &Fun) -> Option<GateStack> {
    let leaves = leaves_of(std::any::type_name::<Fun>());
    let n_ins = f.n_ins();
    match leaves.as_slice() {
        [(id, k)] if n_ins % k == 0 => Some(GateStack { gate: vec![*id], repeat: vec![(n_ins / k) as u32] }),
        [(id0, k0), (id1, k1)] if n_ins >= *k0 && (n_ins - k0) % k1 == 0 => {
            let r1 = ((n_ins - k0) / k1) as u32;
            if r1 == 0 {
                Some(GateStack { gate: vec![*id0], repeat: vec![1] }) // RepeatedAlgFn(.., 0): triangle layer 0
            } else {
                Some(GateStack { gate: vec![*id0, *id1], repeat: vec![1, r1] })
            }
        }
        _ => None,
    }
}

/// single device gate id for the objects that take one (`gkr_so_create_dense` with GKR_SO_EQ_GAMMA, `gkr_so_create_deg2_vecvec`):
/// a plain leaf, or Stacked(affine L1, Repeated(BitCheck, 2)) = GATE_AFF_L1_BITCHECK2 (bintree_add.rs:259-273)
pub fn single_gate(s: &GateStack) -> Option<c_int> {
    match (s.gate.as_slice(), s.repeat.as_slice()) {
        ([g], [1]) => Some(*g),
        ([GATE_AFF_L1, GATE_BITCHECK], [1, 2]) => Some(GATE_AFF_L1_BITCHECK2),
        _ => None,
    }
}

#[cfg(test)]
mod tests {
    use super::*;
    #[test]
    fn leaf_order_and_prefixes() {
        let n = "GKR_MSM::cleanup::utils::algfn::StackedAlgFn<Fr, GKR_MSM::cleanup::utils::twisted_edwards_ops::algfns::triangle_twisted_edwards_add_l1<Fr>, \
                 GKR_MSM::cleanup::utils::algfn::RepeatedAlgFn<Fr, GKR_MSM::cleanup::utils::twisted_edwards_ops::algfns::twisted_edwards_add_l1<Fr>>>";
        assert_eq!(leaves_of(n), vec![(GATE_TRI_L1, 12), (GATE_PRJ_L1, 6)]);
        assert_eq!(leaves_of("x::affine_twisted_edwards_add_l3<Fr>"), vec![(GATE_AFF_L3, 3)]);
    }
}
