//! Raw bindings: one `extern "C"` item per entry of `include/gkr_msm_b200.h` that the shim uses (same names, same argument
//! order).  All field elements cross the boundary as 4 little-endian u64 limbs in Montgomery form -- the in-memory layout of
//! `ark_bls12_381::Fr` (`Fp256<MontBackend<FrConfig, 4>>`), so a `&[Fr]` is passed as a pointer without conversion.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int, c_void};

macro_rules! opaque { ($($n:ident),*) => { $( #[repr(C)] pub struct $n { _p: [u8; 0] } )* } }
opaque!(gkr_ctx, gkr_table, gkr_so, gkr_vecvec, gkr_srs, gkr_knuckles, gkr_u32buf);

pub const GKR_OK: c_int = 0;
pub const GKR_SO_PLAIN: c_int = 0;
pub const GKR_SO_EQ_GAMMA: c_int = 1;
// enum gkr_gate_id
pub const GATE_AFF_L1: c_int = 0;
pub const GATE_AFF_L2: c_int = 1;
pub const GATE_AFF_L3: c_int = 2;
pub const GATE_PRJ_L1: c_int = 3;
pub const GATE_PRJ_L2: c_int = 4;
pub const GATE_PRJ_L3: c_int = 5;
pub const GATE_TRI_L1: c_int = 6;
pub const GATE_BITCHECK: c_int = 7;
pub const GATE_LOGUP_LAYER: c_int = 8;
pub const GATE_ADD_INVERSES: c_int = 9;
pub const GATE_PROD3: c_int = 10;
pub const GATE_FOLDED_PROD: c_int = 11;
pub const GATE_ID: c_int = 12;
pub const GATE_AFF_L1_BITCHECK2: c_int = 13;

extern "C" {
    pub fn gkr_ctx_create(device: c_int, out: *mut *mut gkr_ctx) -> c_int;
    pub fn gkr_ctx_destroy(ctx: *mut gkr_ctx);
    pub fn gkr_ctx_peer_pool(ctx: *mut gkr_ctx, n_devices: c_int, out_stats: *mut u64) -> c_int;
    pub fn gkr_last_error(ctx: *const gkr_ctx) -> *const c_char;
    pub fn gkr_ctx_sync(ctx: *mut gkr_ctx) -> c_int;

    pub fn gkr_table_upload(ctx: *mut gkr_ctx, limbs: *const u64, n: u64, out: *mut *mut gkr_table) -> c_int;
    pub fn gkr_table_download(ctx: *mut gkr_ctx, t: *const gkr_table, limbs_out: *mut u64) -> c_int;
    pub fn gkr_table_len(t: *const gkr_table) -> u64;
    pub fn gkr_table_free(t: *mut gkr_table);
    pub fn gkr_eq_table(ctx: *mut gkr_ctx, point: *const u64, n: u32, mult: *const u64, out: *mut *mut gkr_table) -> c_int;

    pub fn gkr_so_create_dense(ctx: *mut gkr_ctx, so_kind: c_int, gate: c_int, gate_param: u32, gate_consts: *const u64, n_consts: u32,
                               tables: *const *mut gkr_table, n_polys: u32, num_vars: u32, claim: *const u64, out: *mut *mut gkr_so) -> c_int;
    pub fn gkr_so_create_deg2_dense(ctx: *mut gkr_ctx, part_gate: *const c_int, part_repeat: *const u32, n_parts: u32, tables: *const *mut gkr_table,
                                    n_polys: u32, gamma_pows: *const u64, claim: *const u64, point: *const u64, num_vars: u32, out: *mut *mut gkr_so)
                                    -> c_int;
    pub fn gkr_so_create_deg2_vecvec(ctx: *mut gkr_ctx, gate: c_int, polys: *const *mut gkr_vecvec, n_polys: u32, gamma_pows: *const u64,
                                     claim: *const u64, point: *const u64, num_vars: u32, col_logsize: u32, out: *mut *mut gkr_so) -> c_int;
    pub fn gkr_so_unipoly(so: *mut gkr_so, evals_out: *mut u64, n_evals: *mut u32) -> c_int;
    pub fn gkr_so_bind(so: *mut gkr_so, t: *const u64) -> c_int;
    pub fn gkr_so_final_evals(so: *mut gkr_so, out: *mut u64) -> c_int;
    pub fn gkr_so_claim(so: *const gkr_so, out: *mut u64) -> c_int;
    pub fn gkr_so_num_polys(so: *const gkr_so) -> u32;
    pub fn gkr_so_destroy(so: *mut gkr_so);
    pub fn gkr_so_set_prelaunch(so: *mut gkr_so, on: c_int) -> c_int;

    pub fn gkr_vecvec_upload(ctx: *mut gkr_ctx, flat: *const u64, row_len: *const u32, n_rows: u32, row_pad: *const u64, col_pad: *const u64,
                             row_logsize: u32, col_logsize: u32, out: *mut *mut gkr_vecvec) -> c_int;
    pub fn gkr_vecvec_total_len(v: *const gkr_vecvec) -> u64;
    pub fn gkr_vecvec_num_rows(v: *const gkr_vecvec) -> u32;
    pub fn gkr_vecvec_download(ctx: *mut gkr_ctx, v: *const gkr_vecvec, flat_out: *mut u64, row_len_out: *mut u32, row_pad: *mut u64, col_pad: *mut u64,
                               row_logsize: *mut u32, col_logsize: *mut u32) -> c_int;
    pub fn gkr_vecvec_free(v: *mut gkr_vecvec);

    /// split_kind < 0: Vec::algfn_map; 0: algfn_map_split at SplitIdx::LO(var_idx); 1: at SplitIdx::HI(var_idx)
    pub fn gkr_map_dense(ctx: *mut gkr_ctx, part_gate: *const c_int, part_repeat: *const u32, n_parts: u32, input: *const *mut gkr_table, n_in: u32,
                         split_kind: c_int, var_idx: u32, bundle_size: u32, out: *mut *mut gkr_table, n_out: *mut u32) -> c_int;
    /// mode 0: vecvec_map -> gkr_vecvec*; 1: vecvec_map_split at LO(0) -> gkr_vecvec*; 2: vecvec_map_split_to_dense -> gkr_table*
    pub fn gkr_map_vecvec(ctx: *mut gkr_ctx, part_gate: *const c_int, part_repeat: *const u32, n_parts: u32, input: *const *mut gkr_vecvec, n_in: u32,
                          mode: c_int, bundle_size: u32, out: *mut *mut c_void, n_out: *mut u32) -> c_int;

    pub fn gkr_srs_upload(ctx: *mut gkr_ctx, points: *const u64, n: u64, projective: c_int, out: *mut *mut gkr_srs) -> c_int;
    pub fn gkr_srs_free(s: *mut gkr_srs);
    pub fn gkr_msm_g1(ctx: *mut gkr_ctx, srs: *const gkr_srs, first: u64, scalars: *const gkr_table, n: u64, out_xy: *mut u64) -> c_int;
    pub fn gkr_poly_div_by_linear(ctx: *mut gkr_ctx, poly: *const gkr_table, pt: *const u64, quotient: *mut *mut gkr_table, rem: *mut u64) -> c_int;
    pub fn gkr_poly_eval(ctx: *mut gkr_ctx, poly: *const gkr_table, x: *const u64, out: *mut u64) -> c_int;
    pub fn gkr_knuckles_create(ctx: *mut gkr_ctx, num_vars: u32, k: *const u64, out: *mut *mut gkr_knuckles) -> c_int;
    pub fn gkr_knuckles_free(key: *mut gkr_knuckles);
    pub fn gkr_knuckles_compute_t(ctx: *mut gkr_ctx, key: *const gkr_knuckles, poly: *const gkr_table, point: *const u64, n_point: u32,
                                  t_out: *mut *mut gkr_table, opening: *mut u64) -> c_int;
}
