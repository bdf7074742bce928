//! Commitment side: `KzgProvingKey::{commit, open}` (src/commitments/kzg.rs:123-133) and
//! `KnucklesProvingKey::compute_t` (src/commitments/knuckles.rs:111-154) with the SRS resident in HBM.
//! G1 affine points cross the boundary as 12 u64: x then y, Fq Montgomery limbs -- the in-memory layout of the two
//! coordinates of `ark_bls12_381::G1Affine`; (0, 0) encodes the point at infinity.
use super::ctx::*;
use super::ffi::*;
use ark_bls12_381::{Fq, Fr, G1Affine};
use ark_ec::AffineRepr;
use ark_ff::{BigInt, Fp};
use std::ptr;
use std::rc::Rc;

fn fq_limbs(x: &Fq) -> [u64; 6] {
    (x.0).0
}
fn fq_from_limbs(l: &[u64]) -> Fq {
    Fp::new_unchecked(BigInt([l[0], l[1], l[2], l[3], l[4], l[5]]))
}
fn g1_from_limbs(xy: &[u64; 12]) -> G1Affine {
    if xy.iter().all(|&w| w == 0) {
        return G1Affine::zero();
    }
    G1Affine::new_unchecked(fq_from_limbs(&xy[..6]), fq_from_limbs(&xy[6..]))
}

pub struct GpuKzgKey {
    ctx: Rc<GpuCtx>,
    srs: *mut gkr_srs,
    len: usize,
}

impl GpuKzgKey {
    /// from `KzgProvingKey::ptau_1()` (kzg.rs:107-109)
    pub fn new(ctx: &Rc<GpuCtx>, ptau_1: &[G1Affine]) -> Self {
        let mut flat = Vec::with_capacity(12 * ptau_1.len());
        for p in ptau_1 {
            match p.xy() {
                Some((x, y)) => {
                    flat.extend_from_slice(&fq_limbs(x));
                    flat.extend_from_slice(&fq_limbs(y));
                }
                None => flat.extend_from_slice(&[0u64; 12]),
            }
        }
        let mut srs = ptr::null_mut();
        ctx.check(unsafe { gkr_srs_upload(ctx.raw, flat.as_ptr(), ptau_1.len() as u64, 0, &mut srs) });
        Self { ctx: ctx.clone(), srs, len: ptau_1.len() }
    }

    /// KzgProvingKey::commit  kzg.rs:123-126
    pub fn commit_table(&self, poly: &DeviceTable) -> G1Affine {
        assert!(poly.len() <= self.len, "Vector is too large.");
        let mut out = [0u64; 12];
        self.ctx.check(unsafe { gkr_msm_g1(self.ctx.raw, self.srs, 0, poly.raw, poly.len() as u64, out.as_mut_ptr()) });
        g1_from_limbs(&out)
    }
    pub fn commit(&self, poly: &[Fr]) -> G1Affine {
        self.commit_table(&DeviceTable::upload(&self.ctx, poly))
    }

    /// KzgProvingKey::open  kzg.rs:129-132: (commitment to the quotient by x - pt, remainder)
    pub fn open_table(&self, poly: &DeviceTable, pt: Fr) -> (G1Affine, Fr) {
        let mut q = ptr::null_mut();
        let mut rem = [0u64; 4];
        let p = fr_limbs(&pt);
        self.ctx.check(unsafe { gkr_poly_div_by_linear(self.ctx.raw, poly.raw, p.as_ptr(), &mut q, rem.as_mut_ptr()) });
        let q = DeviceTable::from_raw(&self.ctx, q);
        (self.commit_table(&q), fr_from_limbs(rem))
    }
    pub fn open(&self, poly: &[Fr], pt: Fr) -> (G1Affine, Fr) {
        self.open_table(&DeviceTable::upload(&self.ctx, poly), pt)
    }

    /// `ev`  kzg.rs:142-150
    pub fn ev(&self, poly: &DeviceTable, x: Fr) -> Fr {
        let mut out = [0u64; 4];
        let xl = fr_limbs(&x);
        self.ctx.check(unsafe { gkr_poly_eval(self.ctx.raw, poly.raw, xl.as_ptr(), out.as_mut_ptr()) });
        fr_from_limbs(out)
    }
}

impl Drop for GpuKzgKey {
    fn drop(&mut self) {
        unsafe { gkr_srs_free(self.srs) }
    }
}

pub struct GpuKnucklesKey {
    pub kzg: GpuKzgKey,
    key: *mut gkr_knuckles,
    pub num_vars: usize,
}

impl GpuKnucklesKey {
    /// KnucklesProvingKey::new(kzg_pk, num_vars, k)  knuckles.rs:65-81 (the 2N - 1 inverses are computed on the device)
    pub fn new(kzg: GpuKzgKey, num_vars: usize, k: Fr) -> Self {
        assert!(kzg.len >= 2 * (1usize << num_vars) - 1, "SRS is too short.");
        let mut key = ptr::null_mut();
        let kl = fr_limbs(&k);
        kzg.ctx.check(unsafe { gkr_knuckles_create(kzg.ctx.raw, num_vars as u32, kl.as_ptr(), &mut key) });
        Self { kzg, key, num_vars }
    }

    /// KnucklesProvingKey::compute_t  knuckles.rs:111-154: (T, opening)
    pub fn compute_t(&self, poly: &DeviceTable, point: &[Fr]) -> (DeviceTable, Fr) {
        assert_eq!(point.len(), self.num_vars);
        let mut t = ptr::null_mut();
        let mut opening = [0u64; 4];
        let ctx = &self.kzg.ctx;
        ctx.check(unsafe { gkr_knuckles_compute_t(ctx.raw, self.key, poly.raw, limbs(point), point.len() as u32, &mut t, opening.as_mut_ptr()) });
        (DeviceTable::from_raw(ctx, t), fr_from_limbs(opening))
    }
}

impl Drop for GpuKnucklesKey {
    fn drop(&mut self) {
        unsafe { gkr_knuckles_free(self.key) }
    }
}
