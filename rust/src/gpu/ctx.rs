//! Device context, resident tables, and the (zero-cost) boundary view of `ark_bls12_381::Fr`.
use super::ffi::*;
use ark_bls12_381::Fr;
use ark_ff::{BigInt, Fp};
use std::ffi::CStr;
use std::ptr;
use std::rc::Rc;

/// One context per calling thread (protocol code is single-threaded: the transcript is `&mut`).
pub struct GpuCtx {
    pub(crate) raw: *mut gkr_ctx,
}

impl GpuCtx {
    pub fn new(device: i32) -> Rc<Self> {
        let mut raw = ptr::null_mut();
        let rc = unsafe { gkr_ctx_create(device, &mut raw) };
        // the reference never returns Result on this path -- it panics; so does the shim (no CPU fallback behind a failed create)
        assert!(rc == GKR_OK && !raw.is_null(), "gkr_ctx_create failed ({rc}): no CUDA device / library");
        Rc::new(Self { raw })
    }

    /// thread-local default context on device `GKR_DEVICE` (default 0); `GKR_PEER_POOL=N` lends it the HBM of GPUs 0..N
    /// (instances beyond one GPU's memory: `examples/pippenger --x-logsize 24` with full-width scalars needs 430 GiB)
    pub fn current() -> Rc<Self> {
        thread_local! { static CTX: Rc<GpuCtx> = {
            let c = GpuCtx::new(std::env::var("GKR_DEVICE").ok().and_then(|v| v.parse().ok()).unwrap_or(0));
            if let Some(n) = std::env::var("GKR_PEER_POOL").ok().and_then(|v| v.parse::<i32>().ok()) { c.peer_pool(n); }
            c
        }; }
        CTX.with(|c| c.clone())
    }

    /// `gkr_ctx_peer_pool`: tables that do not fit the home GPU are placed on the other GPUs of the box (NVLink peer access);
    /// returns (bytes on peers now, their peak)
    pub fn peer_pool(&self, n_devices: i32) -> (u64, u64) {
        let mut st = [0u64; 2];
        self.check(unsafe { gkr_ctx_peer_pool(self.raw, n_devices, st.as_mut_ptr()) });
        (st[0], st[1])
    }

    #[track_caller]
    pub(crate) fn check(&self, rc: i32) {
        if rc != GKR_OK {
            let msg = unsafe { CStr::from_ptr(gkr_last_error(self.raw)) }.to_string_lossy().into_owned();
            panic!("gkr-msm-b200: status {rc}: {msg}");
        }
    }
}

impl Drop for GpuCtx {
    fn drop(&mut self) {
        unsafe { gkr_ctx_destroy(self.raw) }
    }
}

/// `&[Fr]` as the boundary layout: `Fp256<MontBackend<FrConfig, 4>>` is `#[repr(transparent)]`-like over `BigInt<4>([u64; 4])`
/// holding the Montgomery representation, so the slice IS n x 4 little-endian u64 limbs.
#[inline]
pub fn limbs(v: &[Fr]) -> *const u64 {
    debug_assert_eq!(std::mem::size_of::<Fr>(), 32);
    v.as_ptr() as *const u64
}
#[inline]
pub fn limbs_mut(v: &mut [Fr]) -> *mut u64 {
    v.as_mut_ptr() as *mut u64
}
#[inline]
pub fn fr_from_limbs(l: [u64; 4]) -> Fr {
    Fp::new_unchecked(BigInt(l)) // already Montgomery form: no conversion
}
#[inline]
pub fn fr_limbs(x: &Fr) -> [u64; 4] {
    (x.0).0
}

/// `Vec<Fr>` resident in HBM.  Sumcheck objects never modify the tables they were created from (the first fold writes a
/// fresh half-size buffer), so "clone before prove" (pushforward.rs:682-685) costs nothing on the device.
pub struct DeviceTable {
    pub(crate) ctx: Rc<GpuCtx>,
    pub(crate) raw: *mut gkr_table,
}

impl DeviceTable {
    pub fn upload(ctx: &Rc<GpuCtx>, v: &[Fr]) -> Self {
        let mut raw = ptr::null_mut();
        ctx.check(unsafe { gkr_table_upload(ctx.raw, limbs(v), v.len() as u64, &mut raw) });
        Self { ctx: ctx.clone(), raw }
    }
    pub fn len(&self) -> usize {
        unsafe { gkr_table_len(self.raw) as usize }
    }
    pub fn download(&self) -> Vec<Fr> {
        let mut out = vec![Fr::from(0u64); self.len()];
        self.ctx.check(unsafe { gkr_table_download(self.ctx.raw, self.raw, limbs_mut(&mut out)) });
        out
    }
    /// eq_poly_sequence_from_multiplier(mult, point).last()  (src/utils.rs:222-262), built on the device
    pub fn eq(ctx: &Rc<GpuCtx>, point: &[Fr], mult: Fr) -> Self {
        let mut raw = ptr::null_mut();
        let m = fr_limbs(&mult);
        ctx.check(unsafe { gkr_eq_table(ctx.raw, limbs(point), point.len() as u32, m.as_ptr(), &mut raw) });
        Self { ctx: ctx.clone(), raw }
    }
    pub(crate) fn from_raw(ctx: &Rc<GpuCtx>, raw: *mut gkr_table) -> Self {
        Self { ctx: ctx.clone(), raw }
    }
}

impl Drop for DeviceTable {
    fn drop(&mut self) {
        unsafe { gkr_table_free(self.raw) }
    }
}

/// `VecVecPolynomial<Fr>` (src/cleanup/polys/vecvec.rs:149-160) resident in HBM in CSR form.
pub struct DeviceVecVec {
    pub(crate) ctx: Rc<GpuCtx>,
    pub(crate) raw: *mut gkr_vecvec,
}

impl DeviceVecVec {
    pub fn upload(ctx: &Rc<GpuCtx>, p: &crate::cleanup::polys::vecvec::VecVecPolynomial<Fr>) -> Self {
        let row_len: Vec<u32> = p.data.iter().map(|r| r.len() as u32).collect();
        let flat: Vec<Fr> = p.data.iter().flat_map(|r| r.iter().copied()).collect();
        let (rp, cp) = (fr_limbs(&p.row_pad), fr_limbs(&p.col_pad));
        let mut raw = ptr::null_mut();
        ctx.check(unsafe {
            gkr_vecvec_upload(ctx.raw, limbs(&flat), row_len.as_ptr(), row_len.len() as u32, rp.as_ptr(), cp.as_ptr(), p.row_logsize as u32,
                              p.col_logsize as u32, &mut raw)
        });
        Self { ctx: ctx.clone(), raw }
    }
    pub fn download(&self) -> crate::cleanup::polys::vecvec::VecVecPolynomial<Fr> {
        let n_rows = unsafe { gkr_vecvec_num_rows(self.raw) } as usize;
        let total = unsafe { gkr_vecvec_total_len(self.raw) } as usize;
        let mut flat = vec![Fr::from(0u64); total];
        let mut row_len = vec![0u32; n_rows];
        let (mut rp, mut cp, mut rl, mut cl) = ([0u64; 4], [0u64; 4], 0u32, 0u32);
        self.ctx.check(unsafe {
            gkr_vecvec_download(self.ctx.raw, self.raw, limbs_mut(&mut flat), row_len.as_mut_ptr(), rp.as_mut_ptr(), cp.as_mut_ptr(), &mut rl, &mut cl)
        });
        let mut data = Vec::with_capacity(n_rows);
        let mut off = 0usize;
        for &l in &row_len {
            data.push(flat[off..off + l as usize].to_vec());
            off += l as usize;
        }
        crate::cleanup::polys::vecvec::VecVecPolynomial::new_unchecked(data, fr_from_limbs(rp), fr_from_limbs(cp), rl as usize, cl as usize)
    }
    pub(crate) fn from_raw(ctx: &Rc<GpuCtx>, raw: *mut gkr_vecvec) -> Self {
        Self { ctx: ctx.clone(), raw }
    }
}

impl Drop for DeviceVecVec {
    fn drop(&mut self) {
        unsafe { gkr_vecvec_free(self.raw) }
    }
}
