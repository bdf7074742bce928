//! sm_100a backend of the data-parallel hot path, behind the C ABI of gkr-msm-b200 (`include/gkr_msm_b200.h`).
//!
//! Host code stays in Rust: the Fiat-Shamir transcript (`cleanup::proof_transcript`), challenge sampling and protocol
//! orchestration (`Protocol2` impls) are the crate's own; only the objects behind `Sumcheckable`, `MapSplit` and
//! `KzgProvingKey::commit` live on the device.  Per round, at most four field elements cross PCIe (device -> host) and one
//! 128-bit challenge goes back.
//!
//! NOT COMPILED in the image this was written in (no cargo); written against the trait definitions by inspection.
pub mod ffi;
pub mod ctx;
pub mod gates;
pub mod sumcheckable;
pub mod map_split;
pub mod commit;
pub mod dispatch;
