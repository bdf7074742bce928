//! Reference-side parity harness: proves ONE seeded `examples/pippenger` instance with the crate's stock CPU path, dumps
//! (inputs, SRS seed, proof bytes) as files, and -- when the crate is built with `--features gpu` (the sumcheck objects then
//! live on the device, see rust/reference.patch) -- asserts the proof bytes are identical to a CPU-only run recorded earlier.
//!
//!   cargo +nightly run --release --features "parallel"      --bin gpu_parity -- -x 12 -d 6 -s 128 --dump fixture/   # CPU: mint
//!   cargo +nightly run --release --features "parallel gpu"  --bin gpu_parity -- -x 12 -d 6 -s 128 --check fixture/  # GPU: compare
//!
//! `KzgProvingKey::{load, dump}` are `todo!()` in the reference (src/commitments/kzg.rs:99-105), so the SRS is not stored: it
//! is the mock setup of (tau, g0, h0) (kzg.rs:84-97) and those three values are written instead, ark-serialize uncompressed.
//! The files under `fixture/` are exactly what this repo's tests/golden/ needs to pin its oracle to the real reference:
//!   points.bin  coefs.bin  r.bin  tau.bin  g0.bin  h0.bin  config.txt  proof.bin
use ark_bls12_381::{Bls12_381 as Ctx, Fr, G1Affine, G2Affine};
use ark_ec::twisted_edwards::Affine;
use ark_ec::CurveConfig;
use ark_ed_on_bls12_381_bandersnatch::BandersnatchConfig;
use ark_ff::{BigInteger256, PrimeField};
use ark_serialize::{CanonicalDeserialize, CanonicalSerialize};
use ark_std::{log2, UniformRand};
use rand::{rngs::StdRng, SeedableRng};
use std::fs;
use std::path::Path;
use GKR_MSM::cleanup::proof_transcript::{ProofTranscript2, TProofTranscript2};
use GKR_MSM::cleanup::protocols::pippenger::benchutils::{run_pippenger, PippengerConfig, PippengerData};
use GKR_MSM::commitments::knuckles::KnucklesProvingKey;
use GKR_MSM::commitments::kzg::KzgProvingKey;

type Fs = <BandersnatchConfig as CurveConfig>::ScalarField;
type Fb = <BandersnatchConfig as CurveConfig>::BaseField;

fn write<T: CanonicalSerialize>(dir: &Path, name: &str, v: &T) {
    let mut buf = vec![];
    v.serialize_uncompressed(&mut buf).unwrap();
    fs::write(dir.join(name), buf).unwrap();
}
fn read<T: CanonicalDeserialize>(dir: &Path, name: &str) -> T {
    T::deserialize_uncompressed(&fs::read(dir.join(name)).unwrap()[..]).unwrap()
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    let get = |flag: &str, default: usize| args.iter().position(|a| a == flag).map(|i| args[i + 1].parse().unwrap()).unwrap_or(default);
    let path = |flag: &str| args.iter().position(|a| a == flag).map(|i| std::path::PathBuf::from(&args[i + 1]));
    let (x_logsize, d_logsize, num_bits, clm, seed) = (get("-x", 10), get("-d", 6), get("-s", 128), get("-c", 0), get("--seed", 1) as u64);

    // the draw order of build_pippenger_data (pippenger.rs:462-497), from a seedable generator instead of thread_rng
    let rng = &mut StdRng::seed_from_u64(seed);
    let (points, coefs, r, tau, g0, h0): (Vec<Affine<BandersnatchConfig>>, Vec<Fs>, Vec<Fb>, Fr, G1Affine, G2Affine) = match path("--check") {
        Some(dir) => (read(&dir, "points.bin"), read(&dir, "coefs.bin"), read(&dir, "r.bin"), read(&dir, "tau.bin"), read(&dir, "g0.bin"), read(&dir, "h0.bin")),
        None => {
            let points = (0..1usize << x_logsize).map(|_| Affine::<BandersnatchConfig>::rand(rng)).collect();
            let coefs = (0..1usize << x_logsize)
                .map(|_| Fs::from_le_bytes_mod_order(&ark_ff::BigInteger::to_bytes_le(&BigInteger256::rand(rng))[..num_bits / 8]))
                .collect();
            let y_logsize = log2((num_bits + d_logsize - 1) / d_logsize) as usize;
            let r = (0..y_logsize).map(|_| Fb::rand(rng)).collect();
            (points, coefs, r, Fr::rand(rng), G1Affine::rand(rng), G2Affine::rand(rng))
        }
    };
    let y_size = (num_bits + d_logsize - 1) / d_logsize;
    let y_logsize = log2(y_size) as usize;
    let comm_size = 1usize << (clm + x_logsize);
    let kzg_pk = KzgProvingKey::<Ctx>::mock_setup(tau, g0, h0, 2 * comm_size - 1);
    let commitment_key = KnucklesProvingKey::new(kzg_pk, clm + x_logsize, Fr::from(2u64));
    let data = PippengerData {
        points: points.clone(),
        coefs: coefs.clone(),
        config: PippengerConfig { y_size, y_logsize, d_logsize, x_logsize, commitment_log_multiplicity: clm },
        r: r.clone(),
        vkey: commitment_key.verifying_key(),
        commitment_key,
    };
    let mut transcript = ProofTranscript2::start_prover(b"fgstglsp");
    let t0 = std::time::Instant::now();
    let _output = run_pippenger(&mut transcript, data);
    let elapsed = t0.elapsed();
    let proof = transcript.end();
    eprintln!("run_pippenger: {:.1} ms, proof {} bytes, gpu feature: {}", elapsed.as_secs_f64() * 1e3, proof.len(), cfg!(feature = "gpu"));

    if let Some(dir) = path("--dump") {
        fs::create_dir_all(&dir).unwrap();
        write(&dir, "points.bin", &points);
        write(&dir, "coefs.bin", &coefs);
        write(&dir, "r.bin", &r);
        write(&dir, "tau.bin", &tau);
        write(&dir, "g0.bin", &g0);
        write(&dir, "h0.bin", &h0);
        fs::write(dir.join("config.txt"), format!("x_logsize {x_logsize}\nd_logsize {d_logsize}\nnum_bits {num_bits}\nclm {clm}\nseed {seed}\n")).unwrap();
        fs::write(dir.join("proof.bin"), &proof).unwrap();
    }
    if let Some(dir) = path("--check") {
        let want = fs::read(dir.join("proof.bin")).unwrap();
        assert!(want == proof, "proof bytes differ from the recorded CPU run ({} vs {} bytes)", want.len(), proof.len());
        eprintln!("proof bytes identical to {}", dir.join("proof.bin").display());
    }
}
