/* gkr_msm_b200.h -- C ABI of the B200-native prover backend for the GKR-MSM hot path.
 *
 * The reference (morgana-proofs/GKR-MSM) is a pure-Rust crate with NO FFI; the narrowest waist every hot
 * loop sits behind is a set of Rust traits.  Each entry point below replaces one trait method / free
 * function of the reference (cited as file:line relative to the reference root) and is what a
 * `#[cfg(feature = "gpu")]` shim in the crate would bind with `extern "C"`.
 *
 * Conventions
 *   - Field elements cross the boundary exactly as the reference stores them: `Fr` =
 *     ark_ff::Fp256<MontBackend<FrConfig,4>> = 4 little-endian u64 limbs in Montgomery form (R = 2^256),
 *     canonical (< r).  A `Vec<Fr>` of n elements is n*4 contiguous u64.  No torch / C++ types here.
 *   - Every function returns 0 on success and a negative gkr_status otherwise; `gkr_last_error` gives
 *     the message.  The reference never returns Result on this path -- it panics (e.g.
 *     src/cleanup/protocols/sumcheck.rs:271-274); the Rust shim turns a non-zero status into panic!().
 *   - One calling thread per context (the reference's protocol code is single-threaded: the transcript
 *     is `&mut`).  Kernels are enqueued on the context's stream; `unipoly`/`final_evals` synchronise.
 *   - There is no CPU fallback: without a CUDA device `gkr_ctx_create` fails with GKR_ERR_CUDA.
 */
#ifndef GKR_MSM_B200_H
#define GKR_MSM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gkr_ctx gkr_ctx;               /* device context: stream, scratch, pinned result slots  */
typedef struct gkr_table gkr_table;           /* device-resident dense table (`Vec<Fr>`)                */
typedef struct gkr_so gkr_so;                 /* a `Sumcheckable` object                                */
typedef struct gkr_transcript gkr_transcript; /* host-side ProofTranscript2 (merlin)                    */
typedef struct gkr_u32buf gkr_u32buf;         /* device-resident u32 array (digits, counters, indices)   */

typedef enum gkr_status {
    GKR_OK = 0,
    GKR_ERR_CUDA = -1,      /* CUDA runtime error / no device                                          */
    GKR_ERR_ARG = -2,       /* invalid argument (reference: assert!/assert_eq! on construction)        */
    GKR_ERR_PROTOCOL = -3,  /* call-order violation (reference: panic in bind()/unipoly())             */
    GKR_ERR_UNSUPPORTED = -4
} gkr_status;

/* Closed gate enum (the reference's `AlgFn` is an open generic, src/cleanup/utils/algfn.rs:20-34). */
typedef enum gkr_gate_id {
    GKR_GATE_AFF_L1 = 0,        /* affine_twisted_edwards_add_l1    4->3  twisted_edwards_ops.rs:10-14  */
    GKR_GATE_AFF_L2 = 1,        /* affine_twisted_edwards_add_l2    3->3  :16-20                        */
    GKR_GATE_AFF_L3 = 2,        /* affine_twisted_edwards_add_l3    3->3  :22-29                        */
    GKR_GATE_PRJ_L1 = 3,        /* twisted_edwards_add_l1           6->4  :31-40                        */
    GKR_GATE_PRJ_L2 = 4,        /* twisted_edwards_add_l2           4->4  :43-52                        */
    GKR_GATE_PRJ_L3 = 5,        /* twisted_edwards_add_l3           4->3  :54-65                        */
    GKR_GATE_TRI_L1 = 6,        /* triangle_twisted_edwards_add_l1 12->12 :67-80                        */
    GKR_GATE_BITCHECK = 7,      /* BitCheckFn                       1->1  algfn.rs:262-291              */
    GKR_GATE_LOGUP_LAYER = 8,   /* LogupLayerFn                     4->2  logup_mainphase.rs:42-61      */
    GKR_GATE_ADD_INVERSES = 9,  /* AddInversesFn                    2->2  pushforward.rs:266-281        */
    GKR_GATE_PROD3 = 10,        /* Prod3Fn (single output, deg 3)   3->1  pushforward.rs:38-50          */
    GKR_GATE_FOLDED_PROD = 11,  /* FoldedProdAlgFn(gamma, nargs)  2n->1  multiopen_reduction.rs:13-41   */
    GKR_GATE_ID = 12,           /* IdAlgFn(n)                       n->n  algfn.rs:130-164              */
    GKR_GATE_AFF_L1_BITCHECK2 = 13 /* Stacked(aff_l1, Repeated(BitCheck,2)) 6->5 bintree_add.rs:259-273 */
} gkr_gate_id;

/* ---- context ------------------------------------------------------------------------------- */
int gkr_ctx_create(int device, gkr_ctx** out);
void gkr_ctx_destroy(gkr_ctx* ctx);
/* release the cached large device blocks and the idle part of the memory pool (between proofs of very different sizes) */
int gkr_ctx_trim(gkr_ctx* ctx);
const char* gkr_last_error(const gkr_ctx* ctx);
int gkr_ctx_sync(gkr_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's `gpu_launches`) */
uint64_t gkr_ctx_launch_count(const gkr_ctx* ctx);
/* raw cudaStream_t of the context (so callers can record CUDA events on the launching stream) */
void* gkr_ctx_stream(gkr_ctx* ctx);
int gkr_version(void);
/* measurement hook (bench.py only): per-launch CUDA-event timing on the context stream.  The integer-pipe probes and
 * the kernel-variant lab live in a separate measurement-only library (csrc/lab/, `make lab`), not in this one. */
int gkr_ctx_timing_enable(gkr_ctx* ctx, int on);
/* test hook: on = 0 makes DenseSumcheckObjectSO::bind always use the full Montgomery product instead of the
 * 128-bit-challenge fold (both are bit-exact; tests compare them at sizes the oracle cannot reach). Default on. */
int gkr_ctx_set_fast_fold(gkr_ctx* ctx, int on);
/* test / experiment hook: the kernel-selection knobs that the environment sets at context creation (GKR_DENSE_SMALL_MAX,
 * GKR_DEG2_COMPACT_MAX, GKR_MSM_SIGNED), by name ("dense_small_max", "deg2_compact_max", "msm_signed").  Every setting is
 * bit-exact; the tests run parity cases under each of them.  Unknown key: GKR_ERR_ARG. */
int gkr_ctx_set_tuning(gkr_ctx* ctx, const char* key, long long value);
/* host-side latency accounting: out = {ns spent inside kernel-launch calls of the round kernels, ns spent waiting for
 * round results, number of waits, kernels launched}; reset != 0 clears the first three. */
int gkr_ctx_host_stats(gkr_ctx* ctx, uint64_t out[4], int reset);
/* Memory pooling over the GPUs of one box: devices [0, n_devices) other than the context's own lend their HBM -- when the home
 * GPU is full, tables of >= 4 MiB are placed on the peer with the most free memory and the kernels (which all run on the home
 * GPU) reach them through NVLink peer access.  Same kernels, same proofs; what spills over runs at NVLink speed.  This is how
 * one prover holds instances beyond 180 GB (BASELINE config[3] at x = 23 / 24).  out_stats (may be NULL): {bytes on peers now,
 * their peak}.  n_devices <= 1 only reads the statistics. */
int gkr_ctx_peer_pool(gkr_ctx* ctx, int n_devices, uint64_t* out_stats);
int gkr_ctx_timing_read(gkr_ctx* ctx, int* kernel_id, uint64_t* n_items, float* ms, int max_n);

/* ---- dense tables: `Vec<Fr>` resident in HBM ---------------------------------------------------
 * Ownership mirrors the reference: sumcheck objects take their tables by value
 * (`ProverInput = Vec<Vec<F>>`, dense_eq.rs:193) but the dense objects never modify the originals on
 * the device (the first fold writes a fresh half-size buffer), so a table stays valid and resident
 * after an object was created from it ("clone before prove", pushforward.rs:682-685, costs nothing). */
int gkr_table_upload(gkr_ctx* ctx, const uint64_t* limbs, uint64_t n, gkr_table** out);
int gkr_table_download(gkr_ctx* ctx, const gkr_table* t, uint64_t* limbs_out);
int gkr_table_alloc(gkr_ctx* ctx, uint64_t n, gkr_table** out);
/* synthetic table: element i = (SplitMix64 stream `seed`, 4 outputs starting at 4*(first_index+i)) mod r,
 * taken as Montgomery limbs.  Used by bench.py so device-resident runs need no PCIe traffic; first_index
 * lets each GPU generate its own shard of one global table. */
int gkr_table_synth(gkr_ctx* ctx, uint64_t seed, uint64_t first_index, uint64_t n, gkr_table** out);
uint64_t gkr_table_len(const gkr_table* t);
void* gkr_table_device_ptr(gkr_table* t);
void gkr_table_free(gkr_table* t);

/* eq_poly_sequence_from_multiplier(mult, point).last()   src/utils.rs:222-262
 * (== EqPoly::evals, src/cleanup/protocols/verifier_polys.rs:31-33, with mult = 1).
 * point: n elements; last coordinate <-> least-significant index bit.  out has 2^n entries. */
int gkr_eq_table(gkr_ctx* ctx, const uint64_t* point, uint32_t n, const uint64_t mult[4], gkr_table** out);

/* sum_i f(tables[0][i], .., tables[P-1][i])  for a single-output gate -- the `claim_hint` the callers of
 * DenseSumcheckObjectSO::new compute on the CPU (e.g. sumcheck.rs:951, pushforward.rs:765-776). */
int gkr_dense_gate_sum(gkr_ctx* ctx, int so_kind, int gate, uint32_t gate_param, const uint64_t* gate_consts,
                       uint32_t n_consts, gkr_table* const* tables, uint32_t n_polys, uint64_t out[4]);

/* ---- trait Sumcheckable  (src/cleanup/protocols/sumchecks/vecvec_eq.rs:218-225) -----------------
 * so_kind selects how the gate is wrapped into a single-output function:                           */
typedef enum gkr_so_kind {
    GKR_SO_PLAIN = 0,    /* gate is already single-output: PROD3, FOLDED_PROD(gate_param = nargs)     */
    GKR_SO_EQ_GAMMA = 1  /* EqWrapper(GammaWrapper(gate, gamma)): last table is the eq table          */
} gkr_so_kind;

/* DenseSumcheckObjectSO::new(polys, f, num_vars, claim_hint)   src/cleanup/protocols/sumcheck.rs:252-261
 *   gate_consts: n_consts field elements; for GKR_SO_EQ_GAMMA element i is gamma^i (i = 0..n_outs-1,
 *   element 0 ignored); for FOLDED_PROD element i is gammas[i] (make_gamma_pows(gamma, nargs)).
 *   tables: n_polys device tables of 2^num_vars entries each (reference asserts both, :255-258). */
int gkr_so_create_dense(gkr_ctx* ctx, int so_kind, int gate, uint32_t gate_param, const uint64_t* gate_consts,
                        uint32_t n_consts, gkr_table* const* tables, uint32_t n_polys, uint32_t num_vars,
                        const uint64_t claim[4], gkr_so** out);

/* Sumcheckable::unipoly: writes the round polynomial as its evaluations at 0..deg ((deg+1)*4 limbs);
 * the shim applies liblasso UniPoly::from_evals.  Dense object: a second call returns the cached
 * value (sumcheck.rs:280-281).  Calling after the last round -> GKR_ERR_PROTOCOL (:278). */
int gkr_so_unipoly(gkr_so* so, uint64_t* evals_out, uint32_t* n_evals);
/* Sumcheckable::bind(t): requires a preceding unipoly() (sumcheck.rs:271-274) else GKR_ERR_PROTOCOL. */
int gkr_so_bind(gkr_so* so, const uint64_t t[4]);
/* Sumcheckable::final_evals: n_polys*4 limbs; only after the last round (sumcheck.rs:336). */
int gkr_so_final_evals(gkr_so* so, uint64_t* out);
/* current running claim (DenseSumcheckObjectSO::claim) */
int gkr_so_claim(const gkr_so* so, uint64_t out[4]);
uint32_t gkr_so_degree(const gkr_so* so);
uint32_t gkr_so_num_polys(const gkr_so* so);
uint32_t gkr_so_round(const gkr_so* so);
void gkr_so_destroy(gkr_so* so);
/* Latency of small rounds.  on != 0 is the caller's promise that the object is driven like GenericSumcheckProtocol::prove
 * (sumcheck.rs:101-123): unipoly, bind, unipoly, bind ... with NO other work on the context -- no table operation and no round of
 * ANOTHER sumcheck object, so not the two-object loop of pushforward.rs:781-806 -- between a unipoly and its bind.  The
 * object may then enqueue the kernel of the next small round while the current one runs and pass it the challenge through a
 * mailbox in mapped host memory, which takes the launch (~5 us of a 12-20 us round) off the critical path; a challenge that does
 * not fit the pre-launched kernel, a destroyed object or a host that stalls for seconds cancel the launch.  Results are the same
 * bits.  gkr_sumcheck_prove and the sharded drivers switch it on for their own loops; GKR_PRELAUNCH=0 disables it globally. */
int gkr_so_set_prelaunch(gkr_so* so, int on);

/* ---- VecVecPolynomial<F> resident in HBM (CSR)   src/cleanup/polys/vecvec.rs:149-206 -------------------
 * gkr_vecvec_upload == VecVecPolynomial::new: `flat` holds the rows back to back (row r has row_len[r]
 * elements); rows of odd length get one row_pad appended (vecvec.rs:183-185). */
typedef struct gkr_vecvec gkr_vecvec;
int gkr_vecvec_upload(gkr_ctx* ctx, const uint64_t* flat, const uint32_t* row_len, uint32_t n_rows, const uint64_t row_pad[4],
                      const uint64_t col_pad[4], uint32_t row_logsize, uint32_t col_logsize, gkr_vecvec** out);
/* the same over rows gathered from a resident table: row r = src[idx[..]] for its row_len[r] consecutive entries of the
 * host array `idx` (src == NULL: the all-ones table) -- the bucket images of PushForwardState::new (pushforward.rs:363-396). */
int gkr_vecvec_gather(gkr_ctx* ctx, const gkr_table* src, const uint32_t* idx, const uint32_t* row_len, uint32_t n_rows,
                      const uint64_t row_pad[4], const uint64_t col_pad[4], uint32_t row_logsize, uint32_t col_logsize, gkr_vecvec** out);
/* n_src polynomials over the same rows in one call (one index upload): srcs[k] may be NULL (all ones); row_pads / col_pads
 * hold n_src x 4 u64; outs receives n_src handles. */
int gkr_vecvec_gather_multi(gkr_ctx* ctx, const gkr_table* const* srcs, uint32_t n_src, const uint32_t* idx, const uint32_t* row_len,
                            uint32_t n_rows, const uint64_t* row_pads, const uint64_t* col_pads, uint32_t row_logsize,
                            uint32_t col_logsize, gkr_vecvec** outs);
/* The same with the padded gather index resident on the device (output of gkr_pushforward_bucketize_dev). */
int gkr_vecvec_gather_multi_dev(gkr_ctx* ctx, const gkr_table* const* srcs, uint32_t n_src, const gkr_u32buf* padded_idx,
                                const uint32_t* row_len, uint32_t n_rows, const uint64_t* row_pads, const uint64_t* col_pads,
                                uint32_t row_logsize, uint32_t col_logsize, gkr_vecvec** outs);
uint32_t gkr_vecvec_num_rows(const gkr_vecvec* v);
uint64_t gkr_vecvec_total_len(const gkr_vecvec* v); /* elements after even-padding */
int gkr_vecvec_download(gkr_ctx* ctx, const gkr_vecvec* v, uint64_t* flat_out, uint32_t* row_len_out, uint64_t row_pad[4],
                        uint64_t col_pad[4], uint32_t* row_logsize, uint32_t* col_logsize);
void gkr_vecvec_free(gkr_vecvec* v);

/* DenseDeg2SumcheckObjectSO::new(polys, func, gamma_pows, claim, point)  sumchecks/dense_eq.rs:75-95
 * The gate is a stack Stacked(Repeated(g_0, r_0), Repeated(g_1, r_1), ..) of base gates (algfn.rs:187-259),
 * e.g. triangle layer k: {TRI_L1 x1, PRJ_L1 xk} (triangle_add.rs:199-231).  gamma_pows: n_outs elements
 * = make_gamma_pows(gamma, n_outs) (src/utils.rs:126-135).  unipoly() returns 4 evaluations (nodes 0..3, the
 * output of UnivarFormat::from12, vecvec_eq.rs:197-216); a second unipoly() in a round is an error like the
 * reference's panic (dense_eq.rs:109-111). */
int gkr_so_create_deg2_dense(gkr_ctx* ctx, const int* part_gate, const uint32_t* part_repeat, uint32_t n_parts,
                             gkr_table* const* tables, uint32_t n_polys, const uint64_t* gamma_pows, const uint64_t claim[4],
                             const uint64_t* point, uint32_t num_vars, gkr_so** out);
/* VecVecDeg2SumcheckObjectSO::new(polys, func, gamma_pows, claim, point, col_logsize)  sumchecks/vecvec_eq.rs:94-118
 * Sparse stage over the row variables, then bind_into_dense (:157-190) hands over to a DenseSumcheckObjectSO on
 * EqWrapper(GammaWrapper(func, gamma)); final_evals() then has n_polys + 1 entries (the eq table last, :443-445).
 * gamma_pows: max(n_outs, 2) elements. */
int gkr_so_create_deg2_vecvec(gkr_ctx* ctx, int gate, gkr_vecvec* const* polys, uint32_t n_polys, const uint64_t* gamma_pows,
                              const uint64_t claim[4], const uint64_t* point, uint32_t num_vars, uint32_t col_logsize, gkr_so** out);

/* ---- trait MapSplit / AlgFnUtils: witness generation  (src/cleanup/polys/common.rs:23-35) ------------------
 * The gate is a stack of repeated base gates as above; GKR_GATE_ID with repeat n is IdAlgFn(n).
 * gkr_map_dense: Vec::algfn_map (dense.rs:141-184) when split_kind < 0, else Vec::algfn_map_split
 *   (dense.rs:115-139) with split_kind 0 = SplitIdx::LO(var_idx), 1 = SplitIdx::HI(var_idx); AlgFnUtils::map_split_hi
 *   (algfn.rs:82-89) is split_kind 1, var_idx 0, bundle_size = n_outs.  `out` receives n_outs (2*n_outs when
 *   splitting) fresh tables in the reference's output order (chunks of bundle_size interleaved left/right).
 * gkr_map_vecvec: mode 0 = vecvec_map (vecvec.rs:480-540) -> gkr_vecvec*; mode 1 = vecvec_map_split at LO(0)
 *   (:542-606) -> gkr_vecvec* with row_logsize - 1; mode 2 = vecvec_map_split_to_dense (:608-654) -> gkr_table*. */
int gkr_map_dense(gkr_ctx* ctx, const int* part_gate, const uint32_t* part_repeat, uint32_t n_parts, gkr_table* const* in,
                  uint32_t n_in, int split_kind, uint32_t var_idx, uint32_t bundle_size, gkr_table** out, uint32_t* n_out);
int gkr_map_vecvec(gkr_ctx* ctx, const int* part_gate, const uint32_t* part_repeat, uint32_t n_parts, gkr_vecvec* const* in,
                   uint32_t n_in, int mode, uint32_t bundle_size, void** out, uint32_t* n_out);

/* ---- commitments: MSM over BLS12-381 G1 ------------------------------------------------------------------
 * gkr_srs_upload: bases resident in HBM.  projective = 0: affine points, 12 u64 each (x then y, Fq Montgomery limbs,
 *   ark `Fp384<MontBackend<FqConfig,6>>`; (0,0) = point at infinity) -- `KzgProvingKey::ptau_1` (kzg.rs:18-22);
 *   projective = 1: Jacobian (X, Y, Z), 18 u64 each -- the bases of msm_nonaff (src/msm_nonaffine.rs:34-38).
 * gkr_msm_g1: <bases[first..first+n), scalars> with `scalars` a device table of Fr (Montgomery form), i.e.
 *   KzgProvingKey::commit(poly) (kzg.rs:123-126; fails like its assert when the slice is too long).  out_xy: the
 *   affine result, 12 u64 (all zero = infinity). */
typedef struct gkr_srs gkr_srs;
int gkr_srs_upload(gkr_ctx* ctx, const uint64_t* points, uint64_t n, int projective, gkr_srs** out);
uint64_t gkr_srs_len(const gkr_srs* s);
void gkr_srs_free(gkr_srs* s);
int gkr_msm_g1(gkr_ctx* ctx, const gkr_srs* srs, uint64_t first, const gkr_table* scalars, uint64_t n, uint64_t* out_xy);
/* KzgProvingKey::mock_setup(tau, g0, _, size).ptau_1  (kzg.rs:84-97): bases[i] = tau^i * g0 as affine points,
 * generated on the device (fixed-base: sum of the 2^j g0 selected by the bits of tau^i). g0_xy: 12 u64, tau: Fr. */
int gkr_srs_mock_setup(gkr_ctx* ctx, const uint64_t tau[4], const uint64_t* g0_xy, uint64_t n, gkr_srs** out);

/* bucket accumulation of the c / d commitments: B[b] = sum_{k: bucket_idx[k] == b} bases[point_idx[k]]
 * (PushForwardState::new, pushforward.rs:398-429, 433-456; Pullback::bucketed_msm, src/pullback.rs:28-59).
 * The result (n_buckets projective points, resident) is a valid `gkr_srs` for gkr_msm_g1 -- that call is then
 * msm_nonaff over the bucket bases (pushforward.rs:598-604).  gkr_g1_weighted_bucket_sum = sum_i i*B[i], the
 * running-sum commitment of pushforward.rs:504-524.  gkr_g1_download_affine normalises points for inspection. */
int gkr_g1_bucket_sums(gkr_ctx* ctx, const gkr_srs* srs, const uint32_t* point_idx, const uint32_t* bucket_idx, uint64_t n,
                       uint32_t n_buckets, gkr_srs** out);
/* the same for the digit / counter matrices of PushForwardState::new resident on the device: idx holds rows of 2^x_logsize
 * entries (row y, column x); incidence (y, x) adds bases[x + 2^x_logsize * (y mod 2^clm)] to bucket
 * ((y >> clm) << group_log) | idx[y][x] (pushforward.rs:401-429 with the row merge of :433-456 folded in). */
int gkr_g1_bucket_sums_rows(gkr_ctx* ctx, const gkr_srs* srs, const gkr_u32buf* idx, uint32_t x_logsize, uint32_t clm, uint32_t group_log,
                            gkr_srs** out);
int gkr_g1_weighted_bucket_sum(gkr_ctx* ctx, const gkr_srs* buckets, uint64_t* out_xy);
/* batched forms for the commitment chunks of one proof (y_size / 2^clm of them): bucket ids (chunk << group_log) | bucket
 * turn all chunks into ONE gkr_g1_bucket_sums call; gkr_g1_weighted_bucket_sums then returns `count` running-sum
 * commitments (count x 12 u64), one per group of 2^group_log buckets starting at `first`; gkr_msm_g1_batch is `n_problems`
 * MSMs with the same scalars over the base ranges [first + p * problem_stride, + n) (msm_nonaff per chunk, :598-604). */
int gkr_g1_weighted_bucket_sums(gkr_ctx* ctx, const gkr_srs* buckets, uint64_t first, uint32_t group_log, uint32_t count,
                                uint64_t* out_xy);
int gkr_msm_g1_batch(gkr_ctx* ctx, const gkr_srs* srs, uint64_t first, uint64_t problem_stride, uint32_t n_problems,
                     const gkr_table* scalars, uint64_t n, uint64_t* out_xy);
int gkr_g1_download_affine(gkr_ctx* ctx, const gkr_srs* pts, uint64_t* out_xy);
/* k commitments with different scalar tables over the SAME base range (the phase-1 commitments p_0, p_1, ac_c, ac_d of
 * pushforward.rs:534-537; the two quotient commitments of opening.rs:77-83): below 2^19 points they share one digit sort, one
 * bucket accumulation and one reduction; larger ones run one by one.  Results equal k gkr_msm_g1 calls.  out_xy: k x 12 u64. */
int gkr_msm_g1_multi(gkr_ctx* ctx, const gkr_srs* srs, uint64_t first, const gkr_table* const* scalars, const uint64_t* n, uint32_t k,
                     uint64_t* out_xy);
/* Old API: commitments to bit columns, src/binary_msm.rs:19-54 (CommitmentKey::commit_bitvec,
 * gkr_msm_simple.rs:62-68).  gkr_binary_msm_prepare = prepare_bases: per chunk of `gamma` bases the 2^gamma - 1 subset sums, affine,
 * laid out [chunk][i - 1]; gkr_binary_msm = binary_msm over prepare_coefs' bytes (HOST array, one per chunk; 0 selects nothing). */
int gkr_binary_msm_prepare(gkr_ctx* ctx, const gkr_srs* bases, uint32_t gamma, gkr_srs** prepared);
int gkr_binary_msm(gkr_ctx* ctx, const gkr_srs* prepared, uint32_t gamma, const uint8_t* coefs, uint64_t n_chunks, uint64_t* out_xy);
/* Proving-key preprocessing (not part of a proof): fixed-base window table T[k][i] = 2^(c k) P_i of an affine SRS, (ceil(255 / c) - 1) n
 * more points in HBM.  With it gkr_msm_g1 sends all windows of a scalar into ONE set of 2^c buckets (13 instead of 16 additions per
 * 255-bit scalar at c = 20, no per-window reductions, no Horner tail).  Used for MSMs with at least 4 * 2^c / windows points. */
int gkr_srs_precompute(gkr_ctx* ctx, gkr_srs* srs, int c);
/* Commitment MSM split by point range over the GPUs of one box (csrc/msm_team.cu).  One process per GPU, every
 * rank holds the same SRS.  The leader (rank 0, `create` = 1, opened first) attaches the team to its context: from then on
 * every gkr_msm_g1 of at least 2^18 points over affine bases is cut into `world` slices; the workers answer from
 * gkr_msm_team_serve (returns GKR_OK after gkr_msm_team_quit, an error after `idle_timeout_s` without a command).  The result
 * is the same point and the same limbs as the single-GPU call. */
typedef struct gkr_msm_team gkr_msm_team;
int gkr_msm_team_open(gkr_ctx* ctx, const char* name, int rank, int world, uint64_t max_n, int create, gkr_msm_team** out);
int gkr_msm_team_serve(gkr_ctx* ctx, gkr_msm_team* team, const gkr_srs* srs, double idle_timeout_s);
int gkr_msm_team_wait_ready(gkr_ctx* ctx, gkr_msm_team* team, double timeout_s);
void gkr_msm_team_quit(gkr_msm_team* team);
int gkr_msm_team_world(const gkr_msm_team* team);
void gkr_msm_team_set_min_n(gkr_ctx* ctx, uint64_t n); /* smallest MSM shared with the team (default 2^18 points) */
void gkr_msm_team_close(gkr_ctx* ctx, gkr_msm_team* team);
/* test hook, host only: the O(windows) tail of every MSM (Horner over the extended-Jacobian window sums X, Y, ZZ, ZZZ --
 * 24 u64 each -- with c doublings per window, then one inversion to affine) runs on the CPU; see csrc/host_g1.hpp. */
int gkr_host_g1_horner(const uint64_t* window_sums, int c, int n_windows, uint64_t* out_xy);
/* host only: 48-byte compressed encoding (ark-bls12-381 / zcash format) of an affine point, as written into the proof */
int gkr_host_g1_serialize(const uint64_t* xy, uint8_t* out48);

/* ---- univariate / element-wise table algebra -----------------------------------------
 * gkr_u32buf: digit / counter arrays resident on the device.
 * gkr_table_from_u32: F::from(v) per entry, negated for the access counts (pushforward.rs:479-500).
 * gkr_table_gather:   out[i] = src[idx[i]] (c_pull / d_pull, pushforward.rs:585-595).
 * gkr_table_lincomb:  out = 0; out[dst_off_k + i] += coef_k * src_k[src_off_k + i] -- p_0 + gamma p_1, combined_witness,
 *                     folded_witness (pippenger.rs:209-231, 274-279), p_lt = lambda t + p (opening.rs:65-75).  A NULL
 *                     src_k is the all-ones table (constant shifts and pads of c_adj / d_adj, pushforward.rs:700-710).
 * gkr_poly_eval / gkr_poly_div_by_linear: `ev`, `div_by_linear` (kzg.rs:73-81, 142-150).
 * gkr_knuckles_*: KnucklesProvingKey::new inverses and compute_t (knuckles.rs:65-81, 111-154). */
typedef struct gkr_knuckles gkr_knuckles;
int gkr_u32_upload(gkr_ctx* ctx, const uint32_t* vals, uint64_t n, gkr_u32buf** out);
int gkr_u32_download(gkr_ctx* ctx, const gkr_u32buf* b, uint32_t* out);
uint64_t gkr_u32_len(const gkr_u32buf* b);
void gkr_u32_free(gkr_u32buf* b);
int gkr_table_from_u32(gkr_ctx* ctx, const gkr_u32buf* v, int negate, gkr_table** out);
int gkr_table_gather(gkr_ctx* ctx, const gkr_table* src, const gkr_u32buf* idx, gkr_table** out);
int gkr_table_lincomb(gkr_ctx* ctx, uint32_t n_terms, gkr_table* const* src, const uint64_t* coefs, const uint64_t* src_off,
                      const uint64_t* dst_off, const uint64_t* len, uint64_t out_len, gkr_table** out);
int gkr_poly_eval(gkr_ctx* ctx, const gkr_table* poly, const uint64_t x[4], uint64_t out[4]);
int gkr_poly_div_by_linear(gkr_ctx* ctx, const gkr_table* poly, const uint64_t pt[4], gkr_table** quotient, uint64_t rem[4]);
int gkr_knuckles_create(gkr_ctx* ctx, uint32_t num_vars, const uint64_t k[4], gkr_knuckles** out);
uint32_t gkr_knuckles_num_vars(const gkr_knuckles* key);
int gkr_knuckles_k(const gkr_knuckles* key, uint64_t out[4]);
void gkr_knuckles_free(gkr_knuckles* key);
int gkr_knuckles_compute_t(gkr_ctx* ctx, const gkr_knuckles* key, const gkr_table* poly, const uint64_t* point, uint32_t n_point,
                           gkr_table** t_out, uint64_t opening[4]);

/* PushForwardState::new index bookkeeping (pushforward.rs:351-396), host only: digit / counter matrices ([y_size][n]), the
 * stable digit order of every row (= bucket contents back to back) and the bucket sizes ([y_size][2^d]).  coefs: n x 4 plain
 * little-endian u64 (the integer value of the Bandersnatch scalar, not Montgomery form). */
int gkr_pushforward_bucketize(const uint64_t* coefs, uint64_t n, uint32_t y_size, uint32_t d_logsize, uint32_t* digits,
                              uint32_t* counter, uint32_t* order, uint32_t* lens);
/* The same bookkeeping on the device (csrc/bucketize.cu: a stable counting sort per digit row): only the scalars cross PCIe.
 * digits / counter: [y_size][n] device arrays; padded_order: the bucket contents back to back with every bucket padded to
 * even length by 0xffffffff (the gather index gkr_vecvec_gather_multi_dev takes); lens: HOST output [y_size][2^d].
 * Supports 1 <= d_logsize <= 13 (GKR_ERR_UNSUPPORTED above: use the host version). */
int gkr_pushforward_bucketize_dev(gkr_ctx* ctx, const uint64_t* coefs, uint64_t n, uint32_t y_size, uint32_t d_logsize, gkr_u32buf** digits,
                                  gkr_u32buf** counter, gkr_u32buf** padded_order, uint32_t* lens);

/* ---- host-side protocol mirror (stand-in for the Rust host while no Rust toolchain exists) --------
 * ProofTranscript2  src/cleanup/proof_transcript.rs:76-147 (merlin 3.0 STROBE-128, label b"" per message) */
int gkr_transcript_new(const uint8_t* label, size_t label_len, gkr_transcript** out);
void gkr_transcript_free(gkr_transcript* t);
int gkr_transcript_write_scalars(gkr_transcript* t, const uint64_t* limbs, uint32_t n);
int gkr_transcript_write_raw(gkr_transcript* t, const uint8_t* msg, size_t len);
int gkr_transcript_challenge(gkr_transcript* t, uint32_t bitsize, uint64_t out[4]);
int gkr_transcript_raw_challenge(gkr_transcript* t, uint8_t* out, size_t len);
size_t gkr_transcript_proof_len(const gkr_transcript* t);
/* old API transcript: `impl TranscriptReceiver / TranscriptSender for merlin::Transcript`  src/transcript.rs:78-101
 * (benches/bintree.rs, gkr_msm_simple): labelled merlin messages outside any proof byte string; append_scalars = one message
 * per scalar with label b""; challenge_scalar(label) = from_le_bytes_mod_order of 64 challenge bytes. */
int gkr_transcript_append_message(gkr_transcript* t, const uint8_t* label, size_t label_len, const uint8_t* msg, size_t len);
int gkr_transcript_append_scalars_old(gkr_transcript* t, const uint64_t* limbs, uint32_t n);
int gkr_transcript_challenge_scalar_old(gkr_transcript* t, const uint8_t* label, size_t label_len, uint64_t out[4]);
int gkr_transcript_proof(const gkr_transcript* t, uint8_t* out);

/* GenericSumcheckProtocol::prove   src/cleanup/protocols/sumcheck.rs:101-123
 * Runs `num_rounds` rounds of (unipoly -> compress -> write_scalars -> challenge(128) -> bind).
 * out_claim: final claim; out_point: num_rounds elements, already reversed (:120);
 * out_final_evals: n_polys elements. */
int gkr_sumcheck_prove(gkr_transcript* t, gkr_so* so, uint32_t num_rounds, uint64_t out_claim[4],
                       uint64_t* out_point, uint64_t* out_final_evals);

/* benchutils::run_pippenger   src/cleanup/protocols/pippenger.rs:499-559 -- the whole prover of `examples/pippenger`:
 * PippengerWG::new (bucketing, images, phase-1 commitments, bintree + triangle witness), the output claims at `r`, and
 * Pippenger::prove (ending GKR, second phase, pushforward with the logup main phase, multiopen reduction, Knuckles opening)
 * written to `transcript`.  Host orchestration in C++ (csrc/protocol.cu), every table-sized step on the device.
 *   srs / g0_xy / knuckles: KnucklesProvingKey (kzg.rs:17-22, knuckles.rs:42-81) with num_vars = x_logsize + clm;
 *   points_x / points_y: 2^x_logsize affine Bandersnatch coordinates (Fr Montgomery limbs); coefs: 2^x_logsize x 4 plain
 *   little-endian u64 (scalars of num_bits bits); r: y_logsize elements, y_size = ceil(num_bits / d_logsize);
 *   dense_output: 3 (d_logsize + 1) tables of 2^y_logsize elements (PippengerOutput::output); claim_evs: their evaluations
 *   at r; pair_xy: the two G1 points of the final pairing check (24 u64).  Output pointers may be NULL. */
int gkr_run_pippenger(gkr_ctx* ctx, gkr_transcript* transcript, const gkr_srs* srs, const uint64_t* g0_xy, const gkr_knuckles* knuckles,
                      const uint64_t* points_x, const uint64_t* points_y, const uint64_t* coefs, uint32_t d_logsize, uint32_t x_logsize,
                      uint32_t num_bits, uint32_t clm, const uint64_t* r, uint64_t* dense_output, uint64_t* claim_evs, uint64_t* pair_xy);

/* ---- multi-GPU: hypercube sharded by its top index bits, one process per GPU ----------
 * gkr_exchange: all-gather of a few field elements between the ranks of one box through POSIX shared
 * memory (the per-round partial sums must reach the host-side transcript anyway).                      */
typedef struct gkr_exchange gkr_exchange;
int gkr_exchange_open(const char* name, int rank, int world, int create, gkr_exchange** out);
void gkr_exchange_close(gkr_exchange* ex);
int gkr_exchange_allgather(gkr_exchange* ex, const uint64_t* mine, uint32_t n_elems, uint64_t* all);
/* GenericSumcheckProtocol::prove over a sharded DenseSumcheckObjectSO: `so` covers this rank's slice
 * (local_rounds variables); ex == NULL means a single GPU.  The gate description is repeated because the
 * last log2(world) rounds run on a small replicated object built from the gathered survivors. */
int gkr_sumcheck_prove_sharded(gkr_transcript* t, gkr_so* so, gkr_exchange* ex, uint32_t local_rounds, int so_kind, int gate,
                               uint32_t gate_param, const uint64_t* gate_consts, uint32_t n_consts,
                               const uint64_t global_claim[4], uint64_t out_claim[4], uint64_t* out_point,
                               uint64_t* out_final_evals);
int gkr_exchange_world(const gkr_exchange* ex);
/* VecVecDeg2SumcheckObjectSO sharded by bucket ROWS (the column variables are the most significant ones and are
 * bound last, src/cleanup/polys/vecvec.rs:150-160): shard g of n_shards (a power of two) holds the rows
 * [g R / n_shards, (g + 1) R / n_shards) of every polynomial, uploaded / gathered / mapped as VecVec handles with
 * col_logsize - log2(n_shards) column variables -- the witness maps (gkr_map_vecvec) are row-local, so each shard builds its
 * layers from its own rows.  point / num_vars / col_logsize describe the WHOLE object. */
int gkr_so_create_deg2_vecvec_shard(gkr_ctx* ctx, int gate, gkr_vecvec* const* polys, uint32_t n_polys, const uint64_t* gamma_pows,
                                    const uint64_t* point, uint32_t num_vars, uint32_t col_logsize, uint32_t shard, uint32_t n_shards,
                                    gkr_so** out);
/* VecVecDeg2Sumcheck::prove (vecvec_eq.rs:447-500 over GenericSumcheckProtocol::prove) with one shard per rank: per sparse
 * round the two eq-weighted totals of every shard are added through the exchange, from12 and Fiat-Shamir run replicated on
 * every rank; the dense tail is gkr_sumcheck_prove_sharded.  ex == NULL: one shard.  The proof bytes, the point
 * (num_vars challenges, reversed) and the n_polys + 1 final evaluations equal the single-GPU object's. */
int gkr_sumcheck_prove_sharded_vecvec(gkr_transcript* t, gkr_so* so, gkr_exchange* ex, const uint64_t global_claim[4],
                                      uint64_t out_claim[4], uint64_t* out_point, uint64_t* out_final_evals);

#ifdef __cplusplus
}
#endif
#endif /* GKR_MSM_B200_H */
