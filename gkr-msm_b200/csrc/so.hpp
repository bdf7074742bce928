// Base class behind the opaque `gkr_so` handle: the device-side counterpart of
// `trait Sumcheckable<F>` (src/cleanup/protocols/sumchecks/vecvec_eq.rs:218-225).
#pragma once
#include "common.cuh"

struct gkr_so {
    gkr_ctx* ctx = nullptr;
    virtual ~gkr_so() {}
    // evals at nodes 0..deg (deg+1 values); the reference's UniPoly::from_evals is applied by the caller
    virtual int unipoly(gkr::FrH* evals, uint32_t* n_evals) = 0;
    virtual int bind(const gkr::FrH& t) = 0;
    virtual int final_evals(gkr::FrH* out) = 0;
    // Multi-GPU drivers (sharded.cu, deg2.cu): the LINEAR part of this round's message, i.e. what the shards of one object add up
    // before the host derives the round polynomial.  Default: the sums at the nodes 1..deg (dense objects); the Deg2 objects
    // return their two eq-weighted totals (the inputs of from12).  Counts as this round's unipoly() for the call protocol.
    virtual int partial_sums(gkr::FrH* out, uint32_t* n) {
        gkr::FrH ev[GKR_MAX_DEG + 1];
        uint32_t ne = 0;
        int rc = unipoly(ev, &ne);
        if (rc) return rc;
        for (uint32_t s = 1; s < ne; s++) out[s - 1] = ev[s];
        *n = ne - 1;
        return GKR_OK;
    }
    // The caller promises the strict unipoly() -> bind() alternation of GenericSumcheckProtocol::prove with NO other work on the
    // context's stream in between (gkr_sumcheck_prove and the sharded drivers do): the object may then enqueue the kernel of the
    // next small round while the current one runs and hand it the challenge through its mailbox (common.cuh, GkrMailbox).
    virtual void set_prelaunch(bool) {}
    virtual gkr::FrH claim() const = 0;
    virtual uint32_t degree() const = 0;
    virtual uint32_t num_polys() const = 0;
    virtual uint32_t round() const = 0;
};

// factories implemented in the .cu files
int gkr_make_dense_so(gkr_ctx* ctx, int so_kind, int gate, uint32_t gate_param, const gkr::FrH* consts, uint32_t n_consts,
                      gkr_table* const* tables, uint32_t n_polys, uint32_t num_vars, const gkr::FrH& claim, gkr_so** out);
