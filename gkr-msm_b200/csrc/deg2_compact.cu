// Latency flavour of the Deg2 round kernel: the same source as deg2.cu's (deg2_kernel.cuh) compiled with the field
// multiplier as an out-of-line call.  Used for rounds of at most GKR_DEG2_COMPACT_MAX_PAIRS pairs (deg2.cu), where one
// launch is bounded by streaming the straight-line multi-precision code through a cold instruction cache.
#define GKR_COMPACT_FIELD
#include "deg2_kernel.cuh"

int gkr_launch_deg2_round_compact(int uniform_gate, const Deg2RoundArgs& a, dim3 grid, unsigned threads, cudaStream_t stream) {
    deg2_compact::launch_deg2_round(uniform_gate, a, grid, threads, stream);
    return 0;
}
