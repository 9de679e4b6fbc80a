// fused.cuh - what count.cu and select.cu share for dvs_count_select (counting with the selection rounds trailing)
#pragma once
#include <stdlib.h>

#include "common.cuh"

namespace dvs {
struct TrailArgs {
    const unsigned* d_ready;   // device word: positions of `order` below it have their rows, entropies and flags
    unsigned limit0;           // its value once `first_ready` has happened
    cudaEvent_t first_ready;   // the first chunk of records has been published
    cudaEvent_t count_done;    // all of them have
    const unsigned* d_resident;  // device word: counting CTAs resident right now (placement of the trailing kernel)
};
// DVS_TRAIL_SHAPE: how counting and the trailing selection share the GPU.
//   0..2 = one selection CTA beside the counting CTA on every SM, the SM's 65,536 registers split as
//          1,024 x 56 + 128 x 64 / 1,024 x 48 + 256 x 64 / 1,024 x 40 + 384 x 56;
//   3    = the selection owns DVS_TRAIL_SMS whole SMs (default 36; measured 24..40 on the bench set): the stand-alone 512-thread kernel on a
//          high-priority stream takes the SMs the counting CTAs of a finished launch leave, and the counting goes on
//          with the rest (its CTAs are work-stealing loops, so the ones that find no SM just start late and exit).
inline int trail_shape() {
    const char* e = getenv("DVS_TRAIL_SHAPE");
    const int v = e ? atoi(e) : 3;
    return v < 0 || v > 3 ? 3 : v;
}
inline int trail_count_regs() { return 56 - 8 * (trail_shape() % 3); }
inline unsigned trail_sms(int sm_count) {
    const char* e = getenv("DVS_TRAIL_SMS");
    const int v = e ? atoi(e) : 36;
    return (unsigned)(v < 2 ? 2 : (v > sm_count / 2 ? sm_count / 2 : v));
}
int select_with_trail(dvs_ctx* ctx, const dvs_kfreqs* f, const uint32_t* order, uint32_t num, int mode, uint32_t min_size,
                      uint32_t max_size, uint32_t* sel_idx, double* sel_delta, double* stats5, uint32_t* size_out,
                      const TrailArgs* trail);
}  // namespace dvs
