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
// DVS_TRAIL_SHAPE (A/B measurements): how the SM's 65,536 registers are split between the 1,024 counting threads and
// the trailing selection CTA: 0 = 56 registers + 128 x 64, 1 = 48 + 256 x 64, 2 = 40 + 384 x 64
inline int trail_shape() {
    const char* e = getenv("DVS_TRAIL_SHAPE");
    const int v = e ? atoi(e) : 0;
    return v < 0 || v > 2 ? 0 : v;
}
inline int trail_count_regs() { return 56 - 8 * trail_shape(); }
int select_with_trail(dvs_ctx* ctx, const dvs_kfreqs* f, const uint32_t* order, uint32_t num, int mode, uint32_t min_size,
                      uint32_t max_size, uint32_t* sel_idx, double* sel_delta, double* stats5, uint32_t* size_out,
                      const TrailArgs* trail);
}  // namespace dvs
