// Euclidean distance matrix over k-mer frequency rows.
//
// Replaces euclidean_distances / euclidean_distance (/root/reference/diverse_seq/distance.py:294-336;
// parallel variant cluster.py:647-680): d(i,j) = ||f_i - f_j||_2 = np.linalg.norm(f_i - f_j),
// symmetric, zero diagonal.  numpy's summation order is unspecified, so parity is 1e-9 relative.
//
// FP64 SIMT tile kernel: a CTA (256 threads) owns a 128x128 tile of pairs, every thread an 8x8
// register tile (64 FP64 accumulators); the frequency rows are staged through double-buffered
// shared memory in 16-column slabs, stored [column][row] so that the inner loop reads its operands
// with conflict-free 16-byte LDS (thread (ty,tx) owns rows {2ty,2ty+1}+32j and columns
// {2tx,2tx+1}+32j).  The distance is accumulated in difference form sum((a-b)^2) — one DADD + one
// DFMA per pair-element, FP64-pipe bound — which has no cancellation (the Gram form
// ||a||^2+||b||^2-2ab loses all digits once d^2 << ||a||^2).  Only tiles on or
// below the diagonal are computed when both row ranges are in the shard; results are mirrored.
#include "common.cuh"

namespace dvs {

constexpr int kEuT = 128;  // tile edge (pairs)
constexpr int kEuK = 16;   // columns per slab
constexpr int kEuPad = 2;  // row padding of the [col][row] slabs (doubles)

__global__ void __launch_bounds__(256, 1)
k_euclid_tiles(const double* __restrict__ F, uint64_t dim, uint32_t n, uint32_t row_begin, uint32_t row_end,
               double* __restrict__ out /* [(row_end-row_begin)][n] */) {
    extern __shared__ double eu_smem[];
    double(*sa)[kEuK][kEuT + kEuPad] = reinterpret_cast<double(*)[kEuK][kEuT + kEuPad]>(eu_smem);
    double(*sb)[kEuK][kEuT + kEuPad] =
        reinterpret_cast<double(*)[kEuK][kEuT + kEuPad]>(eu_smem + 2 * kEuK * (kEuT + kEuPad));
    const uint32_t ti = blockIdx.y, tj = blockIdx.x;
    const uint32_t i0 = row_begin + ti * kEuT, j0 = tj * kEuT;
    if (i0 >= row_end) return;
    // a tile strictly above the diagonal whose transpose is also produced by this launch is skipped
    // (tiles of rows and columns only coincide when row_begin is tile aligned)
    const bool mirror_in_range = (row_begin % kEuT == 0) && (j0 >= row_begin) && (j0 < row_end);
    if (mirror_in_range && j0 > i0) return;
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    // loader mapping: thread -> (row lr of the tile, 8-column half lh of the slab)
    const int lr = t & 127, lh = t >> 7;
    const uint32_t ra = i0 + lr, rb = j0 + lr;
    const bool va = ra < row_end, vb = rb < n;
    const double* pa = F + (size_t)(va ? ra : 0) * dim;
    const double* pb = F + (size_t)(vb ? rb : 0) * dim;

    double acc[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = 0.0;

    double ga[8], gb[8];
    auto gload = [&](uint64_t c0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint64_t col = c0 + lh * 8 + q;
            const bool vc = col < dim;
            ga[q] = (va && vc) ? pa[col] : 0.0;
            gb[q] = (vb && vc) ? pb[col] : 0.0;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            sa[buf][lh * 8 + q][lr] = ga[q];
            sb[buf][lh * 8 + q][lr] = gb[q];
        }
    };
    const uint64_t nslab = (dim + kEuK - 1) / kEuK;
    gload(0);
    sstore(0);
    __syncthreads();
    for (uint64_t sl = 0; sl < nslab; ++sl) {
        const int buf = (int)(sl & 1);
        if (sl + 1 < nslab) gload((sl + 1) * kEuK);  // global loads in flight during the FP64 work
#pragma unroll
        for (int c = 0; c < kEuK; ++c) {
            double av[8], bv[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double2 x = *reinterpret_cast<const double2*>(&sa[buf][c][2 * ty + 32 * j]);
                const double2 y = *reinterpret_cast<const double2*>(&sb[buf][c][2 * tx + 32 * j]);
                av[2 * j] = x.x; av[2 * j + 1] = x.y;
                bv[2 * j] = y.x; bv[2 * j + 1] = y.y;
            }
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const double d = av[a] - bv[b];
                    acc[a][b] = fma(d, d, acc[a][b]);
                }
        }
        if (sl + 1 < nslab) sstore(buf ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const uint32_t i = i0 + 2 * ty + 32 * (a >> 1) + (a & 1);
            const uint32_t j = j0 + 2 * tx + 32 * (b >> 1) + (b & 1);
            if (i >= row_end || j >= n) continue;
            const double d = (i == j) ? 0.0 : sqrt(acc[a][b]);
            out[(size_t)(i - row_begin) * n + j] = d;
            if (mirror_in_range && j >= row_begin && j < row_end && i < n)
                out[(size_t)(j - row_begin) * n + i] = d;
        }
}

constexpr size_t kEuSmemBytes = 4ull * kEuK * (kEuT + kEuPad) * sizeof(double);

}  // namespace dvs

using namespace dvs;

extern "C" int dvs_euclid_distances(dvs_ctx* ctx, const dvs_kfreqs* f, uint32_t row_begin, uint32_t row_end,
                                    double* dist) {
    if (!ctx || !f || !dist || row_begin > row_end || row_end > f->nrec) {
        set_error("dvs_euclid_distances: bad argument");
        return DVS_ERR_ARG;
    }
    const size_t nrows = row_end - row_begin, n = f->nrec;
    if (nrows == 0 || n == 0) return DVS_OK;
    DVS_CUDA_TRY(dvs::enter(ctx));
    DevBuf<double> d_out;
    DVS_TRY(d_out.alloc(nrows * n));
    dim3 grid((unsigned)((n + kEuT - 1) / kEuT), (unsigned)((nrows + kEuT - 1) / kEuT));
    PhaseTimer pt(ctx, DVS_PHASE_EUCLID);
    DVS_CUDA_TRY(cudaFuncSetAttribute(k_euclid_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEuSmemBytes));
    k_euclid_tiles<<<grid, 256, kEuSmemBytes, ctx->stream>>>(f->freqs.p, f->dim, (uint32_t)n, row_begin, row_end, d_out.p);
    pt.stop();
    DVS_LAUNCHED(ctx);
    DVS_CUDA_TRY(cudaMemcpyAsync(dist, d_out.p, nrows * n * sizeof(double), cudaMemcpyDefault, ctx->stream));
    DVS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return DVS_OK;
}
