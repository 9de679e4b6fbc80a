// Euclidean distance matrix over k-mer frequency rows.
//
// Replaces euclidean_distances / euclidean_distance (/root/reference/diverse_seq/distance.py:294-336;
// parallel variant cluster.py:647-680): d(i,j) = ||f_i - f_j||_2 = np.linalg.norm(f_i - f_j),
// symmetric, zero diagonal.  numpy's summation order is unspecified, so parity is 1e-9 relative.
//
// FP64 SIMT tile kernel: a CTA owns a 64x64 tile of pairs, threads own 4x4 sub-tiles, the
// frequency rows are staged through shared memory in 32-column slabs.  The distance is
// accumulated in difference form sum((a-b)^2) — one DADD + one DFMA per pair-element — which has
// no cancellation (the Gram form ||a||^2+||b||^2-2ab loses all digits once d^2 << ||a||^2, see
// DESIGN.md §euclid).  Only tiles on or below the diagonal are computed when both row ranges
// are in the shard; results are mirrored.
#include "common.cuh"

namespace dvs {

constexpr int kEuT = 64;   // tile edge (pairs)
constexpr int kEuK = 32;   // columns per slab

__global__ void __launch_bounds__(256)
k_euclid_tiles(const double* __restrict__ F, uint64_t dim, uint32_t n, uint32_t row_begin, uint32_t row_end,
               double* __restrict__ out /* [(row_end-row_begin)][n] */) {
    __shared__ double sa[kEuK][kEuT + 1];
    __shared__ double sb[kEuK][kEuT + 1];
    const uint32_t ti = blockIdx.y, tj = blockIdx.x;
    const uint32_t i0 = row_begin + ti * kEuT, j0 = tj * kEuT;
    if (i0 >= row_end) return;
    // a tile strictly above the diagonal whose transpose is also produced by this launch is skipped
    // (tiles of rows and columns only coincide when row_begin is tile aligned)
    const bool mirror_in_range = (row_begin % kEuT == 0) && (j0 >= row_begin) && (j0 < row_end);
    if (mirror_in_range && j0 > i0) return;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;

    for (uint64_t c0 = 0; c0 < dim; c0 += kEuK) {
        // stage 64 rows x 32 cols of each operand, transposed to [col][row]
        for (int e = threadIdx.x; e < kEuT * kEuK; e += 256) {
            const int r = e / kEuK, c = e % kEuK;
            const uint64_t col = c0 + c;
            const uint32_t ri = i0 + r, rj = j0 + r;
            sa[c][r] = (ri < row_end && col < dim) ? F[(size_t)ri * dim + col] : 0.0;
            sb[c][r] = (rj < n && col < dim) ? F[(size_t)rj * dim + col] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int c = 0; c < kEuK; ++c) {
            double av[4], bv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) av[a] = sa[c][ty * 4 + a];
#pragma unroll
            for (int b = 0; b < 4; ++b) bv[b] = sb[c][tx * 4 + b];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const double d = av[a] - bv[b];
                    acc[a][b] = fma(d, d, acc[a][b]);
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const uint32_t i = i0 + ty * 4 + a, j = j0 + tx * 4 + b;
            if (i >= row_end || j >= n) continue;
            const double d = (i == j) ? 0.0 : sqrt(acc[a][b]);
            out[(size_t)(i - row_begin) * n + j] = d;
            if (mirror_in_range && j >= row_begin && j < row_end && i < n)
                out[(size_t)(j - row_begin) * n + i] = d;
        }
}

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
    k_euclid_tiles<<<grid, 256, 0, ctx->stream>>>(f->freqs.p, f->dim, (uint32_t)n, row_begin, row_end, d_out.p);
    pt.stop();
    DVS_LAUNCHED(ctx);
    DVS_CUDA_TRY(cudaMemcpyAsync(dist, d_out.p, nrows * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DVS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return DVS_OK;
}
