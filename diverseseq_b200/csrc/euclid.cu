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
// {2tx,2tx+1}+32j); global loads are whole 128-byte lines (8 lanes x 16 bytes per row segment).  The
// distance is accumulated in difference form sum((a-b)^2) — one DADD + one
// DFMA per pair-element, FP64-pipe bound — which has no cancellation (the Gram form
// ||a||^2+||b||^2-2ab loses all digits once d^2 << ||a||^2).  Only tiles on or
// below the diagonal are computed when both row ranges are in the shard; results are mirrored.
#include <math.h>
#include <string.h>

#include "comm.cuh"
#include "common.cuh"

namespace dvs {

constexpr int kEuT = 128;  // tile edge (pairs)
constexpr int kEuK = 16;   // columns per slab
constexpr int kEuPad = 2;  // row padding of the [col][row] slabs (doubles)

// where a tile's distances go: one matrix (plain call) or the same place in the matrix of every GPU
// (dvs_euclid_distances_sharded: the tiles of the lower triangle are dealt over the GPUs and each kernel
// stores its results straight into every peer's window - compute and all-gather in one kernel)
struct EuOuts {
    double* p[kCommMaxWorld];
    int count;
};

// SHARDED: blockIdx.x enumerates this GPU's tiles t = tile_first + tile_step * blockIdx.x of the lower
// triangle (t = ti (ti + 1) / 2 + tj, tj <= ti) of the whole n x n matrix
template <bool SHARDED>
__global__ void __launch_bounds__(256, 1)
k_euclid_tiles(const double* __restrict__ F, uint64_t dim, uint32_t n, uint32_t row_begin, uint32_t row_end,
               const EuOuts outs /* each [(row_end-row_begin)][n] */, uint32_t tile_first, uint32_t tile_step) {
    extern __shared__ double eu_smem[];
    double(*sa)[kEuK][kEuT + kEuPad] = reinterpret_cast<double(*)[kEuK][kEuT + kEuPad]>(eu_smem);
    double(*sb)[kEuK][kEuT + kEuPad] =
        reinterpret_cast<double(*)[kEuK][kEuT + kEuPad]>(eu_smem + 2 * kEuK * (kEuT + kEuPad));
    uint32_t ti = blockIdx.y, tj = blockIdx.x;
    if (SHARDED) {
        const uint64_t t = (uint64_t)tile_first + (uint64_t)tile_step * blockIdx.x;
        ti = (uint32_t)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
        while ((uint64_t)ti * (ti + 1) / 2 > t) --ti;
        while ((uint64_t)(ti + 1) * (ti + 2) / 2 <= t) ++ti;
        tj = (uint32_t)(t - (uint64_t)ti * (ti + 1) / 2);
    }
    const uint32_t i0 = row_begin + ti * kEuT, j0 = tj * kEuT;
    if (i0 >= row_end) return;
    // a tile strictly above the diagonal whose transpose is also produced by this launch is skipped
    // (tiles of rows and columns only coincide when row_begin is tile aligned)
    const bool mirror_in_range = (row_begin % kEuT == 0) && (j0 >= row_begin) && (j0 < row_end);
    if (mirror_in_range && j0 > i0) return;
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    // loader mapping: a slab row segment is 16 doubles = one 128-byte line; chunk = 16 bytes (2 columns).
    // Thread t moves chunks t, t+256, t+512, t+768 of each operand: 8 consecutive lanes read one whole
    // line of one row (4 lines per warp-wide LDG.128).  (The first version gave each thread 8 consecutive
    // doubles of its own row: 32 rows per warp instruction, 16x the L1 wavefronts, and the kernel sat
    // on long-scoreboard stalls at 60 % FP64-pipe utilisation.)
    const bool vec_ok = (dim % 2 == 0);  // 16-byte alignment of every row
    const int lc2 = t & 7;               // chunk within the row segment -> columns 2*lc2, 2*lc2+1
    const double* parow[4];
    const double* pbrow[4];
    bool varow[4], vbrow[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int lr = (t >> 3) + 32 * q;
        const uint32_t ra = i0 + lr, rb = j0 + lr;
        varow[q] = ra < row_end;
        vbrow[q] = rb < n;
        parow[q] = F + (size_t)(varow[q] ? ra : 0) * dim;
        pbrow[q] = F + (size_t)(vbrow[q] ? rb : 0) * dim;
    }

    double acc[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = 0.0;

    double2 ga[4], gb[4];
    auto gload = [&](uint64_t c0) {
        const uint64_t col = c0 + 2 * lc2;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (vec_ok && col + 1 < dim) {
                ga[q] = varow[q] ? *reinterpret_cast<const double2*>(parow[q] + col) : make_double2(0.0, 0.0);
                gb[q] = vbrow[q] ? *reinterpret_cast<const double2*>(pbrow[q] + col) : make_double2(0.0, 0.0);
            } else {
                ga[q].x = (varow[q] && col < dim) ? parow[q][col] : 0.0;
                ga[q].y = (varow[q] && col + 1 < dim) ? parow[q][col + 1] : 0.0;
                gb[q].x = (vbrow[q] && col < dim) ? pbrow[q][col] : 0.0;
                gb[q].y = (vbrow[q] && col + 1 < dim) ? pbrow[q][col + 1] : 0.0;
            }
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int lr = (t >> 3) + 32 * q;
            sa[buf][2 * lc2][lr] = ga[q].x;
            sa[buf][2 * lc2 + 1][lr] = ga[q].y;
            sb[buf][2 * lc2][lr] = gb[q].x;
            sb[buf][2 * lc2 + 1][lr] = gb[q].y;
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
            const bool mirror = mirror_in_range && j >= row_begin && j < row_end && i < n;
            for (int r = 0; r < (SHARDED ? outs.count : 1); ++r) {
                double* out = outs.p[r];
                out[(size_t)(i - row_begin) * n + j] = d;
                if (mirror) out[(size_t)(j - row_begin) * n + i] = d;
            }
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
    DVS_CUDA_TRY(cudaFuncSetAttribute(k_euclid_tiles<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEuSmemBytes));
    EuOuts outs;
    memset(&outs, 0, sizeof outs);
    outs.p[0] = d_out.p;
    outs.count = 1;
    k_euclid_tiles<false><<<grid, 256, kEuSmemBytes, ctx->stream>>>(f->freqs.p, f->dim, (uint32_t)n, row_begin, row_end, outs, 0, 0);
    pt.stop();
    DVS_LAUNCHED(ctx);
    DVS_CUDA_TRY(cudaMemcpyAsync(dist, d_out.p, nrows * n * sizeof(double), cudaMemcpyDefault, ctx->stream));
    DVS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return DVS_OK;
}

// all ranks hold the same rows (dvs_count_kmers_sharded / dvs_kfreqs_allgather); every rank gets the whole matrix
extern "C" int dvs_euclid_distances_sharded(dvs_ctx* ctx, dvs_comm* c, const dvs_kfreqs* f_all, double* dist) {
    if (!ctx || !c || !c->connected || !f_all || !dist) {
        set_error("dvs_euclid_distances_sharded: bad argument / communicator not connected");
        return DVS_ERR_ARG;
    }
    const size_t n = f_all->nrec;
    if (n == 0) return DVS_OK;
    DVS_CUDA_TRY(dvs::enter(ctx));
    cudaStream_t st = ctx->stream;
    uint64_t hoff = 0;
    DVS_TRY(comm_heap_alloc(c, n * n * sizeof(double), &hoff));
    int rc = comm_barrier(ctx, c);  // nobody still uses this block of its window
    cudaError_t e = cudaSuccess;
    if (rc == DVS_OK) {
        const uint64_t nt = (n + kEuT - 1) / kEuT, tiles = nt * (nt + 1) / 2;
        const uint64_t mine = tiles > (uint64_t)c->rank ? (tiles - c->rank + c->world - 1) / c->world : 0;
        EuOuts outs;
        memset(&outs, 0, sizeof outs);
        for (int r = 0; r < c->world; ++r) outs.p[r] = reinterpret_cast<double*>(c->peer[(c->rank + r) % c->world] + hoff);
        outs.count = c->world;
        PhaseTimer pt(ctx, DVS_PHASE_EUCLID);
        e = cudaFuncSetAttribute(k_euclid_tiles<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEuSmemBytes);
        if (e == cudaSuccess && mine) {
            k_euclid_tiles<true><<<(unsigned)mine, 256, kEuSmemBytes, st>>>(f_all->freqs.p, f_all->dim, (uint32_t)n, 0,
                                                                           (uint32_t)n, outs, (uint32_t)c->rank,
                                                                           (uint32_t)c->world);
            ctx->launches++;
            e = cudaGetLastError();
        }
        pt.stop();
        if (e == cudaSuccess) rc = comm_barrier(ctx, c);  // every GPU's tiles have landed in this window
        if (e == cudaSuccess && rc == DVS_OK)
            e = cudaMemcpyAsync(dist, c->window + hoff, n * n * sizeof(double), cudaMemcpyDefault, st);
        if (e == cudaSuccess && rc == DVS_OK) rc = comm_check_error(ctx, c, "dvs_euclid_distances_sharded");
    }
    comm_heap_free(c, hoff);
    if (e != cudaSuccess) {
        set_error("dvs_euclid_distances_sharded: %s", cudaGetErrorString(e));
        return DVS_ERR_CUDA;
    }
    return rc;
}
