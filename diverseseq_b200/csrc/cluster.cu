// Average-linkage agglomerative clustering of a precomputed distance matrix on the device
// (SURVEY.md §8(f) rank 4): the tail of `dvs ctree`.
//
// Reference: diverse_seq/cluster.py:191-237 make_cluster_tree ->
//   sklearn AgglomerativeClustering(metric="precomputed", linkage="average").fit(D).children_
// which (sklearn/cluster/_agglomerative.py linkage_tree, connectivity=None) takes the upper triangle
// D[i, j], i < j, and calls scipy.cluster.hierarchy.linkage(method="average"): the nearest-neighbour
// chain algorithm (scipy/cluster/_hierarchy.pyx nn_chain), a stable sort of the merges by height and a
// union-find relabelling (label()).  Restated here with the same scan order and tie rules so that
// children_ is identical, ties included:
//   * the chain grows from the lowest-numbered live cluster; the neighbour of x is the FIRST live
//     i != x with the smallest D[x, i], except that the previous chain element wins ties;
//   * merged pair (x < y): y keeps the merged cluster, x dies,
//     D[i, y] = (n_x * D[i, x] + n_y * D[i, y]) / (n_x + n_y)  (two products, one sum, one divide; no FMA).
// One persistent CTA walks the chain (it is inherently sequential: ~3n nearest-neighbour scans and n
// row updates, each a pass over one matrix row by 1024 threads); the working matrix is kept fully
// symmetric in HBM so every scan reads one contiguous row.
#include <algorithm>
#include <numeric>

#include "common.cuh"

using dvs::DevBuf;
using dvs::set_error;

namespace {

constexpr int kClThreads = 1024;

struct Merge {
    uint32_t x, y;  // x < y: cluster slots (y survives)
    uint32_t size;
    uint32_t pad;
    double height;
};

// lower triangle <- upper triangle (what sklearn's triu_indices extraction reads), zero diagonal ignored
__global__ void k_cl_symmetrise(double* __restrict__ w, uint32_t n) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t i = blockIdx.y; i < n; i += gridDim.y)
        if (j < i) w[(uint64_t)i * n + j] = w[(uint64_t)j * n + i];
}

struct Best {
    double d;
    uint32_t i;
};

__device__ __forceinline__ Best better(Best a, Best b) { return (b.d < a.d || (b.d == a.d && b.i < a.i)) ? b : a; }

__global__ void __launch_bounds__(kClThreads)
k_cl_nn_chain(double* __restrict__ w, uint32_t n, uint32_t* __restrict__ size_g, uint32_t* __restrict__ chain,
              Merge* __restrict__ merges, int size_in_smem) {
    extern __shared__ uint32_t size_s[];  // [n] cluster sizes, [n] chain (when they fit)
    __shared__ Best red[kClThreads / 32];
    __shared__ uint32_t sh_x, sh_y, sh_len, sh_done, sh_nx, sh_ny, sh_cur, sh_prev;
    __shared__ double sh_dprev;
    volatile uint32_t* sz = size_in_smem ? size_s : size_g;  // volatile: thread 0 updates it between barriers
    volatile uint32_t* vchain = size_in_smem ? size_s + n : chain;  // thread 0 only
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (uint32_t i = tid; i < n; i += kClThreads) sz[i] = 1;
    if (tid == 0) {
        sh_len = 0;
        merges[0].pad = 0;
    }
    uint32_t first_alive = 0;  // thread 0 only: the lowest live slot never moves backwards
    __syncthreads();

    for (uint32_t k = 0; k + 1 < n; ++k) {
        if (tid == 0 && sh_len == 0) {
            while (sz[first_alive] == 0) ++first_alive;
            vchain[0] = first_alive;
            sh_len = 1;
            sh_cur = first_alive;
        }
        __syncthreads();
        while (true) {
            const uint32_t len = sh_len;
            const uint32_t x = sh_cur;  // == chain[len - 1]
            const double* row = w + (uint64_t)x * n;
            if (tid == kClThreads - 1 && len > 1) {  // distance to the previous chain element, fetched alongside the scan
                const uint32_t prev = vchain[len - 2];
                sh_prev = prev;
                sh_dprev = __ldcg(row + prev);
            }
            Best b{__longlong_as_double(0x7ff0000000000000LL), 0xffffffffu};
            for (uint32_t i0 = tid; i0 < n; i0 += 4 * kClThreads) {
                double d[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {  // four independent loads in flight per thread
                    const uint32_t i = i0 + u * kClThreads;
                    const bool ok = i < n && sz[i] != 0 && i != x;
                    d[u] = ok ? __ldcg(row + i) : __longlong_as_double(0x7ff8000000000000LL);  // NaN never wins
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (d[u] < b.d) b = Best{d[u], i0 + u * kClThreads};
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                Best c;
                c.d = __shfl_xor_sync(0xffffffffu, b.d, o);
                c.i = __shfl_xor_sync(0xffffffffu, b.i, o);
                b = better(b, c);
            }
            if (lane == 0) red[wid] = b;
            __syncthreads();
            if (wid == 0) {
                b = red[lane];
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    Best c;
                    c.d = __shfl_xor_sync(0xffffffffu, b.d, o);
                    c.i = __shfl_xor_sync(0xffffffffu, b.i, o);
                    b = better(b, c);
                }
                if (lane == 0) {
                    uint32_t y = b.i;
                    double cur = b.d;
                    uint32_t done = 0;
                    if (len > 1) {
                        const uint32_t prev = sh_prev;
                        const double dprev = sh_dprev;
                        if (!(b.d < dprev)) {  // the previous chain element wins ties
                            y = prev;
                            cur = dprev;
                        }
                        done = (y == prev);
                    }
                    if (y == 0xffffffffu) {  // no finite candidate (NaN / inf everywhere): give up
                        merges[0].pad = 1;
                        sh_done = 2;
                    } else if (!done) {
                        vchain[len] = y;
                        sh_len = len + 1;
                        sh_cur = y;
                    } else {
                        sh_len = len - 2;
                        if (len > 2) sh_cur = vchain[len - 3];
                        const uint32_t lo = min(x, y), hi = max(x, y);
                        const uint32_t nx = sz[lo], ny = sz[hi];
                        merges[k] = Merge{lo, hi, nx + ny, 0, cur};
                        sz[lo] = 0;
                        sz[hi] = nx + ny;
                        sh_x = lo;
                        sh_y = hi;
                        sh_nx = nx;
                        sh_ny = ny;
                    }
                    if (y != 0xffffffffu) sh_done = done;
                }
            }
            __syncthreads();
            if (sh_done) break;
        }
        if (sh_done == 2) return;
        // distances of every live cluster to the merged one (row and column y)
        const uint32_t x = sh_x, y = sh_y;
        const double fx = (double)sh_nx, fy = (double)sh_ny, fs = (double)(sh_nx + sh_ny);
        const double* rx = w + (uint64_t)x * n;
        double* ry = w + (uint64_t)y * n;
        for (uint32_t i0 = tid; i0 < n; i0 += 4 * kClThreads) {
            double dx[4], dy[4];
            bool ok[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t i = i0 + u * kClThreads;
                ok[u] = i < n && sz[i] != 0 && i != y;
                dx[u] = ok[u] ? __ldcg(rx + i) : 0.0;
                dy[u] = ok[u] ? __ldcg(ry + i) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (!ok[u]) continue;
                const uint32_t i = i0 + u * kClThreads;
                const double v = __ddiv_rn(__dadd_rn(__dmul_rn(fx, dx[u]), __dmul_rn(fy, dy[u])), fs);
                ry[i] = v;
                w[(uint64_t)i * n + y] = v;
            }
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int dvs_linkage_average(dvs_ctx* ctx, const double* dist, uint32_t n, int32_t* children, double* heights,
                                   uint32_t* counts) {
    if (!ctx || (n > 1 && (!dist || !children))) {
        set_error("dvs_linkage_average: NULL argument");
        return DVS_ERR_ARG;
    }
    if (n < 2) return DVS_OK;  // nothing to merge
    if (n > (1u << 30)) {
        set_error("dvs_linkage_average: too many observations");
        return DVS_ERR_ARG;
    }
    DVS_CUDA_TRY(dvs::enter(ctx));
    cudaStream_t st = ctx->stream;
    const uint64_t nn = (uint64_t)n * n;
    DevBuf<double> w;
    DevBuf<uint32_t> d_size, d_chain;
    DevBuf<Merge> d_merges;
    DVS_TRY(w.alloc(nn));
    DVS_TRY(d_size.alloc(n));
    DVS_TRY(d_chain.alloc(n));
    DVS_TRY(d_merges.alloc(n - 1));
    // `dist` may live on the host or on the device (unified addressing resolves the direction)
    DVS_CUDA_TRY(cudaMemcpyAsync(w.p, dist, nn * sizeof(double), cudaMemcpyDefault, st));
    PhaseTimer pt(ctx, DVS_PHASE_CLUSTER);
    k_cl_symmetrise<<<dim3((n + 255) / 256, std::min<uint32_t>(n, 32768)), 256, 0, st>>>(w.p, n);
    DVS_LAUNCHED(ctx);
    const size_t smem = 2 * (size_t)n * sizeof(uint32_t);
    const int in_smem = smem <= ctx->smem_optin - 4096;
    if (in_smem && smem > 40 * 1024)
        DVS_CUDA_TRY(cudaFuncSetAttribute(k_cl_nn_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_cl_nn_chain<<<1, kClThreads, in_smem ? smem : 0, st>>>(w.p, n, d_size.p, d_chain.p, d_merges.p, in_smem);
    DVS_LAUNCHED(ctx);
    pt.stop();
    std::vector<Merge> m(n - 1);
    DVS_CUDA_TRY(cudaMemcpyAsync(m.data(), d_merges.p, (n - 1) * sizeof(Merge), cudaMemcpyDeviceToHost, st));
    DVS_CUDA_TRY(cudaStreamSynchronize(st));
    if (m[0].pad) {
        set_error("dvs_linkage_average: the distance matrix has no finite entry to merge on (NaN?)");
        return DVS_ERR_VALUE;
    }

    // stable sort by height, then scipy's label(): union-find that numbers merged clusters n, n+1, ...
    std::vector<uint32_t> order(n - 1);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return m[a].height < m[b].height; });
    std::vector<uint32_t> parent(2 * (size_t)n - 1), csize(2 * (size_t)n - 1, 1);
    std::iota(parent.begin(), parent.end(), 0u);
    auto find = [&](uint32_t x) {
        uint32_t p = x;
        while (parent[x] != x) x = parent[x];
        while (parent[p] != x) {
            const uint32_t nxt = parent[p];
            parent[p] = x;
            p = nxt;
        }
        return x;
    };
    uint32_t next_label = n;
    for (uint32_t r = 0; r + 1 < n; ++r) {
        const Merge& e = m[order[r]];
        const uint32_t xr = find(e.x), yr = find(e.y);
        children[2 * r] = (int32_t)std::min(xr, yr);
        children[2 * r + 1] = (int32_t)std::max(xr, yr);
        parent[xr] = parent[yr] = next_label;
        csize[next_label] = csize[xr] + csize[yr];
        if (heights) heights[r] = e.height;
        if (counts) counts[r] = csize[next_label];
        ++next_label;
    }
    return DVS_OK;
}
