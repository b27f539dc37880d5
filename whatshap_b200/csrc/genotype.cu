// whmec_genotype: forward-backward genotype likelihoods (the reference's GenotypeDPTable,
// src/genotypedptable.cpp) on the device.  Sibling DP of the MEC sweep: same packer,
// same column / projection structure, sums of products in double precision instead of u32 min-plus.
//
// One thread per DP cell (bipartition x of a column); a cell adds its contribution to the projection entry it
// maps to with a double-precision atomicAdd (several cells share an entry when reads end / start in the column).
// Per column: backward pass  gl_backward_kernel + gl_scale_kernel,  forward pass  gl_forward_kernel + gl_scale_kernel.
// The backward tables of all columns of a group stay in HBM (the reference keeps every sqrt(n)-th and recomputes,
// genotypedptable.cpp:139-166,326-343).  First correct version: launch-bound (4 small launches per column);
// batching the chains of a single individual into one launch per step is the next cut.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <string>
#include <vector>

#include "gl_pack.h"
#include "tile.cuh"  // device_available_bytes

using namespace whmec;

namespace {

#define CUDA_TRY(expr)                                                    \
    do {                                                                  \
        cudaError_t _e = (expr);                                          \
        if (_e != cudaSuccess) {                                          \
            msg = std::string(#expr) + ": " + cudaGetErrorString(_e);     \
            return WHMEC_ERR_CUDA;                                        \
        }                                                                 \
    } while (0)

constexpr int GL_THREADS = 256;

struct AtomicAdd {
    __device__ void operator()(double *addr, double v) const { atomicAdd(addr, v); }
};

__global__ void __launch_bounds__(GL_THREADS) gl_backward_kernel(GlView v, uint32_t k, const double *beta_k, double *out) {
    const uint64_t x = (uint64_t)blockIdx.x * GL_THREADS + threadIdx.x;
    if (x >= ((uint64_t)1 << v.cols[k].a)) return;
    gl_backward_cell(v, k, (uint32_t)x, beta_k, out, AtomicAdd());
}

// lacc_out: [n_ind * 3] posterior accumulators of column k
__global__ void __launch_bounds__(GL_THREADS) gl_forward_kernel(GlView v, uint32_t k, const double *prev, double *cur, const double *beta_k,
                                                                double *lacc_out) {
    __shared__ double s_acc[GL_MAX_IND * 3];
    const uint32_t n_acc = v.n_ind * 3;
    if (threadIdx.x < n_acc) s_acc[threadIdx.x] = 0.0;
    __syncthreads();
    double lacc[GL_MAX_IND * 3];
    for (uint32_t e = 0; e < n_acc; ++e) lacc[e] = 0.0;
    const uint64_t x = (uint64_t)blockIdx.x * GL_THREADS + threadIdx.x;
    if (x < ((uint64_t)1 << v.cols[k].a)) gl_forward_cell(v, k, (uint32_t)x, prev, cur, beta_k, lacc, AtomicAdd());
    for (uint32_t e = 0; e < n_acc; ++e) {
        double s = lacc[e];
        for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xFFFFFFFFu, s, off);
        if ((threadIdx.x & 31) == 0 && s != 0.0) atomicAdd(&s_acc[e], s);
    }
    __syncthreads();
    if (threadIdx.x < n_acc && s_acc[threadIdx.x] != 0.0) atomicAdd(&lacc_out[threadIdx.x], s_acc[threadIdx.x]);
}

// One block: divides the finished projection column v[0..n) by its largest entry (gl_scale_host) and clears the
// buffer the next column accumulates into.
__global__ void __launch_bounds__(1024) gl_scale_kernel(double *v, uint64_t n, double *clear, uint64_t n_clear) {
    __shared__ double s_max[32];
    double mx = 0.0;
    for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) mx = fmax(mx, v[i]);
    for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_down_sync(0xFFFFFFFFu, mx, off));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x < 32) {
        mx = threadIdx.x < (blockDim.x >> 5) ? s_max[threadIdx.x] : 0.0;
        for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_down_sync(0xFFFFFFFFu, mx, off));
        if (threadIdx.x == 0) s_max[0] = mx;
    }
    __syncthreads();
    mx = s_max[0];
    if (mx > 0.0) {
        const double inv = 1.0 / mx;
        for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) v[i] *= inv;
    }
    for (uint64_t i = threadIdx.x; i < n_clear; i += blockDim.x) clear[i] = 0.0;
}

struct Buffers {  // freed on every exit path
    cudaStream_t stream = nullptr;
    std::vector<void *> ptrs;
    template <class Tp>
    cudaError_t alloc(Tp **p, size_t count) {
        cudaError_t e = cudaMallocAsync((void **)p, std::max<size_t>(count, 1) * sizeof(Tp), stream);
        if (e == cudaSuccess) ptrs.push_back((void *)*p);
        return e;
    }
    ~Buffers() {
        for (void *p : ptrs) cudaFreeAsync(p, stream);
        if (stream) {
            cudaStreamSynchronize(stream);
            cudaStreamDestroy(stream);
        }
    }
};

int genotype_impl(const whmec_problem *p, double *likelihoods, int device, whmec_stats *st, std::string &msg) {
    if (!p || !likelihoods) {
        msg = "null argument";
        return WHMEC_ERR_INPUT;
    }
    Packed pk;
    GlPacked g;
    int rc = gl_pack(p, pk, g, msg);
    if (rc != WHMEC_OK) return rc;
    if (st) *st = pk.stats;
    const uint32_t n = pk.n, T = pk.T, n_ind = pk.n_ind;
    if (n == 0) return WHMEC_OK;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        msg = "no CUDA device (whatshap_b200 has no CPU path)";
        return WHMEC_ERR_CUDA;
    }
    CUDA_TRY(cudaSetDevice(device));
    Buffers B;
    CUDA_TRY(cudaStreamCreateWithFlags(&B.stream, cudaStreamNonBlocking));
    cudaStream_t s = B.stream;

    // groups of whole tables (T == 1: chains; otherwise the one table) whose backward tables fit the device together
    const size_t free_b = device_available_bytes();
    const uint64_t fixed = (uint64_t)n * sizeof(GlCol) + g.eps.size() * 10 + (g.trans.size() + g.q.size()) * 8 + 2 * g.max_proj * 8 +
                           (uint64_t)n * n_ind * 3 * 8 + (512ull << 20);
    if (fixed > free_b) {
        msg = "genotyping: problem exceeds the free HBM of this device";
        return WHMEC_ERR_UNSUPPORTED;
    }
    const uint64_t budget = (free_b - fixed) / 8;  // doubles available for backward tables
    std::vector<uint32_t> group_begin{0};
    {
        uint64_t run = 0;
        uint32_t table_begin = 0;
        for (uint32_t k = 0; k < n; ++k) {
            const uint64_t cost = g.cols[k].last ? 0 : ((uint64_t)1 << g.cols[k].f) * T;
            run += cost;
            if (g.cols[k].last) {  // a table ends here
                uint64_t table = 0;
                for (uint32_t q = table_begin; q <= k; ++q) table += g.cols[q].last ? 0 : ((uint64_t)1 << g.cols[q].f) * T;
                if (table > budget) {
                    msg = "genotyping: the backward tables of one chain exceed the free HBM of this device";
                    return WHMEC_ERR_UNSUPPORTED;
                }
                if (run > budget) {  // close the group before this table
                    group_begin.push_back(table_begin);
                    run = table;
                }
                table_begin = k + 1;
            }
        }
        group_begin.push_back(n);
    }
    uint64_t max_group = 0;
    for (size_t q = 0; q + 1 < group_begin.size(); ++q) {
        const uint32_t lo = group_begin[q], hi = group_begin[q + 1];
        const uint64_t end = g.cols[hi - 1].beta_off;  // the last column of a group ends a table: it owns no entries
        max_group = std::max(max_group, end - g.cols[lo].beta_off);
    }

    GlCol *d_cols = nullptr;
    double *d_eps = nullptr, *d_trans = nullptr, *d_q = nullptr, *d_beta = nullptr, *d_F[2] = {nullptr, nullptr}, *d_acc = nullptr;
    uint8_t *d_allele = nullptr, *d_ind = nullptr;
    int8_t *d_h2p = nullptr;
    CUDA_TRY(B.alloc(&d_cols, n));
    CUDA_TRY(B.alloc(&d_eps, g.eps.size()));
    CUDA_TRY(B.alloc(&d_allele, pk.act_allele.size()));
    CUDA_TRY(B.alloc(&d_ind, pk.act_ind.size()));
    CUDA_TRY(B.alloc(&d_h2p, pk.h2p.size()));
    CUDA_TRY(B.alloc(&d_trans, g.trans.size()));
    CUDA_TRY(B.alloc(&d_q, g.q.size()));
    CUDA_TRY(B.alloc(&d_beta, max_group + 1));
    CUDA_TRY(B.alloc(&d_F[0], g.max_proj));
    CUDA_TRY(B.alloc(&d_F[1], g.max_proj));
    CUDA_TRY(B.alloc(&d_acc, (size_t)n * n_ind * 3));
    uint64_t h2d = 0;
    auto up = [&](void *dst, const void *src, size_t bytes) {
        h2d += bytes;
        return bytes ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s) : cudaSuccess;
    };
    CUDA_TRY(up(d_cols, g.cols.data(), (size_t)n * sizeof(GlCol)));
    CUDA_TRY(up(d_eps, g.eps.data(), g.eps.size() * 8));
    CUDA_TRY(up(d_allele, pk.act_allele.data(), pk.act_allele.size()));
    CUDA_TRY(up(d_ind, pk.act_ind.data(), pk.act_ind.size()));
    CUDA_TRY(up(d_h2p, pk.h2p.data(), pk.h2p.size()));
    CUDA_TRY(up(d_trans, g.trans.data(), g.trans.size() * 8));
    CUDA_TRY(up(d_q, g.q.data(), g.q.size() * 8));
    CUDA_TRY(cudaMemsetAsync(d_acc, 0, (size_t)n * n_ind * 3 * 8, s));
    CUDA_TRY(cudaMemsetAsync(d_F[0], 0, g.max_proj * 8, s));
    CUDA_TRY(cudaMemsetAsync(d_F[1], 0, g.max_proj * 8, s));
    const GlView v{d_cols, d_eps, d_allele, d_ind, d_h2p, d_trans, d_q, T, pk.P, n_ind};

    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    CUDA_TRY(cudaEventCreate(&ev0));
    CUDA_TRY(cudaEventCreate(&ev1));
    CUDA_TRY(cudaEventRecord(ev0, s));
    uint32_t launches = 0;
    auto blocks_of = [&](uint32_t k) { return (unsigned)((((uint64_t)1 << g.cols[k].a) + GL_THREADS - 1) / GL_THREADS); };
    for (size_t q = 0; q + 1 < group_begin.size(); ++q) {
        const uint32_t lo = group_begin[q], hi = group_begin[q + 1];
        const uint64_t base = g.cols[lo].beta_off;
        // column records hold offsets into the layout of ALL backward tables; the device holds this group's, from `base` on
        auto table_of = [&](uint32_t k) { return d_beta + (g.cols[k].beta_off - base); };
        CUDA_TRY(cudaMemsetAsync(d_beta, 0, (g.cols[hi - 1].beta_off - base + 1) * 8, s));
        for (uint32_t k = hi - 1; k > lo; --k) {
            if (g.cols[k].first) continue;  // nothing enters the first column of a table from the left
            double *out = table_of(k - 1);
            gl_backward_kernel<<<blocks_of(k), GL_THREADS, 0, s>>>(v, k, table_of(k), out);
            gl_scale_kernel<<<1, 1024, 0, s>>>(out, ((uint64_t)1 << g.cols[k - 1].f) * T, nullptr, 0);
            launches += 2;
        }
        for (uint32_t k = lo; k < hi; ++k) {
            double *cur = d_F[k & 1], *prev = d_F[(k + 1) & 1];
            gl_forward_kernel<<<blocks_of(k), GL_THREADS, 0, s>>>(v, k, prev, cur, table_of(k), d_acc + (size_t)k * n_ind * 3);
            // scale F_k (the last column of a table writes none) and clear the buffer column k+1 accumulates into
            gl_scale_kernel<<<1, 1024, 0, s>>>(cur, g.cols[k].last ? 0 : ((uint64_t)1 << g.cols[k].f) * T, prev, g.max_proj);
            launches += 2;
        }
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(ev1, s));
    std::vector<double> acc((size_t)n * n_ind * 3);
    CUDA_TRY(cudaMemcpyAsync(acc.data(), d_acc, acc.size() * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, ev0, ev1));
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    gl_normalise(acc.data(), n, n_ind, likelihoods);
    if (st) {
        st->kernel_launches = launches;
        st->sweep_ms = ms;
        st->h2d_bytes = h2d;
        st->d2h_bytes = acc.size() * 8;
        st->backptr_bytes = g.beta_doubles * 8;  // here: bytes of backward tables kept in HBM
        st->path_kind = 4;
    }
    return WHMEC_OK;
}

}  // namespace

extern "C" int whmec_genotype(const whmec_problem *p, double *likelihoods, int device, whmec_stats *st, char *err, size_t errlen) {
    std::string msg;
    int rc;
    try {
        rc = genotype_impl(p, likelihoods, device, st, msg);
    } catch (const std::exception &e) {  // no C++ exception crosses the C boundary
        msg = e.what();
        rc = WHMEC_ERR_INPUT;
    }
    if (rc != WHMEC_OK && err && errlen) {
        std::strncpy(err, msg.c_str(), errlen - 1);
        err[errlen - 1] = 0;
    }
    return rc;
}
