// C ABI (include/whmec.h) + CUDA kernels of the weighted-MEC / PedMEC column sweep for sm_100a.
//
// Replaces, behind plain C entry points, the work of PedigreeDPTable's constructor
// (src/pedigreedptable.cpp:15-37 -> compute_table :84-174 -> compute_column :177-335) and of
// get_super_reads / get_optimal_partitioning / get_optimal_score (:338-406).
//
// There is deliberately NO CPU execution path in this library: if CUDA is unavailable every
// entry point that computes returns WHMEC_ERR_CUDA.
#include <cuda_runtime.h>
#include <malloc.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/whmec.h"
#include "common.h"
#include "dp_device.h"
#include "pack.h"
#include "tile.cuh"

using namespace whmec;

namespace {

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            msg = std::string(#expr) + ": " + cudaGetErrorString(_e);                               \
            return WHMEC_ERR_CUDA;                                                                  \
        }                                                                                           \
    } while (0)

void set_err(char *err, size_t errlen, const std::string &msg) {
    if (err && errlen) {
        std::strncpy(err, msg.c_str(), errlen - 1);
        err[errlen - 1] = 0;
    }
}

// ------------------------------------------------------------------------------------------
// Column kernel (general path: any pedigree, any read structure).  One launch per column.
// ------------------------------------------------------------------------------------------

// Pack `width`-bit back-pointers of 32 consecutive entries (one per lane) into 32-bit words.
__device__ __forceinline__ void bp_store_warp(uint32_t *arena, uint64_t off_words, uint32_t width, uint64_t e,
                                              uint32_t value, bool valid) {
    if (width == 0) return;  // uniform per launch
    const uint32_t per = 32u / width;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t sub = lane % per;
    uint32_t word = valid ? (width == 32 ? value : (value << (sub * width))) : 0u;
    for (uint32_t off = 1; off < per; off <<= 1) word |= __shfl_xor_sync(0xFFFFFFFFu, word, off);
    if (sub == 0 && valid) arena[off_words + e / per] = word;
}

__device__ __forceinline__ ColView make_view(const ColMeta *m, uint32_t T, uint32_t tb, const uint32_t *fn_c0,
                                             const int32_t *fn_delta, const uint32_t *fn_group, const uint32_t *prev,
                                             uint32_t i) {
    ColView v;
    v.m = m;
    v.T = T;
    v.tb = tb;
    const uint32_t g0 = fn_group[m->grp_off + i], g1 = fn_group[m->grp_off + i + 1];
    v.fn_c0 = fn_c0 + m->fn_off + g0;
    v.fn_delta = fn_delta + (size_t)(m->fn_off + g0) * FN_STRIDE;
    v.nf = g1 - g0;
    v.prev = prev;
    return v;
}

// Thread e = o*T + i evaluates all 2^d candidates of its projection entry (d small).
__global__ void __launch_bounds__(256) col_direct_kernel(const ColMeta *__restrict__ cols, uint32_t k, uint32_t T,
                                                         uint32_t tb, const uint32_t *__restrict__ fn_c0,
                                                         const int32_t *__restrict__ fn_delta,
                                                         const uint32_t *__restrict__ fn_group,
                                                         const uint32_t *__restrict__ prev, uint32_t *__restrict__ out,
                                                         uint32_t *__restrict__ arena) {
    __shared__ ColMeta sm;
    if (threadIdx.x < sizeof(ColMeta) / 4) ((uint32_t *)&sm)[threadIdx.x] = ((const uint32_t *)&cols[k])[threadIdx.x];
    __syncthreads();
    const uint64_t nent = ((uint64_t)1 << sm.f) * T;
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t bpv = 0;
    if (e < nent) {
        const uint32_t o = (uint32_t)(e >> tb), i = (uint32_t)e & (T - 1);
        ColView v = make_view(&sm, T, tb, fn_c0, fn_delta, fn_group, prev, i);
        const uint64_t key = eval_candidates(v, o, i, 0u, 1u << sm.d);
        out[e] = (uint32_t)(key >> 32);
        bpv = (uint32_t)key & low_mask(sm.d + tb);
    }
    bp_store_warp(arena, sm.bp_off, sm.bp_width, e, bpv, e < nent);
}

// Many dropped reads (chain ends): the 2^d candidates of an entry are split over 2^log_chunks
// threads; partial minima meet in a 64-bit atomicMin on the key.
__global__ void __launch_bounds__(256) col_chunk_kernel(const ColMeta *__restrict__ cols, uint32_t k, uint32_t T,
                                                        uint32_t tb, const uint32_t *__restrict__ fn_c0,
                                                        const int32_t *__restrict__ fn_delta,
                                                        const uint32_t *__restrict__ fn_group,
                                                        const uint32_t *__restrict__ prev, uint32_t log_chunks,
                                                        unsigned long long *__restrict__ keys) {
    __shared__ ColMeta sm;
    if (threadIdx.x < sizeof(ColMeta) / 4) ((uint32_t *)&sm)[threadIdx.x] = ((const uint32_t *)&cols[k])[threadIdx.x];
    __syncthreads();
    const uint64_t nent = ((uint64_t)1 << sm.f) * T;
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t e = gid >> log_chunks;
    const uint32_t c = (uint32_t)(gid & ((1ull << log_chunks) - 1));
    unsigned long long key = KEY_INF;
    if (e < nent) {
        const uint32_t o = (uint32_t)(e >> tb), i = (uint32_t)e & (T - 1);
        const uint32_t per = 1u << (sm.d - log_chunks);
        ColView v = make_view(&sm, T, tb, fn_c0, fn_delta, fn_group, prev, i);
        key = eval_candidates(v, o, i, c * per, (c + 1) * per);
    }
    if (log_chunks >= 5) {  // a whole warp works on the same entry: reduce before the atomic
        for (int off = 16; off > 0; off >>= 1) {
            unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, key, off);
            key = other < key ? other : key;
        }
        if ((threadIdx.x & 31) == 0 && e < nent) atomicMin(&keys[e], key);
    } else if (e < nent) {
        atomicMin(&keys[e], key);
    }
}

__global__ void __launch_bounds__(256) col_finalize_kernel(const ColMeta *__restrict__ cols, uint32_t k, uint32_t T,
                                                           uint32_t tb, const unsigned long long *__restrict__ keys,
                                                           uint32_t *__restrict__ out, uint32_t *__restrict__ arena) {
    const ColMeta &m = cols[k];
    const uint64_t nent = ((uint64_t)1 << m.f) * T;
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t bpv = 0;
    if (e < nent) {
        const unsigned long long key = keys[e];
        out[e] = (uint32_t)(key >> 32);
        bpv = (uint32_t)key & low_mask(m.d + tb);
    }
    bp_store_warp(arena, m.bp_off, m.bp_width, e, bpv, e < nent);
}

// Backtrace on the device: the packed back-pointers stay in HBM, only the path comes back.
// T == 1: DP-independent chains are traced by independent threads.  T > 1: one thread walks the
// whole table (transmission values couple the chains, pedigreedptable.cpp:272-297).
__global__ void backtrace_kernel(const ColMeta *__restrict__ cols, const uint32_t *__restrict__ arena, uint32_t T,
                                 uint32_t tb, const uint32_t *__restrict__ chain_begin, uint32_t n_chains, uint32_t n,
                                 const uint32_t *__restrict__ last_vals, uint32_t *__restrict__ path_index,
                                 uint32_t *__restrict__ path_tv, uint32_t *__restrict__ result) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    BtView bv{cols, arena, T, tb};
    if (T == 1) {
        if (c >= n_chains) return;
        const uint32_t k_first = chain_begin[c], k_last = chain_begin[c + 1] - 1;
        const ColMeta &m = cols[k_last];
        const uint32_t bp = bp_load(arena, m.bp_off, m.bp_width, 0);
        const uint32_t x = candidate_index(m, 0, bp);
        backtrace_range(bv, k_last, k_first, x, 0, 0, path_index, path_tv);
        if (k_last == n - 1) result[0] = last_vals[0];
    } else {
        if (c != 0) return;
        uint32_t cost, x, tv, ptv;
        pick_optimum(cols[n - 1], last_vals, arena, T, tb, &cost, &x, &tv, &ptv);
        backtrace_range(bv, n - 1, 0, x, tv, ptv, path_index, path_tv);
        result[0] = cost;
    }
}

// Device buffers come from the device's stream-ordered memory pool with an unlimited release
// threshold: a process that phases many chromosomes pays cudaMalloc for its largest problem once.
template <class Tp>
struct DevBuf {
    Tp *p = nullptr;
    size_t count = 0;
    cudaStream_t stream = nullptr;
    cudaError_t alloc(size_t n, cudaStream_t s) {
        count = n;
        stream = s;
        return cudaMallocAsync((void **)&p, std::max<size_t>(n, 1) * sizeof(Tp), s);
    }
    void release() {
        if (p) cudaFreeAsync(p, stream);
        p = nullptr;
    }
};

void keep_pool_memory(int device) {
    static bool done[64] = {false};
    if (device < 0 || device >= 64 || done[device]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t threshold = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
    }
    done[device] = true;
}

}  // namespace

struct whmec_plan {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    Packed pk;
    whmec_stats stats{};
    bool swept = false;
    bool use_tiles = false;
    // column-kernel buffers
    DevBuf<ColMeta> d_cols;
    DevBuf<uint32_t> d_fn_c0, d_fn_group, d_val[2], d_arena, d_chain_begin, d_path_index, d_path_tv, d_result;
    DevBuf<int32_t> d_fn_delta;
    DevBuf<unsigned long long> d_keys;
    uint32_t last_buf = 0;
    // tile path
    TilePlan tiles;

    ~whmec_plan() {
        cudaSetDevice(device);
        d_cols.release(); d_fn_c0.release(); d_fn_group.release(); d_val[0].release(); d_val[1].release();
        d_arena.release(); d_chain_begin.release(); d_path_index.release(); d_path_tv.release();
        d_result.release(); d_fn_delta.release(); d_keys.release();
        tiles.release(stream);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (stream) {
            cudaStreamSynchronize(stream);
            cudaStreamDestroy(stream);
        }
    }
};

namespace {

// The packed problem and the tile schedule are tens of MB of freshly touched host memory per call;
// returning them to the OS on every free (glibc's default for large blocks) makes each call pay the
// page faults again.  Keep freed blocks in the process heap instead (opt out: WHMEC_KEEP_HOST_MEMORY=0).
void keep_host_memory() {
    static bool done = false;
    if (done) return;
    done = true;
    const char *e = std::getenv("WHMEC_KEEP_HOST_MEMORY");
    if (e && e[0] == '0') return;
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
}

int plan_create_impl(const whmec_problem *p, int device, whmec_plan *pl, std::string &msg) {
    keep_host_memory();
    int rc = pack_problem(p, pl->pk, msg);
    if (rc != WHMEC_OK) return rc;
    Packed &pk = pl->pk;
    pl->device = device;
    pl->stats = pk.stats;
    if (pk.n == 0) return WHMEC_OK;
    for (const ColMeta &m : pk.cols)
        if (m.d + pk.tb > 32) {
            msg = "unsupported: dropped reads + transmission bits exceed 32";
            return WHMEC_ERR_UNSUPPORTED;
        }
    CUDA_TRY(cudaSetDevice(device));
    keep_pool_memory(device);
    CUDA_TRY(cudaStreamCreateWithFlags(&pl->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&pl->ev0));
    CUDA_TRY(cudaEventCreate(&pl->ev1));
    const uint32_t n = pk.n;

    CUDA_TRY(pl->d_path_index.alloc(n, pl->stream));
    CUDA_TRY(pl->d_path_tv.alloc(n, pl->stream));
    CUDA_TRY(pl->d_result.alloc(4, pl->stream));
    CUDA_TRY(cudaEventRecord(pl->ev0, pl->stream));
    uint64_t h2d = 0;

    const char *force = std::getenv("WHMEC_FORCE_COLUMN_KERNEL");  // test hook: exercise the general path on T == 1
    pl->use_tiles = !(force && force[0] == '1') && pl->tiles.plan(pk);
    if (pl->use_tiles) {
        rc = pl->tiles.create(pk, pl->stream, h2d, msg);
        if (rc != WHMEC_OK) return rc;
        pl->stats.path_kind = 1;
        pl->stats.backptr_bytes = pl->tiles.backptr_bytes;
    } else {
        const size_t free_b = device_available_bytes();
        uint64_t max_ent = 1;
        for (const ColMeta &m : pk.cols) max_ent = std::max<uint64_t>(max_ent, ((uint64_t)1 << m.f) * pk.T);
        uint64_t need = pk.bp_words * 4 + max_ent * (4 * 2 + 8) + pk.fn_delta.size() * 4 + (uint64_t)n * sizeof(ColMeta);
        if (need + (512ull << 20) > free_b) {
            msg = "back-pointer storage exceeds the free HBM of this device";
            return WHMEC_ERR_UNSUPPORTED;
        }
        CUDA_TRY(pl->d_cols.alloc(n, pl->stream));
        CUDA_TRY(pl->d_fn_c0.alloc(pk.fn_c0.size(), pl->stream));
        CUDA_TRY(pl->d_fn_delta.alloc(pk.fn_delta.size(), pl->stream));
        CUDA_TRY(pl->d_fn_group.alloc(pk.fn_group.size(), pl->stream));
        CUDA_TRY(pl->d_chain_begin.alloc(pk.chain_begin.size(), pl->stream));
        CUDA_TRY(pl->d_val[0].alloc(max_ent, pl->stream));
        CUDA_TRY(pl->d_val[1].alloc(max_ent, pl->stream));
        CUDA_TRY(pl->d_keys.alloc(max_ent, pl->stream));
        CUDA_TRY(pl->d_arena.alloc(pk.bp_words + 1, pl->stream));
        auto up = [&](void *dst, const void *src, size_t bytes) {
            h2d += bytes;
            return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, pl->stream);
        };
        CUDA_TRY(up(pl->d_cols.p, pk.cols.data(), (size_t)n * sizeof(ColMeta)));
        CUDA_TRY(up(pl->d_fn_c0.p, pk.fn_c0.data(), pk.fn_c0.size() * 4));
        CUDA_TRY(up(pl->d_fn_delta.p, pk.fn_delta.data(), pk.fn_delta.size() * 4));
        CUDA_TRY(up(pl->d_fn_group.p, pk.fn_group.data(), pk.fn_group.size() * 4));
        CUDA_TRY(up(pl->d_chain_begin.p, pk.chain_begin.data(), pk.chain_begin.size() * 4));
        pl->stats.path_kind = 2;
    }
    CUDA_TRY(cudaEventRecord(pl->ev1, pl->stream));
    CUDA_TRY(cudaStreamSynchronize(pl->stream));
    CUDA_TRY(cudaEventElapsedTime(&pl->stats.h2d_ms, pl->ev0, pl->ev1));
    pl->stats.h2d_bytes = h2d;
    return WHMEC_OK;
}

int column_sweep(whmec_plan *pl, std::string &msg) {
    Packed &pk = pl->pk;
    const uint32_t n = pk.n, T = pk.T, tb = pk.tb;
    uint32_t launches = 0;
    uint32_t curb = 0;
    for (uint32_t k = 0; k < n; ++k) {
        const ColMeta &m = pk.cols[k];
        const uint64_t nent = ((uint64_t)1 << m.f) * T;
        const uint32_t *prev = pl->d_val[curb ^ 1].p;
        uint32_t *out = pl->d_val[curb].p;
        if (m.d <= 6) {
            const unsigned blocks = (unsigned)((nent + 255) / 256);
            col_direct_kernel<<<blocks, 256, 0, pl->stream>>>(pl->d_cols.p, k, T, tb, pl->d_fn_c0.p, pl->d_fn_delta.p,
                                                               pl->d_fn_group.p, prev, out, pl->d_arena.p);
            launches += 1;
        } else {
            const uint32_t log_chunks = m.d - 6;
            CUDA_TRY(cudaMemsetAsync(pl->d_keys.p, 0xFF, nent * 8, pl->stream));
            const uint64_t threads = nent << log_chunks;
            const unsigned blocks = (unsigned)((threads + 255) / 256);
            col_chunk_kernel<<<blocks, 256, 0, pl->stream>>>(pl->d_cols.p, k, T, tb, pl->d_fn_c0.p, pl->d_fn_delta.p,
                                                              pl->d_fn_group.p, prev, log_chunks, pl->d_keys.p);
            col_finalize_kernel<<<(unsigned)((nent + 255) / 256), 256, 0, pl->stream>>>(pl->d_cols.p, k, T, tb, pl->d_keys.p,
                                                                                          out, pl->d_arena.p);
            launches += 2;
        }
        curb ^= 1;
    }
    pl->last_buf = curb ^ 1;
    CUDA_TRY(cudaGetLastError());
    pl->stats.kernel_launches = launches;
    return WHMEC_OK;
}

int plan_sweep_impl(whmec_plan *pl, std::string &msg) {
    if (pl->pk.n == 0) {
        pl->swept = true;
        return WHMEC_OK;
    }
    CUDA_TRY(cudaSetDevice(pl->device));
    CUDA_TRY(cudaEventRecord(pl->ev0, pl->stream));
    int rc;
    if (pl->use_tiles) {
        rc = pl->tiles.sweep(pl->pk, pl->stream, msg);
        pl->stats.kernel_launches = pl->tiles.launches;
        pl->stats.state_bytes = pl->tiles.state_bytes;
    } else {
        rc = column_sweep(pl, msg);
    }
    if (rc != WHMEC_OK) return rc;
    CUDA_TRY(cudaEventRecord(pl->ev1, pl->stream));
    CUDA_TRY(cudaStreamSynchronize(pl->stream));
    CUDA_TRY(cudaEventElapsedTime(&pl->stats.sweep_ms, pl->ev0, pl->ev1));
    pl->swept = true;
    return WHMEC_OK;
}

int plan_finish_impl(whmec_plan *pl, whmec_solution *s, std::string &msg) {
    Packed &pk = pl->pk;
    const uint32_t n = pk.n;
    if (n == 0) {  // pedigreedptable.cpp:88-92
        s->cost = 0;
        if (s->partition) std::memset(s->partition, 1, pk.n_reads);
        return WHMEC_OK;
    }
    if (!pl->swept) {
        msg = "whmec_plan_finish called before whmec_plan_sweep";
        return WHMEC_ERR_INPUT;
    }
    CUDA_TRY(cudaSetDevice(pl->device));
    CUDA_TRY(cudaEventRecord(pl->ev0, pl->stream));
    if (pl->use_tiles) {
        int rc = pl->tiles.backtrace(pk, pl->stream, pl->d_path_index.p, pl->d_result.p, msg);
        if (rc != WHMEC_OK) return rc;
        CUDA_TRY(cudaMemsetAsync(pl->d_path_tv.p, 0, (size_t)n * 4, pl->stream));
    } else {
        const uint32_t n_chains = (uint32_t)pk.chain_begin.size() - 1;
        const uint32_t threads = pk.T == 1 ? n_chains : 1;
        backtrace_kernel<<<(threads + 63) / 64, 64, 0, pl->stream>>>(pl->d_cols.p, pl->d_arena.p, pk.T, pk.tb,
                                                                      pl->d_chain_begin.p, n_chains, n, pl->d_val[pl->last_buf].p,
                                                                      pl->d_path_index.p, pl->d_path_tv.p, pl->d_result.p);
        CUDA_TRY(cudaGetLastError());
    }
    std::vector<uint32_t> pidx(n), ptv(n);
    uint32_t result[4] = {0, 0, 0, 0};
    CUDA_TRY(cudaMemcpyAsync(pidx.data(), pl->d_path_index.p, (size_t)n * 4, cudaMemcpyDeviceToHost, pl->stream));
    CUDA_TRY(cudaMemcpyAsync(ptv.data(), pl->d_path_tv.p, (size_t)n * 4, cudaMemcpyDeviceToHost, pl->stream));
    CUDA_TRY(cudaMemcpyAsync(result, pl->d_result.p, 16, cudaMemcpyDeviceToHost, pl->stream));
    CUDA_TRY(cudaEventRecord(pl->ev1, pl->stream));
    CUDA_TRY(cudaStreamSynchronize(pl->stream));
    CUDA_TRY(cudaEventElapsedTime(&pl->stats.d2h_ms, pl->ev0, pl->ev1));
    pl->stats.d2h_bytes = (uint64_t)n * 8 + 16;
    s->cost = result[0];
    return build_outputs(pk, pidx.data(), ptv.data(), s, msg);
}

}  // namespace

extern "C" {

int whmec_abi_version(void) { return WHMEC_ABI_VERSION; }

const char *whmec_build_info(void) { return "whmec: CUDA sm_100a weighted-MEC/PedMEC DP (column + tile kernels), no CPU path"; }

int whmec_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int whmec_plan_create(const whmec_problem *p, int device, whmec_plan **out, char *err, size_t errlen) {
    std::string msg;
    whmec_plan *pl = new whmec_plan();
    int rc = plan_create_impl(p, device, pl, msg);
    if (rc != WHMEC_OK) {
        set_err(err, errlen, msg);
        delete pl;
        *out = nullptr;
        return rc;
    }
    *out = pl;
    return WHMEC_OK;
}

int whmec_plan_sweep(whmec_plan *plan, char *err, size_t errlen) {
    std::string msg;
    int rc = plan_sweep_impl(plan, msg);
    if (rc != WHMEC_OK) set_err(err, errlen, msg);
    return rc;
}

int whmec_plan_finish(whmec_plan *plan, whmec_solution *s, char *err, size_t errlen) {
    std::string msg;
    int rc = plan_finish_impl(plan, s, msg);
    if (rc != WHMEC_OK) set_err(err, errlen, msg);
    return rc;
}

int whmec_plan_stats(const whmec_plan *plan, whmec_stats *st) {
    *st = plan->stats;
    return WHMEC_OK;
}

void whmec_plan_destroy(whmec_plan *plan) { delete plan; }

int whmec_solve(const whmec_problem *p, whmec_solution *s, int device, whmec_stats *st, char *err, size_t errlen) {
    whmec_plan *pl = nullptr;
    int rc = whmec_plan_create(p, device, &pl, err, errlen);
    if (rc != WHMEC_OK) return rc;
    rc = whmec_plan_sweep(pl, err, errlen);
    if (rc == WHMEC_OK) rc = whmec_plan_finish(pl, s, err, errlen);
    if (st) *st = pl->stats;
    whmec_plan_destroy(pl);
    return rc;
}

}  // extern "C"
