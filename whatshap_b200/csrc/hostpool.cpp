// Persistent host worker pool (see hostpool.h).
#include "hostpool.h"

#include <pthread.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <exception>
#include <mutex>
#include <new>
#include <thread>

namespace whmec {

namespace {
std::atomic<void *(*)(size_t)> g_stage_alloc{nullptr};
std::atomic<bool (*)(void *)> g_stage_release{nullptr};
}  // namespace

void set_stage_hooks(const StageHooks &hooks) {
    g_stage_release.store(hooks.release);
    g_stage_alloc.store(hooks.alloc);
}

void *stage_alloc(size_t bytes) {
    if (auto fn = g_stage_alloc.load(std::memory_order_acquire))
        if (void *p = fn(bytes)) return p;
    return ::operator new(bytes);
}

void stage_free(void *p) {
    if (!p) return;
    if (auto fn = g_stage_release.load(std::memory_order_acquire))
        if (fn(p)) return;
    ::operator delete(p);
}

bool stage_flush_enabled() {
    static const bool on = [] {
        const char *e = std::getenv("WHMEC_FLUSH_UPLOAD");
        return e && e[0] == '1';
    }();
    return on && g_stage_alloc.load(std::memory_order_relaxed) != nullptr;  // only page-locked arrays are read by the DMA engine directly
}

void stage_flush(const void *p, size_t bytes) {
#if defined(__x86_64__)
    const uintptr_t a = (uintptr_t)p & ~(uintptr_t)63, e = (uintptr_t)p + bytes;
    for (uintptr_t q = a; q < e; q += 64) __builtin_ia32_clflush((const void *)q);
    __builtin_ia32_sfence();
#else
    (void)p;
    (void)bytes;
#endif
}

uint32_t host_threads(uint32_t cap) {
    uint32_t cores = std::max(1u, std::thread::hardware_concurrency());
    // one process per GPU (torchrun exports LOCAL_WORLD_SIZE): the ranks of a box share its cores
    if (const char *e = std::getenv("LOCAL_WORLD_SIZE")) {
        const int ranks = std::atoi(e);
        if (ranks > 1) cores = std::max(1u, cores / (uint32_t)ranks);
    }
    uint32_t hw = std::min(cap, cores);
    if (const char *e = std::getenv("WHMEC_HOST_THREADS")) hw = (uint32_t)std::max(1, std::atoi(e));
    return hw;
}

namespace {

constexpr uint32_t MAX_WORKERS = 63;

struct Job {
    const std::function<void(uint32_t)> *fn;
    uint32_t n;
    std::atomic<uint32_t> next{0};
    std::atomic<int> slots{0};  // workers that may still join this job
    std::atomic<bool> failed{false};
    std::mutex err_m;
    std::exception_ptr err;  // first exception thrown by a task (rethrown on the caller once every worker has checked out)
};

// A task that throws (std::bad_alloc from a packer chunk, ...) must neither terminate a pool worker nor unwind the caller
// while workers still hold the stack-allocated Job: the first exception is kept, the remaining tasks are skipped.
void drain(Job &j) {
    for (uint32_t t = j.next.fetch_add(1, std::memory_order_relaxed); t < j.n; t = j.next.fetch_add(1, std::memory_order_relaxed)) {
        if (j.failed.load(std::memory_order_relaxed)) continue;
        try {
            (*j.fn)(t);
        } catch (...) {
            std::lock_guard<std::mutex> lk(j.err_m);
            if (!j.err) j.err = std::current_exception();
            j.failed.store(true, std::memory_order_relaxed);
        }
    }
}

void rethrow(Job &j) {
    if (j.err) std::rethrow_exception(j.err);
}

struct Pool {
    std::mutex entry;  // one job at a time
    std::mutex m;
    std::condition_variable cv, done_cv;
    uint32_t n_workers = 0;
    Job *job = nullptr;
    std::atomic<uint64_t> gen{0};
    uint32_t running = 0;  // workers that have not yet checked out of the current job

    void worker(uint64_t seen) {
        for (;;) {
            Job *j;
            {
                // the host phases of one solve follow each other within microseconds: poll briefly before sleeping
                const auto t0 = std::chrono::steady_clock::now();
                while (gen.load(std::memory_order_acquire) == seen &&
                       std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(150)) {
#if defined(__x86_64__)
                    __builtin_ia32_pause();
#endif
                }
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [&] { return gen.load() != seen; });
                seen = gen.load();
                j = job;
            }
            if (j->slots.fetch_sub(1, std::memory_order_relaxed) > 0) drain(*j);
            {
                std::lock_guard<std::mutex> lk(m);
                if (--running == 0) done_cv.notify_one();
            }
        }
    }
};

std::atomic<Pool *> g_pool{nullptr};
std::once_flag g_atfork;

// The workers do not exist in a forked child: start over with an empty pool there (the old one is leaked).
void after_fork_in_child() { g_pool.store(nullptr); }

Pool *pool() {
    Pool *p = g_pool.load(std::memory_order_acquire);
    if (p) return p;
    std::call_once(g_atfork, [] { pthread_atfork(nullptr, nullptr, after_fork_in_child); });
    Pool *fresh = new Pool();
    if (g_pool.compare_exchange_strong(p, fresh)) return fresh;
    delete fresh;
    return p;
}

void run_on_temporary_threads(uint32_t n_threads, Job &j) {
    std::vector<std::thread> th;
    for (uint32_t t = 0; t + 1 < n_threads; ++t) th.emplace_back([&j] { drain(j); });
    drain(j);
    for (auto &t : th) t.join();
    rethrow(j);
}

}  // namespace

void parallel_tasks(uint32_t n_tasks, uint32_t n_threads, const std::function<void(uint32_t)> &fn) {
    if (n_tasks == 0) return;
    n_threads = std::min(std::min(n_threads, n_tasks), MAX_WORKERS + 1);
    if (n_threads <= 1) {
        for (uint32_t t = 0; t < n_tasks; ++t) fn(t);
        return;
    }
    Job j;
    j.fn = &fn;
    j.n = n_tasks;
    Pool *P = pool();
    std::unique_lock<std::mutex> entry(P->entry, std::try_to_lock);
    if (!entry.owns_lock()) {
        run_on_temporary_threads(n_threads, j);
        return;
    }
    j.slots.store((int)n_threads - 1);
    {
        std::lock_guard<std::mutex> lk(P->m);
        while (P->n_workers < n_threads - 1) {
            std::thread(&Pool::worker, P, P->gen.load()).detach();
            ++P->n_workers;
        }
        P->job = &j;
        P->running = P->n_workers;
        ++P->gen;
    }
    P->cv.notify_all();
    drain(j);
    {
        std::unique_lock<std::mutex> lk(P->m);
        P->done_cv.wait(lk, [&] { return P->running == 0; });
        P->job = nullptr;
    }
    entry.unlock();
    rethrow(j);
}

}  // namespace whmec
