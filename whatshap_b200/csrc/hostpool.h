// Host-side helpers of the packer / planner / output pass: a persistent worker pool (the three host
// phases of one whmec_solve are each a fraction of a millisecond per thread, so spawning and joining
// threads per phase cost as much as the work) and vectors that do not zero-fill on resize (the big
// per-column arrays are written exactly once, by the workers).
#pragma once
#include <stdint.h>

#include <functional>
#include <memory>
#include <utility>
#include <vector>

namespace whmec {

// Threads a host phase may use: WHMEC_HOST_THREADS if set, else min(cap, hardware concurrency); >= 1.
uint32_t host_threads(uint32_t cap);

// Runs fn(task) for task = 0 .. n_tasks-1 on up to n_threads threads (the caller is one of them), tasks
// handed out dynamically.  Returns when all tasks are done.  Re-entrant: a second caller that finds the
// pool busy runs on short-lived threads of its own.  If a task throws, the remaining tasks are skipped and the first
// exception is rethrown on the caller after every worker has let go of the job.
void parallel_tasks(uint32_t n_tasks, uint32_t n_threads, const std::function<void(uint32_t)> &fn);

// std::allocator whose value-initialisation is default-initialisation: resize() leaves trivial
// elements untouched instead of writing zeros over memory the packer fills right afterwards.
template <class T>
struct NoInitAlloc : std::allocator<T> {
    template <class U>
    struct rebind {
        using other = NoInitAlloc<U>;
    };
    NoInitAlloc() = default;
    template <class U>
    NoInitAlloc(const NoInitAlloc<U> &) {}
    template <class U>
    void construct(U *p) {
        ::new ((void *)p) U;
    }
    template <class U, class... A>
    void construct(U *p, A &&...a) {
        ::new ((void *)p) U(std::forward<A>(a)...);
    }
};
template <class T>
using RawVec = std::vector<T, NoInitAlloc<T>>;

// Arrays that are uploaded to the device as they are (per-column records, cost functions, panels): their memory comes
// from a hook that the CUDA side of the library installs (a pool of page-locked blocks, whmec.cu), so that the packer and the
// planner write straight into DMA-able memory and the upload is one asynchronous copy per array with no staging pass.
// Without a hook (host-only builds: tests/emul) and for small arrays it is plain heap memory.
struct StageHooks {
    void *(*alloc)(size_t bytes);      // nullptr result: fall back to the heap
    bool (*release)(void *p);          // false: `p` is not one of the hook's blocks (heap memory)
};
void set_stage_hooks(const StageHooks &hooks);
void *stage_alloc(size_t bytes);
void stage_free(void *p);

// Writes the cache lines of [p, p + bytes) back to memory (x86: clflush; elsewhere a no-op).  An upload array that up to 128
// host threads have just written sits in dirty lines of many caches (on a two-socket box also the remote socket's); the DMA
// engine then has to pull every line out of a cache.  A writer that flushes its own part when it is done hands the DMA plain
// memory.  Only used when WHMEC_FLUSH_UPLOAD=1 (see stage_flush_enabled).
void stage_flush(const void *p, size_t bytes);
bool stage_flush_enabled();

template <class T>
struct StageAlloc {
    using value_type = T;
    template <class U>
    struct rebind {
        using other = StageAlloc<U>;
    };
    StageAlloc() = default;
    template <class U>
    StageAlloc(const StageAlloc<U> &) {}
    T *allocate(size_t n) { return static_cast<T *>(stage_alloc(n * sizeof(T))); }
    void deallocate(T *p, size_t) { stage_free(p); }
    template <class U>
    void construct(U *p) {
        ::new ((void *)p) U;  // default-initialisation: resize() leaves trivial elements untouched
    }
    template <class U, class... A>
    void construct(U *p, A &&...a) {
        ::new ((void *)p) U(std::forward<A>(a)...);
    }
    template <class U>
    bool operator==(const StageAlloc<U> &) const { return true; }
    template <class U>
    bool operator!=(const StageAlloc<U> &) const { return false; }
};
template <class T>
using StagedVec = std::vector<T, StageAlloc<T>>;

}  // namespace whmec
