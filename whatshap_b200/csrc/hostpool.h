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

}  // namespace whmec
