// TEST-ONLY: the group-wise driver of whmec_solve (csrc/grouped.h) with the emulated kernels as backend,
// so that cutting a problem into groups of chains, slicing the input arrays and merging the results can be
// checked against the reference without a GPU.  Not part of the product.
#include <string>

#include "../../whatshap_b200/csrc/grouped.h"

extern "C" int whemul_solve(const whmec_problem *p, whmec_solution *s, uint32_t chunk, char *err, size_t errlen);
extern "C" int whemul_tile_solve(const whmec_problem *p, whmec_solution *s, uint32_t chunk, uint32_t *n_panels, char *err, size_t errlen);

namespace {
struct EmulBackend {
    struct Handle {
        const whmec_problem *prob;
    };
    bool tiles;
    uint32_t started = 0, finished = 0;
    int start(const whmec_problem &q, Handle *&h, std::string &) {
        h = new Handle{&q};  // the "sweep" runs at finish(): the emulation is synchronous
        ++started;
        return WHMEC_OK;
    }
    int finish(Handle *h, whmec_solution *sub, std::string &msg) {
        char err[256] = {0};
        int rc = tiles ? whemul_tile_solve(h->prob, sub, 0, nullptr, err, sizeof err) : 100;
        if (rc == 100) rc = whemul_solve(h->prob, sub, 0, err, sizeof err);  // planner declined: column kernel
        msg = err;
        ++finished;
        return rc;
    }
    void destroy(Handle *h) { delete h; }
};
}  // namespace

// returns 0 and *handled = 1 when the groups were solved, *handled = 0 when the driver declined
extern "C" int whemul_grouped_solve(const whmec_problem *p, whmec_solution *s, uint32_t groups, int tiles, int *handled,
                                    uint32_t *n_groups) {
    EmulBackend be{tiles != 0};
    std::string msg;
    bool done = false;
    const int rc = whmec::solve_in_groups(p, s, groups, be, msg, &done);
    *handled = done ? 1 : 0;
    *n_groups = be.finished;
    return rc;
}
