// TEST-ONLY host emulation of the genotyping kernels (genotype.cu): the host packer (gl_pack.cpp) and the
// __host__ __device__ per-cell functions of gl_device.h stepped serially, in the kernels' launch order.
#include <cstring>
#include <string>
#include <vector>

#include "../../whatshap_b200/csrc/gl_pack.h"

using namespace whmec;

extern "C" int whemul_genotype_grouped(const whmec_problem *p, double *likelihoods, uint64_t budget_doubles, uint32_t *n_groups,
                                       char *err, size_t errlen) {
    Packed pk;
    GlPacked g;
    std::string msg;
    int rc = gl_pack(p, pk, g, msg);
    if (rc != WHMEC_OK) {
        if (err && errlen) {
            std::strncpy(err, msg.c_str(), errlen - 1);
            err[errlen - 1] = 0;
        }
        return rc;
    }
    const uint32_t n = pk.n, T = pk.T, n_ind = pk.n_ind;
    if (n == 0) return WHMEC_OK;
    const GlView v = g.view(pk);
    auto add = [](double *addr, double val) { *addr += val; };
    std::vector<double> acc((size_t)n * n_ind * 3, 0.0);
    // groups of whole tables, as whmec_genotype forms them when the backward tables do not fit the device together
    // (gl_groups with `budget_doubles`; 0 = no limit); each group runs its launch schedule (gl_schedule)
    std::vector<uint32_t> group_begin;
    if (!gl_groups(g, T, budget_doubles ? budget_doubles : ~0ull, group_begin)) return WHMEC_ERR_UNSUPPORTED;
    if (n_groups) *n_groups = (uint32_t)group_begin.size() - 1;
    for (size_t q = 0; q + 1 < group_begin.size(); ++q) {
        GlSchedule sc;
        gl_schedule(g, T, group_begin[q], group_begin[q + 1], sc);
        std::vector<double> beta(sc.beta_doubles + 1, 0.0), F(sc.f_pool_doubles, 0.0);
        auto table_of = [&](uint32_t k) { return beta.data() + (g.cols[k].beta_off - sc.beta_base); };
        for (size_t l = 0; l + 1 < sc.bwd_begin.size(); ++l) {
            for (uint32_t e = sc.bwd_begin[l]; e < sc.bwd_begin[l + 1]; ++e) {  // the cells of one launch
                const GlStep &st = sc.steps[e];
                for (uint64_t x = 0; x < ((uint64_t)1 << st.cells_log2); ++x)
                    gl_backward_cell(v, st.k, (uint32_t)x, table_of(st.k), beta.data() + st.cur_off, add);
            }
            for (uint32_t e = sc.bwd_begin[l]; e < sc.bwd_begin[l + 1]; ++e) gl_scale_host(beta.data() + sc.steps[e].cur_off, sc.steps[e].n_scale);
        }
        for (size_t l = 0; l + 1 < sc.fwd_begin.size(); ++l) {
            for (uint32_t e = sc.fwd_begin[l]; e < sc.fwd_begin[l + 1]; ++e) {
                const GlStep &st = sc.steps[e];
                for (uint64_t x = 0; x < ((uint64_t)1 << st.cells_log2); ++x)
                    gl_forward_cell(v, st.k, (uint32_t)x, F.data() + st.prev_off, F.data() + st.cur_off, table_of(st.k),
                                    acc.data() + (size_t)st.k * n_ind * 3, add);
            }
            for (uint32_t e = sc.fwd_begin[l]; e < sc.fwd_begin[l + 1]; ++e) {
                const GlStep &st = sc.steps[e];
                gl_scale_host(F.data() + st.cur_off, st.n_scale);
                std::fill(F.begin() + st.prev_off, F.begin() + st.prev_off + st.n_clear, 0.0);
            }
        }
    }
    gl_normalise(acc.data(), n, n_ind, likelihoods);
    return WHMEC_OK;
}

extern "C" int whemul_genotype(const whmec_problem *p, double *likelihoods, char *err, size_t errlen) {
    return whemul_genotype_grouped(p, likelihoods, 0, nullptr, err, errlen);
}
