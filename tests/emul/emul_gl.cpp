// TEST-ONLY host emulation of the genotyping kernels (genotype.cu): the host packer (gl_pack.cpp) and the
// __host__ __device__ per-cell functions of gl_device.h stepped serially, in the kernels' launch order.
#include <cstring>
#include <string>
#include <vector>

#include "../../whatshap_b200/csrc/gl_pack.h"

using namespace whmec;

extern "C" int whemul_genotype(const whmec_problem *p, double *likelihoods, char *err, size_t errlen) {
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
    std::vector<double> beta(g.beta_doubles + 1, 0.0), F[2], acc((size_t)n * n_ind * 3, 0.0);
    F[0].assign(g.max_proj, 0.0);
    F[1].assign(g.max_proj, 0.0);
    for (uint32_t k = n - 1; k >= 1; --k) {
        if (g.cols[k].first) continue;  // nothing enters the first column of a table from the left
        double *out = beta.data() + g.cols[k - 1].beta_off;
        for (uint64_t x = 0; x < ((uint64_t)1 << g.cols[k].a); ++x) gl_backward_cell(v, k, (uint32_t)x, beta.data() + g.cols[k].beta_off, out, add);
        gl_scale_host(out, ((uint64_t)1 << g.cols[k - 1].f) * T);
    }
    for (uint32_t k = 0; k < n; ++k) {
        std::vector<double> &cur = F[k & 1], &prev = F[(k + 1) & 1];
        for (uint64_t x = 0; x < ((uint64_t)1 << g.cols[k].a); ++x)
            gl_forward_cell(v, k, (uint32_t)x, prev.data(), cur.data(), beta.data() + g.cols[k].beta_off, acc.data() + (size_t)k * n_ind * 3, add);
        if (!g.cols[k].last) gl_scale_host(cur.data(), ((uint64_t)1 << g.cols[k].f) * T);
        std::fill(prev.begin(), prev.end(), 0.0);
    }
    gl_normalise(acc.data(), n, n_ind, likelihoods);
    return WHMEC_OK;
}
