// Host packing of the forward-backward genotyping DP (see gl_device.h): per-column records, error
// probabilities, transition and allele-assignment probabilities.  Plain C++ (used by genotype.cu and by the
// test-only emulation).
#pragma once
#include <string>
#include <vector>

#include "gl_device.h"
#include "pack.h"

namespace whmec {

struct GlPacked {
    std::vector<GlCol> cols;
    std::vector<double> eps;          // aligned with Packed::act_allele / act_ind
    std::vector<double> trans, q;
    uint64_t beta_doubles = 0;        // sum over columns but the last of 2^f * T
    uint64_t max_proj = 1;            // largest projection column (doubles)
    GlView view(const Packed &pk) const {
        return GlView{cols.data(), eps.data(), pk.act_allele.data(), pk.act_ind.data(), pk.h2p.data(), trans.data(), q.data(), pk.T, pk.P, pk.n_ind};
    }
};

// Packs `p` (whose gl holds the genotype priors; gt and distrust are ignored) for the genotyping DP.
int gl_pack(const whmec_problem *p, Packed &pk, GlPacked &g, std::string &err);

// Divides a finished projection column by its largest entry (no-op for an all-zero column).
void gl_scale_host(double *v, uint64_t n);

// likelihoods[(ind * n + k) * 3 + g] = acc[(k * n_ind + ind) * 3 + g] / (total of column k)   (genotypedptable.cpp:438-442)
void gl_normalise(const double *acc, uint32_t n, uint32_t n_ind, double *likelihoods);

}  // namespace whmec
