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

// Launch schedule of a group of tables (columns [lo, hi), whole tables): one launch advances EVERY table of the group
// by one column, so the launch count is 4 x (longest table) instead of 4 x (columns).  Shared by genotype.cu and the
// test-only emulation, which executes the same steps with the same buffer offsets.
struct GlStep {           // one column of one table inside a launch (grid.y)
    uint32_t k;           // column
    uint32_t cells_log2;  // a_k
    uint64_t cur_off;     // forward: F_k in the F pool; backward: B_{k-1} in the group's backward store (doubles)
    uint64_t prev_off;    // forward: F_{k-1} in the F pool (cleared after the step); backward: unused
    uint64_t n_scale;     // entries of the finished column at cur_off to rescale (0: none)
    uint64_t n_clear;     // forward: entries at prev_off to clear for the next column
};
struct GlSchedule {
    std::vector<GlStep> steps;
    std::vector<uint32_t> bwd_begin, fwd_begin;  // launch s covers steps [begin[s], begin[s + 1])
    uint64_t beta_base = 0, beta_doubles = 0;    // the group's slice of the backward tables
    uint64_t f_pool_doubles = 0;                 // two projection buffers per table
};
void gl_schedule(const GlPacked &g, uint32_t T, uint32_t lo, uint32_t hi, GlSchedule &out);

// Cuts the columns into groups of whole tables whose backward tables + two projection buffers per table stay within
// `budget` doubles: group q = columns [begin[q], begin[q + 1]).  False if a single table exceeds the budget.
bool gl_groups(const GlPacked &g, uint32_t T, uint64_t budget, std::vector<uint32_t> &begin);

// Packs `p` (whose gl holds the genotype priors; gt and distrust are ignored) for the genotyping DP.
int gl_pack(const whmec_problem *p, Packed &pk, GlPacked &g, std::string &err);

// Divides a finished projection column by its largest entry (no-op for an all-zero column).
void gl_scale_host(double *v, uint64_t n);

// likelihoods[(ind * n + k) * 3 + g] = acc[(k * n_ind + ind) * 3 + g] / (total of column k)   (genotypedptable.cpp:438-442)
void gl_normalise(const double *acc, uint32_t n, uint32_t n_ind, double *likelihoods);

}  // namespace whmec
