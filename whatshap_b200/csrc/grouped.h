// Single-individual problems (T == 1) decompose exactly at columns that no read spans (SURVEY.md
// section 8(e): cost adds, partitioning / super-reads concatenate, ties are unaffected).  This header cuts
// a whmec_problem into G groups of whole chains and drives them through a backend one after the other:
// `start` packs, uploads and ENQUEUES the sweep of a group without waiting for the device, so that the
// host work of group g+1 runs while the GPU sweeps group g; `finish` waits for a group and writes its
// results.  The backend is a template parameter so that the slicing / merging logic is exercised on the
// CPU by the test harness (tests/emul) with the emulated kernels; the product instantiates it with the
// CUDA plan functions (whmec.cu).  The reference has no counterpart: it sweeps one table on one thread
// (src/pedigreedptable.cpp:84-174).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/whmec.h"

namespace whmec {

struct ProblemSlice {
    whmec_problem prob;                // pointers into the caller's arrays except the two rebased ones
    std::vector<uint64_t> read_off;    // rebased to the slice's first entry
    std::vector<uint32_t> ent_col;     // rebased to the slice's first column
    uint32_t col_lo = 0, read_lo = 0;
};

// Column ranges [lo, hi) of G groups of whole chains with about equal DP work (sum of 2^active).
// Returns false when the problem is not a plain sorted single-individual one (the caller then takes the
// ordinary path, which also produces the reference's error texts).
inline bool cut_into_groups(const whmec_problem *p, uint32_t G, std::vector<uint32_t> &cuts) {
    if (!p || p->n_ind != 1 || p->n_trios != 0 || p->n_cols == 0 || p->n_reads == 0 || G < 2) return false;
    if (!p->read_off || !p->ent_col) return false;
    const uint32_t n = p->n_cols;
    std::vector<int32_t> delta(n + 1, 0);    // +1 where a read starts, -1 after it ends
    std::vector<int32_t> span(n + 1, 0);     // reads active in columns k-1 AND k
    uint32_t prev_first = 0;
    for (uint32_t r = 0; r < p->n_reads; ++r) {
        const uint64_t b = p->read_off[r], e = p->read_off[r + 1];
        if (e <= b) return false;
        const uint32_t first = p->ent_col[b], last = p->ent_col[e - 1];
        if (first < prev_first || last < first || last >= n) return false;
        prev_first = first;
        delta[first] += 1;
        delta[last + 1] -= 1;
        if (last > first) {
            span[first + 1] += 1;
            span[last + 1] -= 1;
        }
    }
    std::vector<uint32_t> chain_start;  // columns that start a chain
    std::vector<double> work_before;    // DP cells in columns [0, chain start)
    int32_t active = 0, crossing = 0;
    double work = 0;
    for (uint32_t k = 0; k < n; ++k) {
        crossing += span[k];
        if (k == 0 || crossing == 0) {
            chain_start.push_back(k);
            work_before.push_back(work);
        }
        active += delta[k];
        work += std::ldexp(1.0, std::min(active, 40));
    }
    if (chain_start.size() < 2 * (size_t)G) return false;  // too few chains to be worth it
    cuts.assign(1, 0);
    for (uint32_t g = 1; g < G; ++g) {
        const double want = work * g / G;
        size_t i = std::lower_bound(work_before.begin(), work_before.end(), want) - work_before.begin();
        i = std::min(i, chain_start.size() - 1);
        if (chain_start[i] > cuts.back()) cuts.push_back(chain_start[i]);
    }
    cuts.push_back(n);
    return cuts.size() > 2;
}

// Sub-problem over columns [lo, hi): its reads are those that start in the range (no read crosses a cut).
inline void slice_problem(const whmec_problem *p, uint32_t lo, uint32_t hi, uint32_t read_lo, uint32_t read_hi, ProblemSlice &out) {
    const uint64_t e0 = p->read_off[read_lo], e1 = p->read_off[read_hi];
    out.col_lo = lo;
    out.read_lo = read_lo;
    out.read_off.resize((size_t)(read_hi - read_lo) + 1);
    for (uint32_t r = read_lo; r <= read_hi; ++r) out.read_off[r - read_lo] = p->read_off[r] - e0;
    out.ent_col.resize((size_t)(e1 - e0));
    for (uint64_t e = e0; e < e1; ++e) out.ent_col[e - e0] = p->ent_col[e] - lo;
    whmec_problem &q = out.prob;
    q = *p;
    q.n_cols = hi - lo;
    q.positions = p->positions ? p->positions + lo : nullptr;
    q.n_reads = read_hi - read_lo;
    q.read_off = out.read_off.data();
    q.ent_col = out.ent_col.data();
    q.ent_allele = p->ent_allele ? p->ent_allele + e0 : nullptr;
    q.ent_phred = p->ent_phred ? p->ent_phred + e0 : nullptr;
    q.read_ind = p->read_ind ? p->read_ind + read_lo : nullptr;
    q.recombcost = p->recombcost ? p->recombcost + lo : nullptr;
    q.gt = p->gt ? p->gt + lo : nullptr;                   // one individual: a single row
    q.gl = p->gl ? p->gl + (size_t)lo * 3 : nullptr;
}

// Backend: int start(const whmec_problem &, Handle *&, std::string &);  int finish(Handle *, whmec_solution *, std::string &);
//          void destroy(Handle *).   `handled` false: nothing was done, take the ordinary path.
template <class Backend>
int solve_in_groups(const whmec_problem *p, whmec_solution *s, uint32_t G, Backend &be, std::string &msg, bool *handled) {
    *handled = false;
    std::vector<uint32_t> cuts;
    if (!s || !cut_into_groups(p, G, cuts)) return WHMEC_OK;
    const size_t n_groups = cuts.size() - 1;
    std::vector<ProblemSlice> slices(n_groups);
    std::vector<typename Backend::Handle *> handles(n_groups, nullptr);
    auto cleanup = [&] {
        for (auto *h : handles)
            if (h) be.destroy(h);
    };
    uint32_t read_lo = 0;
    for (size_t g = 0; g < n_groups; ++g) {
        uint32_t read_hi = read_lo;
        while (read_hi < p->n_reads && p->ent_col[p->read_off[read_hi]] < cuts[g + 1]) ++read_hi;
        slice_problem(p, cuts[g], cuts[g + 1], read_lo, read_hi, slices[g]);
        read_lo = read_hi;
        if (be.start(slices[g].prob, handles[g], msg) != WHMEC_OK) {  // any error: the ordinary path reports it in full
            cleanup();
            msg.clear();
            return WHMEC_OK;
        }
    }
    uint64_t cost = 0;
    const uint32_t n = p->n_cols;
    std::vector<uint8_t> alleles;
    for (size_t g = 0; g < n_groups; ++g) {
        const whmec_problem &q = slices[g].prob;
        const uint32_t lo = slices[g].col_lo;
        alleles.assign((size_t)2 * q.n_cols, 0);
        whmec_solution sub;
        sub.cost = 0;
        sub.path_index = s->path_index ? s->path_index + lo : nullptr;
        sub.path_tv = s->path_tv ? s->path_tv + lo : nullptr;
        sub.partition = s->partition ? s->partition + slices[g].read_lo : nullptr;
        sub.sr_allele = s->sr_allele ? alleles.data() : nullptr;       // rows of the whole problem are n columns apart
        sub.sr_quality = s->sr_quality ? s->sr_quality + lo : nullptr;  // one individual: a single row
        const int rc = be.finish(handles[g], &sub, msg);
        if (rc != WHMEC_OK) {
            cleanup();
            msg.clear();
            return WHMEC_OK;
        }
        if (s->sr_allele)
            for (int h = 0; h < 2; ++h) std::memcpy(s->sr_allele + (size_t)h * n + lo, alleles.data() + (size_t)h * q.n_cols, q.n_cols);
        cost += sub.cost;
        be.destroy(handles[g]);
        handles[g] = nullptr;
    }
    s->cost = (uint32_t)cost;
    *handled = true;
    return WHMEC_OK;
}

}  // namespace whmec
