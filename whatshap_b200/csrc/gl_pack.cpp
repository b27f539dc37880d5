// Host packing of the genotyping DP (gl_pack.h).  Restates, once per problem:
//   phred -> error probability          src/genotypecolumncostcomputer.cpp:24-45
//   transmission transition matrix      src/transitionprobabilitycomputer.cpp:18-44
//   allele-assignment priors            src/transitionprobabilitycomputer.cpp:46-89
// in the reference's long double arithmetic, rounded to double for the device.
#include "gl_pack.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>

namespace whmec {

namespace {

long double phred_probability(uint32_t phred) {  // genotypecolumncostcomputer.cpp:24-45
    if (phred == 0) return 0.9999L;
    return powl(10.0L, -(long double)(int)phred / 10.0L);
}

}  // namespace

int gl_pack(const whmec_problem *p, Packed &pk, GlPacked &g, std::string &err) {
    g = GlPacked();
    if (p->n_cols > 0 && !p->gl) {
        // assert(gls != nullptr), transitionprobabilitycomputer.cpp:66
        err = "genotyping needs genotype likelihoods (priors) for every individual and column";
        return WHMEC_ERR_INPUT;
    }
    whmec_problem q = *p;
    q.distrust = 1;  // the column structure does not depend on genotypes; no genotype constraint applies here
    int rc = pack_problem(&q, pk, err, false);
    if (rc != WHMEC_OK) return rc;
    for (uint32_t r = 0; r < pk.n_reads; ++r)
        if (pk.read_first[r] == pk.read_last[r]) {
            // the reference asserts first column < last column (backwardcolumniterator.cpp:41) and aborts
            err = "genotyping: a read covers a single variant (reads need at least two)";
            return WHMEC_ERR_INPUT;
        }
    const uint32_t n = pk.n, T = pk.T, P = pk.P, nA = 1u << P, n_ind = pk.n_ind;
    if (n_ind > GL_MAX_IND || P > GL_MAX_P) {
        err = "unsupported pedigree size for genotyping";
        return WHMEC_ERR_UNSUPPORTED;
    }
    g.cols.resize(n);
    g.eps.resize(pk.act_phred.size());
    for (size_t e = 0; e < pk.act_phred.size(); ++e) g.eps[e] = (double)phred_probability(pk.act_phred[e]);
    g.trans.resize((size_t)n * T * T);
    g.q.resize((size_t)n * T * nA);
    const uint32_t trio_bits = pk.tb;
    for (uint32_t k = 0; k < n; ++k) {
        const ColMeta &m = pk.cols[k];
        GlCol &c = g.cols[k];
        c.a = m.a;
        c.bw = m.bw;
        c.keep = m.keep;
        c.f = m.f;
        // Without transmission values (T == 1) a chain boundary hands over a single number, which cancels in the
        // per-column normalisation: every DP-independent chain is then a table of its own (and they can be
        // processed group by group when the backward tables of all columns do not fit the device).
        c.first = (k == 0) || (T == 1 && pk.cols[k - 1].f == 0);
        c.last = (k + 1 == n) || (T == 1 && m.f == 0);
        c.act_off = pk.act_off[k];
        c.beta_off = g.beta_doubles;
        const uint64_t proj = ((uint64_t)1 << m.f) * T;
        if (!c.last) g.beta_doubles += proj;
        g.max_proj = std::max(g.max_proj, proj);
        c.trans_off = (uint64_t)k * T * T;
        c.q_off = (uint64_t)k * T * nA;
        // transmission transitions
        const long double r = powl(10.0L, -(long double)p->recombcost[k] / 10.0L);
        std::vector<long double> bern(trio_bits + 1);
        for (uint32_t i = 0; i <= trio_bits; ++i) bern[i] = powl(r, (long double)i) * powl(1.0L - r, (long double)(trio_bits - i));
        for (uint32_t i = 0; i < T; ++i) {
            long double norm = 0.0L;
            for (uint32_t j = 0; j < T; ++j) norm += bern[popc32(i ^ j)];
            for (uint32_t j = 0; j < T; ++j) g.trans[c.trans_off + (size_t)i * T + j] = (double)(bern[popc32(i ^ j)] / norm);
        }
        // allele-assignment priors
        for (uint32_t i = 0; i < T; ++i) {
            const int8_t *h = &pk.h2p[(size_t)i * n_ind * 2];
            std::map<std::vector<uint8_t>, size_t> count;
            std::vector<std::vector<uint8_t>> geno(nA);
            std::vector<long double> prob(nA);
            for (uint32_t A = 0; A < nA; ++A) {
                long double pr = 1.0L;
                std::vector<uint8_t> gv(n_ind);
                for (uint32_t ind = 0; ind < n_ind; ++ind) {
                    const uint32_t gi = ((A >> h[2 * ind]) & 1u) + ((A >> h[2 * ind + 1]) & 1u);
                    pr *= p->gl[((size_t)ind * n + k) * 3 + gi];
                    gv[ind] = (uint8_t)gi;
                }
                count[gv] += 1;
                prob[A] = pr;
                geno[A] = gv;
            }
            long double norm = 0.0L;
            for (uint32_t A = 0; A < nA; ++A) {
                prob[A] /= count[geno[A]];
                norm += prob[A];
            }
            for (uint32_t A = 0; A < nA; ++A) g.q[c.q_off + (size_t)i * nA + A] = (double)(prob[A] / norm);
        }
    }
    return WHMEC_OK;
}

void gl_schedule(const GlPacked &g, uint32_t T, uint32_t lo, uint32_t hi, GlSchedule &out) {
    out = GlSchedule();
    out.beta_base = g.cols[lo].beta_off;
    out.beta_doubles = g.cols[hi - 1].beta_off - out.beta_base;  // the group ends a table: its last column owns no entries
    struct Table {
        uint32_t begin, end;  // columns [begin, end)
        uint64_t f_off, f_size;
    };
    std::vector<Table> tables;
    uint32_t longest = 0;
    for (uint32_t k = lo, begin = lo; k < hi; ++k)
        if (g.cols[k].last) {
            uint64_t size = 1;
            for (uint32_t q = begin; q <= k; ++q)
                if (!g.cols[q].last) size = std::max(size, ((uint64_t)1 << g.cols[q].f) * T);
            tables.push_back(Table{begin, k + 1, out.f_pool_doubles, size});
            out.f_pool_doubles += 2 * size;
            longest = std::max(longest, k + 1 - begin);
            begin = k + 1;
        }
    // backward: step s handles the s-th column from the right of every table that still has a column to its left
    for (uint32_t s = 0; s + 1 < longest; ++s) {
        out.bwd_begin.push_back((uint32_t)out.steps.size());
        for (const Table &t : tables) {
            if (t.end - t.begin < s + 2) continue;
            const uint32_t k = t.end - 1 - s;
            GlStep st{};
            st.k = k;
            st.cells_log2 = g.cols[k].a;
            st.cur_off = g.cols[k - 1].beta_off - out.beta_base;
            st.n_scale = ((uint64_t)1 << g.cols[k - 1].f) * T;
            out.steps.push_back(st);
        }
    }
    out.bwd_begin.push_back((uint32_t)out.steps.size());
    for (uint32_t s = 0; s < longest; ++s) {
        out.fwd_begin.push_back((uint32_t)out.steps.size());
        for (const Table &t : tables) {
            if (t.end - t.begin <= s) continue;
            const uint32_t k = t.begin + s;
            GlStep st{};
            st.k = k;
            st.cells_log2 = g.cols[k].a;
            st.cur_off = t.f_off + (uint64_t)(s & 1) * t.f_size;
            st.prev_off = t.f_off + (uint64_t)((s + 1) & 1) * t.f_size;
            st.n_scale = g.cols[k].last ? 0 : ((uint64_t)1 << g.cols[k].f) * T;
            st.n_clear = t.f_size;
            out.steps.push_back(st);
        }
    }
    out.fwd_begin.push_back((uint32_t)out.steps.size());
}

bool gl_groups(const GlPacked &g, uint32_t T, uint64_t budget, std::vector<uint32_t> &begin) {
    begin.assign(1, 0);
    const uint32_t n = (uint32_t)g.cols.size();
    uint64_t run = 0, table = 0, widest = 1;
    uint32_t table_begin = 0;
    for (uint32_t k = 0; k < n; ++k) {
        if (!g.cols[k].last) {
            const uint64_t proj = ((uint64_t)1 << g.cols[k].f) * T;
            table += proj;
            widest = std::max(widest, proj);
            continue;
        }
        const uint64_t cost = table + 2 * widest;  // a table ends here: its backward tables and its two projection buffers
        if (cost > budget) return false;
        if (run + cost > budget) {  // close the group before this table
            begin.push_back(table_begin);
            run = 0;
        }
        run += cost;
        table_begin = k + 1;
        table = 0;
        widest = 1;
    }
    begin.push_back(n);
    return true;
}

void gl_scale_host(double *v, uint64_t n) {
    double mx = 0.0;
    for (uint64_t i = 0; i < n; ++i) mx = v[i] > mx ? v[i] : mx;
    if (!(mx > 0.0)) return;
    const double inv = 1.0 / mx;
    for (uint64_t i = 0; i < n; ++i) v[i] *= inv;
}

void gl_normalise(const double *acc, uint32_t n, uint32_t n_ind, double *likelihoods) {
    for (uint32_t k = 0; k < n; ++k) {
        const double *a0 = acc + (size_t)k * n_ind * 3;
        const double total = a0[0] + a0[1] + a0[2];  // every (x, i, A) lands in exactly one genotype of each individual
        for (uint32_t ind = 0; ind < n_ind; ++ind)
            for (uint32_t gi = 0; gi < 3; ++gi) likelihoods[((size_t)ind * n + k) * 3 + gi] = a0[ind * 3 + gi] / total;
    }
}

}  // namespace whmec

// compute_genotypes (src/genotyper.cpp:12-54): columns are independent; within a column the factors are applied in
// the order of the column iterator (ascending read index) so that every double equals the reference's.
extern "C" int whmec_compute_genotypes(const whmec_problem *p, double *gl, int8_t *gt, char *err, size_t errlen) {
    using namespace whmec;
    std::string msg;
    int rc = WHMEC_OK;
    try {
        if (!p || !gl || !gt) {
            msg = "null argument";
            rc = WHMEC_ERR_INPUT;
        } else {
            whmec_problem q = *p;
            q.n_ind = 1;  // one sample's reads; no pedigree, no genotype constraint
            q.n_trios = 0;
            q.trios = nullptr;
            q.distrust = 0;
            std::vector<uint8_t> het(p->n_cols, 1);
            std::vector<uint32_t> zero_ind(p->n_reads, 0), no_cost(p->n_cols, 0);
            q.gt = het.data();
            q.gl = nullptr;
            q.read_ind = zero_ind.data();
            if (!q.recombcost) q.recombcost = no_cost.data();
            Packed pk;
            rc = pack_problem(&q, pk, msg, false);
            if (rc == WHMEC_OK) {
                const uint32_t n = pk.n;
                const uint32_t n_tasks = n / 2048 + 1, step = (n + n_tasks - 1) / n_tasks;
                parallel_tasks(n_tasks, host_threads(16), [&](uint32_t task) {
                    for (uint32_t k = task * step; k < std::min(n, (task + 1) * step); ++k) {
                        double d[3] = {1.0 / 3.0, 1.0 / 3.0, 1.0 / 3.0};
                        const uint64_t e0 = pk.act_off[k];
                        for (uint32_t j = 0; j < pk.cols[k].a; ++j) {
                            const uint8_t al = pk.act_allele[e0 + j];
                            if (al > 1) continue;
                            const double p_wrong = std::max(0.05, std::pow(10.0, -((double)pk.act_phred[e0 + j]) / 10.0));
                            const double f_same = 2.0 / 3.0 - 1.0 / 3.0 * p_wrong, f_het = 1.0 / 3.0, f_other = 1.0 / 3.0 * p_wrong;
                            const double f[3] = {al == 0 ? f_same : f_other, f_het, al == 0 ? f_other : f_same};
                            double sum = 0.0;  // operator* (genotypedistribution.cpp:56-66)
                            for (int i = 0; i < 3; ++i) {
                                d[i] *= f[i];
                                sum += d[i];
                            }
                            for (int i = 0; i < 3; ++i) d[i] /= sum;
                        }
                        double p_sum = 0.0;  // normalize (:24-36)
                        for (int i = 0; i < 3; ++i) p_sum += d[i];
                        if (p_sum <= 0.0) d[0] = d[1] = d[2] = 1.0 / 3.0;
                        else
                            for (int i = 0; i < 3; ++i) d[i] /= p_sum;
                        int best_index = 0;  // likeliestGenotype / errorProbability (:12-22,37-55)
                        double best = 0.0;
                        for (int i = 0; i < 3; ++i)
                            if (d[i] > best) {
                                best = d[i];
                                best_index = i;
                            }
                        double p_err = 0.0;
                        for (int i = 0; i < 3; ++i)
                            if (i != best_index) p_err += d[i];
                        gt[k] = p_err < 0.1 ? (int8_t)best_index : (int8_t)-1;
                        for (int i = 0; i < 3; ++i) gl[(size_t)k * 3 + i] = d[i];
                    }
                });
            }
        }
    } catch (const std::exception &e) {
        msg = e.what();
        rc = WHMEC_ERR_INPUT;
    }
    if (rc != WHMEC_OK && err && errlen) {
        std::strncpy(err, msg.c_str(), errlen - 1);
        err[errlen - 1] = 0;
    }
    return rc;
}
