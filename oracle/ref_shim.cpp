// TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT PATH.
//
// C-ABI adapter around the UNMODIFIED reference C++ classes (ReadSet, Read, Pedigree,
// PedigreeDPTable) so that tests and bench.py's cpu_baseline / `--impl reference` arm can run
// the real reference on the same flat arrays the product's C ABI (include/whmec.h) takes.
// The reference sources are compiled where they lie under $WHATSHAP_REF/src by oracle/Makefile
// into oracle/_ref/libwhref.so; nothing from the reference is copied into this repository.
//
// Only tests/, __graft_entry__.smoke() and bench.py may load the resulting library.
//
// Reference entry points driven here:
//   PedigreeDPTable ctor            src/pedigreedptable.cpp:15-37
//   get_super_reads                 src/pedigreedptable.cpp:344-388
//   get_optimal_partitioning        src/pedigreedptable.cpp:391-406  (+ core.pyx:414 true->0 mapping)
//   get_optimal_score               src/pedigreedptable.cpp:338-341
//   GenotypeDPTable ctor / get_genotype_likelihoods   src/genotypedptable.cpp:17-48,445-451  (sibling DP)
//   PedMecHeuristic ctor / solve / getOpt*            src/pedmecheuristic.cpp:9-81,121-409         (row-limited heuristic)

#include <chrono>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

// index_path (the optimal bipartition index per column) is a private member of the reference
// class; expose it to this test shim only, without touching the reference sources.
#define private public
#include "pedigreedptable.h"
#undef private
#include "pedigree.h"
#include "read.h"
#include "readset.h"
#include "genotype.h"
#include "phredgenotypelikelihoods.h"
#include "genotypedptable.h"
#include <cassert>
#include "genotypedistribution.h"
#include "genotyper.h"
#include "pedmecheuristic.h"

#include "../include/whmec.h"

namespace {

void set_err(char *err, size_t errlen, const char *msg) {
    if (err && errlen) {
        std::strncpy(err, msg, errlen - 1);
        err[errlen - 1] = 0;
    }
}

struct Built {
    std::unique_ptr<ReadSet> rs;
    std::unique_ptr<Pedigree> ped;
    std::vector<unsigned int> recomb;
    std::vector<unsigned int> positions;
};

void build(const whmec_problem *p, Built &b) {
    b.rs.reset(new ReadSet());
    for (uint32_t r = 0; r < p->n_reads; ++r) {
        char name[32];
        std::snprintf(name, sizeof name, "r%07u", r);
        Read *rd = new Read(name, 60, 0, (int)p->read_ind[r]);
        for (uint64_t e = p->read_off[r]; e < p->read_off[r + 1]; ++e)
            rd->addVariant((int)p->positions[p->ent_col[e]], (int)p->ent_allele[e], (int)p->ent_phred[e]);
        b.rs->add(rd);
    }
    b.ped.reset(new Pedigree());
    for (uint32_t i = 0; i < p->n_ind; ++i) {
        std::vector<Genotype *> gts;
        std::vector<PhredGenotypeLikelihoods *> gls;
        for (uint32_t k = 0; k < p->n_cols; ++k) {
            uint8_t g = p->gt ? p->gt[(size_t)i * p->n_cols + k] : 1;
            if (g == 0) gts.push_back(new Genotype(std::vector<uint32_t>{0, 0}));
            else if (g == 1) gts.push_back(new Genotype(std::vector<uint32_t>{0, 1}));
            else if (g == 2) gts.push_back(new Genotype(std::vector<uint32_t>{1, 1}));
            else gts.push_back(new Genotype());
            if (p->gl) {
                const double *q = p->gl + ((size_t)i * p->n_cols + k) * 3;
                gls.push_back(new PhredGenotypeLikelihoods(std::vector<double>{q[0], q[1], q[2]}, 2, 2));
            } else {
                gls.push_back(nullptr);
            }
        }
        b.ped->addIndividual(i, gts, gls);  // id == index
    }
    for (uint32_t t = 0; t < p->n_trios; ++t)
        b.ped->addRelationship(p->trios[3 * t], p->trios[3 * t + 1], p->trios[3 * t + 2]);
    b.recomb.assign(p->recombcost, p->recombcost + p->n_cols);
    b.positions.assign(p->positions, p->positions + p->n_cols);
}

int solve_one(const whmec_problem *p, whmec_solution *s, double *ctor_seconds, char *err, size_t errlen) {
    try {
        Built b;
        build(p, b);
        auto t0 = std::chrono::steady_clock::now();
        PedigreeDPTable dp(b.rs.get(), b.recomb, b.ped.get(), p->distrust != 0, &b.positions);
        std::vector<ReadSet *> out;
        for (uint32_t i = 0; i < p->n_ind; ++i) out.push_back(new ReadSet());
        std::vector<unsigned int> tv;
        dp.get_super_reads(&out, &tv);
        auto t1 = std::chrono::steady_clock::now();
        if (ctor_seconds) *ctor_seconds = std::chrono::duration<double>(t1 - t0).count();
        if (s) {
            s->cost = dp.get_optimal_score();
            for (uint32_t k = 0; k < p->n_cols; ++k) {
                if (s->path_tv) s->path_tv[k] = tv[k];
                if (s->path_index) s->path_index[k] = dp.index_path[k].index;
            }
            std::unique_ptr<std::vector<bool>> part(dp.get_optimal_partitioning());
            if (s->partition)
                for (uint32_t r = 0; r < p->n_reads; ++r) s->partition[r] = (*part)[r] ? 0 : 1;
            for (uint32_t i = 0; i < p->n_ind; ++i) {
                Read *h0 = out[i]->get(0);
                Read *h1 = out[i]->get(1);
                for (uint32_t k = 0; k < p->n_cols; ++k) {
                    if (s->sr_allele) {
                        s->sr_allele[((size_t)i * 2 + 0) * p->n_cols + k] = (uint8_t)h0->getAllele(k);
                        s->sr_allele[((size_t)i * 2 + 1) * p->n_cols + k] = (uint8_t)h1->getAllele(k);
                    }
                    if (s->sr_quality) s->sr_quality[(size_t)i * p->n_cols + k] = (uint32_t)h1->getVariantQuality(k);
                }
            }
        }
        for (ReadSet *r : out) delete r;
        return WHMEC_OK;
    } catch (const std::exception &e) {
        set_err(err, errlen, e.what());
        if (std::strstr(e.what(), "Mendelian")) return WHMEC_ERR_MENDELIAN;
        return WHMEC_ERR_INPUT;
    }
}

}  // namespace

extern "C" {

// Run the reference DP once.  ctor_seconds (optional) receives the wall time of
// PedigreeDPTable(...) + get_super_reads(), excluding the construction of the inputs.
int whref_solve(const whmec_problem *p, whmec_solution *s, double *ctor_seconds, char *err, size_t errlen) {
    return solve_one(p, s, ctor_seconds, err, errlen);
}

// Run `n` independent reference DPs on `n_threads` host threads (each instance is
// single-threaded, as the reference is).  Returns the wall time of the whole batch in
// *wall_seconds.  Used by bench.py --impl reference to occupy all host cores.
int whref_solve_many(const whmec_problem *const *ps, uint32_t n, uint32_t n_threads, double *wall_seconds,
                     char *err, size_t errlen) {
    std::vector<int> rc(n, 0);
    std::vector<std::string> msgs(n);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    if (n_threads == 0) n_threads = 1;
    for (uint32_t w = 0; w < n_threads; ++w) {
        th.emplace_back([&, w]() {
            for (uint32_t i = w; i < n; i += n_threads) {
                char e[256] = {0};
                rc[i] = solve_one(ps[i], nullptr, nullptr, e, sizeof e);
                msgs[i] = e;
            }
        });
    }
    for (auto &t : th) t.join();
    auto t1 = std::chrono::steady_clock::now();
    if (wall_seconds) *wall_seconds = std::chrono::duration<double>(t1 - t0).count();
    for (uint32_t i = 0; i < n; ++i)
        if (rc[i] != WHMEC_OK) {
            set_err(err, errlen, msgs[i].c_str());
            return rc[i];
        }
    return WHMEC_OK;
}

// Run the reference's forward-backward genotyping DP (GenotypeDPTable) once: out[(i * n_cols + k) * 3 + g] = likelihood of
// genotype index g (0/0, 0/1, 1/1) of individual i at column k, rounded from long double to double.  p->gl holds the
// genotype priors (the reference asserts on missing ones); p->gt is ignored by this DP.
int whref_genotype(const whmec_problem *p, double *out, double *ctor_seconds, char *err, size_t errlen) {
    try {
        if (p->n_cols > 0 && !p->gl) {
            set_err(err, errlen, "genotype priors (gl) are required");
            return WHMEC_ERR_INPUT;
        }
        Built b;
        build(p, b);
        auto t0 = std::chrono::steady_clock::now();
        GenotypeDPTable dp(b.rs.get(), b.recomb, b.ped.get(), &b.positions);
        auto t1 = std::chrono::steady_clock::now();
        if (ctor_seconds) *ctor_seconds = std::chrono::duration<double>(t1 - t0).count();
        for (uint32_t i = 0; i < p->n_ind; ++i)
            for (uint32_t k = 0; k < p->n_cols; ++k) {
                double *q = out + ((size_t)i * p->n_cols + k) * 3;
                std::vector<long double> l = dp.get_genotype_likelihoods(i, k);
                for (int g = 0; g < 3; ++g) q[g] = (double)l[g];
            }
        return WHMEC_OK;
    } catch (const std::exception &e) {
        set_err(err, errlen, e.what());
        return WHMEC_ERR_INPUT;
    }
}

// compute_genotypes (src/genotyper.cpp:12-54) on the reads of `p`: gl[k*3 + g], gt[k] in {0,1,2} or -1 (empty Genotype).
int whref_compute_genotypes(const whmec_problem *p, double *gl, int8_t *gt, char *err, size_t errlen) {
    try {
        Built b;
        build(p, b);
        std::vector<Genotype> genotypes;
        std::vector<GenotypeDistribution> dists;
        std::vector<unsigned int> *positions = new std::vector<unsigned int>(b.positions);  // the callee deletes it
        compute_genotypes(*b.rs, &genotypes, &dists, positions);
        for (uint32_t k = 0; k < p->n_cols; ++k) {
            for (int g = 0; g < 3; ++g) gl[(size_t)k * 3 + g] = dists[k].probabilityOf(g);
            gt[k] = genotypes[k].is_none() ? (int8_t)-1 : (int8_t)genotypes[k].get_index();
        }
        return WHMEC_OK;
    } catch (const std::exception &e) {
        set_err(err, errlen, e.what());
        return WHMEC_ERR_INPUT;
    }
}

const char *whref_info(void) { return "whatshap reference C++ (PedigreeDPTable), compiled in place by oracle/Makefile"; }

}  // extern "C"

// ReadSet::sort() of the reference (src/readset.cpp:42-51, comparator src/readset.h:39-66) on
// reads given by (name, source_id, first position): writes the resulting order.  Lets the tests pin
// the product's ReadSet.sort() tie-breaking to the real thing.
extern "C" int whref_sort_order(uint32_t n, const char *const *names, const int32_t *source_ids, const int32_t *firsts,
                                uint32_t *order) {
    ReadSet rs;
    for (uint32_t i = 0; i < n; ++i) {
        Read *r = new Read(names[i], 0, source_ids[i], (int)i);  // sample id carries the input index
        r->addVariant(firsts[i], 0, 1);
        r->addVariant(firsts[i] + 5, 1, 1);
        rs.add(r);
    }
    rs.sort();
    for (uint32_t i = 0; i < n; ++i) order[i] = (uint32_t)rs.get(i)->getSampleID();
    return 0;
}

// The reference's row-limited heuristic solver (src/pedmecheuristic.cpp) on the same flat arrays: outputs in the layout of
// whmec_heuristic_solution (include/whmec.h).  Sample ids are the pedigree indices (the reference asks for zero-indexed ids).
extern "C" int whref_heuristic(const whmec_problem *p, uint32_t row_limit, int allow_mutations, whmec_heuristic_solution *s, char *err,
                               size_t errlen) {
    try {
        Built b;
        build(p, b);
        PedMecHeuristic h(b.rs.get(), b.recomb, b.ped.get(), p->distrust != 0, &b.positions, row_limit, allow_mutations != 0, 0);
        h.solve();
        std::unique_ptr<Bipartition> part(h.getOptBipartition());
        std::unique_ptr<std::vector<Transmission>> tv(h.getOptTransmission());
        const auto haps = h.getOptHaplotypes();
        std::unique_ptr<std::vector<std::vector<std::pair<uint32_t, uint32_t>>>> mut(h.getMutations());
        s->score = h.getOptScore();
        s->n_samples = (uint32_t)haps.size();
        if (s->partition)
            for (uint32_t r = 0; r < p->n_reads; ++r) s->partition[r] = (*part)[r] ? 1 : 0;
        if (s->transmission)
            for (uint32_t k = 0; k < p->n_cols; ++k) s->transmission[k] = (*tv)[k];
        const size_t n = p->n_cols;
        if (s->mutated) std::memset(s->mutated, 0, haps.size() * 2 * n);
        for (size_t sid = 0; sid < haps.size(); ++sid) {
            for (int hp = 0; hp < 2; ++hp)
                for (size_t k = 0; k < n; ++k)
                    if (s->haplotypes) s->haplotypes[(sid * 2 + hp) * n + k] = haps[sid][hp][k];
            if (s->mutated)
                for (const auto &m : (*mut)[sid]) s->mutated[(sid * 2 + m.first) * n + m.second] = 1;
        }
        return WHMEC_OK;
    } catch (const std::exception &e) {
        set_err(err, errlen, e.what());
        return WHMEC_ERR_INPUT;
    }
}
