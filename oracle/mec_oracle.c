/*
 * TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT PATH.
 *
 * Plain-C restatement of WhatsHap's weighted-MEC / PedMEC dynamic program, used as the
 * checker for the CUDA path (tests/, __graft_entry__.smoke(), bench.py cpu_baseline "port").
 * Nothing under whatshap_b200/ links, imports or executes this file.
 *
 * Parity status: PINNED.  This restatement is validated bit-for-bit against the compiled,
 * unmodified reference (oracle/_ref/libwhref.so, built by oracle/Makefile) by
 * tests/test_oracle.py and against the reference-generated golden vectors in tests/golden/.
 *
 * Each function cites the reference code it follows (paths relative to the whatshap tree).
 * The restatement keeps every table of every column (no sqrt(n) checkpointing,
 * src/pedigreedptable.cpp:103-134) — checkpointing changes memory use, not results.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#include "../include/whmec.h"

#define UMAX 0xFFFFFFFFu
#define MAX_IND 16
#define MAX_PART 10

static void set_err(char *err, size_t errlen, const char *msg) {
    if (err && errlen) {
        strncpy(err, msg, errlen - 1);
        err[errlen - 1] = 0;
    }
}

/* src/pedigreepartitions.cpp:7-42 — haplotype -> IBD partition map for transmission value tv. */
static int h2p_rec(const whmec_problem *p, uint32_t tv, const int *triple_of, int (*h2p)[2], uint32_t i, int depth) {
    if (h2p[i][0] != -1) return 0;
    int tr = triple_of[i];
    if (tr < 0 || depth > (int)p->n_ind) return -1;
    uint32_t f = p->trios[3 * tr], m = p->trios[3 * tr + 1];
    if (h2p_rec(p, tv, triple_of, h2p, f, depth + 1)) return -1;
    if (h2p_rec(p, tv, triple_of, h2p, m, depth + 1)) return -1;
    /* pedigreepartitions.cpp:38-41: index is !(bit) */
    h2p[i][0] = h2p[f][((tv >> (2 * tr)) & 1) ? 0 : 1];
    h2p[i][1] = h2p[m][((tv >> (2 * tr + 1)) & 1) ? 0 : 1];
    return 0;
}

static int build_h2p(const whmec_problem *p, uint32_t tv, int (*h2p)[2]) {
    int triple_of[MAX_IND];
    for (uint32_t i = 0; i < p->n_ind; ++i) {
        triple_of[i] = -1;
        h2p[i][0] = h2p[i][1] = -1;
    }
    for (uint32_t r = 0; r < p->n_trios; ++r) triple_of[p->trios[3 * r + 2]] = (int)r;
    int q = 0;
    for (uint32_t i = 0; i < p->n_ind; ++i)
        if (triple_of[i] == -1) { /* founders, pedigreepartitions.cpp:18-23 */
            h2p[i][0] = q;
            h2p[i][1] = q + 1;
            q += 2;
        }
    for (uint32_t i = 0; i < p->n_ind; ++i)
        if (h2p_rec(p, tv, triple_of, h2p, i, 0)) return -1;
    return 0;
}

typedef struct {
    uint32_t n;          /* number of allowed assignments */
    uint32_t asg[1 << MAX_PART];
    uint32_t base[1 << MAX_PART];
} asg_list;

/* src/pedigreecolumncostcomputer.cpp:14-50 — allowed allele assignments of column k under h2p. */
static void build_assignments(const whmec_problem *p, uint32_t k, int (*h2p)[2], uint32_t P, asg_list *L) {
    L->n = 0;
    for (uint32_t A = 0; A < (1u << P); ++A) {
        int ok = 1;
        unsigned int cost = 0;
        for (uint32_t i = 0; i < p->n_ind; ++i) {
            uint32_t a0 = (A >> h2p[i][0]) & 1, a1 = (A >> h2p[i][1]) & 1;
            if (p->distrust) {
                /* :33-37  unsigned += double  (converted to double, added, truncated) */
                double g = p->gl[((size_t)i * p->n_cols + k) * 3 + (a0 + a1)];
                cost = (unsigned int)((double)cost + g);
            } else {
                /* :39-43 genotype (sorted multiset) must equal the pedigree's */
                uint8_t gt = p->gt[(size_t)i * p->n_cols + k];
                if (gt != a0 + a1) {
                    ok = 0;
                    break;
                }
            }
        }
        if (ok) {
            L->asg[L->n] = A;
            L->base[L->n] = cost;
            L->n++;
        }
    }
}

typedef struct {
    uint32_t a;                 /* active reads */
    uint32_t reads[32];         /* read index of bit j */
    uint8_t allele[32];         /* 0/1/2(blank) */
    uint32_t phred[32];
    uint32_t ind[32];
} column_t;

/* src/pedigreecolumncostcomputer.cpp:53-76 (set_partitioning): cost_partition[p][b] for index x. */
static void partition_costs(const column_t *c, uint32_t x, int (*h2p)[2], uint32_t P, uint32_t cp[][2]) {
    for (uint32_t q = 0; q < P; ++q) cp[q][0] = cp[q][1] = 0;
    for (uint32_t j = 0; j < c->a; ++j) {
        uint32_t bit = (x >> j) & 1;
        int part = h2p[c->ind[j]][bit];
        if (c->allele[j] == 0) cp[part][1] += c->phred[j];       /* REF: cost if partition allele is 1 */
        else if (c->allele[j] == 1) cp[part][0] += c->phred[j];  /* ALT */
    }
}

/* src/pedigreecolumncostcomputer.cpp:101-114 (get_cost). */
static uint32_t cell_cost(const asg_list *L, uint32_t P, uint32_t cp[][2]) {
    uint32_t best = UMAX;
    for (uint32_t n = 0; n < L->n; ++n) {
        uint32_t cost = L->base[n];
        for (uint32_t q = 0; q < P; ++q) cost += cp[q][(L->asg[n] >> q) & 1];
        if (cost < best) best = cost;
    }
    return best;
}

static uint32_t popcount32(uint32_t x) {
    uint32_t c = 0;
    for (; x; x >>= 1) c += x & 1;
    return c;
}

int whoracle_solve(const whmec_problem *p, whmec_solution *s, char *err, size_t errlen) {
    const uint32_t n = p->n_cols;
    if (p->n_ind > MAX_IND || p->n_trios > p->n_ind) {
        set_err(err, errlen, "oracle: too many individuals");
        return WHMEC_ERR_UNSUPPORTED;
    }
    const uint32_t P = 2 * (p->n_ind - p->n_trios);
    if (P > MAX_PART || 2 * p->n_trios > 8) {
        set_err(err, errlen, "oracle: pedigree too large");
        return WHMEC_ERR_UNSUPPORTED;
    }
    const uint32_t T = 1u << (2 * p->n_trios); /* pedigreedptable.cpp:27,191 */

    /* --- read spans; ColumnIterator ctor checks (src/columniterator.cpp:25-33) --- */
    uint32_t *first = (uint32_t *)malloc(sizeof(uint32_t) * (p->n_reads + 1));
    uint32_t *last = (uint32_t *)malloc(sizeof(uint32_t) * (p->n_reads + 1));
    uint32_t prev_first = 0;
    for (uint32_t r = 0; r < p->n_reads; ++r) {
        uint64_t b = p->read_off[r], e = p->read_off[r + 1];
        if (e <= b) {
            set_err(err, errlen, "No variants present");
            free(first); free(last);
            return WHMEC_ERR_INPUT;
        }
        for (uint64_t q = b + 1; q < e; ++q)
            if (p->ent_col[q] <= p->ent_col[q - 1]) {
                set_err(err, errlen, "ColumnIterator: encountered read with unsorted variants.");
                free(first); free(last);
                return WHMEC_ERR_INPUT;
            }
        first[r] = p->ent_col[b];
        last[r] = p->ent_col[e - 1];
        if (first[r] < prev_first) {
            set_err(err, errlen, "ColumnIterator: reads in ReadSet are not sorted.");
            free(first); free(last);
            return WHMEC_ERR_INPUT;
        }
        prev_first = first[r];
        if (last[r] >= n) {
            set_err(err, errlen, "oracle: entry column out of range");
            free(first); free(last);
            return WHMEC_ERR_INPUT;
        }
    }

    int (*h2p)[MAX_IND][2] = (int (*)[MAX_IND][2])malloc(sizeof(int) * T * MAX_IND * 2);
    for (uint32_t t = 0; t < T; ++t)
        if (build_h2p(p, t, h2p[t])) {
            set_err(err, errlen, "oracle: malformed pedigree");
            free(first); free(last); free(h2p);
            return WHMEC_ERR_INPUT;
        }

    int rc = WHMEC_OK;
    if (n == 0) { /* pedigreedptable.cpp:88-92 */
        s->cost = 0;
        if (s->partition) memset(s->partition, 1, p->n_reads);
        free(first); free(last); free(h2p);
        return WHMEC_OK;
    }

    /* --- column structure: ColumnIterator::get_next (columniterator.cpp:91-139) and
     *     ColumnIndexingScheme (columnindexingscheme.cpp:7-34,62-85) --- */
    column_t *cols = (column_t *)calloc(n, sizeof(column_t));
    uint32_t *bw = (uint32_t *)calloc(n, sizeof(uint32_t));
    uint32_t *keep = (uint32_t *)calloc(n, sizeof(uint32_t)); /* bit j set: read of bit j also active in k+1 */
    uint64_t *cursor = (uint64_t *)malloc(sizeof(uint64_t) * (p->n_reads + 1));
    for (uint32_t r = 0; r < p->n_reads; ++r) cursor[r] = p->read_off[r];
    for (uint32_t k = 0; k < n && rc == WHMEC_OK; ++k) {
        column_t *c = &cols[k];
        for (uint32_t r = 0; r < p->n_reads; ++r) {
            if (first[r] > k) break; /* reads sorted by first */
            if (last[r] < k) continue;
            if (c->a >= 32) { /* graycodes.cpp:12 */
                set_err(err, errlen, "oracle: more than 32 active reads in a column");
                rc = WHMEC_ERR_UNSUPPORTED;
                break;
            }
            while (p->ent_col[cursor[r]] < k) cursor[r]++;
            uint32_t j = c->a++;
            c->reads[j] = r;
            c->ind[j] = p->read_ind[r];
            if (p->ent_col[cursor[r]] == k) {
                c->allele[j] = p->ent_allele[cursor[r]];
                c->phred[j] = p->ent_phred[cursor[r]];
                if (c->allele[j] > 2) {
                    set_err(err, errlen, "oracle: allele must be 0, 1 or 2 (blank)");
                    rc = WHMEC_ERR_INPUT;
                }
            } else { /* gap inside the read: BLANK, phred 0 (columniterator.cpp:131) */
                c->allele[j] = 2;
                c->phred[j] = 0;
            }
            if (c->ind[j] >= p->n_ind) {
                set_err(err, errlen, "oracle: read_ind out of range");
                rc = WHMEC_ERR_INPUT;
            }
        }
    }
    for (uint32_t k = 0; k < n && rc == WHMEC_OK; ++k) {
        if (k > 0) { /* backward projection width = |reads(k) ∩ reads(k-1)| */
            uint32_t i = 0, j = 0, w = 0;
            while (i < cols[k - 1].a && j < cols[k].a) {
                if (cols[k - 1].reads[i] == cols[k].reads[j]) { w++; i++; j++; }
                else if (cols[k - 1].reads[i] < cols[k].reads[j]) i++;
                else j++;
            }
            bw[k] = w;
        }
        if (k + 1 < n) {
            uint32_t i = 0, j = 0;
            while (i < cols[k + 1].a && j < cols[k].a) {
                if (cols[k + 1].reads[i] == cols[k].reads[j]) { keep[k] |= 1u << j; i++; j++; }
                else if (cols[k + 1].reads[i] < cols[k].reads[j]) i++;
                else j++;
            }
        }
    }

    /* --- forward sweep, pedigreedptable.cpp:177-335 --- */
    uint32_t **proj = (uint32_t **)calloc(n, sizeof(uint32_t *)); /* [f][T] */
    uint32_t **idxb = (uint32_t **)calloc(n, sizeof(uint32_t *));
    uint32_t **tvb = (uint32_t **)calloc(n, sizeof(uint32_t *));
    asg_list *L = (asg_list *)malloc(sizeof(asg_list) * T);
    uint32_t opt = UMAX, opt_index = 0, opt_tv = 0, opt_prev_tv = 0;
    uint32_t dpv[256], argj[256], cur[256];

    for (uint32_t k = 0; k < n && rc == WHMEC_OK; ++k) {
        const column_t *c = &cols[k];
        const int lastcol = (k + 1 == n);
        uint32_t fbits = popcount32(keep[k]);
        if (!lastcol) {
            size_t sz = ((size_t)1 << fbits) * T;
            proj[k] = (uint32_t *)malloc(sizeof(uint32_t) * sz);
            idxb[k] = (uint32_t *)malloc(sizeof(uint32_t) * sz);
            tvb[k] = (uint32_t *)malloc(sizeof(uint32_t) * sz);
            for (size_t q = 0; q < sz; ++q) proj[k][q] = idxb[k][q] = tvb[k][q] = UMAX; /* :213-229 */
        }
        for (uint32_t t = 0; t < T; ++t) build_assignments(p, k, h2p[t], P, &L[t]);
        const uint32_t bmask = (bw[k] >= 32) ? UMAX : ((1u << bw[k]) - 1);
        const uint64_t ncell = (uint64_t)1 << c->a;
        for (uint64_t r = 0; r < ncell && rc == WHMEC_OK; ++r) {
            /* Gray-code order (graycodes.cpp:26-43): the r-th index visited is r ^ (r >> 1) */
            uint32_t x = (uint32_t)(r ^ (r >> 1));
            uint32_t b = (k > 0) ? (x & bmask) : 0; /* columnindexingiterator.cpp:58-61 */
            int any = 0;
            for (uint32_t i = 0; i < T; ++i) {
                uint32_t cp[MAX_PART][2];
                partition_costs(c, x, h2p[i], P, cp);
                cur[i] = cell_cost(&L[i], P, cp);
                if (cur[i] < UMAX) any = 1;
            }
            for (uint32_t i = 0; i < T; ++i) { /* :264-300 */
                uint32_t mn = UMAX, mj = 0;
                for (uint32_t j = 0; j < T; ++j) {
                    uint32_t prev = (k > 0) ? proj[k - 1][(size_t)b * T + j] : 0;
                    uint32_t val;
                    if (cur[i] < UMAX && prev < UMAX) val = cur[i] + prev;
                    else val = UMAX;
                    if (val < UMAX) val += popcount32(i ^ j) * p->recombcost[k];
                    if (val < mn) { mn = val; mj = j; }
                }
                dpv[i] = mn;
                argj[i] = mj;
            }
            if (!any) { /* :301-303 */
                set_err(err, errlen, "Error: Mendelian conflict");
                rc = WHMEC_ERR_MENDELIAN;
                break;
            }
            if (lastcol) { /* :306-315 */
                for (uint32_t i = 0; i < T; ++i)
                    if (dpv[i] < opt) {
                        opt = dpv[i];
                        opt_index = x;
                        opt_tv = i;
                        opt_prev_tv = argj[i];
                    }
            } else { /* :317-325; forward projection = pext(x, keep) (columnindexingiterator.cpp:26-49) */
                uint32_t f = 0, o = 0;
                for (uint32_t j = 0; j < c->a; ++j)
                    if ((keep[k] >> j) & 1) f |= ((x >> j) & 1) << o++;
                for (uint32_t i = 0; i < T; ++i)
                    if (dpv[i] < proj[k][(size_t)f * T + i]) {
                        proj[k][(size_t)f * T + i] = dpv[i];
                        idxb[k][(size_t)f * T + i] = x;
                        tvb[k][(size_t)f * T + i] = argj[i];
                    }
            }
        }
    }

    /* --- backtrace, pedigreedptable.cpp:137-173 --- */
    if (rc == WHMEC_OK) {
        uint32_t *pidx = (uint32_t *)malloc(sizeof(uint32_t) * n);
        uint32_t *ptv = (uint32_t *)malloc(sizeof(uint32_t) * n);
        uint32_t vi = opt_index, vt = opt_tv, prev_tv = opt_prev_tv;
        pidx[n - 1] = vi;
        ptv[n - 1] = vt;
        for (uint32_t i = n - 1; i > 0; --i) {
            uint32_t bmask = (bw[i] >= 32) ? UMAX : ((1u << bw[i]) - 1);
            uint32_t b = vi & bmask;
            vi = idxb[i - 1][(size_t)b * T + prev_tv];
            vt = prev_tv;
            prev_tv = tvb[i - 1][(size_t)b * T + vt];
            pidx[i - 1] = vi;
            ptv[i - 1] = vt;
        }
        s->cost = opt;
        for (uint32_t k = 0; k < n; ++k) {
            if (s->path_index) s->path_index[k] = pidx[k];
            if (s->path_tv) s->path_tv[k] = ptv[k];
        }
        /* get_optimal_partitioning, pedigreedptable.cpp:391-406 + core.pyx:414 */
        if (s->partition) {
            memset(s->partition, 1, p->n_reads);
            for (uint32_t k = 0; k < n; ++k)
                for (uint32_t j = 0; j < cols[k].a; ++j)
                    if (((pidx[k] >> j) & 1) == 0) s->partition[cols[k].reads[j]] = 0;
        }
        /* get_super_reads, pedigreedptable.cpp:344-388 with get_alleles,
         * pedigreecolumncostcomputer.cpp:117-175 */
        for (uint32_t k = 0; k < n && rc == WHMEC_OK; ++k) {
            uint32_t t = ptv[k];
            asg_list *A = &L[0];
            build_assignments(p, k, h2p[t], P, A);
            uint32_t cp[MAX_PART][2];
            partition_costs(&cols[k], pidx[k], h2p[t], P, cp);
            uint32_t best = UMAX;
            uint32_t call[MAX_IND][2];
            uint32_t bfa[MAX_IND][2][2];
            for (uint32_t i = 0; i < p->n_ind; ++i) {
                call[i][0] = call[i][1] = 2; /* BLANK default ctor */
                bfa[i][0][0] = bfa[i][0][1] = bfa[i][1][0] = bfa[i][1][1] = UMAX;
            }
            for (uint32_t q = 0; q < A->n; ++q) {
                uint32_t cost = A->base[q];
                for (uint32_t pp = 0; pp < P; ++pp) cost += cp[pp][(A->asg[q] >> pp) & 1];
                int new_best = 0;
                if (cost <= best) { /* :131  '<=' : last tie wins */
                    best = cost;
                    new_best = 1;
                }
                for (uint32_t i = 0; i < p->n_ind; ++i) {
                    uint32_t a0 = (A->asg[q] >> h2p[t][i][0]) & 1, a1 = (A->asg[q] >> h2p[t][i][1]) & 1;
                    if (new_best) { call[i][0] = a0; call[i][1] = a1; }
                    if (cost < bfa[i][0][a0]) bfa[i][0][a0] = cost;
                    if (cost < bfa[i][1][a1]) bfa[i][1][a1] = cost;
                }
            }
            if (best == UMAX) { /* :155-157 */
                set_err(err, errlen, "Error: Mendelian conflict");
                rc = WHMEC_ERR_MENDELIAN;
                break;
            }
            for (uint32_t i = 0; i < p->n_ind; ++i) {
                uint32_t quality = 0;
                for (uint32_t h = 0; h < 2; ++h) { /* :160-171 */
                    int q = abs((int)bfa[i][h][0] - (int)bfa[i][h][1]);
                    quality = (uint32_t)q;
                    if (q == 0) call[i][h] = 3; /* EQUAL_SCORES */
                }
                if (s->sr_allele) {
                    s->sr_allele[((size_t)i * 2 + 0) * n + k] = (uint8_t)call[i][0];
                    s->sr_allele[((size_t)i * 2 + 1) * n + k] = (uint8_t)call[i][1];
                }
                if (s->sr_quality) s->sr_quality[(size_t)i * n + k] = quality;
            }
        }
        free(pidx);
        free(ptv);
    }

    for (uint32_t k = 0; k < n; ++k) { free(proj[k]); free(idxb[k]); free(tvb[k]); }
    free(proj); free(idxb); free(tvb); free(L);
    free(cols); free(bw); free(keep); free(cursor);
    free(first); free(last); free(h2p);
    return rc;
}

const char *whoracle_info(void) { return "plain-C restatement of PedigreeDPTable (oracle/mec_oracle.c)"; }
