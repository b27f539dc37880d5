// TEST-ONLY host emulation of the column kernel's driver loop.
//
// Runs pack.cpp and the __host__ __device__ per-thread functions of dp_device.h (the same code
// the CUDA kernels execute) serially on the CPU so that their logic can be checked against the
// oracle in the GPU-less authoring container.  It is built into tests/emul/libwhemul.so by
// tests/emul/Makefile, loaded only by tests/test_emulation.py, and is NOT a fallback: the product
// library (whatshap_b200/csrc) contains no CPU execution path.
#include <cstring>
#include <string>
#include <vector>

#include "../../whatshap_b200/csrc/pack.h"
#include "../../whatshap_b200/csrc/dp_device.h"

using namespace whmec;

extern "C" int whemul_solve(const whmec_problem *p, whmec_solution *s, uint32_t chunk, char *err, size_t errlen) {
    Packed pk;
    std::string msg;
    int rc = pack_problem(p, pk, msg);
    auto fail = [&](int code) {
        if (err && errlen) {
            std::strncpy(err, msg.c_str(), errlen - 1);
            err[errlen - 1] = 0;
        }
        return code;
    };
    if (rc != WHMEC_OK) return fail(rc);
    const uint32_t n = pk.n, T = pk.T, tb = pk.tb;
    if (n == 0) {
        s->cost = 0;
        if (s->partition) std::memset(s->partition, 1, p->n_reads);
        return WHMEC_OK;
    }
    std::vector<uint32_t> arena(pk.bp_words + 1, 0), prev, cur, prevm;
    std::vector<uint8_t> prevarg;
    bool have_xform = false;
    for (uint32_t k = 0; k < n; ++k) {
        const ColMeta &m = pk.cols[k];
        if (m.d + tb > 32) { msg = "d + tb > 32"; return fail(WHMEC_ERR_UNSUPPORTED); }
        uint64_t nout = (uint64_t)1 << m.f, ncand = (uint64_t)1 << m.d;
        cur.assign(nout * T, UMAX);
        // per-column lookup tables, as a thread block builds them in shared memory
        const uint32_t nf_col = pk.fn_group[m.grp_off + T];
        std::vector<uint32_t> pd_lo(TAB_SIZE), pd_hi(TAB_SIZE);
        std::vector<int32_t> tlo((size_t)nf_col * TAB_SIZE), thi((size_t)nf_col * TAB_SIZE);
        const uint32_t keep_lo = lowest_set_bits(m.keep, TAB_BITS), keep_hi = lowest_set_bits(m.keep & ~keep_lo, TAB_BITS);
        for (uint32_t vv = 0; vv < TAB_SIZE; ++vv) {
            pd_lo[vv] = pdep32(vv, keep_lo);
            pd_hi[vv] = pdep32(vv, keep_hi);
        }
        for (uint32_t F = 0; F < nf_col; ++F)
            for (uint32_t half = 0; half < 2; ++half)
                for (uint32_t hi4 = 0; hi4 < 16; ++hi4)
                    build_cost_table_run(pk.fn_delta.data() + (size_t)(m.fn_off + F) * FN_STRIDE, half, hi4,
                                         (half ? thi.data() : tlo.data()) + (size_t)F * TAB_SIZE + 16 * hi4);
        ColTables tab{pd_lo.data(), pd_hi.data(), m.keep & ~keep_lo & ~keep_hi, tlo.data(), thi.data()};
        const bool use_tab = (k % 2) == 0;  // exercise both initialisation paths
        for (uint64_t o = 0; o < nout; ++o)
            for (uint32_t i = 0; i < T; ++i) {
                ColView v;
                v.m = &m; v.T = T; v.tb = tb;
                uint32_t g0 = pk.fn_group[m.grp_off + i], g1 = pk.fn_group[m.grp_off + i + 1];
                v.fn_c0 = pk.fn_c0.data() + m.fn_off + g0;
                v.fn_delta = pk.fn_delta.data() + (size_t)(m.fn_off + g0) * FN_STRIDE;
                v.nf = g1 - g0;
                v.prev = prev.data();
                v.tab = use_tab ? &tab : nullptr;
                v.tab_fn0 = g0;
                v.prevm = have_xform ? prevm.data() : nullptr;
                v.prevarg = have_xform ? prevarg.data() : nullptr;
                uint64_t best = KEY_INF;
                uint64_t step = chunk ? chunk : ncand;
                for (uint64_t r0 = 0; r0 < ncand; r0 += step) {  // chunked like the atomic kernel
                    uint64_t r1 = r0 + step < ncand ? r0 + step : ncand;
                    uint64_t key = eval_candidates(v, (uint32_t)o, i, (uint32_t)r0, (uint32_t)r1);
                    if (key < best) best = key;
                }
                cur[o * T + i] = (uint32_t)(best >> 32);
                bp_store_serial(arena.data(), m.bp_off, m.bp_width, o * T + i, (uint32_t)best & low_mask(m.d + tb));
            }
        prev.swap(cur);
        // every third hand-over keeps raw values, the others pre-apply the transition minimum of the next
        // column (as the batched pedigree kernels do in their epilogue)
        have_xform = (k + 1 < n) && (k % 3 != 2);
        if (have_xform) {
            prevm.assign(nout * T, UMAX);
            prevarg.assign(nout * T, 0);
            for (uint64_t o = 0; o < nout; ++o)
                for (uint32_t i = 0; i < T; ++i) {
                    uint32_t arg;
                    prevm[o * T + i] = transition_min(&prev[o * T], T, i, pk.cols[k + 1].rc, &arg);
                    prevarg[o * T + i] = (uint8_t)arg;
                }
        }
    }
    std::vector<uint32_t> pidx(n), ptv(n);
    BtView bv{pk.cols.data(), arena.data(), T, tb};
    uint32_t cost, x, tv, ptvv;
    pick_optimum(pk.cols[n - 1], prev.data(), arena.data(), T, tb, &cost, &x, &tv, &ptvv);
    backtrace_range(bv, n - 1, 0, x, tv, ptvv, pidx.data(), ptv.data());
    s->cost = cost;
    rc = build_outputs(pk, pidx.data(), ptv.data(), s, msg);
    if (rc != WHMEC_OK) return fail(rc);
    return WHMEC_OK;
}
