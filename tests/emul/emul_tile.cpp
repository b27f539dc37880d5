// TEST-ONLY host emulation of the tile path: runs the host planner (tile_plan.cpp) and the
// __host__ __device__ per-thread functions of tile_device.h serially.  See emul.cpp.
#include <cstring>
#include <string>
#include <vector>

#include "../../whatshap_b200/csrc/pack.h"
#include "../../whatshap_b200/csrc/tile_plan.h"
#include "../../whatshap_b200/csrc/tile_device.h"
#include "../../whatshap_b200/csrc/tile_fast.h"

#include <cstdlib>

using namespace whmec;

// Steady-state columns (TileCol::pad0 != 0) run the kernel's column_fast code thread by thread, as tile_panel_kernel
// dispatches it; the warp ballots that carry the back-pointer bits are assembled from the lanes' recorded predicates.
// WHEMUL_TILE_FAST=0 sends these columns through the generic tile_eval loop instead (the two must agree).
namespace {

uint64_t g_last_fast_columns = 0, g_last_packed_columns = 0;

struct RecordEmit {
    uint32_t *words;  // back-pointer words of this warp (ballot order) ...
    uint32_t lane;
    uint32_t *tile_words;  // ... or of the whole tile (thread-packed layout)
    uint32_t tid, bits_per_thread;
    void operator()(uint32_t word, bool bit) const {
        if (bit) words[word] |= 1u << lane;
    }
    void store(uint32_t bits) const {  // element `tid` of `bits_per_thread` bits, little endian like the kernel's u8 / u16 stores
        const uint32_t at = tid * bits_per_thread;
        tile_words[at >> 5] |= (bits & low_mask(bits_per_thread)) << (at & 31u);
    }
    uint32_t section;  // words from the tile's own section to its mirror section (mirrored panels with km != 0)
    void mirror(uint32_t word, bool bit) const {
        if (bit) words[section + word] |= 1u << lane;
    }
    void store_mirror(uint32_t bits) const {
        const uint32_t at = tid * bits_per_thread;
        tile_words[section + (at >> 5)] |= (bits & low_mask(bits_per_thread)) << (at & 31u);
    }
};

template <int LG, bool SHARE>
void fast_column(const TileCol &tc, const int32_t *TW, const int32_t *T5, uint32_t cg, const uint32_t *Sin, uint32_t *Sout, uint32_t *bpw) {
    constexpr uint32_t IT = 1u << LG;
    const bool mirror = tc.half && tc.km != 0;
    for (uint32_t w = 0; w < tc.bp_tile_stride; ++w) bpw[w] = 0;
    for (uint32_t tid = 0; tid < 1024; ++tid) {
        RecordEmit emit{bpw + (tid >> 5) * IT, tid & 31u, bpw, tid, tile_fast_bits_per_thread(tc), tc.bp_tile_words};
        const bool k0 = tc.K0 < TILE_KINF, packed = (tc.pad2 & 1u) != 0;
        if constexpr (SHARE && (LG == 2 || LG == 3)) {
            // the shape of the kernel's steady-state panels: run the kernel's own preparation of such a column (SteadyCol record +
            // constants of the panel, tile.cu: steady_columns) instead of column_fast_prep
            if (packed && !k0) {
                const SteadyCol sc = steady_col(tc, cg);
                const uint32_t obase = (tid >> 5) * (IT * 32u) + (tid & 31u);
                FastPrep<LG> pr;
                steady_prep<LG>(pr, sc, TW[tid >> 5], T5[tid & 31u], Sin, Sout, obase, obase & ((1u << (tc.l_in - 1)) - 1u),
                                1u << (tc.l_out - 1), popc32(obase));
                if (mirror) column_fast_body<LG, false, true, true, true>(pr, emit);
                else column_fast_body<LG, false, true, true, false>(pr, emit);
                continue;
            }
        }
#define EMUL_FAST(HK, PK, MR) column_fast<LG, HK, SHARE, PK, MR>(tc, TW, T5, cg, Sin, Sout, emit, tid)
        if (mirror) {
            if (packed) { if (k0) EMUL_FAST(true, true, true); else EMUL_FAST(false, true, true); }
            else { if (k0) EMUL_FAST(true, false, true); else EMUL_FAST(false, false, true); }
        } else {
            if (packed) { if (k0) EMUL_FAST(true, true, false); else EMUL_FAST(false, true, false); }
            else { if (k0) EMUL_FAST(true, false, false); else EMUL_FAST(false, false, false); }
        }
#undef EMUL_FAST
    }
}

void run_fast_column(const TileCol &tc, uint32_t tile, const uint32_t *Sin, uint32_t *Sout, uint32_t *bpw) {
    int32_t TW[32], T5[32];
    for (uint32_t i = 0; i < 32; ++i) {
        TW[i] = tile_fast_warp_entry(tc, tile, i);
        T5[i] = tile_fast_lane_entry(tc, i);
    }
    const uint32_t cg = tile_cg(tc, tile);
    if (tc.pad0 == 2) {
        switch (tc.pad1) {
            case 0: fast_column<0, true>(tc, TW, T5, cg, Sin, Sout, bpw); break;
            case 1: fast_column<1, true>(tc, TW, T5, cg, Sin, Sout, bpw); break;
            case 2: fast_column<2, true>(tc, TW, T5, cg, Sin, Sout, bpw); break;
            default: fast_column<3, true>(tc, TW, T5, cg, Sin, Sout, bpw); break;
        }
    } else {
        switch (tc.pad1) {
            case 0: fast_column<0, false>(tc, TW, T5, cg, Sin, Sout, bpw); break;
            case 1: fast_column<1, false>(tc, TW, T5, cg, Sin, Sout, bpw); break;
            case 2: fast_column<2, false>(tc, TW, T5, cg, Sin, Sout, bpw); break;
            case 3: fast_column<3, false>(tc, TW, T5, cg, Sin, Sout, bpw); break;
            default: fast_column<4, false>(tc, TW, T5, cg, Sin, Sout, bpw); break;
        }
    }
}

}  // namespace

// returns 100 when the planner declares the problem not eligible for the tile path
extern "C" int whemul_tile_solve(const whmec_problem *p, whmec_solution *s, uint32_t chunk, uint32_t *n_panels,
                                 char *err, size_t errlen) {
    Packed pk;
    std::string msg;
    auto fail = [&](int code) {
        if (err && errlen) {
            std::strncpy(err, msg.c_str(), errlen - 1);
            err[errlen - 1] = 0;
        }
        return code;
    };
    int rc = pack_problem(p, pk, msg);
    if (rc != WHMEC_OK) return fail(rc);
    const uint32_t n = pk.n;
    if (n == 0) {
        s->cost = 0;
        if (s->partition) std::memset(s->partition, 1, p->n_reads);
        return WHMEC_OK;
    }
    TileSchedule ts;
    plan_tiles(pk, ts);
    if (!ts.eligible) {
        msg = ts.why;
        return fail(100);
    }
    if (n_panels) *n_panels = (uint32_t)ts.panels.size();
    std::vector<uint32_t> state(ts.state_words + 1, 0xDEADBEEF), arena(ts.bp_words + 1, 0);
    const uint32_t n_chains = (uint32_t)pk.chain_begin.size() - 1;
    std::vector<uint64_t> chain_key(n_chains, KEY_INF);
    std::vector<uint32_t> bufA(1u << TILE_SMAX), bufB(1u << TILE_SMAX);
    std::vector<int32_t> TL(TILE_TL_SIZE), TH(TILE_TH_SIZE);
    const char *fast_env = std::getenv("WHEMUL_TILE_FAST");
    const bool use_fast = !(fast_env && fast_env[0] == '0');
    uint64_t fast_columns = 0, packed_columns = 0;
    for (size_t r = 0; r + 1 < ts.round_begin.size(); ++r)
        for (uint32_t pi = ts.round_begin[r]; pi < ts.round_begin[r + 1]; ++pi) {
            const Panel &P = ts.panels[pi];
            for (uint32_t t = 0; t < (1u << (P.g - P.half)); ++t) {  // mirrored panel: tiles with top tile-id bit 0 only
                uint32_t *Sin = bufA.data(), *Sout = bufB.data();
                if (P.fresh) Sin[0] = 0;
                else if (P.in_layout == 1) {
                    const uint32_t told = t & low_mask(P.in_gold);
                    for (uint32_t i = 0; i < (1u << P.s_in); ++i) {
                        uint32_t tA = i & low_mask(P.in_gA), at = (told << P.in_j) + (i >> P.in_gA);  // producer tile, its local index
                        if (P.in_half && ((tA >> (P.in_gA - 1)) & 1u)) {  // not computed: complemented index of the mirror tile
                            tA = ~tA & low_mask(P.in_gA);
                            at = ~at & low_mask(P.in_sA);
                        }
                        Sin[i] = state[P.in_off + ((uint64_t)tA << P.in_sA) + at];
                    }
                } else {
                    const uint32_t gpart = pdep32(t, P.gmask_in);
                    const uint32_t fmask_in = P.lmask_in | P.gmask_in;
                    for (uint32_t l = 0; l < (1u << P.s_in); ++l) {
                        uint32_t e = pdep32(l, P.lmask_in) | gpart;
                        if (P.in_half && (e & P.in_top)) e = ~e & fmask_in;
                        Sin[l] = state[P.in_off + e];
                    }
                }
                for (uint32_t k = P.col_begin; k < P.col_end; ++k) {
                    const TileCol &tc = ts.cols[k];
                    for (uint32_t i = 0; i < TILE_TL_SIZE; ++i) TL[i] = tile_tl_entry(tc, i);
                    for (uint32_t i = 0; i < TILE_TH_SIZE; ++i) TH[i] = tile_th_entry(tc, t, i);
                    TileCtx c{&tc, t, TL.data(), TH.data(), tile_cg(tc, t), Sin};
                    const uint32_t m = tc.l_in + tc.n_new;
                    if (tc.kind == 1) {
                        const uint32_t gpart = pdep32(t, ~tc.lmask_col & low_mask(pk.cols[k].a));
                        uint64_t best = KEY_INF;
                        uint32_t step = chunk ? chunk : (1u << m);
                        for (uint32_t x0 = 0; x0 < (1u << m); x0 += step) {
                            uint32_t x1 = x0 + step < (1u << m) ? x0 + step : (1u << m);
                            uint64_t key = tile_eval_end(c, gpart, x0, x1);
                            if (key < best) best = key;
                        }
                        if (best < chain_key[P.chain]) chain_key[P.chain] = best;
                    } else if (tc.pad0 && use_fast) {
                        run_fast_column(tc, t, Sin, Sout, arena.data() + tc.bp_off + (uint64_t)t * tc.bp_tile_stride);
                        ++fast_columns;
                        packed_columns += (tc.pad2 & 1u) != 0;
                        std::swap(Sin, Sout);
                    } else {
                        const uint32_t ncand = 1u << tc.d;
                        for (uint32_t o = 0; o < (1u << tc.l_out); ++o) {
                            uint64_t best = KEY_INF, mbest = KEY_INF;
                            uint32_t step = chunk ? chunk : ncand;
                            for (uint32_t r0 = 0; r0 < ncand; r0 += step) {
                                uint32_t r1 = r0 + step < ncand ? r0 + step : ncand;
                                uint64_t mkey;
                                uint64_t key = tile_eval(c, o, r0, r1, &mkey);
                                if (key < best) best = key;
                                if (mkey < mbest) mbest = mkey;
                            }
                            Sout[o] = (uint32_t)(best >> 32);
                            // WHEMUL_TILE_FAST=0 on a column the planner marked thread-packed: same bit positions as the fast code
                            const uint32_t at = (tc.pad0 && (tc.pad2 & 1u)) ? tile_packed_bit_index(tc, o) : o;
                            const uint64_t slice = tc.bp_off + (uint64_t)t * tc.bp_tile_stride;
                            bp_store_serial(arena.data(), slice, tc.bp_width, at, (uint32_t)best);
                            if (tc.half && tc.km != 0) bp_store_serial(arena.data(), slice + tc.bp_tile_words, tc.bp_width, at, (uint32_t)mbest);
                        }
                        std::swap(Sin, Sout);
                    }
                }
                if (!P.ends_chain && P.out_layout == 1) {
                    for (uint32_t l = 0; l < (1u << P.s_out); ++l) state[P.out_off + ((uint64_t)t << P.s_out) + l] = Sin[l];
                } else if (!P.ends_chain) {
                    const uint32_t gpart = pdep32(t, P.gmask_out);
                    for (uint32_t l = 0; l < (1u << P.s_out); ++l) state[P.out_off + (pdep32(l, P.lmask_out) | gpart)] = Sin[l];
                }
            }
        }
    std::vector<uint32_t> pidx(n), ptv(n, 0);
    uint64_t total = 0;
    for (uint32_t c = 0; c < n_chains; ++c) {
        total += chain_key[c] >> 32;
        tile_backtrace_chain(pk.cols.data(), ts.cols.data(), arena.data(), pk.chain_begin[c], pk.chain_begin[c + 1] - 1,
                             chain_key[c], pidx.data());
    }
    s->cost = (uint32_t)total;
    g_last_fast_columns = fast_columns;
    g_last_packed_columns = packed_columns;
    rc = build_outputs(pk, pidx.data(), ptv.data(), s, msg);
    if (rc != WHMEC_OK) return fail(rc);
    return WHMEC_OK;
}

// (tile, column) pairs the last whemul_tile_solve ran through column_fast
extern "C" uint64_t whemul_last_fast_columns(void) { return g_last_fast_columns; }
// ... of which with thread-packed back-pointer bits (WHMEC_TILE_PACKED_BP=1)
extern "C" uint64_t whemul_last_packed_columns(void) { return g_last_packed_columns; }

// planner statistics only (no DP): panels, rounds, total tiles, max tiles per round, state/bp words
extern "C" int whemul_plan_info(const whmec_problem *p, uint64_t *out8) {
    Packed pk;
    std::string msg;
    int rc = pack_problem(p, pk, msg);
    if (rc != WHMEC_OK) return rc;
    TileSchedule ts;
    plan_tiles(pk, ts);
    if (!ts.eligible) return 100;
    out8[0] = ts.panels.size();
    out8[1] = ts.round_tiles.size();
    uint64_t tot = 0, mx = 0;
    for (uint32_t t : ts.round_tiles) { tot += t; mx = std::max<uint64_t>(mx, t); }
    out8[2] = tot;
    out8[3] = mx;
    out8[4] = ts.state_words;
    out8[5] = ts.bp_words;
    out8[6] = ts.state_traffic_bytes;
    out8[7] = pk.stats.algorithmic_bytes;
    return 0;
}

// number of panel hand-offs that use the tile-major state layout / total hand-offs
extern "C" int whemul_plan_handoffs(const whmec_problem *p, uint64_t *out2) {
    Packed pk;
    std::string msg;
    int rc = pack_problem(p, pk, msg);
    if (rc != WHMEC_OK) return rc;
    TileSchedule ts;
    plan_tiles(pk, ts);
    if (!ts.eligible) return 100;
    out2[0] = out2[1] = 0;
    for (const Panel &P : ts.panels) {
        if (!P.fresh) out2[1]++;
        if (P.in_layout == 1) out2[0]++;
    }
    return 0;
}

#include <malloc.h>

#include <chrono>
// host-side timing of the packer and the tile planner (milliseconds)
static void keep_heap_like_the_library() {
    static bool once = false;
    if (once) return;
    once = true;
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
}

extern "C" int whemul_time_host(const whmec_problem *p, double *out2) {
    keep_heap_like_the_library();
    Packed pk;
    std::string msg;
    auto t0 = std::chrono::steady_clock::now();
    int rc = pack_problem(p, pk, msg);
    auto t1 = std::chrono::steady_clock::now();
    if (rc != WHMEC_OK) return rc;
    TileSchedule ts;
    plan_tiles(pk, ts);
    auto t2 = std::chrono::steady_clock::now();
    out2[0] = std::chrono::duration<double, std::milli>(t1 - t0).count();
    out2[1] = std::chrono::duration<double, std::milli>(t2 - t1).count();
    return 0;
}

// host-side timing as whmec_solve sees it for a single individual: packer without deltas, tile planner,
// super-read construction on an arbitrary path (milliseconds); heap retention as in the library
extern "C" int whemul_time_host_product(const whmec_problem *p, double *out3) {
    keep_heap_like_the_library();
    Packed pk;
    std::string msg;
    auto t0 = std::chrono::steady_clock::now();
    int rc = pack_problem(p, pk, msg, false);
    auto t1 = std::chrono::steady_clock::now();
    if (rc != WHMEC_OK) return rc;
    TileSchedule ts;
    plan_tiles(pk, ts);
    auto t2 = std::chrono::steady_clock::now();
    std::vector<uint32_t> pidx(pk.n), ptv(pk.n, 0);
    for (uint32_t k = 0; k < pk.n; ++k) pidx[k] = (k * 2654435761u) & low_mask(pk.cols[k].a);
    std::vector<uint8_t> part(p->n_reads), sra((size_t)p->n_ind * 2 * pk.n);
    std::vector<uint32_t> srq((size_t)p->n_ind * pk.n);
    whmec_solution s{};
    s.partition = part.data();
    s.sr_allele = sra.data();
    s.sr_quality = srq.data();
    auto t3 = std::chrono::steady_clock::now();
    build_outputs(pk, pidx.data(), ptv.data(), &s, msg);
    auto t4 = std::chrono::steady_clock::now();
    out3[0] = std::chrono::duration<double, std::milli>(t1 - t0).count();
    out3[1] = std::chrono::duration<double, std::milli>(t2 - t1).count();
    out3[2] = std::chrono::duration<double, std::milli>(t4 - t3).count();
    return 0;
}

// Host phases of a pedigree solve: packer WITH the cost-function deltas + output pass on an arbitrary path (out2: ms)
extern "C" int whemul_time_host_pedigree(const whmec_problem *p, double *out2) {
    keep_heap_like_the_library();
    Packed pk;
    std::string msg;
    auto t0 = std::chrono::steady_clock::now();
    int rc = pack_problem(p, pk, msg, true);
    auto t1 = std::chrono::steady_clock::now();
    if (rc != WHMEC_OK) return rc;
    std::vector<uint32_t> pidx(pk.n), ptv(pk.n, 0);
    for (uint32_t k = 0; k < pk.n; ++k) pidx[k] = (k * 2654435761u) & low_mask(pk.cols[k].a);
    std::vector<uint8_t> part(p->n_reads), sra((size_t)p->n_ind * 2 * pk.n);
    std::vector<uint32_t> srq((size_t)p->n_ind * pk.n);
    whmec_solution s{};
    s.partition = part.data();
    s.sr_allele = sra.data();
    s.sr_quality = srq.data();
    auto t3 = std::chrono::steady_clock::now();
    build_outputs(pk, pidx.data(), ptv.data(), &s, msg);
    auto t4 = std::chrono::steady_clock::now();
    out2[0] = std::chrono::duration<double, std::milli>(t1 - t0).count();
    out2[1] = std::chrono::duration<double, std::milli>(t4 - t3).count();
    return 0;
}

// FNV-1a digest of everything the planner produces (to hold planner rewrites to byte-identical schedules)
extern "C" int whemul_plan_digest(const whmec_problem *p, uint64_t *digest) {
    Packed pk;
    std::string msg;
    int rc = pack_problem(p, pk, msg, false);
    if (rc != WHMEC_OK) return rc;
    TileSchedule ts;
    plan_tiles(pk, ts);
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void *ptr, size_t bytes) {
        const unsigned char *b = (const unsigned char *)ptr;
        for (size_t i = 0; i < bytes; ++i) h = (h ^ b[i]) * 1099511628211ull;
    };
    const uint64_t head[5] = {ts.eligible, ts.state_words, ts.bp_words, ts.state_traffic_bytes, ts.panels.size()};
    mix(head, sizeof head);
    mix(ts.why.data(), ts.why.size());
    if (ts.eligible) {
        mix(ts.cols.data(), ts.cols.size() * sizeof(TileCol));
        mix(ts.panels.data(), ts.panels.size() * sizeof(Panel));
        mix(ts.round_begin.data(), ts.round_begin.size() * 4);
        mix(ts.round_tiles.data(), ts.round_tiles.size() * 4);
    }
    mix(pk.cols.data(), pk.cols.size() * sizeof(ColMeta));
    mix(pk.act_read.data(), pk.act_read.size() * 4);
    mix(pk.act_phred.data(), pk.act_phred.size() * 4);
    mix(pk.act_allele.data(), pk.act_allele.size());
    mix(pk.fn_c0.data(), pk.fn_c0.size() * 4);
    mix(pk.fn_asg.data(), pk.fn_asg.size() * 4);
    mix(pk.fn_group.data(), pk.fn_group.size() * 4);
    *digest = h;
    return 0;
}
