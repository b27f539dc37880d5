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

uint64_t g_last_fast_columns = 0, g_last_packed_columns = 0, g_last_u16_columns = 0;

struct StoreEmit {
    uint32_t *slot;
    void operator()(uint32_t, bool) const {}
    void store(uint32_t bits) const { *slot = bits; }
};

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
};

template <int LG, bool SHARE>
void fast_column(const TileCol &tc, const int32_t *TW, const int32_t *T5, uint32_t cg, const uint32_t *Sin, uint32_t *Sout, uint32_t *bpw) {
    constexpr uint32_t IT = 1u << LG;
    for (uint32_t w = 0; w < ((1u << tc.l_out) + 31) / 32; ++w) bpw[w] = 0;
    for (uint32_t tid = 0; tid < 1024; ++tid) {
        RecordEmit emit{bpw + (tid >> 5) * IT, tid & 31u, bpw, tid, tile_fast_bits_per_thread(tc)};
        if (tc.pad2 & 1u) {
            if (tc.K0 >= TILE_KINF) column_fast<LG, false, SHARE, true>(tc, TW, T5, cg, Sin, Sout, emit, tid);
            else column_fast<LG, true, SHARE, true>(tc, TW, T5, cg, Sin, Sout, emit, tid);
        } else if (tc.K0 >= TILE_KINF) column_fast<LG, false, SHARE>(tc, TW, T5, cg, Sin, Sout, emit, tid);
        else column_fast<LG, true, SHARE>(tc, TW, T5, cg, Sin, Sout, emit, tid);
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
    uint64_t fast_columns = 0, packed_columns = 0, u16_columns = 0;
    for (size_t r = 0; r + 1 < ts.round_begin.size(); ++r)
        for (uint32_t pi = ts.round_begin[r]; pi < ts.round_begin[r + 1]; ++pi) {
            const Panel &P = ts.panels[pi];
            for (uint32_t t = 0; t < (1u << P.g); ++t) {
                uint32_t *Sin = bufA.data(), *Sout = bufB.data();
                if (P.fresh) Sin[0] = 0;
                else if (P.in_layout == 1) {
                    const uint32_t told = t & low_mask(P.in_gold);
                    for (uint32_t i = 0; i < (1u << P.s_in); ++i) {
                        const uint32_t tA = i & low_mask(P.in_gA), ll = i >> P.in_gA;
                        Sin[i] = state[P.in_off + ((uint64_t)told << P.in_j) + ((uint64_t)tA << P.in_sA) + ll];
                    }
                } else {
                    const uint32_t gpart = pdep32(t, P.gmask_in);
                    for (uint32_t l = 0; l < (1u << P.s_in); ++l) Sin[l] = state[P.in_off + (pdep32(l, P.lmask_in) | gpart)];
                }
                if (P.pad >> 31) {
                    // ---- packed 16-bit panel (experimental): rotate + convert the tile, sweep every column with
                    //      column_fast16, convert back into the canonical u32 order the write-back below expects
                    const uint32_t s_in = P.s_in, xp0 = s_in - 1, spread = P.pad & 0x7FFFFFFFu;
                    const uint32_t base = Sin[0] - spread;
                    std::vector<uint32_t> W0(1u << (s_in - 1), 0), W1(1u << (s_in - 1), 0);
                    for (uint32_t i = 0; i < (1u << s_in); ++i) {
                        const uint32_t rel = Sin[i] - base;
                        if (rel >= 32768u) { msg = "u16 panel: input beyond the planner's range bound"; return fail(101); }
                        const uint32_t xr = tile_u16_rotate(i, xp0);
                        W0[xr >> 1] |= rel << (16 * (xr & 1u));
                    }
                    uint32_t *Win = W0.data(), *Wout = W1.data();
                    for (uint32_t k = P.col_begin; k < P.col_end; ++k) {
                        const TileCol &tc = ts.cols[k];
                        if (!tile_is_u16(tc) || tile_u16_xpos(tc) != xp0 - (k - P.col_begin)) { msg = "u16 panel: column not marked"; return fail(103); }
                        TileCol16 c16;
                        tile_col16_from(tc, t, c16);
                        uint32_t TW2[32], T52[32];
                        for (uint32_t i = 0; i < 32; ++i) {
                            TW2[i] = tile_fast16_warp_entry(c16, i);
                            T52[i] = tile_fast16_lane_entry(c16, i);
                        }
                        const uint32_t cg = tile_cg(tc, t), nbits = 4u << (tc.l_out - 12);
                        uint32_t *bpw = arena.data() + tc.bp_off + (uint64_t)t * tc.bp_tile_words;
                        for (uint32_t w = 0; w < tc.bp_tile_words; ++w) bpw[w] = 0;
                        for (uint32_t tid = 0; tid < 1024; ++tid) {
                            uint32_t bits = 0;
                            StoreEmit emit{&bits};
                            if (tc.l_out == 14) column_fast16<2>(c16, TW2, T52, cg, Win, Wout, emit, tid);
                            else column_fast16<1>(c16, TW2, T52, cg, Win, Wout, emit, tid);
                            const uint32_t at = tid * nbits;  // element tid, as the kernel's u8 / u16 store
                            bpw[at >> 5] |= bits << (at & 31u);
                        }
                        for (uint32_t w = 0; w < (1u << (tc.l_out - 1)); ++w)
                            if ((Wout[w] & 0x8000u) || (Wout[w] & 0x80000000u)) { msg = "u16 panel: value beyond 2^15"; return fail(102); }
                        std::swap(Win, Wout);
                        ++fast_columns;
                        ++u16_columns;
                    }
                    const uint32_t xpo = xp0 - (P.col_end - P.col_begin);  // canonical position of X after the panel
                    for (uint32_t i = 0; i < (1u << P.s_out); ++i) {
                        const uint32_t xr = tile_u16_rotate(i, xpo);
                        Sin[i] = ((Win[xr >> 1] >> (16 * (xr & 1u))) & 0xFFFFu) + base;
                    }
                } else
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
                        run_fast_column(tc, t, Sin, Sout, arena.data() + tc.bp_off + (uint64_t)t * tc.bp_tile_words);
                        ++fast_columns;
                        packed_columns += (tc.pad2 & 1u) != 0;
                        std::swap(Sin, Sout);
                    } else {
                        const uint32_t ncand = 1u << tc.d;
                        for (uint32_t o = 0; o < (1u << tc.l_out); ++o) {
                            uint64_t best = KEY_INF;
                            uint32_t step = chunk ? chunk : ncand;
                            for (uint32_t r0 = 0; r0 < ncand; r0 += step) {
                                uint32_t r1 = r0 + step < ncand ? r0 + step : ncand;
                                uint64_t key = tile_eval(c, o, r0, r1);
                                if (key < best) best = key;
                            }
                            Sout[o] = (uint32_t)(best >> 32);
                            // WHEMUL_TILE_FAST=0 on a column the planner marked thread-packed: same bit positions as the fast code
                            const uint32_t at = (tc.pad0 && (tc.pad2 & 1u)) ? tile_packed_bit_index(tc, o) : o;
                            bp_store_serial(arena.data(), tc.bp_off + (uint64_t)t * tc.bp_tile_words, tc.bp_width, at, (uint32_t)best);
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
    g_last_u16_columns = u16_columns;
    rc = build_outputs(pk, pidx.data(), ptv.data(), s, msg);
    if (rc != WHMEC_OK) return fail(rc);
    return WHMEC_OK;
}

// (tile, column) pairs the last whemul_tile_solve ran through column_fast
extern "C" uint64_t whemul_last_fast_columns(void) { return g_last_fast_columns; }
// ... of which with thread-packed back-pointer bits (WHMEC_TILE_PACKED_BP=1)
extern "C" uint64_t whemul_last_packed_columns(void) { return g_last_packed_columns; }
// ... and (tile, column) pairs swept by column_fast16 inside packed 16-bit panels (WHMEC_TILE_U16=1)
extern "C" uint64_t whemul_last_u16_columns(void) { return g_last_u16_columns; }

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

// ---- the experimental packed 16-bit column (column_fast16, tile_fast.h) against column_fast on one random column:
// same tile state (canonical u32 layout vs the rotated u16 layout), same costs, same tie-breaking.  Returns the number of
// outputs compared, or a negative code on the first difference.
namespace {


template <int LG16>
long fast16_check(uint32_t seed, uint32_t cg, uint32_t pX) {
    constexpr uint32_t IT = 1u << LG16, l_out = 12 + LG16, l_in = l_out, m = l_in + 1;
    uint64_t rs = seed * 0x9E3779B97F4A7C15ull + 12345;
    auto rnd = [&](uint32_t n) { rs = rs * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)((rs >> 33) % n); };
    if (pX < 1 || pX >= l_in) return -100;
    // reads of the column: phred and allele per canonical cell bit (bit 0 ends here, bit l_in starts here)
    int32_t w[16];
    uint32_t K1 = 0, K2 = 0;
    const uint32_t max_phred = 1 + rnd(40);
    for (uint32_t q = 0; q < m; ++q) {
        const uint32_t ph = rnd(max_phred + 1), al = rnd(2);
        w[q] = al == 0 ? (int32_t)ph : -(int32_t)ph;
        (al == 0 ? K1 : K2) += ph;
    }
    TileCol tc;
    std::memset(&tc, 0, sizeof tc);
    tc.l_in = l_in; tc.n_new = 1; tc.d = 1; tc.l_out = l_out; tc.kind = 0; tc.g = 0;
    tc.dropmask = 1; tc.K0 = TILE_KINF; tc.K12 = K1 + K2; tc.K2 = (int32_t)K2; tc.dpos[0] = 0;
    tc.pad0 = 2; tc.pad1 = (uint8_t)(l_out - 11);
    for (uint32_t q = 0; q < m; ++q) tc.w_local[q] = w[q];
    // previous projection values of the tile (tie-heavy: small range)
    const uint32_t base = 1000000 + rnd(1000), range = 1 + rnd(rnd(2) ? 6 : 3000);
    std::vector<uint32_t> S32(1u << l_in), out32(1u << l_out), bp32((1u << l_out) / 32, 0);
    for (auto &v : S32) v = base + rnd(range);
    {
        int32_t TW[32], T5[32];
        for (uint32_t i = 0; i < 32; ++i) {
            TW[i] = tile_fast_warp_entry(tc, 0, i);
            T5[i] = tile_fast_lane_entry(tc, i);
        }
        constexpr int LG32 = LG16 + 1;
        for (uint32_t tid = 0; tid < 1024; ++tid) {
            RecordEmit emit{bp32.data() + (tid >> 5) * (1u << LG32), tid & 31u, bp32.data(), tid, 0};
            column_fast<LG32, false, true>(tc, TW, T5, cg, S32.data(), out32.data(), emit, tid);
        }
    }
    // the rotated u16 tile: X first, the ending read second, the others in order
    auto others_of = [&](uint32_t x, uint32_t bits) {  // canonical index without bit 0 and bit pX, compacted
        uint32_t o = 0, k = 0;
        for (uint32_t q = 1; q < bits; ++q)
            if (q != pX) o |= ((x >> q) & 1u) << k++;
        return o;
    };
    std::vector<uint32_t> Win(1u << (l_in - 1), 0), Wout(1u << (l_out - 1), 0), tbits(1024, 0);
    for (uint32_t x = 0; x < (1u << l_in); ++x) {
        const uint32_t xr = ((x >> pX) & 1u) | ((x & 1u) << 1) | (others_of(x, l_in) << 2);
        const uint32_t v = S32[x] - base + 7;  // tile-relative, a small positive offset
        Win[xr >> 1] |= v << (16 * (xr & 1u));
    }
    TileCol16 t16;
    std::memset(&t16, 0, sizeof t16);
    auto both = [](int32_t v) { return ((uint32_t)v & 0xFFFFu) * 0x00010001u; };
    t16.k12x2 = both((int32_t)(K1 + K2)); t16.k2x2 = both((int32_t)K2);
    t16.wp2 = both(w[0]); t16.nwp2 = both(-w[0]);
    t16.wn2 = both(w[l_in]); t16.nwn2 = both(-w[l_in]);
    t16.wx_hi = ((uint32_t)w[pX] & 0xFFFFu) << 16; t16.nwx_hi = ((uint32_t)(-w[pX]) & 0xFFFFu) << 16;
    t16.l_out = l_out;
    {
        uint32_t k = 0;
        for (uint32_t q = 1; q < l_in; ++q)
            if (q != pX) { t16.w2[k] = both(w[q]); t16.nw2[k] = both(-w[q]); ++k; }
    }
    uint32_t TW2[32], T52[32];
    for (uint32_t i = 0; i < 32; ++i) {
        uint32_t a = 0, b = 0;
        for (uint32_t bit = 0; bit < 5; ++bit) {
            if ((i >> bit) & 1u) {
                a = whmec_vadd2_host(a, t16.w2[5 + LG16 + bit]);
                b = whmec_vadd2_host(b, t16.w2[bit]);
            }
        }
        TW2[i] = a;
        T52[i] = b;
    }
    for (uint32_t tid = 0; tid < 1024; ++tid)
        column_fast16<LG16>(t16, TW2, T52, cg, Win.data(), Wout.data(), StoreEmit{&tbits[tid]}, tid);
    // compare every output
    const uint32_t halfq = 1u << (l_out - 2), N = 4 * IT;
    for (uint32_t o = 0; o < (1u << l_out); ++o) {  // canonical output index: cell index without bit 0
        const uint32_t cell = o << 1;
        const uint32_t X = (cell >> pX) & 1u, nw = (cell >> l_in) & 1u, oth = others_of(cell & ((1u << l_in) - 1u), l_in);
        const uint32_t q = oth | (nw ? halfq : 0), qm = oth;
        const uint32_t got = (Wout[q] >> (16 * X)) & 0xFFFFu;
        if (got != out32[o] - base + 7) return -(long)(1 + o);
        // the bit the backtrace would read (tile_u16_bit_index) against the ballot layout's bit of the same output
        const uint32_t idx = tile_u16_bit_index(l_out, pX - 1, o);
        const uint32_t r16 = (tbits[idx / N] >> (idx % N)) & 1u;
        const uint32_t r32 = (bp32[o >> 5] >> (o & 31u)) & 1u;
        if (r16 != r32) return -(long)(1000000 + o);
        (void)qm;
    }
    return (long)(1u << l_out);
}

}  // namespace

extern "C" long whemul_fast16_column_check(uint32_t seed, uint32_t lg16, uint32_t cg, uint32_t pX) {
    switch (lg16) {
        case 0: return fast16_check<0>(seed, cg, pX);
        case 1: return fast16_check<1>(seed, cg, pX);
        default: return fast16_check<2>(seed, cg, pX);
    }
}
