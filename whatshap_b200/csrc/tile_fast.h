// Steady-state column of the tile kernel (see tile.cu: fast_kind), as host/device code: the CUDA kernel emits the
// back-pointer bits as warp ballots, the test-only emulation (tests/emul) records them lane by lane.
#pragma once
#include "tile_device.h"

namespace whmec {

#if defined(__CUDACC__)
using TilePair = uint2;
#define WHMEC_UMIN(a, b) min(a, b)
#define WHMEC_POPC(x) popc32(x)
#else
struct alignas(8) TilePair {
    uint32_t x, y;
};
#define WHMEC_UMIN(a, b) ((a) < (b) ? (a) : (b))
#define WHMEC_POPC(x) popc32(x)
#endif

// The two 32-entry tables of a fast column: per warp  K2 + E(global reads of this tile) + output bits 5+LG .. 9+LG,
// per lane the weights of output bits 0..4  (output bit q <-> local bit q + 1: local bit 0 is the read that ends).
WHMEC_HD int32_t tile_fast_warp_entry(const TileCol &tc, uint32_t tile, uint32_t warp) {
    const uint32_t lg = tc.pad1;
    int32_t s = tc.K2;
    for (uint32_t b = 0; b < tc.g; ++b)
        if ((tile >> b) & 1u) s += tc.w_global[b];
#pragma unroll
    for (uint32_t q = 0; q < 5; ++q)
        if ((warp >> q) & 1u) s += tc.w_local[6 + lg + q];
    return s;
}

WHMEC_HD int32_t tile_fast_lane_entry(const TileCol &tc, uint32_t lane) {
    int32_t s = 0;
#pragma unroll
    for (uint32_t q = 0; q < 5; ++q)
        if ((lane >> q) & 1u) s += tc.w_local[q + 1];
    return s;
}

// (bits << 1) | (t >> 31): one SHF.L.W on the device
WHMEC_HD uint32_t tile_shift_in_sign(uint32_t bits, uint32_t t) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(t, bits, 1);
#else
    return (bits << 1) | (t >> 31);
#endif
}

// plain expressions (no recursion, no calls) so that they fold to constants once the loops are unrolled
#define cx_ctz(x) (((x) & 1) ? 0 : ((x) & 2) ? 1 : ((x) & 4) ? 2 : ((x) & 8) ? 3 : 4)
#define cx_parity(x) ((((x) >> 0) ^ ((x) >> 1) ^ ((x) >> 2) ^ ((x) >> 3) ^ ((x) >> 4)) & 1)

// Fast path of column_drop1 for dropped bit 0 (see fast_kind).  Warp w owns 32 * 2^LG consecutive
// outputs; every shared-memory address inside the loop is a per-thread base plus a compile-time
// offset, the E() of the 2^LG outputs of a thread are subset sums built with one add each, the two
// candidate cells of an output share one 64-bit load (and with SHARE the two outputs that differ
// only in the newly started read share it too), back-pointers leave as warp ballots.
// `emit(word, bit)`: this thread's bit of back-pointer word `word` of its warp (words are numbered from the warp's first).
// PACKED (TileCol::pad2 == 1, experimental): the thread keeps the bits of its own 2^LG (x 2 with SHARE) outputs in a
// register and hands them to `emit.store(bits)` once per column.  No predicate is formed: with all values below 2^28
// (the tile path's precondition) "candidate 1 wins" is the sign of v1 - v0 - par, and one funnel shift moves that sign
// bit into the register — IADD3 + SHF per output instead of IADD + ISETP + VOTE + STG.  The outputs arrive in the order
// (it = 0, twin of 0, it = 1, ...), so the first one ends up in the highest of the thread's bits (tile_packed_bit_index).
template <int LG, bool HASK0, bool SHARE, bool PACKED = false, class Emit>
WHMEC_HD void column_fast(const TileCol &tc, const int32_t *__restrict__ TW, const int32_t *__restrict__ T5,
                          uint32_t cg, const uint32_t *__restrict__ Sin, uint32_t *__restrict__ Sout, Emit emit, uint32_t tid) {
    constexpr int IT = 1 << LG;
    const uint32_t lane = tid & 31u, warp = tid >> 5;
    const uint32_t obase = warp * (IT * 32u) + lane;      // o = obase + 32*it  (+ nout/2 for the shared twin)
    const uint32_t pmask = (1u << (tc.l_in - 1)) - 1u;    // candidate pairs of the previous projection
    const TilePair *sin2 = reinterpret_cast<const TilePair *>(Sin) + (obase & pmask);
    uint32_t *so = Sout + obase;
    const uint32_t half = 1u << (tc.l_out - 1);
    const uint32_t wp = (uint32_t)tc.w_local[0];
    const uint32_t wn = SHARE ? (uint32_t)tc.w_local[tc.l_out] : 0u;  // the read that starts in this column
    const uint32_t K0 = tc.K0, K12 = tc.K12;
    const uint32_t par0 = (WHMEC_POPC(obase) + (cg & 1u)) & 1u;  // parity of the bits above the dropped one
    uint32_t bits = 0;
    uint32_t ue[IT];
    ue[0] = (uint32_t)(TW[warp] + T5[lane]);
#pragma unroll
    for (int it = 1; it < IT; ++it) ue[it] = ue[it & (it - 1)] + (uint32_t)tc.w_local[6 + cx_ctz(it)];
#pragma unroll
    for (int it = 0; it < IT; ++it) {
        const TilePair s = sin2[it * 32];
        const uint32_t par = par0 ^ (uint32_t)cx_parity(it);
        {
            const uint32_t u0 = ue[it], u1 = u0 + wp;
            uint32_t c0 = WHMEC_UMIN(u0, K12 - u0), c1 = WHMEC_UMIN(u1, K12 - u1);
            if (HASK0) { c0 = WHMEC_UMIN(c0, K0); c1 = WHMEC_UMIN(c1, K0); }
            const uint32_t v0 = c0 + s.x, v1 = c1 + s.y;
            const bool pick1 = v1 < v0 + par;  // par == 0: candidate 0 is visited first and keeps ties
            so[it * 32] = WHMEC_UMIN(v0, v1);
            if (PACKED) bits = tile_shift_in_sign(bits, v1 - v0 - par);  // sign set <=> pick1
            else emit((uint32_t)it, pick1 != (par != 0));
        }
        if (SHARE) {  // twin output: the new read on side 1 (one more bit above the dropped one)
            const uint32_t u0 = ue[it] + wn, u1 = u0 + wp;
            uint32_t c0 = WHMEC_UMIN(u0, K12 - u0), c1 = WHMEC_UMIN(u1, K12 - u1);
            if (HASK0) { c0 = WHMEC_UMIN(c0, K0); c1 = WHMEC_UMIN(c1, K0); }
            const uint32_t v0 = c0 + s.x, v1 = c1 + s.y;
            const uint32_t parb = par ^ 1u;
            const bool pick1 = v1 < v0 + parb;
            so[half + it * 32] = WHMEC_UMIN(v0, v1);
            if (PACKED) bits = tile_shift_in_sign(bits, v1 - v0 - parb);
            else emit((half >> 5) + (uint32_t)it, pick1 != (parb != 0));
        }
    }
    if (PACKED) {
        // stored bit = pick1 ^ par, par = par0 ^ parity(it) ^ twin: fold the compile-time part and the thread's par0 in at once
        constexpr uint32_t N = SHARE ? 2 * IT : IT;
        uint32_t cm = 0;
#pragma unroll
        for (int it = 0; it < IT; ++it) {
            if (SHARE) cm |= ((uint32_t)cx_parity(it) << (N - 1 - 2 * it)) | (((uint32_t)cx_parity(it) ^ 1u) << (N - 2 - 2 * it));
            else cm |= (uint32_t)cx_parity(it) << (N - 1 - it);
        }
        const uint32_t all = N >= 32 ? 0xFFFFFFFFu : ((1u << N) - 1u);
        emit.store((bits ^ cm ^ (par0 ? all : 0u)) & all);
    }
}

}  // namespace whmec

// ---- experimental building block (not used by the kernel yet): the steady-state column on packed 16-bit
// values.  Layout of a u16 tile: local bit 0 = a read X that outlives the panel, bit 1 = the read that ends in this
// column, then the other local reads in canonical order, the read that starts in this column on top.  One 32-bit word
// holds the same candidate of the two outputs that differ in X; the next word the other candidate of the same outputs.
// Values are tile-relative and every value (costs included) stays below 2^15.  Exercised against column_fast by
// tests/emul (whemul_fast16_column_check).
namespace whmec {

struct TileCol16 {
    uint32_t k12x2;    // K1 + K2 in both halves
    uint32_t k2x2;     // K2 (+ E of the tile's global reads) in both halves
    uint32_t wp2;      // weight of the ending read (bit 1) in both halves, and its negation mod 2^16
    uint32_t nwp2;
    uint32_t wn2;      // weight of the starting read (top bit) in both halves, and its negation
    uint32_t nwn2;
    uint32_t wx_hi;    // weight of X in the high half only (output B = output A + X)
    uint32_t nwx_hi;
    uint32_t w2[16];   // (w, w) of pair-index bit q (rotated output bit q + 1)
    uint32_t nw2[16];
    uint32_t l_out;    // log2 outputs (twins included)
};

#if defined(__CUDA_ARCH__)
#define WHMEC_VADD2(a, b) __vadd2(a, b)
#define WHMEC_VMINU2(a, b) __vminu2(a, b)
#else
WHMEC_HD uint32_t whmec_vadd2_host(uint32_t a, uint32_t b) { return ((a + b) & 0xFFFFu) | (((a >> 16) + (b >> 16)) << 16); }
WHMEC_HD uint32_t whmec_vminu2_host(uint32_t a, uint32_t b) {
    const uint32_t lo = (a & 0xFFFFu) < (b & 0xFFFFu) ? (a & 0xFFFFu) : (b & 0xFFFFu);
    const uint32_t hi = (a >> 16) < (b >> 16) ? (a >> 16) : (b >> 16);
    return lo | (hi << 16);
}
#define WHMEC_VADD2(a, b) whmec_vadd2_host(a, b)
#define WHMEC_VMINU2(a, b) whmec_vminu2_host(a, b)
#endif

// Thread `tid` produces the 2^LG pair-iterations x 2 twins x 2 halves outputs of pair indices
// qm = warp * 2^LG * 32 + it * 32 + lane (and qm + 2^(l_out - 2) for the twin).  TW2 / T52: packed (value, value) sums of
// the weights selected by the warp / lane bits of qm.  Returns nothing; bits go to emit.store: for iteration `it` the order
// is (main pair: high half, low half; twin pair: high half, low half), first shifted in = highest.
template <int LG, class Emit>
WHMEC_HD void column_fast16(const TileCol16 &tc, const uint32_t *__restrict__ TW2, const uint32_t *__restrict__ T52, uint32_t cg,
                            const uint32_t *__restrict__ Win, uint32_t *__restrict__ Wout, Emit emit, uint32_t tid) {
    constexpr int IT = 1 << LG;
    const uint32_t lane = tid & 31u, warp = tid >> 5;
    const uint32_t qbase = warp * (IT * 32u) + lane;
    const TilePair *win2 = reinterpret_cast<const TilePair *>(Win) + qbase;
    uint32_t *wo = Wout + qbase;
    const uint32_t halfq = 1u << (tc.l_out - 2);  // pairs without the starting read
    // parity of the other bits of output A (X = 0) above the ending read; output B has one more bit set (X)
    const uint32_t par0 = (WHMEC_POPC(qbase) + (cg & 1u)) & 1u;
    uint32_t ue[IT], nue[IT];
    ue[0] = WHMEC_VADD2(WHMEC_VADD2(tc.k2x2, WHMEC_VADD2(TW2[warp], T52[lane])), tc.wx_hi);
    nue[0] = tc.k12x2 - ue[0];  // both halves are true costs: K1 - E >= 0, no borrow
#pragma unroll
    for (int it = 1; it < IT; ++it) {
        ue[it] = WHMEC_VADD2(ue[it & (it - 1)], tc.w2[5 + cx_ctz(it)]);
        nue[it] = WHMEC_VADD2(nue[it & (it - 1)], tc.nw2[5 + cx_ctz(it)]);
    }
    uint32_t bits = 0;
#pragma unroll
    for (int it = 0; it < IT; ++it) {
        const TilePair s = win2[it * 32];
        const uint32_t parA = par0 ^ (uint32_t)cx_parity(it);
#pragma unroll
        for (int twin = 0; twin < 2; ++twin) {
            const uint32_t u0 = twin ? WHMEC_VADD2(ue[it], tc.wn2) : ue[it];
            const uint32_t n0 = twin ? WHMEC_VADD2(nue[it], tc.nwn2) : nue[it];
            const uint32_t c0 = WHMEC_VMINU2(u0, n0);
            const uint32_t c1 = WHMEC_VMINU2(WHMEC_VADD2(u0, tc.wp2), WHMEC_VADD2(n0, tc.nwp2));
            const uint32_t v0 = WHMEC_VADD2(c0, s.x), v1 = WHMEC_VADD2(c1, s.y);
            wo[twin * halfq + it * 32] = WHMEC_VMINU2(v0, v1);
            // per half: bit 15 of (v1 | 0x8000) - v0 - par is set iff v1 >= v0 + par; the halves have opposite parity (X)
            const uint32_t pa = parA ^ (uint32_t)twin;  // parity of output A of this pair
            const uint32_t d = (v1 | 0x80008000u) - v0 - (pa ? 0x00000001u : 0x00010000u);
            bits = tile_shift_in_sign(bits, d);        // high half (output B)
            bits = tile_shift_in_sign(bits, d << 16);  // low half (output A)
        }
    }
    // the raw bit says "v1 >= v0 + par"; the back-pointer is the rank of the winner in visiting order, pick1 ^ par with
    // pick1 = !raw: flip output A's bit unless its parity is set, output B's bit if it is (B has one more bit set: X)
    constexpr uint32_t N = 4 * IT;
    uint32_t cm = 0;  // par0 == 0
#pragma unroll
    for (int it = 0; it < IT; ++it)
#pragma unroll
        for (int twin = 0; twin < 2; ++twin) {
            const uint32_t pa = (uint32_t)cx_parity(it) ^ (uint32_t)twin;
            const int j = it * 4 + twin * 2;
            cm |= pa << (N - 1 - j);               // output B: raw ^ par_A
            cm |= (pa ^ 1u) << (N - 1 - (j + 1));  // output A: raw ^ 1 ^ par_A
        }
    const uint32_t all = N >= 32 ? 0xFFFFFFFFu : ((1u << N) - 1u);
    emit.store((bits ^ cm ^ (par0 ? all : 0u)) & all);
}

// TileCol::pad2 of a column inside a packed 16-bit panel: bit 8 set, bits 16..23 = position of X among the column's
// canonical cell bits.
WHMEC_HD bool tile_is_u16(const TileCol &tc) { return (tc.pad2 & 0x100u) != 0; }
WHMEC_HD uint32_t tile_u16_xpos(const TileCol &tc) { return (tc.pad2 >> 16) & 0xFFu; }

WHMEC_HD uint32_t tile_both_halves(int32_t v) { return ((uint32_t)v & 0xFFFFu) * 0x00010001u; }

// The packed constants of one column of tile `tile` from its TileCol (canonical cell order: bit 0 ends here, bit l_in
// starts here, X at tile_u16_xpos) -- everything column_fast16 needs except the two 32-entry tables.
WHMEC_HD void tile_col16_from(const TileCol &tc, uint32_t tile, TileCol16 &out) {
    const uint32_t xpos = tile_u16_xpos(tc), l_in = tc.l_in;
    int32_t k2 = tc.K2;
    for (uint32_t b = 0; b < tc.g; ++b)
        if ((tile >> b) & 1u) k2 += tc.w_global[b];
    out.k12x2 = tile_both_halves((int32_t)tc.K12);
    out.k2x2 = tile_both_halves(k2);
    out.wp2 = tile_both_halves(tc.w_local[0]);
    out.nwp2 = tile_both_halves(-tc.w_local[0]);
    out.wn2 = tile_both_halves(tc.w_local[l_in]);
    out.nwn2 = tile_both_halves(-tc.w_local[l_in]);
    out.wx_hi = ((uint32_t)tc.w_local[xpos] & 0xFFFFu) << 16;
    out.nwx_hi = ((uint32_t)(-tc.w_local[xpos]) & 0xFFFFu) << 16;
    out.l_out = tc.l_out;
    uint32_t k = 0;
    for (uint32_t q = 1; q < l_in; ++q)
        if (q != xpos) {
            out.w2[k] = tile_both_halves(tc.w_local[q]);
            out.nw2[k] = tile_both_halves(-tc.w_local[q]);
            ++k;
        }
    for (; k < 16; ++k) out.w2[k] = out.nw2[k] = 0;
}

// The two 32-entry tables of column_fast16: sums of the packed weights selected by the warp / lane bits of the pair index.
WHMEC_HD uint32_t tile_fast16_warp_entry(const TileCol16 &c, uint32_t warp) {
    const uint32_t lg = c.l_out - 12;
    uint32_t s = 0;
    for (uint32_t b = 0; b < 5; ++b)
        if ((warp >> b) & 1u) s = WHMEC_VADD2(s, c.w2[5 + lg + b]);
    return s;
}
WHMEC_HD uint32_t tile_fast16_lane_entry(const TileCol16 &c, uint32_t lane) {
    uint32_t s = 0;
    for (uint32_t b = 0; b < 5; ++b)
        if ((lane >> b) & 1u) s = WHMEC_VADD2(s, c.w2[b]);
    return s;
}

// Rotated index of canonical local index i of `bits` bits with X at canonical position xp: X first, the rest in order.
WHMEC_HD uint32_t tile_u16_rotate(uint32_t i, uint32_t xp) {
    const uint32_t low = i & ((1u << xp) - 1u), high = i >> (xp + 1);
    return ((i >> xp) & 1u) | (low << 1) | (high << (xp + 1));
}

}  // namespace whmec
