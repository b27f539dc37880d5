// PROBE ONLY (not part of the library): instruction count of a packed 16-bit steady-state column body.
// Layout assumed: local bit 0 = a read that outlives the panel (X), local bit 1 = the read that ends in this column.
// One 32-bit word of the u16 state = the SAME candidate of two outputs (X on side 0 / 1); the next word = the other
// candidate of the same two outputs: one LDS.64 per output pair.  Values are tile-relative and below 2^15.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -cubin scripts/probes/column_fast16_probe.cu; cuobjdump -sass
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int LG = 3, IT = 1 << LG;  // 8 output pairs = 16 outputs per thread, as the cfg3 variant of column_fast

struct Col16 {
    uint32_t k12x2;      // K12 in both halves
    uint32_t wp2;        // weight of the ending read in both halves
    uint32_t nwp2;       // its negation (mod 2^16) in both halves
    uint32_t wx_hi;      // weight of X in the high half (output B = A + X)
    uint32_t w2[8];      // packed weights (w, w) of the output bits that vary inside a thread
    uint32_t nw2[8];     // their negations
    uint32_t l_out;
};

__device__ __forceinline__ uint32_t sign_bits(uint32_t bits, uint32_t v1, uint32_t v0, uint32_t par2) {
    // per half: bit 15 of ((v1 | 0x8000) - v0 - par) is set iff v1 >= v0 + par (no borrow between the halves below 2^15);
    // par == 0: "v1 >= v0" = candidate 1 does not win; par == 1: "v1 > v0" = candidate 0 wins -- either way the stored bit
    // (pick1 ^ par) is this bit XOR a per-thread constant
    const uint32_t d = (v1 | 0x80008000u) - v0 - par2;
    bits = __funnelshift_l(d, bits, 1);        // high half's bit
    return __funnelshift_l(d << 16, bits, 1);  // low half's bit
}

__global__ void __launch_bounds__(1024, 1) probe16(const __grid_constant__ Col16 c, const uint32_t *TW2, const uint32_t *T52, uint32_t cg, uint32_t *bpw) {
    extern __shared__ __align__(16) uint32_t sm[];  // u16 state: words of two entries
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t pbase = warp * (IT * 32u) + lane;  // output pair index
    const uint2 *sin2 = reinterpret_cast<const uint2 *>(sm) + pbase;
    uint32_t *so = sm + 8192 + pbase;                 // next column's words
    const uint32_t par0 = (__popc(pbase) + (cg & 1u)) & 1u;
    uint32_t ue[IT], nue[IT];
    ue[0] = __vadd2(TW2[warp], T52[lane]) + c.wx_hi;  // (E_A, E_A + w_X)
    nue[0] = c.k12x2 - ue[0];                         // per half below 2^15: no borrow
#pragma unroll
    for (int it = 1; it < IT; ++it) {
        const int q = (it & 1) ? 0 : (it & 2) ? 1 : 2;
        ue[it] = __vadd2(ue[it & (it - 1)], c.w2[q]);
        nue[it] = __vadd2(nue[it & (it - 1)], c.nw2[q]);
    }
    uint32_t bits = 0;
#pragma unroll
    for (int it = 0; it < IT; ++it) {
        const uint2 s = sin2[it * 32];
        const uint32_t par = par0 ^ (uint32_t)(((it >> 0) ^ (it >> 1) ^ (it >> 2)) & 1);
        const uint32_t c0 = __vminu2(ue[it], nue[it]);
        const uint32_t c1 = __viaddmin_u16x2(ue[it], c.wp2, __vadd2(nue[it], c.nwp2));
        const uint32_t v0 = __vadd2(c0, s.x), v1 = __vadd2(c1, s.y);
        so[it * 32] = __vminu2(v0, v1);
        bits = sign_bits(bits, v1, v0, par * 0x00010001u);
    }
    // stored bit = pick1 ^ par = !(second >= first) ^ ... : one XOR with a per-thread constant mask
    reinterpret_cast<uint16_t *>(bpw)[tid] = (uint16_t)(~bits ^ (par0 ? 0x6996u : 0x9669u));
}
