#include <cuda_runtime.h>
#include "../../whatshap_b200/csrc/tile_fast.h"
using namespace whmec;
struct BallotEmit { uint32_t *bp; __device__ __forceinline__ void operator()(uint32_t w, bool b) const { bp[w] = __ballot_sync(0xFFFFFFFFu, b); } __device__ __forceinline__ void store(uint32_t) const {} };
template <int BITS> struct PackedEmit { uint32_t *bpw; uint32_t tid; __device__ __forceinline__ void operator()(uint32_t, bool) const {}
  __device__ __forceinline__ void store(uint32_t bits) const { if (BITS == 8) reinterpret_cast<uint8_t *>(bpw)[tid] = (uint8_t)bits; else reinterpret_cast<uint16_t *>(bpw)[tid] = (uint16_t)bits; } };
#ifndef VARIANT
#define VARIANT 0
#endif
__global__ void __launch_bounds__(1024, 1) probe(const TileCol *tcp, const int32_t *TW, const int32_t *T5, uint32_t cg, uint32_t *bpw) {
    extern __shared__ uint32_t sm[];
    __shared__ TileCol tc; if (threadIdx.x == 0) tc = *tcp; __syncthreads();
    const uint32_t tid = threadIdx.x;
#if VARIANT == 0
    column_fast<3, false, true>(tc, TW, T5, cg, sm, sm + 16384, BallotEmit{bpw + (tid >> 5) * 8}, tid);
#else
    column_fast<3, false, true, true>(tc, TW, T5, cg, sm, sm + 16384, PackedEmit<16>{bpw, tid}, tid);
#endif
}
