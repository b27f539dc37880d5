// PROBE ONLY: instruction count of the packed 16-bit steady-state column body (column_fast16 in csrc/tile_fast.h — the
// building block that tests/emul holds to column_fast bit for bit).  16 outputs per thread and column, as
// the cfg3 variant of column_fast.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -cubin scripts/probes/column_fast16_probe.cu; cuobjdump -sass
#include <cuda_runtime.h>
#include "../../whatshap_b200/csrc/tile_fast.h"
using namespace whmec;
struct Store16 {
    uint32_t *bpw;
    uint32_t tid;
    __device__ __forceinline__ void operator()(uint32_t, bool) const {}
    __device__ __forceinline__ void store(uint32_t bits) const { reinterpret_cast<uint16_t *>(bpw)[tid] = (uint16_t)bits; }
};
__global__ void __launch_bounds__(1024, 1) probe16(const __grid_constant__ TileCol16 c, const uint32_t *TW2, const uint32_t *T52, uint32_t cg, uint32_t *bpw) {
    extern __shared__ __align__(16) uint32_t sm[];
    column_fast16<2>(c, TW2, T52, cg, sm, sm + 8192, Store16{bpw, threadIdx.x}, threadIdx.x);
}
