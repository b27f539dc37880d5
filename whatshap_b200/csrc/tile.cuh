// Tile kernel path (T == 1): declared here, defined in tile.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "pack.h"

namespace whmec {

bool tile_path_eligible(const Packed &pk);

struct TilePlan {
    uint64_t backptr_bytes = 0;
    uint64_t state_bytes = 0;
    uint32_t launches = 0;
    void *impl = nullptr;
    int create(const Packed &pk, cudaStream_t stream, uint64_t &h2d_bytes, std::string &msg);
    int sweep(const Packed &pk, cudaStream_t stream, std::string &msg);
    int backtrace(const Packed &pk, cudaStream_t stream, uint32_t *d_path_index, uint32_t *d_result, std::string &msg);
    void release();
};

}  // namespace whmec
