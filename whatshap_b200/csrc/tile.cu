#include "tile.cuh"
#include "../../include/whmec.h"
namespace whmec {
bool tile_path_eligible(const Packed &) { return false; }
int TilePlan::create(const Packed &, cudaStream_t, uint64_t &, std::string &msg) { msg = "tile path not built"; return WHMEC_ERR_UNSUPPORTED; }
int TilePlan::sweep(const Packed &, cudaStream_t, std::string &msg) { msg = "tile path not built"; return WHMEC_ERR_UNSUPPORTED; }
int TilePlan::backtrace(const Packed &, cudaStream_t, uint32_t *, uint32_t *, std::string &msg) { msg = "tile path not built"; return WHMEC_ERR_UNSUPPORTED; }
void TilePlan::release() {}
}
