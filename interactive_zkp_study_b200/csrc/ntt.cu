// placeholder until the NTT lands (next commit)
#include "registry.cuh"
using namespace zkp;
static int nyi(const char* f) { set_last_error(std::string(f) + ": not implemented yet"); return ZKP_ERR_INVALID_ARGUMENT; }
extern "C" {
int zkp_fr_ntt(uint8_t*, uint32_t, const uint8_t*, int, const uint8_t*) { return nyi("zkp_fr_ntt"); }
int zkp_fr_ntt_dev(uint64_t, uint64_t, uint32_t, const uint8_t*, int, const uint8_t*) { return nyi("zkp_fr_ntt_dev"); }
}
