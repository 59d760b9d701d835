// placeholder until the polynomial ops land (next commit)
#include "registry.cuh"
using namespace zkp;
static int nyi(const char* f) { set_last_error(std::string(f) + ": not implemented yet"); return ZKP_ERR_INVALID_ARGUMENT; }
extern "C" {
int zkp_fr_vec_op(int, const uint8_t*, const uint8_t*, uint64_t, uint8_t*) { return nyi("zkp_fr_vec_op"); }
int zkp_fr_batch_inverse(const uint8_t*, uint64_t, uint8_t*) { return nyi("zkp_fr_batch_inverse"); }
int zkp_fr_poly_eval(const uint8_t*, uint64_t, const uint8_t*, uint8_t*) { return nyi("zkp_fr_poly_eval"); }
int zkp_groth16_quotient(const uint8_t*, const uint8_t*, const uint8_t*, uint64_t, const uint8_t*, uint64_t, uint8_t*, uint8_t*) { return nyi("zkp_groth16_quotient"); }
int zkp_fr_poly_mul(const uint8_t*, uint64_t, const uint8_t*, uint64_t, uint8_t*) { return nyi("zkp_fr_poly_mul"); }
int zkp_fr_poly_divmod(const uint8_t*, uint64_t, const uint8_t*, uint64_t, uint8_t*, uint8_t*) { return nyi("zkp_fr_poly_divmod"); }
}
