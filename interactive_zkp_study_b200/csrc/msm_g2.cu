// G2 instantiation of the MSM / table entry points (include/zkp_b200.h).
#include "msm_api.cuh"

using namespace zkp;
using Api = GroupApi<Fp2>;

// zkp_table_download (msm_g1.cu) needs the G2 download path
template struct zkp::GroupApi<Fp2>;

extern "C" {

int zkp_g2_msm(const uint8_t* pts, const uint8_t* scalars, uint64_t n, uint8_t out_xy[128], int* out_is_inf) {
  return Api::msm_host(pts, scalars, n, out_xy, out_is_inf);
}
int zkp_g2_table_load(const uint8_t* pts, uint64_t n, uint64_t* handle) { return Api::table_load(pts, n, handle); }
int zkp_g2_msm_table(uint64_t table, uint64_t offset, const uint8_t* scalars, uint64_t n, uint8_t out_xy[128],
                     int* out_is_inf) {
  return Api::msm_table(table, offset, scalars, n, out_xy, out_is_inf);
}
int zkp_g2_msm_dev(uint64_t table, uint64_t offset, uint64_t scalars, uint64_t sc_offset, uint64_t n,
                   uint8_t out_xy[128], int* out_is_inf) {
  return Api::msm_dev(table, offset, scalars, sc_offset, n, out_xy, out_is_inf, false);
}
int zkp_g2_fixed_base_mul_dev(const uint8_t base_xy[128], uint64_t scalars, uint64_t n, uint64_t* out_table) {
  return Api::fixed_base_dev(base_xy, scalars, n, out_table);
}
int zkp_g2_table_precompute(uint64_t table, int window_bits) { return Api::table_precompute(table, window_bits); }
int zkp_g2_fixed_base_mul(const uint8_t base_xy[128], const uint8_t* scalars, uint64_t n, uint64_t* out_table) {
  return Api::fixed_base_host(base_xy, scalars, n, out_table);
}

}  // extern "C"
