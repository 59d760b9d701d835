// G1 instantiation of the MSM / table entry points (include/zkp_b200.h).
#include "msm_api.cuh"

namespace zkp {
int g_force_window_bits = 0;
}
using namespace zkp;
using Api = GroupApi<Fp>;

extern "C" {

int zkp_msm_set_window_bits(int c) {
  if (c != 0 && (c < 2 || c > MSM_MAX_C)) {
    set_last_error("zkp_msm_set_window_bits: c must be 0 or in [2,20]");
    return ZKP_ERR_INVALID_ARGUMENT;
  }
  g_force_window_bits = c;
  return ZKP_OK;
}

int zkp_g1_msm(const uint8_t* pts, const uint8_t* scalars, uint64_t n, uint8_t out_xy[64], int* out_is_inf) {
  return Api::msm_host(pts, scalars, n, out_xy, out_is_inf);
}
int zkp_g1_table_load(const uint8_t* pts, uint64_t n, uint64_t* handle) { return Api::table_load(pts, n, handle); }
int zkp_g1_msm_table(uint64_t table, uint64_t offset, const uint8_t* scalars, uint64_t n, uint8_t out_xy[64],
                     int* out_is_inf) {
  return Api::msm_table(table, offset, scalars, n, out_xy, out_is_inf);
}
int zkp_g1_msm_dev(uint64_t table, uint64_t offset, uint64_t scalars, uint64_t sc_offset, uint64_t n,
                   uint8_t out_xy[64], int* out_is_inf) {
  return Api::msm_dev(table, offset, scalars, sc_offset, n, out_xy, out_is_inf, false);
}
int zkp_g1_msm_dev_partial(uint64_t table, uint64_t offset, uint64_t scalars, uint64_t sc_offset, uint64_t n,
                           uint8_t out_xyzz[128]) {
  return Api::msm_dev(table, offset, scalars, sc_offset, n, out_xyzz, nullptr, true);
}
int zkp_g1_combine_partials(const uint8_t* partials, uint32_t count, uint8_t out_xy[64], int* out_is_inf) {
  return Api::combine(partials, count, out_xy, out_is_inf);
}
int zkp_g1_fixed_base_mul(const uint8_t base_xy[64], const uint8_t* scalars, uint64_t n, uint64_t* out_table) {
  return Api::fixed_base_host(base_xy, scalars, n, out_table);
}
int zkp_g1_fixed_base_mul_dev(const uint8_t base_xy[64], uint64_t scalars, uint64_t n, uint64_t* out_table) {
  return Api::fixed_base_dev(base_xy, scalars, n, out_table);
}

int zkp_g1_msm_dev_batch(uint64_t table, uint32_t count, const uint64_t* scalars, const uint64_t* sc_offsets,
                         const uint64_t* offsets, const uint64_t* lens, uint8_t* out_xy, int* out_is_inf) {
  return Api::msm_batch(table, count, scalars, sc_offsets, offsets, lens, out_xy, out_is_inf);
}
int zkp_g1_table_precompute(uint64_t table, int window_bits) { return Api::table_precompute(table, window_bits); }

// G1 or G2 table -> canonical affine points on the host
int zkp_table_download(uint64_t table, uint64_t offset, uint64_t n, uint8_t* out_pts) {
  return guarded([&](Context& c) {
    if (n && !out_pts) throw InvalidArgument("zkp_table_download: null output");
    if (Resource* t = registry().get(table, HandleKind::G1Table)) {
      GroupApi<Fp>::download(c, t, offset, n, out_pts);
    } else if (Resource* t2 = registry().get(table, HandleKind::G2Table)) {
      GroupApi<Fp2>::download(c, t2, offset, n, out_pts);
    } else {
      throw BadHandle("zkp_table_download: not a point table");
    }
  });
}

// window width of a precomputed table (0 = plain layout)
int zkp_table_window_bits(uint64_t table, int* window_bits) {
  return guarded([&](Context&) {
    if (!window_bits) throw InvalidArgument("zkp_table_window_bits: null output");
    Resource* t = registry().get(table, HandleKind::G1Table);
    if (!t) t = registry().get(table, HandleKind::G2Table);
    if (!t) throw BadHandle("zkp_table_window_bits: not a point table");
    *window_bits = t->pre_c;
  });
}

}  // extern "C"
