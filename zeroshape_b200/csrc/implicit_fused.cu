// Fused tcgen05 implicit-decoder kernel (K1).  Placeholder entry points until the kernel lands:
// they report ZS_ERR_UNSUPPORTED so that callers fail loudly instead of silently falling back.
#include "common.cuh"

extern "C" size_t zs_implicit_packed_bytes(void) { return 0; }
extern "C" int zs_implicit_pack(const ZsImplicitWeights*, void*, void*) {
  zs::set_error("zs_implicit_pack: fused kernel not built in this revision");
  return ZS_ERR_UNSUPPORTED;
}
extern "C" size_t zs_implicit_kv_bytes(int, int) { return 0; }
extern "C" int zs_implicit_kv_pack(const float*, const float*, const float*, const float*, int, int, void*, void*) {
  zs::set_error("zs_implicit_kv_pack: fused kernel not built in this revision");
  return ZS_ERR_UNSUPPORTED;
}
extern "C" int zs_implicit_fused_fwd(const void*, const void*, int, const float*, int, int64_t, int, float, float, int,
                                     int, float*, int, int, void*) {
  zs::set_error("zs_implicit_fused_fwd: fused kernel not built in this revision");
  return ZS_ERR_UNSUPPORTED;
}
