// Entry points declared in include/starcop_b200.h whose kernels are not written yet.
// They fail loudly (SC_ERR_UNSUPPORTED); nothing falls back to another implementation.
#include "common.cuh"
extern "C" int64_t sc_mag1c_workspace_bytes(int, int, int, int) { return -1; }
extern "C" int sc_mag1c_filter(const void*, int64_t, const void*, const uint8_t*, void*, void*, int, int, int, int,
                               double, int, void*, void*) { return SC_ERR_UNSUPPORTED; }
extern "C" int64_t sc_ratio_workspace_bytes(int, int64_t) { return -1; }
extern "C" int sc_ratio_product(const float*, const float*, float*, int, int64_t, float, float, void*, void*) { return SC_ERR_UNSUPPORTED; }
