// placeholder until the tcgen05 kernel lands
#include "pstb_common.cuh"
using namespace pstb;
extern "C" int64_t pstb_kernel_workspace_bytes(int64_t, int64_t) { return 16; }
extern "C" int pstb_snp_kernel(const uint8_t*, int64_t, int64_t, int64_t, pstb_axis, pstb_axis, int, int, double, double, int, double*, float*, int, int, void*, int64_t, int64_t, void*) { return fail("pstb_snp_kernel: not built"); }
extern "C" int pstb_syrk_planes(const void*, const void*, int64_t, int64_t, int64_t, float*, int64_t, int, float, void*) { return fail("not built"); }
extern "C" int pstb_mirror_lower(float*, int64_t, int64_t, void*) { return fail("not built"); }
extern "C" int pstb_convert_kernel(const float*, int64_t, void*, int, double, void*) { return fail("not built"); }
