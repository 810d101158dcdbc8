// force_v3.cu -- instantiates the pair kernels of arithmetic variant 3 (force_kernels.cuh: 0 = x86 order and the scalar laws,
// 1 = fused, 2 = fused + culling, 3 = fused with rsqrt(s^3), 4 = that + culling).
#include "force_kernels.cuh"

namespace haccsr {
template int launch_force<6, 0, true, 3>(haccsr_ctx *, const ForceParams &, int, bool);
template int launch_force<6, 0, false, 3>(haccsr_ctx *, const ForceParams &, int, bool);
template int launch_force<7, 0, true, 3>(haccsr_ctx *, const ForceParams &, int, bool);
template int launch_force<7, 0, false, 3>(haccsr_ctx *, const ForceParams &, int, bool);
}  // namespace haccsr
