// k_fused_warp_units.cu -- instantiations of k2u_f32 (one warp per muscle, pairs of units; per-instance tables).
#include "mb_launchers.cuh"
#include "mb_fused_warp.cuh"
namespace mb {
#define MB_FORCES(FIB, CEN, NPU_) \
    (forces ? launch_fused(k2u_f32<FIB, CEN, NPU_, true>, a, g, st) : launch_fused(k2u_f32<FIB, CEN, NPU_, false>, a, g, st))
#define MB_NPU(FIB, CEN) (npu == 2 ? MB_FORCES(FIB, CEN, 2) : MB_FORCES(FIB, CEN, 1))
#define MB_CEN(FIB) (cen == 0 ? MB_NPU(FIB, 0) : (cen == 1 ? MB_NPU(FIB, 1) : MB_NPU(FIB, 2)))
int launch_k2u_f32(const K2Args& a, const FusedGeom& g, int fib, int cen, bool forces, cudaStream_t st) {
    const int npu = (a.n + 63) / 64;
    return fib == MB_FIBERS_POTVIN ? MB_CEN(MB_FIBERS_POTVIN) : MB_CEN(MB_FIBERS_PYMUSCLE);
}
}  // namespace mb
