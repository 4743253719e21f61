// k_fused_warp_units.cu -- instantiations of k2u_f32 (one warp per muscle, pairs of units; per-instance tables).
#include "mb_launchers.cuh"
#include "mb_fused_warp.cuh"
namespace mb {
#define MB_FORCES(FIB, CEN, NPU_) \
    (forces ? launch_fused(k2u_f32<FIB, CEN, NPU_, true>, xa, g, st) : launch_fused(k2u_f32<FIB, CEN, NPU_, false>, xa, g, st))
#define MB_NPU(FIB, CEN) \
    (npu == 4 ? MB_FORCES(FIB, CEN, 4) : (npu == 3 ? MB_FORCES(FIB, CEN, 3) : (npu == 2 ? MB_FORCES(FIB, CEN, 2) : MB_FORCES(FIB, CEN, 1))))
#define MB_CEN(FIB) (cen == 0 ? MB_NPU(FIB, 0) : (cen == 1 ? MB_NPU(FIB, 1) : MB_NPU(FIB, 2)))
// g.ns = the longest muscle of the batch (uniform: n); pairs per lane = ceil(g.ns / 64) <= 4
int launch_k2u_f32(const K2xArgs& xa, const FusedGeom& g, int fib, int cen, bool forces, cudaStream_t st) {
    const int npu = (g.ns + 63) / 64;
    return fib == MB_FIBERS_POTVIN ? MB_CEN(MB_FIBERS_POTVIN) : MB_CEN(MB_FIBERS_PYMUSCLE);
}
}  // namespace mb
