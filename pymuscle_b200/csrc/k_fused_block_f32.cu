// k_fused_block_f32.cu -- FP32-mode instantiations of the fused block kernel (k2_fused_f32, mb_fused.cuh).
#include "mb_launchers.cuh"
namespace mb {
#define MB_FORCES(FIB, CEN, UU) \
    (forces ? launch_fused(k2_fused_f32<FIB, CEN, UU, true>, a, g, st) : launch_fused(k2_fused_f32<FIB, CEN, UU, false>, a, g, st))
#define MB_CEN(FIB, UU) (cen == 0 ? MB_FORCES(FIB, 0, UU) : (cen == 1 ? MB_FORCES(FIB, 1, UU) : MB_FORCES(FIB, 2, UU)))
#define MB_FIB(UU) (fib == MB_FIBERS_POTVIN ? MB_CEN(MB_FIBERS_POTVIN, UU) : MB_CEN(MB_FIBERS_PYMUSCLE, UU))
// four instances per thread: only the shapes that fit 64 registers (no decimated force output, no PyMuscle fibers
// with central fatigue); fused_geometry() never picks U = 4 for the others
#define MB_CEN4(FIB) (cen == 0 ? launch_fused(k2_fused_f32<FIB, 0, 4, false>, a, g, st) \
                               : (cen == 1 ? launch_fused(k2_fused_f32<FIB, 1, 4, false>, a, g, st) \
                                           : launch_fused(k2_fused_f32<FIB, 2, 4, false>, a, g, st)))
int launch_k2_block_f32(const K2Args& a, const FusedGeom& g, int fib, int cen, bool forces, cudaStream_t st) {
    if (g.U == 4) {
        if (forces || (fib == MB_FIBERS_PYMUSCLE && cen != 0))
            return fail(MB_EINVAL, "fused block kernel: 4 instances per thread are not built for this configuration");
        return fib == MB_FIBERS_POTVIN ? MB_CEN4(MB_FIBERS_POTVIN) : launch_fused(k2_fused_f32<MB_FIBERS_PYMUSCLE, 0, 4, false>, a, g, st);
    }
    if (g.U == 2) return MB_FIB(2);
    return MB_FIB(1);
}
}  // namespace mb
