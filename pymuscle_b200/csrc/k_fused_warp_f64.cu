// k_fused_warp_f64.cu -- instantiations of k2w_f64 (FP64 mode, one warp per muscle; mb_fused_warp64.cuh).
#include "mb_launchers.cuh"
#include "mb_fused_warp64.cuh"
namespace mb {
#define MB_FORCES(FIB, CEN, UPL_) \
    (forces ? launch_fused(k2w_f64<FIB, CEN, UPL_, true>, a, g, st) : launch_fused(k2w_f64<FIB, CEN, UPL_, false>, a, g, st))
#define MB_UPL(FIB, CEN) \
    (upl == 4 ? MB_FORCES(FIB, CEN, 4) : (upl == 3 ? MB_FORCES(FIB, CEN, 3) : (upl == 2 ? MB_FORCES(FIB, CEN, 2) : MB_FORCES(FIB, CEN, 1))))
#define MB_CEN(FIB) (cen == 0 ? MB_UPL(FIB, 0) : (cen == 1 ? MB_UPL(FIB, 1) : MB_UPL(FIB, 2)))
int launch_k2w_f64(const K2Args& a, const FusedGeom& g, int fib, int cen, bool forces, cudaStream_t st) {
    const int upl = (a.n + 31) / 32;
    return fib == MB_FIBERS_POTVIN ? MB_CEN(MB_FIBERS_POTVIN) : MB_CEN(MB_FIBERS_PYMUSCLE);
}
}  // namespace mb
