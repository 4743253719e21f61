// k_fused_warp_pair.cu -- instantiations of k2w_f32 (one warp per muscle pair, mb_fused_warp.cuh).
#include "mb_launchers.cuh"
#include "mb_fused_warp.cuh"
namespace mb {
#define MB_FORCES(FIB, CEN, UPL_) \
    (forces ? launch_fused(k2w_f32<FIB, CEN, UPL_, true>, a, g, st) : launch_fused(k2w_f32<FIB, CEN, UPL_, false>, a, g, st))
#define MB_UPL(FIB, CEN) \
    (upl == 4 ? MB_FORCES(FIB, CEN, 4) : (upl == 3 ? MB_FORCES(FIB, CEN, 3) : (upl == 2 ? MB_FORCES(FIB, CEN, 2) : MB_FORCES(FIB, CEN, 1))))
#define MB_CEN(FIB) (cen == 0 ? MB_UPL(FIB, 0) : (cen == 1 ? MB_UPL(FIB, 1) : MB_UPL(FIB, 2)))
#define MB_P_UPL(FIB, CEN) \
    (upl == 4 ? launch_fused(k2w_f32<FIB, CEN, 4, false, true>, a, g, st) \
              : (upl == 3 ? launch_fused(k2w_f32<FIB, CEN, 3, false, true>, a, g, st) \
                          : (upl == 2 ? launch_fused(k2w_f32<FIB, CEN, 2, false, true>, a, g, st) \
                                      : launch_fused(k2w_f32<FIB, CEN, 1, false, true>, a, g, st))))
#define MB_P_CEN(FIB) (cen == 0 ? MB_P_UPL(FIB, 0) : (cen == 1 ? MB_P_UPL(FIB, 1) : MB_P_UPL(FIB, 2)))
// g.kernel == FK_WARP_PAIR_PERSIST: the short-window variant (persistent warps, prefetched state; no force capture)
int launch_k2w_f32(const K2Args& a, const FusedGeom& g, int fib, int cen, bool forces, cudaStream_t st) {
    const int upl = (a.n + 31) / 32;
    if (g.kernel == FK_WARP_PAIR_PERSIST) return fib == MB_FIBERS_POTVIN ? MB_P_CEN(MB_FIBERS_POTVIN) : MB_P_CEN(MB_FIBERS_PYMUSCLE);
    return fib == MB_FIBERS_POTVIN ? MB_CEN(MB_FIBERS_POTVIN) : MB_CEN(MB_FIBERS_PYMUSCLE);
}
}  // namespace mb
