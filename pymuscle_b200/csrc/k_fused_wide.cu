// k_fused_wide.cu -- instantiations of k2x_f32 (one block per muscle: wide muscles and ragged rows; mb_fused_wide.cuh).
#include "mb_launchers.cuh"
#include "mb_fused_wide.cuh"
namespace mb {
template <typename KernelT>
static int launch_wide(KernelT kernel, const K2xArgs& xa, const FusedGeom& g, cudaStream_t st) {
    if (g.smem > 40 * 1024)
        MB_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem));
    kernel<<<(unsigned)g.blocks, g.threads, g.smem, st>>>(xa);
    MB_CUDA_OK(cudaGetLastError());
    return MB_OK;
}
#define MB_FORCES(FIB, CEN, NPU_, T_) \
    (forces ? launch_wide(k2x_f32<FIB, CEN, NPU_, true, T_>, xa, g, st) : launch_wide(k2x_f32<FIB, CEN, NPU_, false, T_>, xa, g, st))
#define MB_NPU(FIB, CEN, T_) \
    (npu == 4 ? MB_FORCES(FIB, CEN, 4, T_) : (npu == 3 ? MB_FORCES(FIB, CEN, 3, T_) : (npu == 2 ? MB_FORCES(FIB, CEN, 2, T_) : MB_FORCES(FIB, CEN, 1, T_))))
#define MB_CEN(FIB, T_) (cen == 0 ? MB_NPU(FIB, 0, T_) : (cen == 1 ? MB_NPU(FIB, 1, T_) : MB_NPU(FIB, 2, T_)))
#define MB_FIB(T_) (fib == MB_FIBERS_POTVIN ? MB_CEN(MB_FIBERS_POTVIN, T_) : MB_CEN(MB_FIBERS_PYMUSCLE, T_))
// g.threads = 256 or 512; g.U = pairs per thread (1..4)
int launch_k2x_f32(const K2xArgs& xa, const FusedGeom& g, int fib, int cen, bool forces, cudaStream_t st) {
    const int npu = g.U;
    if (g.threads == 256) return MB_FIB(256);
    return MB_FIB(512);
}
}  // namespace mb
