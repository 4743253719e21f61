// mb_launchers.cuh -- what each kernel family's translation unit exports to the C-ABI file.
#pragma once
#include "mb_host.cuh"
#include "mb_fused.cuh"      // K2Args
#include "mb_stream.cuh"     // K1sArgs
#include "mb_fused_wide.cuh"  // K2xArgs

namespace mb {

// ---- k_stream_f32.cu / k_stream_f64.cu: the streaming single-step kernel (k1_stream) ------------
#define MB_K1S_PARAMS                                                                                              \
    const MbConfig &c, const void *params, int64_t rows, int64_t L, int32_t S, const int32_t *seg_off,            \
        const double *seg_div, const void *in, int32_t in_per_unit, double *cpf, double *dur, void *forces,       \
        void *totals, const K1sEpilogue &ep, double dt, cudaStream_t st
int launch_k1s_f32(MB_K1S_PARAMS);
int launch_k1s_f64(MB_K1S_PARAMS);
// ---- k_flat_f32.cu / k_flat_f64.cu: the flat-tiled variant for rows of several muscles (k1_flat) ------
int launch_k1f_f32(MB_K1S_PARAMS);
int launch_k1f_f64(MB_K1S_PARAMS);

// ---- k_fused_*.cu: the fused K-step kernels.  fib = MbFibers, cen = 0 / 1 / 2 (see CoreF32) ------
int launch_k2_block_f32(const K2Args& a, const FusedGeom& g, int fib, int cen, bool forces, cudaStream_t st);
int launch_k2_block_f64(const K2Args& a, const FusedGeom& g, int fib, int cen, bool forces, cudaStream_t st);
int launch_k2w_f32(const K2Args& a, const FusedGeom& g, int fib, int cen, bool forces, cudaStream_t st);
int launch_k2u_f32(const K2xArgs& xa, const FusedGeom& g, int fib, int cen, bool forces, cudaStream_t st);
int launch_k2w_f64(const K2Args& a, const FusedGeom& g, int fib, int cen, bool forces, cudaStream_t st);
int launch_k2x_f32(const K2xArgs& xa, const FusedGeom& g, int fib, int cen, bool forces, cudaStream_t st);

// launch with the opt-in to large dynamic shared memory (cheap; static bytes count towards the 48 KB default too)
template <typename KernelT, typename ArgsT>
static int launch_fused(KernelT kernel, const ArgsT& a, const FusedGeom& g, cudaStream_t st) {
    if (g.smem > 40 * 1024)
        MB_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem));
    kernel<<<(unsigned)g.blocks, g.threads, g.smem, st>>>(a);
    MB_CUDA_OK(cudaGetLastError());
    return MB_OK;
}

}  // namespace mb
