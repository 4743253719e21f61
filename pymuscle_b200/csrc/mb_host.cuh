// mb_host.cuh -- host-side plumbing shared by the translation units of libmuscle_b200.so:
// error reporting, cached tuning knobs, device queries, the fused kernels' launch geometry and
// the launcher entry points each kernel family's .cu file exports to the C-ABI file
// (muscle_b200.cu).  One family per translation unit keeps a full rebuild at the time of the
// slowest family instead of the sum of all of them (`make -j`).
#pragma once
#include "../../include/muscle_b200.h"
#include "mb_core.cuh"

#include <cmath>
#include <cstdint>

namespace mb {

// ---- errors (thread-local message behind mb_last_error) -------------------------------------
int fail(int code, const char* fmt, ...);

#define MB_CUDA_OK(expr)                                                                            \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess) return ::mb::fail(MB_ECUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

// ---- tuning / debugging knobs: environment variables, read ONCE per process -------------------
struct Knobs {
    int k1_generic;      // MB_K1_GENERIC=1     mb_step never uses the streaming kernels
    int k1_stages;       // MB_K1_STAGES        streaming kernel: pipeline depth (0 = default)
    int k1_w;            // MB_K1_W             streaming kernel: tile width cap (0 = default)
    int k1_bps;          // MB_K1_BPS           streaming kernel: cap on resident blocks per SM (0 = none)
    int k1_flat;         // MB_K1_FLAT          -1 auto, 0 never, 1 whenever eligible: flat-tiled kernel for ragged rows
    int k1_flat_group;   // MB_K1_FLAT_GROUP    flat-tiled kernel: target units per group of muscles (0 = 16 tiles)
    int fused_block;     // MB_FUSED_BLOCK=1    mb_step_fused: block kernel where it would pick a warp kernel
    int fused_no_bulk;   // MB_FUSED_NO_BULK=1  block kernel stages excitations with plain loads
    int fused_general;   // MB_FUSED_GENERAL=1  fused kernels keep every clamp of the reference
    int f64_block;       // MB_F64_BLOCK=1      FP64 mode: block kernel where it would pick the warp kernel
    int fused_persist;   // MB_FUSED_PERSIST    -1 auto (windows of <= 24 steps), 0 never, 1 always: persistent short-window warp kernel
    int fused_units;     // MB_FUSED_UNITS      FP32 mode, 129..256 units, shared table: warp-per-muscle kernel instead of the block
                         //                     kernel: -1 auto (without central fatigue), 0 never, 1 always
    int fused_wide;      // MB_FUSED_WIDE=1     FP32 mode, n > 128: block-per-muscle kernel where it would pick the block kernel
};
const Knobs& knobs();

int current_device();
int sm_count();

// ---- per-launch scalars derived from MbConfig (host, float64) ---------------------------------------
constexpr double kLog2e = 1.4426950408889634074;
inline double lin_scale() { return 1.0 - std::exp(-2.0 * (0.4 * 0.4 * 0.4)); }   // fibers:241

inline ScalF scal_f(const MbConfig& c) {
    ScalF s;
    s.gain_e = (float)(c.firing_gain * c.input_scale);
    s.argmin = (float)(-c.max_duration * kLog2e / c.adaptation_time_constant);
    s.c25k = (float)(lin_scale() / 0.4 / kKappa);
    s.ythr = (float)(0.4 * kKappa);
    s.inv_out = (float)(1.0 / c.output_divisor);
    return s;
}
inline ScalD scal_d(const MbConfig& c) {
    ScalD s;
    s.input_scale = c.input_scale;
    s.min_rate = c.min_firing_rate;
    s.gain = c.firing_gain;
    s.mitau = -1.0 / c.adaptation_time_constant;
    s.max_dur = c.max_duration;
    s.c25 = lin_scale() / 0.4;
    s.out_div = c.output_divisor;
    return s;
}

// ---- launch geometry of the fused kernels ---------------------------------------------------------
struct FusedGeom {
    int U, q, G, Gp, T, ns, npg, ptp, threads, smem;
    int64_t blocks;
    int kernel;          // FusedKernel
};
enum FusedKernel {
    FK_BLOCK = 0,        // k2_fused_f32 / k2_fused_f64: block per muscle groups, thread = unit
    FK_WARP_PAIR = 1,    // k2w_f32: one warp per muscle pair (12..128 units)
    FK_WARP_UNITS = 2,   // k2u_f32: one warp per muscle, pairs of units (per-instance tables)
    FK_WARP_F64 = 3,     // k2w_f64: one warp per muscle, FP64 mode (n <= 128)
    FK_WIDE = 4,         // k2x_f32: one block per muscle, any n (wide muscles, ragged rows)
    FK_WARP_PAIR_PERSIST = 5   // k2w_f32<PERSIST>: short windows -- resident warps walk the pairs, state prefetched by cp.async
};

}  // namespace mb
