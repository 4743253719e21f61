// k_flat_impl.cuh -- launcher of the flat-tiled single-step kernel for ragged rows (k1_flat, mb_flat.cuh).
#pragma once
#include "mb_launchers.cuh"
#include "mb_flat.cuh"

namespace mb {

template <typename T, bool CENTRAL, bool RECOVERY, bool FORCES, bool FATIGUE>
static int launch_k1f_variant(const K1fArgs<T>& fa, int threads, int smem, int64_t max_items, cudaStream_t st) {
    auto kern = k1_flat<T, CENTRAL, RECOVERY, FORCES, FATIGUE>;
    static thread_local int c_dev = -1, c_threads = -1, c_smem = -1, c_per_sm = 0;
    const int dev = current_device();
    if (c_dev != dev || c_threads != threads || c_smem != smem) {
        if (smem > 40 * 1024) MB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int q = 0;
        MB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, kern, threads, smem));
        if (q < 1) return fail(MB_ECUDA, "k1_flat does not fit on an SM (threads %d, smem %d)", threads, smem);
        c_dev = dev; c_threads = threads; c_smem = smem; c_per_sm = q;
    }
    int per_sm = c_per_sm;
    const int cap = knobs().k1_bps;
    if (cap > 0 && per_sm > cap) per_sm = cap;
    int64_t blocks = (int64_t)sm_count() * per_sm;
    if (blocks > max_items) blocks = max_items;
    kern<<<(unsigned)blocks, threads, smem, st>>>(fa);
    MB_CUDA_OK(cudaGetLastError());
    return MB_OK;
}

template <typename T>
static int launch_k1f(MB_K1S_PARAMS, const typename StreamTraits<T>::S& sc) {
    K1fArgs<T> fa;
    K1sArgs<T>& a = fa.s;
    a.params = (const typename StreamTraits<T>::P*)params;
    a.param_stride = 0;
    a.rows = rows; a.L = L; a.S = S; a.seg_off = seg_off; a.seg_div = seg_div;
    a.peripheral = c.peripheral_fatigue != 0;
    a.in = (const T*)in; a.in_per_unit = 0; a.cpf = cpf; a.dur = dur; a.forces = (T*)forces; a.totals = (T*)totals;
    a.out_gain = (const T*)ep.out_gain; a.fatigue = ep.fatigue; a.fat_div = c.output_divisor;
    a.hill_rest = (const T*)ep.hill_rest; a.hill_cur = (const T*)ep.hill_cur; a.hill_prev = (const T*)ep.hill_prev;
    a.hill_width = ep.hill_width; a.hill_opt = ep.hill_opt; a.hill_vmax = ep.hill_vmax; a.hill_ecc = ep.hill_ecc;
    a.dt = dt;
    a.adk_d = -kLog2e / c.adaptation_time_constant;
    a.max_dur = c.max_duration;
    a.sc = sc;
    a.seg_in_smem = 1;
    double* const fatigue = ep.fatigue;
    const bool central = c.central_fatigue != 0, recovery = c.fibers_kind == MB_FIBERS_PYMUSCLE;
    const int wcap = central ? StreamTraits<T>::kMaxWCentral : StreamTraits<T>::kMaxW;
    int W = (int)((L + 31) / 32) * 32;                        // a row shorter than the widest tile: one tile per row chunk
    if (knobs().k1_w > 0 && W > knobs().k1_w) W = knobs().k1_w;
    if (W > wcap) W = wcap;
    if (W < 32) W = 32;
    int NS = knobs().k1_stages;
    if (NS <= 0) NS = StreamTraits<T>::kDefStages;
    if (NS > 8) NS = 8;
    if (NS < 2) NS = 2;
    a.W = W; a.NS = NS;
    fa.group_units = knobs().k1_flat_group > 0 ? knobs().k1_flat_group : 16 * W;
    const int threads = W + 32, ncw = W / 32;
    const int stage_bytes = kStreamRows * (sizeof(T) == 4 ? wcap + 2 : W + 2) * 8 * (central ? 2 : 1) + StreamTraits<T>::G * W * 16;   // compile-time row pitch
    const int smem = 128 + NS * stage_bytes + NS * kStreamRows * kFlatMaxSegs * (int)sizeof(T) +
                     2 * ncw * kFlatMaxSegs * kStreamRows * (fatigue ? 2 : 1) * (int)sizeof(T) + 8 +
                     S * 8 + (S + 1) * 4 + (kFlatMaxGroups + 1) * 4 + 16;
    const int64_t max_items = ((rows + kStreamRows - 1) / kStreamRows) * (int64_t)S;   // (an upper bound: groups <= muscles)
#define MB_K1F_F(C_, R_, F_) (fatigue ? launch_k1f_variant<T, C_, R_, F_, true>(fa, threads, smem, max_items, st) \
                                       : launch_k1f_variant<T, C_, R_, F_, false>(fa, threads, smem, max_items, st))
#define MB_K1F(C_, R_) (forces ? MB_K1F_F(C_, R_, true) : MB_K1F_F(C_, R_, false))
    if (central) return recovery ? MB_K1F(true, true) : MB_K1F(true, false);
    return recovery ? MB_K1F(false, true) : MB_K1F(false, false);
#undef MB_K1F_F
#undef MB_K1F
}

}  // namespace mb
