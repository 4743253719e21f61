// k_stream_impl.cuh -- launcher of the streaming single-step kernel (k1_stream, mb_stream.cuh) for one
// precision mode; k_stream_f32.cu / k_stream_f64.cu instantiate it.
#pragma once
#include "mb_launchers.cuh"

namespace mb {

template <typename T, bool CENTRAL, bool RECOVERY, bool FORCES, bool FATIGUE, bool PERROW = false>
static int launch_k1s_variant(const K1sArgs<T>& a, int threads, int smem, int64_t items, cudaStream_t st) {
    auto kern = k1_stream<T, CENTRAL, RECOVERY, FORCES, FATIGUE, PERROW>;
    // persistent grid: exactly the blocks that are resident at once, each walking the work list.
    // The opt-in and the occupancy query are cached per (instantiation, geometry): a batch-of-1 drop-in
    // step is only ~20 us long, so two driver calls per launch would show.
    // The function attribute is per device, so the device is part of the key.
    static thread_local int c_dev = -1, c_threads = -1, c_smem = -1, c_per_sm = 0;
    const int dev = current_device();
    if (c_dev != dev || c_threads != threads || c_smem != smem) {
        // (the 48 KB default limit covers static + dynamic shared memory; the kernel has a few hundred static bytes)
        if (smem > 40 * 1024) MB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int q = 0;
        MB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, kern, threads, smem));
        if (q < 1) return fail(MB_ECUDA, "k1_stream does not fit on an SM (threads %d, smem %d)", threads, smem);
        c_dev = dev; c_threads = threads; c_smem = smem; c_per_sm = q;
    }
    int per_sm = c_per_sm;
    const int cap = knobs().k1_bps;
    if (cap > 0 && per_sm > cap) per_sm = cap;
    int64_t blocks = (int64_t)sm_count() * per_sm;
    if (blocks > items) blocks = items;
    kern<<<(unsigned)blocks, threads, smem, st>>>(a);
    MB_CUDA_OK(cudaGetLastError());
    return MB_OK;
}

template <typename T>
static int launch_k1s(MB_K1S_PARAMS, const typename StreamTraits<T>::S& sc) {
    K1sArgs<T> a;
    a.params = (const typename StreamTraits<T>::P*)params;
    a.param_stride = c.param_rows > 1 ? (int64_t)StreamTraits<T>::G * L : 0;
    a.rows = rows; a.L = L; a.S = S; a.seg_off = seg_off; a.seg_div = seg_div;
    a.peripheral = c.peripheral_fatigue != 0;
    a.in = (const T*)in; a.in_per_unit = in_per_unit; a.cpf = cpf; a.dur = dur; a.forces = (T*)forces; a.totals = (T*)totals;
    a.out_gain = (const T*)ep.out_gain; a.fatigue = ep.fatigue; a.fat_div = c.output_divisor;
    a.hill_rest = (const T*)ep.hill_rest; a.hill_cur = (const T*)ep.hill_cur; a.hill_prev = (const T*)ep.hill_prev;
    a.hill_width = ep.hill_width; a.hill_opt = ep.hill_opt; a.hill_vmax = ep.hill_vmax; a.hill_ecc = ep.hill_ecc;
    a.dt = dt;
    a.adk_d = -kLog2e / c.adaptation_time_constant;
    a.max_dur = c.max_duration;
    a.sc = sc;
    double* const fatigue = ep.fatigue;
    const bool central = c.central_fatigue != 0, recovery = c.fibers_kind == MB_FIBERS_PYMUSCLE;
    const int64_t mean_n = (L + S - 1) / S;
    int W = (int)((mean_n + 31) / 32) * 32;
    // per-instance tables put R parameter tiles into every stage: narrower tiles keep two stages within reach
    const int wcap = central ? StreamTraits<T>::kMaxWCentral : StreamTraits<T>::kMaxW;   // the instantiation's launch bound
    int wmax = knobs().k1_w > 0 ? knobs().k1_w : (c.param_rows > 1 ? 64 : wcap);
    if (wmax > wcap) wmax = wcap;
    if (W > wmax) W = wmax;
    if (W < 32) W = 32;
    const int stage_bytes = kStreamRows * (sizeof(T) == 4 ? wcap + 2 : W + 2) * 8 * (central ? 2 : 1) +                             // state rows (compile-time pitch)
                            StreamTraits<T>::G * W * 16 * (c.param_rows > 1 ? kStreamRows : 1) +         // + unit parameters
                            (in_per_unit ? kStreamRows * (W + 2 * (16 / (int)sizeof(T))) * (int)sizeof(T) : 0);   // + per-unit inputs
    int NS = knobs().k1_stages;
    if (NS <= 0) NS = StreamTraits<T>::kDefStages;            // measured: FP32 mode: double buffering x 2 blocks/SM saturates HBM, deeper rings
                                                              // only cost occupancy; FP64 mode (3-4 narrower blocks per SM): three stages
    if (NS > 8) NS = 8;
    if (NS < 2) NS = 2;
    a.W = W; a.NS = NS;
    const int threads = W + 32;
    a.seg_in_smem = (seg_off != nullptr && S <= 2048) ? 1 : 0;
    const int smem = 128 + NS * stage_bytes + 2 * kStreamRows * 8 * (8 + (int)sizeof(T)) + (a.seg_in_smem ? S * 8 + (S + 1) * 4 + 8 : 0);
    const int64_t items = ((rows + kStreamRows - 1) / kStreamRows) * S;
#define MB_K1S_F(C_, R_, F_) (fatigue ? launch_k1s_variant<T, C_, R_, F_, true>(a, threads, smem, items, st) \
                                       : launch_k1s_variant<T, C_, R_, F_, false>(a, threads, smem, items, st))
#define MB_K1S(C_, R_) (forces ? MB_K1S_F(C_, R_, true) : MB_K1S_F(C_, R_, false))
    if (c.param_rows > 1) {                                   // per-instance tables: forces on, no fatigue epilogue
        if (central) return recovery ? launch_k1s_variant<T, true, true, true, false, true>(a, threads, smem, items, st)
                                     : launch_k1s_variant<T, true, false, true, false, true>(a, threads, smem, items, st);
        return recovery ? launch_k1s_variant<T, false, true, true, false, true>(a, threads, smem, items, st)
                        : launch_k1s_variant<T, false, false, true, false, true>(a, threads, smem, items, st);
    }
    if (central) return recovery ? MB_K1S(true, true) : MB_K1S(true, false);
    return recovery ? MB_K1S(false, true) : MB_K1S(false, false);
#undef MB_K1S_F
#undef MB_K1S
}

}  // namespace mb
