// mb_fused_wide.cuh -- K2x: the fused K-step kernel for WIDE muscles and ragged rows, FP32 mode.
//
// The warp kernels keep a muscle inside one warp (<= 128 units) and the block kernel inside 480
// threads; StandardMuscle makes far wider muscles first-class (muscle.py:144-147, :189-243: 527 N gives
// 2,000 units) and an arm model is a row of 30 muscles of 256..3,127 units.  Here ONE BLOCK owns one
// muscle -- (row, segment) of the ragged layout -- for the whole window:
//
//   * thread t holds the unit pairs (t + 2jT, t + (2j+1)T), j < NPU (T = blockDim.x): the two units of a
//     pair are the two halves of the packed FFMA2 / FMUL2 / FADD2 operands (the K2u organisation), so a
//     thread carries NPU independent chains and consecutive lanes touch consecutive state elements;
//   * per step a thread adds its pair forces and stores one float into its warp's private tile FT[32][34];
//     after 32 steps lane s of every warp sums row s, the per-warp partials of the 32 steps meet in a
//     double-buffered [32][warps] array, ONE block barrier per 32 steps, and one warp (rotating) adds the
//     warps' partials in a fixed order and writes the muscle's 32 totals;
//   * pairs beyond the muscle's length are skipped by block-uniform predicates, so a short muscle of a
//     ragged row costs only the pairs it has;
//   * the peak twitch forces (only needed when the compensated capacity pair is renormalised, once per 32
//     steps) are re-read from the L2-resident table instead of living in registers.
//
// HBM traffic: state in/out once per window + 8 bytes per (step, muscle) of excitation and total.
#pragma once
#include "mb_fused.cuh"
#include <type_traits>

namespace mb {

constexpr int kXTile = 32;         // steps per tile
constexpr int kXPitch = 34;        // row pitch of a warp's float force tile (even: 64-bit row reads, conflict free)


template <int FIB, int CENTRAL, int NPU, bool FORCES, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? 2 : 1) k2x_f32(const K2xArgs xa) {
    const K2Args& a = xa.k;
    constexpr int NW = THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* ftw = (float*)smem_raw + warp * (kXTile * kXPitch + 2 * kXTile);      // [32][34] this warp's forces
    float* excw = ftw + kXTile * kXPitch;                                         // [2][32] this warp's excitations
    float* pt = (float*)smem_raw + NW * (kXTile * kXPitch + 2 * kXTile);          // [2][32][NW + 1] warp partials

    const int64_t m = blockIdx.x;                          // muscle = (row, segment)
    const int64_t row = m / xa.S;
    const int seg = (int)(m - row * xa.S);
    const int off0 = xa.seg_off ? xa.seg_off[seg] : 0;
    const int n = xa.seg_off ? xa.seg_off[seg + 1] - off0 : (int)xa.L;
    const int64_t base = row * xa.L + off0;                // first state element of this muscle
    const float inv_out = xa.seg_div ? (float)(1.0 / xa.seg_div[seg]) : a.sf.inv_out;
    const float4* table = (const float4*)a.params + (a.param_stride ? m * a.param_stride : 0);
    const int64_t tabL = xa.L;                             // table row length ([group][L])
    const int npa = (n + 2 * THREADS - 1) / (2 * THREADS); // pairs per thread this muscle needs (block-uniform)
    MB_ASSERT(npa <= NPU && m < a.M && seg < xa.S && base + n <= xa.total_units && off0 + n <= xa.L);
    const int ntiles = (a.K + kXTile - 1) / kXTile;
    const float* exc = (const float*)a.exc;
    const PairScal sc = pair_scal(a);

    PairParams<float2> pp[NPU];
    PairState st[NPU];
#pragma unroll
    for (int i = 0; i < NPU; ++i) {
        if (i >= npa) continue;
        const int u0 = tid + 2 * i * THREADS, u1 = u0 + THREADS;
        const UnitF a0 = load_unit(table, tabL, off0 + (u0 < n ? u0 : 0));
        const UnitF a1 = load_unit(table, tabL, off0 + (u1 < n ? u1 : 0));
        const PairParams<float> q0 = pair_params(a0, a), q1 = pair_params(a1, a);
        const float inf = __int_as_float(0x7f800000);
        pp[i].crit = make_float2(u0 < n ? q0.crit : inf, u1 < n ? q1.crit : inf);   // idle slot: never recruited, force 0
        pp[i].thr_g = make_float2(q0.thr_g, q1.thr_g);   pp[i].peak = make_float2(q0.peak, q1.peak);
        pp[i].phir = make_float2(q0.phir, q1.phir);      pp[i].phir6 = make_float2(q0.phir6, q1.phir6);
        pp[i].ctA = make_float2(q0.ctA, q1.ctA);         pp[i].ctB = make_float2(q0.ctB, q1.ctB);
        pp[i].fatdt = make_float2(q0.fatdt, q1.fatdt);   pp[i].rdidt = make_float2(q0.rdidt, q1.rdidt);
        pp[i].ptf = make_float2(q0.ptf, q1.ptf);         pp[i].ptf_lo = make_float2(q0.ptf_lo, q1.ptf_lo);
        pair_init(st[i], pp[i], 0, u0 < n ? a.cpf[base + u0] : unit_peak(a0),
                  (CENTRAL && u0 < n) ? a.dur[base + u0] : 0.0, a.adk_d);
        pair_init(st[i], pp[i], 1, u1 < n ? a.cpf[base + u1] : unit_peak(a1),
                  (CENTRAL && u1 < n) ? a.dur[base + u1] : 0.0, a.adk_d);
    }

    // excitation of (tile, step = lane); 0 beyond the window.  exc_stride_k == 0: constant
    auto fetch = [&](int tile) -> float {
        const int k = tile * kXTile + lane;
        return (k < a.K) ? __ldg(exc + (int64_t)(k / a.exc_hold) * a.exc_stride_k + m) : 0.0f;
    };
    float e_next = fetch(0);

    for (int tile = 0; tile < ntiles; ++tile) {
        const int k0 = tile * kXTile;
        const int steps = min(kXTile, a.K - k0);
        float* es = excw + (tile & 1) * kXTile;
        es[lane] = e_next;
        __syncwarp();
        if (tile + 1 < ntiles) e_next = fetch(tile + 1);
#pragma unroll
        for (int i = 0; i < NPU; ++i)
            if (i < npa) pair_begin_tile<CENTRAL>(st[i], sc);

        int fphase = FORCES ? k0 % a.forces_every : 0;        // steps since the last captured one
        auto do_step = [&](int s, auto last_tag) {
            constexpr bool LAST = decltype(last_tag)::value;
            const float e = es[s];
            float2 fsum = make_float2(0.0f, 0.0f);
            bool write_now = false;
            if (FORCES) {                                     // every forces_every-th step, counted (a per-step modulo is ~20 instructions)
                write_now = ++fphase == a.forces_every;
                if (write_now) fphase = 0;
            }
#pragma unroll
            for (int i = 0; i < NPU; ++i) {
                if (i >= npa) continue;
                bool oa, ob;
                const float2 F = pair_step<FIB, CENTRAL>(st[i], pp[i], sc, make_float2(e, e), oa, ob);
                fsum = (i == 0) ? F : __fadd2_rn(fsum, F);
                if (FORCES || LAST) {
                    const int u0 = tid + 2 * i * THREADS, u1 = u0 + THREADS;
                    if (FORCES && write_now) {
                        const int64_t fb = (int64_t)((k0 + s) / a.forces_every) * xa.total_units + base;   // [record][rows][L]
                        float* fd = (float*)a.forces_dec;
                        if (u0 < n) fd[fb + u0] = F.x;
                        if (u1 < n) fd[fb + u1] = F.y;
                        if (a.mask_dec) {
                            if (u0 < n) a.mask_dec[fb + u0] = oa ? 1 : 0;
                            if (u1 < n) a.mask_dec[fb + u1] = ob ? 1 : 0;
                        }
                    }
                    if (LAST) {
                        float* fl = (float*)a.forces_last + base;
                        if (u0 < n) fl[u0] = F.x;
                        if (u1 < n) fl[u1] = F.y;
                    }
                }
            }
            ftw[s * kXPitch + lane] = fsum.x + fsum.y;
        };
        const bool has_last = a.forces_last != nullptr && tile == ntiles - 1;
        const int fast_steps = has_last ? steps - 1 : steps;
        for (int s = 0; s < fast_steps; ++s) do_step(s, std::false_type{});
        if (has_last) do_step(steps - 1, std::true_type{});
#pragma unroll
        for (int i = 0; i < NPU; ++i) {
            if (i >= npa) continue;
            // the peaks are only needed here, once per tile: re-read them (L2) rather than keep 4 registers per pair
            const int u0 = tid + 2 * i * THREADS, u1 = u0 + THREADS;
            const float4 c0 = __ldg(table + 2 * tabL + off0 + (u0 < n ? u0 : 0));
            const float4 c1 = __ldg(table + 2 * tabL + off0 + (u1 < n ? u1 : 0));
            pp[i].ptf = make_float2(c0.x, c1.x);
            pp[i].ptf_lo = make_float2(c0.w, c1.w);
            pair_renormalise<FIB>(st[i], pp[i]);
        }
        __syncwarp();
        // ---- stage 1: lane s sums this warp's 32 lanes of step k0 + s ---------------------------------------
        float* ptb = pt + (tile & 1) * (kXTile * (NW + 1));
        {
            const float2* rp = (const float2*)(ftw + lane * kXPitch);
            float2 acc0 = rp[0], acc1 = rp[1];
#pragma unroll
            for (int c = 2; c < 16; c += 2) {
                acc0 = __fadd2_rn(acc0, rp[c]);
                acc1 = __fadd2_rn(acc1, rp[c + 1]);
            }
            const float2 acc = __fadd2_rn(acc0, acc1);
            ptb[lane * (NW + 1) + warp] = acc.x + acc.y;
        }
        __syncthreads();                                    // the tile's partials are complete (one barrier per 32 steps:
                                                            // pt is double-buffered, see the header)
        // ---- stage 2: one warp (rotating) adds the warps' partials in a fixed order ----------------------------
        if (warp == tile % NW) {
            const int kl = k0 + lane;
            if (lane < steps && a.totals && (a.totals_every == 1 || (kl + 1) % a.totals_every == 0)) {
                const float* q = ptb + lane * (NW + 1);
                float v = q[0];
#pragma unroll
                for (int w = 1; w < NW; ++w) v += q[w];
                v *= inv_out;
                const int64_t krow = kl / a.totals_every;
                ((float*)a.totals)[krow * a.totals_stride_k + m] = v;
                if (a.n_peers > 0) {
                    const int64_t op = (m >> 1) * a.peer_stride_pair + krow * a.peer_stride_k + (m & 1) * a.peer_stride_odd;
                    for (int pr = 0; pr < a.n_peers; ++pr) peer_store((float*)a.totals_peer[pr] + op, v, a.peer_multicast);
                }
            }
        }
    }

#pragma unroll
    for (int i = 0; i < NPU; ++i) {
        if (i >= npa) continue;
        const int u0 = tid + 2 * i * THREADS, u1 = u0 + THREADS;
        if (u0 < n) pair_store<CENTRAL != 0>(st[i], 0, a, base + u0);
        if (u1 < n) pair_store<CENTRAL != 0>(st[i], 1, a, base + u1);
    }
}

}  // namespace mb
