// mb_fused_warp64.cuh -- K2w64: the fused K-step kernel of FP64 mode for muscles of up to 128 units.
//
// One WARP owns one muscle for the whole window and never talks to another warp (the FP32 kernel K2w's
// organisation): lane l holds units l, l+32, l+64, l+96 (UPL = ceil(n/32) of them), each with its own
// parameter set and state in registers, which gives a thread UPL independent FP64 dependency chains.
// Compared with the block kernel (k2_fused_f64: thread = one unit of four muscles, a muscle spread over
// 3.75 warps) this removes
//   * the cross-warp reduction: no partial-sum ring, no mbarriers -- ncu showed 12 % of that kernel's warp
//     samples asleep on the ring and 16 % of its instructions in per-tile code;
//   * per-instance work that is per-muscle work here: the excitation is read (one broadcast LDS.64) and
//     scaled (muscle.py:275) once per step and lane instead of once per instance.
// Per step a thread adds its UPL unit forces and stores one double into the warp's private tile
// FT[32][33]; after 32 steps (one __syncwarp) lane s sums row s and writes the muscle's total of step
// k0 + s.  Lane s prefetches the excitation of step k0 + s of the NEXT tile meanwhile.
//
// Arithmetic: CoreF64's (mb_fused.cuh) -- reference operation order for the recruitment mask, G = 1 -
// exp(-dur/tau) by recurrence with an integer on-step counter (fast path) or the reference's literal
// clamps (general path), the table-based exp_neg_fast for the force sigmoid.  The chains of a thread advance in
// three phases (rates, exponentials, forces + fatigue) with ONE branch around the exponentials, so the
// scheduler can interleave them; groups of 32 units without a recruited unit are skipped (see do_step).
#pragma once
#include "mb_fused.cuh"
#include <type_traits>

namespace mb {

#ifndef MB_W64_SKIP
#define MB_W64_SKIP 1              // 0: every group every step; 1: skip muscles entirely at rest; 2: skip every group at rest
                                   // (measured: 1 wins -- the per-step switch of 2 costs more than the groups it skips)
#endif
#ifndef MB_W64_UNROLL
#define MB_W64_UNROLL 2
#endif
constexpr int kW64Unroll = MB_W64_UNROLL;   // unroll factor of the step loop
constexpr int kW64Tile = 32;       // steps per tile
constexpr int kW64Pitch = 33;      // row pitch of the force tile in doubles (odd => conflict free)
constexpr int kW64Warps = 8;       // warps (muscles) per block

// The launcher routes here only when firing_gain == 1 (the reference's default, pool:57): r0 * 1.0 is r0 bit for
// bit, so the multiply of pool:173 is dropped; other gains run the block kernel.
template <int FIB, int CENTRAL, int UPL, bool FORCES>
__global__ void __launch_bounds__(kW64Warps * 32, (FIB == MB_FIBERS_PYMUSCLE && CENTRAL != 0 && UPL == 4) ? 1 : 2)
k2w_f64(const K2Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double exp2_tab[32];                        // exp_neg_fast's table
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 32) exp2_tab[threadIdx.x] = kExp2Tab[threadIdx.x];
    __syncthreads();                                       // (the only block-wide barrier: before any warp may leave)
    double* ftw = (double*)smem_raw + warp * (kW64Tile * kW64Pitch + 2 * kW64Tile);   // [32][33]
    double* excw = ftw + kW64Tile * kW64Pitch;                                          // [2][32]

    const int n = a.n;
    const int64_t m = (int64_t)blockIdx.x * kW64Warps + warp;
    if (m >= a.M) return;                                  // warps are independent: no barrier below
    MB_ASSERT(n <= 32 * UPL);
    const int ntiles = (a.K + kW64Tile - 1) / kW64Tile;
    const double* exc = (const double*)a.exc;
    const double2* table = (const double2*)a.params + m * a.param_stride;        // (per-instance tables: this muscle's own)

    // per-launch scalars
    const ScalD sc = a.sd;
    const double dt = a.dt, max_dur = a.max_dur, decay = a.decay_d;
    const double omd = -expm1(a.dt * sc.mitau);            // 1 - decay
    const double Gmax = CENTRAL ? -expm1(a.max_dur * sc.mitau) : 0.0;
    const double c25s = sc.c25 / kCbrt2, xlin = kLinThrD * kCbrt2;

    // per-unit parameters and state
    double thr[UPL], peak[UPL], phir[UPL], phir6[UPL], ctA2[UPL], ctB2[UPL], fatdt[UPL], rdidt[UPL], ptf[UPL];
    double cpf[UPL], G[UPL];
    double dur[CENTRAL == 2 ? UPL : 1];                     // general path only: the fast path counts on-steps instead
    int cnt[UPL];
#pragma unroll
    for (int i = 0; i < UPL; ++i) {
        const int u = lane + 32 * i;
        const bool valid = u < n;
        const UnitD p = load_unit(table, n, valid ? u : 0);
        thr[i] = valid ? p.thr : __longlong_as_double(0x7ff0000000000000LL);      // idle slot: never recruited, force 0
        peak[i] = p.peak; phir[i] = p.phir; phir6[i] = p.phir6;
        ctA2[i] = p.ctA * kCbrt2; ctB2[i] = p.ctB * kCbrt2;
        fatdt[i] = a.peripheral ? p.fat * a.dt : 0.0;
        rdidt[i] = a.peripheral ? p.rdi * a.dt : 0.0;
        ptf[i] = p.ptf;
        cpf[i] = valid ? a.cpf[m * n + u] : p.ptf;
        const double d0 = (CENTRAL && valid) ? a.dur[m * n + u] : 0.0;
        if (CENTRAL == 2) dur[i] = d0;
        G[i] = CENTRAL ? -expm1(d0 * sc.mitau) : 0.0;                            // pool:217-218
        cnt[i] = 0;
    }

    auto fetch = [&](int tile) -> double {
        const int k = tile * kW64Tile + lane;
        return (k < a.K) ? __ldg(exc + (int64_t)(k / a.exc_hold) * a.exc_stride_k + m) : 0.0;   // exc_stride_k == 0: constant
    };
    double e_next = fetch(0);

    for (int tile = 0; tile < ntiles; ++tile) {
        const int k0 = tile * kW64Tile;
        const int steps = min(kW64Tile, a.K - k0);
        double* es = excw + (tile & 1) * kW64Tile;
        es[lane] = e_next;
        __syncwarp();
        if (tile + 1 < ntiles) e_next = fetch(tile + 1);
        if (CENTRAL == 2) {                                 // general path: resynchronise with a full-precision exp per tile
#pragma unroll
            for (int i = 0; i < UPL; ++i) G[i] = -expm1(dur[i] * sc.mitau);
        }

        int fphase = FORCES ? k0 % a.forces_every : 0;        // steps since the last captured one
        auto do_step = [&](int s, auto last_tag) {
            constexpr bool LAST = decltype(last_tag)::value;
            const double e = __dmul_rn(es[s], sc.input_scale);                    // muscle.py:275, once per muscle-step
            double r0[UPL];
            bool on[UPL];
            unsigned gmask = 0;
#pragma unroll
            for (int i = 0; i < UPL; ++i) {
                r0[i] = __dadd_rn(__dsub_rn(e, thr[i]), sc.min_rate);            // pool:169-170
                on[i] = !(r0[i] < sc.min_rate);                                   // pool:171
                gmask |= (__any_sync(0xffffffffu, on[i]) ? 1u : 0u) << i;
            }
            // A unit that is not recruited produces no force and (Potvin fibers) keeps its state, so a whole GROUP of 32
            // units without a recruited one costs nothing beyond the three operations above.  Thresholds grow with the
            // unit index, so the recruited groups are the leading ones: `kon` = groups to advance (warp-uniform; any
            // table order is handled, a hole just computes a group of zeros).  One branch per step selects straight-
            // line code for kon groups, inside which the chains interleave as before.
#if MB_W64_SKIP == 2
            const int kon = 32 - __clz(gmask);
#elif MB_W64_SKIP == 1
            const int kon = gmask ? UPL : 0;                                     // only a muscle entirely at rest is skipped
#else
            const int kon = UPL;
#endif
            bool write_now = false;
            if (FORCES) {                                     // every forces_every-th step, counted (a per-step modulo is ~20 instructions)
                write_now = ++fphase == a.forces_every;
                if (write_now) fphase = 0;
            }
            auto emit = [&](int i, double F, bool oni) {                        // per-unit outputs of the FORCES / LAST variants
                if (FORCES || LAST) {
                    const int u = lane + 32 * i;
                    if (u < n) {
                        if (FORCES && write_now) {
                            const int64_t idx = (int64_t)((k0 + s) / a.forces_every) * a.M * n + m * n + u;
                            ((double*)a.forces_dec)[idx] = F;
                            if (a.mask_dec) a.mask_dec[idx] = oni ? 1 : 0;
                        }
                        if (LAST) ((double*)a.forces_last)[m * n + u] = F;
                    }
                }
            };
            auto body = [&](auto kon_tag) {
                constexpr int KON = decltype(kon_tag)::value;
                double x[KON > 0 ? KON : 1], nf[KON > 0 ? KON : 1];
                bool lin[KON > 0 ? KON : 1];
                bool all_lin = true;
#pragma unroll
                for (int i = 0; i < KON; ++i) {
                    double fr;
                    if (CENTRAL == 1) {
                        // (plain compare + select: fmin() drags NaN handling along -- DSETP.MIN + SEL + FSEL + LOP3 + moves)
                        const double r = (r0[i] > peak[i]) ? peak[i] : r0[i];     // pool:173 (gain == 1), :176-177
                        const double q = fma(r, phir[i], -phir6[i]);              // pool:231-232
                        const double fa = fma(-q, G[i], r);                       // pool:217-221, :121 (dur BEFORE the update)
                        fr = on[i] ? fa : 0.0;
                        // on-step counter and G <- G * decay + (1 - decay) of a firing unit (pool:196-197), branch-free:
                        // written as `if (on) {...}` the compiler emits a divergent branch per unit, which serialises the chains
                        const double Gn = fma(G[i], decay, omd);
                        G[i] = on[i] ? Gn : G[i];
                        asm("{\n\t.reg .pred p;\n\tsetp.geu.f64 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}"   // one predicated add
                            : "+r"(cnt[i]) : "d"(r0[i]), "d"(sc.min_rate));
                    } else {
                        double r = (on[i] ? r0[i] : 0.0);                         // pool:172, :173 (gain == 1)
                        r = (r > peak[i]) ? peak[i] : r;                          // pool:176-177
                        fr = r;
                        if (CENTRAL == 2) {
                            const double q = fma(r, phir[i], -phir6[i]);
                            const double ad = q * G[i];                           // pool:217-219 (dur BEFORE the update)
                            fr = r - ((ad < 0.0) ? 0.0 : ad);                     // pool:221, :121
                            if (r > 0.0) {                                        // pool:196-197
                                const double dn = dur[i] + dt;
                                const bool clamp = dn > max_dur;                  // pool:202-203
                                dur[i] = clamp ? max_dur : dn;
                                G[i] = clamp ? Gmax : fma(G[i], decay, omd);
                            }
                        }
                    }
                    x[i] = fma(cpf[i], -ctB2[i], ctA2[i]) * fr;                   // cbrt(2) * (fibers:220 with the fatigued contraction time)
                    lin[i] = x[i] <= xlin;                                        // fibers:233-236
                    all_lin = all_lin && lin[i];
                }
                if (!all_lin) {                                                   // one branch for the group of chains (see header)
                    double ex[KON > 0 ? KON : 1];
#pragma unroll
                    for (int i = 0; i < KON; ++i) {
                        ex[i] = exp_neg_fast<CENTRAL != 1>(-(x[i] * x[i]) * x[i], exp2_tab, a.ec);   // exp(-2 x^3)
                        // (opaque use: otherwise the compiler sinks each exponential into a divergent per-unit branch
                        //  `lin ? line : sigmoid`, and the four Horner chains run one after the other)
                        asm volatile("" : "+d"(ex[i]));
                    }
#pragma unroll
                    for (int i = 0; i < KON; ++i) nf[i] = lin[i] ? x[i] * c25s : 1.0 - ex[i];   // fibers:233-246
                } else {
#pragma unroll
                    for (int i = 0; i < KON; ++i) nf[i] = x[i] * c25s;
                }
                double fsum = 0.0;
#pragma unroll
                for (int i = 0; i < KON; ++i) {
                    const double F = nf[i] * cpf[i];                              // fibers:258 (capacity BEFORE the update)
                    cpf[i] = fatigue_update<FIB == MB_FIBERS_PYMUSCLE>(cpf[i], nf[i], fatdt[i], rdidt[i], ptf[i]);
                    fsum = (i == 0) ? F : fsum + F;
                    emit(i, F, on[i]);
                }
#pragma unroll
                for (int i = KON; i < UPL; ++i) {                                 // groups at rest
                    if (FIB == MB_FIBERS_PYMUSCLE) cpf[i] = fatigue_update<true>(cpf[i], 0.0, fatdt[i], rdidt[i], ptf[i]);   // recovery, pmf:121-141
                    emit(i, 0.0, false);
                }
                ftw[s * kW64Pitch + lane] = fsum;
            };
            switch (kon) {
                case 0: body(std::integral_constant<int, 0>{}); break;
                case 1: body(std::integral_constant<int, 1>{}); break;
                case 2: if constexpr (UPL >= 2) body(std::integral_constant<int, 2>{}); break;
                case 3: if constexpr (UPL >= 3) body(std::integral_constant<int, 3>{}); break;
                default: if constexpr (UPL >= 4) body(std::integral_constant<int, 4>{}); break;
            }
        };
        const bool has_last = a.forces_last != nullptr && tile == ntiles - 1;
        const int fast_steps = has_last ? steps - 1 : steps;
#pragma unroll (kW64Unroll)
        for (int s = 0; s < fast_steps; ++s) do_step(s, std::false_type{});
        if (has_last) do_step(steps - 1, std::true_type{});
        __syncwarp();
        // ---- lane s sums the 32 lanes' forces of step k0 + s ----------------------------------------------
        const int kl = k0 + lane;
        if (lane < steps && a.totals && (a.totals_every == 1 || (kl + 1) % a.totals_every == 0)) {
            const double* rp = ftw + lane * kW64Pitch;
            double acc0 = rp[0], acc1 = rp[1], acc2 = rp[2], acc3 = rp[3];
#pragma unroll
            for (int c = 4; c < 32; c += 4) {
                acc0 += rp[c]; acc1 += rp[c + 1]; acc2 += rp[c + 2]; acc3 += rp[c + 3];
            }
            const double v = ((acc0 + acc1) + (acc2 + acc3)) / sc.out_div;        // muscle.py:278
            const int64_t krow = kl / a.totals_every;
            MB_ASSERT(krow < (a.K + a.totals_every - 1) / a.totals_every);
            ((double*)a.totals)[krow * a.totals_stride_k + m] = v;
            if (a.n_peers > 0) {                                                  // peer GPUs' gather buffers (NVLink stores)
                const int64_t op = (m >> 1) * a.peer_stride_pair + krow * a.peer_stride_k + (m & 1) * a.peer_stride_odd;
                for (int pr = 0; pr < a.n_peers; ++pr) peer_store((double*)a.totals_peer[pr] + op, v, a.peer_multicast);
            }
        }
        __syncwarp();
    }

#pragma unroll
    for (int i = 0; i < UPL; ++i) {
        const int u = lane + 32 * i;
        if (u < n) {
            const int64_t idx = m * n + u;
            if (a.peripheral) a.cpf[idx] = cpf[i];
            if (CENTRAL == 1) a.dur[idx] = fmin(fma((double)cnt[i], dt, a.dur[idx]), max_dur);   // pool:196-203 (monotone: clamp once)
            if (CENTRAL == 2) a.dur[idx] = dur[i];
        }
    }
}

}  // namespace mb
