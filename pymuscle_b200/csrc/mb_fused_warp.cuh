// mb_fused_warp.cuh -- K2w: the fused K-step kernel for muscles of 12..128 units, FP32 mode.
//
// One WARP owns two consecutive muscles for the whole window and never talks to another
// warp: lane l holds units l, l+32, l+64, l+96 (UPL = ceil(n/32) of them) of both muscles.
// The two muscles of a unit form one packed pair (FFMA2 / FMUL2 / FADD2, CoreF32 with U = 2),
// so a thread carries UPL independent pair chains with UPL parameter sets in registers.
//
// Per step a thread adds its UPL unit forces (packed adds) and stores one float2 into the
// warp's private tile FTw[32][33] (shared memory).  Time is cut into tiles of 32 steps; after a
// tile (one __syncwarp) lane s sums row s -- 32 bank-conflict-free 64-bit LDS -- and writes the
// two muscles' totals of step k0+s.  There is no cross-warp reduction, no mbarrier and no
// block barrier in the loop; warps drift freely.
//
// Excitations: lane s fetches the pair's two values of step k0+s of the NEXT tile while the
// current tile is being computed (one load per lane per 32 steps), parks them in a private
// shared-memory row, and every step is one broadcast LDS.64.
//
// exp(-dur/tau) - 1 is carried by recurrence and re-synchronised once per tile (CoreF32,
// CENTRAL == 1); the compensated capacity pair is renormalised once per tile.
#pragma once
#include "mb_fused.cuh"
#include <type_traits>

namespace mb {

constexpr int kWTile = 32;       // steps per tile: lanes of stage 1 = steps
constexpr int kWPitch = 33;      // row pitch of FTw in float2 (odd => conflict free)
constexpr int kWarpsPerBlock = 8;

#ifndef MB_W_MINB
#define MB_W_MINB 2
#endif

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}

// PERSIST (short windows): with K = 8..128 steps per launch a muscle pair is a few microseconds of arithmetic between
// two round trips to HBM -- state in, state out -- and ncu shows a third of the warp samples waiting on exactly those
// loads (K = 32: 258 us against 182 us of arithmetic).  The persistent variant launches only the resident warps; each
// walks a strided list of muscle pairs and, while it advances pair j, has the state of pair j+1 copied into its own
// shared-memory stage by cp.async (the stage is free as soon as pair j's state is in registers).  At the end of a pair
// nothing is read back: capacities are stored from the compensated float pair (a unit that sits at its peak gets the
// exact float64 peak, so rested units keep their bits) and on-durations are advanced by a fire-and-forget
// red.add.f64 of count * dt.
template <int FIB, int CENTRAL, int UPL, bool FORCES, bool PERSIST = false>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, MB_W_MINB) k2w_f32(const K2Args a) {
    using Core = CoreF32<FIB, CENTRAL, 2>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kPerWarp = kWTile * kWPitch + 2 * kWTile;                      // float2 elements: force tile + excitations
    float2* ftw = (float2*)smem_raw + warp * kPerWarp;                           // [32][33]
    float2* excw = ftw + kWTile * kWPitch;                                       // [2][32]
    const int n = a.n;
    // PERSIST: per-warp stage [2][2n] doubles (capacities | durations of the NEXT pair), behind all warps' tiles
    double* stage = (double*)((float2*)smem_raw + kWarpsPerBlock * kPerWarp) + (size_t)warp * 4 * ((n + 1) & ~1);
    const int stage_n = 2 * ((n + 1) & ~1);                                      // doubles per array in the stage

    // (32-bit pair indices: the launcher takes batches of fewer than 2^31 muscles here; the persistent loop is short of
    //  registers, and every 64-bit loop-carried value costs the step loop's schedule)
    const int npairs = (int)((a.M + 1) >> 1);
    const int first = (int)(blockIdx.x * kWarpsPerBlock + warp);
    if (first >= npairs) return;                          // warps are independent: no block barrier below
    const int ntiles = (a.K + kWTile - 1) / kWTile;
    const float* exc = (const float*)a.exc;

    auto prefetch = [&](int pair) {                       // state of `pair` -> this warp's stage (cp.async, one group)
        const int64_t m0 = (int64_t)pair * 2;
        const int cnt = (m0 + 1 < a.M) ? 2 * n : n;       // units to fetch; m0 even => the source is 16-byte aligned
        const uint32_t sc = smem_u32(stage), sd = smem_u32(stage + stage_n);
        for (int i = lane; i < (cnt >> 1); i += 32) {
            cp_async16(sc + 16u * i, a.cpf + m0 * n + 2 * i);
            if (CENTRAL) cp_async16(sd + 16u * i, a.dur + m0 * n + 2 * i);
        }
        if ((cnt & 1) && lane == 0) {
            cp_async8(sc + 8u * (cnt - 1), a.cpf + m0 * n + cnt - 1);
            if (CENTRAL) cp_async8(sd + 8u * (cnt - 1), a.dur + m0 * n + cnt - 1);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (PERSIST) prefetch(first);

    // unit parameters: the same for every pair this warp advances -- loaded (L2) and converted once
    Core core[UPL];
#pragma unroll
    for (int i = 0; i < UPL; ++i) {
        const int u = lane + 32 * i;
        core[i].load(a, u < n ? u : 0, 0);
        if (u >= n) core[i].disable();                                  // idle slot: never recruited, force 0
    }

    auto run_pair = [&](const int pair) {
    const int64_t m0 = (int64_t)pair * 2;
    const bool two = m0 + 1 < a.M;
    MB_ASSERT(m0 < a.M && (int64_t)(m0 + (two ? 2 : 1)) * n <= a.M * (int64_t)n);
    if (PERSIST) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
    }
#pragma unroll
    for (int i = 0; i < UPL; ++i) {
        const int u = lane + 32 * i;
        if (u < n) {
            if (PERSIST) {
                core[i].init(0, a, stage[u], CENTRAL ? stage[stage_n + u] : 0.0);
                // (a batch with an odd muscle count: the idle second half mirrors the first, its outputs are never stored)
                core[i].init(1, a, stage[two ? n + u : u], CENTRAL ? stage[stage_n + (two ? n + u : u)] : 0.0);
            } else {
                core[i].init(0, a, a.cpf[m0 * n + u], CENTRAL ? a.dur[m0 * n + u] : 0.0);
                if (two) core[i].init(1, a, a.cpf[(m0 + 1) * n + u], CENTRAL ? a.dur[(m0 + 1) * n + u] : 0.0);
            }
        }
    }
    if (PERSIST) {
        __syncwarp();                                     // every lane has its state in registers: the stage is free
        if (pair + (int)(gridDim.x * kWarpsPerBlock) < npairs) prefetch(pair + (int)(gridDim.x * kWarpsPerBlock));
    }

    // excitations of (tile, step = lane) for the two muscles; 0 beyond the window / batch
    auto fetch = [&](int tile) -> float2 {
        float2 e = make_float2(0.0f, 0.0f);
        const int k = tile * kWTile + lane;
        if (k < a.K) {
            const float* src = exc + (int64_t)(k / a.exc_hold) * a.exc_stride_k + m0;   // exc_stride_k == 0: constant
            e.x = __ldg(src);
            if (two) e.y = __ldg(src + 1);
        }
        return e;
    };
    float2 e_next = fetch(0);

    for (int tile = 0; tile < ntiles; ++tile) {
        const int k0 = tile * kWTile;
        const int steps = min(kWTile, a.K - k0);
        float2* es = excw + (tile & 1) * kWTile;
        es[lane] = e_next;
        __syncwarp();
        if (tile + 1 < ntiles) e_next = fetch(tile + 1);
#pragma unroll
        for (int i = 0; i < UPL; ++i) core[i].begin_tile();

        // one step of the warp's two muscles; LAST = the window's final step (current_forces wanted)
        int fphase = FORCES ? k0 % a.forces_every : 0;        // steps since the last captured one
        auto do_step = [&](int s, auto last_tag) {
            constexpr bool LAST = decltype(last_tag)::value;
            const float2 e = es[s];
            float2 fsum;
            bool write_now = false;
            if (FORCES) {                                     // every forces_every-th step, counted (a per-step modulo is ~20 instructions)
                write_now = ++fphase == a.forces_every;
                if (write_now) fphase = 0;
            }
#pragma unroll
            for (int i = 0; i < UPL; ++i) {
                bool oa, ob;
                const float2 F = core[i].step_pair(0, e, oa, ob);
                fsum = (i == 0) ? F : __fadd2_rn(fsum, F);
                if (FORCES || LAST) {
                    const int u = lane + 32 * i;
                    if (u < n) {
                        if (FORCES && write_now) {
                            const int64_t base = (int64_t)((k0 + s) / a.forces_every) * a.M * n;
                            float* fd = (float*)a.forces_dec;
                            fd[base + m0 * n + u] = F.x;
                            if (two) fd[base + (m0 + 1) * n + u] = F.y;
                            if (a.mask_dec) {
                                a.mask_dec[base + m0 * n + u] = oa ? 1 : 0;
                                if (two) a.mask_dec[base + (m0 + 1) * n + u] = ob ? 1 : 0;
                            }
                        }
                        if (LAST) {
                            float* fl = (float*)a.forces_last;
                            fl[m0 * n + u] = F.x;
                            if (two) fl[(m0 + 1) * n + u] = F.y;
                        }
                    }
                }
            }
            ftw[s * kWPitch + lane] = fsum;
        };
        const bool has_last = a.forces_last != nullptr && tile == ntiles - 1;
        const int fast_steps = has_last ? steps - 1 : steps;
        for (int s = 0; s < fast_steps; ++s) do_step(s, std::false_type{});
        if (has_last) do_step(steps - 1, std::true_type{});
#pragma unroll
        for (int i = 0; i < UPL; ++i) core[i].renormalise();
        __syncwarp();
        // ---- stage 1 (and only): lane s sums the 32 lanes' forces of step k0 + s -----------------
        const int kl = k0 + lane;                             // the step whose totals this lane writes
        if (lane < steps && (a.totals_every == 1 || (kl + 1) % a.totals_every == 0)) {
            const int64_t krow = kl / a.totals_every;
            const float2* rp = ftw + lane * kWPitch;
            float2 acc0 = rp[0], acc1 = rp[1];
#pragma unroll
            for (int c = 2; c < 32; c += 2) {
                acc0 = __fadd2_rn(acc0, rp[c]);
                acc1 = __fadd2_rn(acc1, rp[c + 1]);
            }
            const float2 acc = __fadd2_rn(acc0, acc1);
            if (a.totals) {
                const int64_t o = krow * a.totals_stride_k + m0;
                MB_ASSERT(krow < (a.K + a.totals_every - 1) / a.totals_every && m0 < a.M);
                const float t0 = core[0].out_scale(acc.x), t1 = core[0].out_scale(acc.y);
                float* t = (float*)a.totals + o;
                t[0] = t0;
                if (two) t[1] = t1;
                if (a.n_peers > 0) {
                    // the same totals into the peer GPUs' gather buffers.  With the pair-major layout
                    // (stride_k == 2) the 32 lanes of the warp write 256 contiguous bytes per tile
                    const int64_t op = (m0 >> 1) * a.peer_stride_pair + krow * a.peer_stride_k;
                    const bool vec = two && a.peer_stride_odd == 1 && (((a.peer_stride_pair | a.peer_stride_k) & 1) == 0);
                    for (int pr = 0; pr < a.n_peers; ++pr) {
                        float* tp = (float*)a.totals_peer[pr] + op;
                        if (vec) {
                            peer_store2(tp, t0, t1, a.peer_multicast);
                        } else {
                            peer_store(tp, t0, a.peer_multicast);
                            if (two) peer_store(tp + a.peer_stride_odd, t1, a.peer_multicast);
                        }
                    }
                }
            }
        }
        __syncwarp();
    }

#pragma unroll
    for (int i = 0; i < UPL; ++i) {
        const int u = lane + 32 * i;
        if (u < n) {
            if (PERSIST) {
                const float lo2 = __ldg((const float4*)a.params + 2 * (int64_t)n + u).y;   // third float of the peak (L2)
                core[i].store_direct(0, a, m0 * n + u, lo2);
                if (two) core[i].store_direct(1, a, (m0 + 1) * n + u, lo2);
            } else {
                core[i].store(0, a, m0 * n + u);
                if (two) core[i].store(1, a, (m0 + 1) * n + u);
            }
        }
    }
    };   // run_pair
    if constexpr (PERSIST) {
        const int stride = (int)(gridDim.x * kWarpsPerBlock);
        for (int pair = first; pair < npairs; pair += stride) run_pair(pair);
    } else {
        run_pair(first);                                  // exactly one pair per warp
    }
}

// ======================================================================================
// K2u: one warp per MUSCLE, pairs of UNITS -- the fused kernel for per-instance parameter tables.
//
// Lane l holds units l, l+32 (one pair) and, for n > 64, l+64, l+96 (a second pair) of a single
// muscle; the two units of a pair are the two halves of the packed operands, so the per-unit
// parameters are register pairs too and every muscle can have its own table (domain
// randomisation, MbConfig::param_rows) at no extra cost.  ~80 registers, three blocks per SM.
// Everything else is K2w: private force tile, lane s sums row s, no cross-warp traffic.
// (With one shared table it runs at 1.22e12 unit-steps/s against K2w's 1.33e12 on the headline
// shape -- per-step overhead is spread over 4 instead of 8 unit-instances -- so K2w keeps that case.)
// ======================================================================================
constexpr int kUPitch = 34;      // row pitch of the float force tile (even: 64-bit row reads, conflict free)

__device__ __forceinline__ float2 P2(float a, float b) { return make_float2(a, b); }

// Ragged rows (round 2): the same kernel takes one muscle (row, segment) per warp of a ragged batch whose muscles have
// at most 256 units (NPU <= 4 pairs per lane; pairs beyond a muscle's length are skipped by warp-uniform predicates) --
// a hopper leg's 74 / 120 / 169 / 120 units -- where the block-per-muscle kernel would give a 74-unit muscle 256 threads.
// (PyMuscle fibers with two or more pairs per lane need 100+ registers: two blocks per SM instead of three)
template <int FIB, int CENTRAL, int NPU, bool FORCES>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, ((FIB == MB_FIBERS_PYMUSCLE && NPU == 2) || NPU > 2) ? 2 : 3)
k2u_f32(const K2xArgs xa) {
    const K2Args& a = xa.k;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* ftw = (float*)smem_raw + warp * (kWTile * kUPitch + 2 * kWTile);     // [32][34]
    float* excw = ftw + kWTile * kUPitch;                                        // [2][32]

    const int64_t m = (int64_t)blockIdx.x * kWarpsPerBlock + warp;              // muscle = (row, segment)
    if (m >= a.M) return;                                 // warps are independent: no barrier below
    const int64_t row = m / xa.S;
    const int seg = (int)(m - row * xa.S);
    const int off0 = xa.seg_off ? xa.seg_off[seg] : 0;
    const int n = xa.seg_off ? xa.seg_off[seg + 1] - off0 : (int)xa.L;
    const int64_t base = row * xa.L + off0;                // first state element of this muscle
    const int npa = (n + 63) >> 6;                         // pairs per lane this muscle needs (warp-uniform)
    MB_ASSERT(npa <= NPU && base + n <= xa.total_units);
    const int ntiles = (a.K + kWTile - 1) / kWTile;
    const float* exc = (const float*)a.exc;
    const PairScal sc = pair_scal(a);
    const float inv_out = xa.seg_div ? (float)(1.0 / xa.seg_div[seg]) : a.sf.inv_out;
    // per-instance tables: this muscle's own table (uniform batches); else the row's shared table at the muscle's offset
    const float4* table = (const float4*)a.params + (a.param_stride ? m * a.param_stride : off0);
    const int64_t tabL = xa.L;

    PairParams<float2> pp[NPU];
    PairState st[NPU];
#pragma unroll
    for (int i = 0; i < NPU; ++i) {
        if (i >= npa) continue;
        const int u0 = lane + 64 * i, u1 = u0 + 32;
        const UnitF a0 = load_unit(table, tabL, u0 < n ? u0 : 0);
        const UnitF a1 = load_unit(table, tabL, u1 < n ? u1 : 0);
        const PairParams<float> q0 = pair_params(a0, a), q1 = pair_params(a1, a);
        const float inf = __int_as_float(0x7f800000);
        pp[i].crit = P2(u0 < n ? q0.crit : inf, u1 < n ? q1.crit : inf);         // idle slot: never recruited, force 0
        pp[i].thr_g = P2(q0.thr_g, q1.thr_g);   pp[i].peak = P2(q0.peak, q1.peak);
        pp[i].phir = P2(q0.phir, q1.phir);      pp[i].phir6 = P2(q0.phir6, q1.phir6);
        pp[i].ctA = P2(q0.ctA, q1.ctA);         pp[i].ctB = P2(q0.ctB, q1.ctB);
        pp[i].fatdt = P2(q0.fatdt, q1.fatdt);   pp[i].rdidt = P2(q0.rdidt, q1.rdidt);
        pp[i].ptf = P2(q0.ptf, q1.ptf);         pp[i].ptf_lo = P2(q0.ptf_lo, q1.ptf_lo);
        pair_init(st[i], pp[i], 0, u0 < n ? a.cpf[base + u0] : unit_peak(a0),
                  (CENTRAL && u0 < n) ? a.dur[base + u0] : 0.0, a.adk_d);
        pair_init(st[i], pp[i], 1, u1 < n ? a.cpf[base + u1] : unit_peak(a1),
                  (CENTRAL && u1 < n) ? a.dur[base + u1] : 0.0, a.adk_d);
    }

    // excitation of (tile, step = lane); 0 beyond the window.  exc_stride_k == 0: constant
    auto fetch = [&](int tile) -> float {
        const int k = tile * kWTile + lane;
        return (k < a.K) ? __ldg(exc + (int64_t)(k / a.exc_hold) * a.exc_stride_k + m) : 0.0f;
    };
    float e_next = fetch(0);

    for (int tile = 0; tile < ntiles; ++tile) {
        const int k0 = tile * kWTile;
        const int steps = min(kWTile, a.K - k0);
        float* es = excw + (tile & 1) * kWTile;
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
                    const int u0 = lane + 64 * i, u1 = u0 + 32;
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
            ftw[s * kUPitch + lane] = fsum.x + fsum.y;
        };
        const bool has_last = a.forces_last != nullptr && tile == ntiles - 1;
        const int fast_steps = has_last ? steps - 1 : steps;
        for (int s = 0; s < fast_steps; ++s) do_step(s, std::false_type{});
        if (has_last) do_step(steps - 1, std::true_type{});
#pragma unroll
        for (int i = 0; i < NPU; ++i)
            if (i < npa) pair_renormalise<FIB>(st[i], pp[i]);
        __syncwarp();
        // ---- lane s sums the 32 lanes' forces of step k0 + s -----------------------------------
        const int kl = k0 + lane;                             // the step whose total this lane writes
        if (lane < steps && (a.totals_every == 1 || (kl + 1) % a.totals_every == 0)) {
            const int64_t krow = kl / a.totals_every;
            const float2* rp = (const float2*)(ftw + lane * kUPitch);
            float2 acc0 = rp[0], acc1 = rp[1];
#pragma unroll
            for (int c = 2; c < 16; c += 2) {
                acc0 = __fadd2_rn(acc0, rp[c]);
                acc1 = __fadd2_rn(acc1, rp[c + 1]);
            }
            const float2 acc = __fadd2_rn(acc0, acc1);
            if (a.totals) {
                const float v = (acc.x + acc.y) * inv_out;
                const int64_t o = krow * a.totals_stride_k + m;
                ((float*)a.totals)[o] = v;
                if (a.n_peers > 0) {
                    const int64_t op = (m >> 1) * a.peer_stride_pair + krow * a.peer_stride_k + (m & 1) * a.peer_stride_odd;
                    for (int pr = 0; pr < a.n_peers; ++pr) peer_store((float*)a.totals_peer[pr] + op, v, a.peer_multicast);
                }
            }
        }
        __syncwarp();
    }

#pragma unroll
    for (int i = 0; i < NPU; ++i) {
        if (i >= npa) continue;
        const int u0 = lane + 64 * i, u1 = u0 + 32;
        if (u0 < n) pair_store<CENTRAL != 0>(st[i], 0, a, base + u0);
        if (u1 < n) pair_store<CENTRAL != 0>(st[i], 1, a, base + u1);
    }
}

}  // namespace mb
