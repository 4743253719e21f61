// mb_fused.cuh -- K2: K fused steps per launch, unit state in registers.
//
// Uniform batch: M muscles x n units.  A block owns G = q*U consecutive muscles.
// Muscle group tq (U muscles) occupies `ns` consecutive threads, ns = n rounded up to a
// multiple of 8; thread tq*ns + u handles unit u of muscles tq*U .. +U-1: its U
// unit-instances share one set of per-unit parameters (registers, loaded once per window
// as 128-bit vectors) and give U independent dependency chains.  Per step: one vector LDS
// (U excitations), the arithmetic of mb_core.cuh, one vector STS (U forces).
//
// Per-(step, muscle) totals without any block-wide barrier.  Time is cut into tiles of
// T = 8 steps and the sum over a muscle's units is done in two stages:
//
//   stage 1  warp-local.  During a tile every warp writes its lanes' forces into its own
//            tile FTw[T][32] (row pitch 33 vectors).  After the tile (one __syncwarp) lane
//            (part = lane >> 3, s = lane & 7) sums the 8 columns part*8 .. +7 of row s with
//            8 independent 128-bit LDS (bank-conflict free) and stores the partial into
//            PT[tile & 3][s][4*warp + part].  Group boundaries are multiples of 8 threads, so
//            every 8-column part belongs to exactly one muscle group.
//   stage 2  one warp per tile (rotating: warp (tile % warps)) waits on the mbarrier that
//            counts the warps' stage-1 arrivals, sums the ns/8 partials of every
//            (step, muscle) and stores the totals to HBM (coalesced).
//
// PT is a ring of 4 tiles guarded by mbarriers (pfull: 1 arrival per warp; pfree: 1 arrival
// by the stage-2 warp), so warps may drift up to 3 tiles apart and never wait for each
// other in steady state.  Warp 0 additionally stages the per-step excitations two tiles
// ahead with cp.async.bulk (1-D bulk copy per step row, mbarrier complete_tx) into an
// 8-slot ring; consumers wait on the slot's mbarrier.
//
// HBM traffic per window: state in/out (32 B/unit) plus excitations and totals
// (2*sizeof(T)/n B per unit-step) -- see DESIGN.md.
#pragma once
#include "mb_core.cuh"

namespace mb {

constexpr int kExcSlots = 8;     // excitation ring (tiles)
constexpr int kPtSlots = 4;      // partial-sum ring (tiles)
constexpr int kTile = 8;         // steps per tile (stage 1 maps lane & 7 to the step)
constexpr int kFtPitch = 33;     // row pitch of a warp's force tile, in vectors (odd => conflict free)

struct K2Args {
    const void* params;
    int64_t M;
    int32_t n, ns, q, K, T, G, Gp, npg, ptp;   // ns: threads per muscle group; npg = ns/8 parts per group; ptp: PT row pitch
    const void* exc;
    int64_t exc_stride_k;
    int32_t exc_mode;            // 0 = bulk copy, 1 = element-wise loads, 2 = constant over the window
    double* cpf;
    double* dur;
    void* totals;
    int64_t totals_stride_k;
    void* forces_dec;
    uint8_t* mask_dec;
    int32_t forces_every;
    void* forces_last;
    int32_t peripheral;
    double dt, adk_d, max_dur;
    double decay_d;              // exp(-dt/tau): per-step factor of exp(-dur/tau) for a firing unit
    ScalF sf;
    ScalD sd;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// 1-D bulk async copy global -> shared (SASS UBLKCP); bytes % 16 == 0, both addresses 16-B aligned.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// U-wide vector of T in shared memory (one LDS/STS of 4*U or 8*U bytes)
template <typename T, int U> struct alignas(sizeof(T) * U) VecU { T v[U]; };

// ======================================================================================
// Per-thread arithmetic of the two precision modes.  Each "core" holds the unit's
// parameters and the state of its U instances in registers.
// ======================================================================================

// ---- FP32 mode ------------------------------------------------------------------------
// Capacity is carried as hi + lo (two floats): `hi` is frozen for a tile, per-step
// decrements accumulate in the small `lo` (so their rounding error is relative to the
// tile's decrement, not to the capacity) and the pair is renormalised (error-free
// TwoSum) once per tile.  On-duration is dur0 + count*dt with an exact float step
// counter.  Plain float accumulators miss rtol 1e-4 over 10,000 steps.
//
// CENTRAL: 0 = pool.apply_fatigue off (adaptation is the identity)
//          1 = on, fast path: the launcher has checked that (a) the adaptation of a
//              recruited unit cannot be negative and (b) max_duration/tau is so large
//              that the duration clamp cannot change exp(-dur/tau); both clamps drop out.
//              E = exp(-dur/tau) is carried by recurrence -- a firing step multiplies it by
//              exp(-dt/tau) -- and re-synchronised with one ex2 per tile from the exact
//              counter, which replaces the per-step MUFU of the adaptation curve
//          2 = on, general path with both clamps (pool:202-206, :221)
template <int FIB, int CENTRAL, int U>
struct CoreF32 {
    using T = float;
    static constexpr bool kCentral = CENTRAL != 0;
    using ParamVec = float4;
    UnitF p;
    ScalF sc;
    float fatdt, rdidt, dtk, decay;
    float hi[U], lo[U], cpf[U], cnt[U], arg0[U], room[U], E[U];

    __device__ __forceinline__ void load(const K2Args& a, int u) {
        p = load_unit((const float4*)a.params, a.n, u);
        sc = a.sf;
        fatdt = a.peripheral ? (float)((double)p.fat * a.dt) : 0.0f;
        rdidt = a.peripheral ? (float)((double)p.rdi * a.dt) : 0.0f;
        dtk = (float)(a.dt * a.adk_d);
        decay = (float)a.decay_d;
    }
    __device__ __forceinline__ double rest_capacity() const { return (double)p.ptf + (double)p.ptf_lo; }
    __device__ __forceinline__ void init(int j, const K2Args& a, double c, double d0) {
        hi[j] = (float)c;
        lo[j] = (float)(c - (double)hi[j]);
        cpf[j] = hi[j] + lo[j];
        room[j] = (p.ptf - hi[j]) + p.ptf_lo;            // peak - hi: the upper bound of lo (pmf:104-105)
        cnt[j] = 0.0f;
        arg0[j] = (float)(d0 * a.adk_d);
        E[j] = 1.0f;
    }
    __device__ __forceinline__ void begin_tile() {
        if (CENTRAL == 1) {
#pragma unroll
            for (int j = 0; j < U; ++j) E[j] = ex2_approx(fmaf(cnt[j], dtk, arg0[j]));
        }
    }
    // one step of instance j; returns the unit force
    __device__ __forceinline__ float step(int j, float e, bool& on) {
        const float r = pool_rate_raw(p, sc, e, on);
        float fr;
        if (CENTRAL == 0) {
            fr = on ? r : 0.0f;
        } else if (CENTRAL == 1) {
            const float q = fmaf(r, p.phir, -p.phir6);                     // pool:231-232
            fr = r - fmaf(-q, E[j], q);                                    // pool:217-221, :121 (dur BEFORE the update)
            fr = on ? fr : 0.0f;
            if (on) {                                                      // pool:196-197
                cnt[j] += 1.0f;
                E[j] *= decay;
            }
        } else {
            float arg = fmaf(cnt[j], dtk, arg0[j]);
            if (on) cnt[j] += 1.0f;
            arg = fmaxf(arg, sc.argmin);
            const float r0 = on ? r : 0.0f;
            fr = r0 - fmaxf(pool_adaptation(p, r0, arg), 0.0f);
        }
        float nf;
        const float F = fibers_force(p, sc, fr, cpf[j], nf);
        // capacity update on the compensated pair (fibers:110-114, pmf:94-105)
        float l = fmaf(nf, -fatdt, lo[j]);
        if (FIB == MB_FIBERS_PYMUSCLE) {
            const float gap = room[j] - lo[j];                             // peak - capacity
            l = fmaf((nf <= 0.0f) ? gap : 0.0f, rdidt, l);                 // pmf:121-141
        }
        l = fmaxf(l, -hi[j]);
        if (FIB == MB_FIBERS_PYMUSCLE) l = fminf(l, room[j]);
        lo[j] = l;
        cpf[j] = hi[j] + l;
        return F;
    }
    // renormalise (TwoSum) so that |lo| <= ulp(hi)/2 again
    __device__ __forceinline__ void end_tile() {
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const float s1 = hi[j] + lo[j];
            const float bb = s1 - hi[j];
            const float err = (hi[j] - (s1 - bb)) + (lo[j] - bb);
            hi[j] = s1;
            lo[j] = err;
            room[j] = (p.ptf - s1) + p.ptf_lo;
        }
    }
    __device__ __forceinline__ void store(int j, const K2Args& a, int64_t idx) {
        if (a.peripheral) {
            // add the window's change to the float64 state: an untouched unit keeps its
            // exact bits, and the low bits of the incoming capacity are not truncated
            const double c0 = a.cpf[idx];
            const float h0 = (float)c0;
            const float l0 = (float)(c0 - (double)h0);
            a.cpf[idx] = c0 + (((double)hi[j] - (double)h0) + ((double)lo[j] - (double)l0));
        }
        if (CENTRAL) {
            double d = fma((double)cnt[j], a.dt, a.dur[idx]);
            d = (d > a.max_dur) ? a.max_dur : d;
            a.dur[idx] = (d < 0.0) ? 0.0 : d;
        }
    }
    __device__ __forceinline__ float out_scale(float v) const { return v * sc.inv_out; }
};

// ---- FP64 mode ------------------------------------------------------------------------
// Durations are accumulated exactly as the reference does (dur += dt, clamped).  The
// adaptation factor E = exp(-dur/tau) is carried by recurrence (a firing step multiplies it
// by exp(-dt/tau); a clamped duration pins it to exp(-max_duration/tau)) and recomputed
// with a full-precision exp at every tile start, so its error stays at a few ulps.
template <int FIB, int CENTRAL, int U>
struct CoreF64 {
    using T = double;
    static constexpr bool kCentral = CENTRAL != 0;
    using ParamVec = double2;
    UnitD p;
    ScalD sc;
    double fatdt, rdidt, dt, max_dur, decay, Emax;
    double cpf[U], dur[U], E[U];

    __device__ __forceinline__ void load(const K2Args& a, int u) {
        p = load_unit((const double2*)a.params, a.n, u);
        sc = a.sd;
        fatdt = a.peripheral ? p.fat * a.dt : 0.0;
        rdidt = a.peripheral ? p.rdi * a.dt : 0.0;
        dt = a.dt;
        max_dur = a.max_dur;
        decay = a.decay_d;
        Emax = CENTRAL ? exp(a.max_dur * a.sd.mitau) : 0.0;
    }
    __device__ __forceinline__ double rest_capacity() const { return p.ptf; }
    __device__ __forceinline__ void init(int j, const K2Args&, double c, double d0) {
        cpf[j] = c;
        dur[j] = d0;
        E[j] = 1.0;
    }
    __device__ __forceinline__ void begin_tile() {
        if (CENTRAL) {
#pragma unroll
            for (int j = 0; j < U; ++j) E[j] = exp(dur[j] * sc.mitau);     // pool:217-218
        }
    }
    __device__ __forceinline__ double step(int j, double e, bool& on) {
        const double r = pool_rate(p, sc, e, on);
        double fr = r;
        if (CENTRAL) {
            const double q = fma(r, p.phir, -p.phir6);                     // pool:231-232
            const double ad = q * (1.0 - E[j]);                            // pool:217-219 (dur BEFORE the update)
            fr = r - ((ad < 0.0) ? 0.0 : ad);                              // pool:221, :121
            if (r > 0.0) {                                                 // pool:196-197
                const double dn = dur[j] + dt;
                const bool clamp = dn > max_dur;                           // pool:202-203
                dur[j] = clamp ? max_dur : dn;
                E[j] = clamp ? Emax : E[j] * decay;
            }
        }
        double nf;
        const double F = fibers_force(p, sc, fr, cpf[j], nf);
        cpf[j] = fatigue_update<FIB == MB_FIBERS_PYMUSCLE>(cpf[j], nf, fatdt, rdidt, p.ptf);
        return F;
    }
    __device__ __forceinline__ void end_tile() {}
    __device__ __forceinline__ void store(int j, const K2Args& a, int64_t idx) {
        if (a.peripheral) a.cpf[idx] = cpf[j];
        if (CENTRAL) a.dur[idx] = dur[j];
    }
    __device__ __forceinline__ double out_scale(double v) const { return v / sc.out_div; }   // muscle.py:278
};

// ======================================================================================
// excitation staging (warp 0): tile `tile` of the window -> ring slot tile & 7
// ======================================================================================
template <typename T>
__device__ __forceinline__ void publish_exc(const K2Args& a, T* exc_s, uint64_t* exc_full, int tile, int64_t m0,
                                            int gcount, int lane) {
    const int k0 = tile * a.T;
    const int steps = min(a.T, a.K - k0);
    const int slot = tile & (kExcSlots - 1);
    T* dst = exc_s + slot * a.T * a.Gp;
    const T* src = (const T*)a.exc + (int64_t)k0 * a.exc_stride_k + m0;
    if (a.exc_mode == 0) {
        if (lane == 0) {
            const uint32_t row_bytes = (uint32_t)(gcount * sizeof(T));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&exc_full[slot], row_bytes * steps);
            for (int s = 0; s < steps; ++s)
                bulk_g2s(dst + s * a.Gp, src + (int64_t)s * a.exc_stride_k, row_bytes, &exc_full[slot]);
        }
    } else {
        for (int i = lane; i < steps * a.G; i += 32) {
            const int s = i / a.G, g = i - s * a.G;
            if (g < gcount) dst[s * a.Gp + g] = __ldg(src + (int64_t)s * a.exc_stride_k + g);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&exc_full[slot]);
    }
}

// ======================================================================================
// the kernel
// ======================================================================================
template <typename Core, int U, bool FORCES>
__device__ __forceinline__ void k2_body(const K2Args& a) {
    using T = typename Core::T;
    using V = VecU<T, U>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* exc_full = (uint64_t*)smem_raw;            // [8]
    uint64_t* pfull = exc_full + kExcSlots;              // [4]  stage-1 partials of a tile complete
    uint64_t* pfree = pfull + kPtSlots;                  // [4]  stage 2 has consumed them
    T* exc_s = (T*)(smem_raw + 128);                     // [8][T][Gp]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    V* ft = (V*)(exc_s + kExcSlots * a.T * a.Gp);        // [nwarps][kTile][kFtPitch]
    V* pt = ft + nwarps * kTile * kFtPitch;              // [4][kTile][ptp]
    const int pt_tile = kTile * a.ptp;

    const int n = a.n;
    const int64_t m0 = (int64_t)blockIdx.x * a.G;
    const int gcount = (int)min((int64_t)a.G, a.M - m0);
    const int ntiles = (a.K + a.T - 1) / a.T;

    if (tid == 0) {
        for (int i = 0; i < kExcSlots; ++i) mbar_init(&exc_full[i], 1);
        for (int i = 0; i < kPtSlots; ++i) {
            mbar_init(&pfull[i], nwarps);
            mbar_init(&pfree[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // muscles beyond the batch read 0 excitation (never recruited); idle lanes contribute 0 force
    for (int i = tid; i < kExcSlots * a.T * a.Gp; i += blockDim.x) exc_s[i] = T(0);
    {
        T* z = (T*)ft;
        for (int i = tid; i < nwarps * kTile * kFtPitch * U; i += blockDim.x) z[i] = T(0);
    }
    __syncthreads();
    if (a.exc_mode == 2) {                                // constant excitation: fill slot 0 once
        const T* src = (const T*)a.exc + m0;
        for (int i = tid; i < a.T * a.Gp; i += blockDim.x) {
            const int g = i % a.Gp;
            if (g < gcount) exc_s[i] = src[g];
        }
        __syncthreads();
    } else if (warp == 0) {
        for (int t = 0; t < min(2, ntiles); ++t) publish_exc<T>(a, exc_s, exc_full, t, m0, gcount, lane);
    }

    const int tq = tid / a.ns;
    const int u_raw = tid - tq * a.ns;
    const bool has_unit = tq < a.q && u_raw < n;
    const int u = has_unit ? u_raw : 0;
    Core core;
    core.load(a, u);

    T F[U];
    bool valid[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
        const int g = tq * U + j;
        valid[j] = has_unit && g < gcount;
        const int64_t idx = (m0 + g) * n + u;
        const double c = valid[j] ? a.cpf[idx] : core.rest_capacity();
        const double d0 = (valid[j] && Core::kCentral) ? a.dur[idx] : 0.0;
        core.init(j, a, c, d0);
        F[j] = 0;
    }

    // stage 2: totals of tile `t` from the partials PT[t & 3]; executed by one whole warp
    auto stage2 = [&](int t) {
        const int pb = t & (kPtSlots - 1);
        mbar_wait(&pfull[pb], (uint32_t)((t / kPtSlots) & 1));
        if (a.totals) {
            const int k0 = t * a.T;
            const int steps = min(a.T, a.K - k0);
            const T* pp = (const T*)(pt + pb * pt_tile);
            T* totals = (T*)a.totals;
            for (int idx = lane; idx < steps * a.G; idx += 32) {          // idx = s*G + g, g = tqo*U + j
                const int s = idx / a.G, g = idx - s * a.G;
                const int tqo = g / U, j = g - tqo * U;
                const T* src = pp + (s * a.ptp + tqo * a.npg) * U + j;
                T acc = src[0];
                for (int q2 = 1; q2 < a.npg; ++q2) acc += src[q2 * U];
                if (g < gcount) totals[(int64_t)(k0 + s) * a.totals_stride_k + m0 + g] = core.out_scale(acc);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&pfree[pb]);
    };

    V* ftw = ft + warp * (kTile * kFtPitch);              // this warp's force tile
    V* const fwr = ftw + lane;                            // compute: row s at fwr + s*kFtPitch
    const V* const frd = ftw + (lane & 7) * kFtPitch + (lane >> 3) * 8;   // stage 1: row lane&7, part lane>>3
    const int pt_col = warp * 4 + (lane >> 3);
    const bool part_used = pt_col < a.q * a.npg;

    for (int tile = 0; tile < ntiles; ++tile) {
        const int k0 = tile * a.T;
        const int steps = min(a.T, a.K - k0);
        const int slot = (a.exc_mode == 2) ? 0 : (tile & (kExcSlots - 1));
        if (a.exc_mode != 2) {
            if (warp == 0 && tile + 2 < ntiles) publish_exc<T>(a, exc_s, exc_full, tile + 2, m0, gcount, lane);
            mbar_wait(&exc_full[slot], (uint32_t)((tile / kExcSlots) & 1));
        }
        if (has_unit) {
            core.begin_tile();
            const T* es = exc_s + slot * a.T * a.Gp + tq * U;
            V* fp = fwr;
#pragma unroll 2
            for (int s = 0; s < steps; ++s, es += a.Gp, fp += kFtPitch) {
                const V e = *(const V*)es;
                bool write_now = false;
                int64_t dec_base = 0;
                if (FORCES) {
                    const int k = k0 + s;
                    write_now = ((k + 1) % a.forces_every) == 0;
                    dec_base = (int64_t)(k / a.forces_every) * a.M * n;
                }
                V fv;
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    bool on;
                    F[j] = core.step(j, e.v[j], on);
                    fv.v[j] = F[j];
                    if (FORCES) {
                        if (write_now && valid[j]) {
                            const int64_t idx = dec_base + (m0 + tq * U + j) * n + u;
                            ((T*)a.forces_dec)[idx] = F[j];
                            if (a.mask_dec) a.mask_dec[idx] = on ? 1 : 0;
                        }
                    }
                }
                *fp = fv;
            }
            if ((tile & 3) == 3) core.end_tile();
        }
        __syncwarp();
        // ---- stage 1: this warp's 4 column parts x 8 step rows -> PT[tile & 3] ----------
        const int pb = tile & (kPtSlots - 1);
        if (tile >= kPtSlots) mbar_wait(&pfree[pb], (uint32_t)((tile / kPtSlots - 1) & 1));
        {
            V acc = frd[0];
#pragma unroll
            for (int i = 1; i < 8; ++i) {
                const V v = frd[i];
#pragma unroll
                for (int j = 0; j < U; ++j) acc.v[j] += v.v[j];
            }
            if (part_used) pt[pb * pt_tile + (lane & 7) * a.ptp + pt_col] = acc;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&pfull[pb]);
        // ---- stage 2 of the previous tile, by one warp in turn ------------------------------
        if (tile > 0 && warp == (tile - 1) % nwarps) stage2(tile - 1);
    }
    if (warp == (ntiles - 1) % nwarps) stage2(ntiles - 1);

#pragma unroll
    for (int j = 0; j < U; ++j) {
        if (!valid[j]) continue;
        const int64_t idx = (m0 + tq * U + j) * n + u;
        core.store(j, a, idx);
        if (a.forces_last) ((T*)a.forces_last)[idx] = F[j];
    }
}

template <int FIB, int CENTRAL, int U, bool FORCES>
__global__ void __launch_bounds__(480, 2) k2_fused_f32(const K2Args a) {
    k2_body<CoreF32<FIB, CENTRAL, U>, U, FORCES>(a);
}

template <int FIB, int CENTRAL, int U, bool FORCES>
__global__ void __launch_bounds__(480, 1) k2_fused_f64(const K2Args a) {
    k2_body<CoreF64<FIB, CENTRAL, U>, U, FORCES>(a);
}

}  // namespace mb
