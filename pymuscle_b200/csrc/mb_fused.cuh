// mb_fused.cuh -- K2: K fused steps per launch, unit state in registers.
//
// Uniform batch: M muscles x n units.  A block owns G = q*U consecutive muscles.
// Muscle group tq (U muscles) occupies `ns` consecutive threads, ns = n rounded up to a
// multiple of 8; thread tq*ns + u handles unit u of muscles tq*U .. +U-1: its U
// unit-instances share one set of per-unit parameters (registers, loaded once per window
// as 128-bit vectors) and give U independent dependency chains.  Per step: one vector LDS
// (U excitations), the arithmetic below, one vector STS (U forces).
//
// FP32 mode packs two instances into one 64-bit register pair and uses the sm_100 packed
// FP32 instructions (FFMA2 / FMUL2 / FADD2, one issue slot for two lanes of FMA-pipe work);
// compares, min/max and selects stay scalar (ALU pipe), ex2 is one MUFU per instance-step.
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
//   stage 2  one warp per tile (rotating) waits on the mbarrier that counts the warps'
//            stage-1 arrivals, sums the ns/8 partials of every (step, muscle) and stores the
//            totals to HBM (coalesced).
//
// PT is a ring of 4 tiles guarded by mbarriers (pfull: 1 arrival per warp; pfree: 1 arrival
// by the stage-2 warp), so warps may drift up to 4 tiles apart and never wait for each
// other in steady state.  Warp 0 additionally stages the per-step excitations kExcAhead
// tiles ahead with cp.async.bulk (1-D bulk copy per step row, mbarrier complete_tx) into a
// 16-slot ring; consumers wait on the slot's mbarrier.
//
// HBM traffic per window: state in/out (32 B/unit) plus excitations and totals
// (2*sizeof(T)/n B per unit-step) -- see DESIGN.md.
#pragma once
#include "mb_core.cuh"
#include "mb_async.cuh"

namespace mb {

constexpr int kMaxPeers = 8;     // destinations of the fused totals gather besides `totals`
constexpr int kExcSlots = 16;    // excitation ring (tiles)
constexpr int kExcAhead = 6;     // tiles staged ahead of the staging warp's own position (> PT drift bound)
constexpr int kPtSlots = 4;      // partial-sum ring (tiles)
constexpr int kTile = 8;         // steps per tile (stage 1 maps lane & 7 to the step)
constexpr int kFtPitch = 33;     // row pitch of a warp's force tile, in vectors (odd => conflict free)

struct K2Args {
    const void* params;
    int64_t param_stride;        // 0: one shared table; else 128-bit vectors between the tables of consecutive muscles
    int64_t M;
    int32_t n, ns, q, K, T, G, Gp, npg, ptp;   // ns: threads per muscle group; npg = ns/8 parts per group; ptp: PT row pitch
    const void* exc;
    int64_t exc_stride_k;
    int32_t exc_mode;            // 0 = bulk copy, 1 = element-wise loads, 2 = constant over the window
    int32_t exc_hold;            // h >= 1: step k reads excitation row k / h (a control signal held for h steps)
    int32_t totals_every;        // t >= 1: only the totals of steps t-1, 2t-1, ... are stored, at row k / t
    double* cpf;
    double* dur;
    void* totals;
    int64_t totals_stride_k;
    void* totals_peer[kMaxPeers];  // further copies of the totals (same stride): peer GPUs' gather buffers, written
    int32_t n_peers;               // straight from the kernel over NVLink (the all-gather fused into the epilogue)
    int64_t peer_stride_k, peer_stride_pair, peer_stride_odd;   // peer element (k, m) at (m >> 1)*stride_pair + k*stride_k + (m & 1)*stride_odd
    int32_t peer_multicast;        // 1: totals_peer[0] is an NVLink MULTICAST address (NVSwitch replicates one store to every GPU)
    void* forces_dec;
    uint8_t* mask_dec;
    int32_t forces_every;
    void* forces_last;
    int32_t peripheral;
    double dt, adk_d, max_dur;
    double decay_d;              // exp(-dt/tau): per-step factor of exp(-dur/tau) for a firing unit
    float dtk;                   // (float)(dt * adk_d)
    float decay, dm1;            // (float)decay_d and decay - 1 (exact in float)
    ScalF sf;
    ScalD sd;
    ExpConsts ec;                // exp_neg_fast's constants (FP64 mode)
};

// Ragged layouts (rows of several muscles): the kernels that take one muscle per block / per warp address muscle
// m = (row, segment) through this
struct K2xArgs {
    K2Args k;                      // M = rows * S muscles; n unused (per-muscle lengths come from seg_off)
    int64_t L;                     // units per row
    int64_t total_units;           // rows * L
    int32_t S;
    const int32_t* seg_off;        // [S + 1] or NULL (one muscle of L units per row)
    const double* seg_div;         // [S] output divisors or NULL (k.sf.inv_out)
};

// One total into a peer GPU's gather buffer.  `mc`: the address is a multicast mapping (NVLS) -- one multimem.st leaves
// this GPU and the NVSwitch delivers it to every GPU of the group, instead of one store per peer.
__device__ __forceinline__ void peer_store(float* p, float v, int mc) {
    if (mc) asm volatile("multimem.st.weak.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
    else *p = v;
}
__device__ __forceinline__ void peer_store(double* p, double v, int mc) {
    if (mc) asm volatile("multimem.st.weak.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
    else *p = v;
}
__device__ __forceinline__ void peer_store2(float* p, float v0, float v1, int mc) {      // p 8-byte aligned
    if (mc) asm volatile("multimem.st.weak.global.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v0), "f"(v1) : "memory");
    else *(float2*)p = make_float2(v0, v1);
}

// U-wide vector of T in shared memory (one LDS/STS of 4*U or 8*U bytes)
template <typename T, int U> struct alignas(sizeof(T) * U) VecU { T v[U]; };


// ======================================================================================
// Per-thread arithmetic of the two precision modes.  Each "core" holds the unit's
// parameters and the state of its U instances in registers.
// ======================================================================================

// ---- FP32 mode ------------------------------------------------------------------------
// Work is done on PAIRS: two unit-instances share one 64-bit register pair and the FMA-pipe
// work is issued as packed FFMA2 / FMUL2 / FADD2.  A pair is either two muscles of the same
// unit (parameters are scalars, broadcast operands: PT = float) or two units of the same
// muscle (parameters are pairs: PT = float2).
//
// Capacity is carried as hi + lo (two floats): `hi` is frozen between renormalisations,
// per-step decrements accumulate in the small `lo` (so their rounding error is relative to
// the accumulated decrement, not to the capacity) and the pair is renormalised (error-free
// TwoSum) every 32 steps.  On-duration is dur0 + count*dt with an exact float step
// counter.  Plain float accumulators miss rtol 1e-4 over 10,000 steps.
//
// CENTRAL: 0 = pool.apply_fatigue off (adaptation is the identity)
//          1 = on, fast path: the launcher has checked that (a) the adaptation of a
//              recruited unit cannot be negative and (b) max_duration/tau is so large
//              that the duration clamp cannot change exp(-dur/tau); both clamps drop out.
//              nD = exp(-dur/tau) - 1 is carried by recurrence -- a firing step maps it to
//              nD*decay + (decay - 1) -- and re-synchronised with one ex2 per tile from the
//              exact counter, which replaces the per-step MUFU of the adaptation curve
//          2 = on, general path with both clamps (pool:202-206, :221)
__device__ __forceinline__ float2 B2(float v) { return make_float2(v, v); }
__device__ __forceinline__ float2 B2(float2 v) { return v; }
__device__ __forceinline__ float2 N2(float v) { return make_float2(-v, -v); }
__device__ __forceinline__ float2 N2(float2 v) { return make_float2(-v.x, -v.y); }
__device__ __forceinline__ float LX(float v) { return v; }
__device__ __forceinline__ float LX(float2 v) { return v.x; }
__device__ __forceinline__ float LY(float v) { return v; }
__device__ __forceinline__ float LY(float2 v) { return v.y; }

template <typename PT>
struct PairParams {            // what a step needs of the unit(s) of a pair
    PT crit, thr_g, peak, phir, phir6, ctA, ctB, fatdt, rdidt, ptf, ptf_lo;
};
struct PairState {
    float2 hi, lo, cpf, cnt, nD, room, arg0;
};
constexpr float kLinTop = 0.12014662085535621f;   // 1 - exp(-2 * 0.4^3): the two pieces of the force curve meet here

struct PairScal {              // per-launch scalars
    float gain_e, c25k, ythr, argmin, dtk, decay, dm1;
};

__device__ __forceinline__ float& el(float2& v, int k) { return k ? v.y : v.x; }

template <typename PT>
__device__ __forceinline__ void pair_init(PairState& st, const PairParams<PT>& p, int k, double c, double d0, double adk_d) {
    const float h = (float)c;
    const float l = (float)(c - (double)h);
    const float ptf = k ? LY(p.ptf) : LX(p.ptf), ptf_lo = k ? LY(p.ptf_lo) : LX(p.ptf_lo);
    el(st.hi, k) = h;
    el(st.lo, k) = l;
    el(st.cpf, k) = h + l;
    el(st.room, k) = (ptf - h) + ptf_lo;                 // peak - hi: the upper bound of lo (pmf:104-105)
    el(st.cnt, k) = 0.0f;
    el(st.nD, k) = 0.0f;
    el(st.arg0, k) = (float)(d0 * adk_d);
}

template <int CENTRAL>
__device__ __forceinline__ void pair_begin_tile(PairState& st, const PairScal& sc) {
    if (CENTRAL == 1) {
        st.nD.x = ex2_approx(fmaf(st.cnt.x, sc.dtk, st.arg0.x)) - 1.0f;
        st.nD.y = ex2_approx(fmaf(st.cnt.y, sc.dtk, st.arg0.y)) - 1.0f;
    }
}

// One step of a pair; x = the two excitations; returns the two unit forces.
// Without adaptation in the loop (CENTRAL != 1) recruitment enters the arithmetic as a float
// mask onf (1 = recruited): the firing rate is multiplied by it and the recovery term is blended
// with it, so the FMA-pipe work stays packed and the only scalar (ALU-pipe) instructions are the
// two compares, the peak clamp, the curve select and the capacity clamp.
template <int FIB, int CENTRAL, typename PT>
__device__ __forceinline__ float2 pair_step(PairState& st, const PairParams<PT>& p, const PairScal& sc, float2 x,
                                            bool& on_a, bool& on_b) {
    // ---- pool: rate and recruitment (pool:150-179) ----
    float2 r = __ffma2_rn(x, B2(sc.gain_e), N2(p.thr_g));
    float2 onf;
    onf.x = (x.x >= LX(p.crit)) ? 1.0f : 0.0f;                              // pool:171 (exact mask: see mb_params_pack)
    onf.y = (x.y >= LY(p.crit)) ? 1.0f : 0.0f;
    r.x = fminf(r.x, LX(p.peak));
    r.y = fminf(r.y, LY(p.peak));
    float2 fr;
    if (CENTRAL == 0) {
        fr = __fmul2_rn(r, onf);                                            // pool:172
    } else if (CENTRAL == 1) {
        // measured on B200: with the adaptation recurrences in the loop the FMA pipe is the busiest
        // unit, so here the recruitment mask is used as a predicate (scalar, predicated updates)
        // rather than blended arithmetically (which costs two more FMA-pipe operations per step)
        const float2 q = __ffma2_rn(r, B2(p.phir), N2(p.phir6));            // pool:231-232
        fr = __ffma2_rn(q, st.nD, r);                                       // r - q*(1 - exp(-dur/tau)), dur BEFORE the update
        const bool oa = x.x >= LX(p.crit), ob = x.y >= LY(p.crit);
        fr.x = oa ? fr.x : 0.0f;                                            // pool:171-172
        fr.y = ob ? fr.y : 0.0f;
        if (oa) {                                                           // pool:196-197
            st.cnt.x += 1.0f;
            st.nD.x = fmaf(st.nD.x, sc.decay, sc.dm1);
        }
        if (ob) {
            st.cnt.y += 1.0f;
            st.nD.y = fmaf(st.nD.y, sc.decay, sc.dm1);
        }
    } else {
        const bool oa = onf.x != 0.0f, ob = onf.y != 0.0f;
        const float2 arg = __ffma2_rn(st.cnt, B2(sc.dtk), st.arg0);
        st.cnt = __fadd2_rn(st.cnt, onf);
        const float ra = oa ? r.x : 0.0f, rb = ob ? r.y : 0.0f;
        const float qa = fmaf(ra, LX(p.phir), -LX(p.phir6)), qb = fmaf(rb, LY(p.phir), -LY(p.phir6));
        const float aa = fmaf(-qa, ex2_approx(fmaxf(arg.x, sc.argmin)), qa);            // pool:217-219
        const float ab = fmaf(-qb, ex2_approx(fmaxf(arg.y, sc.argmin)), qb);
        fr.x = ra - fmaxf(aa, 0.0f);                                        // pool:221, :121
        fr.y = rb - fmaxf(ab, 0.0f);
    }
    on_a = onf.x != 0.0f;
    on_b = onf.y != 0.0f;
    // ---- fibers (fibers:211-259) on ny = -kappa*x: exp(-2x^3) = ex2(ny^3) ----
    const float2 ctkn = __ffma2_rn(st.cpf, B2(p.ctB), N2(p.ctA));           // -kappa*ct_cur/1000
    const float2 ny = __fmul2_rn(ctkn, fr);
    const float2 lin = __fmul2_rn(ny, B2(-sc.c25k));
    const float2 y2 = __fmul2_rn(ny, ny);
    const float2 ny3 = __fmul2_rn(y2, ny);
    float2 ex;
    ex.x = ex2_approx(ny3.x);
    ex.y = ex2_approx(ny3.y);
    const float2 sig = __ffma2_rn(ex, B2(-1.0f), B2(1.0f));
    // piecewise curve without a predicate: below x = 0.4 the line lies above the sigmoid and below its
    // value at 0.4 (kLinTop = 1 - exp(-2*0.4^3), where the pieces meet); above, the sigmoid is the larger
    float2 nf;
    nf.x = fmaxf(fminf(lin.x, kLinTop), sig.x);                             // fibers:233-246
    nf.y = fmaxf(fminf(lin.y, kLinTop), sig.y);
    const float2 F = __fmul2_rn(nf, st.cpf);                                // fibers:258 (capacity BEFORE the update)
    // ---- capacity update on the compensated pair (fibers:110-114, pmf:94-105) ----
    float2 l = __ffma2_rn(nf, N2(p.fatdt), st.lo);
    if (FIB == MB_FIBERS_PYMUSCLE) {
        // recovery of a unit that does not fire (nf <= 0  <=>  not recruited):
        // l += (peak - capacity) * (1 - onf) * rdidt                        pmf:121-141
        float2 firing = onf;
        if (CENTRAL == 2) {                                                 // general parameters: test nf itself
            firing.x = (nf.x > 0.0f) ? 1.0f : 0.0f;
            firing.y = (nf.y > 0.0f) ? 1.0f : 0.0f;
        }
        const float2 gap = __fadd2_rn(st.room, N2(st.lo));
        const float2 gn = __ffma2_rn(gap, N2(firing), gap);
        l = __ffma2_rn(gn, B2(p.rdidt), l);
    }
    l.x = fmaxf(l.x, -st.hi.x);                                             // fibers:114
    l.y = fmaxf(l.y, -st.hi.y);
    // (the upper clamp at the peak, pmf:104-105, is applied in pair_renormalise(): recovery approaches
    //  the peak from below and can overshoot it only by rounding)
    st.lo = l;
    st.cpf = __fadd2_rn(st.hi, l);
    return F;
}

// renormalise (TwoSum) so that |lo| <= ulp(hi)/2 again; applies the upper clamp of PyMuscle fibers
template <int FIB, typename PT>
__device__ __forceinline__ void pair_renormalise(PairState& st, const PairParams<PT>& p) {
    if (FIB == MB_FIBERS_PYMUSCLE) {                                        // pmf:104-105
        st.lo.x = fminf(st.lo.x, st.room.x);
        st.lo.y = fminf(st.lo.y, st.room.y);
    }
    const float2 s1 = __fadd2_rn(st.hi, st.lo);
    const float2 bb = __fadd2_rn(s1, N2(st.hi));
    const float2 t1 = __fadd2_rn(s1, N2(bb));
    const float2 e1 = __fadd2_rn(st.hi, N2(t1));
    const float2 e2 = __fadd2_rn(st.lo, N2(bb));
    st.hi = s1;
    st.lo = __fadd2_rn(e1, e2);
    st.room = __fadd2_rn(__fadd2_rn(B2(p.ptf), N2(s1)), B2(p.ptf_lo));
}

// write element k of a pair back to the float64 state
template <bool CENTRAL>
__device__ __forceinline__ void pair_store(PairState& st, int k, const K2Args& a, int64_t idx) {
    if (a.peripheral) {
        // add the window's change to the float64 state: an untouched unit keeps its
        // exact bits, and the low bits of the incoming capacity are not truncated
        const double c0 = a.cpf[idx];
        const float h0 = (float)c0;
        const float l0 = (float)(c0 - (double)h0);
        a.cpf[idx] = c0 + (((double)el(st.hi, k) - (double)h0) + ((double)el(st.lo, k) - (double)l0));
    }
    if (CENTRAL) {
        double d = fma((double)el(st.cnt, k), a.dt, a.dur[idx]);
        d = (d > a.max_dur) ? a.max_dur : d;
        a.dur[idx] = (d < 0.0) ? 0.0 : d;
    }
}

// write element k of a pair back WITHOUT reading the old state (persistent short-window kernel): the capacity is the
// exact sum of the float pair (48 bits), except that a unit sitting at its peak -- rested, or recovered up to the clamp
// pmf:104-105 -- gets the exact float64 peak; the on-duration is advanced by a fire-and-forget atomic add of count * dt
// (no clamp can trigger far below max_duration; near it, the slow read-modify-write path runs).
template <bool CENTRAL, typename PT>
__device__ __forceinline__ void pair_store_direct(PairState& st, const PairParams<PT>& p, int k, const K2Args& a, int64_t idx,
                                                  float ptf_lo2) {
    if (a.peripheral) {
        const float hi = el(st.hi, k), lo = el(st.lo, k);
        const float ptf = k ? LY(p.ptf) : LX(p.ptf), ptf_lo = k ? LY(p.ptf_lo) : LX(p.ptf_lo);
        const double exact_peak = ((double)ptf + (double)ptf_lo) + (double)ptf_lo2;
        a.cpf[idx] = (hi == ptf && lo == ptf_lo) ? exact_peak : (double)hi + (double)lo;
    }
    if (CENTRAL) {
        const float cnt = el(st.cnt, k);
        if (cnt > 0.0f) {
            const double add = (double)cnt * a.dt;
            const double d0_est = (double)el(st.arg0, k) / a.adk_d;            // start value to float accuracy
            if (d0_est + add < 0.99 * a.max_dur) {
                atomicAdd(&a.dur[idx], add);                                   // (result unused: RED.ADD.F64)
            } else {
                double d = fma((double)cnt, a.dt, a.dur[idx]);
                a.dur[idx] = (d > a.max_dur) ? a.max_dur : d;                  // pool:202-203
            }
        }
    }
}

__device__ __forceinline__ PairScal pair_scal(const K2Args& a) {
    PairScal s;
    s.gain_e = a.sf.gain_e; s.c25k = a.sf.c25k; s.ythr = a.sf.ythr; s.argmin = a.sf.argmin;
    s.dtk = a.dtk; s.decay = a.decay; s.dm1 = a.dm1;
    return s;
}
__device__ __forceinline__ PairParams<float> pair_params(const UnitF& u, const K2Args& a) {
    PairParams<float> p;
    p.crit = u.crit; p.thr_g = u.thr_g; p.peak = u.peak; p.phir = u.phir; p.phir6 = u.phir6;
    p.ctA = u.ctA; p.ctB = u.ctB; p.ptf = u.ptf; p.ptf_lo = u.ptf_lo;
    p.fatdt = a.peripheral ? (float)((double)u.fat * a.dt) : 0.0f;
    p.rdidt = a.peripheral ? (float)((double)u.rdi * a.dt) : 0.0f;
    return p;
}

// U instances (muscles) of ONE unit: pairs of muscles, scalar parameters
template <int FIB, int CENTRAL, int U>
struct CoreF32 {
    using T = float;
    using V = VecU<float, U>;
    static constexpr bool kCentral = CENTRAL != 0;
    static constexpr int NP = (U + 1) / 2;           // pairs
    UnitF u;
    PairParams<float> p;
    PairScal sc;
    float inv_out;
    PairState st[NP];

    __device__ __forceinline__ void load(const K2Args& a, int unit, int64_t muscle) {
        u = load_unit((const float4*)a.params + muscle * a.param_stride, a.n, unit);
        p = pair_params(u, a);
        sc = pair_scal(a);
        inv_out = a.sf.inv_out;
#pragma unroll
        for (int j = 0; j < 2 * NP; ++j) init(j, a, rest_capacity(), 0.0);   // the idle half of an odd U
    }
    __device__ __forceinline__ void disable() { p.crit = __int_as_float(0x7f800000); }   // never recruited, force 0
    __device__ __forceinline__ double rest_capacity() const { return unit_peak(u); }
    __device__ __forceinline__ void init(int j, const K2Args& a, double c, double d0) {
        pair_init(st[j >> 1], p, j & 1, c, d0, a.adk_d);
    }
    __device__ __forceinline__ void begin_tile() {
#pragma unroll
        for (int pi = 0; pi < NP; ++pi) pair_begin_tile<CENTRAL>(st[pi], sc);
    }
    __device__ __forceinline__ float2 step_pair(int pi, float2 x, bool& on_a, bool& on_b) {
        return pair_step<FIB, CENTRAL>(st[pi], p, sc, x, on_a, on_b);
    }
    // one step of all U instances: e -> f (forces); onbits = recruitment mask of the instances
    __device__ __forceinline__ void step_vec(const V& e, V& f, unsigned& onbits) {
        onbits = 0;
#pragma unroll
        for (int pi = 0; pi < NP; ++pi) {
            const int j0 = 2 * pi, j1 = (2 * pi + 1 < U) ? 2 * pi + 1 : 2 * pi;
            bool oa, ob;
            const float2 F = step_pair(pi, make_float2(e.v[j0], e.v[j1]), oa, ob);
            f.v[j0] = F.x;
            if (2 * pi + 1 < U) f.v[j1] = F.y;
            onbits |= (oa ? 1u : 0u) << j0;
            if (2 * pi + 1 < U) onbits |= (ob ? 1u : 0u) << j1;
        }
    }
    __device__ __forceinline__ void set_table(const double*) {}
    __device__ __forceinline__ void renormalise() {
#pragma unroll
        for (int pi = 0; pi < NP; ++pi) pair_renormalise<FIB>(st[pi], p);
    }
    __device__ __forceinline__ void store(int j, const K2Args& a, int64_t idx) {
        pair_store<CENTRAL != 0>(st[j >> 1], j & 1, a, idx);
    }
    __device__ __forceinline__ void store_direct(int j, const K2Args& a, int64_t idx, float ptf_lo2) {
        pair_store_direct<CENTRAL != 0>(st[j >> 1], p, j & 1, a, idx, ptf_lo2);
    }
    __device__ __forceinline__ float out_scale(float v) const { return v * inv_out; }
};

// ---- FP64 mode ------------------------------------------------------------------------
// The adaptation factor E = exp(-dur/tau) is carried by recurrence: a firing step multiplies it
// by exp(-dt/tau).  General path (CENTRAL == 2): durations are accumulated exactly as the
// reference does (dur += dt, clamped), a clamped duration pins E to exp(-max_duration/tau), and E
// is recomputed with a full-precision exp at every tile start.  Fast path (CENTRAL == 1): see
// begin_tile().
template <int FIB, int CENTRAL, int U>
struct CoreF64 {
    using T = double;
    using V = VecU<double, U>;
    static constexpr bool kCentral = CENTRAL != 0;
    UnitD p;
    ScalD sc;
    double fatdt, rdidt, dt, max_dur, decay, omd /* 1 - decay */, Gmax;
    double c25s;                 // slope of the linear part over cbrt(2)
    double ctA2, ctB2;           // contraction-time factors times cbrt(2): (cbrt(2) x)^3 = 2 x^3 without a multiply
    double cpf[U], dur[U], G[U]; // G = 1 - exp(-dur/tau), the adaptation factor (pool:217-219)
    int cnt[U];                  // CENTRAL == 1: on-steps since the window started (dur[] stays at its start value)
    const double* tab;           // shared-memory copy of kExp2Tab

    __device__ __forceinline__ void load(const K2Args& a, int u, int64_t muscle) {
        p = load_unit((const double2*)a.params + muscle * a.param_stride, a.n, u);
        sc = a.sd;
        fatdt = a.peripheral ? p.fat * a.dt : 0.0;
        rdidt = a.peripheral ? p.rdi * a.dt : 0.0;
        dt = a.dt;
        max_dur = a.max_dur;
        decay = a.decay_d;
        omd = -expm1(a.dt * a.sd.mitau);
        Gmax = CENTRAL ? -expm1(a.max_dur * a.sd.mitau) : 0.0;
        ctA2 = p.ctA * kCbrt2;
        ctB2 = p.ctB * kCbrt2;
        c25s = a.sd.c25 / kCbrt2;
#pragma unroll
        for (int j = 0; j < U; ++j) init(j, a, p.ptf, 0.0);                // instances beyond the batch rest
    }
    __device__ __forceinline__ double rest_capacity() const { return p.ptf; }
    __device__ __forceinline__ void init(int j, const K2Args&, double c, double d0) {
        cpf[j] = c;
        dur[j] = d0;
        cnt[j] = 0;
        G[j] = CENTRAL ? -expm1(d0 * sc.mitau) : 0.0;                      // pool:217-218
    }
    // Fast path: G = 1 - exp(-dur/tau) is carried by the recurrence alone for the whole window -- every on-step
    // maps it to G * decay + (1 - decay), decay = exp(-dt/tau), half an ulp of relative error each, i.e. ~1e-13 after 1,000 steps against the
    // mode's 1e-9 bar -- and the duration is the start value plus an integer count of on-steps (the reference's
    // min(dur + dt, max) per on-step, pool:196-203, is monotone, so the clamp can wait until the state is stored).
    // General path: resynchronised with a full-precision exp at every tile start.
    __device__ __forceinline__ void begin_tile() {
        if (CENTRAL == 2) {
#pragma unroll
            for (int j = 0; j < U; ++j) G[j] = -expm1(dur[j] * sc.mitau);  // pool:217-218
        }
    }
    // CENTRAL: 0 = no adaptation, 2 = general, 1 = fast path (as in FP32 mode): the launcher has
    // checked that the adaptation of a recruited unit cannot be negative, that gain and minimal rate
    // are positive (so "rate > 0" is the recruitment mask), that the duration clamp cannot change
    // exp(-dur/tau) in float64, and that the sigmoid's exponent stays above -700
    __device__ __forceinline__ double rate(int j, double e, bool& on) {
        double r, fr;
        if (CENTRAL == 1) {
            // pool_rate() without zeroing the rate of an unrecruited unit: the select below does it
            const double r0 = __dadd_rn(__dsub_rn(__dmul_rn(e, sc.input_scale), p.thr), sc.min_rate);   // muscle.py:275, pool:169-170
            on = !(r0 < sc.min_rate);                                      // pool:171
            r = fmin(r0 * sc.gain, p.peak);                                // pool:173, :176-177
            const double q = fma(r, p.phir, -p.phir6);                     // pool:231-232
            fr = on ? fma(-q, G[j], r) : 0.0;                              // pool:217-221, :121 (dur BEFORE the update)
            if (on) {
                cnt[j] += 1;
                G[j] = fma(G[j], decay, omd);
            }
        } else {
            r = pool_rate(p, sc, e, on);
            fr = r;
        }
        if (CENTRAL == 2) {
            const double q = fma(r, p.phir, -p.phir6);                     // pool:231-232
            const double ad = q * G[j];                                    // pool:217-219 (dur BEFORE the update)
            fr = r - ((ad < 0.0) ? 0.0 : ad);                              // pool:221, :121
            if (r > 0.0) {                                                 // pool:196-197
                const double dn = dur[j] + dt;
                const bool clamp = dn > max_dur;                           // pool:202-203
                dur[j] = clamp ? max_dur : dn;
                G[j] = clamp ? Gmax : fma(G[j], decay, omd);
            }
        }
        return fr;
    }
    // The U instances of a thread advance in three phases so that their FP64 chains interleave: a branch
    // per instance around the sigmoid's exponential (as in fibers_force_fast) would serialise them -- the
    // scheduler cannot move instructions across the branches, and ncu showed 40 % of the stall samples on
    // fixed-latency dependencies of the lone Horner chain.  Here one branch decides for all U instances, and
    // inside it the U exponentials are straight-line code.
    __device__ __forceinline__ void step_vec(const V& e, V& f, unsigned& onbits) {
        onbits = 0;
        double x[U];
        bool need = false;
#pragma unroll
        for (int j = 0; j < U; ++j) {
            bool on;
            const double fr = rate(j, e.v[j], on);
            onbits |= (on ? 1u : 0u) << j;
            x[j] = fma(cpf[j], -ctB2, ctA2) * fr;                          // cbrt(2) * (fibers:220 with the fatigued contraction time)
            need |= x[j] > kLinThrD * kCbrt2;
        }
        double ex[U];
        if (need) {
#pragma unroll
            for (int j = 0; j < U; ++j) ex[j] = exp_neg_fast<CENTRAL != 1>(-(x[j] * x[j]) * x[j], tab);   // exp(-2 x^3)
        } else {
#pragma unroll
            for (int j = 0; j < U; ++j) ex[j] = 1.0;
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const double nf = (x[j] <= kLinThrD * kCbrt2) ? x[j] * c25s : 1.0 - ex[j];   // fibers:247-252
            f.v[j] = nf * cpf[j];                                                  // fibers:258
            cpf[j] = fatigue_update<FIB == MB_FIBERS_PYMUSCLE>(cpf[j], nf, fatdt, rdidt, p.ptf);
        }
    }
    __device__ __forceinline__ void renormalise() {}
    __device__ __forceinline__ void set_table(const double* t) { tab = t; }
    __device__ __forceinline__ void store(int j, const K2Args& a, int64_t idx) {
        if (a.peripheral) a.cpf[idx] = cpf[j];
        if (CENTRAL == 1) a.dur[idx] = fmin(fma((double)cnt[j], dt, dur[j]), max_dur);       // pool:196-203
        if (CENTRAL == 2) a.dur[idx] = dur[j];
    }
    __device__ __forceinline__ double out_scale(double v) const { return v / sc.out_div; }   // muscle.py:278
};

// ======================================================================================
// excitation staging (warp 0): tile `tile` of the window -> ring slot tile & 15
// ======================================================================================
template <typename T>
__device__ __forceinline__ void publish_exc(const K2Args& a, T* exc_s, uint32_t exc_bar0, int tile, int64_t m0,
                                            int gcount, int lane) {
    const int k0 = tile * a.T;
    const int steps = min(a.T, a.K - k0);
    const int slot = tile & (kExcSlots - 1);
    T* dst = exc_s + slot * a.T * a.Gp;
    const T* src = (const T*)a.exc + m0;                  // row of step k: k / exc_hold
    const int hold = a.exc_hold;
    const uint32_t bar = exc_bar0 + 8u * slot;
    if (a.exc_mode == 0) {
        if (lane == 0) {
            const uint32_t row_bytes = (uint32_t)(gcount * sizeof(T));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar, row_bytes * steps);
            const uint32_t d0 = smem_u32(dst);
            for (int s = 0; s < steps; ++s)
                bulk_g2s(d0 + s * a.Gp * (int)sizeof(T), src + (int64_t)((k0 + s) / hold) * a.exc_stride_k, row_bytes, bar);
        }
    } else {
        for (int i = lane; i < steps * a.G; i += 32) {
            const int s = i / a.G, g = i - s * a.G;
            if (g < gcount) dst[s * a.Gp + g] = __ldg(src + (int64_t)((k0 + s) / hold) * a.exc_stride_k + g);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar);
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
    // mbarriers: exc_full[16] | pfull[4] (stage-1 partials of a tile complete) | pfree[4] (stage 2 consumed them)
    const uint32_t exc_bar0 = smem_u32(smem_raw);
    const uint32_t pfull0 = exc_bar0 + 8u * kExcSlots;
    const uint32_t pfree0 = pfull0 + 8u * kPtSlots;
    T* exc_s = (T*)(smem_raw + 256);                     // [16][T][Gp]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    V* ft = (V*)(exc_s + kExcSlots * a.T * a.Gp);        // [nwarps][kTile][kFtPitch]
    V* pt = ft + nwarps * kTile * kFtPitch;              // [4][kTile][ptp]
    const int pt_tile = kTile * a.ptp;

    const int n = a.n;
    const int64_t m0 = (int64_t)blockIdx.x * a.G;
    const int gcount = (int)min((int64_t)a.G, a.M - m0);
    const int ntiles = (a.K + a.T - 1) / a.T;
    __shared__ double exp2_tab[32];                       // for exp_neg_fast (FP64 mode)
    if (tid < 32) exp2_tab[tid] = kExp2Tab[tid];

    if (tid == 0) {
        for (int i = 0; i < kExcSlots; ++i) mbar_init(exc_bar0 + 8u * i, 1);
        for (int i = 0; i < kPtSlots; ++i) {
            mbar_init(pfull0 + 8u * i, nwarps);
            mbar_init(pfree0 + 8u * i, 1);
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
        for (int t = 0; t < min(kExcAhead, ntiles); ++t) publish_exc<T>(a, exc_s, exc_bar0, t, m0, gcount, lane);
    }

    const int tq = tid / a.ns;
    const int u_raw = tid - tq * a.ns;
    const bool has_unit = tq < a.q && u_raw < n;
    const int u = has_unit ? u_raw : 0;
    Core core;
    core.set_table(exp2_tab);
    core.load(a, u, min(m0 + (int64_t)tq * U, a.M - 1));     // (per-instance tables: U == 1, the muscle's own table)
#pragma unroll
    for (int j = 0; j < U; ++j) {
        const int g = tq * U + j;
        if (has_unit && g < gcount) {
            const int64_t idx = (m0 + g) * n + u;
            MB_ASSERT(idx < a.M * (int64_t)n && u < n);
            core.init(j, a, a.cpf[idx], Core::kCentral ? a.dur[idx] : 0.0);
        }
    }

    // stage 2: totals of tile `t` from the partials PT[t & 3]; executed by one whole warp
    auto stage2 = [&](int t) {
        const int pb = t & (kPtSlots - 1);
        mbar_wait(pfull0 + 8u * pb, (uint32_t)((t / kPtSlots) & 1));
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
#pragma unroll 5
                for (int q2 = 1; q2 < a.npg; ++q2) acc += src[q2 * U];
                const int k = k0 + s;
                if (g < gcount && (a.totals_every == 1 || (k + 1) % a.totals_every == 0)) {
                    const int64_t krow = k / a.totals_every;
                    const int64_t o = krow * a.totals_stride_k + m0 + g;
                    const T v = core.out_scale(acc);
                    totals[o] = v;
                    if (a.n_peers > 0) {                               // peer GPUs (NVLink stores)
                        const int64_t mg = m0 + g;
                        const int64_t op = (mg >> 1) * a.peer_stride_pair + krow * a.peer_stride_k + (mg & 1) * a.peer_stride_odd;
                        for (int pr = 0; pr < a.n_peers; ++pr) peer_store((T*)a.totals_peer[pr] + op, v, a.peer_multicast);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(pfree0 + 8u * pb);
    };

    V* const fwr = ft + warp * (kTile * kFtPitch) + lane;                                   // compute: row s at fwr + s*kFtPitch
    const V* const frd = ft + warp * (kTile * kFtPitch) + (lane & 7) * kFtPitch + (lane >> 3) * 8;   // stage 1
    V* const pwr = pt + (lane & 7) * a.ptp + warp * 4 + (lane >> 3);                        // + pb*pt_tile
    const bool part_used = warp * 4 + (lane >> 3) < a.q * a.npg;
    const T* const es0 = exc_s + tq * U;
    int s2warp = 0;                                       // the warp that runs stage 2 of tile-1 this iteration: (tile-1) % nwarps

    for (int tile = 0; tile < ntiles; ++tile) {
        const int k0 = tile * a.T;
        const int steps = min(a.T, a.K - k0);
        int slot = 0;
        if (a.exc_mode != 2) {
            slot = tile & (kExcSlots - 1);
            if (warp == 0 && tile + kExcAhead < ntiles)
                publish_exc<T>(a, exc_s, exc_bar0, tile + kExcAhead, m0, gcount, lane);
            mbar_wait(exc_bar0 + 8u * slot, (uint32_t)((tile / kExcSlots) & 1));
        }
        if (has_unit) {
            core.begin_tile();
            int fphase = FORCES ? k0 % a.forces_every : 0;        // steps since the last captured one
            const T* es = es0 + slot * a.T * a.Gp;
            V* fp = fwr;
#pragma unroll 2
            for (int s = 0; s < steps; ++s, es += a.Gp, fp += kFtPitch) {
                const V e = *(const V*)es;
                V fv;
                unsigned onbits;
                core.step_vec(e, fv, onbits);
                *fp = fv;
                if (FORCES) {
                    const int k = k0 + s;
                    if (++fphase == a.forces_every) {             // every forces_every-th step, counted instead of k % f
                        fphase = 0;
                        const int64_t dec_base = (int64_t)(k / a.forces_every) * a.M * n;
#pragma unroll
                        for (int j = 0; j < U; ++j) {
                            if (tq * U + j < gcount) {
                                const int64_t idx = dec_base + (m0 + tq * U + j) * n + u;
                                ((T*)a.forces_dec)[idx] = fv.v[j];
                                if (a.mask_dec) a.mask_dec[idx] = (uint8_t)((onbits >> j) & 1u);
                            }
                        }
                    }
                }
            }
            if ((tile & 3) == 3 || tile == ntiles - 1) core.renormalise();
        }
        __syncwarp();
        // ---- stage 1: this warp's 4 column parts x 8 step rows -> PT[tile & 3] ----------
        const int pb = tile & (kPtSlots - 1);
        if (tile >= kPtSlots) mbar_wait(pfree0 + 8u * pb, (uint32_t)((tile / kPtSlots - 1) & 1));
        {
            V acc = frd[0];
#pragma unroll
            for (int i = 1; i < 8; ++i) {
                const V v = frd[i];
#pragma unroll
                for (int j = 0; j < U; ++j) acc.v[j] += v.v[j];
            }
            if (part_used) pwr[pb * pt_tile] = acc;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(pfull0 + 8u * pb);
        // ---- stage 2 of the previous tile, by one warp in turn ------------------------------
        if (tile > 0) {
            if (warp == s2warp) stage2(tile - 1);
            if (++s2warp == nwarps) s2warp = 0;
        }
    }
    if (warp == s2warp) stage2(ntiles - 1);

    if (has_unit) {
        const int last_row = (a.K - 1) - (ntiles - 1) * a.T;
        const V fl = fwr[last_row * kFtPitch];            // this lane's forces of the window's last step
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const int g = tq * U + j;
            if (g >= gcount) continue;
            const int64_t idx = (m0 + g) * n + u;
            core.store(j, a, idx);
            if (a.forces_last) ((T*)a.forces_last)[idx] = fl.v[j];
        }
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
