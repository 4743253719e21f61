// mb_core.cuh -- per-motor-unit arithmetic shared by the single-step (K1) and
// fused (K2) kernels.  One call = one unit, one step; SURVEY.md Appendix A.2
// lists the reference operations line by line.  Citations: pool: =
// pymuscle/potvin_fuglevand_2017_motor_neuron_pool.py, fibers: =
// pymuscle/potvin_fuglevand_2017_muscle_fibers.py, pmf: = pymuscle/pymuscle_fibers.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// MB_BOUNDS build (make bounds -> libmuscle_b200_bounds.so): device-side asserts on the index of every state / output
// access and on the shared-memory slots that depend on the batch layout.  compute-sanitizer is not available on the
// GPU pool, so one pytest target (tests/test_gpu_bounds.py) runs the odd-sized and ragged shapes through this build.
#ifdef MB_BOUNDS
#include <cassert>
#define MB_ASSERT(cond) assert(cond)
#else
#define MB_ASSERT(cond) ((void)0)
#endif

namespace mb {

// ----------------------------------------------------------------------------------
// Packed per-unit parameters.  The table is L units long and stored as groups of
// 128-bit vectors, [group][unit], so that a warp of consecutive units issues fully
// coalesced LDG.128s.  mb_params_pack() (host) is the only writer.
// ----------------------------------------------------------------------------------
constexpr int kGroupsF32 = 3;   // float4  x 3 per unit
constexpr int kGroupsF64 = 5;   // double2 x 5 per unit

struct UnitF {
    float crit;    // smallest float32 input for which the reference's float64 predicate recruits (pool:169-172)
    float thr_g;   // gain * (thr - minR):            r = x*gain_e - thr_g            (pool:169-173)
    float peak;    // peak firing rate                                                (pool:176-177)
    float phir;    // phi * ratio,  ratio = (thr-1)/(RR-1)                            (pool:231-232)
    float phir6;   // phi * ratio * (minR - d):       q = r*phir - phir6
    float ctA;     // kappa * ct * (1 + ccr) / 1000        (kappa: see kKappa below)
    float ctB;     // kappa * ct * ccr / (1000 * ptf):  kappa*ct_cur/1000 = ctA - ctB*cpf  (fibers:123-125, :220)
    float fat;     // nominal fatigability                                             (fibers:110)
    float ptf;     // peak twitch force (upper clamp)                                  (pmf:104-105)
    float ptf_lo2; // third float of the peak: ptf + ptf_lo + ptf_lo2 is the float64 peak exactly
    float rdi;     // rec / ptf:  recovery = (ptf - cpf) * rdi * dt                    (pmf:137-140)
    float ptf_lo;  // (float)(peak - ptf): ptf + ptf_lo carries 48 bits of the float64 peak
};

struct UnitD {
    double thr, peak, phir, phir6, ctA, ctB, fat, ptf, rec, rdi;
};

__device__ __forceinline__ UnitF make_unit(const float4 a, const float4 b, const float4 c) {
    UnitF u;
    u.crit = a.x; u.thr_g = a.y; u.peak = a.z; u.phir = a.w;
    u.phir6 = b.x; u.ctA = b.y; u.ctB = b.z; u.fat = b.w;
    u.ptf = c.x; u.ptf_lo2 = c.y; u.rdi = c.z; u.ptf_lo = c.w;
    return u;
}
__device__ __forceinline__ UnitD make_unit(const double2 a, const double2 b, const double2 c, const double2 d,
                                           const double2 e) {
    UnitD u;
    u.thr = a.x; u.peak = a.y; u.phir = b.x; u.phir6 = b.y; u.ctA = c.x; u.ctB = c.y;
    u.fat = d.x; u.ptf = d.y; u.rec = e.x; u.rdi = e.y;
    return u;
}

// from the global table [group][L] (read-only path)
__device__ __forceinline__ UnitF load_unit(const float4* __restrict__ P, int64_t L, int64_t i) {
    return make_unit(__ldg(P + i), __ldg(P + L + i), __ldg(P + 2 * L + i));
}
__device__ __forceinline__ UnitD load_unit(const double2* __restrict__ P, int64_t L, int64_t i) {
    return make_unit(__ldg(P + i), __ldg(P + L + i), __ldg(P + 2 * L + i), __ldg(P + 3 * L + i), __ldg(P + 4 * L + i));
}
// from a staged tile [group][W] in shared memory
__device__ __forceinline__ UnitF load_unit_staged(const float4* P, int W, int i) {
    return make_unit(P[i], P[W + i], P[2 * W + i]);
}
__device__ __forceinline__ UnitD load_unit_staged(const double2* P, int W, int i) {
    return make_unit(P[i], P[W + i], P[2 * W + i], P[3 * W + i], P[4 * W + i]);
}

// Scalars derived from MbConfig by the launcher (host, float64) -- passed by value.
struct ScalF {
    float gain_e;     // firing_gain * input_scale
    float argmin;     // -max_duration*log2(e)/tau : lower bound of the exponent
    float c25k;       // ((1 - exp(-2*0.4^3)) / 0.4) / kappa      (fibers:240-241)
    float ythr;       // 0.4 * kappa                               (fibers:233-236)
    float inv_out;    // 1 / output_divisor
};
struct ScalD {
    double input_scale, min_rate, gain, mitau /* -1/tau */, max_dur, c25, out_div;
};

// FP32 mode folds the sigmoid's constant into the contraction-time factors: with
// kappa = cbrt(2*log2(e)) and y = kappa*x,  exp(-2*x^3) = ex2(-(y*y*y)).  UnitF::ctA/ctB
// are stored pre-multiplied by kappa.
constexpr double kKappa = 1.4236443587288568;   // cbrt(2 * 1.4426950408889634)
constexpr double kLinThrD = 0.4;
constexpr double kCbrt2 = 1.2599210498948732;     // FP64 fused kernel: (cbrt(2) x)^3 = 2 x^3

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ----------------------------------------------------------------------------------
// FP32 mode
// ----------------------------------------------------------------------------------

// Recruitment decision and the un-thresholded rate min(gain*(e - thr + minR), peak).
// The mask is one compare against the host-derived critical input, which reproduces
// the float64 predicate bit for bit; the rate value itself only
// needs float accuracy.  The caller zeroes the rate (or what follows from it) when
// the unit is not recruited.
__device__ __forceinline__ float pool_rate_raw(const UnitF& p, const ScalF& s, float x, bool& on) {
    on = (x >= p.crit);
    return fminf(fmaf(x, s.gain_e, -p.thr_g), p.peak);
}

// Adaptation a = q * (1 - exp(-dur/tau)),  q = phi*(r - minR + d)*ratio   (pool:208-233).
// dur_arg = dur * adk (<= 0), dur BEFORE this step's update.  NOT yet clamped at 0.
__device__ __forceinline__ float pool_adaptation(const UnitF& p, float r, float dur_arg) {
    const float q = fmaf(r, p.phir, -p.phir6);
    const float ex = ex2_approx(dur_arg);
    return fmaf(-q, ex, q);
}

// Normalised force and unit force for firing rate fr at capacity cpf
// (fibers:211-259).  Branch-free: both curve pieces are evaluated.
__device__ __forceinline__ float fibers_force(const UnitF& p, const ScalF& s, float fr, float cpf, float& nf) {
    const float ctk = fmaf(cpf, -p.ctB, p.ctA);      // kappa * ct_cur / 1000
    const float y = ctk * fr;
    const float lin = y * s.c25k;
    const float y2 = y * y;
    const float sig = 1.0f - ex2_approx(-(y2 * y));
    nf = (y <= s.ythr) ? lin : sig;
    return nf * cpf;
}

// ----------------------------------------------------------------------------------
// FP64 mode.  The three operations that define the recruitment mask keep the
// reference's order and are protected from FMA contraction; everything else may
// be re-associated (< 5e-14 over 10,000 steps).
// ----------------------------------------------------------------------------------
__device__ __forceinline__ double pool_rate(const UnitD& p, const ScalD& s, double x, bool& on) {
    const double e = __dmul_rn(x, s.input_scale);                       // muscle.py:275
    const double r0 = __dadd_rn(__dsub_rn(e, p.thr), s.min_rate);       // pool:169-170
    on = !(r0 < s.min_rate);                                            // pool:171
    const double r = (on ? r0 : 0.0) * s.gain;                          // pool:172-173
    return (r > p.peak) ? p.peak : r;                                   // pool:176-177
}

__device__ __forceinline__ double pool_adapt(const UnitD& p, const ScalD& s, double r, double dur) {
    const double q = fma(r, p.phir, -p.phir6);
    const double a = q * (1.0 - exp(dur * s.mitau));
    return r - ((a < 0.0) ? 0.0 : a);
}

// 2^(j/32), j = 0..31 (table of the fast exponential below)
static __constant__ double kExp2Tab[32] = {
    0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0,
    0x1.172b83c7d517bp+0, 0x1.1d4873168b9aap+0, 0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0,
    0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0, 0x1.3dea64c123422p+0, 0x1.44e086061892dp+0,
    0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0, 0x1.6247eb03a5585p+0,
    0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0, 0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0,
    0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0,
    0x1.ae89f995ad3adp+0, 0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0,
    0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0, 0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0};

// exp(x) for -700 <= x <= 0 in ten FP64 instructions (the library exp needs ~3x as many):
// x = n*ln2/32 + r, |r| <= ln2/64, exp(x) = 2^(n>>5) * 2^((n&31)/32) * P4(r).
//   * reduction: ONE fma with ln2/32 rounded to double -- the product is exact inside the fma, so the only error
//     is the constant's 2^-53 relative error: |x| * 1.1e-16 <= 8e-14 absolute in r, i.e. relative in exp(x);
//   * P4 = degree-4 Taylor: truncation r^5/120 <= 1.3e-12.
// Relative error < 2e-12 against the FP64 mode's 1e-9 bar (errors in the normalised force enter the capacity
// decrement linearly, they are not amplified over the steps).  `tab` is a shared-memory copy of kExp2Tab
// (divergent index).
// The four constants that are not 32-bit immediates.  Written as literals the compiler rebuilds each of them with two
// moves per use (8 issue slots per exponential); a kernel that passes them in through its argument struct (constant
// bank -> uniform registers, loaded once) gets them for free.
struct ExpConsts {
    double k32_ln2;      // 32 / ln 2
    double nln2_32;      // -(ln 2) / 32
    double c24, c6;      // 1/24, 1/6
};
__host__ __device__ __forceinline__ ExpConsts exp_consts() {
    return ExpConsts{0x1.71547652b82fep+5, -0x1.62e42fefa39efp-6, 1.0 / 24.0, 1.0 / 6.0};
}

template <bool CLAMP>
__device__ __forceinline__ double exp_neg_fast(double x, const double* tab, const ExpConsts& k) {
    if (CLAMP) x = fmax(x, -700.0);                                     // (the launcher proves it redundant for most tables)
    const double kMagic = 6755399441055744.0;                          // 1.5 * 2^52: round to integer
    const double t = fma(x, k.k32_ln2, kMagic);
    const int n = __double2loint(t);
    const double nf = t - kMagic;
    const double r = fma(nf, k.nln2_32, x);
    double p = fma(r, k.c24, k.c6);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double s = tab[n & 31] * p;
    return __hiloint2double(__double2hiint(s) + ((n >> 5) << 20), __double2loint(s));
}
template <bool CLAMP>
__device__ __forceinline__ double exp_neg_fast(double x, const double* tab) {
    return exp_neg_fast<CLAMP>(x, tab, exp_consts());
}

// fibers_force with the fast exponential (fused kernel)
template <bool CLAMP>
__device__ __forceinline__ double fibers_force_fast(const UnitD& p, const ScalD& s, double fr, double cpf, double& nf,
                                                    const double* tab) {
    const double ctk = fma(cpf, -p.ctB, p.ctA);
    const double x = ctk * fr;
    if (x <= kLinThrD) {
        nf = x * s.c25;
    } else {
        nf = 1.0 - exp_neg_fast<CLAMP>(-2.0 * (x * x * x), tab);
    }
    return nf * cpf;
}

__device__ __forceinline__ double fibers_force(const UnitD& p, const ScalD& s, double fr, double cpf, double& nf) {
    const double ctk = fma(cpf, -p.ctB, p.ctA);
    const double x = ctk * fr;
    if (x <= kLinThrD) {
        nf = x * s.c25;
    } else {
        nf = 1.0 - exp(-2.0 * (x * x * x));
    }
    return nf * cpf;
}

// Capacity update, float64 state (used by K1 in both modes and by K2 in FP64 mode).
// fatdt = fat*dt, rdidt = (rec/ptf)*dt (both zero when peripheral fatigue is off).  The
// recovery term is written (ptf - c) * rdidt so that it is exactly 0 for a rested unit,
// as in the reference's rec * ((ptf - c)/ptf) * dt.
// c < 0 ? 0 : c on the bit pattern: three integer instructions and no predicate instead of an FP64
// compare and two selects (-0.0 becomes +0.0, which compares equal)
__device__ __forceinline__ double zero_if_negative(double c) {
    const int hi = __double2hiint(c);
    const int keep = ~(hi >> 31);
    return __hiloint2double(hi & keep, __double2loint(c) & keep);
}

template <bool RECOVERY>
__device__ __forceinline__ double fatigue_update(double cpf, double nf, double fatdt, double rdidt, double ptf) {
    double c = fma(-nf, fatdt, cpf);                                    // fibers:110-111
    if (RECOVERY) {
        if (nf <= 0.0) c = fma(ptf - c, rdidt, c);                      // pmf:121-141
    }
    c = zero_if_negative(c);                                            // fibers:114
    if (RECOVERY) c = (c > ptf) ? ptf : c;                              // pmf:104-105
    return c;
}

__device__ __forceinline__ double duration_update(double dur, bool firing, double dt, double max_dur) {
    double d = firing ? dur + dt : dur;                                 // pool:196-197
    d = (d > max_dur) ? max_dur : d;                                    // pool:202-203
    return (d < 0.0) ? 0.0 : d;                                         // pool:205-206
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Eight independent warp sums at once ("transposing" butterfly): every exchange halves the number of
// values a lane carries, so 4 + 2 + 1 + 2 = 9 shuffles replace the 40 of eight separate warp_sum calls.
// Afterwards every lane holds the warp total of row (lane >> 2) & 7.  Fixed order => deterministic.
template <typename T>
__device__ __forceinline__ T warp_sum_rows8(const T (&v)[8], int lane) {
    const bool up16 = lane & 16, up8 = lane & 8, up4 = lane & 4;
    T a[4], b[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {                                      // lanes with bit 4 keep rows 4..7
        const T keep = up16 ? v[i + 4] : v[i], send = up16 ? v[i] : v[i + 4];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {                                      // bit 3 picks rows +2, +3
        const T keep = up8 ? a[i + 2] : a[i], send = up8 ? a[i] : a[i + 2];
        b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    const T keep = up4 ? b[1] : b[0], send = up4 ? b[0] : b[1];        // bit 2 picks row +1
    T r = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    r += __shfl_xor_sync(0xffffffffu, r, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;
}

// per-muscle output scaling (muscle.py:278): FP32 mode multiplies by the reciprocal, FP64 mode divides
__device__ __forceinline__ float  out_scale(const ScalF& s, float v)  { return v * s.inv_out; }
__device__ __forceinline__ double out_scale(const ScalD& s, double v) { return v / s.out_div; }
__device__ __forceinline__ float  out_div(float v, double d)  { return (float)((double)v / d); }
__device__ __forceinline__ double out_div(double v, double d) { return v / d; }

// the float64 peak twitch force, exactly (a rested unit has capacity == peak, and must keep it)
__device__ __forceinline__ double unit_peak(const UnitF& p) { return ((double)p.ptf + (double)p.ptf_lo) + (double)p.ptf_lo2; }
__device__ __forceinline__ double unit_peak(const UnitD& p) { return p.ptf; }

}  // namespace mb
