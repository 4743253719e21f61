// muscle_b200.cu -- sm_100a kernels + C ABI (include/muscle_b200.h) for the
// PyMuscle per-step hot path.  See DESIGN.md for the data layout and the
// roofline of each kernel.
//
//   K1  k1_step<T>        one step, one launch, thread per motor unit, HBM-bound
//   K2  k2_fused_*        K steps per launch, state in registers, FP-pipe-bound
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared ...
#include "mb_launchers.cuh"
#include "mb_fused_warp64.cuh"
#include "mb_fused_wide.cuh"
#include "mb_flat.cuh"
#include "mb_fused_warp.cuh"     // tile constants of the warp kernels (geometry); nothing is instantiated here

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>

using namespace mb;

// =====================================================================================
// error plumbing
// =====================================================================================
static thread_local char g_err[512] = "";

int mb::fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

// the environment is read once per process (not on every launch)
const Knobs& mb::knobs() {
    static const Knobs k = [] {
        Knobs q;
        q.k1_generic = env_int("MB_K1_GENERIC", 0);
        q.k1_stages = env_int("MB_K1_STAGES", 0);
        q.k1_w = env_int("MB_K1_W", 0);
        q.k1_bps = env_int("MB_K1_BPS", 0);
        q.k1_flat = env_int("MB_K1_FLAT", -1);
        q.k1_flat_group = env_int("MB_K1_FLAT_GROUP", 0);
        q.fused_block = env_int("MB_FUSED_BLOCK", 0);
        q.fused_no_bulk = env_int("MB_FUSED_NO_BULK", 0);
        q.fused_general = env_int("MB_FUSED_GENERAL", 0);
        q.f64_block = env_int("MB_F64_BLOCK", 0);
        q.fused_wide = env_int("MB_FUSED_WIDE", 0);
        q.fused_persist = env_int("MB_FUSED_PERSIST", -1);
        q.fused_units = env_int("MB_FUSED_UNITS", -1);
        return q;
    }();
    return k;
}

int mb::current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    return dev;
}

int mb::sm_count() {
    // cached per device: one process may drive several GPUs
    static thread_local int c_dev = -1, n = 0;
    const int dev = current_device();
    if (dev != c_dev) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        c_dev = dev;
    }
    return n;
}

extern "C" int mb_abi_version(void) { return MB_ABI_VERSION; }
extern "C" const char* mb_last_error(void) { return g_err; }

static int check_cfg(const MbConfig* c) {
    if (!c) return fail(MB_EINVAL, "cfg is NULL");
    if (c->precision != MB_F32 && c->precision != MB_F64) return fail(MB_EINVAL, "bad precision %d", c->precision);
    if (c->fibers_kind != MB_FIBERS_POTVIN && c->fibers_kind != MB_FIBERS_PYMUSCLE)
        return fail(MB_EINVAL, "bad fibers_kind %d", c->fibers_kind);
    if (!(c->adaptation_time_constant > 0)) return fail(MB_EINVAL, "adaptation_time_constant must be > 0");
    if (!(c->firing_gain > 0)) return fail(MB_EINVAL, "firing_gain must be > 0");
    if (!(c->output_divisor > 0)) return fail(MB_EINVAL, "output_divisor must be > 0");
    if (!(c->input_scale > 0)) return fail(MB_EINVAL, "input_scale must be > 0");
    return MB_OK;
}

// =====================================================================================
// parameter packing (host)
// =====================================================================================
extern "C" size_t mb_params_bytes(const MbConfig* cfg, int64_t L) {
    if (!cfg || L <= 0) return 0;
    return cfg->precision == MB_F32 ? (size_t)L * kGroupsF32 * sizeof(float4)
                                    : (size_t)L * kGroupsF64 * sizeof(double2);
}

// float <-> monotone unsigned key (total order of the finite floats and infinities)
static inline uint32_t f2key(float f) {
    uint32_t b;
    memcpy(&b, &f, 4);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
static inline float key2f(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    float f;
    memcpy(&f, &b, 4);
    return f;
}

// The reference's recruitment decision for a float32 input x, in float64
// (muscle.py:275 then pool:169-172).  `volatile` keeps the compiler from
// contracting or re-associating the three roundings.
static bool ref_recruited(float x, double input_scale, double thr, double min_rate) {
    volatile double e = (double)x * input_scale;
    volatile double d = e - thr;
    volatile double r = d + min_rate;
    return !(r < min_rate);
}

// Smallest float32 x with ref_recruited(x); the predicate is monotone in x.
static float critical_input(double input_scale, double thr, double min_rate) {
    uint32_t lo = f2key(-std::numeric_limits<float>::max());
    uint32_t hi = f2key(std::numeric_limits<float>::infinity());
    if (ref_recruited(key2f(lo), input_scale, thr, min_rate)) return -std::numeric_limits<float>::infinity();
    if (!ref_recruited(key2f(hi), input_scale, thr, min_rate)) return std::numeric_limits<float>::quiet_NaN();
    while (hi - lo > 1) {                       // invariant: lo not recruited, hi recruited
        uint32_t mid = lo + (hi - lo) / 2;
        if (ref_recruited(key2f(mid), input_scale, thr, min_rate)) hi = mid; else lo = mid;
    }
    return key2f(hi);
}

extern "C" int mb_params_pack(const MbConfig* cfg, int64_t L, const double* thr, const double* peak,
                              const double* ptf, const double* ct, const double* fat, const double* rec,
                              void* out_host, int32_t* table_props) {
    if (int rc = check_cfg(cfg)) return rc;
    if (L <= 0 || !thr || !peak || !ptf || !ct || !fat || !out_host) return fail(MB_EINVAL, "NULL table or L <= 0");
    if (cfg->fibers_kind == MB_FIBERS_PYMUSCLE && !rec) return fail(MB_EINVAL, "rec table required for PyMuscle fibers");
    const double rr = cfg->max_recruitment_threshold, phi = cfg->adaptation_magnitude;
    const double minr = cfg->min_firing_rate, dd = cfg->derecruitment_delta, g = cfg->firing_gain;
    const double ccr = cfg->contraction_time_change_ratio;
    // A recruited unit fires at >= gain*minR, so q = phi*ratio*(r - (minR - d)) >= 0 for every
    // recruited unit iff phi*ratio >= 0 and gain*minR clears minR - d (with float slack).
    bool adapt_nonneg = (g * minr) * (1.0 - 1e-5) > (minr - dd) && g > 0.0 && minr > 0.0;
    double xmax = 0.0;                        // largest normalised firing rate ct_cur * rate / 1000 any unit can reach
    for (int64_t i = 0; i < L; ++i) {
        const double ratio = (thr[i] - 1.0) / (rr - 1.0);                 // pool:231
        const double phir = phi * ratio;
        const double phir6 = phir * (minr - dd);
        adapt_nonneg = adapt_nonneg && (phir >= 0.0);
        // ... and the peak clamp (pool:176-177) must not pull a recruited unit's rate below minR - d either:
        // q = phir * (min(r, peak) - (minR - d)) would turn negative and pool:221 would clamp the adaptation to 0
        adapt_nonneg = adapt_nonneg && (peak[i] * (1.0 - 1e-5) > (minr - dd));
        const double ctA = ct[i] * (1.0 + ccr) / 1000.0;
        if (ctA * peak[i] > xmax) xmax = ctA * peak[i];
        const double ctB = ct[i] * ccr / (1000.0 * ptf[i]);
        const double r = rec ? rec[i] : 0.0;
        const double rdi = r / ptf[i];
        if (cfg->precision == MB_F32) {
            float4* P = (float4*)out_host;
            P[i] = make_float4(critical_input(cfg->input_scale, thr[i], minr), (float)(g * (thr[i] - minr)),
                               (float)peak[i], (float)phir);
            P[L + i] = make_float4((float)phir6, (float)(ctA * kKappa), (float)(ctB * kKappa), (float)fat[i]);
            const float p0 = (float)ptf[i];
            const float p1 = (float)(ptf[i] - (double)p0);
            const float p2 = (float)((ptf[i] - (double)p0) - (double)p1);          // p0 + p1 + p2 == ptf[i] exactly
            P[2 * L + i] = make_float4(p0, p2, (float)rdi, p1);
        } else {
            double2* P = (double2*)out_host;
            P[i] = make_double2(thr[i], peak[i]);
            P[L + i] = make_double2(phir, phir6);
            P[2 * L + i] = make_double2(ctA, ctB);
            P[3 * L + i] = make_double2(fat[i], ptf[i]);
            P[4 * L + i] = make_double2(r, rdi);
        }
    }
    if (table_props)
        *table_props = (adapt_nonneg ? MB_PROP_ADAPT_NONNEG : 0) | ((2.0 * xmax * xmax * xmax < 600.0) ? MB_PROP_EXP_BOUNDED : 0);
    return MB_OK;
}

// =====================================================================================
// K1: single step.  blockDim = (TX, TY): TY muscles per block iteration, TX threads
// (a multiple of 32) striding over one muscle's units.  Per-muscle total = warp
// shuffle tree + one shared-memory stage, in a fixed order (deterministic).
// =====================================================================================
template <typename T> struct Vec;
template <> struct Vec<float>  { using P = float4;  using U = UnitF; using S = ScalF; };
template <> struct Vec<double> { using P = double2; using U = UnitD; using S = ScalD; };

template <typename T>
struct K1Args {
    const typename Vec<T>::P* params;
    int64_t param_stride;          // 0: one shared table; else vectors between the tables of consecutive rows
    int64_t rows, L;
    int32_t S;
    const int32_t* seg_off;
    const double* seg_div;
    int32_t stage, in_per_unit, recovery, central, peripheral;
    const T* in;
    double* cpf;
    double* dur;
    T* forces;
    T* totals;
    uint8_t* mask;
    T* rates;
    double dt;
    double adk_d;      // -log2e/tau (FP32 mode exponent scaling, applied in float64)
    double max_dur;
    typename Vec<T>::S sc;
};

// Per-unit arithmetic of one step.  Loads and stores are done by the caller (all loads of a
// thread's R rows are issued before any arithmetic so that they overlap); this only maps
// (input x, capacity c, duration d) -> (force, new capacity, new duration, mask, rate).
struct K1Out {
    double c_new, d_new;
    bool on;
};

// FP32 mode: float arithmetic, float64 state
__device__ __forceinline__ float k1_compute(const K1Args<float>& a, const UnitF& p, float x, double c, double d,
                                            K1Out& o, float& rate) {
    float fr = x;
    o.on = false;
    o.d_new = d;
    o.c_new = c;
    if (a.stage != MB_STAGE_FIBERS) {
        float r = pool_rate_raw(p, a.sc, x, o.on);
        r = o.on ? r : 0.0f;                                            // pool:171-172
        fr = r;
        if (a.central) {
            fr = r - fmaxf(pool_adaptation(p, r, (float)(d * a.adk_d)), 0.0f);   // pool:221, :121
            o.d_new = duration_update(d, r > 0.0f, a.dt, a.max_dur);
        }
        rate = fr;
        if (a.stage == MB_STAGE_POOL) return 0.0f;
    }
    float nf;
    const float F = fibers_force(p, a.sc, fr, (float)c, nf);
    if (a.peripheral) {
        const double fatdt = (double)p.fat * a.dt, rdidt = (double)p.rdi * a.dt;
        const double ptf = unit_peak(p);
        o.c_new = a.recovery ? fatigue_update<true>(c, (double)nf, fatdt, rdidt, ptf)
                             : fatigue_update<false>(c, (double)nf, fatdt, rdidt, ptf);
    }
    return F;
}

// FP64 mode
__device__ __forceinline__ double k1_compute(const K1Args<double>& a, const UnitD& p, double x, double c, double d,
                                             K1Out& o, double& rate) {
    double fr = x;
    o.on = false;
    o.d_new = d;
    o.c_new = c;
    if (a.stage != MB_STAGE_FIBERS) {
        const double r = pool_rate(p, a.sc, x, o.on);
        fr = r;
        if (a.central) {
            fr = pool_adapt(p, a.sc, r, d);
            o.d_new = duration_update(d, r > 0.0, a.dt, a.max_dur);
        }
        rate = fr;
        if (a.stage == MB_STAGE_POOL) return 0.0;
    }
    double nf;
    const double F = fibers_force(p, a.sc, fr, c, nf);
    if (a.peripheral) {
        const double fatdt = p.fat * a.dt, rdidt = p.rdi * a.dt;
        o.c_new = a.recovery ? fatigue_update<true>(c, nf, fatdt, rdidt, p.ptf)
                             : fatigue_update<false>(c, nf, fatdt, rdidt, p.ptf);
    }
    return F;
}

// One block = one muscle (segment) of R consecutive rows at a time.  A thread owns unit
// positions tx, tx+TX, ... of the segment and advances that unit in all R rows: the unit's
// parameters (48 / 80 bytes from L2) are loaded once per R unit-steps, so L2 traffic stays
// below the HBM traffic of the state (20-40 bytes per unit-step), and the R rows give R
// independent HBM loads in flight per thread.
template <typename T, int R, int MINB>
__global__ void __launch_bounds__(256, MINB) k1_step(const K1Args<T> a) {
    __shared__ T red[R][8];
    const int tx = threadIdx.x, TX = blockDim.x, lane = tx & 31, warp = tx >> 5, nwarps = TX >> 5;
    const int64_t chunks = (a.rows + R - 1) / R;
    const int64_t items = chunks * a.S;
    const bool need_c = a.stage != MB_STAGE_POOL;
    const bool need_d = a.stage != MB_STAGE_FIBERS && a.central;
    for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
        const int64_t chunk = item / a.S;
        const int seg = (int)(item - chunk * a.S);
        const int64_t r0 = chunk * R;
        const int off0 = a.seg_off ? a.seg_off[seg] : 0;
        const int n = a.seg_off ? a.seg_off[seg + 1] - off0 : (int)a.L;
        const int nrows = (int)min((int64_t)R, a.rows - r0);
        T xm[R], partial[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            partial[r] = 0;
            xm[r] = (!a.in_per_unit && r < nrows) ? a.in[(r0 + r) * a.S + seg] : T(0);
        }
        for (int u = tx; u < n; u += TX) {
            auto p = load_unit(a.params + r0 * a.param_stride, a.L, off0 + u);
            const int64_t i0 = r0 * a.L + off0 + u;
            MB_ASSERT(off0 + u < a.L && i0 + (int64_t)(nrows - 1) * a.L < a.rows * a.L);
            double c[R], d[R];
            T x[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {                    // all loads first
                const int64_t i = i0 + (int64_t)r * a.L;
                const bool v = r < nrows;
                c[r] = (v && need_c) ? a.cpf[i] : 0.0;
                d[r] = (v && need_d) ? a.dur[i] : 0.0;
                x[r] = a.in_per_unit ? (v ? a.in[i] : T(0)) : xm[r];
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (r >= nrows) continue;
                const int64_t i = i0 + (int64_t)r * a.L;
                if (a.param_stride && r > 0) p = load_unit(a.params + (r0 + r) * a.param_stride, a.L, off0 + u);
                K1Out o;
                T rate = 0;
                const T F = k1_compute(a, p, x[r], c[r], d[r], o, rate);
                if (need_d) a.dur[i] = o.d_new;
                if (a.stage != MB_STAGE_FIBERS) {
                    if (a.mask) a.mask[i] = o.on ? 1 : 0;
                    if (a.rates) a.rates[i] = rate;
                }
                if (need_c) {
                    if (a.forces) a.forces[i] = F;
                    if (a.peripheral) a.cpf[i] = o.c_new;
                    partial[r] += F;
                }
            }
        }
        if (a.totals && need_c) {                             // block-uniform condition
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const T w = warp_sum(partial[r]);
                if (lane == 0) red[r][warp] = w;
            }
            __syncthreads();
            if (tx < nrows) {
                T s = red[tx][0];
                for (int w = 1; w < nwarps; ++w) s += red[tx][w];
                a.totals[(r0 + tx) * a.S + seg] = a.seg_div ? out_div(s, a.seg_div[seg]) : out_scale(a.sc, s);
            }
            __syncthreads();
        }
    }
}

// =====================================================================================
// K3: per-muscle capacity sums (StandardMuscle.get_peripheral_fatigue, muscle.py:248-256)
// =====================================================================================
// fatigue = sum(peak - capacity) / max_arb_output: algebraically 1 - sum(capacity)/max_arb_output for a StandardMuscle
// (max_arb_output = sum(peak), muscle.py:187), exactly 0 at rest whatever the summation order, never negative.
template <typename PT>
__device__ __forceinline__ double table_peak(const PT* params, int64_t L, int64_t i);
template <>
__device__ __forceinline__ double table_peak<float4>(const float4* P, int64_t L, int64_t i) {
    const float4 c = __ldg(P + 2 * L + i);                // (ptf, ptf_lo2, rdi, ptf_lo)
    return ((double)c.x + (double)c.w) + (double)c.y;
}
template <>
__device__ __forceinline__ double table_peak<double2>(const double2* P, int64_t L, int64_t i) {
    return __ldg(P + 3 * L + i).y;                        // (fat, ptf)
}

template <typename PT>
__global__ void __launch_bounds__(256) k3_fatigue(const PT* params, int64_t param_stride, int64_t rows, int64_t L, int32_t S,
                                                  const int32_t* seg_off, const double* seg_div, const double* cpf, double* out,
                                                  double out_div) {
    __shared__ double red[8];
    const int64_t n_muscles = rows * S;
    for (int64_t mu = blockIdx.x; mu < n_muscles; mu += gridDim.x) {
        const int64_t row = mu / S;
        const int seg = (int)(mu - row * S);
        const int off0 = seg_off ? seg_off[seg] : 0;
        const int n = seg_off ? seg_off[seg + 1] - off0 : (int)L;
        const PT* table = params + row * param_stride;
        double part = 0.0;
        for (int u = threadIdx.x; u < n; u += blockDim.x) part += table_peak<PT>(table, L, off0 + u) - cpf[row * L + off0 + u];
        part = warp_sum(part);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
            out[mu] = s / (seg_div ? seg_div[seg] : out_div);
        }
        __syncthreads();
    }
}

// =====================================================================================
// peak probes: dependent FMA / MUFU chains, 8 independent chains per thread
// =====================================================================================
template <int KIND>
__global__ void __launch_bounds__(1024) k_peak_probe(int iters, double* sink) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (KIND == 0) {
        float v[8];
        const float b = 1.0f - 1e-7f * (float)(t & 7), c = 1e-9f * (float)(t & 3);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 1.0f + 0.01f * (float)i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], b, c);
        }
        float s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[i];
        sink[t] = (double)s;
    } else if (KIND == 1) {
        double v[8];
        const double b = 1.0 - 1e-12 * (double)(t & 7), c = 1e-15 * (double)(t & 3);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 1.0 + 0.01 * (double)i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = fma(v[i], b, c);
        }
        double s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[i];
        sink[t] = s;
    } else if (KIND == 2) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.1f * (float)(i + 1) + 1e-3f * (float)(t & 15);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = ex2_approx(-v[i]);
        }
        float s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[i];
        sink[t] = (double)s;
    } else if (KIND == 3) {
        // packed FP32x2 FMA (sm_100 fma.rn.f32x2 -> SASS FFMA2): 8 chains of 2 floats
        unsigned long long v[8], b, c;
        {
            const float bf = 1.0f - 1e-7f * (float)(t & 7), cf = 1e-9f * (float)(t & 3);
            asm("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(bf));
            asm("mov.b64 %0, {%1, %1};" : "=l"(c) : "f"(cf));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float x = 1.0f + 0.01f * (float)i;
            asm("mov.b64 %0, {%1, %1};" : "=l"(v[i]) : "f"(x));
        }
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(b), "l"(c));
        }
        float s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float lo, hi;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[i]));
            s += lo + hi;
        }
        sink[t] = (double)s;
    } else if (KIND == 4) {
        // ALU pipe: FMNMX chains
        float v[8];
        const float lo = 0.25f + 1e-3f * (float)(t & 7), hi = 0.75f;
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.1f * (float)(i + 1);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = fminf(v[i] + 0.0f * 0.0f, hi);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], lo);
            }
        }
        float s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[i];
        sink[t] = (double)s;
    } else if (KIND == 6 || KIND == 7) {
        // packed FP32x2 multiply (6) / add (7): 8 chains of 2 floats
        unsigned long long v[8], b;
        {
            const float bf = (KIND == 6) ? 1.0f - 1e-7f * (float)(t & 7) : 1e-9f * (float)(t & 3);
            asm("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(bf));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float x = 1.0f + 0.01f * (float)i;
            asm("mov.b64 %0, {%1, %1};" : "=l"(v[i]) : "f"(x));
        }
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (KIND == 6) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(v[i]) : "l"(b));
                    else asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v[i]) : "l"(b));
                }
        }
        float s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float lo, hi;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[i]));
            s += lo + hi;
        }
        sink[t] = (double)s;
    } else {
        // issue mix of the fused kernel: per 8 "instructions" 5 FFMA, 2 FMNMX, 1 MUFU
        float v[8];
        const float b = 1.0f - 1e-7f * (float)(t & 7), c = 1e-3f * (float)(t & 3), hi = 0.75f;
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.1f * (float)(i + 1);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float x = v[i];
                x = fmaf(x, b, c); x = fmaf(x, b, c); x = fminf(x, hi); x = fmaf(x, b, c);
                x = ex2_approx(-x); x = fmaf(x, b, c); x = fmaxf(x, c); x = fmaf(x, b, c);
                v[i] = x;
            }
        }
        float s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[i];
        sink[t] = (double)s;
    }
}

extern "C" int mb_peak_probe(int32_t kind, int32_t blocks, int32_t threads, int32_t iters, void* sink_dev,
                             int64_t* lane_ops, void* stream) {
    if (kind < 0 || kind > 7 || blocks <= 0 || threads <= 0 || threads > 1024 || iters <= 0 || !sink_dev)
        return fail(MB_EINVAL, "mb_peak_probe: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (kind == 0) k_peak_probe<0><<<blocks, threads, 0, st>>>(iters, (double*)sink_dev);
    else if (kind == 1) k_peak_probe<1><<<blocks, threads, 0, st>>>(iters, (double*)sink_dev);
    else if (kind == 2) k_peak_probe<2><<<blocks, threads, 0, st>>>(iters, (double*)sink_dev);
    else if (kind == 3) k_peak_probe<3><<<blocks, threads, 0, st>>>(iters, (double*)sink_dev);
    else if (kind == 4) k_peak_probe<4><<<blocks, threads, 0, st>>>(iters, (double*)sink_dev);
    else if (kind == 5) k_peak_probe<5><<<blocks, threads, 0, st>>>(iters, (double*)sink_dev);
    else if (kind == 6) k_peak_probe<6><<<blocks, threads, 0, st>>>(iters, (double*)sink_dev);
    else k_peak_probe<7><<<blocks, threads, 0, st>>>(iters, (double*)sink_dev);
    MB_CUDA_OK(cudaGetLastError());
    if (lane_ops) *lane_ops = (int64_t)blocks * threads * 64 * (int64_t)iters;
    return MB_OK;
}

// =====================================================================================
// launchers
// =====================================================================================
template <typename T>
static int launch_k1(const MbConfig& c, const void* params, int64_t rows, int64_t L, int32_t S,
                     const int32_t* seg_off, const double* seg_div, int32_t stage, const void* in, int32_t in_per_unit, double* cpf,
                     double* dur, void* forces, void* totals, uint8_t* mask, void* rates, double dt,
                     cudaStream_t st, const typename Vec<T>::S& sc) {
    K1Args<T> a;
    a.params = (const typename Vec<T>::P*)params;
    a.param_stride = c.param_rows > 1 ? (int64_t)(sizeof(T) == 4 ? kGroupsF32 : kGroupsF64) * L : 0;
    a.rows = rows; a.L = L; a.S = S; a.seg_off = seg_off; a.seg_div = seg_div;
    a.stage = stage; a.in_per_unit = in_per_unit;
    a.recovery = c.fibers_kind == MB_FIBERS_PYMUSCLE;
    a.central = c.central_fatigue != 0;
    a.peripheral = c.peripheral_fatigue != 0;
    a.in = (const T*)in; a.cpf = cpf; a.dur = dur;
    a.forces = (T*)forces; a.totals = (T*)totals; a.mask = mask; a.rates = (T*)rates;
    a.dt = dt;
    a.adk_d = -kLog2e / c.adaptation_time_constant;
    a.max_dur = c.max_duration;
    a.sc = sc;
    // threads per block: enough for the mean segment (a multiple of 32, at most 256)
    const int64_t mean_n = (L + S - 1) / S;
    int tx = (int)((mean_n + 31) / 32) * 32;
    if (tx > 256) tx = 256;
    if (tx < 32) tx = 32;
    constexpr int R = 8;                                      // rows advanced together by a block
    const int64_t items = ((rows + R - 1) / R) * S;
    int64_t blocks = items;
    const int64_t cap = (int64_t)sm_count() * (2048 / tx) * 4;   // a few waves of resident blocks
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    k1_step<T, R, (sizeof(T) == 8 ? 1 : 2)><<<(unsigned)blocks, tx, 0, st>>>(a);    // FP64 mode: 128 registers spill
    MB_CUDA_OK(cudaGetLastError());
    return MB_OK;
}

// The streaming kernel covers the batched hot case: whole-muscle stage, one input per muscle or per unit,
// no mask / rate outputs, 16-byte aligned state arrays (any row length).
template <typename T>
static bool k1s_eligible(const MbConfig& c, int64_t L, int32_t stage, int32_t in_per_unit, const void* in,
                         const double* cpf, const double* dur, const uint8_t* mask, const void* rates,
                         const void* forces = (const void*)1, const void* fatigue = nullptr) {
    if (knobs().k1_generic) return false;
    // per-instance tables: FP32 mode with the force output and without the fatigue epilogue (one instantiation set)
    if (c.param_rows > 1 && (sizeof(T) != 4 || !forces || fatigue)) return false;
    if (stage != MB_STAGE_MUSCLE || mask || rates || !c.peripheral_fatigue) return false;
    // per-unit inputs are staged with bulk copies too (any row length: the copies start at the 16-byte
    // boundary at or below each row's tile start)
    if (in_per_unit) {
        if (((uintptr_t)in % 16) != 0) return false;
        cudaPointerAttributes at;                             // (the drop-in classes pass pinned HOST memory)
        if (cudaPointerGetAttributes(&at, in) != cudaSuccess || at.type != cudaMemoryTypeDevice) {
            cudaGetLastError();
            return false;
        }
    }
    if (((uintptr_t)cpf % 16) != 0) return false;
    if (c.central_fatigue && ((uintptr_t)dur % 16) != 0) return false;
    return true;
}

// Rows of several muscles with one input per muscle: the flat-tiled kernel (mb_flat.cuh)
// Measured on B200 (scripts/r2_ragged.py, % of HBM peak, k1_stream -> k1_flat).  FP32 mode: flat tiling wins on every
// ragged layout but one -- hopper legs 50 -> 77 % (env step 45 -> 64 %), 8 x 256 units 91 -> 94 % (72 -> 89 %), the arm
// 81 -> 85 % (79 -> 84 %), 30 x 1,137 units 89.5 -> 87.5 % (84.9 -> 85.5 %) -- so it takes all of them.  FP64 mode: only
// where muscles are short against a tile (hopper 36 -> 51 %); with muscles of several tiles k1_stream's grouped rows win.
static bool use_flat(const MbConfig& c, int64_t L, int32_t S, int32_t in_per_unit) {
    if (!(S > 1 && S <= kFlatMaxS && !in_per_unit && c.param_rows <= 1) || knobs().k1_flat == 0) return false;
    if (knobs().k1_flat > 0 || c.precision == MB_F32) return true;
    return L / S <= 192;
}

extern "C" int mb_step(const MbConfig* cfg, const void* params_dev, int64_t rows, int64_t L, int32_t S,
                       const int32_t* seg_off_dev, const double* seg_div_dev, int32_t stage, const void* in_dev, int32_t in_per_unit,
                       double* cpf_dev, double* dur_dev, void* forces_dev, void* totals_dev, uint8_t* mask_dev,
                       void* rates_dev, double step_size, void* stream) {
    if (int rc = check_cfg(cfg)) return rc;
    if (!params_dev || !in_dev) return fail(MB_EINVAL, "mb_step: params/in is NULL");
    if (rows <= 0 || L <= 0 || S <= 0) return fail(MB_EINVAL, "mb_step: rows, L, S must be positive");
    if (S > 1 && !seg_off_dev) return fail(MB_EINVAL, "mb_step: seg_off required when S > 1");
    if (stage < MB_STAGE_MUSCLE || stage > MB_STAGE_FIBERS) return fail(MB_EINVAL, "mb_step: bad stage %d", stage);
    if (stage != MB_STAGE_POOL && !cpf_dev) return fail(MB_EINVAL, "mb_step: cpf is NULL");
    if (stage != MB_STAGE_FIBERS && cfg->central_fatigue && !dur_dev) return fail(MB_EINVAL, "mb_step: dur is NULL");
    if (stage == MB_STAGE_FIBERS && !in_per_unit) return fail(MB_EINVAL, "mb_step: firing rates must be per unit");
    if (cfg->param_rows > 1 && (S != 1 || cfg->param_rows != rows))
        return fail(MB_EINVAL, "mb_step: per-instance tables need a uniform batch with param_rows == rows");
    cudaStream_t st = (cudaStream_t)stream;
    const K1sEpilogue no_epilogue;
    if (cfg->precision == MB_F32) {
        if (k1s_eligible<float>(*cfg, L, stage, in_per_unit, in_dev, cpf_dev, dur_dev, mask_dev, rates_dev, forces_dev))
            return (use_flat(*cfg, L, S, in_per_unit) ? launch_k1f_f32 : launch_k1s_f32)(*cfg, params_dev, rows, L, S, seg_off_dev, seg_div_dev, in_dev, in_per_unit, cpf_dev,
                                  dur_dev, forces_dev, totals_dev, no_epilogue, step_size, st);
        return launch_k1<float>(*cfg, params_dev, rows, L, S, seg_off_dev, seg_div_dev, stage, in_dev, in_per_unit, cpf_dev,
                                dur_dev, forces_dev, totals_dev, mask_dev, rates_dev, step_size, st, scal_f(*cfg));
    }
    if (k1s_eligible<double>(*cfg, L, stage, in_per_unit, in_dev, cpf_dev, dur_dev, mask_dev, rates_dev, forces_dev))
        return (use_flat(*cfg, L, S, in_per_unit) ? launch_k1f_f64 : launch_k1s_f64)(*cfg, params_dev, rows, L, S, seg_off_dev, seg_div_dev, in_dev, in_per_unit, cpf_dev,
                              dur_dev, forces_dev, totals_dev, no_epilogue, step_size, st);
    return launch_k1<double>(*cfg, params_dev, rows, L, S, seg_off_dev, seg_div_dev, stage, in_dev, in_per_unit, cpf_dev, dur_dev,
                             forces_dev, totals_dev, mask_dev, rates_dev, step_size, st, scal_d(*cfg));
}

// Whole-muscle step with a fused epilogue: totals *= out_gain (per muscle) and / or the Hill-type factors of
// the muscle's length history and, in the same pass, the peripheral fatigue of the NEW capacities.
// Streaming kernel only (see k1s_eligible).
static int step_env_impl(const MbConfig* cfg, const void* params_dev, int64_t rows, int64_t L, int32_t S,
                         const int32_t* seg_off_dev, const double* seg_div_dev, const void* in_dev,
                         double* cpf_dev, double* dur_dev, void* forces_dev, void* totals_dev,
                         const K1sEpilogue& ep, double step_size, void* stream) {
    if (int rc = check_cfg(cfg)) return rc;
    if (!params_dev || !in_dev || !cpf_dev) return fail(MB_EINVAL, "mb_step_env: params/in/cpf is NULL");
    if (rows <= 0 || L <= 0 || S <= 0) return fail(MB_EINVAL, "mb_step_env: rows, L, S must be positive");
    if (S > 1 && !seg_off_dev) return fail(MB_EINVAL, "mb_step_env: seg_off required when S > 1");
    if (cfg->central_fatigue && !dur_dev) return fail(MB_EINVAL, "mb_step_env: dur is NULL");
    if (!(step_size > 0) && ep.hill_rest) return fail(MB_EINVAL, "mb_step_env: the force-velocity curve needs step_size > 0");
    cudaStream_t st = (cudaStream_t)stream;
    if (cfg->precision == MB_F32) {
        if (!k1s_eligible<float>(*cfg, L, MB_STAGE_MUSCLE, 0, in_dev, cpf_dev, dur_dev, nullptr, nullptr, forces_dev, ep.fatigue))
            return fail(MB_EUNSUPPORTED, "mb_step_env needs 16-byte aligned state arrays and peripheral fatigue on");
        return (use_flat(*cfg, L, S, 0) ? launch_k1f_f32 : launch_k1s_f32)(*cfg, params_dev, rows, L, S, seg_off_dev, seg_div_dev,
                                                                        in_dev, 0, cpf_dev, dur_dev, forces_dev, totals_dev,
                                                                        ep, step_size, st);
    }
    if (!k1s_eligible<double>(*cfg, L, MB_STAGE_MUSCLE, 0, in_dev, cpf_dev, dur_dev, nullptr, nullptr, forces_dev, ep.fatigue))
        return fail(MB_EUNSUPPORTED, "mb_step_env needs 16-byte aligned state arrays and peripheral fatigue on");
    return (use_flat(*cfg, L, S, 0) ? launch_k1f_f64 : launch_k1s_f64)(*cfg, params_dev, rows, L, S, seg_off_dev, seg_div_dev, in_dev,
                                                                    0, cpf_dev, dur_dev, forces_dev, totals_dev, ep, step_size, st);
}

extern "C" int mb_step_env(const MbConfig* cfg, const void* params_dev, int64_t rows, int64_t L, int32_t S,
                           const int32_t* seg_off_dev, const double* seg_div_dev, const void* in_dev,
                           double* cpf_dev, double* dur_dev, void* forces_dev, void* totals_dev,
                           const void* out_gain_dev, double* fatigue_dev, double step_size, void* stream) {
    K1sEpilogue ep;
    ep.out_gain = out_gain_dev;
    ep.fatigue = fatigue_dev;
    return step_env_impl(cfg, params_dev, rows, L, S, seg_off_dev, seg_div_dev, in_dev, cpf_dev, dur_dev, forces_dev,
                         totals_dev, ep, step_size, stream);
}

extern "C" int mb_step_env_hill(const MbConfig* cfg, const void* params_dev, int64_t rows, int64_t L, int32_t S,
                                const int32_t* seg_off_dev, const double* seg_div_dev, const void* in_dev,
                                double* cpf_dev, double* dur_dev, void* forces_dev, void* totals_dev,
                                const void* out_gain_dev, double* fatigue_dev, const MbHill* hill, double step_size,
                                void* stream) {
    if (!hill || !hill->rest_dev || !hill->cur_dev || !hill->prev_dev)
        return fail(MB_EINVAL, "mb_step_env_hill: hill and its three length arrays are required");
    if (!(hill->max_velocity > 0)) return fail(MB_EINVAL, "mb_step_env_hill: max_velocity must be > 0");
    K1sEpilogue ep;
    ep.out_gain = out_gain_dev;
    ep.fatigue = fatigue_dev;
    ep.hill_rest = hill->rest_dev; ep.hill_cur = hill->cur_dev; ep.hill_prev = hill->prev_dev;
    ep.hill_width = hill->curve_width_factor; ep.hill_opt = 1.10;        // hill_type.py:38 pins the optimum at 1.10
    ep.hill_vmax = hill->max_velocity; ep.hill_ecc = hill->max_eccentric_multiple;
    return step_env_impl(cfg, params_dev, rows, L, S, seg_off_dev, seg_div_dev, in_dev, cpf_dev, dur_dev, forces_dev,
                         totals_dev, ep, step_size, stream);
}

extern "C" int mb_peripheral_fatigue(const MbConfig* cfg, const void* params_dev, int64_t rows, int64_t L, int32_t S,
                                     const int32_t* seg_off_dev, const double* seg_div_dev, const double* cpf_dev, double* fatigue_dev,
                                     void* stream) {
    if (int rc = check_cfg(cfg)) return rc;
    if (rows <= 0 || L <= 0 || S <= 0 || !params_dev || !cpf_dev || !fatigue_dev) return fail(MB_EINVAL, "mb_peripheral_fatigue: bad argument");
    if (S > 1 && !seg_off_dev) return fail(MB_EINVAL, "mb_peripheral_fatigue: seg_off required when S > 1");
    int64_t blocks = rows * S;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (cfg->precision == MB_F32)
        k3_fatigue<float4><<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(
            (const float4*)params_dev, cfg->param_rows > 1 ? (int64_t)kGroupsF32 * L : 0, rows, L, S, seg_off_dev, seg_div_dev,
            cpf_dev, fatigue_dev, cfg->output_divisor);
    else
        k3_fatigue<double2><<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(
            (const double2*)params_dev, cfg->param_rows > 1 ? (int64_t)kGroupsF64 * L : 0, rows, L, S, seg_off_dev, seg_div_dev,
            cpf_dev, fatigue_dev, cfg->output_divisor);
    MB_CUDA_OK(cudaGetLastError());
    return MB_OK;
}

// ---- fused geometry ---------------------------------------------------------------------
static const int kWarpUnitsMax = 256;      // warp-per-muscle kernel (k2u_f32): 4 pairs of units per lane
static const int kFusedMaxUnits = 480;     // threads per block <= 480 (15 warps), 2 blocks per SM in FP32 mode
static const int kWideMaxUnits = 4096;     // block-per-muscle kernel: 512 threads x 4 pairs of units

static int fused_geometry(const MbConfig& c, int64_t M, int32_t n, int32_t K, int32_t upt, FusedGeom* g,
                          bool forces = false, bool state_aligned16 = true) {
    if (n < 2 || n > kFusedMaxUnits)
        return fail(MB_EUNSUPPORTED, "fused kernel supports 2 <= n <= %d units per muscle (got %d)", kFusedMaxUnits, n);
    if (M <= 0 || K <= 0) return fail(MB_EINVAL, "M and K must be positive");
    const bool f32 = c.precision == MB_F32;
    const int esz = f32 ? 4 : 8;
    g->kernel = FK_BLOCK;
    const bool per_row = c.param_rows > 1;
    if (per_row && c.param_rows != M) return fail(MB_EINVAL, "per-instance tables need param_rows == M");
    // shared tables, 129..256 units: the warp-per-muscle kernel beats the block kernel without central fatigue (measured:
    // StandardMuscle n = 169 1.14e12 vs 9.2e11, n = 256 1.28e12 vs 1.20e12; Potvin n = 200 with its adaptation state: equal)
    const bool units_mid = f32 && !per_row && upt == 0 && n > 128 && n <= kWarpUnitsMax &&
                           (knobs().fused_units > 0 || (knobs().fused_units < 0 && !c.central_fatigue));
    if (f32 && ((per_row && n <= 128) || units_mid)) {
        // one warp per muscle, pairs of units: per-instance tables, and mid-sized muscles
        g->kernel = FK_WARP_UNITS;
        g->U = 1; g->q = kWarpsPerBlock; g->G = kWarpsPerBlock; g->Gp = g->G; g->T = kWTile;
        g->ns = n; g->npg = 0; g->ptp = 0;
        g->threads = kWarpsPerBlock * 32;
        g->smem = kWarpsPerBlock * (kWTile * kUPitch + 2 * kWTile) * (int)sizeof(float);
        g->blocks = (M + g->G - 1) / g->G;
        return MB_OK;
    }
    if (!f32 && upt == 0 && n >= 12 && n <= 128 && c.firing_gain == 1.0 && !knobs().f64_block) {
        // FP64 mode, 12..128 units: one warp per muscle, no cross-warp reduction; every muscle may have its own table
        g->kernel = FK_WARP_F64;
        g->U = 1; g->q = kW64Warps; g->G = kW64Warps; g->Gp = g->G; g->T = kW64Tile;
        g->ns = n; g->npg = 0; g->ptp = 0;
        g->threads = kW64Warps * 32;
        g->smem = kW64Warps * (kW64Tile * kW64Pitch + 2 * kW64Tile) * (int)sizeof(double);
        g->blocks = (M + g->G - 1) / g->G;
        return MB_OK;
    }
    if (per_row) upt = 1;                                    // block kernels: a thread's instances share one table
    if (f32 && upt == 0 && n >= 12 && n <= 128 && !knobs().fused_block) {
        // FP32 mode, 12..128 units: one warp per pair of muscles, no cross-warp reduction (below 32 units
        // lanes idle, but the block kernel's 100+ muscles per block do worse: 3.8e11 at n = 16)
        g->kernel = FK_WARP_PAIR;
        g->U = 2; g->q = kWarpsPerBlock; g->G = 2 * kWarpsPerBlock; g->Gp = g->G; g->T = kWTile;
        g->ns = n; g->npg = 0; g->ptp = 0;
        g->threads = kWarpsPerBlock * 32;
        g->smem = kWarpsPerBlock * (kWTile * kWPitch + 2 * kWTile) * (int)sizeof(float2);
        g->blocks = (M + g->G - 1) / g->G;
        // very short windows without force capture: resident warps walk the pairs and prefetch the next pair's state.
        // Measured (Potvin 120 x 65,536, us per window, one-pair-per-warp / persistent): K = 8: 162 / 128, K = 16: 199 / 173,
        // K = 32: 264 / 266, K = 64: 418 / 464 -- the persistent loop is ~9 % slower per step, so it only pays while the
        // window is dominated by the state round trip
        const int persist = knobs().fused_persist;
        const int64_t resident = 2LL * sm_count();
        if (!forces && state_aligned16 && c.param_rows <= 1 && g->blocks > resident &&
            (persist > 0 || (persist < 0 && K <= 24))) {
            g->kernel = FK_WARP_PAIR_PERSIST;
            g->smem += kWarpsPerBlock * 4 * ((n + 1) & ~1) * (int)sizeof(double);
            g->blocks = resident;
        }
        return MB_OK;
    }
    const int ns = (n + 7) / 8 * 8;                          // threads per muscle group (stage-1 parts are 8 wide)
    int q = kFusedMaxUnits / ns;
    if (upt != 0 && upt != 1 && upt != 2 && upt != 4) return fail(MB_EINVAL, "units_per_thread must be 0, 1, 2 or 4");
    int U = upt;
    const bool no_u4 = forces || (f32 && c.fibers_kind == MB_FIBERS_PYMUSCLE && c.central_fatigue);
    if (U == 4 && no_u4) U = 2;
    if (U == 0) {
        // largest U that still leaves >= 2 blocks per SM; small batches spread out instead
        // (four instances per thread with the decimated force output, or with PyMuscle fibers + central fatigue in
        //  FP32 mode, do not fit the block kernels' register budget: those shapes run two per thread)
        U = no_u4 ? 2 : 4;
        while (U > 1 && (M + (int64_t)q * U - 1) / ((int64_t)q * U) < 2LL * sm_count()) U >>= 1;
    }
    // a block should not own more muscles than exist
    while (q > 1 && (int64_t)(q - 1) * U >= M) --q;
    g->U = U; g->q = q; g->G = q * U; g->ns = ns; g->npg = ns / 8;
    const int vl = 16 / esz;
    g->Gp = (g->G + vl - 1) / vl * vl;
    g->threads = (q * ns + 31) / 32 * 32;
    const int warps = g->threads / 32;
    const int vec_bytes = U * esz;
    g->ptp = (warps * 4) | 1;                                // odd PT row pitch (vectors)
    int T = kTile;
    if (T > K) T = K;
    g->T = T;
    g->smem = 256 + kExcSlots * T * g->Gp * esz + warps * kTile * kFtPitch * vec_bytes +
              kPtSlots * kTile * g->ptp * vec_bytes;
    if (g->smem > 220 * 1024) return fail(MB_EUNSUPPORTED, "fused kernel: tile does not fit in shared memory");
    g->blocks = (M + g->G - 1) / g->G;
    return MB_OK;
}

extern "C" int mb_fused_geometry(const MbConfig* cfg, int64_t M, int32_t n, int32_t K, int32_t units_per_thread,
                                 int32_t* muscles_per_block, int32_t* threads, int64_t* blocks, int32_t* smem_bytes) {
    if (int rc = check_cfg(cfg)) return rc;
    FusedGeom g;
    if (int rc = fused_geometry(*cfg, M, n, K, units_per_thread, &g)) return rc;
    if (muscles_per_block) *muscles_per_block = g.G;
    if (threads) *threads = g.threads;
    if (blocks) *blocks = g.blocks;
    if (smem_bytes) *smem_bytes = g.smem;
    return MB_OK;
}

// The fused window, every option (MbFusedArgs); mb_step_fused / mb_step_fused_gather are views of it.
extern "C" int mb_step_fused_ex(const MbConfig* cfg, const void* params_dev, const MbFusedArgs* w, double step_size,
                                void* stream) {
    if (int rc = check_cfg(cfg)) return rc;
    if (!w) return fail(MB_EINVAL, "mb_step_fused_ex: args is NULL");
    if (!params_dev || !w->exc_dev || !w->cpf_dev) return fail(MB_EINVAL, "mb_step_fused: params/exc/cpf is NULL");
    if (cfg->central_fatigue && !w->dur_dev) return fail(MB_EINVAL, "mb_step_fused: dur is NULL");
    if (w->rows <= 0 || w->L <= 0 || w->K <= 0) return fail(MB_EINVAL, "mb_step_fused: rows, L and K must be positive");
    const int32_t S = w->S > 0 ? w->S : 1;
    if (S > 1 && !w->seg_off_dev) return fail(MB_EINVAL, "mb_step_fused: seg_off required when S > 1");
    if (w->L > INT32_MAX) return fail(MB_EINVAL, "mb_step_fused: L too large");
    const int64_t M = w->rows * S;
    const int32_t n = (int32_t)w->L, K = w->K;
    const int32_t hold = w->exc_hold > 0 ? w->exc_hold : 1, tev = w->totals_every > 0 ? w->totals_every : 1;
    if (w->forces_every < 0 || (w->forces_every > 0 && !w->forces_dec_dev)) return fail(MB_EINVAL, "mb_step_fused: forces_every without buffer");
    if (w->exc_stride_k != 0 && w->exc_stride_k < M) return fail(MB_EINVAL, "mb_step_fused: exc_stride_k < M");
    if (w->totals_dev && w->totals_stride_k < M) return fail(MB_EINVAL, "mb_step_fused: totals_stride_k < M");
    // ---- which kernel: rows of several muscles and muscles wider than the block kernel's 480 units run one block
    // per muscle (FP32 mode; up to 4,096 units per muscle)
    const int32_t maxn = S > 1 ? (w->max_segment_units > 0 ? w->max_segment_units : n) : n;
    const bool wide = cfg->precision == MB_F32 && maxn >= 2 && maxn <= kWideMaxUnits &&
                      (S > 1 || n > kFusedMaxUnits || (knobs().fused_wide && n > 128));
    if (!wide && S > 1)
        return fail(MB_EUNSUPPORTED, "mb_step_fused: rows of several muscles need FP32 mode and muscles of at most %d units",
                    kWideMaxUnits);
    FusedGeom g;
    if (wide && S > 1 && maxn <= kWarpUnitsMax) {
        // rows of SHORT muscles: one warp per muscle (k2u_f32's arrangement, pairs of units, up to 256 units)
        g.kernel = FK_WARP_UNITS;
        g.U = 1; g.q = kWarpsPerBlock; g.G = kWarpsPerBlock; g.Gp = g.G; g.T = kWTile;
        g.ns = maxn; g.npg = 0; g.ptp = 0;
        g.threads = kWarpsPerBlock * 32;
        g.smem = kWarpsPerBlock * (kWTile * kUPitch + 2 * kWTile) * (int)sizeof(float);
        g.blocks = (M + g.G - 1) / g.G;
    } else if (wide) {
        if (cfg->param_rows > 1 && (S != 1 || cfg->param_rows != M)) return fail(MB_EINVAL, "per-instance tables need a uniform batch with param_rows == M");
        g.kernel = FK_WIDE;
        g.threads = maxn <= 2048 ? 256 : 512;
        g.U = (maxn + 2 * g.threads - 1) / (2 * g.threads);  // pairs per thread
        const int nw = g.threads / 32;
        g.q = 1; g.G = 1; g.Gp = 1; g.T = kXTile; g.ns = maxn; g.npg = 0; g.ptp = 0;
        g.smem = (nw * (kXTile * kXPitch + 2 * kXTile) + 2 * kXTile * (nw + 1)) * (int)sizeof(float);
        g.blocks = M;
    } else if (int rc = fused_geometry(*cfg, M, n, K, w->units_per_thread, &g, w->forces_every > 0,
                                       ((uintptr_t)w->cpf_dev % 16 == 0) && ((uintptr_t)w->dur_dev % 16 == 0))) {
        return rc;
    }
    const bool f32 = cfg->precision == MB_F32;
    const int esz = f32 ? 4 : 8;

    K2Args a;
    memset(&a, 0, sizeof(a));
    a.params = params_dev; a.param_stride = cfg->param_rows > 1 ? (int64_t)(f32 ? kGroupsF32 : kGroupsF64) * n : 0;
    a.M = M; a.n = n; a.ns = g.ns; a.q = g.q; a.K = K; a.T = g.T; a.G = g.G; a.Gp = g.Gp;
    a.npg = g.npg; a.ptp = g.ptp;
    a.exc = w->exc_dev; a.exc_stride_k = w->exc_stride_k;
    a.exc_hold = hold; a.totals_every = tev;
    if (w->exc_stride_k == 0) {
        a.exc_mode = 2;
    } else {
        // bulk copies need 16-byte aligned rows: base pointer, row pitch, block offset and row length
        const bool aligned = ((uintptr_t)w->exc_dev % 16 == 0) && ((w->exc_stride_k * esz) % 16 == 0) &&
                             ((g.G * esz) % 16 == 0) && ((M * esz) % 16 == 0);
        a.exc_mode = (aligned && !knobs().fused_no_bulk) ? 0 : 1;
    }
    a.cpf = w->cpf_dev; a.dur = w->dur_dev; a.totals = w->totals_dev; a.totals_stride_k = w->totals_stride_k;
    const int32_t n_peers = w->n_peers;
    if (n_peers < 0 || n_peers > kMaxPeers) return fail(MB_EINVAL, "mb_step_fused_gather: at most %d peer buffers", kMaxPeers);
    if (n_peers > 0 && (!w->totals_dev || !w->peer_totals_dev)) return fail(MB_EINVAL, "mb_step_fused_gather: totals / peer list is NULL");
    a.n_peers = n_peers;
    a.peer_stride_k = w->peer_stride_k;
    a.peer_stride_pair = w->peer_stride_pair;
    a.peer_stride_odd = w->peer_stride_odd > 0 ? w->peer_stride_odd : 1;
    a.peer_multicast = w->peer_multicast != 0;
    if (a.peer_multicast && n_peers != 1) return fail(MB_EINVAL, "mb_step_fused_gather: a multicast gather takes exactly one (multicast) buffer");
    if (n_peers > 0 && (w->peer_stride_k <= 0 || w->peer_stride_pair <= 0)) return fail(MB_EINVAL, "mb_step_fused_gather: peer strides must be positive");
    for (int i = 0; i < n_peers; ++i) {
        if (!w->peer_totals_dev[i]) return fail(MB_EINVAL, "mb_step_fused_gather: peer buffer %d is NULL", i);
        a.totals_peer[i] = w->peer_totals_dev[i];
    }
    a.forces_dec = w->forces_dec_dev; a.mask_dec = w->mask_dec_dev; a.forces_every = w->forces_every;
    a.forces_last = w->forces_last_dev;
    a.peripheral = cfg->peripheral_fatigue != 0;
    a.dt = step_size;
    a.adk_d = -kLog2e / cfg->adaptation_time_constant;
    a.max_dur = cfg->max_duration;
    a.decay_d = std::exp(-step_size / cfg->adaptation_time_constant);
    a.dtk = (float)(a.dt * a.adk_d);
    a.decay = (float)a.decay_d;
    a.dm1 = a.decay - 1.0f;                                  // exact in float (Sterbenz)
    a.sf = scal_f(*cfg);
    a.sd = scal_d(*cfg);
    a.ec = exp_consts();

    // central-fatigue variant (see k2_fused_f32): the fast path needs the packed table's
    // MB_PROP_ADAPT_NONNEG guarantee and a duration clamp that exp(-dur/tau) cannot see
    int cen = 0;
    if (cfg->central_fatigue) {
        const bool clamp_invisible = cfg->max_duration / cfg->adaptation_time_constant > 110.0;
        const bool nonneg = (cfg->table_props & MB_PROP_ADAPT_NONNEG) != 0;
        cen = (clamp_invisible && nonneg && !knobs().fused_general) ? 1 : 2;
    }
    const int fib = cfg->fibers_kind;
    const bool forces = w->forces_every > 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (g.kernel == FK_WIDE || g.kernel == FK_WARP_UNITS) {
        K2xArgs xa;
        xa.k = a;
        xa.L = w->L; xa.total_units = w->rows * w->L; xa.S = S;
        xa.seg_off = S > 1 ? w->seg_off_dev : nullptr;
        xa.seg_div = S > 1 ? w->seg_div_dev : nullptr;
        return g.kernel == FK_WIDE ? launch_k2x_f32(xa, g, fib, cen, forces, st) : launch_k2u_f32(xa, g, fib, cen, forces, st);
    }
    if (g.kernel == FK_WARP_PAIR || g.kernel == FK_WARP_PAIR_PERSIST) return launch_k2w_f32(a, g, fib, cen, forces, st);
    if (f32) return launch_k2_block_f32(a, g, fib, cen, forces, st);
    // FP64: the fast path additionally needs the sigmoid's exponent bounded (no clamp in exp_neg_fast);
    // without central fatigue the clamp is kept (CENTRAL == 0 instantiates it)
    if (cen == 1 && !(cfg->table_props & MB_PROP_EXP_BOUNDED)) cen = 2;
    if (g.kernel == FK_WARP_F64) return launch_k2w_f64(a, g, fib, cen, forces, st);
    return launch_k2_block_f64(a, g, fib, cen, forces, st);
}

extern "C" int mb_step_fused(const MbConfig* cfg, const void* params_dev, int64_t M, int32_t n, int32_t K,
                             const void* exc_dev, int64_t exc_stride_k, double* cpf_dev, double* dur_dev,
                             void* totals_dev, int64_t totals_stride_k, void* forces_dec_dev, uint8_t* mask_dec_dev,
                             int32_t forces_every, void* forces_last_dev, double step_size,
                             int32_t units_per_thread, void* stream) {
    MbFusedArgs w;
    memset(&w, 0, sizeof(w));
    w.rows = M; w.L = n; w.S = 1; w.K = K;
    w.exc_dev = exc_dev; w.exc_stride_k = exc_stride_k; w.exc_hold = 1;
    w.cpf_dev = cpf_dev; w.dur_dev = dur_dev;
    w.totals_dev = totals_dev; w.totals_stride_k = totals_stride_k; w.totals_every = 1;
    w.forces_dec_dev = forces_dec_dev; w.mask_dec_dev = mask_dec_dev; w.forces_every = forces_every;
    w.forces_last_dev = forces_last_dev;
    w.units_per_thread = units_per_thread;
    return mb_step_fused_ex(cfg, params_dev, &w, step_size, stream);
}

extern "C" int mb_step_fused_gather(const MbConfig* cfg, const void* params_dev, int64_t M, int32_t n, int32_t K,
                                    const void* exc_dev, int64_t exc_stride_k, double* cpf_dev, double* dur_dev,
                                    void* totals_dev, int64_t totals_stride_k, void* const* peer_totals_dev,
                                    int32_t n_peers, int64_t peer_stride_k, int64_t peer_stride_pair,
                                    void* forces_last_dev, double step_size, void* stream) {
    MbFusedArgs w;
    memset(&w, 0, sizeof(w));
    w.rows = M; w.L = n; w.S = 1; w.K = K;
    w.exc_dev = exc_dev; w.exc_stride_k = exc_stride_k; w.exc_hold = 1;
    w.cpf_dev = cpf_dev; w.dur_dev = dur_dev;
    w.totals_dev = totals_dev; w.totals_stride_k = totals_stride_k; w.totals_every = 1;
    w.forces_last_dev = forces_last_dev;
    w.peer_totals_dev = peer_totals_dev; w.n_peers = n_peers;
    w.peer_stride_k = peer_stride_k; w.peer_stride_pair = peer_stride_pair;
    return mb_step_fused_ex(cfg, params_dev, &w, step_size, stream);
}
