// mb_fused.cuh -- K2: K fused steps per launch, unit state in registers.
//
// Uniform batch: M muscles x n units.  A block owns G = q*U consecutive muscles.
//
//   compute warps   thread t < q*n handles unit (t % n) of muscles (t / n)*U .. +U-1:
//                   its U unit-instances share one set of per-unit parameters (registers,
//                   loaded once per window as 128-bit vectors) and give U independent
//                   dependency chains.  Per step: one vector LDS (U excitations), the
//                   arithmetic of mb_core.cuh, one vector STS (U forces).
//   service warp    one extra warp per block.  It (a) stages the per-step excitations
//                   two tiles ahead with cp.async.bulk (1-D bulk copy per step row,
//                   mbarrier complete_tx) and (b) turns the force tile of the previous
//                   T steps into per-step, per-muscle totals while the compute warps are
//                   already producing the next tile.
//
// Shared memory:
//   exc ring    [4][T][Gp]           excitations, one slot per tile of T steps
//   force tile  [2][T][q][pitch][U]  unit forces, double-buffered; pitch >= n is chosen so
//                                    that the service warp's row-parallel 128-bit loads are
//                                    bank-conflict free
// Synchronisation is mbarrier-only inside the loop (exc_full[4], ft_full[2], ft_free[2]):
// compute warps never wait for each other, only for data.
//
// HBM traffic per window: state in/out (32 B/unit) plus excitations and totals
// (2*sizeof(T)/n B per unit-step) -- see DESIGN.md.
#pragma once
#include "mb_core.cuh"

namespace mb {

constexpr int kExcSlots = 4;

struct K2Args {
    const void* params;
    int64_t M;
    int32_t n, q, K, T, P, G, Gp, pitch, ncomp;
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
//              that the duration clamp cannot change exp(-dur/tau); both clamps drop out
//          2 = on, general path with both clamps (pool:202-206, :221)
template <int FIB, int CENTRAL, int U>
struct CoreF32 {
    using T = float;
    static constexpr bool kCentral = CENTRAL != 0;
    using ParamVec = float4;
    UnitF p;
    ScalF sc;
    float fatdt, rdidt, dtk;
    float hi[U], lo[U], cpf[U], cnt[U], arg0[U], room[U];

    __device__ __forceinline__ void load(const K2Args& a, int u) {
        p = load_unit((const float4*)a.params, a.n, u);
        sc = a.sf;
        fatdt = a.peripheral ? (float)((double)p.fat * a.dt) : 0.0f;
        rdidt = a.peripheral ? (float)((double)p.rdi * a.dt) : 0.0f;
        dtk = (float)(a.dt * a.adk_d);
    }
    __device__ __forceinline__ double rest_capacity() const { return (double)p.ptf + (double)p.ptf_lo; }
    __device__ __forceinline__ void init(int j, const K2Args& a, double c, double d0) {
        hi[j] = (float)c;
        lo[j] = (float)(c - (double)hi[j]);
        cpf[j] = hi[j] + lo[j];
        room[j] = (p.ptf - hi[j]) + p.ptf_lo;            // peak - hi: the upper bound of lo (pmf:104-105)
        cnt[j] = 0.0f;
        arg0[j] = (float)(d0 * a.adk_d);
    }
    // one step of instance j; returns the unit force
    __device__ __forceinline__ float step(int j, float e, bool& on) {
        const float r = pool_rate_raw(p, sc, e, on);
        float fr;
        if (CENTRAL == 0) {
            fr = on ? r : 0.0f;
        } else {
            float arg = fmaf(cnt[j], dtk, arg0[j]);
            if (on) cnt[j] += 1.0f;                                        // pool:196-197
            if (CENTRAL == 1) {
                fr = r - pool_adaptation(p, r, arg);
                fr = on ? fr : 0.0f;
            } else {
                arg = fmaxf(arg, sc.argmin);
                const float r0 = on ? r : 0.0f;
                fr = r0 - fmaxf(pool_adaptation(p, r0, arg), 0.0f);
            }
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
template <int FIB, int CENTRAL, int U>
struct CoreF64 {
    using T = double;
    static constexpr bool kCentral = CENTRAL != 0;
    using ParamVec = double2;
    UnitD p;
    ScalD sc;
    double fatdt, rdidt, dt, max_dur;
    double cpf[U], dur[U];

    __device__ __forceinline__ void load(const K2Args& a, int u) {
        p = load_unit((const double2*)a.params, a.n, u);
        sc = a.sd;
        fatdt = a.peripheral ? p.fat * a.dt : 0.0;
        rdidt = a.peripheral ? p.rdi * a.dt : 0.0;
        dt = a.dt;
        max_dur = a.max_dur;
    }
    __device__ __forceinline__ double rest_capacity() const { return p.ptf; }
    __device__ __forceinline__ void init(int j, const K2Args&, double c, double d0) {
        cpf[j] = c;
        dur[j] = d0;
    }
    __device__ __forceinline__ double step(int j, double e, bool& on) {
        const double r = pool_rate(p, sc, e, on);
        double fr = r;
        if (CENTRAL) {
            fr = pool_adapt(p, sc, r, dur[j]);
            dur[j] = duration_update(dur[j], r > 0.0, dt, max_dur);
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
// service warp: excitation staging + per-(step, muscle) totals
// ======================================================================================
template <typename T>
__device__ __forceinline__ void publish_exc(const K2Args& a, T* exc_s, uint64_t* exc_full, int tile, int64_t m0,
                                            int gcount, int lane) {
    const int k0 = tile * a.T;
    const int steps = min(a.T, a.K - k0);
    const int slot = tile & (kExcSlots - 1);
    T* dst = exc_s + (size_t)slot * a.T * a.Gp;
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

// Rows = steps*q, each `n` vectors of U forces (U muscles side by side), row pitch
// `pitch` vectors.  P lanes (a power of two) share a row, 32/P rows are summed per pass.
template <typename Core, int U>
__device__ __forceinline__ void tile_row_sums(const K2Args& a, const Core& core, const typename Core::T* ftb,
                                              int steps, int k0, int64_t m0, int gcount, int lane) {
    using T = typename Core::T;
    using V = VecU<T, U>;
    const int P = a.P, n = a.n;
    const int rows = steps * a.q;
    const int rows_per_pass = 32 / P;
    const int part = lane & (P - 1);
    T* totals = (T*)a.totals;
    for (int r0 = 0; r0 < rows; r0 += rows_per_pass) {
        const int row = r0 + lane / P;
        const bool active = row < rows;
        const V* rp = (const V*)ftb + (size_t)(active ? row : 0) * a.pitch;
        T acc[U];
#pragma unroll
        for (int j = 0; j < U; ++j) acc[j] = 0;
#pragma unroll 4
        for (int c = part; c < n; c += P) {
            const V v = rp[c];
#pragma unroll
            for (int j = 0; j < U; ++j) acc[j] += v.v[j];
        }
        for (int o = P >> 1; o > 0; o >>= 1) {
#pragma unroll
            for (int j = 0; j < U; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
        }
        if (active && part == 0 && totals) {
            const int s = row / a.q, tq = row - s * a.q;
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const int g = tq * U + j;
                if (g < gcount) totals[(int64_t)(k0 + s) * a.totals_stride_k + m0 + g] = core.out_scale(acc[j]);
            }
        }
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
    uint64_t* exc_full = (uint64_t*)smem_raw;            // [4]
    uint64_t* ft_full = exc_full + kExcSlots;            // [2]
    uint64_t* ft_free = ft_full + 2;                     // [2]
    T* exc_s = (T*)(smem_raw + 64);                      // [4][T][Gp]
    T* ft = exc_s + kExcSlots * a.T * a.Gp;              // [2][T][q][pitch][U]

    const int tid = threadIdx.x, lane = tid & 31;
    const int n = a.n;
    const bool is_service = tid >= a.ncomp;
    const int64_t m0 = (int64_t)blockIdx.x * a.G;
    const int gcount = (int)min((int64_t)a.G, a.M - m0);
    const int ntiles = (a.K + a.T - 1) / a.T;
    const int tile_vecs = a.T * a.q * a.pitch;           // vectors per force-tile buffer

    if (tid == 0) {
        for (int i = 0; i < kExcSlots; ++i) mbar_init(&exc_full[i], 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&ft_full[i], a.ncomp / 32);         // one arrival per compute warp
            mbar_init(&ft_free[i], 1);                    // one arrival by the service warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // columns of muscles beyond the batch read as 0 excitation (never recruited)
    for (int i = tid; i < kExcSlots * a.T * a.Gp; i += blockDim.x) exc_s[i] = T(0);
    __syncthreads();
    if (a.exc_mode == 2) {                                // constant excitation: fill slot 0 once
        const T* src = (const T*)a.exc + m0;
        for (int i = tid; i < a.T * a.Gp; i += blockDim.x) {
            const int g = i % a.Gp;
            if (g < gcount) exc_s[i] = src[g];
        }
        __syncthreads();
    }

    Core core;
    if (is_service) {
        // ------------------------------ service warp ------------------------------------
        core.load(a, 0);                                   // only the output scaling is used here
        if (a.exc_mode != 2)
            for (int t = 0; t < min(2, ntiles); ++t) publish_exc<T>(a, exc_s, exc_full, t, m0, gcount, lane);
        for (int tile = 0; tile < ntiles; ++tile) {
            if (a.exc_mode != 2 && tile + 2 < ntiles) publish_exc<T>(a, exc_s, exc_full, tile + 2, m0, gcount, lane);
            const int b = tile & 1;
            const int k0 = tile * a.T;
            mbar_wait(&ft_full[b], (uint32_t)((tile >> 1) & 1));
            tile_row_sums<Core, U>(a, core, (const T*)((const V*)ft + (size_t)b * tile_vecs), min(a.T, a.K - k0), k0,
                                   m0, gcount, lane);
            __syncwarp();
            if (lane == 0) mbar_arrive(&ft_free[b]);
        }
        return;
    }

    // ---------------------------------- compute warps -----------------------------------
    const bool has_unit = tid < a.q * n;
    const int u = has_unit ? tid % n : 0;
    const int tq = has_unit ? tid / n : 0;
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

    const int step_stride = a.q * a.pitch;               // vectors per step in the force tile
    for (int tile = 0; tile < ntiles; ++tile) {
        const int k0 = tile * a.T;
        const int steps = min(a.T, a.K - k0);
        const int slot = (a.exc_mode == 2) ? 0 : (tile & (kExcSlots - 1));
        const int b = tile & 1;
        if (a.exc_mode != 2) mbar_wait(&exc_full[slot], (uint32_t)((tile >> 2) & 1));
        if (tile >= 2) mbar_wait(&ft_free[b], (uint32_t)(((tile >> 1) - 1) & 1));
        if (has_unit) {
            const T* es = exc_s + (size_t)slot * a.T * a.Gp + tq * U;
            V* fp = (V*)ft + (size_t)b * tile_vecs + tq * a.pitch + u;
#pragma unroll 2
            for (int s = 0; s < steps; ++s, es += a.Gp, fp += step_stride) {
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
            core.end_tile();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ft_full[b]);          // release: this warp's part of the tile is written
    }

#pragma unroll
    for (int j = 0; j < U; ++j) {
        if (!valid[j]) continue;
        const int64_t idx = (m0 + tq * U + j) * n + u;
        core.store(j, a, idx);
        if (a.forces_last) ((T*)a.forces_last)[idx] = F[j];
    }
}

template <int FIB, int CENTRAL, int U, bool FORCES>
__global__ void __launch_bounds__(512, 2) k2_fused_f32(const K2Args a) {
    k2_body<CoreF32<FIB, CENTRAL, U>, U, FORCES>(a);
}

template <int FIB, int CENTRAL, int U, bool FORCES>
__global__ void __launch_bounds__(512, 1) k2_fused_f64(const K2Args a) {
    k2_body<CoreF64<FIB, CENTRAL, U>, U, FORCES>(a);
}

}  // namespace mb
