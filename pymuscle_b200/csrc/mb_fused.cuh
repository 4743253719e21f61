// mb_fused.cuh -- K2: K fused steps per launch, state in registers.
//
// Uniform batch: M muscles x n units.  A block owns G = q*U consecutive muscles;
// thread t < q*n handles unit (t % n) of muscles (t / n)*U .. +U-1, so its U
// unit-instances share one set of per-unit parameters (held in registers for the
// whole window) and give U independent dependency chains.
//
// Shared memory:
//   exc tile   [2][T][Gp]     per-step, per-muscle excitations, filled by cp.async.bulk
//                             (one 1-D bulk copy per step row, mbarrier complete_tx),
//                             double-buffered one tile (T steps) ahead; a thread reads
//                             its U excitations with one vector LDS per step
//   force tile [T][q][n][U]   unit forces of the current tile, one vector STS per thread
//                             and step; after T steps every (step, q) row is summed over
//                             the unit axis by P lanes (vector LDS, shuffle tree), so the
//                             per-step, per-muscle reduction costs one barrier per T steps
// HBM traffic per window: state in/out (32 B/unit) plus excitations and totals
// (2*sizeof(T)/n B per unit-step) -- see DESIGN.md.
#pragma once
#include "mb_core.cuh"

namespace mb {

struct K2Args {
    const void* params;
    int64_t M;
    int32_t n, q, K, T, P, G, Gp;
    const void* exc;
    int64_t exc_stride_k;
    int32_t exc_mode;            // 0 = bulk copy, 1 = per-element cp.async, 2 = constant over the window
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
__device__ __forceinline__ void cp_async_4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// ---- excitation staging (shared by the FP32 and FP64 kernels) -------------------------
template <typename T>
struct ExcStager {
    const K2Args& a;
    T* tiles;              // [2][T][Gp]
    uint64_t* bars;        // [2]
    int64_t m0;
    int gcount, tid, nthreads;

    __device__ void init() {
        if (a.exc_mode == 0 && tid == 0) {
            mbar_init(&bars[0], 1);
            mbar_init(&bars[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        // columns of muscles beyond the batch read as 0 excitation (never recruited)
        for (int i = tid; i < 2 * a.T * a.Gp; i += nthreads) tiles[i] = T(0);
        __syncthreads();
        if (a.exc_mode == 2) {
            const T* src = (const T*)a.exc + m0;
            for (int i = tid; i < a.T * a.Gp; i += nthreads) {
                const int g = i % a.Gp;
                if (g < gcount) tiles[i] = src[g];
            }
            __syncthreads();
        }
    }
    // start filling buffer `buf` with the excitations of steps [k0, k0+steps)
    __device__ void issue(int buf, int k0, int steps) {
        if (a.exc_mode == 2) return;
        T* dst = tiles + (size_t)buf * a.T * a.Gp;
        const T* src = (const T*)a.exc + (int64_t)k0 * a.exc_stride_k + m0;
        if (a.exc_mode == 0) {
            if (tid == 0) {
                const uint32_t row_bytes = (uint32_t)(gcount * sizeof(T));
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&bars[buf], row_bytes * steps);
                for (int s = 0; s < steps; ++s)
                    bulk_g2s(dst + s * a.Gp, src + (int64_t)s * a.exc_stride_k, row_bytes, &bars[buf]);
            }
        } else {
            for (int i = tid; i < steps * a.G; i += nthreads) {
                const int s = i / a.G, g = i - s * a.G;
                if (g < gcount) {
                    if (sizeof(T) == 4) cp_async_4(dst + s * a.Gp + g, src + (int64_t)s * a.exc_stride_k + g);
                    else                cp_async_8(dst + s * a.Gp + g, src + (int64_t)s * a.exc_stride_k + g);
                }
            }
        }
    }
    // block until buffer `buf` (its `use`-th fill, counting from 0) is readable
    __device__ void wait(int buf, int use) {
        if (a.exc_mode == 0) {
            mbar_wait(&bars[buf], (uint32_t)(use & 1));
        } else if (a.exc_mode == 1) {
            cp_async_wait_all();
            __syncthreads();
        }
    }
};

// U-wide vector of T in shared memory (one LDS/STS of 4*U or 8*U bytes)
template <typename T, int U> struct alignas(sizeof(T) * U) VecU { T v[U]; };

// ---- per-step, per-muscle totals from the force tile ----------------------------------
// Rows = steps*q, each n vectors of U forces (U muscles side by side).  P lanes
// (a power of two <= 32) stride over a row, accumulate U partial sums each, and a
// butterfly leaves the row's U totals in every lane of the group.
template <typename T, int U>
__device__ __forceinline__ void tile_row_sums(const K2Args& a, const T* ft, int steps, int k0, int64_t m0,
                                              int gcount, int tid, int nthreads, T scale_mul, T scale_div) {
    using V = VecU<T, U>;
    const int P = a.P, n = a.n;
    const int ntasks = steps * a.q * P;
    T* totals = (T*)a.totals;
    for (int base = 0; base < ntasks; base += nthreads) {
        const int task = base + tid;
        const bool active = task < ntasks;
        const int row = active ? task / P : 0;
        const int part = task & (P - 1);
        const V* rp = (const V*)ft + (size_t)row * n;
        T acc[U];
#pragma unroll
        for (int j = 0; j < U; ++j) acc[j] = 0;
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
                if (g < gcount) totals[(int64_t)(k0 + s) * a.totals_stride_k + m0 + g] = acc[j] * scale_mul / scale_div;
            }
        }
    }
}

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
template <int FIB, int CENTRAL, int U, bool FORCES>
__global__ void __launch_bounds__(512, 2) k2_fused_f32(const K2Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using V = VecU<float, U>;
    uint64_t* bars = (uint64_t*)smem_raw;
    float* exc_s = (float*)(smem_raw + 16);
    float* ft = exc_s + 2 * a.T * a.Gp;

    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int n = a.n;
    const bool has_unit = tid < a.q * n;
    const int u = has_unit ? tid % n : 0;
    const int tq = has_unit ? tid / n : 0;
    const int64_t m0 = (int64_t)blockIdx.x * a.G;
    const int gcount = (int)min((int64_t)a.G, a.M - m0);

    ExcStager<float> stager{a, exc_s, bars, m0, gcount, tid, nthreads};
    stager.init();

    const UnitF p = load_unit((const float4*)a.params, n, u);
    const ScalF sc = a.sf;
    const float fatdt = a.peripheral ? (float)((double)p.fat * a.dt) : 0.0f;
    const float rdidt = a.peripheral ? (float)((double)p.rdi * a.dt) : 0.0f;
    const float dtk = (float)(a.dt * a.adk_d);

    float hi[U], lo[U], cpf[U], cnt[U], arg0[U], room[U], F[U];
    bool valid[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
        const int g = tq * U + j;
        valid[j] = has_unit && g < gcount;
        const int64_t idx = (m0 + g) * n + u;
        const double c = valid[j] ? a.cpf[idx] : (double)p.ptf;
        const double d0 = (valid[j] && CENTRAL) ? a.dur[idx] : 0.0;
        hi[j] = (float)c;
        lo[j] = (float)(c - (double)hi[j]);
        cpf[j] = hi[j] + lo[j];
        room[j] = (p.ptf - hi[j]) + p.ptf_lo;            // peak - hi: the upper bound of lo (pmf:104-105)
        cnt[j] = 0.0f;
        arg0[j] = (float)(d0 * a.adk_d);
        F[j] = 0.0f;
    }

    const int ntiles = (a.K + a.T - 1) / a.T;
    const int step_stride = a.q * n;                 // vectors per step in the force tile
    stager.issue(0, 0, min(a.T, a.K));
    for (int tile = 0; tile < ntiles; ++tile) {
        const int k0 = tile * a.T;
        const int steps = min(a.T, a.K - k0);
        const int buf = (a.exc_mode == 2) ? 0 : (tile & 1);
        stager.wait(buf, tile >> 1);
        if (tile + 1 < ntiles) stager.issue(buf ^ 1, k0 + a.T, min(a.T, a.K - k0 - a.T));

        if (has_unit) {
            const float* es = exc_s + (size_t)buf * a.T * a.Gp + tq * U;
            V* fp = (V*)ft + tid;                    // == tq*n + u
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
                    const float r = pool_rate_raw(p, sc, e.v[j], on);
                    float fr;
                    if (CENTRAL == 0) {
                        fr = on ? r : 0.0f;
                    } else {
                        float arg = fmaf(cnt[j], dtk, arg0[j]);
                        if (on) cnt[j] += 1.0f;                       // pool:196-197
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
                    F[j] = fibers_force(p, sc, fr, cpf[j], nf);
                    fv.v[j] = F[j];
                    if (FORCES) {
                        if (write_now && valid[j]) {
                            const int64_t idx = dec_base + (m0 + tq * U + j) * n + u;
                            ((float*)a.forces_dec)[idx] = F[j];
                            if (a.mask_dec) a.mask_dec[idx] = on ? 1 : 0;
                        }
                    }
                    // capacity update on the compensated pair (fibers:110-114, pmf:94-105)
                    float l = fmaf(nf, -fatdt, lo[j]);
                    if (FIB == MB_FIBERS_PYMUSCLE) {
                        const float gap = room[j] - lo[j];                 // peak - capacity
                        l = fmaf((nf <= 0.0f) ? gap : 0.0f, rdidt, l);     // pmf:121-141
                    }
                    l = fmaxf(l, -hi[j]);
                    if (FIB == MB_FIBERS_PYMUSCLE) l = fminf(l, room[j]);
                    lo[j] = l;
                    cpf[j] = hi[j] + l;
                }
                *fp = fv;
            }
            // renormalise (TwoSum) so that |lo| <= ulp(hi)/2 again
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
        __syncthreads();                                   // force tile complete
        tile_row_sums<float, U>(a, ft, steps, k0, m0, gcount, tid, nthreads, sc.inv_out, 1.0f);
        __syncthreads();                                   // force tile free again
    }

#pragma unroll
    for (int j = 0; j < U; ++j) {
        if (!valid[j]) continue;
        const int64_t idx = (m0 + tq * U + j) * n + u;
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
        if (a.forces_last) ((float*)a.forces_last)[idx] = F[j];
    }
}

// ---- FP64 mode ------------------------------------------------------------------------
template <int FIB, int CENTRAL, int U, bool FORCES>
__global__ void __launch_bounds__(512, 1) k2_fused_f64(const K2Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using V = VecU<double, U>;
    uint64_t* bars = (uint64_t*)smem_raw;
    double* exc_s = (double*)(smem_raw + 16);
    double* ft = exc_s + 2 * a.T * a.Gp;

    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int n = a.n;
    const bool has_unit = tid < a.q * n;
    const int u = has_unit ? tid % n : 0;
    const int tq = has_unit ? tid / n : 0;
    const int64_t m0 = (int64_t)blockIdx.x * a.G;
    const int gcount = (int)min((int64_t)a.G, a.M - m0);

    ExcStager<double> stager{a, exc_s, bars, m0, gcount, tid, nthreads};
    stager.init();

    const UnitD p = load_unit((const double2*)a.params, n, u);
    const ScalD sc = a.sd;
    const double fatdt = a.peripheral ? p.fat * a.dt : 0.0;
    const double rdidt = a.peripheral ? p.rdi * a.dt : 0.0;

    double cpf[U], dur[U], F[U];
    bool valid[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
        const int g = tq * U + j;
        valid[j] = has_unit && g < gcount;
        const int64_t idx = (m0 + g) * n + u;
        cpf[j] = valid[j] ? a.cpf[idx] : p.ptf;
        dur[j] = (valid[j] && CENTRAL) ? a.dur[idx] : 0.0;
        F[j] = 0.0;
    }

    const int ntiles = (a.K + a.T - 1) / a.T;
    const int step_stride = a.q * n;
    stager.issue(0, 0, min(a.T, a.K));
    for (int tile = 0; tile < ntiles; ++tile) {
        const int k0 = tile * a.T;
        const int steps = min(a.T, a.K - k0);
        const int buf = (a.exc_mode == 2) ? 0 : (tile & 1);
        stager.wait(buf, tile >> 1);
        if (tile + 1 < ntiles) stager.issue(buf ^ 1, k0 + a.T, min(a.T, a.K - k0 - a.T));

        if (has_unit) {
            const double* es = exc_s + (size_t)buf * a.T * a.Gp + tq * U;
            V* fp = (V*)ft + tid;
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
                    const double r = pool_rate(p, sc, e.v[j], on);
                    double fr = r;
                    if (CENTRAL) {
                        fr = pool_adapt(p, sc, r, dur[j]);
                        dur[j] = duration_update(dur[j], r > 0.0, a.dt, a.max_dur);
                    }
                    double nf;
                    F[j] = fibers_force(p, sc, fr, cpf[j], nf);
                    fv.v[j] = F[j];
                    if (FORCES) {
                        if (write_now && valid[j]) {
                            const int64_t idx = dec_base + (m0 + tq * U + j) * n + u;
                            ((double*)a.forces_dec)[idx] = F[j];
                            if (a.mask_dec) a.mask_dec[idx] = on ? 1 : 0;
                        }
                    }
                    cpf[j] = fatigue_update<FIB == MB_FIBERS_PYMUSCLE>(cpf[j], nf, fatdt, rdidt, p.ptf);
                }
                *fp = fv;
            }
        }
        __syncthreads();
        tile_row_sums<double, U>(a, ft, steps, k0, m0, gcount, tid, nthreads, 1.0, sc.out_div);
        __syncthreads();
    }

#pragma unroll
    for (int j = 0; j < U; ++j) {
        if (!valid[j]) continue;
        const int64_t idx = (m0 + tq * U + j) * n + u;
        if (a.peripheral) a.cpf[idx] = cpf[j];
        if (CENTRAL) a.dur[idx] = dur[j];
        if (a.forces_last) ((double*)a.forces_last)[idx] = F[j];
    }
}

}  // namespace mb
