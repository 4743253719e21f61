// mb_stream.cuh -- K1s: the single-step kernel as an HBM streaming pipeline.
//
// The per-step state (capacities, and on-durations when central fatigue is on) is the
// only HBM traffic that matters: 8 B in / 8 B out per unit per state array plus the
// 4-8 B force write.  Arithmetic per byte is low, so the kernel is organised around
// keeping enough bytes in flight rather than around occupancy:
//
//   producer warp   walks the block's work list -- (muscle segment, chunk of R rows), cut
//                   into column tiles of W units -- and copies each tile's state rows
//                   global -> shared with cp.async.bulk (one bulk copy per row, issued by
//                   lane r; SASS UBLKCP), completion counted on the stage's `full` mbarrier
//                   (expect_tx).  NS stages deep: (NS-1) * R * W * 8 B per state array are
//                   in flight per block regardless of what the consumer warps are doing.
//                   The tile's packed unit parameters ([groups][W] 128-bit vectors, L2-resident
//                   table) ride along in the same stage, one bulk copy per group.
//   consumer warps  thread tx owns unit u0 + tx of the tile in all R rows: it waits for the
//                   stage, reads its parameters and its R state values from shared memory (no
//                   exposed global-load latency in the consumers), runs the step arithmetic
//                   (mb_core.cuh) and writes forces / new state straight to HBM with
//                   coalesced stores.  Per-muscle partial sums stay in registers across the
//                   segment's column tiles and are reduced once per work item (warp
//                   shuffles + one shared-memory stage, named barrier over the consumers).
//
// Bulk copies need 16-byte aligned global addresses: rows are 8-byte elements, so each row
// copy starts at the even element at or below the tile start and the consumers skip `shift`
// = 0 or 1 elements.  With an even row length the shift is the same in all rows; with an odd
// one it alternates from row to row (a chunk starts at a multiple of 8 rows), which costs the
// consumers nothing because the row index is a compile-time constant of the unrolled loop.
// The pool-only / fibers-only stages and the mask / rate outputs use the generic k1_step kernel.
#pragma once
#include "mb_core.cuh"
#include "mb_async.cuh"

namespace mb {

constexpr int kStreamRows = 8;          // R: rows advanced together by a block

template <typename T> struct StreamTraits;
// kMaxW / kMaxWCentral: widest column tile (= consumer threads per block) without / with central fatigue;
// kMinBlocks / kMinBlocksCentral: resident blocks the register budget is sized for; kDefStages: ring depth.
// FP64 mode carries twice the registers per row value: with the FP32 geometry (288 threads x 2 blocks = 96
// registers) its consumers spilled up to 300 bytes per thread, so it runs narrower blocks with more registers
// (128 without central fatigue, 168 with it) and a deeper ring to keep the bytes in flight.  FP32 mode with
// central fatigue needs 128 registers too (narrower tiles, three blocks).
template <> struct StreamTraits<float>  { using U = UnitF; using P = float4;  using S = ScalF; static constexpr int G = kGroupsF32;
                                          static constexpr int kMaxW = 256, kMinBlocks = 2;
                                          static constexpr int kMaxWCentral = 128, kMinBlocksCentral = 3;   // 128 registers: at 96 the central variants spill, and how much
                                                                                                            // depends on unrelated code (16..72 bytes seen, up to 2x in time)
                                          static constexpr int kDefStages = 2; };
template <> struct StreamTraits<double> { using U = UnitD; using P = double2; using S = ScalD; static constexpr int G = kGroupsF64;
                                          static constexpr int kMaxW = 128, kMinBlocks = 3;
                                          static constexpr int kMaxWCentral = 64, kMinBlocksCentral = 4;
                                          static constexpr int kDefStages = 3; };

// Optional per-muscle epilogue of the streaming step (mb_step_env): all arrays are [rows, S].
struct K1sEpilogue {
    const void* out_gain = nullptr;      // multiplier of the totals (mode's element type)
    double* fatigue = nullptr;           // 1 - sum(new capacity)/divisor (muscle.py:248-256)
    // Hill-type factors (hill_type.py:15-66) evaluated in the epilogue from muscle lengths; hill_rest == NULL: off
    const void* hill_rest = nullptr;     // resting lengths
    const void* hill_cur = nullptr;      // current lengths
    const void* hill_prev = nullptr;     // lengths one step earlier
    double hill_width = 17.33, hill_opt = 1.10, hill_vmax = 3.0, hill_ecc = 1.8;
};

// force-length x force-velocity multiplier of one muscle (hill_type.py:15-41, :44-66)
__device__ __forceinline__ double hill_gain(double rest, double cur, double prev, double dt, double width, double opt,
                                            double vmax, double ecc) {
    const double d = fabs(cur / rest - opt);
    const double fl = exp(width * (-1.0 * (d * d * d)));                  // hill_type.py:39-41
    const double vel = ((prev - cur) / rest) / dt;                        // hill_type.py:60
    const double e = exp((0.04 - vel / vmax) / 0.18);                     // hill_type.py:62-63
    return fl * (ecc - ecc / (1.0 + e));                                  // hill_type.py:65
}
__device__ __forceinline__ float hill_gain(float rest, float cur, float prev, double dt, double width, double opt,
                                           double vmax, double ecc) {
    const float d = fabsf(cur / rest - (float)opt);
    const float fl = __expf(-(float)width * (d * d * d));
    const float vel = ((prev - cur) / rest) / (float)dt;
    const float e = __expf((0.04f - vel / (float)vmax) / 0.18f);
    return fl * ((float)ecc - (float)ecc / (1.0f + e));
}

// stage row pitch in elements: compile-time (the instantiation's widest tile) in FP32 mode, W + 2 in FP64 mode
template <typename T, bool CENTRAL>
__host__ __device__ __forceinline__ constexpr int stream_row_pitch(int W) {
    return sizeof(T) == 4 ? (CENTRAL ? StreamTraits<T>::kMaxWCentral : StreamTraits<T>::kMaxW) + 2 : W + 2;
}

template <typename T>
struct K1sArgs {
    const typename StreamTraits<T>::P* params;
    int64_t param_stride;        // 0: one shared table; else vectors between the tables of consecutive rows (per-instance)
    int64_t rows, L;
    int32_t S, W, NS;
    int32_t seg_in_smem;         // 1: seg_off / seg_div are copied to shared memory at kernel start
    const int32_t* seg_off;
    const double* seg_div;
    int32_t peripheral;
    const T* in;                 // [rows, S] one input per muscle, or [rows, L] one per unit (in_per_unit)
    int32_t in_per_unit;         // 1: the ndarray form of Muscle.step (muscle.py:67, :89): staged like the state rows
    double* cpf;
    double* dur;
    T* forces;                   // may be NULL
    T* totals;                   // may be NULL
    const T* out_gain;           // [rows, S] multiplier of the totals (e.g. Hill-type factors); may be NULL
    double* fatigue;             // [rows, S] 1 - sum(new capacity)/divisor (muscle.py:248-256); may be NULL
    const T* hill_rest;          // [rows, S] Hill-type epilogue (K1sEpilogue); NULL = off
    const T* hill_cur;
    const T* hill_prev;
    double hill_width, hill_opt, hill_vmax, hill_ecc;
    double dt, adk_d, max_dur;
    double fat_div;              // output divisor of a uniform batch (seg_div == NULL)
    typename StreamTraits<T>::S sc;
};

// one unit, one step; state in, state out (no memory access)
// FP32 mode: the capacity CHANGE of the step -- decay -fat*nf*dt (fibers:110-111) or, for a unit at rest, recovery
// rec*((peak - c)/peak)*dt (pmf:121-141) -- is computed in float and added to the float64 state, which is what
// "float per-step arithmetic, float64 state" means (the fused kernels do the same on their compensated pairs).  The
// float64 form cost five FP64 instructions (two issue cycles each on B200) and two conversions per unit-step in a
// kernel that ncu shows short of issue slots, not of HBM bandwidth.  fatdt / rdidt here are the float products
// fat*dt and (rec/peak)*dt; a rested unit (c == peak) still gets exactly 0: (float)c == (float)peak.
template <bool CENTRAL, bool RECOVERY>
__device__ __forceinline__ float k1s_unit(const K1sArgs<float>& a, const UnitF& p, float x, double c, double d,
                                          float fatdt, float rdidt, double ptf, double& cn, double& dn, const double*) {
    bool on;
    float r = pool_rate_raw(p, a.sc, x, on);
    r = on ? r : 0.0f;                                                  // pool:171-172
    float fr = r;
    if (CENTRAL) {
        fr = r - fmaxf(pool_adaptation(p, r, (float)(d * a.adk_d)), 0.0f);   // pool:221, :121
        dn = duration_update(d, r > 0.0f, a.dt, a.max_dur);
    }
    float nf;
    const float cf = (float)c;
    const float F = fibers_force(p, a.sc, fr, cf, nf);
    float delta = -nf * fatdt;                                          // fibers:110-111
    if (RECOVERY) delta = (nf <= 0.0f) ? (p.ptf - cf) * rdidt : delta;  // pmf:121-141 (nf <= 0: the decay term is 0)
    double cnew = zero_if_negative(c + (double)delta);                  // fibers:114
    if (RECOVERY) cnew = (cnew > ptf) ? ptf : cnew;                     // pmf:104-105
    cn = cnew;
    return F;
}
template <bool CENTRAL, bool RECOVERY>
__device__ __forceinline__ double k1s_unit(const K1sArgs<double>& a, const UnitD& p, double x, double c, double d,
                                           double fatdt, double rdidt, double ptf, double& cn, double& dn,
                                           const double* exp_tab) {
    // both exponentials (adaptation exp(-dur/tau), pool:213-221; force sigmoid, fibers:247-252) go through the
    // table-based exp_neg_fast (relative error < 1e-14 against the mode's 1e-9 bar): the library exp costs
    // ~2.5x the FP64 instructions and this kernel is FP64-pipe bound, not HBM bound
    bool on;
    const double r = pool_rate(p, a.sc, x, on);
    double fr = r;
    if (CENTRAL) {
        const double q = fma(r, p.phir, -p.phir6);
        const double ad = q * (1.0 - exp_neg_fast<true>(d * a.sc.mitau, exp_tab));
        fr = r - ((ad < 0.0) ? 0.0 : ad);                               // pool:205-206, :121
        dn = duration_update(d, r > 0.0, a.dt, a.max_dur);
    }
    double nf;
    const double F = fibers_force_fast<true>(p, a.sc, fr, c, nf, exp_tab);
    cn = fatigue_update<RECOVERY>(c, nf, fatdt, rdidt, ptf);
    return F;
}

// per-step rate constant x step size in the mode's arithmetic
__device__ __forceinline__ float rate_dt(float rate, double dt) { return rate * (float)dt; }
__device__ __forceinline__ double rate_dt(double rate, double dt) { return rate * dt; }

// FP64 mode, first half of a unit-step: the adapted firing rate (and the new on-duration)
template <bool CENTRAL>
__device__ __forceinline__ double k1s_rate(const K1sArgs<double>& a, const UnitD& p, double x, double d, double& dn,
                                           const double* exp_tab) {
    bool on;
    const double r = pool_rate(p, a.sc, x, on);
    double fr = r;
    if (CENTRAL) {
        const double q = fma(r, p.phir, -p.phir6);
        const double ad = q * (1.0 - exp_neg_fast<true>(d * a.sc.mitau, exp_tab));
        fr = r - ((ad < 0.0) ? 0.0 : ad);                               // pool:205-206, :121
        dn = duration_update(d, r > 0.0, a.dt, a.max_dur);
    }
    return fr;
}

template <typename T, bool CENTRAL, bool RECOVERY, bool FORCES, bool FATIGUE, bool PERROW = false>
__global__ void __launch_bounds__((CENTRAL ? StreamTraits<T>::kMaxWCentral : StreamTraits<T>::kMaxW) + 32,
                  CENTRAL ? StreamTraits<T>::kMinBlocksCentral : StreamTraits<T>::kMinBlocks)
k1_stream(const K1sArgs<T> a) {
    constexpr int R = kStreamRows;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int W = a.W, NS = a.NS;
    // stage row pitch in elements (shift + round-up slack).  Compile-time -- the instantiation's widest tile, whatever W a
    // launch uses -- so that the 8 rows' shared-memory offsets are immediates: with a run-time pitch every row's read
    // cost three integer instructions in a kernel that is short of issue slots (ncu), not of shared memory
    // (FP32 mode only: in FP64 mode the immediates let the compiler hoist more and the 128-register consumers spill:
    //  config #5 in FP64 mode 580 -> 624 us)
    const int rowp = stream_row_pitch<T, CENTRAL>(W);
    const uint32_t bar0 = smem_u32(smem_raw);             // full[NS] | empty[NS]
    using P = typename StreamTraits<T>::P;
    constexpr int G = StreamTraits<T>::G;
    double* cpf_s = (double*)(smem_raw + 128);            // [NS][R][rowp]
    double* dur_s = cpf_s + (size_t)NS * R * rowp;        // [NS][R][rowp]   (CENTRAL only)
    P* par_s = (P*)(CENTRAL ? dur_s + (size_t)NS * R * rowp : dur_s);     // [NS][PR][G][W] staged unit parameters
    constexpr int kInAlign = 16 / (int)sizeof(T);         // input elements per 16 bytes (bulk-copy granularity)
    const int rowpi = W + 2 * kInAlign;                   // input row pitch (shift + round-up slack), a multiple of kInAlign
    constexpr int PR = PERROW ? R : 1;                    // parameter tiles per stage: one, or one per row (per-instance tables)
    T* in_s = (T*)(par_s + (size_t)NS * PR * G * W);      // [NS][R][rowpi] staged per-unit inputs (in_per_unit only)
    T* redc0 = in_s + (a.in_per_unit ? (size_t)NS * R * rowpi : 0);       // [2 item parities][R][8] capacity deficits
    T* red0 = redc0 + 2 * R * 8;                                          // [2 item parities][R][8] force sums
    // ragged rows: segment offsets and divisors are read once per work item by every thread ->
    // keep them in shared memory instead of paying a dependent global load per item
    double* segdiv_s = (double*)(red0 + 2 * R * 8);       // [S]
    int32_t* segoff_s = (int32_t*)(segdiv_s + (a.seg_in_smem ? a.S : 0));   // [S + 1]
    const int32_t* seg_off = a.seg_off;
    const double* seg_div = a.seg_div;
    if (a.seg_in_smem) {
        for (int i = threadIdx.x; i <= a.S; i += blockDim.x) segoff_s[i] = a.seg_off[i];
        if (a.seg_div)
            for (int i = threadIdx.x; i < a.S; i += blockDim.x) segdiv_s[i] = a.seg_div[i];
        seg_off = segoff_s;
        if (a.seg_div) seg_div = segdiv_s;
    }

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ double exp2_tab[32];                       // exp_neg_fast's table (FP64 mode)
    if (sizeof(T) == 8 && tid < 32) exp2_tab[tid] = kExp2Tab[tid];
    const int ncons = blockDim.x - 32;                    // consumer threads (== W)
    const int ncw = ncons >> 5;
    if (tid == 0) {
        for (int i = 0; i < NS; ++i) {
            mbar_init(bar0 + 8u * i, 1);                  // full: the producer's expect_tx arrival
            mbar_init(bar0 + 8u * (NS + i), ncw);         // empty: one arrival per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int64_t chunks = (a.rows + R - 1) / R;
    const int64_t items = chunks * a.S;
    // work item -> (chunk, segment) without a division per item: a block's items are gridDim.x apart
    const int step_chunks = (int)(gridDim.x / (unsigned)a.S), step_segs = (int)(gridDim.x % (unsigned)a.S);

    if (warp == ncw) {
        // ------------------------------ producer warp ----------------------------------------
        uint32_t slot = 0, phase = 0;                     // ring position (it % NS, (it / NS) & 1 without the divisions)
        int64_t chunk = (int64_t)(blockIdx.x / (unsigned)a.S);
        int seg = (int)(blockIdx.x % (unsigned)a.S);
        for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
            const int64_t r0 = chunk * R;
            const int off0 = seg_off ? seg_off[seg] : 0;
            const int n = seg_off ? seg_off[seg + 1] - off0 : (int)a.L;
            const int nrows = (int)min((int64_t)R, a.rows - r0);
            for (int u0 = 0; u0 < n; u0 += W) {
                const uint32_t s = slot, ph = phase;
                if (++slot == (uint32_t)NS) { slot = 0; phase ^= 1u; }
                mbar_wait(bar0 + 8u * (NS + s), ph ^ 1);  // stage free (passes at once in the first round)
                const int count = min(W, n - u0);
                const int64_t e0 = off0 + u0;             // element offset inside a row
                const uint32_t pbytes = (uint32_t)count * 16u;           // one parameter group of the tile
                // every lane works out what it will copy; the stage's transaction count is the warp's sum
                const bool state_lane = lane < nrows;
                const bool input_lane = a.in_per_unit && lane >= 16 && lane < 16 + nrows;   // lanes 16 .. 16+R-1: the input rows
                const int64_t ge = (r0 + (input_lane ? lane - 16 : lane)) * a.L + e0;       // this lane's row: first element of the tile
                const int64_t a0 = ge & ~(int64_t)1, a1 = (ge + count + 1) & ~(int64_t)1;   // 16-byte aligned span (8-byte state)
                const uint32_t bytes = (uint32_t)(a1 - a0) * 8u;
                const int64_t i0a = ge & ~(int64_t)(kInAlign - 1);                          // ... and of the per-unit inputs
                const int64_t i1a = (ge + count + kInAlign - 1) & ~(int64_t)(kInAlign - 1);
                const uint32_t ibytes = (uint32_t)(i1a - i0a) * (uint32_t)sizeof(T);
                uint32_t mine = state_lane ? bytes * (CENTRAL ? 2u : 1u) : (input_lane ? ibytes : 0u);
                if (PERROW) {
                    if (lane == 0) mine += pbytes * G * (uint32_t)nrows;
                } else if (lane >= R && lane < R + G) {
                    mine += pbytes;
                }
                const uint32_t total = __reduce_add_sync(0xffffffffu, mine);
                if (lane == 0) mbar_expect_tx(bar0 + 8u * s, total);
                __syncwarp();
                if (state_lane) {
                    const uint32_t so = (uint32_t)((s * R + lane) * rowp) * 8u;
                    bulk_g2s(smem_u32(cpf_s) + so, a.cpf + a0, bytes, bar0 + 8u * s);
                    if (CENTRAL) bulk_g2s(smem_u32(dur_s) + so, a.dur + a0, bytes, bar0 + 8u * s);
                }
                if (input_lane)
                    bulk_g2s(smem_u32(in_s + (size_t)(s * R + lane - 16) * rowpi), a.in + i0a, ibytes, bar0 + 8u * s);
                if (!PERROW) {
                    if (lane >= R && lane < R + G) {
                        const int grp = lane - R;                        // lanes R .. R+G-1: the parameter groups
                        bulk_g2s(smem_u32(par_s + (size_t)(s * G + grp) * W), a.params + (int64_t)grp * a.L + e0, pbytes,
                                 bar0 + 8u * s);
                    }
                } else {
                    // per-instance tables: one tile per (row, group), spread over the 32 lanes
                    for (int j = lane; j < nrows * G; j += 32) {
                        const int r = j / G, grp = j - r * G;
                        bulk_g2s(smem_u32(par_s + (size_t)((s * R + r) * G + grp) * W),
                                 a.params + (r0 + r) * a.param_stride + (int64_t)grp * a.L + e0, pbytes, bar0 + 8u * s);
                    }
                }
            }
            chunk += step_chunks;
            seg += step_segs;
            if (seg >= a.S) { seg -= a.S; ++chunk; }
        }
        return;
    }

    // ---------------------------------- consumer warps ----------------------------------------
    // the per-muscle inputs of the NEXT work item are fetched while the current one is processed
    auto fetch_inputs = [&](int64_t chunk, int seg, T (&x)[R]) {
        const int64_t r0 = chunk * R;
#pragma unroll
        for (int r = 0; r < R; ++r)
            x[r] = (!a.in_per_unit && r0 + r < a.rows) ? a.in[(r0 + r) * a.S + seg] : T(0);
    };
    uint32_t slot = 0, phase = 0, item_no = 0;            // ring position as counters: NS is a run-time value
    T xnext[R];
    int64_t chunk_next = (int64_t)(blockIdx.x / (unsigned)a.S);
    int seg_next = (int)(blockIdx.x % (unsigned)a.S);
    fetch_inputs(chunk_next, seg_next, xnext);
    for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
        const int64_t chunk = chunk_next;
        const int seg = seg_next;
        chunk_next += step_chunks;
        seg_next += step_segs;
        if (seg_next >= a.S) { seg_next -= a.S; ++chunk_next; }
        const int64_t r0 = chunk * R;
        const int off0 = seg_off ? seg_off[seg] : 0;
        const int n = seg_off ? seg_off[seg + 1] - off0 : (int)a.L;
        const int nrows = (int)min((int64_t)R, a.rows - r0);
        T xm[R], partial[R];
        T csum[R];                                            // lost capacity sum(peak - new capacity): 0 exactly for a rested muscle
#pragma unroll
        for (int r = 0; r < R; ++r) {
            partial[r] = 0;
            csum[r] = 0;
            xm[r] = xnext[r];
        }
        constexpr bool want_fatigue = FATIGUE;                // fused get_peripheral_fatigue epilogue
        fetch_inputs(chunk_next, seg_next, xnext);                // (rows past the end read as 0)
        for (int u0 = 0; u0 < n; u0 += W) {
            const uint32_t s = slot, ph = phase;
            if (++slot == (uint32_t)NS) { slot = 0; phase ^= 1u; }
            const int u = u0 + tid;
            const bool active = u < n;
            // first wanted element of a staged row: r0 * L is even (r0 is a multiple of 8), so row r starts
            // at parity (r * L + off0 + u0) & 1 -- the same in all rows for even L, alternating for odd L
            const int sh0 = (off0 + u0) & 1, sh1 = sh0 ^ (int)(a.L & 1);
            mbar_wait(bar0 + 8u * s, ph);
            if (active) {
                const P* pst = par_s + (size_t)(s * PR * G) * W;
                auto p = load_unit_staged(pst, W, tid);
                auto fatdt = rate_dt(p.fat, a.dt), rdidt = rate_dt(p.rdi, a.dt);   // float in FP32 mode, double in FP64 mode
                double ptf = unit_peak(p);
                const double* cs = cpf_s + (s * R) * rowp + tid;
                const double* ds = dur_s + (s * R) * rowp + tid;
                // state of row r, read from the stage where it is used (the stage stays valid until this warp's
                // arrival on its `empty` barrier below): loading all 2R values up front pinned 16-32 registers
                // for the whole tile and pushed the variants with an epilogue into spills
                auto ldc = [&](int r) { return cs[r * rowp + ((r & 1) ? sh1 : sh0)]; };
                auto ldd = [&](int r) { return CENTRAL ? ds[r * rowp + ((r & 1) ? sh1 : sh0)] : 0.0; };
                if (a.in_per_unit) {                                 // this unit's own excitation in every row
                    const T* xs = in_s + (size_t)(s * R) * rowpi + tid;
                    const int em = (off0 + u0) & (kInAlign - 1), lm = (int)(a.L & (kInAlign - 1));
#pragma unroll
                    for (int r = 0; r < R; ++r) xm[r] = xs[r * rowpi + ((r * lm + em) & (kInAlign - 1))];
                }
                const int64_t i0 = r0 * a.L + off0 + u;
                MB_ASSERT(u < n && off0 + u < a.L && i0 + (int64_t)(nrows - 1) * a.L < a.rows * a.L && tid < W);
                T* fp = FORCES ? a.forces + i0 : nullptr;
                double* cp = a.cpf + i0;
                double* dp = CENTRAL ? a.dur + i0 : nullptr;
                auto row = [&](int r) {
                    if (PERROW && r > 0) {                           // per-instance tables: this row's own parameters
                        p = load_unit_staged(pst + (size_t)(r * G) * W, W, tid);
                        fatdt = rate_dt(p.fat, a.dt);
                        rdidt = rate_dt(p.rdi, a.dt);
                        ptf = unit_peak(p);
                    }
                    const double c = ldc(r), d = ldd(r);
                    double cn, dn = d;
                    const T F = k1s_unit<CENTRAL, RECOVERY>(a, p, xm[r], c, d, fatdt, rdidt, ptf, cn, dn, exp2_tab);
                    if (FORCES) *fp = F;
                    *cp = cn;
                    if (CENTRAL) *dp = dn;
                    partial[r] += F;
                    if (want_fatigue) csum[r] += (T)(ptf - cn);
                    if (FORCES) fp += a.L;
                    cp += a.L;
                    if (CENTRAL) dp += a.L;
                };
                if constexpr (sizeof(T) == 8 && !PERROW && (CENTRAL || !FATIGUE)) {
                    // FP64 mode: GR rows at a time in three phases, so that their FP64 chains interleave.  A
                    // branch per row around the sigmoid's exponential (fibers_force_fast) serialises the rows --
                    // the scheduler cannot move instructions across it, and ncu showed 40 % of the stall samples
                    // on fixed-latency dependencies.  One branch decides for the group; rows past the end of a
                    // partial chunk are computed on stale shared memory and not stored.  The group size is what
                    // the register budget of the instantiation allows (measured: larger groups spill and lose; the
                    // variant with the fatigue epilogue has no room at all and keeps the row-by-row form).
                    constexpr int GR = CENTRAL ? 8 : 2;
#pragma unroll
                    for (int g = 0; g < R; g += GR) {
                        double xr[GR], dn[GR], ex[GR], c[GR];
                        bool need = false;
#pragma unroll
                        for (int i = 0; i < GR; ++i) {
                            const double d = ldd(g + i);
                            c[i] = ldc(g + i);
                            dn[i] = d;
                            const double fr = k1s_rate<CENTRAL>(a, p, xm[g + i], d, dn[i], exp2_tab);
                            xr[i] = fma(c[i], -p.ctB, p.ctA) * fr;                 // fibers:220
                            need |= xr[i] > kLinThrD;
                        }
                        if (need) {
#pragma unroll
                            for (int i = 0; i < GR; ++i) ex[i] = exp_neg_fast<true>(-2.0 * (xr[i] * xr[i] * xr[i]), exp2_tab);
                        } else {
#pragma unroll
                            for (int i = 0; i < GR; ++i) ex[i] = 1.0;
                        }
#pragma unroll
                        for (int i = 0; i < GR; ++i) {
                            const int r = g + i;
                            const double nf = (xr[i] <= kLinThrD) ? xr[i] * a.sc.c25 : 1.0 - ex[i];   // fibers:247-252
                            const double F = nf * c[i];                                        // fibers:258
                            const double cn = fatigue_update<RECOVERY>(c[i], nf, fatdt, rdidt, ptf);
                            if (nrows == R || r < nrows) {
                                if (FORCES) *fp = F;
                                *cp = cn;
                                if (CENTRAL) *dp = dn[i];
                                partial[r] += F;
                                if (want_fatigue) csum[r] += (T)(ptf - cn);
                            }
                            if (FORCES) fp += a.L;
                            cp += a.L;
                            if (CENTRAL) dp += a.L;
                        }
                    }
                } else if (nrows == R) {
#pragma unroll
                    for (int r = 0; r < R; ++r) row(r);
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (r < nrows) row(r);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + 8u * (NS + s));    // this warp is done with the stage
        }
        if (a.totals || want_fatigue) {
            // scratch alternates between consecutive items, so ONE barrier per item is enough: a warp
            // reaches the next write of the same half only after the barrier of the item in between
            T* red = red0 + (item_no & 1) * (R * 8);
            T* redc = redc0 + (item_no & 1) * (R * 8);
            ++item_no;
            static_assert(R == 8, "warp_sum_rows8 reduces exactly eight rows");
            const int rr = (lane >> 2) & 7;                       // the row whose warp total this lane ends up with
            const T wf = warp_sum_rows8(partial, lane);
            if ((lane & 3) == 0) red[rr * 8 + warp] = wf;
            if (want_fatigue) {
                const T wc = warp_sum_rows8(csum, lane);
                if ((lane & 3) == 0) redc[rr * 8 + warp] = wc;
            }
            asm volatile("bar.sync 1, %0;" ::"r"(ncons) : "memory");
            if (tid < nrows) {
                const int64_t mu = (r0 + tid) * a.S + seg;
                MB_ASSERT(mu < a.rows * a.S && seg < a.S);
                if (a.totals) {
                    T sum = red[tid * 8];
                    for (int w = 1; w < ncw; ++w) sum += red[tid * 8 + w];
                    sum = seg_div ? out_div(sum, seg_div[seg]) : out_scale(a.sc, sum);
                    if (a.out_gain) sum *= a.out_gain[mu];                       // epilogue: caller-supplied multiplier
                    if (a.hill_rest)                                             // ... and the Hill-type factors of this muscle
                        sum *= hill_gain(a.hill_rest[mu], a.hill_cur[mu], a.hill_prev[mu], a.dt, a.hill_width, a.hill_opt,
                                         a.hill_vmax, a.hill_ecc);
                    a.totals[mu] = sum;
                }
                if (want_fatigue) {
                    // 1 - sum(capacity)/max_arb_output (muscle.py:248-256) as sum(peak - capacity)/max_arb_output: the same
                    // number for a StandardMuscle (max_arb_output = sum(peak), muscle.py:187) up to rounding, but exactly 0
                    // for a rested muscle whatever the order of the parallel sum, and never negative
                    T cs = redc[tid * 8];
                    for (int w = 1; w < ncw; ++w) cs += redc[tid * 8 + w];
                    a.fatigue[mu] = (double)cs / (seg_div ? seg_div[seg] : a.fat_div);
                }
            }
        }
    }
}

}  // namespace mb
