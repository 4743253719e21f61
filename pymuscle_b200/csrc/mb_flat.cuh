// mb_flat.cuh -- K1f: the single-step streaming kernel for rows made of SEVERAL muscles (ragged rows:
// the 4 muscles of a hopper leg, the 30 of an arm; examples/envs/muscled_hopper.py:10-23, :33-34).
//
// k1_stream (mb_stream.cuh) makes every (muscle, 8 rows) its own work item: column tiles never cross a
// muscle boundary, so short muscles leave tiles mostly empty (a hopper's 74 + 120 + 169 + 120 units fill
// five 256-wide tiles to 38 %), every item pays a reduction, a named barrier and an epilogue, and items
// differ in size by the muscles' lengths (8.8 % load imbalance on the arm).  Here
//
//   * a work item is (GROUP of whole adjacent muscles, 8 rows); groups are cut greedily to ~kFlatGroupUnits
//     units and at most kFlatMaxSegs muscles (every block derives the same grouping from seg_off at start),
//     so items are nearly equal and a block's static share of them cycles through all groups;
//   * the group's columns are tiled FLAT: W-wide tiles run across muscle boundaries (hopper: 2 tiles, 94 %
//     full; arm: 137 tiles per row chunk instead of 152);
//   * per-muscle sums are SEGMENTED and deterministic: a warp whose 32 columns lie in one muscle keeps
//     accumulating in registers across tiles and flushes (transposing butterfly, 9 shuffles for 8 rows)
//     only when its muscle changes; a warp that straddles a boundary reduces once per muscle it touches
//     with the other lanes masked.  Flushes go to the warp's OWN slots acc[warp][muscle][row] (plain
//     read-modify-write by one lane, program order), and after ONE named barrier per item thread (row,
//     muscle) adds the warps' slots in a fixed order and runs the per-muscle epilogue (divisor, gain,
//     Hill-type factors, fatigue readout);
//   * the per-muscle inputs of an item ride in every stage next to the state rows (the producer warp keeps
//     them in registers, fetched one item ahead), so consumers never touch global memory for them.
//
// Everything else is k1_stream: producer warp + cp.async.bulk (SASS UBLKCP) into an NS-stage ring with
// full/empty mbarriers, 16-byte-aligned row copies with a 0/1 element shift, consumers thread = column.
#pragma once
#include "mb_stream.cuh"

namespace mb {

constexpr int kFlatMaxSegs = 16;        // muscles per group (bounds the per-warp accumulator slots)
constexpr int kFlatMaxGroups = 256;     // groups per row
constexpr int kFlatMaxS = 1024;         // muscles per row the kernel takes (group table is built by one thread)

__device__ __forceinline__ float fma_t(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fma_t(double a, double b, double c) { return fma(a, b, c); }

template <typename T>
struct K1fArgs {
    K1sArgs<T> s;                       // in_per_unit = 0, param_stride = 0, S > 1
    int32_t group_units;                // target units per group
};

template <typename T, bool CENTRAL, bool RECOVERY, bool FORCES, bool FATIGUE>
__global__ void __launch_bounds__((CENTRAL ? StreamTraits<T>::kMaxWCentral : StreamTraits<T>::kMaxW) + 32,
                  CENTRAL ? StreamTraits<T>::kMinBlocksCentral : StreamTraits<T>::kMinBlocks)
k1_flat(const K1fArgs<T> fa) {
    const K1sArgs<T>& a = fa.s;
    constexpr int R = kStreamRows;
    constexpr int NSG = kFlatMaxSegs;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int W = a.W, NS = a.NS, S = a.S;
    const int rowp = stream_row_pitch<T, CENTRAL>(W);     // compile-time in FP32 mode (mb_stream.cuh)
    const uint32_t bar0 = smem_u32(smem_raw);             // full[NS] | empty[NS]
    using P = typename StreamTraits<T>::P;
    constexpr int G = StreamTraits<T>::G;
    double* cpf_s = (double*)(smem_raw + 128);            // [NS][R][rowp]
    double* dur_s = cpf_s + (size_t)NS * R * rowp;        // [NS][R][rowp]   (CENTRAL only)
    P* par_s = (P*)(CENTRAL ? dur_s + (size_t)NS * R * rowp : dur_s);     // [NS][G][W]
    T* xin_s = (T*)(par_s + (size_t)NS * G * W);          // [NS][R][NSG] the item's per-muscle inputs
    const int ncons = blockDim.x - 32, ncw = ncons >> 5;
    T* accc = xin_s + (size_t)NS * R * NSG;               // [2][ncw][NSG][R] capacity deficits (FATIGUE)
    T* accf = accc + (FATIGUE ? 2 * ncw * NSG * R : 0);   // [2][ncw][NSG][R] force sums
    double* segdiv_s = (double*)(accf + 2 * ncw * NSG * R + ((2 * ncw * NSG * R) & 1));   // [S] (8-byte aligned)
    int32_t* segoff_s = (int32_t*)(segdiv_s + S);         // [S + 1]
    int32_t* grp_s = segoff_s + S + 1;                    // [ng + 1] first muscle of every group
    __shared__ int ng_s;
    __shared__ double exp2_tab[32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (sizeof(T) == 8 && tid < 32) exp2_tab[tid] = kExp2Tab[tid];
    for (int i = tid; i <= S; i += blockDim.x) segoff_s[i] = a.seg_off[i];
    if (a.seg_div)
        for (int i = tid; i < S; i += blockDim.x) segdiv_s[i] = a.seg_div[i];
    if (tid == 0) {
        for (int i = 0; i < NS; ++i) {
            mbar_init(bar0 + 8u * i, 1);                  // full: the producer's expect_tx arrival
            mbar_init(bar0 + 8u * (NS + i), ncw);         // empty: one arrival per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {                                       // greedy grouping, identical in every block
        int ng = 0, acc = 0, cnt = 0;
        grp_s[0] = 0;
        for (int sgi = 0; sgi < S; ++sgi) {
            const int len = segoff_s[sgi + 1] - segoff_s[sgi];
            if (cnt > 0 && (acc + len > fa.group_units || cnt == NSG) && ng + 1 < kFlatMaxGroups) {
                grp_s[++ng] = sgi;
                acc = 0;
                cnt = 0;
            }
            acc += len;
            ++cnt;
        }
        grp_s[++ng] = S;
        ng_s = ng;
    }
    __syncthreads();
    const int ng = ng_s;
    const double* seg_div = a.seg_div ? segdiv_s : nullptr;

    const int64_t chunks = (a.rows + R - 1) / R;
    const int64_t items = chunks * ng;                    // item = chunk * ng + group (chunk-major: neighbours in DRAM)
    const int step_chunks = (int)(gridDim.x / (unsigned)ng), step_grps = (int)(gridDim.x % (unsigned)ng);

    if (warp == ncw) {
        // ------------------------------ producer warp ----------------------------------------
        uint32_t slot = 0, phase = 0;                     // ring position (it % NS, (it / NS) & 1 without the divisions)
        int64_t chunk = (int64_t)(blockIdx.x / (unsigned)ng);
        int grp = (int)(blockIdx.x % (unsigned)ng);
        // the item's inputs, 4 per lane: value j of lane l is (row, muscle) = ((l + 32 j) / NSG, (l + 32 j) % NSG)
        auto fetch_inputs = [&](int64_t ch, int g, T (&x)[4]) {
            const int sg0 = grp_s[g], nsg = grp_s[g + 1] - sg0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int idx = lane + 32 * j, r = idx / NSG, sl = idx % NSG;
                const int64_t row = ch * R + r;
                x[j] = (ch < chunks && row < a.rows && sl < nsg) ? a.in[row * S + sg0 + sl] : T(0);
            }
        };
        T xcur[4], xnext[4];
        fetch_inputs(chunk, grp, xnext);
        for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
#pragma unroll
            for (int j = 0; j < 4; ++j) xcur[j] = xnext[j];
            int64_t chunk_n = chunk + step_chunks;
            int grp_n = grp + step_grps;
            if (grp_n >= ng) { grp_n -= ng; ++chunk_n; }
            fetch_inputs(chunk_n, grp_n, xnext);           // one item ahead: the latency hides behind this item's tiles
            const int64_t r0 = chunk * R;
            const int c0 = segoff_s[grp_s[grp]], c1 = segoff_s[grp_s[grp + 1]];
            const int nrows = (int)min((int64_t)R, a.rows - r0);
            for (int u0 = c0; u0 < c1; u0 += W) {
                const uint32_t s = slot, ph = phase;
                if (++slot == (uint32_t)NS) { slot = 0; phase ^= 1u; }
                mbar_wait(bar0 + 8u * (NS + s), ph ^ 1);  // stage free (passes at once in the first round)
                const int count = min(W, c1 - u0);
                const uint32_t pbytes = (uint32_t)count * 16u;
                const bool state_lane = lane < nrows;
                const int64_t ge = (r0 + lane) * a.L + u0;
                const int64_t a0 = ge & ~(int64_t)1, a1 = (ge + count + 1) & ~(int64_t)1;
                const uint32_t bytes = (uint32_t)(a1 - a0) * 8u;
                uint32_t mine = state_lane ? bytes * (CENTRAL ? 2u : 1u) : 0u;
                if (lane >= R && lane < R + G) mine += pbytes;
                const uint32_t total = __reduce_add_sync(0xffffffffu, mine);
                T* xs = xin_s + (size_t)s * R * NSG;
#pragma unroll
                for (int j = 0; j < 4; ++j) xs[lane + 32 * j] = xcur[j];
                __syncwarp();                              // the input stores are ordered before lane 0's (releasing) arrival
                if (lane == 0) mbar_expect_tx(bar0 + 8u * s, total);
                __syncwarp();
                if (state_lane) {
                    const uint32_t so = (uint32_t)((s * R + lane) * rowp) * 8u;
                    bulk_g2s(smem_u32(cpf_s) + so, a.cpf + a0, bytes, bar0 + 8u * s);
                    if (CENTRAL) bulk_g2s(smem_u32(dur_s) + so, a.dur + a0, bytes, bar0 + 8u * s);
                }
                if (lane >= R && lane < R + G) {
                    const int gq = lane - R;
                    bulk_g2s(smem_u32(par_s + (size_t)(s * G + gq) * W), a.params + (int64_t)gq * a.L + u0, pbytes, bar0 + 8u * s);
                }
            }
            chunk = chunk_n;
            grp = grp_n;
        }
        return;
    }

    // ---------------------------------- consumer warps ----------------------------------------
    uint32_t slot = 0, phase = 0, item_no = 0;            // ring position as counters: NS is a run-time value
    int64_t chunk = (int64_t)(blockIdx.x / (unsigned)ng);
    int grp = (int)(blockIdx.x % (unsigned)ng);
    const int rr = (lane >> 2) & 7;                       // the row whose warp total this lane holds after a butterfly
    for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
        const int64_t r0 = chunk * R;
        const int sg0 = grp_s[grp], nsg = grp_s[grp + 1] - sg0;
        const int c0 = segoff_s[sg0], c1 = segoff_s[sg0 + nsg];
        const int nrows = (int)min((int64_t)R, a.rows - r0);
        const int buf = item_no & 1;
        T* af = accf + (size_t)(buf * ncw + warp) * NSG * R;            // this warp's slots [NSG][R]
        T* ac = accc + (size_t)(buf * ncw + warp) * NSG * R;
        for (int i = lane; i < NSG * R; i += 32) {
            af[i] = T(0);
            if (FATIGUE) ac[i] = T(0);
        }
        __syncwarp();
        T partial[R], csum[R];                            // forces; lost capacity sum(peak - new capacity)
#pragma unroll
        for (int r = 0; r < R; ++r) { partial[r] = 0; csum[r] = 0; }
        int wseg = -1;                                    // the muscle this warp's registers are accumulating (-1: none)
        // Clearing the accumulators is folded into the first accumulation after a flush: acc = acc * keep + F with keep = 0
        // (an FMA in place of the add).  As a conditional assignment the clear made the compiler merge 16 registers on
        // both sides of the branch -- 32 moves per tile in a kernel that is short of issue slots
        // (FP32 mode only: in FP64 mode the extra live value costs the 128-register consumers spills)
        constexpr bool kFoldClear = sizeof(T) == 4;
        T keep = T(1);
        int seg = sg0;                                    // muscle of this thread's column (advances monotonically)
        // registers -> this warp's slots of local muscle sl; `mine`: the lanes whose column belongs to that muscle
        auto flush_masked = [&](int sl, bool mine) {
            T t[R];
#pragma unroll
            for (int r = 0; r < R; ++r) t[r] = mine ? partial[r] : T(0);
            const T wf = warp_sum_rows8(t, lane);
            MB_ASSERT(sl >= 0 && sl < NSG);
            if ((lane & 3) == 0) af[sl * R + rr] += wf;
            if (FATIGUE) {
                T tc[R];
#pragma unroll
                for (int r = 0; r < R; ++r) tc[r] = mine ? csum[r] : T(0);
                const T wc = warp_sum_rows8(tc, lane);
                if ((lane & 3) == 0) ac[sl * R + rr] += wc;
            }
        };
        auto clear = [&]() {
#pragma unroll
            for (int r = 0; r < R; ++r) { partial[r] = 0; csum[r] = 0; }
        };
        for (int u0 = c0; u0 < c1; u0 += W) {
            const uint32_t s = slot, ph = phase;
            if (++slot == (uint32_t)NS) { slot = 0; phase ^= 1u; }
            const int c = u0 + tid;
            const bool active = c < c1;
            const int cc = active ? c : c1 - 1;           // idle lanes shadow the group's last column (their forces are 0)
            while (cc >= segoff_s[seg + 1]) ++seg;
            const int sh0 = u0 & 1, sh1 = sh0 ^ (int)(a.L & 1);
            const int seg_lo = __shfl_sync(0xffffffffu, seg, 0), seg_hi = __shfl_sync(0xffffffffu, seg, 31);
            if (wseg >= 0 && (seg_lo != wseg || seg_hi != wseg)) {      // the warp leaves its muscle: flush first
                flush_masked(wseg - sg0, true);
                if constexpr (kFoldClear) keep = T(0); else clear();
                wseg = -1;
            }
            mbar_wait(bar0 + 8u * s, ph);
            if (active) {
                const P* pst = par_s + (size_t)(s * G) * W;
                const auto p = load_unit_staged(pst, W, tid);
                const auto fatdt = rate_dt(p.fat, a.dt), rdidt = rate_dt(p.rdi, a.dt);   // float in FP32 mode, double in FP64 mode
                const double ptf = unit_peak(p);
                const double* cs = cpf_s + (s * R) * rowp + tid;
                const double* ds = dur_s + (s * R) * rowp + tid;
                const T* xs = xin_s + (size_t)s * R * NSG + (seg - sg0);
                const int64_t i0 = r0 * a.L + c;
                MB_ASSERT(c < a.L && seg >= sg0 && seg - sg0 < nsg && nsg <= NSG && tid < W &&
                          i0 + (int64_t)(nrows - 1) * a.L < a.rows * a.L);
                T* fp = FORCES ? a.forces + i0 : nullptr;
                double* cp = a.cpf + i0;
                double* dp = CENTRAL ? a.dur + i0 : nullptr;
                auto row = [&](int r) {
                    const int sh = (r & 1) ? sh1 : sh0;
                    const double cv = cs[r * rowp + sh], dv = CENTRAL ? ds[r * rowp + sh] : 0.0;
                    double cn, dn = dv;
                    const T F = k1s_unit<CENTRAL, RECOVERY>(a, p, xs[r * NSG], cv, dv, fatdt, rdidt, ptf, cn, dn, exp2_tab);
                    if (FORCES) fp[(int64_t)r * a.L] = F;
                    cp[(int64_t)r * a.L] = cn;
                    if (CENTRAL) dp[(int64_t)r * a.L] = dn;
                    if constexpr (kFoldClear) {
                        partial[r] = fma_t(partial[r], keep, F);
                        if (FATIGUE) csum[r] = fma_t(csum[r], keep, (T)(ptf - cn));
                    } else {
                        partial[r] += F;
                        if (FATIGUE) csum[r] += (T)(ptf - cn);
                    }
                };
                if (nrows == R) {                                 // full chunk: straight-line code for the 8 rows
#pragma unroll
                    for (int r = 0; r < R; ++r) row(r);
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        if (r < nrows) {
                            row(r);
                        } else if (kFoldClear) {
                            partial[r] *= keep;
                            csum[r] *= keep;
                        }
                    }
                }
            } else if (kFoldClear) {                              // idle lanes of a group's last tile
#pragma unroll
                for (int r = 0; r < R; ++r) { partial[r] *= keep; csum[r] *= keep; }
            }
            keep = T(1);
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + 8u * (NS + s));    // this warp is done with the stage
            if (seg_lo == seg_hi) {                               // the whole warp inside one muscle: keep summing in registers
                wseg = seg_lo;
            } else {                                              // the warp straddles muscle boundaries: one masked reduction each
                for (int sq = seg_lo; sq <= seg_hi; ++sq) flush_masked(sq - sg0, active && seg == sq);
                if constexpr (kFoldClear) keep = T(0); else clear();
            }
        }
        if (wseg >= 0) flush_masked(wseg - sg0, true);
        // ---- per-muscle totals of the item: thread (row, local muscle) adds the warps' slots in a fixed order --------
        asm volatile("bar.sync 1, %0;" ::"r"(ncons) : "memory");
        for (int q = tid; q < nrows * nsg; q += ncons) {
            const int r = q / nsg, sl = q - r * nsg;
            const int sg = sg0 + sl;
            const int64_t mu = (r0 + r) * S + sg;
            MB_ASSERT(sg < S && mu < a.rows * S && sl < NSG);
            const T* f0 = accf + (size_t)(buf * ncw) * NSG * R + sl * R + r;
            if (a.totals) {
                T sum = f0[0];
                for (int w = 1; w < ncw; ++w) sum += f0[(size_t)w * NSG * R];
                sum = seg_div ? out_div(sum, seg_div[sg]) : out_scale(a.sc, sum);
                if (a.out_gain) sum *= a.out_gain[mu];
                if (a.hill_rest)
                    sum *= hill_gain(a.hill_rest[mu], a.hill_cur[mu], a.hill_prev[mu], a.dt, a.hill_width, a.hill_opt,
                                     a.hill_vmax, a.hill_ecc);
                a.totals[mu] = sum;
            }
            if (FATIGUE) {
                const T* q0 = accc + (size_t)(buf * ncw) * NSG * R + sl * R + r;
                T cs = q0[0];
                for (int w = 1; w < ncw; ++w) cs += q0[(size_t)w * NSG * R];
                a.fatigue[mu] = (double)cs / (seg_div ? seg_div[sg] : a.fat_div);     // lost capacity / max_arb_output (see mb_stream.cuh)
            }
        }
        ++item_no;
        chunk += step_chunks;
        grp += step_grps;
        if (grp >= ng) { grp -= ng; ++chunk; }
    }
}

}  // namespace mb
