/*
 * muscle_b200.h -- C ABI of libmuscle_b200.so: the PyMuscle per-step hot path
 * (motor-neuron pool -> muscle fibers -> per-muscle total) as hand-written
 * sm_100a CUDA kernels.
 *
 * The reference (iandanforth/pymuscle v0.1.2) is pure Python/NumPy and has no
 * FFI of its own; the boundary it offers is the Python class surface
 *     Muscle.step                      pymuscle/muscle.py:65-92
 *     StandardMuscle.step              pymuscle/muscle.py:258-279
 *     PotvinFuglevand2017MotorNeuronPool.step   pymuscle/potvin_fuglevand_2017_motor_neuron_pool.py:283-295
 *     PotvinFuglevand2017MuscleFibers.step      pymuscle/potvin_fuglevand_2017_muscle_fibers.py:284-300
 *     PyMuscleFibers._update_fatigue            pymuscle/pymuscle_fibers.py:80-141
 * Each entry point below names the reference method(s) it replaces.  The Python
 * package pymuscle_b200 binds these with ctypes (see INTEGRATION.md for the
 * stub a maintainer of the reference would add).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross the boundary;
 *   - every `*_dev` pointer is a DEVICE pointer owned by the caller (a torch
 *     tensor kept alive by the Python side); nothing is allocated, freed or
 *     synchronised here; work is enqueued on `stream` (a cudaStream_t passed as
 *     void*; NULL = legacy default stream);
 *   - return value 0 = success, negative MB_E* otherwise; mb_last_error() gives
 *     a thread-local human-readable message for the last failure;
 *   - the library is stateless: all simulator state lives in the caller's
 *     `cpf` / `dur` arrays (float64 in both precision modes).
 *
 * Batch layout ("rows x segments")
 *   A batch is `rows` independent rows; each row holds `S` muscles (segments)
 *   laid out back to back, `L` = sum of their unit counts.  Per-unit arrays are
 *   [rows, L] row-major, per-muscle arrays [rows, S].  seg_off[S+1] (int32,
 *   device) gives the unit offset of each segment inside a row; NULL means one
 *   muscle of L units per row (the uniform case, e.g. 65,536 x 120).  The
 *   per-unit parameter table has L entries and is shared by all rows.
 */
#ifndef MUSCLE_B200_H
#define MUSCLE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB_ABI_VERSION 2

enum MbStatus { MB_OK = 0, MB_EINVAL = -1, MB_EUNSUPPORTED = -2, MB_ECUDA = -3 };

/* Arithmetic mode.  MB_F32: float per-step arithmetic, compensated / float64
 * state (rtol 1e-4 over 10,000 steps); MB_F64: float64 throughout (rtol 1e-9).
 * Excitations, forces and totals have the mode's element type; cpf/dur are
 * float64 in both.  The recruitment mask is bit-exact in both modes. */
enum MbPrecision { MB_F32 = 0, MB_F64 = 1 };

/* Which fibers model updates the capacities: Potvin & Fuglevand 2017
 * (fibers:96-115) or PyMuscleFibers with recovery + upper clamp (pmf:80-141). */
enum MbFibers { MB_FIBERS_POTVIN = 0, MB_FIBERS_PYMUSCLE = 1 };

/* Which half of Muscle.step a single-step launch runs. */
enum MbStage {
    MB_STAGE_MUSCLE = 0,   /* excitation -> forces, totals  (Muscle.step, muscle.py:65-92)            */
    MB_STAGE_POOL   = 1,   /* excitation -> adapted firing rates       (Pool.step, pool:283-295)       */
    MB_STAGE_FIBERS = 2    /* firing rates -> forces, totals           (Fibers.step, fibers:284-300)   */
};

/* Scalar hyper-parameters: the constructor arguments of the reference models
 * that are not per-unit tables (pool:53-66, fibers:46-56, muscle.py:152-187). */
typedef struct MbConfig {
    int32_t precision;            /* MbPrecision */
    int32_t fibers_kind;          /* MbFibers */
    int32_t central_fatigue;      /* Pool(apply_fatigue=...)    pool:65,193 */
    int32_t peripheral_fatigue;   /* Fibers(apply_fatigue=...)  fibers:55,279 */
    double max_recruitment_threshold;   /* RR   pool:56  */
    double firing_gain;                 /* g    pool:57  */
    double min_firing_rate;             /* minR pool:58  */
    double derecruitment_delta;         /* d    pool:61  */
    double adaptation_magnitude;        /* phi  pool:62  */
    double adaptation_time_constant;    /* tau  pool:63  */
    double max_duration;                /*      pool:64  */
    double contraction_time_change_ratio; /*    fibers:54 */
    double input_scale;    /* 1, or Pool.max_excitation for StandardMuscle (muscle.py:275) */
    double output_divisor; /* 1, or StandardMuscle.max_arb_output          (muscle.py:278) */
    int32_t table_props;   /* MbTableProps bits returned by mb_params_pack for the table in use */
    int32_t param_rows;    /* 0 or 1: one parameter table shared by all rows (the reference's case: identical muscles);
                              N > 1: per-instance hyper-parameters -- params_dev holds N consecutive packed tables and
                              row r of a uniform batch (S == 1, rows == N) uses table r (domain randomisation) */
} MbConfig;

/* Facts about a packed table that let the fused kernel drop provably dead clamps. */
enum MbTableProps {
    MB_PROP_ADAPT_NONNEG = 1,  /* the adaptation of a recruited unit is never negative (pool:221 is a no-op), gain and
                                  minimal firing rate are positive (rate > 0 <=> recruited) */
    MB_PROP_EXP_BOUNDED = 2    /* 2 * (ct_cur * rate / 1000)^3 stays below 600 for every unit: exp(-2x^3) cannot underflow */
};

int mb_abi_version(void);
const char* mb_last_error(void);

/* ---- parameter tables (host) ------------------------------------------------------
 * Packs the reference's per-unit float64 tables (built by the caller with NumPy so
 * that thresholds are bit-identical to pool:263-281 etc.) into the mode-specific
 * 128-bit-vector layout the kernels load.  For MB_F32 this also derives, per unit,
 * the critical float32 input at which the reference's float64 recruitment predicate
 * (pool:169-172, after muscle.py:275's scaling) flips, so the FP32 mask is exact.
 * `rec` may be NULL for MB_FIBERS_POTVIN.  `out_host` receives mb_params_bytes()
 * bytes which the caller uploads to the device; `table_props` receives MbTableProps bits
 * that the caller copies into MbConfig.table_props for later launches. */
size_t mb_params_bytes(const MbConfig* cfg, int64_t L);
int mb_params_pack(const MbConfig* cfg, int64_t L,
                   const double* thr, const double* peak, const double* ptf,
                   const double* ct, const double* fat, const double* rec,
                   void* out_host, int32_t* table_props /* out, may be NULL */);

/* ---- K1: one step for the whole batch, one launch (HBM-bound) -----------------------
 * Replaces, for every muscle of the batch at once, Muscle.step / StandardMuscle.step
 * (stage MUSCLE), Pool.step (stage POOL) or Fibers.step (stage FIBERS).
 *   in_dev        stage MUSCLE/POOL: excitations; stage FIBERS: firing rates.
 *                 [rows, S] when in_per_unit == 0 (one value per muscle, the scalar form of
 *                 muscle.py:81-86), [rows, L] when in_per_unit == 1 (the ndarray form).
 *   cpf_dev       [rows, L] float64 current peak forces   (fibers:63, R/W unless stage POOL)
 *   dur_dev       [rows, L] float64 recruitment durations (pool:79,   R/W unless stage FIBERS)
 *   forces_dev    [rows, L] per-unit forces = Muscle.current_forces (fibers:258); may be NULL
 *   seg_div_dev   [S] float64 per-segment output divisors (each StandardMuscle of a ragged row
 *                 has its own max_arb_output, muscle.py:187); NULL = cfg->output_divisor for all
 *   totals_dev    [rows, S] per-muscle totals (fibers:276, / output divisor); may be NULL
 *   mask_dev      [rows, L] uint8 recruitment mask (pool:171); may be NULL
 *   rates_dev     [rows, L] adapted firing rates (Pool.step's return); NULL unless wanted
 * The streaming kernel (stage MUSCLE without mask / rates, device-resident inputs, 16-byte aligned
 * cpf / dur / per-unit input base pointers) reads the state and per-unit input arrays in whole
 * 16-byte granules: with an odd rows * L it touches (and ignores) the padding up to the next 16-byte
 * boundary after the last element, which every CUDA allocator provides.
 */
int mb_step(const MbConfig* cfg, const void* params_dev,
            int64_t rows, int64_t L, int32_t S, const int32_t* seg_off_dev,
            const double* seg_div_dev,
            int32_t stage, const void* in_dev, int32_t in_per_unit,
            double* cpf_dev, double* dur_dev,
            void* forces_dev, void* totals_dev, uint8_t* mask_dev, void* rates_dev,
            double step_size, void* stream);

/* ---- K1 with a fused per-muscle epilogue (vector-environment step) ----------------------
 * One launch replaces, for every muscle of every row, the loop bodies of the reference's
 * environment examples: outputs = [muscle.step(a[i], dt) for i, muscle in enumerate(muscles)]
 * (examples/envs/muscled_hopper.py:33-34, examples/envs/pymunk_arm.py:165-166) followed by
 * StandardMuscle.get_peripheral_fatigue() (muscle.py:248-256) as the proprioceptive signal, with
 * an optional per-muscle output multiplier (e.g. the Hill-type force-length / force-velocity
 * factors of pymuscle/hill_type.py:15-66).
 *   in_dev        [rows, S] one excitation per muscle
 *   out_gain_dev  [rows, S] totals are multiplied by it; may be NULL
 *   fatigue_dev   [rows, S] float64, sum(peak - new capacity)/output divisor (see mb_peripheral_fatigue); may be NULL
 * Returns MB_EUNSUPPORTED for layouts the streaming kernel does not cover (state arrays not
 * 16-byte aligned, peripheral fatigue off); the caller then uses mb_step + mb_peripheral_fatigue. */
int mb_step_env(const MbConfig* cfg, const void* params_dev,
                int64_t rows, int64_t L, int32_t S, const int32_t* seg_off_dev, const double* seg_div_dev,
                const void* in_dev, double* cpf_dev, double* dur_dev,
                void* forces_dev, void* totals_dev, const void* out_gain_dev, double* fatigue_dev,
                double step_size, void* stream);

/* ---- mb_step_env with the Hill-type curves evaluated in the epilogue -------------------------------
 * Replaces, per muscle, contractile_element_force_length_curve (pymuscle/hill_type.py:15-41) times
 * contractile_element_force_velocity_curve (hill_type.py:44-66) applied to the muscle's output: the
 * total of muscle (row, s) is multiplied by FL(rest, cur) * FV(rest, cur, prev, step_size) computed
 * inside the kernel from three [rows, S] length arrays (mode's element type).  As in the reference the
 * optimal length is pinned at 1.10 rest lengths (hill_type.py:38).  out_gain_dev (may be NULL) still
 * multiplies on top. */
typedef struct MbHill {
    const void* rest_dev;            /* [rows, S] resting lengths                       */
    const void* cur_dev;             /* [rows, S] current lengths                       */
    const void* prev_dev;            /* [rows, S] lengths one step earlier              */
    double curve_width_factor;       /* 17.33  hill_type.py:18 */
    double max_velocity;             /* 3.0    hill_type.py:49 */
    double max_eccentric_multiple;   /* 1.8    hill_type.py:50 */
} MbHill;
int mb_step_env_hill(const MbConfig* cfg, const void* params_dev,
                     int64_t rows, int64_t L, int32_t S, const int32_t* seg_off_dev, const double* seg_div_dev,
                     const void* in_dev, double* cpf_dev, double* dur_dev,
                     void* forces_dev, void* totals_dev, const void* out_gain_dev, double* fatigue_dev,
                     const MbHill* hill, double step_size, void* stream);

/* ---- K2: K fused steps, state in registers (FP32/FP64-pipe-bound) --------------------
 * Equivalent to K successive Muscle.step calls on each of `M` identical muscles of `n`
 * units (uniform layout only: rows = M, L = n, S = 1).
 *   exc_dev          excitations, element (k, m) at exc_dev[k * exc_stride_k + m];
 *                    exc_stride_k == 0 means one constant value per muscle for the window
 *   totals_dev       totals, element (k, m) at totals_dev[k * totals_stride_k + m]; may be NULL
 *   forces_every     f > 0: per-unit forces (and masks) of steps f-1, 2f-1, ... are written to
 *                    forces_dec_dev[(k / f), m, u] / mask_dec_dev; 0 = never
 *   forces_last_dev  [M, n] forces of the window's last step (current_forces); may be NULL
 * Returns MB_EUNSUPPORTED when n is outside the fused kernel's range (2..480); the
 * caller then issues K mb_step launches instead (still on the GPU).
 */
int mb_step_fused(const MbConfig* cfg, const void* params_dev,
                  int64_t M, int32_t n, int32_t K,
                  const void* exc_dev, int64_t exc_stride_k,
                  double* cpf_dev, double* dur_dev,
                  void* totals_dev, int64_t totals_stride_k,
                  void* forces_dec_dev, uint8_t* mask_dec_dev, int32_t forces_every,
                  void* forces_last_dev,
                  double step_size, int32_t units_per_thread /* 0 = auto */, void* stream);

/* ---- K2 with the gather of totals fused into its epilogue (multi-GPU) -----------------------
 * Same as mb_step_fused, but every per-(step, muscle) total is also stored into `n_peers` further
 * buffers -- in practice the gather buffers of the other GPUs of the box, mapped into this
 * process (CUDA IPC / symmetric memory) -- with plain stores that travel over NVLink while the
 * kernel keeps computing.  The local copy uses totals_stride_k; in a peer buffer element
 * (step k, local muscle m) lives at (m >> 1) * peer_stride_pair + k * peer_stride_k + (m & 1):
 * (stride_k, stride_pair) = (M_total, 2) is the ordinary [K, M] layout, (2, 2 * K_max) the
 * pair-major layout in which a warp's stores of a 32-step tile are 256 contiguous bytes.  The
 * caller passes pointers that already include this rank's region offset and synchronises the
 * ranks afterwards (a barrier) before anyone reads.  The reference has no counterpart
 * (single process); it replaces an all_gather after the window.  `peer_totals_dev` is a HOST
 * array of device pointers, n_peers <= 8. */
int mb_step_fused_gather(const MbConfig* cfg, const void* params_dev,
                         int64_t M, int32_t n, int32_t K,
                         const void* exc_dev, int64_t exc_stride_k,
                         double* cpf_dev, double* dur_dev,
                         void* totals_dev, int64_t totals_stride_k,
                         void* const* peer_totals_dev, int32_t n_peers,
                         int64_t peer_stride_k, int64_t peer_stride_pair,
                         void* forces_last_dev, double step_size, void* stream);

/* ---- the fused window with every option -----------------------------------------------------------
 * mb_step_fused and mb_step_fused_gather are views of this call.  Zero-initialise the struct; 0 in
 * exc_hold / totals_every / S means 1.
 *   rows, L, S, seg_off_dev, seg_div_dev   batch layout as in mb_step; muscle m = row * S + segment.
 *                    Uniform batches (S == 1) with 2 <= L <= 480 run the register-resident warp / block
 *                    kernels; wider muscles and rows of several muscles run the block-per-muscle kernel
 *                    (FP32 mode), so that e.g. 1,000 steps of StandardMuscle(2,000 units) or of a 30-muscle
 *                    arm are ONE launch instead of 1,000 (muscle.py:144-147, :189-243 make such muscles
 *                    first-class).
 *   exc_hold         h: step k reads excitation row k / h -- a control signal held for h simulation
 *                    steps (exc has ceil(K / h) rows), as in a control loop that runs slower than the
 *                    simulation (README.md:122-130 feeds one value for the whole run)
 *   totals_every     t: only the totals of steps t-1, 2t-1, ... are stored, at row k / t (K / t rows) */
typedef struct MbFusedArgs {
    int64_t rows, L;
    int32_t S, K;
    const int32_t* seg_off_dev;
    const double* seg_div_dev;
    const void* exc_dev;
    int64_t exc_stride_k;
    int32_t exc_hold;
    int32_t totals_every;
    double* cpf_dev;
    double* dur_dev;
    void* totals_dev;
    int64_t totals_stride_k;
    void* forces_dec_dev;
    uint8_t* mask_dec_dev;
    int32_t forces_every;
    int32_t units_per_thread;
    void* forces_last_dev;
    void* const* peer_totals_dev;
    int32_t n_peers;
    int32_t max_segment_units;   /* S > 1: the longest muscle of a row (the host knows seg_off; the library only has the device copy) */
    int64_t peer_stride_k, peer_stride_pair;
    int64_t peer_stride_odd;     /* stride of the odd muscle of a pair in a peer buffer (0 = 1): element (k, m) lives at
                                    (m >> 1) * peer_stride_pair + k * peer_stride_k + (m & 1) * peer_stride_odd, so that
                                    (stride_k, stride_pair, stride_odd) = (1, 2 * K_max, K_max) is the muscle-major layout in which
                                    the FP64 warp kernel (one warp per muscle) stores 256 contiguous bytes per 32 steps */
    int32_t peer_multicast;      /* 1: peer_totals_dev[0] is an NVLink multicast address (n_peers == 1): every total leaves this
                                    GPU once (multimem.st) and the NVSwitch replicates it into all GPUs' gather buffers */
    int32_t reserved;
} MbFusedArgs;
int mb_step_fused_ex(const MbConfig* cfg, const void* params_dev, const MbFusedArgs* args,
                     double step_size, void* stream);

/* Launch geometry mb_step_fused would use (for reporting): muscles per block,
 * threads per block, blocks, dynamic shared memory bytes. */
int mb_fused_geometry(const MbConfig* cfg, int64_t M, int32_t n, int32_t K, int32_t units_per_thread,
                      int32_t* muscles_per_block, int32_t* threads, int64_t* blocks, int32_t* smem_bytes);

/* ---- per-muscle lost capacity: StandardMuscle.get_peripheral_fatigue (muscle.py:248-256)
 * fatigue[rows, S] = sum(peak - cpf) / output_divisor, as float64 -- the reference's 1 - sum(cpf)/max_arb_output
 * with max_arb_output = sum(peak) (muscle.py:187) taken inside the sum: equal up to rounding, but exactly 0 for a rested
 * muscle whatever the order of the parallel sum, and never negative.  mb_step_env's fused readout is the same quantity. */
int mb_peripheral_fatigue(const MbConfig* cfg, const void* params_dev, int64_t rows, int64_t L, int32_t S,
                          const int32_t* seg_off_dev, const double* seg_div_dev, const double* cpf_dev,
                          double* fatigue_dev, void* stream);

/* ---- roofline denominators that MEASURED_PEAKS.json does not carry -------------------
 * Dependent-chain microkernels (8 independent chains per thread): each launch executes
 *     blocks * threads * 64 * iters
 * instructions per lane of kind 0 FFMA, 1 DFMA, 2 MUFU.EX2, 3 FFMA2 (packed FP32 pair), 4 FMNMX,
 * 5 a 5 FFMA : 2 FMNMX : 1 MUFU mix, 6 FMUL2, 7 FADD2.  The caller times the launch with CUDA
 * events.  `sink_dev` needs blocks*threads elements of 8 bytes. */
int mb_peak_probe(int32_t kind, int32_t blocks, int32_t threads, int32_t iters,
                  void* sink_dev, int64_t* lane_ops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MUSCLE_B200_H */
