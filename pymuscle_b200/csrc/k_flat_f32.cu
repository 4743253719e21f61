// k_flat_f32.cu -- FP32-mode instantiations of the flat-tiled single-step kernel for ragged rows (mb_flat.cuh).
#include "k_flat_impl.cuh"
namespace mb {
int launch_k1f_f32(MB_K1S_PARAMS) {
    return launch_k1f<float>(c, params, rows, L, S, seg_off, seg_div, in, in_per_unit, cpf, dur, forces, totals, ep, dt, st,
                             scal_f(c));
}
}  // namespace mb
