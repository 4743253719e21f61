// k_stream_f32.cu -- FP32-mode instantiations of the streaming single-step kernel (mb_stream.cuh).
#include "k_stream_impl.cuh"
namespace mb {
int launch_k1s_f32(MB_K1S_PARAMS) {
    return launch_k1s<float>(c, params, rows, L, S, seg_off, seg_div, in, in_per_unit, cpf, dur, forces, totals, ep, dt, st,
                             scal_f(c));
}
}  // namespace mb
