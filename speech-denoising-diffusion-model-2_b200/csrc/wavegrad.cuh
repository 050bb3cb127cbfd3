// WaveGrad (cfg 4): declarations shared by wavegrad.cu (plan, fp32 path, small kernels) and wavegrad_tc.cu (tcgen05 conv).
#pragma once
#include "common.cuh"

namespace sddm {

// One Conv1d of WaveGrad as a tcgen05 GEMM over time-major bf16 activations:
//   acc[s][j] = sum_{tap, ci} A[b][s + toff[tap]][ci] * W[j][tap * Cin + ci]      (rows outside [0, rows) read as zero = the conv's padding)
// A is a (possibly strided) view of an activation that its PRODUCER already stored in the form this conv consumes
// (leaky_relu / FiLM affine applied, nearest down-sampling = a row-strided view).  Up-sampling convs run at the LOW resolution
// in polyphase form: GEMM column j = phase * H + c produces output time t = phases * s + phase, channel c.
struct WgTcConv {
    const __nv_bfloat16* a;            // view base: [B][rows][Cin], row pitch a_row_pitch, batch pitch a_batch_pitch (elements)
    long long a_row_pitch, a_batch_pitch;
    int rows, Cin;                     // Cin: multiple of 64 (32-channel tensors are stored zero-padded to 64)
    int ntaps; int toff[3];
    const __nv_bfloat16* w;            // [Ntot][ntaps * Cin] K-major
    int Ntot;
    // epilogue
    const float* bias;                 // [H]
    int H, phases, L_out;              // Ntot = phases * H; output rows L_out (<= phases * rows)
    const float* add; int add_div;     // nullable fp32 [B][ceil(L_out / add_div)][H]: + add[t / add_div][c]
    int add_rows;                      // rows per batch of `add`
    const __nv_bfloat16* film;         // nullable bf16 [B][L_out][2 H]: FiLM shift | scale
    int act_mode;                      // act16 = 0: not stored, 1: leaky_relu(v), 2: leaky_relu(film_shift + film_scale * v)
    int post_lrelu; const float* pe; int pe_stride;   // v = leaky_relu(v) + pe[b][c]  (FiLM.input_conv)
    float* raw32; __nv_bfloat16* raw16; __nv_bfloat16* act16;   // nullable outputs [B][L_out][ld]; ld16 for the bf16 ones, H for raw32
    int ld16;
    int B;
};
int launch_wg_conv_tc(const WgTcConv& p, cudaStream_t st);

}  // namespace sddm
