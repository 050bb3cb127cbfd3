// DiffWave (cfg 5) kernels: shared declarations between diffwave.cu (plan, fp32 CUDA-core path, small kernels) and
// diffwave_tc.cu (tcgen05 path).  Residual channels C = 64 everywhere (config_diffwave.json / the reference default).
#pragma once
#include "common.cuh"

namespace sddm {

constexpr int DW_C = 64;        // residual channels
constexpr int DW_N = 128;       // 2 C: gate | filter, residual | skip
constexpr int DW_EMB = 512;     // diffusion-embedding width (diffwave.py:30-31)

// one residual layer on the tcgen05 path (diffwave.py:85-108), time tile = 128 samples:
//   acc1 = sum_tap x[t + (tap-1) d] . Wd_tap        (TMEM, fp32)   + cond[t] (cached conditioner + biases) + bias1[v(t)]
//   z    = sigmoid(acc1[:64]) * tanh(acc1[64:])     (bf16, shared memory: operand of the second MMA AND stored to the z cache)
//   acc2 = z . W_res
//   x_out = (x + acc2 + b_res) / sqrt(2)  (bf16)
// The skip branch is NOT evaluated per layer: skip_l = W_skip_l z_l is linear in z_l, so the layer only stores z_l (128 B per
// sample instead of a 512 B fp32 read-modify-write) and launch_dw_final_tc contracts all layers at once (K = 64 L).
struct DwLayerTc {
    const __nv_bfloat16* x_in;    // [B][T][64]
    __nv_bfloat16* x_out;         // [B][T][64]  (a different buffer: other CTAs still read x_in halos)
    const __nv_bfloat16* cond;    // [B][T][128] this layer's cached conditioner (+ dilated_conv bias + conditioner bias)
    __nv_bfloat16* zc;            // [B][T][64]  this layer's slice of the z cache
    const float* bias1;           // [B][layers][4][128]: W_tap . e summed over the in-bounds taps, variant v = (t<d) | (t+d>=T)<<1
    int bias1_row_stride;         // floats between batch rows (layers * 4 * 128)
    const __nv_bfloat16* w1;      // [3][128][64] K-major (tap, n, c)
    const __nv_bfloat16* w2;      // [64][64]    K-major (n = residual channel, c)
    const float* b2;              // [64] output_residual bias
    int B, T, dil;
    int opaque_zero;              // must be 0 (see the release of the operand / conditioner buffers in the kernel)
};
int launch_dw_layer_tc(const DwLayerTc& p, cudaStream_t st);

// skip / output head on the tcgen05 path (diffwave.py:150-153):
//   s = (sum_l z_l . W_skip_l^T + sum_l b_skip_l) / sqrt(L);  eps = w_out . relu(W_sp s + b_sp) + b_out
struct DwFinalTc {
    const __nv_bfloat16* zc;      // [L][B][T][64]
    const __nv_bfloat16* ws;      // [L][64][64] K-major (n, c)
    const __nv_bfloat16* wsp;     // [64][64]    K-major skip_projection
    const float* bsum;            // [64]  sum over layers of the output_projection (skip) biases
    const float* bsp;             // [64]
    const float* wo;              // [64]
    float bo, inv_sqrt_layers;
    float* eps;                   // [B][T]
    int L, B, T;
};
int launch_dw_final_tc(const DwFinalTc& p, cudaStream_t st);

// conditioner projection on the tcgen05 path: out[t][n] = sum_f up[t][f] W[n][f] + bias[n], bf16 output
struct DwCondTc {
    const __nv_bfloat16* up;      // [T][KP] one utterance, K padded with zeros to a multiple of 64
    const __nv_bfloat16* w;       // [128][KP]
    const float* bias;            // [128]
    __nv_bfloat16* out;           // [T][128]
    int T, KP;
};
int launch_dw_cond_tc(const DwCondTc& p, cudaStream_t st);

}  // namespace sddm
