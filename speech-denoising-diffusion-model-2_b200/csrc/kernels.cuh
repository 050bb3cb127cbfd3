// Host-callable launchers of the non-convolution kernels (kernels_misc.cu).
#pragma once
#include "common.cuh"
#include "../../include/sddm_b200.h"

namespace sddm {

// x_T = a * cond + b * z                       (diffusion.py:281-320); a, b are the step-T scalars
// seed_dev (nullable): device {seed, row0} that overrides the by-value pair - a captured CUDA graph is replayed with new seeds
int launch_x_T_coef(int variant, float a, float b, const float* cond, const float* z, uint64_t seed, int64_t row0,
                    float* x_out, int B, int L, cudaStream_t st, const unsigned long long* seed_dev = nullptr);

// posterior update incl. clamp, optionally fused with the overlap-add of the final conv frames.
//   frames != nullptr : eps[s] = frames[a][j] + frames[a-1][j+hop]   (UNetModified2.py:30-41)
//   frames == nullptr : eps read from eps_in
//   do_update == 0    : only eps_out is written (sddm_eps)
struct PostP {
    const float* frames;   // [B][n_frames][F] or nullptr
    const float* eps_in;   // [B][L] or nullptr
    float* eps_out;        // [B][L] or nullptr
    const float* x_in;     // [B][L]
    float* x_out;          // [B][L]
    float* x_trace;        // [B][L] or nullptr
    const float* cond;     // [B][L] (supportive / conditional)
    const float* z;        // [B][L] injected noise or nullptr (Philox)
    uint64_t seed;
    int64_t row0;
    const unsigned long long* seed_dev;   // nullable: device {seed, row0} overriding the two fields above (CUDA-graph replay)
    int variant, t, T, do_update;
    int B, L, F, hop, n_frames;
};
// k8 = {c2, sqrt(alpha_t), noise std, gamma, 1-gamma, c_xt, c_yt, c_epst} of step t (host scalars)
int launch_post_coef(const PostP& p, const float* k8, cudaStream_t st);

struct PostCoef {  // scalars of step t, fetched on the host from the plan's host tables
    float c2, sa, sig, gam, one_m_gam, cx, cy, ce;
};

// one sample of p_transition / _sr3 / _supportive / _conditional incl. the clamp: every product, sum and quotient rounded
// separately, as the reference's eager ops do (diffusion.py:164-222)
__device__ __forceinline__ float post_one(int variant, const PostCoef& k, float x, float e, float c, float z, bool add_noise) {
    float r;
    if (variant == SDDM_VAR_SUPPORTIVE) {
        float mu = __fsub_rn(x, __fmul_rn(k.c2, e));
        r = __fdiv_rn(__fadd_rn(__fmul_rn(k.one_m_gam, mu), __fmul_rn(k.gam, c)), k.sa);
    } else if (variant == SDDM_VAR_CONDITIONAL) {
        r = __fsub_rn(__fadd_rn(__fmul_rn(k.cx, x), __fmul_rn(k.cy, c)), __fmul_rn(k.ce, e));
    } else {
        r = __fdiv_rn(__fsub_rn(x, __fmul_rn(k.c2, e)), k.sa);
    }
    if (add_noise) r = __fadd_rn(r, __fmul_rn(k.sig, z));
    return fminf(fmaxf(r, -1.0f), 1.0f);
}



// forward diffusion q(x_t | x_0) with per-row coefficients (diffusion.py:225-279); see q_sample_kernel
int launch_q_sample(int mode, const float* coef, const float* x0, const float* y, const float* z, uint64_t seed, int64_t row0, float* x_t,
                    float* combined, float* z_out, int B, int L, cudaStream_t st);

int launch_frames(const float* sig, float* frames, int B, int n, int F, int hop, cudaStream_t st);
int launch_overlap_add(const float* frames, float* sig, int B, int n, int F, int hop, cudaStream_t st);

// noise-level embedding: PE -> Linear -> Swish -> Linear -> Swish -> all per-block Linear(inner -> Cout)
// out[row][E]                                   (UNetModified2.py:49-89,168-174)
struct TembP {
    const float* nl;       // [rows] noise levels
    const float* freq;     // [inner/2] PE frequencies
    const float* w1; const float* b1;   // [4*inner][inner]
    const float* w2; const float* b2;   // [inner][4*inner]
    const float* wn; const float* bn;   // [E][inner], [E]
    float* out;            // [rows][E]
    int rows, inner, E;
};
int launch_temb(const TembP& p, cudaStream_t st);

// stem: framing + channel concat + conv3x3(2 -> CO), NHWC output + GroupNorm partial statistics
struct StemP {
    const float* cond; const float* x_t;   // [B][L]
    const float* w;      // [2][9][CO] fp32 pack
    const float* bias;   // [CO]
    float* out;          // [B][H][W][CO]
    float* parts;        // [B][nparts][CO][2], nparts = tiles per sample (16 x 8 pixels each)
    int B, L, H, W, hop, CO, nparts;
    int act16;           // out is bf16 [B][H][W][CO]
    int gn_on;           // finalise the consumer's GroupNorm in this kernel (gn_fuse.cuh)
    GnFuse gn;
};
int launch_stem(const StemP& p, cudaStream_t st);
int launch_stem_row(const StemP& p, const __nv_bfloat16* w_row, uint32_t w_bytes, cudaStream_t st);   // conv_row.cu
int stem_nparts(int H, int W);

// GroupNorm finalize: partial sums -> per-(sample, channel) scale / shift
struct GnP {
    const float* parts[2]; int C[2]; int nparts[2]; int nsrc;
    const float* gamma; const float* beta;   // [Ctot]
    float* scale; float* shift;              // [B][Ctot]
    int B, Ctot, groups, HW;
    float eps;
};
int launch_gn_finalize(const GnP& p, cudaStream_t st);

// final Block: GN-apply + Swish + conv3x3(C -> 1); output frames [B][H][W]
struct FinalP {
    const float* x; const float* scale; const float* shift;   // [B][H][W][C], [B][C]
    const float* w;     // [9][C]
    float bias;
    float* frames;      // [B][H][W]
    int B, H, W, C;
    int act16;          // x is bf16
    int fast_math;      // approximate exp / divide in the Swish (bf16 mode); exact-ish expf / IEEE divide otherwise
};
int launch_final_conv(const FinalP& p, cudaStream_t st);

// dataset edge on the device: utterances back to back in `flat` <-> rows of the [N, 1, T] batch (kernels_misc.cu)
int launch_chunk_rows(const float* flat, const int64_t* sample_off, const int64_t* row_off, int n_utt, int T, int64_t row_lo, int64_t row_hi,
                      float* rows, cudaStream_t st);
int launch_regroup_rows(const float* rows, const int64_t* sample_off, const int64_t* row_off, int n_utt, int T, int64_t row_lo, int64_t row_hi,
                        float* flat_out, cudaStream_t st);

// debug / tests: bf16 -> fp32 copy of an activation tensor
int launch_bf16_to_f32(const void* src, float* dst, size_t n, cudaStream_t st);

}  // namespace sddm
