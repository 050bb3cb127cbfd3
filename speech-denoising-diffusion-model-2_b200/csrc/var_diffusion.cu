// SNR-adaptive diffusion (the diffusion half): every frame n of every row b carries its own linear beta
// schedule whose end value follows from the estimated SNR of that frame.
//
// reference: VariableGaussianDiffusion, model/diffusion.py:329-446.  There, get_beta_schedule (:345-359) rebuilds the whole
// [B, 1, N, T + 1] schedule with a per-row numpy.linspace on the HOST in every call - i.e. in every reverse step, for get_x_T,
// get_noise_level and p_transition alike.  Here the three numbers a step needs (beta_t, alpha_bar_t, alpha_bar_{t-1}) are recomputed
// per frame inside the kernel that uses them: one warp per frame, the <= T-term running product evaluated redundantly by every lane
// (100 multiplies), then the frame's L samples.  No schedule tensor, no host round trip.
//
// Arithmetic follows the reference operation by operation (goldens made on the CPU):
//   end     = (10^(snr / -20) / scale)^2                                   fp32
//   beta_i  = fp32(fp64(i - 1) * step + 1e-6), step = (fp64(end) - 1e-6) / (T - 1), beta_T = end     (numpy.linspace, fp64 -> fp32)
//   ab_i    = fp32(running fp64 product of the fp32 values 1 - beta_j)     (torch.cumprod keeps an fp64 accumulator on the CPU)
// every product / sum / quotient / square root of the update formulas is rounded separately (no FMA contraction).
#include "kernels.cuh"
#include "../../include/sddm_b200.h"

namespace sddm {
namespace {

struct VarCoef { float beta_t, ab_t, ab_tm1; };

// t in [0, T]; ab_tm1 is only meaningful for t >= 1
__device__ __forceinline__ VarCoef var_coef(float snr, int T, float scale, int t, float* betas_out, float* ab_out) {
    const float e = __fdiv_rn(snr, -20.0f);
    const float p = (float)pow(10.0, (double)e);
    float end = __fdiv_rn(p, scale);
    end = __fmul_rn(end, end);
    const double start = 1e-6, stop = (double)end;
    const double step = __ddiv_rn(__dsub_rn(stop, start), (double)(T - 1));
    double prod = 1.0;
    VarCoef c{0.f, 1.f, 1.f};
    if (betas_out) { betas_out[0] = 0.f; ab_out[0] = 1.f; }
    for (int i = 1; i <= t; ++i) {
        const double lin = i == T ? stop : __dadd_rn(__dmul_rn((double)(i - 1), step), start);
        const float beta = (float)lin;
        const float alpha = __fsub_rn(1.0f, beta);
        c.ab_tm1 = (float)prod;
        prod = __dmul_rn(prod, (double)alpha);
        c.beta_t = beta;
        c.ab_t = (float)prod;
        if (betas_out) { betas_out[i] = beta; ab_out[i] = c.ab_t; }
    }
    return c;
}

// whole schedules (verification / callers that want the reference's tensors): one thread per frame
__global__ void __launch_bounds__(128) var_schedule_kernel(const float* __restrict__ snr, int frames, int T, float scale,
                                                           float* __restrict__ betas, float* __restrict__ ab) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= frames) return;
    var_coef(__ldg(snr + f), T, scale, T, betas + (int64_t)f * (T + 1), ab + (int64_t)f * (T + 1));
}

__global__ void __launch_bounds__(128) var_noise_level_kernel(const float* __restrict__ snr, int frames, int T, float scale, int t,
                                                              float* __restrict__ out) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= frames) return;
    out[f] = __fsqrt_rn(var_coef(__ldg(snr + f), T, scale, t, nullptr, nullptr).ab_t);
}

enum { VAR_MIX = 0, VAR_POST = 1 };

// MODE VAR_MIX : out = s a + sqrt(1 - s^2) z, s = sqrt(ab_t)            (get_x_T with t = T, a = condition; q_stochastic, a = x_0)
// MODE VAR_POST: out = clamp((a - beta_t / sqrt(1 - ab_t) b) / sqrt(1 - beta_t) [+ sigma_t z], -1, 1)            (p_transition)
// z == nullptr: Philox4x32-10 normals keyed by the GLOBAL frame id (row0 + b) N + n, draw index `draw`
template <int MODE>
__global__ void __launch_bounds__(256) var_frames_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ snr,
                                                         const float* __restrict__ z, uint64_t seed, int64_t row0, uint32_t draw, int frames,
                                                         int L, int T, float scale, int t, float* __restrict__ out, float* __restrict__ level_out) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int f = blockIdx.x * wpb + (threadIdx.x >> 5); f < frames; f += gridDim.x * wpb) {
        const VarCoef c = var_coef(__ldg(snr + f), T, scale, t, nullptr, nullptr);
        float k0, k1, k2 = 0.f;
        bool add_noise = true;
        if (MODE == VAR_MIX) {
            k0 = __fsqrt_rn(c.ab_t);
            k1 = __fsqrt_rn(__fsub_rn(1.0f, __fmul_rn(k0, k0)));
            if (level_out && lane == 0) level_out[f] = k0;
        } else {
            k0 = __fdiv_rn(c.beta_t, __fsqrt_rn(__fsub_rn(1.0f, c.ab_t)));
            k1 = __fsqrt_rn(__fsub_rn(1.0f, c.beta_t));
            add_noise = t > 1;
            if (add_noise) k2 = __fsqrt_rn(__fmul_rn(__fdiv_rn(__fsub_rn(1.0f, c.ab_tm1), __fsub_rn(1.0f, c.ab_t)), c.beta_t));
        }
        const int64_t base = (int64_t)f * L;
        const uint64_t key = (uint64_t)row0 + (uint64_t)f;   // global frame id: row0 is given in frames (first row of the call * N)
        for (int i0 = lane * 4; i0 < L; i0 += 128) {
            float zz[4] = {0.f, 0.f, 0.f, 0.f};
            if (add_noise && !z) {
                const float4 r = philox_normal4(seed, (uint32_t)(i0 >> 2), key, draw);
                zz[0] = r.x; zz[1] = r.y; zz[2] = r.z; zz[3] = r.w;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = i0 + q;
                if (i >= L) break;
                const float zv = add_noise ? (z ? z[base + i] : zz[q]) : 0.f;
                float o;
                if (MODE == VAR_MIX) {
                    o = __fadd_rn(__fmul_rn(k0, a[base + i]), __fmul_rn(k1, zv));
                } else {
                    o = __fdiv_rn(__fsub_rn(a[base + i], __fmul_rn(k0, b[base + i])), k1);
                    if (add_noise) o = __fadd_rn(o, __fmul_rn(k2, zv));
                    o = fminf(fmaxf(o, -1.0f), 1.0f);
                }
                out[base + i] = o;
            }
        }
    }
}

int check_var(const float* snr, int B, int N, int T, float scale) {
    if (!snr || B <= 0 || N <= 0) { set_error("variable diffusion: null SNR estimate / bad batch"); return SDDM_E_INVALID; }
    if (T < 2 || T > 100000) { set_error("variable diffusion: n_timestep must be in [2, 100000], got %d", T); return SDDM_E_INVALID; }
    if (!(scale > 0.f)) { set_error("variable diffusion: snr_estimate_scale must be positive"); return SDDM_E_INVALID; }
    return SDDM_OK;
}

int frames_grid(int frames) {
    const int g = (frames + 7) / 8;
    const int cap = device_sm_count() * 8;
    return g < cap ? (g < 1 ? 1 : g) : cap;
}

}  // namespace
}  // namespace sddm

using namespace sddm;

extern "C" SDDM_API int sddm_var_schedule(const float* snr, int B, int N, int T, float scale, float* betas, float* alpha_bar, void* stream) {
    int rc = check_var(snr, B, N, T, scale);
    if (rc) return rc;
    if (!betas || !alpha_bar) { set_error("variable diffusion: schedule outputs are null"); return SDDM_E_INVALID; }
    const int frames = B * N;
    var_schedule_kernel<<<(frames + 127) / 128, 128, 0, (cudaStream_t)stream>>>(snr, frames, T, scale, betas, alpha_bar);
    SDDM_LAUNCH_CHECK();
    return SDDM_OK;
}

extern "C" SDDM_API int sddm_var_noise_level(const float* snr, int B, int N, int T, float scale, int t, float* out, void* stream) {
    int rc = check_var(snr, B, N, T, scale);
    if (rc) return rc;
    if (!out || t < 0 || t > T) { set_error("variable diffusion: t must be in [0, %d] and the output non-null", T); return SDDM_E_INVALID; }
    const int frames = B * N;
    var_noise_level_kernel<<<(frames + 127) / 128, 128, 0, (cudaStream_t)stream>>>(snr, frames, T, scale, t, out);
    SDDM_LAUNCH_CHECK();
    return SDDM_OK;
}

extern "C" SDDM_API int sddm_var_mix(const float* x, const float* snr, const float* z, uint64_t seed, int64_t row0, int B, int N, int L, int T,
                                     float scale, int t, float* out, float* noise_level, void* stream) {
    int rc = check_var(snr, B, N, T, scale);
    if (rc) return rc;
    if (!x || !out || L <= 0 || t < 1 || t > T) { set_error("variable diffusion: bad argument (null buffer, L <= 0 or t outside [1, %d])", T); return SDDM_E_INVALID; }
    const int frames = B * N;
    var_frames_kernel<VAR_MIX><<<frames_grid(frames), 256, 0, (cudaStream_t)stream>>>(x, nullptr, snr, z, seed, row0 * N, 0u, frames, L, T, scale, t, out,
                                                                                      noise_level);
    SDDM_LAUNCH_CHECK();
    return SDDM_OK;
}

extern "C" SDDM_API int sddm_var_posterior(const float* x_t, const float* eps, const float* snr, const float* z, uint64_t seed, int64_t row0, int B,
                                           int N, int L, int T, float scale, int t, float* out, void* stream) {
    int rc = check_var(snr, B, N, T, scale);
    if (rc) return rc;
    if (!x_t || !eps || !out || L <= 0 || t < 1 || t > T) { set_error("variable diffusion: bad argument (null buffer, L <= 0 or t outside [1, %d])", T); return SDDM_E_INVALID; }
    const int frames = B * N;
    var_frames_kernel<VAR_POST><<<frames_grid(frames), 256, 0, (cudaStream_t)stream>>>(x_t, eps, snr, z, seed, row0 * N, (uint32_t)(T + 1 - t), frames, L, T,
                                                                                       scale, t, out, nullptr);
    SDDM_LAUNCH_CHECK();
    return SDDM_OK;
}
