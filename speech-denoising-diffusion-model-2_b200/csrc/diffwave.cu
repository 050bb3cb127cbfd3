// cfg 5: DiffWave denoiser (reference model/diffwave.py:22-155) + SDDM_spectrogram.infer (model/model.py:206-257).
//
// Layout in HBM (one row = one utterance of T = 256 * frames samples), time-major so that a time tile is a dense
// [rows = time][K = channels] operand:
//   x     [B][T][64]      residual stream (fp32; bf16 on the tcgen05 path, two buffers ping-pong)
//   skip  [B][T][64]      running sum of the skip branches, fp32 (fp32 path) |  z cache [L][B][T][64] bf16 (tcgen05 path: the gated
//                         activations of every layer; the skip branch is linear in them and is contracted once at the end)
//   cond  [L][B][T][128]  conditioner_projection(upsampled spectrogram) + its bias + the dilated_conv bias, for every layer:
//                         independent of the diffusion step, so it is computed once per batch (sddm_dw_condition) and only
//                         READ by the T_steps x L layer evaluations (1.23 GB per 10 s utterance in bf16; the B200 has room)
//   bias1 [B][L][4][128]  the additive diffusion-step term: conv(x + e) = conv(x) + sum over the in-bounds taps of W_tap e
//                         (zero padding applies to x + e, so rows within `dilation` of an edge drop the tap that falls outside)
// The fp32 path (this file) runs every contraction as a CUDA-core tiled GEMM (parity mode); diffwave_tc.cu holds the
// tcgen05 / TMEM / TMA kernels used when precision = SDDM_PREC_BF16.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/sddm_b200.h"
#include "diffwave.cuh"
#include "kernels.cuh"

namespace sddm {
namespace {

constexpr int BM = 64, BN = 64, BK = 16;
enum { EPI_COND = 0, EPI_GATE = 1, EPI_OUT = 2, EPI_FINAL = 3 };

// packed column n' of the gate/filter GEMM -> channel row n of the reference weight: a 64-column block holds 32 gate channels
// followed by the 32 matching filter channels, so one thread owns both halves of a channel.
__host__ __device__ inline int gate_perm(int np) {
    const int y = np >> 6, j = np & 63;
    return j < 32 ? 32 * y + j : DW_C + 32 * y + (j - 32);
}

struct GemmP {
    const float* A; int lda;        // [B][T][lda]
    int T, ntaps, dil, Kper;        // reduction index k = tap * Kper + c, operand row t + (tap - 1) * dil (zero outside [0, T))
    const float* W; int K, N;       // [K][N]
    const float* bias;              // [N]
    const float* cond;              // GATE: [B][T][N]
    const float* bias1; int bias1_stride;   // GATE: this layer's [4][N] block of row b at bias1 + b * bias1_stride
    float* out;                     // COND: [B][T][N] | GATE: z [B][T][64] | OUT: x [B][T][64], in place | FINAL: eps [B][T]
    float* skip; int first;         // OUT
    const float* wo; float bo;      // FINAL
};

template <int EPI>
__global__ void __launch_bounds__(256) dw_gemm_fp32(GemmP p) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int t0 = blockIdx.x * BM, n0 = blockIdx.y * BN, b = blockIdx.z;
    const float* Ab = p.A + (size_t)b * p.T * p.lda;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int arow = tid >> 2, akq = (tid & 3) * 4;
    const int bk = tid >> 4, bn4 = (tid & 15) * 4;
    for (int k0 = 0; k0 < p.K; k0 += BK) {
        const int tap = k0 / p.Kper, c0 = k0 - tap * p.Kper;
        const int tt = t0 + arow + (p.ntaps == 3 ? (tap - 1) * p.dil : 0);
        float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tt >= 0 && tt < p.T) av = *reinterpret_cast<const float4*>(Ab + (size_t)tt * p.lda + c0 + akq);
        const float4 bv = *reinterpret_cast<const float4*>(p.W + (size_t)(k0 + bk) * p.N + n0 + bn4);
        __syncthreads();
        As[akq + 0][arow] = av.x;
        As[akq + 1][arow] = av.y;
        As[akq + 2][arow] = av.z;
        As[akq + 3][arow] = av.w;
        *reinterpret_cast<float4*>(&Bs[bk][bn4]) = bv;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            float w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = Bs[k][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = t0 + ty * 4 + i;
        const size_t row = (size_t)b * p.T + r;
        if (EPI == EPI_COND) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + tx + 16 * j;
                p.out[row * p.N + n] = acc[i][j] + __ldg(p.bias + n);
            }
        } else if (EPI == EPI_GATE) {
            const int v = (r < p.dil ? 1 : 0) | (r + p.dil >= p.T ? 2 : 0);
            const float* b1 = p.bias1 + (size_t)b * p.bias1_stride + v * p.N;
            float a[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + tx + 16 * j;
                a[j] = acc[i][j] + __ldg(p.cond + row * p.N + n) + __ldg(b1 + n);
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {   // columns j, j + 2 = gate / filter of channel 32 * blockIdx.y + tx + 16 j
                const float z = (1.0f / (1.0f + expf(-a[j]))) * tanhf(a[j + 2]);
                p.out[row * DW_C + 32 * blockIdx.y + tx + 16 * j] = z;
            }
        } else if (EPI == EPI_OUT) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + tx + 16 * j;
                const float v = acc[i][j] + __ldg(p.bias + n);
                if (n < DW_C) {
                    float* xp = p.out + row * DW_C + n;
                    *xp = __fdiv_rn(*xp + v, 1.41421356237309515f);          // (x + residual) / sqrt(2.0), diffwave.py:108
                } else {
                    float* sp = p.skip + row * DW_C + (n - DW_C);
                    *sp = p.first ? v : *sp + v;
                }
            }
        } else {   // EPI_FINAL: eps = output_projection(relu(skip_projection(sum / sqrt(L))))   (diffwave.py:150-153)
            float part = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = tx + 16 * j;
                part = fmaf(fmaxf(acc[i][j] + __ldg(p.bias + n), 0.f), __ldg(p.wo + n), part);
            }
#pragma unroll
            for (int o = 8; o >= 1; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (tx == 0) p.out[row] = part + p.bo;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// diffusion embedding (diffwave.py:33-45): enc = [sin(s v), cos(s v)] -> Linear(128, 512) -> silu -> Linear(512, 512) -> silu
// one CTA per batch row, one warp per output (coalesced weight rows, shuffle reduction)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + expf(-v)); }

__global__ void __launch_bounds__(512) dw_embed_kernel(const float* __restrict__ step, float step_scalar, const float* __restrict__ vec,
                                                       const float* __restrict__ w1, const float* __restrict__ b1,
                                                       const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ h2) {
    __shared__ float enc[128], h1[DW_EMB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.x;
    const float s = step ? step[b] : step_scalar;
    if (tid < 64) {
        const float a = s * vec[tid];
        enc[tid] = sinf(a);
        enc[64 + tid] = cosf(a);
    }
    __syncthreads();
    for (int o = warp; o < DW_EMB; o += 16) {
        float acc = 0.f;
        for (int k = lane; k < 128; k += 32) acc = fmaf(__ldg(w1 + o * 128 + k), enc[k], acc);
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
        if (lane == 0) h1[o] = silu_f(acc + b1[o]);
    }
    __syncthreads();
    for (int o = warp; o < DW_EMB; o += 16) {
        float acc = 0.f;
        for (int k = lane; k < DW_EMB; k += 32) acc = fmaf(__ldg(w2 + o * DW_EMB + k), h1[k], acc);
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
        if (lane == 0) h2[(size_t)b * DW_EMB + o] = silu_f(acc + b2[o]);
    }
}

// e_l = diffusion_projection_l(h2) (diffwave.py:86), then bias1[b][l][v][n'] = sum over the in-bounds taps of W_tap[n] . e_l
// grid (L, B), 128 threads.  permute != 0: column order of the fp32 GEMM (gate_perm), else natural.
__global__ void __launch_bounds__(128) dw_bias1_kernel(const float* __restrict__ h2, const float* __restrict__ wp, const float* __restrict__ bp,
                                                       const float* __restrict__ wd, float* __restrict__ bias1, int L, int permute) {
    __shared__ float h[DW_EMB], e[DW_C];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, l = blockIdx.x, b = blockIdx.y;
    for (int k = tid; k < DW_EMB; k += 128) h[k] = h2[(size_t)b * DW_EMB + k];
    __syncthreads();
    for (int c = warp; c < DW_C; c += 4) {
        const float* w = wp + ((size_t)l * DW_C + c) * DW_EMB;
        float acc = 0.f;
        for (int k = lane; k < DW_EMB; k += 32) acc = fmaf(__ldg(w + k), h[k], acc);
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
        if (lane == 0) e[c] = acc + bp[l * DW_C + c];
    }
    __syncthreads();
    const int n = permute ? gate_perm(tid) : tid;
    const float* w = wd + ((size_t)l * DW_N + n) * DW_C * 3;   // [n][c][tap]
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int c = 0; c < DW_C; ++c) {
        s0 = fmaf(__ldg(w + c * 3 + 0), e[c], s0);   // tap 0 reads x[t - d]
        s1 = fmaf(__ldg(w + c * 3 + 1), e[c], s1);
        s2 = fmaf(__ldg(w + c * 3 + 2), e[c], s2);   // tap 2 reads x[t + d]
    }
    float* o = bias1 + (((size_t)b * L + l) * 4) * DW_N + tid;
    o[0 * DW_N] = s0 + s1 + s2;   // interior
    o[1 * DW_N] = s1 + s2;        // t < d: left tap outside
    o[2 * DW_N] = s0 + s1;        // t + d >= T: right tap outside
    o[3 * DW_N] = s1;             // both outside
}

// x = relu(input_projection(audio))   (diffwave.py:140-141), time-major
template <bool BF16>
__global__ void __launch_bounds__(256) dw_input_kernel(const float* __restrict__ audio, const float* __restrict__ w, const float* __restrict__ bias,
                                                       void* __restrict__ xout, int64_t total /* B * T * 16 */) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i >> 4;
        const int c = (int)(i & 15) * 4;
        const float a = __ldg(audio + row);
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = fmaxf(fmaf(__ldg(w + c + q), a, __ldg(bias + c + q)), 0.f);
        if (BF16) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&lo);
            pk.y = *reinterpret_cast<uint32_t*>(&hi);
            reinterpret_cast<uint2*>(xout)[i] = pk;
        } else {
            reinterpret_cast<float4*>(xout)[i] = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

// SpectrogramUpsampler (diffwave.py:54-61): ConvTranspose2d(1, 1, [3, 32], stride [1, 16], padding [1, 8]) + leaky_relu(0.4), twice.
//   out[f][j] = bias + sum_{kf < 3} sum_{i in {ih, ih - 1}} in[f + 1 - kf][i] * W[kf][j + 8 - 16 i],   ih = (j + 8) >> 4
// stage 1 writes time-major [16 frames][F] so that stage 2 reads it coalesced along f.
__device__ __forceinline__ float lrelu04(float v) { return v > 0.f ? v : 0.4f * v; }

__global__ void __launch_bounds__(256) dw_ups1_kernel(const float* __restrict__ spec, const float* __restrict__ w, float bias,
                                                      float* __restrict__ u1, int B, int F, int frames) {
    const int W1 = 16 * frames;
    const int64_t total = (int64_t)B * W1 * F;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int f = (int)(idx % F);
        const int64_t r = idx / F;
        const int j = (int)(r % W1), b = (int)(r / W1);
        const int ih = (j + 8) >> 4, kh = j + 8 - 16 * ih;
        const float* sp = spec + (size_t)b * F * frames;
        float acc = bias;
#pragma unroll
        for (int kf = 0; kf < 3; ++kf) {
            const int ff = f + 1 - kf;
            if (ff < 0 || ff >= F) continue;
            if (ih < frames) acc = fmaf(__ldg(sp + (size_t)ff * frames + ih), __ldg(w + kf * 32 + kh), acc);
            if (ih >= 1) acc = fmaf(__ldg(sp + (size_t)ff * frames + ih - 1), __ldg(w + kf * 32 + kh + 16), acc);
        }
        u1[idx] = lrelu04(acc);
    }
}

// stage 2 for ONE utterance: u1 [W1][F] -> up [T = 16 W1][KP] (time-major, columns f >= F zero).  One thread = one f and one
// group of 16 output times sharing the same two input columns (ih = g, g - 1): t in [16 g - 8, 16 g + 8).
template <bool BF16>
__global__ void __launch_bounds__(128) dw_ups2_kernel(const float* __restrict__ u1, const float* __restrict__ w, float bias,
                                                      void* __restrict__ up, int F, int KP, int W1) {
    __shared__ float ws[96];
    if (threadIdx.x < 96) ws[threadIdx.x] = w[threadIdx.x];
    __syncthreads();
    const int f = blockIdx.x * 128 + threadIdx.x, g = blockIdx.y;
    if (f >= KP) return;
    const int T = 16 * W1;
    float hi[3] = {0.f, 0.f, 0.f}, lo[3] = {0.f, 0.f, 0.f};
    const bool real = f < F;
    if (real) {
#pragma unroll
        for (int kf = 0; kf < 3; ++kf) {
            const int ff = f + 1 - kf;
            if (ff < 0 || ff >= F) continue;
            if (g < W1) hi[kf] = __ldg(u1 + (size_t)g * F + ff);
            if (g >= 1) lo[kf] = __ldg(u1 + (size_t)(g - 1) * F + ff);
        }
    }
#pragma unroll
    for (int kh = 0; kh < 16; ++kh) {
        const int t = 16 * g - 8 + kh;
        if (t < 0 || t >= T) continue;
        float v = 0.f;
        if (real) {
            float acc = bias;
#pragma unroll
            for (int kf = 0; kf < 3; ++kf) {
                acc = fmaf(hi[kf], ws[kf * 32 + kh], acc);
                acc = fmaf(lo[kf], ws[kf * 32 + kh + 16], acc);
            }
            v = lrelu04(acc);
        }
        if (BF16) reinterpret_cast<__nv_bfloat16*>(up)[(size_t)t * KP + f] = __float2bfloat16_rn(v);
        else reinterpret_cast<float*>(up)[(size_t)t * KP + f] = v;
    }
}

__global__ void __launch_bounds__(256) dw_bf16_to_f32_strided(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int64_t rows,
                                                              int cols, int ld) {
    const int64_t total = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols;
        const int c = (int)(i - r * cols);
        dst[i] = __bfloat162float(src[r * ld + c]);
    }
}

__global__ void __launch_bounds__(256) dw_copy_strided(const float* __restrict__ src, float* __restrict__ dst, int64_t rows, int cols, int ld,
                                                       int unpermute) {
    const int64_t total = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols;
        const int c = (int)(i - r * cols);
        const float v = src[r * ld + c];
        dst[unpermute ? r * cols + gate_perm(c) : i] = v;
    }
}

inline int grid_1d(int64_t n, int block) {
    int64_t g = (n + block - 1) / block;
    if (g > 148 * 32) g = 148 * 32;
    return (int)(g < 1 ? 1 : g);
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace
}  // namespace sddm

using namespace sddm;

// =====================================================================================================
// plan
// =====================================================================================================
struct sddm_dw_plan {
    sddm_dw_config cfg{};
    int L = 0, F = 0, KP = 0, T = 0;
    bool tc = false;
    std::map<std::string, std::vector<int64_t>> expect;
    std::map<std::string, std::vector<float>> host_w;
    bool have_sched = false, finalized = false;
    std::vector<float> sch[12];
    float* d_f32 = nullptr;
    __nv_bfloat16* d_bf16 = nullptr;
    // per-launch CUDA-event timing of the layer / head kernels (bench roofline; off by default)
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev;
    std::vector<int> prof_kind;   // 0 layer, 1 head
    size_t prof_used = 0;
    float* d_bias1_table = nullptr;   // [T+1][L][4][128]: bias1 of every diffusion step of the schedule (sampling never recomputes it)
    // offsets into d_f32
    size_t o_vec = 0, o_ew1 = 0, o_eb1 = 0, o_ew2 = 0, o_eb2 = 0, o_wp = 0, o_bp = 0, o_wd = 0, o_w1 = 0, o_wc = 0, o_bc = 0, o_w2 = 0,
           o_b2 = 0, o_wsp = 0, o_bsp = 0, o_wo = 0, o_inw = 0, o_inb = 0, o_u1w = 0, o_u2w = 0;
    float u1b = 0.f, u2b = 0.f, bo = 0.f;
    // offsets into d_bf16 (tcgen05 path)
    size_t h_w1 = 0, h_w2 = 0, h_wc = 0, h_ws = 0, h_wsp = 0;
    size_t o_b2r = 0, o_bsum = 0;   // tcgen05 path: residual biases [L][64], sum of the skip biases [64]
    // which workspace holds a valid conditioner cache
    const void* cond_ws = nullptr;
    int cond_B = 0, cond_frames = 0;
};

namespace sddm {
namespace {

struct DwLayout {
    size_t x0, x1, z, skip, cond, u1, up, h2, bias1, step, xt, eps, total;
};

DwLayout dw_layout(const sddm_dw_plan* p, int B, int frames) {
    const size_t T = (size_t)p->cfg.hop_samples * frames, BT = (size_t)B * T;
    const size_t esz = p->tc ? 2 : 4;
    DwLayout l{};
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 1024); return r; };
    l.x0 = take(BT * DW_C * esz);
    l.x1 = take(p->tc ? BT * DW_C * esz : 0);
    l.z = take(p->tc ? (size_t)p->L * BT * DW_C * 2 : BT * DW_C * 4);   // tcgen05 path: z cache of every layer
    l.skip = take(p->tc ? 0 : BT * DW_C * 4);
    l.cond = take((size_t)p->L * BT * DW_N * esz);
    l.u1 = take((size_t)B * 16 * frames * p->F * 4);
    l.up = take(T * p->KP * esz);
    l.h2 = take((size_t)B * DW_EMB * 4);
    l.bias1 = take((size_t)B * p->L * 4 * DW_N * 4);
    l.step = take((size_t)B * 4);
    l.xt = take(BT * 4);
    l.eps = take(BT * 4);
    l.total = o;
    return l;
}

template <typename T>
T* at(void* ws, size_t off) { return reinterpret_cast<T*>(reinterpret_cast<unsigned char*>(ws) + off); }

int dw_ready(const sddm_dw_plan* p) {
    if (!p) { set_error("null plan"); return SDDM_E_INVALID; }
    if (!p->finalized) { set_error("plan not finalised (load weights, set schedule, call sddm_dw_plan_finalize)"); return SDDM_E_STATE; }
    return SDDM_OK;
}

int dw_check_ws(const sddm_dw_plan* p, int B, int frames, const void* ws, size_t ws_bytes) {
    if (B <= 0 || frames <= 0) { set_error("batch and frame count must be positive (B=%d frames=%d)", B, frames); return SDDM_E_INVALID; }
    if (!ws) { set_error("null workspace"); return SDDM_E_INVALID; }
    if (reinterpret_cast<uintptr_t>(ws) % 1024) { set_error("workspace must be 1024-byte aligned"); return SDDM_E_INVALID; }
    const size_t need = dw_layout(p, B, frames).total;
    if (ws_bytes < need) { set_error("workspace too small: %zu < %zu bytes", ws_bytes, need); return SDDM_E_WORKSPACE; }
    return SDDM_OK;
}

template <int EPI>
int launch_gemm(const GemmP& g, int B, cudaStream_t st) {
    dim3 grid(g.T / BM, g.N / BN, B);
    dw_gemm_fp32<EPI><<<grid, 256, 0, st>>>(g);
    SDDM_LAUNCH_CHECK();
    return SDDM_OK;
}

// one eps_hat evaluation on a conditioned workspace
int dw_prof_mark(sddm_dw_plan* p, int kind, bool begin, cudaStream_t st) {
    if (!p->prof_on) return SDDM_OK;
    if (begin) {
        if (p->prof_used + 2 > p->prof_ev.size()) {
            for (int i = 0; i < 2; ++i) {
                cudaEvent_t e;
                SDDM_CUDA_TRY(cudaEventCreate(&e));
                p->prof_ev.push_back(e);
            }
            p->prof_kind.push_back(kind);
        } else {
            p->prof_kind[p->prof_used / 2] = kind;
        }
        SDDM_CUDA_TRY(cudaEventRecord(p->prof_ev[p->prof_used], st));
    } else {
        SDDM_CUDA_TRY(cudaEventRecord(p->prof_ev[p->prof_used + 1], st));
        p->prof_used += 2;
    }
    return SDDM_OK;
}

// step_dev: per-row step values (module API) or nullptr => row t of the precomputed table serves every batch row
int dw_forward(sddm_dw_plan* p, const float* audio, const float* step_dev, int t, float* eps_out, int B, int frames, void* ws,
               cudaStream_t st) {
    const DwLayout lay = dw_layout(p, B, frames);
    const int T = p->cfg.hop_samples * frames, L = p->L;
    const float* W = p->d_f32;
    const float* bias1;
    int bias1_stride;
    if (step_dev) {
        float* h2 = at<float>(ws, lay.h2);
        float* b1 = at<float>(ws, lay.bias1);
        dw_embed_kernel<<<B, 512, 0, st>>>(step_dev, 0.f, W + p->o_vec, W + p->o_ew1, W + p->o_eb1, W + p->o_ew2, W + p->o_eb2, h2);
        SDDM_LAUNCH_CHECK();
        dw_bias1_kernel<<<dim3(L, B), 128, 0, st>>>(h2, W + p->o_wp, W + p->o_bp, W + p->o_wd, b1, L, p->tc ? 0 : 1);
        SDDM_LAUNCH_CHECK();
        bias1 = b1;
        bias1_stride = L * 4 * DW_N;
    } else {
        bias1 = p->d_bias1_table + (size_t)t * L * 4 * DW_N;
        bias1_stride = 0;
    }
    const int64_t n16 = (int64_t)B * T * 16;
    float* skip = at<float>(ws, lay.skip);
    if (!p->tc) {
        float* x = at<float>(ws, lay.x0);
        float* z = at<float>(ws, lay.z);
        dw_input_kernel<false><<<grid_1d(n16, 256), 256, 0, st>>>(audio, W + p->o_inw, W + p->o_inb, x, n16);
        SDDM_LAUNCH_CHECK();
        for (int l = 0; l < L; ++l) {
            const int dil = 1 << (l % p->cfg.dilation_cycle_length);
            GemmP g{};
            g.A = x; g.lda = DW_C; g.T = T; g.ntaps = 3; g.dil = dil; g.Kper = DW_C;
            g.W = W + p->o_w1 + (size_t)l * 3 * DW_C * DW_N; g.K = 3 * DW_C; g.N = DW_N;
            g.cond = at<float>(ws, lay.cond) + (size_t)l * B * T * DW_N;
            g.bias1 = bias1 + (size_t)l * 4 * DW_N; g.bias1_stride = bias1_stride;
            g.out = z;
            int rc = launch_gemm<EPI_GATE>(g, B, st);
            if (rc) return rc;
            GemmP o{};
            o.A = z; o.lda = DW_C; o.T = T; o.ntaps = 1; o.dil = 0; o.Kper = DW_C;
            o.W = W + p->o_w2 + (size_t)l * DW_C * DW_N; o.K = DW_C; o.N = DW_N;
            o.bias = W + p->o_b2 + (size_t)l * DW_N;
            o.out = x; o.skip = skip; o.first = l == 0;
            if ((rc = launch_gemm<EPI_OUT>(o, B, st))) return rc;
        }
    } else {
        __nv_bfloat16* xa = at<__nv_bfloat16>(ws, lay.x0);
        __nv_bfloat16* xb = at<__nv_bfloat16>(ws, lay.x1);
        dw_input_kernel<true><<<grid_1d(n16, 256), 256, 0, st>>>(audio, W + p->o_inw, W + p->o_inb, xa, n16);
        SDDM_LAUNCH_CHECK();
        for (int l = 0; l < L; ++l) {
            DwLayerTc q{};
            q.x_in = (l & 1) ? xb : xa;
            q.x_out = (l & 1) ? xa : xb;
            q.cond = at<__nv_bfloat16>(ws, lay.cond) + (size_t)l * B * T * DW_N;
            q.bias1 = bias1 + (size_t)l * 4 * DW_N; q.bias1_row_stride = bias1_stride;
            q.w1 = p->d_bf16 + p->h_w1 + (size_t)l * 3 * DW_N * DW_C;
            q.w2 = p->d_bf16 + p->h_w2 + (size_t)l * DW_C * DW_C;
            q.b2 = W + p->o_b2r + (size_t)l * DW_C;
            q.zc = at<__nv_bfloat16>(ws, lay.z) + (size_t)l * B * T * DW_C;
            q.B = B; q.T = T; q.dil = 1 << (l % p->cfg.dilation_cycle_length);
            int rc = dw_prof_mark(p, 0, true, st);
            if (rc) return rc;
            if ((rc = launch_dw_layer_tc(q, st))) return rc;
            if ((rc = dw_prof_mark(p, 0, false, st))) return rc;
        }
        DwFinalTc f{};
        f.zc = at<__nv_bfloat16>(ws, lay.z);
        f.ws = p->d_bf16 + p->h_ws; f.wsp = p->d_bf16 + p->h_wsp;
        f.bsum = W + p->o_bsum; f.bsp = W + p->o_bsp; f.wo = W + p->o_wo;
        f.bo = p->bo; f.inv_sqrt_layers = (float)(1.0 / std::sqrt((double)L));
        f.eps = eps_out; f.L = L; f.B = B; f.T = T;
        int rc = dw_prof_mark(p, 1, true, st);
        if (rc) return rc;
        if ((rc = launch_dw_final_tc(f, st))) return rc;
        return dw_prof_mark(p, 1, false, st);
    }
    GemmP f{};
    f.A = skip; f.lda = DW_C; f.T = T; f.ntaps = 1; f.dil = 0; f.Kper = DW_C;
    f.W = W + p->o_wsp; f.K = DW_C; f.N = DW_C;
    f.bias = W + p->o_bsp; f.wo = W + p->o_wo; f.bo = p->bo;
    f.out = eps_out;
    return launch_gemm<EPI_FINAL>(f, B, st);
}

void dw_step_coefs(const sddm_dw_plan* p, int t, float* k8) {
    // order of sddm_schedule: betas 0, alphas 1, sqrt_alpha_bar 2, predicted_noise_coeff 3, sigma 4, ...
    k8[0] = p->sch[3][t];
    k8[1] = sqrtf(p->sch[1][t]);
    k8[2] = p->sch[4][t];
    k8[3] = 0.f; k8[4] = 1.f; k8[5] = 0.f; k8[6] = 0.f; k8[7] = 0.f;
}

float dw_step_value(const sddm_dw_plan* p, int t) {
    return p->cfg.noise_condition == SDDM_DW_COND_TIME_STEP ? (float)t : p->sch[2][t];
}

}  // namespace
}  // namespace sddm

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

SDDM_API int sddm_dw_plan_create(const sddm_dw_config* cfg, sddm_dw_plan** out) {
    if (!cfg || !out) { set_error("null argument"); return SDDM_E_INVALID; }
    *out = nullptr;
    if (cfg->residual_channels != DW_C) { set_error("residual_channels must be %d (config_diffwave.json), got %d", DW_C, cfg->residual_channels); return SDDM_E_INVALID; }
    if (cfg->hop_samples != 256) { set_error("hop_samples must be 256 (SpectrogramUpsampler is 16 x 16), got %d", cfg->hop_samples); return SDDM_E_INVALID; }
    if (cfg->freq_bins < 1 || cfg->freq_bins > 4096) { set_error("freq_bins out of range: %d", cfg->freq_bins); return SDDM_E_INVALID; }
    if (cfg->residual_layers < 1 || cfg->residual_layers > 256) { set_error("residual_layers out of range: %d", cfg->residual_layers); return SDDM_E_INVALID; }
    if (cfg->dilation_cycle_length < 1 || cfg->dilation_cycle_length > 12) { set_error("dilation_cycle_length must be in [1, 12], got %d", cfg->dilation_cycle_length); return SDDM_E_INVALID; }
    if (cfg->n_timestep < 1) { set_error("n_timestep must be positive"); return SDDM_E_INVALID; }
    if (cfg->noise_condition != SDDM_DW_COND_SQRT_ALPHA_BAR && cfg->noise_condition != SDDM_DW_COND_TIME_STEP) { set_error("unknown noise_condition %d", cfg->noise_condition); return SDDM_E_INVALID; }
    if (cfg->precision != SDDM_PREC_FP32 && cfg->precision != SDDM_PREC_BF16) { set_error("DiffWave precision must be SDDM_PREC_FP32 or SDDM_PREC_BF16, got %d", cfg->precision); return SDDM_E_INVALID; }
    sddm_dw_plan* p = new sddm_dw_plan();
    p->cfg = *cfg;
    p->L = cfg->residual_layers;
    p->F = cfg->freq_bins;
    p->tc = cfg->precision == SDDM_PREC_BF16;
    p->KP = (int)align_up((size_t)p->F, p->tc ? 64 : 16);
    p->T = cfg->n_timestep;
    auto& e = p->expect;
    e["input_projection.weight"] = {DW_C, 1, 1};
    e["input_projection.bias"] = {DW_C};
    e["diffusion_embedding.embedding_vector"] = {64};
    e["diffusion_embedding.projection1.weight"] = {DW_EMB, 128};
    e["diffusion_embedding.projection1.bias"] = {DW_EMB};
    e["diffusion_embedding.projection2.weight"] = {DW_EMB, DW_EMB};
    e["diffusion_embedding.projection2.bias"] = {DW_EMB};
    for (int i = 1; i <= 2; ++i) {
        e["spectrogram_upsampler.conv" + std::to_string(i) + ".weight"] = {1, 1, 3, 32};
        e["spectrogram_upsampler.conv" + std::to_string(i) + ".bias"] = {1};
    }
    for (int l = 0; l < p->L; ++l) {
        const std::string k = "residual_layers." + std::to_string(l) + ".";
        e[k + "dilated_conv.weight"] = {DW_N, DW_C, 3};
        e[k + "dilated_conv.bias"] = {DW_N};
        e[k + "diffusion_projection.weight"] = {DW_C, DW_EMB};
        e[k + "diffusion_projection.bias"] = {DW_C};
        e[k + "conditioner_projection.weight"] = {DW_N, p->F, 1};
        e[k + "conditioner_projection.bias"] = {DW_N};
        e[k + "output_projection.weight"] = {DW_C, DW_C, 1};
        e[k + "output_projection.bias"] = {DW_C};
        e[k + "output_residual.weight"] = {DW_C, DW_C, 1};
        e[k + "output_residual.bias"] = {DW_C};
    }
    e["skip_projection.weight"] = {DW_C, DW_C, 1};
    e["skip_projection.bias"] = {DW_C};
    e["output_projection.weight"] = {1, DW_C, 1};
    e["output_projection.bias"] = {1};
    *out = p;
    return SDDM_OK;
}

SDDM_API void sddm_dw_plan_destroy(sddm_dw_plan* p) {
    if (!p) return;
    if (p->d_f32) cudaFree(p->d_f32);
    if (p->d_bf16) cudaFree(p->d_bf16);
    if (p->d_bias1_table) cudaFree(p->d_bias1_table);
    for (cudaEvent_t e : p->prof_ev) cudaEventDestroy(e);
    delete p;
}

SDDM_API int sddm_dw_plan_load_weight(sddm_dw_plan* p, const char* name, const void* data, const int64_t* shape, int ndim) {
    if (!p || !name || !data || !shape) { set_error("null argument"); return SDDM_E_INVALID; }
    if (p->finalized) { set_error("plan already finalised"); return SDDM_E_STATE; }
    auto it = p->expect.find(name);
    if (it == p->expect.end()) { set_error("unexpected weight '%s' for this DiffWave configuration", name); return SDDM_E_INVALID; }
    const auto& want = it->second;
    bool ok = (int)want.size() == ndim;
    size_t n = 1;
    for (int i = 0; ok && i < ndim; ++i) { ok = want[i] == shape[i]; n *= (size_t)shape[i]; }
    if (!ok) { set_error("weight '%s': shape mismatch", name); return SDDM_E_INVALID; }
    std::vector<float> v(n);
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, data) == cudaSuccess && (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged)) {
        SDDM_CUDA_TRY(cudaMemcpy(v.data(), data, n * sizeof(float), cudaMemcpyDeviceToHost));
    } else {
        cudaGetLastError();
        memcpy(v.data(), data, n * sizeof(float));
    }
    p->host_w[name] = std::move(v);
    return SDDM_OK;
}

SDDM_API int sddm_dw_plan_set_schedule(sddm_dw_plan* p, const sddm_schedule* s, int n) {
    if (!p || !s) { set_error("null argument"); return SDDM_E_INVALID; }
    if (n != p->cfg.n_timestep + 1) { set_error("schedule length %d != n_timestep + 1 = %d", n, p->cfg.n_timestep + 1); return SDDM_E_INVALID; }
    const float* src[12] = {s->betas, s->alphas, s->sqrt_alpha_bar, s->predicted_noise_coeff, s->sigma, s->supportive_gamma,
                            s->supportive_sigma_hat, s->sqrt_delta, s->c_xt, s->c_yt, s->c_epst, s->sqrt_delta_estimated};
    for (int i = 0; i < 5; ++i) {   // the DiffWave loop uses p_transition 'original' only: tables 0..4
        if (!src[i]) { set_error("schedule table %d is null", i); return SDDM_E_INVALID; }
        p->sch[i].assign(src[i], src[i] + n);
    }
    p->have_sched = true;
    return SDDM_OK;
}

SDDM_API int sddm_dw_plan_finalize(sddm_dw_plan* p) {
    if (!p) { set_error("null plan"); return SDDM_E_INVALID; }
    if (p->finalized) return SDDM_OK;
    if (!p->have_sched) { set_error("schedule not set"); return SDDM_E_STATE; }
    for (auto& kv : p->expect)
        if (!p->host_w.count(kv.first)) { set_error("weight '%s' was not loaded", kv.first.c_str()); return SDDM_E_STATE; }
    const int L = p->L, F = p->F, KP = p->KP;
    std::vector<float> f;
    auto put = [&](const std::vector<float>& v) { size_t o = f.size(); f.insert(f.end(), v.begin(), v.end()); while (f.size() % 4) f.push_back(0.f); return o; };
    auto W = [&](const std::string& k) -> const std::vector<float>& { return p->host_w[k]; };
    p->o_vec = put(W("diffusion_embedding.embedding_vector"));
    p->o_ew1 = put(W("diffusion_embedding.projection1.weight"));
    p->o_eb1 = put(W("diffusion_embedding.projection1.bias"));
    p->o_ew2 = put(W("diffusion_embedding.projection2.weight"));
    p->o_eb2 = put(W("diffusion_embedding.projection2.bias"));
    p->o_inw = put(W("input_projection.weight"));
    p->o_inb = put(W("input_projection.bias"));
    p->o_u1w = put(W("spectrogram_upsampler.conv1.weight"));
    p->o_u2w = put(W("spectrogram_upsampler.conv2.weight"));
    p->u1b = W("spectrogram_upsampler.conv1.bias")[0];
    p->u2b = W("spectrogram_upsampler.conv2.bias")[0];
    p->bo = W("output_projection.bias")[0];
    p->o_wo = put(W("output_projection.weight"));
    p->o_bsp = put(W("skip_projection.bias"));
    {   // skip_projection with the 1 / sqrt(L) of diffwave.py:150 folded in: [c][n]
        std::vector<float> v((size_t)DW_C * DW_C);
        const auto& w = W("skip_projection.weight");
        const double s = 1.0 / std::sqrt((double)L);
        for (int n = 0; n < DW_C; ++n)
            for (int c = 0; c < DW_C; ++c) v[(size_t)c * DW_C + n] = (float)(w[(size_t)n * DW_C + c] * s);
        p->o_wsp = put(v);
    }
    std::vector<float> wp((size_t)L * DW_C * DW_EMB), bp((size_t)L * DW_C), wd((size_t)L * DW_N * DW_C * 3), b2((size_t)L * DW_N);
    std::vector<float> w1, wc, bc, w2;
    std::vector<__nv_bfloat16> h;
    if (!p->tc) {
        w1.resize((size_t)L * 3 * DW_C * DW_N);
        wc.assign((size_t)L * KP * DW_N, 0.f);
        w2.resize((size_t)L * DW_C * DW_N);
    }
    bc.resize((size_t)L * DW_N);
    std::vector<__nv_bfloat16> hw1, hw2, hwc, hws, hwsp;
    std::vector<float> b2r((size_t)L * DW_C), bsum(DW_C, 0.f);
    if (p->tc) {
        hw1.resize((size_t)L * 3 * DW_N * DW_C);
        hw2.resize((size_t)L * DW_C * DW_C);
        hws.resize((size_t)L * DW_C * DW_C);
        hwsp.resize((size_t)DW_C * DW_C);
        hwc.assign((size_t)L * DW_N * KP, __float2bfloat16(0.f));
        const auto& w = W("skip_projection.weight");
        for (size_t i = 0; i < hwsp.size(); ++i) hwsp[i] = __float2bfloat16(w[i]);   // [n][c]; the 1 / sqrt(L) is applied to the operand
    }
    for (int l = 0; l < L; ++l) {
        const std::string k = "residual_layers." + std::to_string(l) + ".";
        const auto& dw = W(k + "dilated_conv.weight");          // [n][c][tap]
        const auto& db = W(k + "dilated_conv.bias");
        const auto& cw = W(k + "conditioner_projection.weight"); // [n][f]
        const auto& cb = W(k + "conditioner_projection.bias");
        const auto& rw = W(k + "output_residual.weight");        // [n][c]
        const auto& rb = W(k + "output_residual.bias");
        const auto& sw = W(k + "output_projection.weight");
        const auto& sb = W(k + "output_projection.bias");
        memcpy(&wp[(size_t)l * DW_C * DW_EMB], W(k + "diffusion_projection.weight").data(), sizeof(float) * DW_C * DW_EMB);
        memcpy(&bp[(size_t)l * DW_C], W(k + "diffusion_projection.bias").data(), sizeof(float) * DW_C);
        memcpy(&wd[(size_t)l * DW_N * DW_C * 3], dw.data(), sizeof(float) * DW_N * DW_C * 3);
        for (int n = 0; n < DW_C; ++n) {
            b2[(size_t)l * DW_N + n] = rb[n];
            b2[(size_t)l * DW_N + DW_C + n] = sb[n];
            b2r[(size_t)l * DW_C + n] = rb[n];
            bsum[n] += sb[n];
        }
        for (int np = 0; np < DW_N; ++np) {
            const int n = p->tc ? np : gate_perm(np);
            bc[(size_t)l * DW_N + np] = cb[n] + db[n];
            if (!p->tc) {
                for (int tap = 0; tap < 3; ++tap)
                    for (int c = 0; c < DW_C; ++c)
                        w1[((size_t)l * 3 * DW_C + tap * DW_C + c) * DW_N + np] = dw[((size_t)n * DW_C + c) * 3 + tap];
                for (int fq = 0; fq < F; ++fq) wc[((size_t)l * KP + fq) * DW_N + np] = cw[(size_t)n * F + fq];
            } else {
                for (int tap = 0; tap < 3; ++tap)
                    for (int c = 0; c < DW_C; ++c)
                        hw1[(((size_t)l * 3 + tap) * DW_N + np) * DW_C + c] = __float2bfloat16(dw[((size_t)n * DW_C + c) * 3 + tap]);
                for (int fq = 0; fq < F; ++fq) hwc[((size_t)l * DW_N + np) * KP + fq] = __float2bfloat16(cw[(size_t)n * F + fq]);
            }
        }
        for (int n = 0; n < DW_N; ++n)
            for (int c = 0; c < DW_C; ++c) {
                const float v = n < DW_C ? rw[(size_t)n * DW_C + c] : sw[(size_t)(n - DW_C) * DW_C + c];
                if (!p->tc) w2[((size_t)l * DW_C + c) * DW_N + n] = v;
                else if (n < DW_C) hw2[((size_t)l * DW_C + n) * DW_C + c] = __float2bfloat16(v);
                else hws[((size_t)l * DW_C + (n - DW_C)) * DW_C + c] = __float2bfloat16(v);
            }
    }
    p->o_wp = put(wp); p->o_bp = put(bp); p->o_wd = put(wd); p->o_b2 = put(b2); p->o_bc = put(bc);
    p->o_b2r = put(b2r); p->o_bsum = put(bsum);
    if (!p->tc) { p->o_w1 = put(w1); p->o_wc = put(wc); p->o_w2 = put(w2); }
    SDDM_CUDA_TRY(cudaMalloc(&p->d_f32, f.size() * sizeof(float)));
    SDDM_CUDA_TRY(cudaMemcpy(p->d_f32, f.data(), f.size() * sizeof(float), cudaMemcpyHostToDevice));
    if (p->tc) {
        p->h_w1 = 0; p->h_w2 = hw1.size(); p->h_ws = p->h_w2 + hw2.size(); p->h_wsp = p->h_ws + hws.size(); p->h_wc = p->h_wsp + hwsp.size();
        h = hw1;
        h.insert(h.end(), hw2.begin(), hw2.end());
        h.insert(h.end(), hws.begin(), hws.end());
        h.insert(h.end(), hwsp.begin(), hwsp.end());
        h.insert(h.end(), hwc.begin(), hwc.end());
        SDDM_CUDA_TRY(cudaMalloc(&p->d_bf16, h.size() * sizeof(__nv_bfloat16)));
        SDDM_CUDA_TRY(cudaMemcpy(p->d_bf16, h.data(), h.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
    }
    {   // bias1 of every step of the schedule (diffwave.py:33-45,86 evaluated at the value SDDM_spectrogram.infer passes)
        const int n = p->T + 1;
        std::vector<float> steps(n);
        for (int t = 0; t < n; ++t) steps[t] = dw_step_value(p, t);
        float *d_steps = nullptr, *d_h2 = nullptr;
        SDDM_CUDA_TRY(cudaMalloc(&d_steps, n * sizeof(float)));
        SDDM_CUDA_TRY(cudaMalloc(&d_h2, (size_t)n * DW_EMB * sizeof(float)));
        SDDM_CUDA_TRY(cudaMalloc(&p->d_bias1_table, (size_t)n * L * 4 * DW_N * sizeof(float)));
        SDDM_CUDA_TRY(cudaMemcpy(d_steps, steps.data(), n * sizeof(float), cudaMemcpyHostToDevice));
        const float* Wt = p->d_f32;
        dw_embed_kernel<<<n, 512>>>(d_steps, 0.f, Wt + p->o_vec, Wt + p->o_ew1, Wt + p->o_eb1, Wt + p->o_ew2, Wt + p->o_eb2, d_h2);
        SDDM_LAUNCH_CHECK();
        dw_bias1_kernel<<<dim3(L, n), 128>>>(d_h2, Wt + p->o_wp, Wt + p->o_bp, Wt + p->o_wd, p->d_bias1_table, L, p->tc ? 0 : 1);
        SDDM_LAUNCH_CHECK();
        SDDM_CUDA_TRY(cudaDeviceSynchronize());
        cudaFree(d_steps);
        cudaFree(d_h2);
    }
    p->host_w.clear();
    p->finalized = true;
    return SDDM_OK;
}

SDDM_API size_t sddm_dw_workspace_bytes(const sddm_dw_plan* p, int B, int frames) {
    if (!p || B <= 0 || frames <= 0) { set_error("bad argument"); return 0; }
    return dw_layout(p, B, frames).total;
}

SDDM_API int sddm_dw_condition(sddm_dw_plan* p, const float* spec, int B, int frames, void* ws, size_t ws_bytes, void* stream) {
    int rc = dw_ready(p);
    if (rc) return rc;
    if ((rc = dw_check_ws(p, B, frames, ws, ws_bytes))) return rc;
    if (!spec) { set_error("null spectrogram"); return SDDM_E_INVALID; }
    if (16 * frames + 1 > 65535 || B > 65535) { set_error("utterance too long / batch too large for one launch (frames=%d, B=%d; limits 4095 frames, 65535 rows)", frames, B); return SDDM_E_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const DwLayout lay = dw_layout(p, B, frames);
    const int T = p->cfg.hop_samples * frames, W1 = 16 * frames, F = p->F, KP = p->KP, L = p->L;
    const float* Wt = p->d_f32;
    float* u1 = at<float>(ws, lay.u1);
    p->cond_ws = nullptr;
    dw_ups1_kernel<<<grid_1d((int64_t)B * W1 * F, 256), 256, 0, st>>>(spec, Wt + p->o_u1w, p->u1b, u1, B, F, frames);
    SDDM_LAUNCH_CHECK();
    for (int b = 0; b < B; ++b) {
        dim3 g2((KP + 127) / 128, W1 + 1);
        if (!p->tc) {
            float* up = at<float>(ws, lay.up);
            dw_ups2_kernel<false><<<g2, 128, 0, st>>>(u1 + (size_t)b * W1 * F, Wt + p->o_u2w, p->u2b, up, F, KP, W1);
            SDDM_LAUNCH_CHECK();
            for (int l = 0; l < L; ++l) {
                GemmP g{};
                g.A = up; g.lda = KP; g.T = T; g.ntaps = 1; g.dil = 0; g.Kper = KP;
                g.W = Wt + p->o_wc + (size_t)l * KP * DW_N; g.K = KP; g.N = DW_N;
                g.bias = Wt + p->o_bc + (size_t)l * DW_N;
                g.out = at<float>(ws, lay.cond) + ((size_t)l * B + b) * T * DW_N;
                if ((rc = launch_gemm<EPI_COND>(g, 1, st))) return rc;
            }
        } else {
            __nv_bfloat16* up = at<__nv_bfloat16>(ws, lay.up);
            dw_ups2_kernel<true><<<g2, 128, 0, st>>>(u1 + (size_t)b * W1 * F, Wt + p->o_u2w, p->u2b, up, F, KP, W1);
            SDDM_LAUNCH_CHECK();
            for (int l = 0; l < L; ++l) {
                DwCondTc q{};
                q.up = up;
                q.w = p->d_bf16 + p->h_wc + (size_t)l * DW_N * KP;
                q.bias = Wt + p->o_bc + (size_t)l * DW_N;
                q.out = at<__nv_bfloat16>(ws, lay.cond) + ((size_t)l * B + b) * T * DW_N;
                q.T = T; q.KP = KP;
                if ((rc = launch_dw_cond_tc(q, st))) return rc;
            }
        }
    }
    p->cond_ws = ws; p->cond_B = B; p->cond_frames = frames;
    return SDDM_OK;
}

static int dw_check_conditioned(const sddm_dw_plan* p, const void* ws, int B, int frames) {
    if (p->cond_ws != ws || p->cond_B != B || p->cond_frames != frames) {
        set_error("workspace is not conditioned for B=%d frames=%d: call sddm_dw_condition first", B, frames);
        return SDDM_E_STATE;
    }
    return SDDM_OK;
}

SDDM_API int sddm_dw_eps(sddm_dw_plan* p, const float* audio, const float* diffusion_step, int t, float* eps_out, int B, int frames, void* ws,
                         size_t ws_bytes, void* stream) {
    int rc = dw_ready(p);
    if (rc) return rc;
    if ((rc = dw_check_ws(p, B, frames, ws, ws_bytes))) return rc;
    if ((rc = dw_check_conditioned(p, ws, B, frames))) return rc;
    if (!audio || !eps_out) { set_error("null buffer"); return SDDM_E_INVALID; }
    if (!diffusion_step && (t < 0 || t > p->T)) { set_error("t=%d out of range [0, %d]", t, p->T); return SDDM_E_INVALID; }
    return dw_forward(p, audio, diffusion_step, t, eps_out, B, frames, ws, (cudaStream_t)stream);
}

SDDM_API int sddm_dw_sample(sddm_dw_plan* p, const float* spec, const float* noises, uint64_t seed, int64_t row0, float* out, float* eps_trace,
                            int B, int frames, void* ws, size_t ws_bytes, void* stream) {
    int rc = sddm_dw_condition(p, spec, B, frames, ws, ws_bytes, stream);
    if (rc) return rc;
    if (!out) { set_error("null output"); return SDDM_E_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const DwLayout lay = dw_layout(p, B, frames);
    const int T = p->T, Ls = p->cfg.hop_samples * frames;
    const size_t BL = (size_t)B * Ls;
    float* x = at<float>(ws, lay.xt);
    float* eps = at<float>(ws, lay.eps);
    if ((rc = launch_x_T_coef(SDDM_VAR_ORIGINAL, 0.f, 1.f, nullptr, noises, seed, row0, x, B, Ls, st))) return rc;   // model.py:216
    for (int t = T; t >= 1; --t) {
        float* e = eps_trace ? eps_trace + (size_t)(T - t) * BL : eps;
        if ((rc = dw_forward(p, x, nullptr, t, e, B, frames, ws, st))) return rc;
        PostP pp{};
        pp.eps_in = e;
        pp.x_in = x;
        pp.x_out = t == 1 ? out : x;
        pp.z = (noises && t > 1) ? noises + (size_t)(T + 1 - t) * BL : nullptr;
        pp.seed = seed; pp.row0 = row0;
        pp.variant = SDDM_VAR_ORIGINAL; pp.t = t; pp.T = T; pp.do_update = 1;
        pp.B = B; pp.L = Ls; pp.F = 4; pp.hop = 4; pp.n_frames = 0;
        float k8[8];
        dw_step_coefs(p, t, k8);
        if ((rc = launch_post_coef(pp, k8, st))) return rc;
    }
    return SDDM_OK;
}

SDDM_API int sddm_dw_profile_enable(sddm_dw_plan* p, int on) {
    if (!p) { set_error("null plan"); return SDDM_E_INVALID; }
    SDDM_CUDA_TRY(cudaDeviceSynchronize());
    p->prof_on = on != 0;
    p->prof_used = 0;
    return SDDM_OK;
}

SDDM_API int sddm_dw_profile_read(sddm_dw_plan* p, int kind, double* total_ms, int64_t* launches) {
    if (!p || !total_ms || !launches) { set_error("null argument"); return SDDM_E_INVALID; }
    SDDM_CUDA_TRY(cudaDeviceSynchronize());
    double tot = 0.0;
    int64_t n = 0;
    for (size_t i = 0; i + 1 < p->prof_used; i += 2) {
        if (p->prof_kind[i / 2] != kind) continue;
        float ms = 0.f;
        SDDM_CUDA_TRY(cudaEventElapsedTime(&ms, p->prof_ev[i], p->prof_ev[i + 1]));
        tot += ms;
        ++n;
    }
    *total_ms = tot;
    *launches = n;
    return SDDM_OK;
}

SDDM_API int sddm_dw_debug_fetch(sddm_dw_plan* p, const char* what, void* ws, int B, int frames, float* out, int64_t* n, void* stream) {
    int rc = dw_ready(p);
    if (rc) return rc;
    if (!what || !ws || !n) { set_error("null argument"); return SDDM_E_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const DwLayout lay = dw_layout(p, B, frames);
    const int64_t T = (int64_t)p->cfg.hop_samples * frames;
    const std::string w(what);
    int64_t rows = 0;
    int cols = 0, ld = 0, unperm = 0;
    size_t off = 0;
    bool is16 = p->tc;
    if (w == "upsampled") { rows = T; cols = p->F; ld = p->KP; off = lay.up; }
    else if (w == "x") { rows = B * T; cols = DW_C; ld = DW_C; off = (p->tc && (p->L & 1)) ? lay.x1 : lay.x0; }
    else if (w == "skip") {
        if (p->tc) { set_error("the tcgen05 path keeps the per-layer z cache instead of a skip sum"); return SDDM_E_INVALID; }
        rows = B * T; cols = DW_C; ld = DW_C; off = lay.skip; is16 = false;
    }
    else if (w[0] == 'z') {   // gated activation: tcgen05 path "z<l>" (z cache of layer l), fp32 path "z" (the last layer evaluated)
        const int l = p->tc ? atoi(w.c_str() + 1) : 0;
        if (l < 0 || l >= p->L) { set_error("no such layer: %s", what); return SDDM_E_INVALID; }
        rows = B * T; cols = DW_C; ld = DW_C; off = lay.z + (p->tc ? (size_t)l * B * T * DW_C * 2 : 0);
    }
    else if (w.rfind("cond", 0) == 0) {
        const int l = atoi(w.c_str() + 4);
        if (l < 0 || l >= p->L) { set_error("no such layer: %s", what); return SDDM_E_INVALID; }
        rows = B * T; cols = DW_N; ld = DW_N; off = lay.cond + (size_t)l * B * T * DW_N * (p->tc ? 2 : 4); unperm = p->tc ? 0 : 1;
    } else { set_error("unknown debug tensor '%s'", what); return SDDM_E_INVALID; }
    *n = rows * cols;
    if (!out) return SDDM_OK;
    if (is16) dw_bf16_to_f32_strided<<<grid_1d(rows * cols, 256), 256, 0, st>>>(at<__nv_bfloat16>(ws, off), out, rows, cols, ld);
    else dw_copy_strided<<<grid_1d(rows * cols, 256), 256, 0, st>>>(at<float>(ws, off), out, rows, cols, ld, unperm);
    SDDM_LAUNCH_CHECK();
    return SDDM_OK;
}

}  // extern "C"
