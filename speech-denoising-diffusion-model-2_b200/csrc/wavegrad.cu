// cfg 4: WaveGrad denoiser (reference model/wavegrad.py:20-179) + SDDM_spectrogram.infer (model/model.py:206-257).
//
// fp32 path (this file): CUDA cores.  Activations are time-major [B][L][C]; every Conv1d of the network (k = 1 / 3,
// any dilation) is ONE launch of a tiled GEMM whose operand loader fuses everything the reference applies to the conv input:
//   nearest-neighbour F.interpolate (x factor or / factor: an index map), the FiLM affine shift + scale * x, leaky_relu(0.2)
//   and the zero padding (applied last, in the post-activation domain),
// and whose epilogue fuses bias, the FiLM-branch leaky_relu + positional encoding, and the block's residual add.
// 56 launches per eps_hat instead of the reference's ~200 eager kernels.  With precision = SDDM_PREC_BF16 the 52 conv launches run
// wavegrad_tc.cu's tcgen05 kernel instead (forward_tc below: activations stored in the form their consumers read).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/sddm_b200.h"
#include "kernels.cuh"
#include "wavegrad.cuh"

namespace sddm {
namespace {

constexpr int BM = 64, BN = 64, BK = 16;
constexpr int WG_HOP = 300, WG_MELS = 128;

struct WgConv {
    const float* in; int Lin, Cin;      // [B][Lin][Cin]
    int L;                              // output length = length of the (virtual) interpolated input
    int up, down;                       // nearest map of F.interpolate: src = tau / up (x up) or tau * down (/ down)
    const float* film;                  // nullable [B][L][2 Cin]: shift = [.., c], scale = [.., Cin + c]   (wavegrad.py:70,99,104,106)
    int pre_lrelu;                      // leaky_relu(0.2) on the conv input
    const float* w; int K, dil;         // [K * Cin][Cout], reduction index tap * Cin + ci; tap reads tau = t + (tap - K/2) * dil
    const float* bias; int Cout;
    int post_lrelu;                     // FiLM.input_conv: leaky_relu after the bias (wavegrad.py:67)
    const float* pe; int pe_stride;     // nullable: + pe[b * pe_stride + co]  (wavegrad.py:68)
    const float* add;                   // nullable [B][L][Cout]: residual / parallel branch
    float* out;                         // [B][L][Cout]
};

__device__ __forceinline__ float lrelu02(float v) { return v > 0.f ? v : __fmul_rn(v, 0.2f); }

__global__ void __launch_bounds__(256) wg_conv_fp32(WgConv p) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int t0 = blockIdx.x * BM, n0 = blockIdx.y * BN, b = blockIdx.z;
    const float* inb = p.in + (size_t)b * p.Lin * p.Cin;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int arow = tid >> 2, akq = (tid & 3) * 4;
    const int bk = tid >> 4, bn4 = (tid & 15) * 4;
    const int t = t0 + arow, Ktot = p.K * p.Cin;
    for (int k0 = 0; k0 < Ktot; k0 += BK) {
        const int tap = k0 / p.Cin, c0 = k0 - tap * p.Cin;
        const int tau = t + (tap - p.K / 2) * p.dil;
        float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < p.L && tau >= 0 && tau < p.L) {
            const int src = p.up > 1 ? tau / p.up : (p.down > 1 ? tau * p.down : tau);
            av = *reinterpret_cast<const float4*>(inb + (size_t)src * p.Cin + c0 + akq);
            if (p.film) {
                const float* fp = p.film + ((size_t)b * p.L + tau) * 2 * p.Cin + c0 + akq;
                const float4 sh = *reinterpret_cast<const float4*>(fp), sc = *reinterpret_cast<const float4*>(fp + p.Cin);
                av.x = __fadd_rn(sh.x, __fmul_rn(sc.x, av.x));
                av.y = __fadd_rn(sh.y, __fmul_rn(sc.y, av.y));
                av.z = __fadd_rn(sh.z, __fmul_rn(sc.z, av.z));
                av.w = __fadd_rn(sh.w, __fmul_rn(sc.w, av.w));
            }
            if (p.pre_lrelu) { av.x = lrelu02(av.x); av.y = lrelu02(av.y); av.z = lrelu02(av.z); av.w = lrelu02(av.w); }
        }
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n0 + bn4 < p.Cout) bv = *reinterpret_cast<const float4*>(p.w + (size_t)(k0 + bk) * p.Cout + n0 + bn4);
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
        if (r >= p.L) continue;
        const size_t row = (size_t)b * p.L + r;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx + 16 * j;
            if (n >= p.Cout) continue;
            float v = acc[i][j] + __ldg(p.bias + n);
            if (p.post_lrelu) v = lrelu02(v);
            if (p.pe) v += __ldg(p.pe + (size_t)b * p.pe_stride + n);
            if (p.add) v += __ldg(p.add + row * p.Cout + n);
            p.out[row * p.Cout + n] = v;
        }
    }
}

// downsample.0: Conv1d(1, 32, 5, padding = 2) on the raw audio (wavegrad.py:144)
__global__ void __launch_bounds__(256) wg_first_kernel(const float* __restrict__ audio, const float* __restrict__ w /* [32][5] */,
                                                       const float* __restrict__ bias, float* __restrict__ out, int B, int L) {
    const int64_t total = (int64_t)B * L * 8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i & 7) * 4;
        const int64_t row = i >> 3;
        const int b = (int)(row / L), t = (int)(row - (int64_t)b * L);
        float a[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const int tt = t + k - 2;
            a[k] = (tt >= 0 && tt < L) ? __ldg(audio + (int64_t)b * L + tt) : 0.f;
        }
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 5; ++k) s = fmaf(__ldg(w + (c + q) * 5 + k), a[k], s);
            v[q] = s + __ldg(bias + c + q);
        }
        reinterpret_cast<float4*>(out)[i] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// tcgen05 path: the same conv, stored as the two bf16 operand tensors its consumers read (raw and leaky_relu'd), channels
// zero-padded from 32 to 64 (one 128-byte TMA / UMMA swizzle atom per row)
__global__ void __launch_bounds__(256) wg_first_kernel_tc(const float* __restrict__ audio, const float* __restrict__ w, const float* __restrict__ bias,
                                                          __nv_bfloat16* __restrict__ raw16, __nv_bfloat16* __restrict__ act16, int B, int L) {
    const int64_t total = (int64_t)B * L * 16;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i & 15) * 4;
        const int64_t row = i >> 4;
        const int b = (int)(row / L), t = (int)(row - (int64_t)b * L);
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (c < 32) {
            float a[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const int tt = t + k - 2;
                a[k] = (tt >= 0 && tt < L) ? __ldg(audio + (int64_t)b * L + tt) : 0.f;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float s = 0.f;
#pragma unroll
                for (int k = 0; k < 5; ++k) s = fmaf(__ldg(w + (c + q) * 5 + k), a[k], s);
                v[q] = s + __ldg(bias + c + q);
            }
        }
        __nv_bfloat162 r0 = __floats2bfloat162_rn(v[0], v[1]), r1 = __floats2bfloat162_rn(v[2], v[3]);
        __nv_bfloat162 a0 = __floats2bfloat162_rn(lrelu02(v[0]), lrelu02(v[1])), a1 = __floats2bfloat162_rn(lrelu02(v[2]), lrelu02(v[3]));
        uint2 pr, pa;
        pr.x = *reinterpret_cast<uint32_t*>(&r0); pr.y = *reinterpret_cast<uint32_t*>(&r1);
        pa.x = *reinterpret_cast<uint32_t*>(&a0); pa.y = *reinterpret_cast<uint32_t*>(&a1);
        reinterpret_cast<uint2*>(raw16)[i] = pr;
        reinterpret_cast<uint2*>(act16)[i] = pa;
    }
}

__global__ void __launch_bounds__(256) wg_transpose_spec_tc(const float* __restrict__ spec, __nv_bfloat16* __restrict__ out, int B, int C, int F) {
    const int64_t total = (int64_t)B * C * F;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int64_t r = i / C;
        const int f = (int)(r % F), b = (int)(r / F);
        out[i] = __float2bfloat16_rn(__ldg(spec + ((int64_t)b * C + c) * F + f));
    }
}

__global__ void __launch_bounds__(256) wg_bf16_to_f32(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int64_t rows, int cols, int ld) {
    const int64_t total = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols;
        dst[i] = __bfloat162float(src[r * ld + (int)(i - r * cols)]);
    }
}

// last_conv: Conv1d(128, 1, 3, padding = 1) (wavegrad.py:165); one warp per output sample, lanes over channels
__global__ void __launch_bounds__(256) wg_last_kernel(const float* __restrict__ x /* [B][L][128] */, const float* __restrict__ w /* [3][128] */,
                                                      float bias, float* __restrict__ out, int B, int L) {
    const int lane = threadIdx.x & 31;
    const int64_t total = (int64_t)B * L;
    for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < total; row += (int64_t)gridDim.x * 8) {
        const int b = (int)(row / L), t = (int)(row - (int64_t)b * L);
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int tt = t + k - 1;
            if (tt < 0 || tt >= L) continue;
            const float4 a = *reinterpret_cast<const float4*>(x + ((int64_t)b * L + tt) * 128 + lane * 4);
            const float4 ww = *reinterpret_cast<const float4*>(w + k * 128 + lane * 4);
            s = fmaf(a.x, ww.x, s); s = fmaf(a.y, ww.y, s); s = fmaf(a.z, ww.z, s); s = fmaf(a.w, ww.w, s);
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
        if (lane == 0) out[row] = s + bias;
    }
}

// PositionalEncoding of the 5 FiLMs (wavegrad.py:44-49): pe[b][off_i + c] = sin / cos(level_b * freq_i[c mod dim/2])
struct PeDims { int dim[5]; int off[5]; int foff[5]; };
__global__ void __launch_bounds__(256) wg_pe_kernel(const float* __restrict__ level, float level_scalar, const float* __restrict__ freq,
                                                     float* __restrict__ pe, PeDims d, int total) {
    const int b = blockIdx.x;
    const float nl = level ? level[b] : level_scalar;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        int f = 0;
        while (f < 4 && i >= d.off[f + 1]) ++f;
        const int c = i - d.off[f], half = d.dim[f] / 2;
        const float a = __fmul_rn(nl, __ldg(freq + d.foff[f] + (c < half ? c : c - half)));
        pe[(size_t)b * total + i] = c < half ? sinf(a) : cosf(a);
    }
}

// [B][128][F] -> [B][F][128]
__global__ void __launch_bounds__(256) wg_transpose_spec(const float* __restrict__ spec, float* __restrict__ out, int B, int C, int F) {
    const int64_t total = (int64_t)B * C * F;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int64_t r = i / C;
        const int f = (int)(r % F), b = (int)(r / F);
        out[i] = __ldg(spec + ((int64_t)b * C + c) * F + f);
    }
}

inline int grid_1d(int64_t n, int block) {
    int64_t g = (n + block - 1) / block;
    if (g > 148 * 32) g = 148 * 32;
    return (int)(g < 1 ? 1 : g);
}
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

const int kDown[4][3] = {{32, 128, 2}, {128, 128, 2}, {128, 256, 3}, {256, 512, 5}};
const int kFilm[5][2] = {{32, 128}, {128, 128}, {128, 256}, {256, 512}, {512, 512}};
const int kUp[5][7] = {{768, 512, 5, 1, 2, 1, 2}, {512, 512, 5, 1, 2, 1, 2}, {512, 256, 3, 1, 2, 4, 8}, {256, 128, 2, 1, 2, 4, 8}, {128, 128, 2, 1, 2, 4, 8}};

struct ConvW { int Cin, Cout, K; size_t w_off, b_off; };

}  // namespace
}  // namespace sddm

using namespace sddm;

struct sddm_wg_plan {
    sddm_wg_config cfg{};
    int T = 0;
    std::map<std::string, std::vector<int64_t>> expect;
    std::map<std::string, std::vector<float>> host_w;
    std::map<std::string, ConvW> convs;     // GEMM convs by module key
    bool have_sched = false, finalized = false;
    std::vector<float> sch[5];
    float* d_f32 = nullptr;
    __nv_bfloat16* d_bf16 = nullptr;          // tcgen05 path: [N][taps * Cin] K-major packs (polyphase-combined for the up-sampling convs)
    std::map<std::string, size_t> tc_w;      // module key -> offset into d_bf16
    bool tc = false;
    size_t o_first_w = 0, o_first_b = 0, o_last_w = 0, o_freq = 0;
    float last_b = 0.f;
    PeDims pe{};
    int pe_total = 0;
    // per-launch CUDA-event timing of the tcgen05 conv launches (bench roofline; off by default)
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev;
    size_t prof_used = 0;
    double prof_flops = 0.0, prof_bytes = 0.0;   // executed GEMM flops / algorithmic operand + result bytes of the timed launches
};

namespace sddm {
namespace {

void wg_expect_conv(sddm_wg_plan* p, const std::string& key, int cin, int cout, int k) {
    p->expect[key + ".weight"] = {cout, cin, k};
    p->expect[key + ".bias"] = {cout};
    p->convs[key] = ConvW{cin, cout, k, 0, 0};
}

// One forward = a fixed sequence of launches over bump-allocated activations.  dry == true only sizes the workspace / records
// the tensors (debug fetch); otherwise it enqueues.
struct WgRun {
    sddm_wg_plan* p;
    int B, frames;
    unsigned char* ws;
    bool dry;
    cudaStream_t st;
    size_t off = 0;
    int rc = SDDM_OK;
    struct Named { void* ptr; int L, C, ld, is16; };
    std::map<std::string, Named> named;

    float* alloc(size_t L, size_t C) {
        const size_t o = off;
        off = align_up(off + (size_t)B * L * C * sizeof(float), 1024);
        return reinterpret_cast<float*>(ws + o);   // ws may be null in a sizing run: the value is then only an offset, never dereferenced
    }
    void name(const char* n, void* ptr, int L, int C, int ld = 0, int is16 = 0) { named[n] = Named{ptr, L, C, ld ? ld : C, is16}; }
    __nv_bfloat16* alloc16(size_t L, size_t ld) {
        const size_t o = off;
        off = align_up(off + (size_t)B * L * ld * 2, 1024);
        return reinterpret_cast<__nv_bfloat16*>(ws + o);
    }

    float* conv(const std::string& key, const float* in, int Lin, int L, int up, int down, const float* film, int pre, int dil, int post,
                const float* pe, const float* add) {
        const ConvW& w = p->convs.at(key);
        float* out = alloc(L, w.Cout);
        if (dry || rc) return out;
        WgConv c{};
        c.in = in; c.Lin = Lin; c.Cin = w.Cin; c.L = L; c.up = up; c.down = down; c.film = film; c.pre_lrelu = pre;
        c.w = p->d_f32 + w.w_off; c.K = w.K; c.dil = dil; c.bias = p->d_f32 + w.b_off; c.Cout = w.Cout;
        c.post_lrelu = post; c.pe = pe; c.pe_stride = p->pe_total; c.add = add; c.out = out;
        dim3 grid((L + BM - 1) / BM, (w.Cout + BN - 1) / BN, B);
        wg_conv_fp32<<<grid, 256, 0, st>>>(c);
        count_launch();
        if (cudaPeekAtLastError() != cudaSuccess) { set_error("wavegrad conv launch failed: %s", cudaGetErrorString(cudaGetLastError())); rc = SDDM_E_CUDA; }
        return out;
    }

    int forward(const float* spec, const float* audio, const float* level_dev, float level_scalar, float* eps_out) {
        return p->tc ? forward_tc(spec, audio, level_dev, level_scalar, eps_out) : forward_fp32(spec, audio, level_dev, level_scalar, eps_out);
    }

    // ---- tcgen05 path ------------------------------------------------------------------------------------
    struct Act { __nv_bfloat16* raw16; __nv_bfloat16* act16; float* raw32; int L, C, ld; };

    // one conv launch; `a` is the operand tensor (already in consumer form), decim: row-strided (nearest down-sampled) view
    void tc(const std::string& key, const __nv_bfloat16* a, int a_ld, int a_L, int decim, int rows, const int* toff, int ntaps, int phases,
            const float* add, int add_div, int add_rows, const __nv_bfloat16* film, int act_mode, bool post_pe, const float* pe, Act& o, int want) {
        // want bits: 1 raw32, 2 raw16, 4 act16
        const ConvW& w = p->convs.at(key);
        const int H = w.Cout;
        o.L = phases * rows; o.C = H; o.ld = H < 64 ? 64 : H;
        o.raw32 = (want & 1) ? alloc(o.L, H) : nullptr;
        o.raw16 = (want & 2) ? alloc16(o.L, o.ld) : nullptr;
        o.act16 = (want & 4) ? alloc16(o.L, o.ld) : nullptr;
        if (dry || rc) return;
        WgTcConv c{};
        c.a = a; c.a_row_pitch = (long long)a_ld * decim; c.a_batch_pitch = (long long)a_L * a_ld; c.rows = rows; c.Cin = a_ld;
        c.ntaps = ntaps;
        for (int i = 0; i < ntaps; ++i) c.toff[i] = toff[i];
        c.w = p->d_bf16 + p->tc_w.at(key); c.Ntot = phases * H;
        c.bias = p->d_f32 + w.b_off; c.H = H; c.phases = phases; c.L_out = o.L;
        c.add = add; c.add_div = add_div; c.add_rows = add_rows; c.film = film; c.act_mode = (want & 4) ? act_mode : 0;
        c.post_lrelu = post_pe ? 1 : 0; c.pe = pe; c.pe_stride = p->pe_total;
        c.raw32 = o.raw32; c.raw16 = o.raw16; c.act16 = o.act16; c.ld16 = o.ld; c.B = B;
        if (p->prof_on) {
            if (p->prof_used + 2 > p->prof_ev.size())
                for (int i = 0; i < 2; ++i) { cudaEvent_t e; cudaEventCreate(&e); p->prof_ev.push_back(e); }
            cudaEventRecord(p->prof_ev[p->prof_used], st);
        }
        rc = launch_wg_conv_tc(c, st);
        if (p->prof_on) {
            cudaEventRecord(p->prof_ev[p->prof_used + 1], st);
            p->prof_used += 2;
            p->prof_flops += 2.0 * B * rows * (double)c.Ntot * ntaps * a_ld;
            p->prof_bytes += (double)B * (rows * (double)a_ld * 2 + (double)o.L * H * ((want & 1 ? 4 : 0) + (want & 2 ? 2 : 0) + (want & 4 ? 2 : 0) +
                                                                                        (add ? 4.0 / add_div : 0) + (film && act_mode == 2 ? 4 : 0))) +
                             (double)c.Ntot * ntaps * a_ld * 2;
        }
    }

    int forward_tc(const float* spec, const float* audio, const float* level_dev, float level_scalar, float* eps_out) {
        const int T = WG_HOP * frames;
        const float* W = dry ? nullptr : p->d_f32;
        float* pe = alloc(1, p->pe_total);
        __nv_bfloat16* spec_t = alloc16(frames, WG_MELS);
        Act d[5], film[5];
        d[0].L = T; d[0].C = 32; d[0].ld = 64; d[0].raw32 = nullptr;
        d[0].raw16 = alloc16(T, 64);
        d[0].act16 = alloc16(T, 64);
        if (!dry) {
            wg_pe_kernel<<<B, 256, 0, st>>>(level_dev, level_scalar, W + p->o_freq, pe, p->pe, p->pe_total);
            count_launch();
            wg_transpose_spec_tc<<<grid_1d((int64_t)B * WG_MELS * frames, 256), 256, 0, st>>>(spec, spec_t, B, WG_MELS, frames);
            count_launch();
            wg_first_kernel_tc<<<grid_1d((int64_t)B * T * 16, 256), 256, 0, st>>>(audio, W + p->o_first_w, W + p->o_first_b, d[0].raw16, d[0].act16, B, T);
            count_launch();
        }
        const int t1[1] = {0}, t3[3] = {-1, 0, 1};
        for (int i = 0; i < 5; ++i) {
            if (i > 0) {
                const int f = kDown[i - 1][2], Lin = d[i - 1].L, L = Lin / f, ld = d[i - 1].ld;
                const std::string k = "downsample." + std::to_string(i) + ".";
                Act r, a, c;
                tc(k + "residual_dense", d[i - 1].raw16, ld, Lin, f, L, t1, 1, 1, nullptr, 1, 0, nullptr, 0, false, nullptr, r, 1);
                tc(k + "conv.0", d[i - 1].act16, ld, Lin, f, L, t3, 3, 1, nullptr, 1, 0, nullptr, 1, false, nullptr, a, 4);
                const int t2[3] = {-2, 0, 2}, t4[3] = {-4, 0, 4};
                tc(k + "conv.1", a.act16, a.ld, L, 1, L, t2, 3, 1, nullptr, 1, 0, nullptr, 1, false, nullptr, c, 4);
                tc(k + "conv.2", c.act16, c.ld, L, 1, L, t4, 3, 1, r.raw32, 1, L, nullptr, 1, false, nullptr, d[i], i < 4 ? 6 : 2);
            }
            const std::string k = "film." + std::to_string(i) + ".";
            Act fa;
            tc(k + "input_conv", d[i].raw16, d[i].ld, d[i].L, 1, d[i].L, t3, 3, 1, nullptr, 1, 0, nullptr, 0, true, pe + p->pe.off[i], fa, 2);
            tc(k + "output_conv", fa.raw16, fa.ld, fa.L, 1, fa.L, t3, 3, 1, nullptr, 1, 0, nullptr, 0, false, nullptr, film[i], 2);   // bf16 [L][2H]: shift | scale
            char nm[8];
            snprintf(nm, sizeof nm, "d%d", i);
            name(nm, d[i].raw16, d[i].L, d[i].C, d[i].ld, 1);
        }
        Act x;
        tc("first_conv", spec_t, WG_MELS, frames, 1, frames, t3, 3, 1, nullptr, 1, 0, nullptr, 1, false, nullptr, x, 6);
        for (int i = 0; i < 5; ++i) {
            const int f = kUp[i][2], Lx = x.L, L = Lx * f;
            const __nv_bfloat16* fl = film[4 - i].raw16;
            const std::string k = "upsample." + std::to_string(i) + ".";
            Act b1, q, xs, r3, xo;
            tc(k + "block1", x.raw16, x.ld, Lx, 1, Lx, t1, 1, 1, nullptr, 1, 0, nullptr, 0, false, nullptr, b1, 1);
            tc(k + "block2.0", x.act16, x.ld, Lx, 1, Lx, t3, 3, f, nullptr, 1, 0, fl, 2, false, nullptr, q, 4);
            const int ta[3] = {-kUp[i][4], 0, kUp[i][4]}, tb[3] = {-kUp[i][5], 0, kUp[i][5]}, tcx[3] = {-kUp[i][6], 0, kUp[i][6]};
            tc(k + "block2.1", q.act16, q.ld, L, 1, L, ta, 3, 1, b1.raw32, f, Lx, fl, 2, false, nullptr, xs, 5);
            tc(k + "block3.0", xs.act16, xs.ld, L, 1, L, tb, 3, 1, nullptr, 1, 0, fl, 2, false, nullptr, r3, 4);
            tc(k + "block3.1", r3.act16, r3.ld, L, 1, L, tcx, 3, 1, xs.raw32, 1, L, nullptr, 1, false, nullptr, xo, i < 4 ? 6 : 1);
            x = xo;
            char nm[8];
            snprintf(nm, sizeof nm, "u%d", i);
            if (i < 4) name(nm, x.raw16, L, x.C, x.ld, 1);
            else name(nm, x.raw32, L, x.C, x.C, 0);
        }
        if (!dry && !rc) {
            wg_last_kernel<<<grid_1d(((int64_t)B * T + 7) / 8, 1), 256, 0, st>>>(x.raw32, W + p->o_last_w, p->last_b, eps_out, B, T);
            count_launch();
            if (cudaPeekAtLastError() != cudaSuccess) { set_error("wavegrad launch failed: %s", cudaGetErrorString(cudaGetLastError())); rc = SDDM_E_CUDA; }
        }
        return rc;
    }

    int forward_fp32(const float* spec, const float* audio, const float* level_dev, float level_scalar, float* eps_out) {
        const int T = WG_HOP * frames;
        const float* W = dry ? nullptr : p->d_f32;
        float* pe = alloc(1, p->pe_total);
        float* spec_t = alloc(frames, WG_MELS);
        if (!dry) {
            wg_pe_kernel<<<B, 256, 0, st>>>(level_dev, level_scalar, W + p->o_freq, pe, p->pe, p->pe_total);
            count_launch();
            wg_transpose_spec<<<grid_1d((int64_t)B * WG_MELS * frames, 256), 256, 0, st>>>(spec, spec_t, B, WG_MELS, frames);
            count_launch();
        }
        // ---- downsampling path + FiLMs (wavegrad.py:170-173)
        float* d[5];
        int Ld[5];
        float* film[5];
        Ld[0] = T;
        d[0] = alloc(T, 32);
        if (!dry) {
            wg_first_kernel<<<grid_1d((int64_t)B * T * 8, 256), 256, 0, st>>>(audio, W + p->o_first_w, W + p->o_first_b, d[0], B, T);
            count_launch();
        }
        for (int i = 0; i < 5; ++i) {
            if (i > 0) {
                const int f = kDown[i - 1][2], Lin = Ld[i - 1], L = Lin / f;
                const std::string k = "downsample." + std::to_string(i) + ".";
                float* r = conv(k + "residual_dense", d[i - 1], Lin, L, 1, f, nullptr, 0, 1, 0, nullptr, nullptr);
                float* a = conv(k + "conv.0", d[i - 1], Lin, L, 1, f, nullptr, 1, 1, 0, nullptr, nullptr);
                float* c = conv(k + "conv.1", a, L, L, 1, 1, nullptr, 1, 2, 0, nullptr, nullptr);
                d[i] = conv(k + "conv.2", c, L, L, 1, 1, nullptr, 1, 4, 0, nullptr, r);
                Ld[i] = L;
            }
            const std::string k = "film." + std::to_string(i) + ".";
            float* fa = conv(k + "input_conv", d[i], Ld[i], Ld[i], 1, 1, nullptr, 0, 1, 1, pe + p->pe.off[i], nullptr);
            film[i] = conv(k + "output_conv", fa, Ld[i], Ld[i], 1, 1, nullptr, 0, 1, 0, nullptr, nullptr);
            char nm[8];
            snprintf(nm, sizeof nm, "d%d", i);
            name(nm, d[i], Ld[i], i == 0 ? 32 : kDown[i - 1][1]);
        }
        // ---- upsampling path (wavegrad.py:175-178)
        float* x = conv("first_conv", spec_t, frames, frames, 1, 1, nullptr, 0, 1, 0, nullptr, nullptr);
        int Lx = frames;
        for (int i = 0; i < 5; ++i) {
            const int f = kUp[i][2], L = Lx * f;
            const float* fl = film[4 - i];
            const std::string k = "upsample." + std::to_string(i) + ".";
            float* b1 = conv(k + "block1", x, Lx, L, f, 1, nullptr, 0, 1, 0, nullptr, nullptr);
            float* q = conv(k + "block2.0", x, Lx, L, f, 1, nullptr, 1, kUp[i][3], 0, nullptr, nullptr);
            float* xs = conv(k + "block2.1", q, L, L, 1, 1, fl, 1, kUp[i][4], 0, nullptr, b1);
            float* r = conv(k + "block3.0", xs, L, L, 1, 1, fl, 1, kUp[i][5], 0, nullptr, nullptr);
            x = conv(k + "block3.1", r, L, L, 1, 1, fl, 1, kUp[i][6], 0, nullptr, xs);
            Lx = L;
            char nm[8];
            snprintf(nm, sizeof nm, "u%d", i);
            name(nm, x, L, kUp[i][1]);
        }
        if (!dry && !rc) {
            wg_last_kernel<<<grid_1d(((int64_t)B * T + 7) / 8, 1) , 256, 0, st>>>(x, W + p->o_last_w, p->last_b, eps_out, B, T);
            count_launch();
            if (cudaPeekAtLastError() != cudaSuccess) { set_error("wavegrad launch failed: %s", cudaGetErrorString(cudaGetLastError())); rc = SDDM_E_CUDA; }
        }
        return rc;
    }
};

int wg_ready(const sddm_wg_plan* p) {
    if (!p) { set_error("null plan"); return SDDM_E_INVALID; }
    if (!p->finalized) { set_error("plan not finalised (load weights, set schedule, call sddm_wg_plan_finalize)"); return SDDM_E_STATE; }
    return SDDM_OK;
}

size_t wg_forward_bytes(const sddm_wg_plan* p, int B, int frames) {
    WgRun r{const_cast<sddm_wg_plan*>(p), B, frames, nullptr, true, nullptr};
    r.forward(nullptr, nullptr, nullptr, 0.f, nullptr);
    return r.off;
}

struct WgLayout { size_t fwd, xt, eps, total; };
WgLayout wg_layout(const sddm_wg_plan* p, int B, int frames) {
    WgLayout l{};
    l.fwd = 0;
    const size_t fb = align_up(wg_forward_bytes(p, B, frames), 1024);
    l.xt = fb;
    l.eps = l.xt + align_up((size_t)B * WG_HOP * frames * 4, 1024);
    l.total = l.eps + align_up((size_t)B * WG_HOP * frames * 4, 1024);
    return l;
}

int wg_check_ws(const sddm_wg_plan* p, int B, int frames, const void* ws, size_t ws_bytes) {
    if (B <= 0 || frames <= 0) { set_error("batch and frame count must be positive (B=%d frames=%d)", B, frames); return SDDM_E_INVALID; }
    if (B > 65535 || (long long)frames * WG_HOP > (1ll << 30)) { set_error("batch too large / utterance too long (B=%d frames=%d)", B, frames); return SDDM_E_INVALID; }
    if (!ws) { set_error("null workspace"); return SDDM_E_INVALID; }
    if (reinterpret_cast<uintptr_t>(ws) % 256) { set_error("workspace must be 256-byte aligned"); return SDDM_E_INVALID; }
    const size_t need = wg_layout(p, B, frames).total;
    if (ws_bytes < need) { set_error("workspace too small: %zu < %zu bytes", ws_bytes, need); return SDDM_E_WORKSPACE; }
    return SDDM_OK;
}

float wg_level(const sddm_wg_plan* p, int t) {
    return p->cfg.noise_condition == SDDM_DW_COND_TIME_STEP ? (float)t : p->sch[2][t];
}

}  // namespace
}  // namespace sddm

extern "C" {

SDDM_API int sddm_wg_plan_create(const sddm_wg_config* cfg, sddm_wg_plan** out) {
    if (!cfg || !out) { set_error("null argument"); return SDDM_E_INVALID; }
    *out = nullptr;
    if (cfg->hop_samples != WG_HOP) { set_error("hop_samples must be %d (WaveGrad's 5*5*3*2*2 upsampling), got %d", WG_HOP, cfg->hop_samples); return SDDM_E_INVALID; }
    if (cfg->precision != SDDM_PREC_FP32 && cfg->precision != SDDM_PREC_BF16) { set_error("WaveGrad precision must be SDDM_PREC_FP32 or SDDM_PREC_BF16, got %d", cfg->precision); return SDDM_E_INVALID; }
    if (cfg->n_timestep < 1) { set_error("n_timestep must be positive"); return SDDM_E_INVALID; }
    if (cfg->noise_condition != SDDM_DW_COND_SQRT_ALPHA_BAR && cfg->noise_condition != SDDM_DW_COND_TIME_STEP) { set_error("unknown noise_condition %d", cfg->noise_condition); return SDDM_E_INVALID; }
    sddm_wg_plan* p = new sddm_wg_plan();
    p->cfg = *cfg;
    p->T = cfg->n_timestep;
    p->tc = cfg->precision == SDDM_PREC_BF16;
    p->expect["downsample.0.weight"] = {32, 1, 5};
    p->expect["downsample.0.bias"] = {32};
    for (int i = 1; i <= 4; ++i) {
        const std::string k = "downsample." + std::to_string(i) + ".";
        const int cin = kDown[i - 1][0], h = kDown[i - 1][1];
        wg_expect_conv(p, k + "residual_dense", cin, h, 1);
        wg_expect_conv(p, k + "conv.0", cin, h, 3);
        wg_expect_conv(p, k + "conv.1", h, h, 3);
        wg_expect_conv(p, k + "conv.2", h, h, 3);
    }
    int off = 0, foff = 0;
    for (int i = 0; i < 5; ++i) {
        const std::string k = "film." + std::to_string(i) + ".";
        wg_expect_conv(p, k + "input_conv", kFilm[i][0], kFilm[i][0], 3);
        wg_expect_conv(p, k + "output_conv", kFilm[i][0], 2 * kFilm[i][1], 3);
        p->expect[k + "encoding.frequencies"] = {kFilm[i][0] / 2};
        p->pe.dim[i] = kFilm[i][0]; p->pe.off[i] = off; p->pe.foff[i] = foff;
        off += kFilm[i][0];
        foff += kFilm[i][0] / 2;
    }
    p->pe_total = off;
    for (int i = 0; i < 5; ++i) {
        const std::string k = "upsample." + std::to_string(i) + ".";
        const int cin = kUp[i][0], h = kUp[i][1];
        wg_expect_conv(p, k + "block1", cin, h, 1);
        wg_expect_conv(p, k + "block2.0", cin, h, 3);
        wg_expect_conv(p, k + "block2.1", h, h, 3);
        wg_expect_conv(p, k + "block3.0", h, h, 3);
        wg_expect_conv(p, k + "block3.1", h, h, 3);
    }
    wg_expect_conv(p, "first_conv", WG_MELS, 768, 3);
    p->expect["last_conv.weight"] = {1, 128, 3};
    p->expect["last_conv.bias"] = {1};
    *out = p;
    return SDDM_OK;
}

SDDM_API void sddm_wg_plan_destroy(sddm_wg_plan* p) {
    if (!p) return;
    if (p->d_f32) cudaFree(p->d_f32);
    if (p->d_bf16) cudaFree(p->d_bf16);
    for (cudaEvent_t e : p->prof_ev) cudaEventDestroy(e);
    delete p;
}

SDDM_API int sddm_wg_plan_load_weight(sddm_wg_plan* p, const char* name, const void* data, const int64_t* shape, int ndim) {
    if (!p || !name || !data || !shape) { set_error("null argument"); return SDDM_E_INVALID; }
    if (p->finalized) { set_error("plan already finalised"); return SDDM_E_STATE; }
    auto it = p->expect.find(name);
    if (it == p->expect.end()) { set_error("unexpected weight '%s' for WaveGrad", name); return SDDM_E_INVALID; }
    const auto& want = it->second;
    bool ok = (int)want.size() == ndim;
    size_t n = 1;
    for (int i = 0; ok && i < ndim; ++i) { ok = want[i] == shape[i]; n *= (size_t)shape[i]; }
    if (!ok) { set_error("weight '%s': shape mismatch", name); return SDDM_E_INVALID; }
    std::vector<float> v(n);
    memcpy(v.data(), data, n * sizeof(float));
    p->host_w[name] = std::move(v);
    return SDDM_OK;
}

SDDM_API int sddm_wg_plan_set_schedule(sddm_wg_plan* p, const sddm_schedule* s, int n) {
    if (!p || !s) { set_error("null argument"); return SDDM_E_INVALID; }
    if (n != p->cfg.n_timestep + 1) { set_error("schedule length %d != n_timestep + 1 = %d", n, p->cfg.n_timestep + 1); return SDDM_E_INVALID; }
    const float* src[5] = {s->betas, s->alphas, s->sqrt_alpha_bar, s->predicted_noise_coeff, s->sigma};
    for (int i = 0; i < 5; ++i) {
        if (!src[i]) { set_error("schedule table %d is null", i); return SDDM_E_INVALID; }
        p->sch[i].assign(src[i], src[i] + n);
    }
    p->have_sched = true;
    return SDDM_OK;
}

SDDM_API int sddm_wg_plan_finalize(sddm_wg_plan* p) {
    if (!p) { set_error("null plan"); return SDDM_E_INVALID; }
    if (p->finalized) return SDDM_OK;
    if (!p->have_sched) { set_error("schedule not set"); return SDDM_E_STATE; }
    for (auto& kv : p->expect)
        if (!p->host_w.count(kv.first)) { set_error("weight '%s' was not loaded", kv.first.c_str()); return SDDM_E_STATE; }
    std::vector<float> f;
    auto put = [&](const std::vector<float>& v) { size_t o = f.size(); f.insert(f.end(), v.begin(), v.end()); while (f.size() % 4) f.push_back(0.f); return o; };
    p->o_first_w = put(p->host_w["downsample.0.weight"]);
    p->o_first_b = put(p->host_w["downsample.0.bias"]);
    {   // last_conv [1][128][3] -> [3][128]
        const auto& w = p->host_w["last_conv.weight"];
        std::vector<float> v(3 * 128);
        for (int c = 0; c < 128; ++c)
            for (int k = 0; k < 3; ++k) v[k * 128 + c] = w[c * 3 + k];
        p->o_last_w = put(v);
        p->last_b = p->host_w["last_conv.bias"][0];
    }
    {
        std::vector<float> fr;
        for (int i = 0; i < 5; ++i) {
            const auto& v = p->host_w["film." + std::to_string(i) + ".encoding.frequencies"];
            fr.insert(fr.end(), v.begin(), v.end());
        }
        p->o_freq = put(fr);
    }
    for (auto& kv : p->convs) {
        ConvW& c = kv.second;
        const auto& w = p->host_w[kv.first + ".weight"];   // [Cout][Cin][K]
        std::vector<float> v((size_t)c.K * c.Cin * c.Cout);
        for (int n = 0; n < c.Cout; ++n)
            for (int ci = 0; ci < c.Cin; ++ci)
                for (int k = 0; k < c.K; ++k) v[((size_t)k * c.Cin + ci) * c.Cout + n] = w[((size_t)n * c.Cin + ci) * c.K + k];
        c.w_off = put(v);
        c.b_off = put(p->host_w[kv.first + ".bias"]);
    }
    if (p->tc) {   // [N][taps * CinP] bf16, K-major; 32-channel inputs zero-padded to 64; up-sampling convs in polyphase form
        std::vector<__nv_bfloat16> h;
        for (auto& kv : p->convs) {
            const ConvW& c = kv.second;
            const auto& w = p->host_w[kv.first + ".weight"];   // [Cout][Cin][K]
            const int cinp = c.Cin < 64 ? 64 : c.Cin;
            int phases = 1;
            if (kv.first.size() > 9 && kv.first.compare(kv.first.size() - 8, 8, "block2.0") == 0) phases = kUp[kv.first[9] - '0'][2];
            const int ktot = c.K * cinp, ntot = phases * c.Cout;
            std::vector<float> m((size_t)ntot * ktot, 0.f);
            for (int ph = 0; ph < phases; ++ph)
                for (int k = 0; k < c.K; ++k) {
                    // tap k reads up-sampled index t + k - K/2 (dilation 1 on these layers); t = phases * s + ph -> source row s + o
                    const int num = ph + k - c.K / 2;
                    const int o = phases == 1 ? 0 : (num >= 0 ? num / phases : -((-num + phases - 1) / phases));
                    const int gt = phases == 1 ? k : o + 1;
                    for (int n = 0; n < c.Cout; ++n)
                        for (int ci = 0; ci < c.Cin; ++ci)
                            m[((size_t)ph * c.Cout + n) * ktot + (size_t)gt * cinp + ci] += w[((size_t)n * c.Cin + ci) * c.K + k];
                }
            while (h.size() % 512) h.push_back(__float2bfloat16(0.f));   // 1024-byte aligned packs (TMA base alignment)
            p->tc_w[kv.first] = h.size();
            for (float v : m) h.push_back(__float2bfloat16(v));
        }
        SDDM_CUDA_TRY(cudaMalloc(&p->d_bf16, h.size() * sizeof(__nv_bfloat16)));
        SDDM_CUDA_TRY(cudaMemcpy(p->d_bf16, h.data(), h.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
    }
    SDDM_CUDA_TRY(cudaMalloc(&p->d_f32, f.size() * sizeof(float)));
    SDDM_CUDA_TRY(cudaMemcpy(p->d_f32, f.data(), f.size() * sizeof(float), cudaMemcpyHostToDevice));
    SDDM_CUDA_TRY(cudaDeviceSynchronize());
    p->host_w.clear();
    p->finalized = true;
    return SDDM_OK;
}

SDDM_API size_t sddm_wg_workspace_bytes(const sddm_wg_plan* p, int B, int frames) {
    if (!p || B <= 0 || frames <= 0) { set_error("bad argument"); return 0; }
    return wg_layout(p, B, frames).total;
}

SDDM_API int sddm_wg_eps(sddm_wg_plan* p, const float* spec, const float* audio, const float* noise_level, int t, float* eps_out, int B,
                         int frames, void* ws, size_t ws_bytes, void* stream) {
    int rc = wg_ready(p);
    if (rc) return rc;
    if ((rc = wg_check_ws(p, B, frames, ws, ws_bytes))) return rc;
    if (!spec || !audio || !eps_out) { set_error("null buffer"); return SDDM_E_INVALID; }
    float lv = 0.f;
    if (!noise_level) {
        if (t < 0 || t > p->T) { set_error("t=%d out of range [0, %d]", t, p->T); return SDDM_E_INVALID; }
        lv = wg_level(p, t);
    }
    WgRun r{p, B, frames, reinterpret_cast<unsigned char*>(ws), false, (cudaStream_t)stream};
    return r.forward(spec, audio, noise_level, lv, eps_out);
}

SDDM_API int sddm_wg_sample(sddm_wg_plan* p, const float* spec, const float* noises, uint64_t seed, int64_t row0, float* out, float* eps_trace,
                            int B, int frames, void* ws, size_t ws_bytes, void* stream) {
    int rc = wg_ready(p);
    if (rc) return rc;
    if ((rc = wg_check_ws(p, B, frames, ws, ws_bytes))) return rc;
    if (!spec || !out) { set_error("null buffer"); return SDDM_E_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const WgLayout lay = wg_layout(p, B, frames);
    const int T = p->T, Ls = WG_HOP * frames;
    const size_t BL = (size_t)B * Ls;
    unsigned char* base = reinterpret_cast<unsigned char*>(ws);
    float* x = reinterpret_cast<float*>(base + lay.xt);
    float* eps = reinterpret_cast<float*>(base + lay.eps);
    if ((rc = launch_x_T_coef(SDDM_VAR_ORIGINAL, 0.f, 1.f, nullptr, noises, seed, row0, x, B, Ls, st))) return rc;   // model.py:216
    for (int t = T; t >= 1; --t) {
        float* e = eps_trace ? eps_trace + (size_t)(T - t) * BL : eps;
        WgRun r{p, B, frames, base, false, st};
        if ((rc = r.forward(spec, x, nullptr, wg_level(p, t), e))) return rc;
        PostP pp{};
        pp.eps_in = e;
        pp.x_in = x;
        pp.x_out = t == 1 ? out : x;
        pp.z = (noises && t > 1) ? noises + (size_t)(T + 1 - t) * BL : nullptr;
        pp.seed = seed; pp.row0 = row0;
        pp.variant = SDDM_VAR_ORIGINAL; pp.t = t; pp.T = T; pp.do_update = 1;
        pp.B = B; pp.L = Ls; pp.F = 4; pp.hop = 4; pp.n_frames = 0;
        const float k8[8] = {p->sch[3][t], sqrtf(p->sch[1][t]), p->sch[4][t], 0.f, 1.f, 0.f, 0.f, 0.f};
        if ((rc = launch_post_coef(pp, k8, st))) return rc;
    }
    return SDDM_OK;
}

SDDM_API int sddm_wg_profile_enable(sddm_wg_plan* p, int on) {
    if (!p) { set_error("null plan"); return SDDM_E_INVALID; }
    SDDM_CUDA_TRY(cudaDeviceSynchronize());
    p->prof_on = on != 0;
    p->prof_used = 0;
    p->prof_flops = p->prof_bytes = 0.0;
    return SDDM_OK;
}

SDDM_API int sddm_wg_profile_read(sddm_wg_plan* p, double* total_ms, int64_t* launches, double* flops, double* bytes) {
    if (!p || !total_ms || !launches || !flops || !bytes) { set_error("null argument"); return SDDM_E_INVALID; }
    SDDM_CUDA_TRY(cudaDeviceSynchronize());
    double tot = 0.0;
    for (size_t i = 0; i + 1 < p->prof_used; i += 2) {
        float ms = 0.f;
        SDDM_CUDA_TRY(cudaEventElapsedTime(&ms, p->prof_ev[i], p->prof_ev[i + 1]));
        tot += ms;
    }
    *total_ms = tot;
    *launches = (int64_t)(p->prof_used / 2);
    *flops = p->prof_flops;
    *bytes = p->prof_bytes;
    return SDDM_OK;
}

SDDM_API int sddm_wg_debug_fetch(sddm_wg_plan* p, const char* what, void* ws, int B, int frames, float* out, int64_t* shape2, void* stream) {
    int rc = wg_ready(p);
    if (rc) return rc;
    if (!what || !ws || !shape2) { set_error("null argument"); return SDDM_E_INVALID; }
    WgRun r{p, B, frames, reinterpret_cast<unsigned char*>(ws), true, nullptr};   // dry: replays the allocation sequence only
    r.forward(nullptr, nullptr, nullptr, 0.f, nullptr);
    auto it = r.named.find(what);
    if (it == r.named.end()) { set_error("unknown debug tensor '%s'", what); return SDDM_E_INVALID; }
    const auto& nm = it->second;
    shape2[0] = nm.L;
    shape2[1] = nm.C;
    if (!out) return SDDM_OK;
    const int64_t rows = (int64_t)B * nm.L;
    if (nm.is16) {
        wg_bf16_to_f32<<<grid_1d(rows * nm.C, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(nm.ptr), out, rows, nm.C, nm.ld);
        SDDM_LAUNCH_CHECK();
    } else {
        SDDM_CUDA_TRY(cudaMemcpyAsync(out, nm.ptr, (size_t)rows * nm.C * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    }
    return SDDM_OK;
}

}  // extern "C"
