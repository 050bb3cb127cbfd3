// HBM-bound kernels of the reverse-diffusion loop: x_T initialisation, posterior update fused with
// overlap-add, framing helpers, noise-level embedding, stem (framing + cat + conv 2->C), GroupNorm
// finalize and the final Block (GN + Swish + conv C->1).  All fp32, all coalesced / vectorised.
#include "kernels.cuh"
#include "../../include/sddm_b200.h"

namespace sddm {

// ===================================================================================================
// x_T                                                   reference: model/diffusion.py:281-320
// ===================================================================================================
__global__ void __launch_bounds__(256) x_T_kernel(int variant, float a, float b, const float4* __restrict__ cond,
                                                  const float4* __restrict__ z, uint64_t seed, int64_t row0,
                                                  float4* __restrict__ out, int B, int L4) {
    const int64_t total = (int64_t)B * L4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(i / L4), e4 = (int)(i - (int64_t)row * L4);
        float4 zz = make_float4(0.f, 0.f, 0.f, 0.f), c = zz;
        if (variant != SDDM_VAR_SUPPORTIVE) zz = z ? z[i] : philox_normal4(seed, (uint32_t)e4, (uint64_t)(row0 + row), 0u);
        if (variant == SDDM_VAR_CONDITION_IN || variant == SDDM_VAR_CONDITIONAL || variant == SDDM_VAR_SUPPORTIVE) c = cond[i];
        float4 o;
        if (variant == SDDM_VAR_CONDITION_IN || variant == SDDM_VAR_CONDITIONAL) {
            // separately rounded mul, mul, add: bit-identical to the eager op sequence of the reference
            o.x = __fadd_rn(__fmul_rn(a, c.x), __fmul_rn(b, zz.x));
            o.y = __fadd_rn(__fmul_rn(a, c.y), __fmul_rn(b, zz.y));
            o.z = __fadd_rn(__fmul_rn(a, c.z), __fmul_rn(b, zz.z));
            o.w = __fadd_rn(__fmul_rn(a, c.w), __fmul_rn(b, zz.w));
        } else if (variant == SDDM_VAR_SUPPORTIVE) {
            o = c;   // model.py:65-67
        } else {
            o = zz;  // model.py:68-70 (pure noise start)
        }
        out[i] = o;
    }
}

static int grid_for(int64_t n_items, int block, int max_blocks = 148 * 16) {
    int64_t g = (n_items + block - 1) / block;
    if (g < 1) g = 1;
    if (g > max_blocks) g = max_blocks;
    return (int)g;
}

int launch_x_T_coef(int variant, float a, float b, const float* cond, const float* z, uint64_t seed, int64_t row0,
                    float* x_out, int B, int L, cudaStream_t st) {
    const int L4 = L / 4;
    x_T_kernel<<<grid_for((int64_t)B * L4, 256), 256, 0, st>>>(variant, a, b, (const float4*)cond, (const float4*)z,
                                                              seed, row0, (float4*)x_out, B, L4);
    SDDM_LAUNCH_CHECK();
    return SDDM_OK;
}

// ===================================================================================================
// posterior update (+ fused overlap-add)                reference: model/diffusion.py:164-222,
//                                                                  model/UNetModified2.py:30-41
// ===================================================================================================
struct PostCoef {  // scalars of step t, fetched on the host from the plan's host tables
    float c2, sa, sig, gam, one_m_gam, cx, cy, ce;
};

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

__global__ void __launch_bounds__(256) post_kernel(PostP p, PostCoef k) {
    const int L4 = p.L / 4;
    const int64_t total = (int64_t)p.B * L4;
    const int K = (p.F + p.hop - 1) / p.hop;   // frames covering one sample
    const bool add_noise = p.t > 1;
    const bool need_cond = p.variant == SDDM_VAR_SUPPORTIVE || p.variant == SDDM_VAR_CONDITIONAL;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(i / L4), e4 = (int)(i - (int64_t)row * L4);
        float4 e;
        if (p.frames) {
            const int s = e4 * 4, a = s / p.hop, j = s - a * p.hop;
            e = make_float4(0.f, 0.f, 0.f, 0.f);
            const float* fr = p.frames + (int64_t)row * p.n_frames * p.F;
            for (int kk = K - 1; kk >= 0; --kk) {   // ascending frame index = the reference's accumulation order
                const int fa = a - kk, fj = j + kk * p.hop;
                if (fa >= 0 && fa < p.n_frames && fj < p.F) {
                    const float4 v = *reinterpret_cast<const float4*>(fr + (int64_t)fa * p.F + fj);
                    e.x = __fadd_rn(e.x, v.x); e.y = __fadd_rn(e.y, v.y); e.z = __fadd_rn(e.z, v.z); e.w = __fadd_rn(e.w, v.w);
                }
            }
        } else {
            e = reinterpret_cast<const float4*>(p.eps_in)[i];
        }
        if (p.eps_out) reinterpret_cast<float4*>(p.eps_out)[i] = e;
        if (!p.do_update) continue;
        const float4 x = reinterpret_cast<const float4*>(p.x_in)[i];
        float4 c = make_float4(0.f, 0.f, 0.f, 0.f), z = c;
        if (need_cond) c = reinterpret_cast<const float4*>(p.cond)[i];
        if (add_noise)
            z = p.z ? reinterpret_cast<const float4*>(p.z)[i]
                    : philox_normal4(p.seed, (uint32_t)e4, (uint64_t)(p.row0 + row), (uint32_t)(p.T + 1 - p.t));
        float4 o;
        o.x = post_one(p.variant, k, x.x, e.x, c.x, z.x, add_noise);
        o.y = post_one(p.variant, k, x.y, e.y, c.y, z.y, add_noise);
        o.z = post_one(p.variant, k, x.z, e.z, c.z, z.z, add_noise);
        o.w = post_one(p.variant, k, x.w, e.w, c.w, z.w, add_noise);
        reinterpret_cast<float4*>(p.x_out)[i] = o;
        if (p.x_trace) reinterpret_cast<float4*>(p.x_trace)[i] = o;
    }
}

int launch_post_coef(const PostP& p, const float* k8, cudaStream_t st) {
    PostCoef k{k8[0], k8[1], k8[2], k8[3], k8[4], k8[5], k8[6], k8[7]};
    post_kernel<<<grid_for((int64_t)p.B * (p.L / 4), 256), 256, 0, st>>>(p, k);
    SDDM_LAUNCH_CHECK();
    return SDDM_OK;
}

// ===================================================================================================
// framing helpers                                       reference: model/UNetModified2.py:5-41
// ===================================================================================================
__global__ void __launch_bounds__(256) frames_kernel(const float* __restrict__ sig, float* __restrict__ fr, int B, int n,
                                                     int F, int hop, int nf) {
    const int64_t total = (int64_t)B * nf * F;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(i % F);
        const int64_t r = i / F;
        const int a = (int)(r % nf), b = (int)(r / nf);
        fr[i] = sig[(int64_t)b * n + (int64_t)a * hop + j];
    }
}

__global__ void __launch_bounds__(256) overlap_add_kernel(const float* __restrict__ fr, float* __restrict__ sig, int B, int n,
                                                          int F, int hop, int nf) {
    const int64_t total = (int64_t)B * n;
    const int K = (F + hop - 1) / hop;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int s = (int)(i % n), b = (int)(i / n);
        const int a = s / hop, j = s - a * hop;
        float acc = 0.f;
        for (int kk = K - 1; kk >= 0; --kk) {
            const int fa = a - kk, fj = j + kk * hop;
            if (fa >= 0 && fa < nf && fj < F) acc = __fadd_rn(acc, fr[((int64_t)b * nf + fa) * F + fj]);
        }
        sig[i] = acc;
    }
}

int launch_frames(const float* sig, float* frames, int B, int n, int F, int hop, cudaStream_t st) {
    const int nf = (n - F) / hop + 1;
    frames_kernel<<<grid_for((int64_t)B * nf * F, 256), 256, 0, st>>>(sig, frames, B, n, F, hop, nf);
    SDDM_LAUNCH_CHECK();
    return SDDM_OK;
}

int launch_overlap_add(const float* frames, float* sig, int B, int n, int F, int hop, cudaStream_t st) {
    const int nf = (n - F) / hop + 1;
    overlap_add_kernel<<<grid_for((int64_t)B * n, 256), 256, 0, st>>>(frames, sig, B, n, F, hop, nf);
    SDDM_LAUNCH_CHECK();
    return SDDM_OK;
}

// ===================================================================================================
// noise-level embedding                                 reference: model/UNetModified2.py:49-89,168-174
// one block per row; accurate sinf/cosf (arguments reach 1e4: never the fast intrinsics)
// ===================================================================================================
__global__ void __launch_bounds__(128) temb_kernel(TembP p) {
    __shared__ float enc[64], h1[256], h2[64];
    const int row = blockIdx.x, tid = threadIdx.x, inner = p.inner, half = inner / 2;
    const float nl = p.nl[row];
    if (tid < half) {
        const float a = __fmul_rn(nl, p.freq[tid]);
        enc[tid] = sinf(a);
        enc[tid + half] = cosf(a);
    }
    __syncthreads();
    for (int j = tid; j < 4 * inner; j += blockDim.x) {
        float acc = p.b1[j];
        for (int i = 0; i < inner; ++i) acc = fmaf(p.w1[j * inner + i], enc[i], acc);
        h1[j] = swish_accurate(acc);
    }
    __syncthreads();
    for (int j = tid; j < inner; j += blockDim.x) {
        float acc = p.b2[j];
        for (int i = 0; i < 4 * inner; ++i) acc = fmaf(p.w2[j * 4 * inner + i], h1[i], acc);
        h2[j] = swish_accurate(acc);
    }
    __syncthreads();
    for (int j = tid; j < p.E; j += blockDim.x) {
        float acc = p.bn[j];
        for (int i = 0; i < inner; ++i) acc = fmaf(p.wn[j * inner + i], h2[i], acc);
        p.out[(int64_t)row * p.E + j] = acc;
    }
}

int launch_temb(const TembP& p, cudaStream_t st) {
    if (p.inner > 64 || (p.inner & 1)) { set_error("temb: inner_channel must be even and <= 64"); return SDDM_E_INVALID; }
    temb_kernel<<<p.rows, 128, 0, st>>>(p);
    SDDM_LAUNCH_CHECK();
    return SDDM_OK;
}

// ===================================================================================================
// stem: SignalToFrames x2 + cat + conv3x3(2 -> CO)      reference: UNetModified2.py:23-28,244-247,177-178
// tile = 16 frames x 8 positions, one pixel per thread, all CO outputs in registers.
// ===================================================================================================
int stem_nparts(int H, int W) { return (H / 16) * (W / 8) * 4; }

template <int CO>
__global__ void __launch_bounds__(128) stem_kernel(StemP p) {
    __shared__ __align__(16) float sw[18 * CO];
    __shared__ float sb[CO];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 18 * CO; i += 128) sw[i] = p.w[i];
    for (int i = tid; i < CO; i += 128) sb[i] = p.bias[i];
    __syncthreads();
    const int tiles_x = p.W / 8, tiles_y = p.H / 16;
    const int tile = blockIdx.x % (tiles_x * tiles_y), n = blockIdx.x / (tiles_x * tiles_y);
    const int y = (tile / tiles_x) * 16 + (tid >> 3), x = (tile % tiles_x) * 8 + (tid & 7);
    float in[2][9];
#pragma unroll
    for (int ci = 0; ci < 2; ++ci) {
        const float* sig = (ci == 0 ? p.cond : p.x_t) + (int64_t)n * p.L;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int iy = y + ky - 1, ix = x + kx - 1;
                in[ci][ky * 3 + kx] = (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) ? __ldg(sig + (int64_t)iy * p.hop + ix) : 0.f;
            }
    }
    float* op = p.out + (((int64_t)n * p.H + y) * p.W + x) * CO;
#pragma unroll
    for (int c0 = 0; c0 < CO; c0 += 32) {
        float acc[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = sb[c0 + c];
#pragma unroll
        for (int ci = 0; ci < 2; ++ci)
#pragma unroll
            for (int tp = 0; tp < 9; ++tp) {
                const float v = in[ci][tp];
                const float4* w4 = reinterpret_cast<const float4*>(sw + (ci * 9 + tp) * CO + c0);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 w = w4[q];
                    acc[4 * q + 0] = fmaf(v, w.x, acc[4 * q + 0]);
                    acc[4 * q + 1] = fmaf(v, w.y, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(v, w.z, acc[4 * q + 2]);
                    acc[4 * q + 3] = fmaf(v, w.w, acc[4 * q + 3]);
                }
            }
#pragma unroll
        for (int q = 0; q < 8; ++q)
            reinterpret_cast<float4*>(op + c0)[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
        float sq[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) sq[c] = acc[c] * acc[c];
        const float s1 = warp_transpose_reduce32(acc, lane);
        const float s2 = warp_transpose_reduce32(sq, lane);
        float* pp = p.parts + (((int64_t)n * p.nparts + tile * 4 + warp) * CO + c0 + lane) * 2;
        pp[0] = s1;
        pp[1] = s2;
    }
}

int launch_stem(const StemP& p, cudaStream_t st) {
    if (p.H % 16 || p.W % 8) { set_error("stem: frame grid %dx%d must be a multiple of 16x8", p.H, p.W); return SDDM_E_INVALID; }
    const int grid = p.B * (p.H / 16) * (p.W / 8);
    if (p.CO == 32) stem_kernel<32><<<grid, 128, 0, st>>>(p);
    else if (p.CO == 64) stem_kernel<64><<<grid, 128, 0, st>>>(p);
    else { set_error("stem: inner_channel %d unsupported (32 or 64)", p.CO); return SDDM_E_INVALID; }
    SDDM_LAUNCH_CHECK();
    return SDDM_OK;
}

// ===================================================================================================
// GroupNorm finalize                                    reference: nn.GroupNorm in UNetModified2.py:116-121
// parts hold per-(sample, tile-part, channel) [sum, sumsq]; combine in fp64 in a fixed order
// (deterministic), emit scale = gamma * rstd, shift = beta - mean * scale per (sample, channel).
// ===================================================================================================
__global__ void __launch_bounds__(64) gn_finalize_kernel(GnP p) {
    const int g = blockIdx.x, n = blockIdx.y, tid = threadIdx.x;
    const int cpg = p.Ctot / p.groups, c_lo = g * cpg;
    const int s = (c_lo < p.C[0]) ? 0 : 1;
    const int cl = c_lo - (s ? p.C[0] : 0);
    const int C = p.C[s], np = p.nparts[s];
    const float* base = p.parts[s] + (int64_t)n * np * C * 2;
    double sum = 0.0, sq = 0.0;
    for (int i = tid; i < np * cpg; i += 64) {
        const int part = i / cpg, j = i - part * cpg;
        const float2 v = *reinterpret_cast<const float2*>(base + ((int64_t)part * C + cl + j) * 2);
        sum += (double)v.x;
        sq += (double)v.y;
    }
    __shared__ double sh[2][2];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    if ((tid & 31) == 0) { sh[tid >> 5][0] = sum; sh[tid >> 5][1] = sq; }
    __syncthreads();
    sum = sh[0][0] + sh[1][0];
    sq = sh[0][1] + sh[1][1];
    const double cnt = (double)cpg * (double)p.HW;
    const double mean = sum / cnt;
    double var = sq / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const double rstd = 1.0 / sqrt(var + (double)p.eps);
    for (int j = tid; j < cpg; j += 64) {
        const int c = c_lo + j;
        const double sc = (double)p.gamma[c] * rstd;
        p.scale[(int64_t)n * p.Ctot + c] = (float)sc;
        p.shift[(int64_t)n * p.Ctot + c] = (float)((double)p.beta[c] - mean * sc);
    }
}

int launch_gn_finalize(const GnP& p, cudaStream_t st) {
    const int cpg = p.Ctot / p.groups;
    if (p.Ctot % p.groups || (p.nsrc == 2 && p.C[0] % cpg)) { set_error("GroupNorm: groups straddle the concat boundary"); return SDDM_E_INVALID; }
    gn_finalize_kernel<<<dim3(p.groups, p.B), 64, 0, st>>>(p);
    SDDM_LAUNCH_CHECK();
    return SDDM_OK;
}

// ===================================================================================================
// final Block: GN-apply + Swish + conv3x3(C -> 1)       reference: UNetModified2.py:235,267 (Block, :113-124)
// tile = 16 x 8 pixels, one pixel per thread; halo tile staged in smem post-activation (zero padded).
// ===================================================================================================
__global__ void __launch_bounds__(128) final_conv_kernel(FinalP p) {
    extern __shared__ __align__(16) float smem[];
    const int C = p.C, CP = C + 4;
    float* sa = smem;                 // [18*10][CP]
    float* sw = smem + 180 * CP;      // [9][C]
    const int tid = threadIdx.x;
    const int tiles_x = p.W / 8, tiles_y = p.H / 16;
    const int tile = blockIdx.x % (tiles_x * tiles_y), n = blockIdx.x / (tiles_x * tiles_y);
    const int y0 = (tile / tiles_x) * 16, x0 = (tile % tiles_x) * 8;
    for (int i = tid; i < 9 * C; i += 128) sw[i] = p.w[i];
    const int c4n = C / 4;
    const float4* sc4 = reinterpret_cast<const float4*>(p.scale + (int64_t)n * C);
    const float4* sh4 = reinterpret_cast<const float4*>(p.shift + (int64_t)n * C);
    for (int i = tid; i < 180 * c4n; i += 128) {
        const int pix = i / c4n, c4 = i - pix * c4n;
        const int hy = pix / 10, hx = pix - hy * 10;
        const int iy = y0 + hy - 1, ix = x0 + hx - 1;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(p.x + (((int64_t)n * p.H + iy) * p.W + ix) * C) + c4);
            const float4 a = __ldg(sc4 + c4), b = __ldg(sh4 + c4);
            v.x = swish_accurate(fmaf(r.x, a.x, b.x));
            v.y = swish_accurate(fmaf(r.y, a.y, b.y));
            v.z = swish_accurate(fmaf(r.z, a.z, b.z));
            v.w = swish_accurate(fmaf(r.w, a.w, b.w));
        }
        *reinterpret_cast<float4*>(sa + pix * CP + c4 * 4) = v;
    }
    __syncthreads();
    const int ty = tid >> 3, tx = tid & 7;
    float acc = p.bias;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const float4* a4 = reinterpret_cast<const float4*>(sa + ((ty + ky) * 10 + tx + kx) * CP);
            const float4* w4 = reinterpret_cast<const float4*>(sw + (ky * 3 + kx) * C);
            for (int c4 = 0; c4 < c4n; ++c4) {
                const float4 a = a4[c4], w = w4[c4];
                acc = fmaf(a.x, w.x, acc);
                acc = fmaf(a.y, w.y, acc);
                acc = fmaf(a.z, w.z, acc);
                acc = fmaf(a.w, w.w, acc);
            }
        }
    p.frames[((int64_t)n * p.H + y0 + ty) * p.W + x0 + tx] = acc;
}

int launch_final_conv(const FinalP& p, cudaStream_t st) {
    if (p.H % 16 || p.W % 8 || p.C % 4) { set_error("final conv: unsupported shape"); return SDDM_E_INVALID; }
    const size_t smem = (size_t)(180 * (p.C + 4) + 9 * p.C) * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(final_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("final conv: smem %zu too large", smem); return SDDM_E_CUDA; }
    }
    final_conv_kernel<<<p.B * (p.H / 16) * (p.W / 8), 128, smem, st>>>(p);
    SDDM_LAUNCH_CHECK();
    return SDDM_OK;
}

}  // namespace sddm
