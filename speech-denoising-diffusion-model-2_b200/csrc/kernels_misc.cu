// HBM-bound kernels of the reverse-diffusion loop: x_T initialisation, posterior update fused with
// overlap-add, framing helpers, noise-level embedding, stem (framing + cat + conv 2->C), GroupNorm
// finalize and the final Block (GN + Swish + conv C->1).  All fp32, all coalesced / vectorised.
#include "kernels.cuh"
#include "gn_fuse.cuh"
#include "../../include/sddm_b200.h"

namespace sddm {

// ===================================================================================================
// x_T                                                   reference: model/diffusion.py:281-320
// ===================================================================================================
__global__ void __launch_bounds__(256) x_T_kernel(int variant, float a, float b, const float4* __restrict__ cond,
                                                  const float4* __restrict__ z, uint64_t seed, int64_t row0,
                                                  float4* __restrict__ out, int B, int L4, const unsigned long long* __restrict__ seed_dev) {
    if (seed_dev) { seed = seed_dev[0]; row0 = (int64_t)seed_dev[1]; }
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
                    float* x_out, int B, int L, cudaStream_t st, const unsigned long long* seed_dev) {
    const int L4 = L / 4;
    x_T_kernel<<<grid_for((int64_t)B * L4, 256), 256, 0, st>>>(variant, a, b, (const float4*)cond, (const float4*)z,
                                                              seed, row0, (float4*)x_out, B, L4, seed_dev);
    SDDM_LAUNCH_CHECK();
    return SDDM_OK;
}

// ===================================================================================================
// posterior update (+ fused overlap-add)                reference: model/diffusion.py:164-222,
//                                                                  model/UNetModified2.py:30-41
// ===================================================================================================
__global__ void __launch_bounds__(256) post_kernel(PostP p, PostCoef k) {
    if (p.seed_dev) { p.seed = p.seed_dev[0]; p.row0 = (int64_t)p.seed_dev[1]; }
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
// forward diffusion q(x_t | x_0)                        reference: model/diffusion.py:225-279
//   mode 0 (q_stochastic):             x_t = a x_0 + b z                     coef[row] = {a, b, -, -}
//   mode 1 (q_stochastic_conditional): g = sd z; c = (m sab) (y - x_0); x_t = (sab x_0 + c) + g; combined = k (c + g)
//                                                                             coef[row] = {sab, m sab, sd, k = 1 / sqrt(1 - alpha_bar)}
// every product / sum rounded separately, as the reference's eager ops do
// ===================================================================================================
__global__ void __launch_bounds__(256) q_sample_kernel(int mode, const float4* __restrict__ coef, const float4* __restrict__ x0,
                                                       const float4* __restrict__ y, const float4* __restrict__ z, uint64_t seed, int64_t row0,
                                                       float4* __restrict__ x_t, float4* __restrict__ combined, float4* __restrict__ z_out, int B, int L4) {
    const int64_t total = (int64_t)B * L4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(i / L4), e4 = (int)(i - (int64_t)row * L4);
        const float4 k = __ldg(coef + row), a = x0[i];
        const float4 n = z ? z[i] : philox_normal4(seed, (uint32_t)e4, (uint64_t)(row0 + row), 0u);
        if (z_out) z_out[i] = n;
        float4 o;
        if (mode == 0) {
            o.x = __fadd_rn(__fmul_rn(k.x, a.x), __fmul_rn(k.y, n.x));
            o.y = __fadd_rn(__fmul_rn(k.x, a.y), __fmul_rn(k.y, n.y));
            o.z = __fadd_rn(__fmul_rn(k.x, a.z), __fmul_rn(k.y, n.z));
            o.w = __fadd_rn(__fmul_rn(k.x, a.w), __fmul_rn(k.y, n.w));
        } else {
            const float4 c = y[i];
            const float av[4] = {a.x, a.y, a.z, a.w}, cv[4] = {c.x, c.y, c.z, c.w}, nv[4] = {n.x, n.y, n.z, n.w};
            float ov[4], mv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float g = __fmul_rn(k.z, nv[q]);
                const float nc = __fmul_rn(k.y, __fsub_rn(cv[q], av[q]));
                ov[q] = __fadd_rn(__fadd_rn(__fmul_rn(k.x, av[q]), nc), g);
                mv[q] = __fmul_rn(k.w, __fadd_rn(nc, g));
            }
            o = make_float4(ov[0], ov[1], ov[2], ov[3]);
            combined[i] = make_float4(mv[0], mv[1], mv[2], mv[3]);
        }
        x_t[i] = o;
    }
}

int launch_q_sample(int mode, const float* coef, const float* x0, const float* y, const float* z, uint64_t seed, int64_t row0, float* x_t,
                    float* combined, float* z_out, int B, int L, cudaStream_t st) {
    q_sample_kernel<<<grid_for((int64_t)B * (L / 4), 256), 256, 0, st>>>(mode, (const float4*)coef, (const float4*)x0, (const float4*)y, (const float4*)z,
                                                                       seed, row0, (float4*)x_t, (float4*)combined, (float4*)z_out, B, L / 4);
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
// HBM-write bound (2 x 64 KB in, CO x 128 KB out per chunk).  Tile = 16 frames x 8 positions; a thread owns 4 output
// channels of one pixel per pass, so the CO/4 threads of a pixel write its CO*4 contiguous bytes and a warp writes whole
// 128-byte lines.  The 18 x 10 x 2 input window is staged in shared memory; weights stay in registers across the
// persistent CTA's tiles.  GroupNorm partial statistics: one (sum, sumsq) per channel per tile, fixed summation order.
// ===================================================================================================
int stem_nparts(int H, int W) { return (H / 16) * (W / 8); }

template <int CO, bool A16>
__global__ void __launch_bounds__(256) stem_kernel(StemP p) {
    constexpr int CG = CO / 4;            // channel groups (threads) per pixel
    constexpr int PPP = 256 / CG;         // pixels per pass
    constexpr int PASSES = 128 / PPP;
    __shared__ float sin_[2][18][10];
    __shared__ float red[8][CO][2];
    __shared__ int gn_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cg = tid % CG, pl = tid / CG;
    float4 w[18];
#pragma unroll
    for (int k = 0; k < 18; ++k) w[k] = __ldg(reinterpret_cast<const float4*>(p.w + (size_t)k * CO) + cg);
    const float4 bias = __ldg(reinterpret_cast<const float4*>(p.bias) + cg);
    const int tiles_x = p.W / 8, tiles_y = p.H / 16, per = tiles_x * tiles_y, ntiles = p.B * per;
    // input window element(s) of this thread: i = tid and tid + 256 of the 2 x 18 x 10 window; the NEXT tile's values are
    // fetched into registers while the current tile is computed (hides the global latency of the persistent loop)
    auto fetch = [&](int tl, int i) -> float {
        if (tl >= ntiles || i >= 360) return 0.f;
        const int n = tl / per, tile = tl - n * per;
        const int y0 = (tile / tiles_x) * 16, x0 = (tile % tiles_x) * 8;
        const int ci = i / 180, r = i - ci * 180, hy = r / 10, hx = r - hy * 10;
        const int iy = y0 + hy - 1, ix = x0 + hx - 1;
        const float* sig = (ci == 0 ? p.cond : p.x_t) + (int64_t)n * p.L;
        return (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) ? __ldg(sig + (int64_t)iy * p.hop + ix) : 0.f;
    };
    float nx0 = fetch(blockIdx.x, tid), nx1 = fetch(blockIdx.x, tid + 256);
    for (int tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
        const int n = tl / per, tile = tl - n * per;
        const int y0 = (tile / tiles_x) * 16, x0 = (tile % tiles_x) * 8;
        __syncthreads();   // previous tile's readers of sin_ / red are done
        (&sin_[0][0][0])[tid] = nx0;
        if (tid + 256 < 360) (&sin_[0][0][0])[tid + 256] = nx1;
        __syncthreads();
        nx0 = fetch(tl + gridDim.x, tid);
        nx1 = fetch(tl + gridDim.x, tid + 256);
        float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ps = 0; ps < PASSES; ++ps) {
            const int pix = ps * PPP + pl, py = pix >> 3, px = pix & 7;
            float4 acc = bias;
#pragma unroll
            for (int ci = 0; ci < 2; ++ci)
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const float v = sin_[ci][py + ky][px + kx];
                        const float4 ww = w[ci * 9 + ky * 3 + kx];
                        acc.x = fmaf(v, ww.x, acc.x); acc.y = fmaf(v, ww.y, acc.y);
                        acc.z = fmaf(v, ww.z, acc.z); acc.w = fmaf(v, ww.w, acc.w);
                    }
            const int64_t oidx = (((int64_t)n * p.H + y0 + py) * p.W + x0 + px) * CO + cg * 4;
            if (A16) {   // bf16 storage: the statistics are those of the stored (rounded) values
                const __nv_bfloat162 lo = __floats2bfloat162_rn(acc.x, acc.y), hi = __floats2bfloat162_rn(acc.z, acc.w);
                uint4 pk;   // neighbouring channel groups pair up: the even one writes 16 bytes (8 channels)
                pk.x = *reinterpret_cast<const uint32_t*>(&lo);
                pk.y = *reinterpret_cast<const uint32_t*>(&hi);
                pk.z = __shfl_down_sync(0xffffffffu, pk.x, 1);
                pk.w = __shfl_down_sync(0xffffffffu, pk.y, 1);
                if ((cg & 1) == 0) *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + oidx) = pk;
                const float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
                acc = make_float4(a.x, a.y, b.x, b.y);
            } else {
                *reinterpret_cast<float4*>(p.out + oidx) = acc;
            }
            s1[0] += acc.x; s1[1] += acc.y; s1[2] += acc.z; s1[3] += acc.w;
            s2[0] = fmaf(acc.x, acc.x, s2[0]); s2[1] = fmaf(acc.y, acc.y, s2[1]);
            s2[2] = fmaf(acc.z, acc.z, s2[2]); s2[3] = fmaf(acc.w, acc.w, s2[3]);
        }
#pragma unroll
        for (int o = CG; o < 32; o <<= 1)   // lanes with the same channel group: fixed butterfly order
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], o);
                s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], o);
            }
        if (lane < CG) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { red[warp][lane * 4 + k][0] = s1[k]; red[warp][lane * 4 + k][1] = s2[k]; }
        }
        __syncthreads();
        if (tid < CO) {
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int wq = 0; wq < 8; ++wq) { a += red[wq][tid][0]; b += red[wq][tid][1]; }
            *reinterpret_cast<float2*>(p.parts + (((int64_t)n * p.nparts + tile) * CO + tid) * 2) = make_float2(a, b);
        }
        if (p.gn_on) {   // publish the tile's partials; the last tile of sample n finalises the consumer's GroupNorm
            __threadfence();
            __syncthreads();
            if (tid == 0) gn_last = atomicAdd(p.gn.counter + n, 1u) == (unsigned)(p.gn.expect - 1);
            __syncthreads();
            if (gn_last) {
                __threadfence();
                if (tid < 128) gn_fused_finalize(p.gn, n, tid, 128);
                if (tid == 0) p.gn.counter[n] = 0u;
            }
        }
    }
}

int launch_stem(const StemP& p, cudaStream_t st) {
    if (p.H % 16 || p.W % 8) { set_error("stem: frame grid %dx%d must be a multiple of 16x8", p.H, p.W); return SDDM_E_INVALID; }
    if (p.nparts != stem_nparts(p.H, p.W)) { set_error("stem: nparts mismatch"); return SDDM_E_INVALID; }
    const int ntiles = p.B * (p.H / 16) * (p.W / 8);
    const int grid = ntiles < 148 * 8 ? ntiles : 148 * 8;
    if (p.CO == 32) { if (p.act16) stem_kernel<32, true><<<grid, 256, 0, st>>>(p); else stem_kernel<32, false><<<grid, 256, 0, st>>>(p); }
    else if (p.CO == 64) { if (p.act16) stem_kernel<64, true><<<grid, 256, 0, st>>>(p); else stem_kernel<64, false><<<grid, 256, 0, st>>>(p); }
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
    pdl_launch_dependents();   // the consumer convolution may set up (barriers, TMEM, weights) while this grid runs
    pdl_wait();                // partial statistics come from the preceding kernel
    const int g = blockIdx.x, n = blockIdx.y, tid = threadIdx.x;
    const int cpg = p.Ctot / p.groups, c_lo = g * cpg;
    // affine parameters first: their latency overlaps the partial-sum loads instead of following the reduction
    float gam = 0.f, bet = 0.f;
    if (tid < cpg) { gam = __ldg(p.gamma + c_lo + tid); bet = __ldg(p.beta + c_lo + tid); }
    double sum = 0.0, sq = 0.0;
    // a group may straddle the concatenation boundary (e.g. cat(64, 32) channels in 32 groups of 3): take from every
    // source the channels of [c_lo, c_lo + cpg) it holds
    int off = 0;
    for (int s = 0; s < p.nsrc; ++s) {
        const int C = p.C[s], np = p.nparts[s];
        const int lo = c_lo > off ? c_lo : off, hi = (c_lo + cpg) < (off + C) ? (c_lo + cpg) : (off + C);
        const int w = hi - lo;   // channels of this group inside source s
        if (w > 0) {
            const float* base = p.parts[s] + (int64_t)n * np * C * 2;
            const int cl = lo - off;
#pragma unroll 4
            for (int i = tid; i < np * w; i += 64) {
                const int part = i / w, j = i - part * w;
                const float2 v = __ldg(reinterpret_cast<const float2*>(base + ((int64_t)part * C + cl + j) * 2));
                sum += (double)v.x;
                sq += (double)v.y;
            }
        }
        off += C;
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
    if (tid < cpg) {
        const int c = c_lo + tid;
        const double sc = (double)gam * rstd;
        p.scale[(int64_t)n * p.Ctot + c] = (float)sc;
        p.shift[(int64_t)n * p.Ctot + c] = (float)((double)bet - mean * sc);
    }
}

int launch_gn_finalize(const GnP& p, cudaStream_t st) {
    const int cpg = p.Ctot / p.groups;
    if (p.Ctot % p.groups) { set_error("GroupNorm: %d channels do not divide into %d groups", p.Ctot, p.groups); return SDDM_E_INVALID; }
    if (cpg > 64) { set_error("GroupNorm: more than 64 channels per group"); return SDDM_E_INVALID; }
    SDDM_CUDA_TRY(launch_pdl(gn_finalize_kernel, dim3(p.groups, p.B), dim3(64), 0, st, p));
    SDDM_LAUNCH_CHECK();
    return SDDM_OK;
}

// ===================================================================================================
// final Block: GN-apply + Swish + conv3x3(C -> 1)       reference: UNetModified2.py:235,267 (Block, :113-124)
// HBM-read bound.  Tile = 16 x 16 pixels; the post-activation, zero-padded 18 x 18 halo is staged in shared memory with
// coalesced 128-bit loads (the C/4 threads of a pixel read its C*4 contiguous bytes).  A thread owns one pixel column and
// 4 channels: it walks the 18 halo rows once (3 x LDS.128 per row) and feeds 16 row accumulators, so every staged value is
// read 3 times instead of 9; the C/4 partial dot products of a pixel are combined with a shuffle butterfly.
// ===================================================================================================
template <int C, bool FAST, bool A16>
__global__ void __launch_bounds__(4 * C) final_conv_kernel(FinalP p) {
    constexpr int CG = C / 4;
    extern __shared__ __align__(16) float smem[];   // [18*18][C]
    const int tid = threadIdx.x, cg = tid % CG, col = tid / CG;
    float4 w[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) w[k] = __ldg(reinterpret_cast<const float4*>(p.w + (size_t)k * C) + cg);
    const int tiles_x = p.W / 16, tiles_y = p.H / 16, per = tiles_x * tiles_y, ntiles = p.B * per;
    for (int tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
        const int n = tl / per, tile = tl - n * per;
        const int y0 = (tile / tiles_x) * 16, x0 = (tile % tiles_x) * 16;
        const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + (int64_t)n * C) + cg);
        const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift + (int64_t)n * C) + cg);
        __syncthreads();   // previous tile's readers are done
        if (A16) {
            // bf16 storage: a thread stages 8 channels of a pixel with one 16-byte load (C/8 threads per pixel)
            constexpr int CH8 = C / 8, PPR = (4 * C) / CH8, U = 4;   // 8-channel chunks per pixel, pixels per round, loads in flight
            const int ch = tid % CH8, pl = tid / CH8;
            const float4 sc0 = __ldg(reinterpret_cast<const float4*>(p.scale + (int64_t)n * C) + 2 * ch), sc1 = __ldg(reinterpret_cast<const float4*>(p.scale + (int64_t)n * C) + 2 * ch + 1);
            const float4 sh0 = __ldg(reinterpret_cast<const float4*>(p.shift + (int64_t)n * C) + 2 * ch), sh1 = __ldg(reinterpret_cast<const float4*>(p.shift + (int64_t)n * C) + 2 * ch + 1);
#pragma unroll 1
            for (int base = pl; base < 18 * 18; base += PPR * U) {
                uint4 rv[U];
                bool okv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int pix = base + PPR * u;
                    const int hy = pix / 18, hx = pix - hy * 18;
                    const int iy = y0 + hy - 1, ix = x0 + hx - 1;
                    okv[u] = pix < 18 * 18 && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
                    if (okv[u]) rv[u] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.x) + (((int64_t)n * p.H + iy) * p.W + ix) * C) + ch);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int pix = base + PPR * u;
                    if (pix >= 18 * 18) break;
                    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                    if (okv[u]) {
                        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rv[u].x)), b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rv[u].y));
                        const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rv[u].z)), d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rv[u].w));
                        v0.x = swish_fast(fmaf(a.x, sc0.x, sh0.x)); v0.y = swish_fast(fmaf(a.y, sc0.y, sh0.y));
                        v0.z = swish_fast(fmaf(b.x, sc0.z, sh0.z)); v0.w = swish_fast(fmaf(b.y, sc0.w, sh0.w));
                        v1.x = swish_fast(fmaf(c.x, sc1.x, sh1.x)); v1.y = swish_fast(fmaf(c.y, sc1.y, sh1.y));
                        v1.z = swish_fast(fmaf(d.x, sc1.z, sh1.z)); v1.w = swish_fast(fmaf(d.y, sc1.w, sh1.w));
                    }
                    float4* dst = reinterpret_cast<float4*>(smem + (size_t)pix * C + ch * 8);
                    dst[0] = v0; dst[1] = v1;
                }
            }
        } else {
        constexpr int U = 7;   // 128-bit loads in flight per thread
#pragma unroll 1
        for (int base = col; base < 18 * 18; base += 16 * U) {
            float4 rv[U];
            bool okv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int pix = base + 16 * u;
                const int hy = pix / 18, hx = pix - hy * 18;
                const int iy = y0 + hy - 1, ix = x0 + hx - 1;
                okv[u] = pix < 18 * 18 && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
                if (okv[u]) {
                    const int64_t idx = (((int64_t)n * p.H + iy) * p.W + ix) * C + cg * 4;
                    if (A16) {
                        const uint2 pk = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.x) + idx));
                        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.x));
                        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.y));
                        rv[u] = make_float4(a.x, a.y, b.x, b.y);
                    } else {
                        rv[u] = __ldg(reinterpret_cast<const float4*>(p.x + idx));
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int pix = base + 16 * u;
                if (pix >= 18 * 18) break;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (okv[u]) {
                    const float4 r = rv[u];
                    if (FAST) {
                        v.x = swish_fast(fmaf(r.x, sc.x, sh.x)); v.y = swish_fast(fmaf(r.y, sc.y, sh.y));
                        v.z = swish_fast(fmaf(r.z, sc.z, sh.z)); v.w = swish_fast(fmaf(r.w, sc.w, sh.w));
                    } else {
                        v.x = swish_accurate(fmaf(r.x, sc.x, sh.x)); v.y = swish_accurate(fmaf(r.y, sc.y, sh.y));
                        v.z = swish_accurate(fmaf(r.z, sc.z, sh.z)); v.w = swish_accurate(fmaf(r.w, sc.w, sh.w));
                    }
                }
                *reinterpret_cast<float4*>(smem + (size_t)pix * C + cg * 4) = v;
            }
        }
        }
        __syncthreads();
        float acc[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) acc[r] = 0.f;
#pragma unroll
        for (int hr = 0; hr < 18; ++hr) {
            float4 a[3];
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) a[kx] = *reinterpret_cast<const float4*>(smem + (size_t)(hr * 18 + col + kx) * C + cg * 4);
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int r = hr - ky;
                if (r < 0 || r >= 16) continue;
                float t = acc[r];
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float4 ww = w[ky * 3 + kx];
                    t = fmaf(a[kx].x, ww.x, t); t = fmaf(a[kx].y, ww.y, t); t = fmaf(a[kx].z, ww.z, t); t = fmaf(a[kx].w, ww.w, t);
                }
                acc[r] = t;
            }
        }
#pragma unroll
        for (int o = 1; o < CG; o <<= 1)
#pragma unroll
            for (int r = 0; r < 16; ++r) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
#pragma unroll
        for (int r = 0; r < 16; ++r)
            if ((r % CG) == cg) p.frames[((int64_t)n * p.H + y0 + r) * p.W + x0 + col] = acc[r] + p.bias;
    }
}

int launch_final_conv(const FinalP& p, cudaStream_t st) {
    if (p.H % 16 || p.W % 16 || (p.C != 32 && p.C != 64)) { set_error("final conv: unsupported shape %dx%dx%d", p.H, p.W, p.C); return SDDM_E_INVALID; }
    const size_t smem = (size_t)18 * 18 * p.C * sizeof(float);
    const int ntiles = p.B * (p.H / 16) * (p.W / 16);
    const int grid = ntiles < 148 * 4 ? ntiles : 148 * 4;
#define SDDM_FINAL_LAUNCH(CC, FF, AA)                                                                                      \
    do {                                                                                                                  \
        SDDM_SET_MAX_SMEM((final_conv_kernel<CC, FF, AA>), 18 * 18 * CC * 4);                                           \
        final_conv_kernel<CC, FF, AA><<<grid, 4 * CC, smem, st>>>(p);                                                     \
    } while (0)
    if (p.act16 && !p.fast_math) { set_error("final conv: bf16 activations come with the tcgen05 path"); return SDDM_E_INVALID; }
    if (p.C == 32) { if (p.act16) SDDM_FINAL_LAUNCH(32, true, true); else if (p.fast_math) SDDM_FINAL_LAUNCH(32, true, false); else SDDM_FINAL_LAUNCH(32, false, false); }
    else { if (p.act16) SDDM_FINAL_LAUNCH(64, true, true); else if (p.fast_math) SDDM_FINAL_LAUNCH(64, true, false); else SDDM_FINAL_LAUNCH(64, false, false); }
#undef SDDM_FINAL_LAUNCH
    SDDM_LAUNCH_CHECK();
    return SDDM_OK;
}

// ===================================================================================================
// dataset edge on the device                            reference: InferDataset.__getitem__ + infer_data_collate,
//                                                       data_loader/data_loaders.py:101-155; regroup loop infer.py:81-120
// The utterances of a batch sit back to back in `flat` (sample_off[u] .. sample_off[u + 1]); utterance u owns the rows
// row_off[u] .. row_off[u + 1] of the [N, 1, T] batch (ceil(len / T) rows, the last one zero padded).
//   chunk  : rows[r - row_lo][:] for r in [row_lo, row_hi)  <- flat            (pad + view + cat of the reference)
//   regroup: flat_out[sample_off[u] + i] <- rows[...]  for i < len(u)          (reshape(1, -1) per file, trimmed to the input length)
// One block per row; the owning utterance is found by bisection of row_off.
// ===================================================================================================
__device__ __forceinline__ int owner_of_row(const int64_t* __restrict__ row_off, int n_utt, int64_t r) {
    int lo = 0, hi = n_utt - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(row_off + mid) <= r) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(256) chunk_rows_kernel(const float* __restrict__ flat, const int64_t* __restrict__ sample_off,
                                                         const int64_t* __restrict__ row_off, int n_utt, int T, int64_t row_lo,
                                                         float* __restrict__ rows) {
    const int64_t r = row_lo + blockIdx.x;
    const int u = owner_of_row(row_off, n_utt, r);
    const int64_t s0 = __ldg(sample_off + u), len = __ldg(sample_off + u + 1) - s0;
    const int64_t first = (r - __ldg(row_off + u)) * T;          // first sample of this row inside its utterance
    float* dst = rows + (int64_t)blockIdx.x * T;
    for (int i = threadIdx.x; i < T; i += blockDim.x) dst[i] = (first + i < len) ? __ldg(flat + s0 + first + i) : 0.f;
}

__global__ void __launch_bounds__(256) regroup_rows_kernel(const float* __restrict__ rows, const int64_t* __restrict__ sample_off,
                                                           const int64_t* __restrict__ row_off, int n_utt, int T, int64_t row_lo,
                                                           float* __restrict__ flat_out) {
    const int64_t r = row_lo + blockIdx.x;
    const int u = owner_of_row(row_off, n_utt, r);
    const int64_t s0 = __ldg(sample_off + u), len = __ldg(sample_off + u + 1) - s0;
    const int64_t first = (r - __ldg(row_off + u)) * T;
    const float* src = rows + (int64_t)blockIdx.x * T;
    for (int i = threadIdx.x; i < T; i += blockDim.x)
        if (first + i < len) flat_out[s0 + first + i] = src[i];
}

int launch_chunk_rows(const float* flat, const int64_t* sample_off, const int64_t* row_off, int n_utt, int T, int64_t row_lo, int64_t row_hi,
                      float* rows, cudaStream_t st) {
    if (row_hi <= row_lo) return SDDM_OK;
    chunk_rows_kernel<<<(unsigned)(row_hi - row_lo), 256, 0, st>>>(flat, sample_off, row_off, n_utt, T, row_lo, rows);
    SDDM_LAUNCH_CHECK();
    return SDDM_OK;
}

int launch_regroup_rows(const float* rows, const int64_t* sample_off, const int64_t* row_off, int n_utt, int T, int64_t row_lo, int64_t row_hi,
                        float* flat_out, cudaStream_t st) {
    if (row_hi <= row_lo) return SDDM_OK;
    regroup_rows_kernel<<<(unsigned)(row_hi - row_lo), 256, 0, st>>>(rows, sample_off, row_off, n_utt, T, row_lo, flat_out);
    SDDM_LAUNCH_CHECK();
    return SDDM_OK;
}

__global__ void __launch_bounds__(256) bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = __bfloat162float(src[i]);
}

int launch_bf16_to_f32(const void* src, float* dst, size_t n, cudaStream_t st) {
    bf16_to_f32_kernel<<<grid_for((int64_t)n, 256), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(src), dst, n);
    SDDM_LAUNCH_CHECK();
    return SDDM_OK;
}

}  // namespace sddm
