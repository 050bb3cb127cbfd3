// CUDA-core fp32 implicit-GEMM 3x3 convolution with the UNet's fusions — the parity ("fp32") mode of
// the denoiser and the on-device yardstick for the tcgen05 path.
//
//   prologue : GroupNorm-apply + Swish on the operand while staging the halo tile (zero padding applied
//              post-activation), two-source channel concat, nearest x2 upsample addressing, stride 2
//   main     : 3x3 taps x Cin in chunks of 8 channels, 4 pixels x 4 couts per thread
//   extra K  : the ResnetBlock's 1x1 res_conv over the raw block input accumulates into the same tile
//   epilogue : + bias (+ noise-level embedding) (+ identity residual), NHWC store, GroupNorm partial
//              statistics (sum, sum of squares per channel per tile), fixed-order => deterministic
//
// reference: Block / ResnetBlock / Downsample / Upsample, model/UNetModified2.py:93-142
#include "common.cuh"
#include "../../include/sddm_b200.h"

namespace sddm {

constexpr int CK = 8;    // input channels per smem stage
constexpr int NT = 32;   // output channels per CTA

template <int TH, int TW, int MODE>
__global__ void __launch_bounds__(TH* TW * 2) conv3x3_fp32_kernel(ConvP p) {
    constexpr int S = (MODE == CONV_S2) ? 2 : 1;
    constexpr int HH = S * (TH - 1) + 3, HW = S * (TW - 1) + 3;
    constexpr int NPG = TH * TW / 4;      // pixel groups (4 consecutive x)
    constexpr int NTHREADS = NPG * 8;
    __shared__ __align__(16) float sA[CK * HH * HW];
    __shared__ __align__(16) float sW[9 * CK * NT];
    __shared__ float red[2][NPG][NT];

    const int tid = threadIdx.x;
    const int tiles_x = p.Wout / TW, tiles_y = p.Hout / TH, tiles = tiles_x * tiles_y;
    const int n = blockIdx.x / tiles, tile = blockIdx.x - n * tiles;
    const int oy0 = (tile / tiles_x) * TH, ox0 = (tile % tiles_x) * TW;
    const int co0 = blockIdx.y * NT;
    const int pg = tid >> 3, cg = tid & 7;
    const int py = pg / (TW / 4), px0 = (pg % (TW / 4)) * 4;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int main_chunks = p.Cin / CK;
    const int res_chunks = p.res_w ? p.res_Cin / CK : 0;
    for (int ch = 0; ch < main_chunks + res_chunks; ++ch) {
        const bool is_res = ch >= main_chunks;
        const int cbase = (is_res ? ch - main_chunks : ch) * CK;   // concatenated channel index of this chunk
        const ConvSrc* srcs = is_res ? p.res_src : p.src;
        const int s = (cbase < srcs[0].C) ? 0 : 1;
        const ConvSrc src = srcs[s];
        const int coff = cbase - (s ? srcs[0].C : 0);
        const int ctot = is_res ? p.res_Cin : p.Cin;
        __syncthreads();
        // ---- stage the (post-activation, zero-padded) halo tile --------------------------------
        for (int idx = tid; idx < HH * HW * CK; idx += NTHREADS) {
            const int c = idx % CK, pix = idx / CK;
            const int hy = pix / HW, hx = pix - hy * HW;
            int iy, ix;
            bool ok;
            if (MODE == CONV_UP) {
                const int uy = oy0 + hy - 1, ux = ox0 + hx - 1;
                ok = uy >= 0 && uy < p.Hout && ux >= 0 && ux < p.Wout;
                iy = uy >> 1;
                ix = ux >> 1;
            } else {
                iy = S * oy0 + hy - 1;
                ix = S * ox0 + hx - 1;
                ok = iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win;
            }
            float v = 0.f;
            if (ok) {
                v = __ldg(src.x + (((int64_t)n * p.Hin + iy) * p.Win + ix) * src.C + coff + c);
                if (src.scale) {
                    const int64_t gi = (int64_t)n * ctot + cbase + c;
                    v = swish_accurate(fmaf(v, __ldg(src.scale + gi), __ldg(src.shift + gi)));
                }
            }
            sA[c * HH * HW + pix] = v;
        }
        // ---- stage the weights of this chunk for this cout tile ---------------------------------
        if (!is_res) {
            const float* wsrc = p.w + (int64_t)ch * 9 * CK * p.Cout + co0;
            for (int idx = tid; idx < 9 * CK * NT; idx += NTHREADS) sW[idx] = __ldg(wsrc + (int64_t)(idx / NT) * p.Cout + (idx % NT));
        } else {
            const float* wsrc = p.res_w + (int64_t)(ch - main_chunks) * CK * p.Cout + co0;
            for (int idx = tid; idx < CK * NT; idx += NTHREADS) sW[idx] = __ldg(wsrc + (int64_t)(idx / NT) * p.Cout + (idx % NT));
        }
        __syncthreads();
        // ---- accumulate -------------------------------------------------------------------------
        if (!is_res) {
#pragma unroll
            for (int c = 0; c < CK; ++c)
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const float* arow = sA + c * HH * HW + (S * py + ky) * HW + S * px0;
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const float4 b = *reinterpret_cast<const float4*>(sW + ((ky * 3 + kx) * CK + c) * NT + cg * 4);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float a = arow[S * i + kx];
                            acc[i][0] = fmaf(a, b.x, acc[i][0]);
                            acc[i][1] = fmaf(a, b.y, acc[i][1]);
                            acc[i][2] = fmaf(a, b.z, acc[i][2]);
                            acc[i][3] = fmaf(a, b.w, acc[i][3]);
                        }
                    }
                }
        } else {  // 1x1 res_conv: centre tap only (always stride 1)
#pragma unroll
            for (int c = 0; c < CK; ++c) {
                const float* arow = sA + c * HH * HW + (py + 1) * HW + px0 + 1;
                const float4 b = *reinterpret_cast<const float4*>(sW + c * NT + cg * 4);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float a = arow[i];
                    acc[i][0] = fmaf(a, b.x, acc[i][0]);
                    acc[i][1] = fmaf(a, b.y, acc[i][1]);
                    acc[i][2] = fmaf(a, b.z, acc[i][2]);
                    acc[i][3] = fmaf(a, b.w, acc[i][3]);
                }
            }
        }
    }

    // ---- epilogue ---------------------------------------------------------------------------------
    const int co = co0 + cg * 4;
    float4 add = *reinterpret_cast<const float4*>(p.bias + co);
    if (p.temb) {
        const float4 t = *reinterpret_cast<const float4*>(p.temb + (int64_t)n * p.temb_stride + co);
        add.x += t.x; add.y += t.y; add.z += t.z; add.w += t.w;
    }
    if (p.res_w) {
        const float4 t = *reinterpret_cast<const float4*>(p.res_bias + co);
        add.x += t.x; add.y += t.y; add.z += t.z; add.w += t.w;
    }
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    const int oy = oy0 + py;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t o = (((int64_t)n * p.Hout + oy) * p.Wout + ox0 + px0 + i) * p.Cout + co;
        float4 v = make_float4(acc[i][0] + add.x, acc[i][1] + add.y, acc[i][2] + add.z, acc[i][3] + add.w);
        if (p.res_identity) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(p.res_src[0].x + o));
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        *reinterpret_cast<float4*>(p.out + o) = v;
        s1[0] += v.x; s1[1] += v.y; s1[2] += v.z; s1[3] += v.w;
        s2[0] = fmaf(v.x, v.x, s2[0]); s2[1] = fmaf(v.y, v.y, s2[1]); s2[2] = fmaf(v.z, v.z, s2[2]); s2[3] = fmaf(v.w, v.w, s2[3]);
    }
    if (p.parts) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            red[0][pg][cg * 4 + j] = s1[j];
            red[1][pg][cg * 4 + j] = s2[j];
        }
        __syncthreads();
        if (tid < 2 * NT) {
            const int which = tid / NT, c = tid % NT;
            float t = 0.f;
            for (int g = 0; g < NPG; ++g) t += red[which][g][c];
            p.parts[(((int64_t)n * p.nparts + tile) * p.Cout + co0 + c) * 2 + which] = t;
        }
    }
}

static bool big_tile(int Hout, int Wout) { return Hout % 16 == 0 && Wout % 8 == 0; }

int conv_fp32_nparts(int Hout, int Wout) { return big_tile(Hout, Wout) ? (Hout / 16) * (Wout / 8) : (Hout / 8) * (Wout / 4); }

template <int TH, int TW>
static int launch_tile(const ConvP& p, cudaStream_t st) {
    dim3 grid(p.B * (p.Hout / TH) * (p.Wout / TW), p.Cout / NT);
    constexpr int NTHR = TH * TW * 2;
    switch (p.mode) {
        case CONV_S1: conv3x3_fp32_kernel<TH, TW, CONV_S1><<<grid, NTHR, 0, st>>>(p); break;
        case CONV_S2: conv3x3_fp32_kernel<TH, TW, CONV_S2><<<grid, NTHR, 0, st>>>(p); break;
        case CONV_UP: conv3x3_fp32_kernel<TH, TW, CONV_UP><<<grid, NTHR, 0, st>>>(p); break;
        default: set_error("conv: bad mode %d", p.mode); return SDDM_E_INVALID;
    }
    SDDM_LAUNCH_CHECK();
    return SDDM_OK;
}

int launch_conv_fp32(const ConvP& p, cudaStream_t st) {
    if (p.Cout % NT || p.Cin % CK || (p.nsrc == 2 && p.src[0].C % CK) || (p.res_w && (p.res_Cin % CK || (p.res_nsrc == 2 && p.res_src[0].C % CK)))) {
        set_error("conv fp32: channel counts must be multiples of %d (in) / %d (out): Cin=%d Cout=%d", CK, NT, p.Cin, p.Cout);
        return SDDM_E_INVALID;
    }
    const int ein_h = p.mode == CONV_S2 ? p.Hout * 2 : (p.mode == CONV_UP ? p.Hout / 2 : p.Hout);
    const int ein_w = p.mode == CONV_S2 ? p.Wout * 2 : (p.mode == CONV_UP ? p.Wout / 2 : p.Wout);
    if (ein_h != p.Hin || ein_w != p.Win) { set_error("conv fp32: inconsistent spatial sizes"); return SDDM_E_INVALID; }
    if (p.parts && p.nparts != conv_fp32_nparts(p.Hout, p.Wout)) { set_error("conv fp32: nparts mismatch"); return SDDM_E_INVALID; }
    if (big_tile(p.Hout, p.Wout)) return launch_tile<16, 8>(p, st);
    if (p.Hout % 8 == 0 && p.Wout % 4 == 0) return launch_tile<8, 4>(p, st);
    set_error("conv fp32: output %dx%d is not a multiple of 8x4", p.Hout, p.Wout);
    return SDDM_E_INVALID;
}

}  // namespace sddm
