// Shared declarations for the sddm_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>

namespace sddm {

// ---------------------------------------------------------------------------------------------------
// error plumbing (thread-local message, see sddm_last_error in api.cu)
// ---------------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define SDDM_CUDA_TRY(expr)                                                                          \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            ::sddm::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, \
                              cudaGetErrorString(_e));                                               \
            return SDDM_E_CUDA;                                                                      \
        }                                                                                            \
    } while (0)

#define SDDM_LAUNCH_CHECK()                                                                          \
    do {                                                                                             \
        ::sddm::count_launch();                                                                      \
        cudaError_t _e = cudaPeekAtLastError();                                                      \
        if (_e != cudaSuccess) {                                                                     \
            ::sddm::set_error("kernel launch failed %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, \
                              __LINE__, cudaGetErrorString(_e));                                     \
            return SDDM_E_CUDA;                                                                      \
        }                                                                                            \
    } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize applies per (kernel, DEVICE): set it once per device, not once per process
// (one process may build plans on several GPUs).  Each expansion site owns its flags.
#define SDDM_SET_MAX_SMEM(kernel, bytes)                                                                          \
    do {                                                                                                         \
        static bool _done[64] = {};                                                                              \
        int _dev = 0;                                                                                            \
        if (cudaGetDevice(&_dev) != cudaSuccess) _dev = -1;                                                      \
        if (_dev < 0 || _dev >= 64 || !_done[_dev]) {                                                            \
            SDDM_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            if (_dev >= 0 && _dev < 64) _done[_dev] = true;                                                      \
        }                                                                                                        \
    } while (0)

// SM count of the CURRENT device (cached per device)
inline int device_sm_count() {
    static int cache[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return 148; }
    if (cache[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 148; }
        cache[dev] = n;
    }
    return cache[dev];
}

// ---------------------------------------------------------------------------------------------------
// convolution op descriptor, shared by the CUDA-core fp32 kernel and the tcgen05 bf16 kernel.
// Activations are NHWC fp32 (or bf16 when act16; the pointers below are then reinterpreted) ("raw", i.e. pre-GroupNorm); a source with scale != nullptr is consumed as
// swish(x * scale[n][c] + shift[n][c]) (GroupNorm-apply + Swish fused into the operand staging); the
// zero padding of the convolution is applied AFTER that transform (UNetModified2.py:116-121).
// ---------------------------------------------------------------------------------------------------
struct ConvSrc {
    const float* x;      // [B][Hin][Win][C]
    const float* scale;  // [B][Cin_total] (indexed with the concatenated channel index) or nullptr = raw
    const float* shift;
    int C;
};

// GroupNorm of the CONSUMER, finalised by the producer kernel's last-arriving CTA per sample (gn_fuse.cuh)
struct GnFuse {
    const float* parts[2];   // partial statistics of every source of the consumer's GroupNorm (this kernel's output first)
    int C[2];
    int nparts[2];
    int nsrc;
    const float* gamma;      // [Ctot]
    const float* beta;
    float* scale;            // [B][Ctot]
    float* shift;
    int Ctot, groups, HW;
    float eps;
    unsigned int* counter;   // [B], zero outside the kernel
    int expect;              // arrivals per sample = tiles per sample of the producer
};

enum ConvMode { CONV_S1 = 0, CONV_S2 = 1, CONV_UP = 2 };  // 3x3 pad 1: stride 1 | stride 2 | nearest x2 then stride 1

struct ConvP {
    ConvSrc src[2];  // channel concatenation [src0, src1] (torch.cat(dim=1), UNetModified2.py:263)
    int nsrc;
    int Cin;         // total input channels
    int Hin, Win;    // source spatial size
    int Hout, Wout;
    int Cout;
    int mode;
    const float* w;             // fp32 pack  [Cin/8][9][8][Cout]
    const __nv_bfloat16* w_tc;  // bf16 pack  [Cin/32][9][4][Cout][8]  (UMMA K-major core-matrix image)
    const float* bias;          // [Cout]
    const float* temb;          // nullable; + temb[n * temb_stride + co]  (FeatureWiseAffine, additive)
    int temb_stride;
    // residual: out += res_conv(x) with x = raw concat of res_src (1x1 conv), or += x when res_identity
    ConvSrc res_src[2];
    int res_nsrc;
    int res_Cin;
    int res_identity;
    const float* res_w;             // fp32 pack [Cin/8][8][Cout]
    const __nv_bfloat16* res_w_tc;  // bf16 pack [Cin/32][4][Cout][8]
    const float* res_bias;
    float* out;    // [B][Hout][Wout][Cout]
    float* parts;  // GroupNorm partial statistics of `out`: [B][nparts][Cout][2] (sum, sum of squares)
    int nparts;
    int B;
    int act16;     // activations (every src / res_src / out tensor) are bf16 in HBM instead of fp32 (tcgen05 path only)
    int x3;        // tcgen05 path, fp32 activations: split-bf16 operands (hi, lo) and three products per MMA step (SDDM_PREC_BF16X3)
    const __nv_bfloat16* w_tc_lo;      // bf16(w - bf16(w)) in the w_tc layout
    const __nv_bfloat16* res_w_tc_lo;
    int gn_on;     // tcgen05 path: finalise the consumer's GroupNorm in this kernel (gn below), no gn_finalize launch
    GnFuse gn;
};

int launch_conv_fp32(const ConvP& p, cudaStream_t st);
int launch_conv_tc(const ConvP& p, cudaStream_t st);   // tcgen05 path; requires Hout%16==0, Wout%8==0, C%32==0
cudaError_t conv_tc_set_hang_buffer(unsigned* dev_ptr);   // debug notes of timed-out waits (sddm_debug_hang); no-op in a normal build
bool conv_tc_supported(const ConvP& p);
int conv_fp32_nparts(int Hout, int Wout);
int conv_tc_nparts(int Hout, int Wout);
int conv_tc_tiles(int Hout, int Wout);
// row-streaming tcgen05 kernel of the 128-wide level (conv_row.cu): bf16 activations, Cout = 32 (or 1: the final Block -> frames)
bool conv_row_supported(const ConvP& p);
int conv_row_nparts(int H, int W);   // partial-statistics slots per sample: (H / 16) blocks x 2 groups x 4 warps (64-wide level: (H / 4) x 2 x 2)
int conv_row_arrivals(int H);    // arrivals per sample of the fused GroupNorm finalisation
struct PostP;
// post != nullptr (final Block only): fuse the overlap-add of the frames + the posterior update described by (post, k8) into the epilogue
int launch_conv_row(const ConvP& p, const __nv_bfloat16* w_row, uint32_t w_bytes, float* frames, float final_bias, cudaStream_t st,
                    const PostP* post, const float* k8);   // tiles per sample (= arrivals per sample of the fused GroupNorm finalisation)

// ---------------------------------------------------------------------------------------------------
// programmatic dependent launch (PDL): a kernel launched with launch_pdl() may start while its predecessor in the stream
// is still running; it must execute pdl_wait() before it touches anything the predecessor produced (and before it writes
// global memory), and a predecessor calls pdl_launch_dependents() to let the successor's CTAs be scheduled early.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float swish_accurate(float v) { return v / (1.0f + expf(-v)); }
__device__ __forceinline__ float swish_fast(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

// Philox4x32-10 (Salmon et al. 2011), counter-based: results depend only on (key, counter).
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

// four N(0,1) samples from one Philox block (Box-Muller on (0,1] uniforms).  The transcendental parts use the hardware
// approximations explicitly (lg2.approx / sin.approx / cos.approx: absolute error ~1e-6, far below what a Gaussian draw needs),
// so every translation unit - whatever its math flags - produces the SAME stream for a given (seed, element, row, draw).
__device__ __forceinline__ float2 box_muller_fast(uint32_t a, uint32_t b) {
    const float k = 2.3283064365386963e-10f;  // 2^-32
    const float u0 = fminf(((float)a + 1.0f) * k, 1.0f), u1 = (float)b * k;
    const float m = sqrtf(-2.0f * __logf(u0));
    float s, c;
    __sincosf(6.283185307179586f * u1, &s, &c);
    return make_float2(m * c, m * s);
}
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint32_t elem4, uint64_t row, uint32_t draw) {
    const uint4 r = philox4x32_10(make_uint4(elem4, (uint32_t)row, draw, (uint32_t)(row >> 32)),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const float2 p0 = box_muller_fast(r.x, r.y), p1 = box_muller_fast(r.z, r.w);
    return make_float4(p0.x, p0.y, p1.x, p1.y);
}
// component `comp` (0..3) of philox_normal4 - only the Box-Muller pair that holds it is evaluated
__device__ __forceinline__ float philox_normal1(uint64_t seed, uint32_t elem4, uint64_t row, uint32_t draw, int comp) {
    const uint4 r = philox4x32_10(make_uint4(elem4, (uint32_t)row, draw, (uint32_t)(row >> 32)),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const float2 p = (comp & 2) ? box_muller_fast(r.z, r.w) : box_muller_fast(r.x, r.y);
    return (comp & 1) ? p.y : p.x;
}

// Warp-wide column sums of 32 per-lane values with 31 shuffles (recursive halving):
// on return, lane l holds in v[0] the sum over all 32 lanes of the value with index l.
__device__ __forceinline__ float warp_transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
    for (int step = 0; step < 5; ++step) {
        const int width = 16 >> step;
        const bool upper = (lane & width) != 0;
#pragma unroll
        for (int i = 0; i < width; ++i) {
            float send = upper ? v[i] : v[i + width];
            float keep = upper ? v[i + width] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, width);
        }
    }
    return v[0];
}

}  // namespace sddm
