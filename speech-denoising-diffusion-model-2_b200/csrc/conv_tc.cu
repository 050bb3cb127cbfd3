// tcgen05 / TMEM implicit-GEMM 3x3 convolution of the UNetModified2 denoiser (sm_100a).
//
//   GEMM view : M = 128 output pixels (a 16 x 8 window of one sample), N = Cout (32..160), K = 9 * Cin (+ res_Cin)
//   operands  : bf16, fp32 accumulation in tensor memory (two accumulator stages: MMA of tile i+1 overlaps the
//               epilogue of tile i)
//   A operand : the zero-padded, post-activation halo window is staged ONCE per 32-channel K slab in shared memory
//               in the no-swizzle K-major core-matrix layout  [k8 plane][halo pixel][8 channels = 16 B];  the 9 taps
//               are 9 shared-memory descriptors with shifted start addresses over the same slab (im2col-free).
//               Producer warps fuse GroupNorm-apply + Swish + concat / nearest-x2 / stride-2 addressing + bf16
//               conversion into the staging pass (zero padding is applied AFTER the activation, UNetModified2.py:116-121).
//   B operand : weights pre-packed on the host as [Cin/16][tap][2][Cout][8] bf16, streamed per 16-channel chunk with
//               cp.async.bulk (or kept resident in shared memory for the whole persistent CTA when they fit).
//   extra K   : the ResnetBlock 1x1 res_conv over the raw block input accumulates into the same TMEM tile.
//   epilogue  : tcgen05.ld -> + bias (+ noise-level embedding, + res bias) (+ identity residual) -> NHWC fp32 store
//               + GroupNorm partial statistics (sum, sum of squares per channel per warp) for the next layer.
//
// Warp roles (448 threads, persistent CTA, static tile schedule):
//   warps 0-3 epilogue (TMEM lane quarter = warp id) | warp 4 MMA issuer + TMEM owner | warp 5 weight loader
//   warps 6-13 A-operand producers
//
// reference: Block / ResnetBlock / Downsample / Upsample, model/UNetModified2.py:93-142
#include <vector>

#include "common.cuh"
#include "../../include/sddm_b200.h"

namespace sddm {
namespace {

constexpr int TH = 16, TW = 8;            // output window of one tile
constexpr int kEpiWarps = 4;
constexpr int kProdWarp0 = 6, kProdWarps = 8;
constexpr int kProdThreads = kProdWarps * 32;
constexpr int kThreads = (kProdWarp0 + kProdWarps) * 32;   // 448
constexpr int kMaxA = 4, kMaxW = 16;
constexpr int kItems = 3;                 // halo items (pixel x 8 channels) per producer thread per batch

// halo geometry per mode ------------------------------------------------------------------------------
template <int MODE> struct Geo;
template <> struct Geo<CONV_S1> { static constexpr int PH = 18, PW = 10, PLANE = 181, SBO = 10 * 16; };
template <> struct Geo<CONV_UP> { static constexpr int PH = 18, PW = 10, PLANE = 181, SBO = 10 * 16; };
template <> struct Geo<CONV_S2> { static constexpr int PH = 33, PW = 17, PLANE = 565, SBO = 2 * 17 * 16; };
// PLANE (slots of 16 B between the k8 planes) is = 5 (mod 8): the four planes of one pixel then fall into disjoint banks.

struct TcArgs {
    ConvP p;
    int tiles_x, tiles_y, ntiles;   // per-sample tile grid, total tiles (B * tiles_x * tiles_y)
    int nA_main, nA_res;            // 32-channel A slabs of the main conv / the 1x1 res_conv
    int NA, NW;                     // ring depths
    int resident;                   // all weight chunks stay in shared memory
    int acc_stride, tmem_cols;
    uint32_t a_stage_bytes, w_stage_bytes;
};

// ---- PTX wrappers -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a pipeline bug must surface as a trap (CUDA error), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    uint32_t spins = 0;
    while (!mbar_try(bar, parity)) {
        if ((++spins & 1023u) == 0 && clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// shared-memory matrix descriptor, SWIZZLE_NONE, K-major: core matrix = 8 rows x 16 B (128 contiguous bytes);
// LBO = byte distance between the two K halves (8 elements each) of one MMA K step, SBO = byte distance between
// consecutive 8-row groups along M / N.  Bit 46 = descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor, kind::f16: D = fp32, A = B = bf16, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane base + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// shared-memory carve-up (dynamic smem, 128-byte aligned base)
struct SmemHdr {
    uint64_t full_a[kMaxA], empty_a[kMaxA];
    uint64_t full_w[kMaxW], empty_w[kMaxW];
    uint64_t tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad[31];
    float addv[256];
};
constexpr int kHdrBytes = 2048;
static_assert(sizeof(SmemHdr) <= kHdrBytes, "header too large");

struct PBatch {   // one producer batch in flight: raw values + the per-channel affine of the slab
    float4 v[kItems][2];
    float4 sc[2], sh[2];
    int slot[kItems];      // halo slot to write (-1: no item)
    uint32_t okmask;       // bit r: item r reads real data (else zero padding)
    int has_affine;
};

// =====================================================================================================
template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) conv3x3_tc_kernel(const TcArgs a) {
    using G = Geo<MODE>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SmemHdr* hdr = reinterpret_cast<SmemHdr*>(smem_raw);
    const uint32_t smem_base = smem_u32(smem_raw);
    const uint32_t a_base = smem_base + kHdrBytes;
    const uint32_t w_base = a_base + (uint32_t)a.NA * a.a_stage_bytes;
    const ConvP& p = a.p;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nA = a.nA_main + a.nA_res;
    const int tiles_per_img = a.tiles_x * a.tiles_y;

    if (tid == 0) {
        for (int i = 0; i < kMaxA; ++i) { mbar_init(smem_u32(&hdr->full_a[i]), kProdThreads); mbar_init(smem_u32(&hdr->empty_a[i]), 1); }
        for (int i = 0; i < kMaxW; ++i) { mbar_init(smem_u32(&hdr->full_w[i]), 1); mbar_init(smem_u32(&hdr->empty_w[i]), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&hdr->tmem_full[i]), 1); mbar_init(smem_u32(&hdr->tmem_empty[i]), kEpiWarps * 32); }
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(smem_u32(&hdr->tmem_base), (uint32_t)a.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = hdr->tmem_base;

    if (warp < kEpiWarps) {
        // ============================== epilogue ==========================================================
        const int m = tid, py = m >> 3, px = m & 7;
        const int nblk = p.Cout >> 5;
        for (int it = 0, tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
            const int as = it & 1;
            const uint32_t aph = (uint32_t)(it >> 1) & 1u;
            const int n = tile / tiles_per_img, trem = tile - n * tiles_per_img;
            const int oy = (trem / a.tiles_x) * TH + py, ox = (trem % a.tiles_x) * TW + px;
            const bool valid = oy < p.Hout && ox < p.Wout;
            epi_bar();   // previous tile's readers of addv are done
            for (int c = tid; c < p.Cout; c += kEpiWarps * 32) {
                float v = __ldg(p.bias + c);
                if (p.temb) v += __ldg(p.temb + (int64_t)n * p.temb_stride + c);
                if (p.res_w_tc) v += __ldg(p.res_bias + c);
                hdr->addv[c] = v;
            }
            epi_bar();
            const int64_t obase = (((int64_t)n * p.Hout + oy) * p.Wout + ox) * p.Cout;
            const float* rptr = (p.res_identity && valid) ? p.res_src[0].x + obase : nullptr;
            float4 rn[8];
            if (rptr) {
#pragma unroll
                for (int q = 0; q < 8; ++q) rn[q] = __ldg(reinterpret_cast<const float4*>(rptr) + q);
            }
            mbar_wait(smem_u32(&hdr->tmem_full[as]), aph);
            tc_fence_after();
            const uint32_t tacc = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(as * a.acc_stride);
            for (int cb = 0; cb < nblk; ++cb) {
                float v[32], q2[32];
                tmem_ld32(tacc + (uint32_t)(cb * 32), v);
                float4 rc[8];
                if (rptr) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) rc[q] = rn[q];
                    if (cb + 1 < nblk) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) rn[q] = __ldg(reinterpret_cast<const float4*>(rptr + (cb + 1) * 32) + q);
                    }
                }
#pragma unroll
                for (int k = 0; k < 32; ++k) v[k] += hdr->addv[cb * 32 + k];
                if (rptr) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        v[4 * q + 0] += rc[q].x; v[4 * q + 1] += rc[q].y; v[4 * q + 2] += rc[q].z; v[4 * q + 3] += rc[q].w;
                    }
                }
                if (valid) {
                    float4* o = reinterpret_cast<float4*>(p.out + obase + cb * 32);
#pragma unroll
                    for (int q = 0; q < 8; ++q) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                } else {
#pragma unroll
                    for (int k = 0; k < 32; ++k) v[k] = 0.f;
                }
                if (p.parts) {
#pragma unroll
                    for (int k = 0; k < 32; ++k) q2[k] = v[k] * v[k];
                    const float s1 = warp_transpose_reduce32(v, lane);
                    const float s2 = warp_transpose_reduce32(q2, lane);
                    float2* dst = reinterpret_cast<float2*>(p.parts) + ((int64_t)n * p.nparts + trem * 4 + warp) * p.Cout + cb * 32 + lane;
                    *dst = make_float2(s1, s2);
                }
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&hdr->tmem_empty[as]));
        }
    } else if (warp == 4) {
        // ============================== MMA issuer (one thread) ===========================================
        if (lane == 0) {
            const uint32_t idesc = make_idesc(p.Cout);
            const uint32_t b_lbo = (uint32_t)p.Cout * 16u, b_sbo = 128u;
            const uint32_t a_lbo = (uint32_t)G::PLANE * 16u, a_sbo = (uint32_t)G::SBO;
            int sa = 0, sw = 0;
            uint32_t pa = 0, pw = 0;
            for (int it = 0, tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
                const int as = it & 1;
                const uint32_t aph = (uint32_t)(it >> 1) & 1u;
                mbar_wait(smem_u32(&hdr->tmem_empty[as]), aph ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * a.acc_stride);
                uint32_t acc = 0;
                for (int ai = 0; ai < nA; ++ai) {
                    const bool is_res = ai >= a.nA_main;
                    mbar_wait(smem_u32(&hdr->full_a[sa]), pa);
                    const uint32_t a_stage = a_base + (uint32_t)sa * a.a_stage_bytes;
                    for (int h = 0; h < 2; ++h) {
                        const int c = 2 * ai + h;
                        int wslot;
                        if (a.resident) { wslot = c; mbar_wait(smem_u32(&hdr->full_w[c]), 0u); }
                        else { wslot = sw; mbar_wait(smem_u32(&hdr->full_w[sw]), pw); }
                        tc_fence_after();
                        const uint32_t w_stage = w_base + (uint32_t)wslot * a.w_stage_bytes;
                        const uint32_t a_half = a_stage + (uint32_t)h * 2u * a_lbo;
                        if (!is_res) {
#pragma unroll
                            for (int tap = 0; tap < 9; ++tap) {
                                const int ky = tap / 3, kx = tap - 3 * ky;
                                uint32_t aoff;
                                if (MODE == CONV_S2) aoff = (uint32_t)(ky * G::PW + (kx == 1 ? 9 : (kx >> 1))) * 16u;
                                else aoff = (uint32_t)(ky * G::PW + kx) * 16u;
                                umma(d_tmem, make_desc(a_half + aoff, a_lbo, a_sbo),
                                     make_desc(w_stage + (uint32_t)tap * 2u * b_lbo, b_lbo, b_sbo), idesc, acc);
                                acc = 1;
                            }
                        } else {   // 1x1 res_conv: centre tap of the raw (stride-1) halo
                            umma(d_tmem, make_desc(a_half + (uint32_t)(G::PW + 1) * 16u, a_lbo, a_sbo),
                                 make_desc(w_stage, b_lbo, b_sbo), idesc, acc);
                            acc = 1;
                        }
                        if (!a.resident) {
                            umma_commit(smem_u32(&hdr->empty_w[sw]));
                            if (++sw == a.NW) { sw = 0; pw ^= 1u; }
                        }
                    }
                    umma_commit(smem_u32(&hdr->empty_a[sa]));
                    if (++sa == a.NA) { sa = 0; pa ^= 1u; }
                }
                umma_commit(smem_u32(&hdr->tmem_full[as]));
            }
        }
    } else if (warp == 5) {
        // ============================== weight loader (one thread) ========================================
        if (lane == 0) {
            const int nchunks = 2 * nA, nmain = 2 * a.nA_main;
            const uint32_t main_bytes = 288u * (uint32_t)p.Cout, res_bytes = 32u * (uint32_t)p.Cout;
            int sw = 0;
            uint32_t pw = 0;
            for (int it = 0, tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
                if (a.resident && it > 0) break;
                for (int c = 0; c < nchunks; ++c) {
                    const bool is_res = c >= nmain;
                    const int wslot = a.resident ? c : sw;
                    if (!a.resident) mbar_wait(smem_u32(&hdr->empty_w[sw]), pw ^ 1u);
                    const uint32_t bytes = is_res ? res_bytes : main_bytes;
                    const void* src = is_res ? (const void*)(p.res_w_tc + (size_t)(c - nmain) * 16 * p.Cout)
                                             : (const void*)(p.w_tc + (size_t)c * 144 * p.Cout);
                    const uint32_t bar = smem_u32(&hdr->full_w[wslot]);
                    mbar_expect_tx(bar, bytes);
                    bulk_g2s(w_base + (uint32_t)wslot * a.w_stage_bytes, src, bytes, bar);
                    if (!a.resident && ++sw == a.NW) { sw = 0; pw ^= 1u; }
                }
            }
        }
    } else {
        // ============================== A-operand producers ===============================================
        constexpr int NPIX = G::PH * G::PW;
        constexpr int RPS = (NPIX * 4 + kProdThreads * kItems - 1) / (kProdThreads * kItems);   // batches per slab
        const int ptid = tid - kProdWarp0 * 32;
        const int j = ptid & 3;          // k8 plane of this thread (channels j*8 .. j*8+7 of the slab)
        const int pix0 = ptid >> 2;      // first halo pixel
        unsigned char* a_gen = smem_raw + kHdrBytes;

        // flattened batch iterator: (tile, slab, round)
        int b_tile = blockIdx.x, b_ai = 0, b_round = 0;
        auto issue = [&](PBatch& B) {
            const int tile = b_tile, ai = b_ai, round = b_round;
            const int n = tile / tiles_per_img, trem = tile - n * tiles_per_img;
            const int oy0 = (trem / a.tiles_x) * TH, ox0 = (trem % a.tiles_x) * TW;
            const bool is_res = ai >= a.nA_main;
            const int cbase = (is_res ? ai - a.nA_main : ai) * 32;
            const ConvSrc* srcs = is_res ? p.res_src : p.src;
            const int s = (cbase < srcs[0].C) ? 0 : 1;
            const float* x = srcs[s].x;
            const int C = srcs[s].C;
            const int coff = cbase - (s ? srcs[0].C : 0) + j * 8;
            const float* scale = srcs[s].scale;
            B.has_affine = (!is_res && scale != nullptr) ? 1 : 0;
            if (B.has_affine) {
                const int ctot = p.Cin;
                const float4* sp = reinterpret_cast<const float4*>(scale + (int64_t)n * ctot + cbase + j * 8);
                const float4* hp = reinterpret_cast<const float4*>(srcs[s].shift + (int64_t)n * ctot + cbase + j * 8);
                B.sc[0] = __ldg(sp); B.sc[1] = __ldg(sp + 1);
                B.sh[0] = __ldg(hp); B.sh[1] = __ldg(hp + 1);
            }
            B.okmask = 0;
#pragma unroll
            for (int r = 0; r < kItems; ++r) {
                const int pix = pix0 + (kProdThreads / 4) * (round * kItems + r);
                B.slot[r] = -1;
                if (pix >= NPIX) continue;
                const int hy = pix / G::PW, hx = pix - hy * G::PW;
                int iy, ix, slot;
                bool ok;
                if (MODE == CONV_UP && !is_res) {
                    const int uy = oy0 + hy - 1, ux = ox0 + hx - 1;
                    ok = uy >= 0 && uy < p.Hout && ux >= 0 && ux < p.Wout;
                    iy = uy >> 1; ix = ux >> 1;
                    slot = pix;
                } else if (MODE == CONV_S2) {
                    iy = 2 * oy0 + hy - 1; ix = 2 * ox0 + hx - 1;
                    ok = iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win;
                    slot = hy * G::PW + ((hx & 1) ? 9 + (hx >> 1) : (hx >> 1));
                } else {
                    iy = oy0 + hy - 1; ix = ox0 + hx - 1;
                    ok = iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win;
                    slot = pix;
                }
                B.slot[r] = slot;
                if (ok) {
                    B.okmask |= 1u << r;
                    const float4* g = reinterpret_cast<const float4*>(x + (((int64_t)n * p.Hin + iy) * p.Win + ix) * C + coff);
                    B.v[r][0] = __ldg(g);
                    B.v[r][1] = __ldg(g + 1);
                }
            }
            // advance the iterator
            if (++b_round == RPS) {
                b_round = 0;
                if (++b_ai == nA) { b_ai = 0; b_tile += gridDim.x; }
            }
        };

        int sa = 0;
        uint32_t pa = 0;
        int c_round = 0;
        auto finish = [&](const PBatch& B) {
            if (c_round == 0) mbar_wait(smem_u32(&hdr->empty_a[sa]), pa ^ 1u);
            unsigned char* stage = a_gen + (size_t)sa * a.a_stage_bytes + (size_t)j * G::PLANE * 16;
#pragma unroll
            for (int r = 0; r < kItems; ++r) {
                if (B.slot[r] < 0) continue;
                float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (B.okmask & (1u << r)) {
                    f[0] = B.v[r][0].x; f[1] = B.v[r][0].y; f[2] = B.v[r][0].z; f[3] = B.v[r][0].w;
                    f[4] = B.v[r][1].x; f[5] = B.v[r][1].y; f[6] = B.v[r][1].z; f[7] = B.v[r][1].w;
                    if (B.has_affine) {
                        const float scv[8] = {B.sc[0].x, B.sc[0].y, B.sc[0].z, B.sc[0].w, B.sc[1].x, B.sc[1].y, B.sc[1].z, B.sc[1].w};
                        const float shv[8] = {B.sh[0].x, B.sh[0].y, B.sh[0].z, B.sh[0].w, B.sh[1].x, B.sh[1].y, B.sh[1].z, B.sh[1].w};
#pragma unroll
                        for (int k = 0; k < 8; ++k) f[k] = swish_fast(fmaf(f[k], scv[k], shv[k]));
                    }
                }
                uint4 o;
                o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]); o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
                *reinterpret_cast<uint4*>(stage + (size_t)B.slot[r] * 16) = o;
            }
            if (++c_round == RPS) {
                c_round = 0;
                fence_async_smem();
                mbar_arrive(smem_u32(&hdr->full_a[sa]));
                if (++sa == a.NA) { sa = 0; pa ^= 1u; }
            }
        };

        int my_tiles = 0;
        for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) ++my_tiles;
        const int total = my_tiles * nA * RPS;
        PBatch cur, nxt;
        if (total > 0) issue(cur);
        for (int b = 0; b < total; ++b) {
            if (b + 1 < total) issue(nxt);
            finish(cur);
            cur = nxt;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        __syncwarp();
        tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
    }
}

// =====================================================================================================
// single-tile descriptor self-test (sddm_debug_umma_probe)
// =====================================================================================================
__global__ void __launch_bounds__(128) umma_probe_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ Bm,
                                                         float* __restrict__ D, int N, int K, uint32_t a_off, uint32_t a_lbo,
                                                         uint32_t a_sbo, int swap_fields) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t sA = smem_u32(smem_raw), a_bytes = a_off + (uint32_t)(K / 8) * a_lbo + 16u * a_sbo + 256u;
    const uint32_t b_off = (a_bytes + 127u) & ~127u, b_lbo = (uint32_t)N * 16u, b_sbo = 128u;
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&tbase), 256);
    for (int i = tid; i < 128 * K; i += 128) {
        const int m = i / K, k = i - m * K;
        *reinterpret_cast<__nv_bfloat16*>(smem_raw + a_off + (k / 8) * a_lbo + (m / 8) * a_sbo + (m % 8) * 16 + (k % 8) * 2) = A[i];
    }
    for (int i = tid; i < N * K; i += 128) {
        const int n = i / K, k = i - n * K;
        *reinterpret_cast<__nv_bfloat16*>(smem_raw + b_off + (k / 8) * b_lbo + (n / 8) * b_sbo + (n % 8) * 16 + (k % 8) * 2) = Bm[i];
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tbase;
    if (tid == 0) {
        const uint32_t idesc = make_idesc(N);
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint32_t aa = sA + a_off + (uint32_t)ks * 2u * a_lbo, bb = sA + b_off + (uint32_t)ks * 2u * b_lbo;
            const uint64_t da = swap_fields ? make_desc(aa, a_sbo, a_lbo) : make_desc(aa, a_lbo, a_sbo);
            const uint64_t db = swap_fields ? make_desc(bb, b_sbo, b_lbo) : make_desc(bb, b_lbo, b_sbo);
            umma(tmem, da, db, idesc, ks > 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    for (int cb = 0; cb < N / 32; ++cb) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cb * 32), v);
        for (int k = 0; k < 32; ++k) D[(size_t)tid * N + cb * 32 + k] = v[k];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { __syncwarp(); tmem_dealloc(tmem, 256); }
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
            cudaGetLastError();
            n = 148;
        }
    }
    return n;
}

constexpr size_t kSmemMax = 232448;   // 227 KB opt-in limit per CTA on sm_100

template <int MODE>
int launch_mode(TcArgs& a, cudaStream_t st) {
    using G = Geo<MODE>;
    const ConvP& p = a.p;
    a.a_stage_bytes = 4u * G::PLANE * 16u;
    a.w_stage_bytes = 288u * (uint32_t)p.Cout;
    a.NA = (MODE == CONV_S2) ? 2 : 4;
    const int nchunks = 2 * (a.nA_main + a.nA_res);
    const size_t budget = kSmemMax - kHdrBytes - (size_t)a.NA * a.a_stage_bytes;
    int nw = (int)(budget / a.w_stage_bytes);
    if (nw > kMaxW) nw = kMaxW;
    if (nw < 2) { set_error("conv tc: Cout=%d leaves no room for a weight ring", p.Cout); return SDDM_E_INVALID; }
    a.resident = nchunks <= nw;
    a.NW = a.resident ? nchunks : nw;
    const size_t smem = kHdrBytes + (size_t)a.NA * a.a_stage_bytes + (size_t)a.NW * a.w_stage_bytes;
    static bool attr_set = false;
    if (!attr_set) {
        SDDM_CUDA_TRY(cudaFuncSetAttribute(conv3x3_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
        attr_set = true;
    }
    const int grid = a.ntiles < num_sms() ? a.ntiles : num_sms();
    conv3x3_tc_kernel<MODE><<<grid, kThreads, smem, st>>>(a);
    SDDM_LAUNCH_CHECK();
    return SDDM_OK;
}

}  // namespace

bool conv_tc_supported(const ConvP& p) {
    if (p.Cout % 32 || p.Cout < 32 || p.Cout > 256 || p.Cin % 32 || p.Cin < 32) return false;
    for (int i = 0; i < p.nsrc; ++i)
        if (p.src[i].C % 32) return false;
    if (p.res_Cin % 32) return false;
    return p.mode == CONV_S1 || p.mode == CONV_S2 || p.mode == CONV_UP;
}

int conv_tc_nparts(int Hout, int Wout) { return ((Hout + TH - 1) / TH) * ((Wout + TW - 1) / TW) * 4; }

int launch_conv_tc(const ConvP& p, cudaStream_t st) {
    if (!conv_tc_supported(p) || !p.w_tc) { set_error("conv tc: unsupported shape Cin=%d Cout=%d mode=%d", p.Cin, p.Cout, p.mode); return SDDM_E_INVALID; }
    const int ein_h = p.mode == CONV_S2 ? p.Hout * 2 : (p.mode == CONV_UP ? p.Hout / 2 : p.Hout);
    const int ein_w = p.mode == CONV_S2 ? p.Wout * 2 : (p.mode == CONV_UP ? p.Wout / 2 : p.Wout);
    if (ein_h != p.Hin || ein_w != p.Win) { set_error("conv tc: inconsistent spatial sizes"); return SDDM_E_INVALID; }
    if (p.res_Cin && !p.res_identity && (!p.res_w_tc || p.mode != CONV_S1)) { set_error("conv tc: res_conv needs packed weights and stride 1"); return SDDM_E_INVALID; }
    TcArgs a{};
    a.p = p;
    a.tiles_x = (p.Wout + TW - 1) / TW;
    a.tiles_y = (p.Hout + TH - 1) / TH;
    a.ntiles = p.B * a.tiles_x * a.tiles_y;
    if (p.parts && p.nparts != a.tiles_x * a.tiles_y * 4) { set_error("conv tc: nparts mismatch"); return SDDM_E_INVALID; }
    a.nA_main = p.Cin / 32;
    a.nA_res = (p.res_w_tc && !p.res_identity) ? p.res_Cin / 32 : 0;
    a.acc_stride = p.Cout <= 32 ? 32 : (p.Cout <= 64 ? 64 : (p.Cout <= 128 ? 128 : 256));
    a.tmem_cols = 2 * a.acc_stride;
    switch (p.mode) {
        case CONV_S1: return launch_mode<CONV_S1>(a, st);
        case CONV_S2: return launch_mode<CONV_S2>(a, st);
        default: return launch_mode<CONV_UP>(a, st);
    }
}

}  // namespace sddm

// ---------------------------------------------------------------------------------------------------
// probe: D[128 x N] = A[128 x K] * B[N x K]^T through one CTA's tcgen05 path, compared on the host.
//   variant % 10 : 0 canonical A (SBO 128, LBO 2048) | 1 stride-1 halo geometry, centre-tap offset | 2 stride-2 geometry
//   variant / 10 : 1 = LBO / SBO descriptor fields swapped (convention cross-check; expected to FAIL)
// ---------------------------------------------------------------------------------------------------
extern "C" SDDM_API int sddm_debug_umma_probe(int variant, int N, int K, float* max_err_host) {
    using namespace sddm;
    if (!max_err_host || N % 32 || N < 32 || N > 256 || K % 16 || K < 16 || K > 64) { set_error("probe: N in 32..256 step 32, K in 16..64 step 16"); return SDDM_E_INVALID; }
    const int geo = variant % 10, swap = variant / 10;
    uint32_t a_off = 0, a_lbo = 2048, a_sbo = 128;
    if (geo == 1) { a_off = 11 * 16; a_lbo = 181 * 16; a_sbo = 160; }
    if (geo == 2) { a_off = (17 + 9) * 16; a_lbo = 565 * 16; a_sbo = 544; }
    std::vector<__nv_bfloat16> hA((size_t)128 * K), hB((size_t)N * K);
    std::vector<float> fA(hA.size()), fB(hB.size());
    uint32_t s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((int)((s >> 8) & 0xFFFF) - 32768) / 32768.0f; };
    for (size_t i = 0; i < hA.size(); ++i) { hA[i] = __float2bfloat16(rnd()); fA[i] = __bfloat162float(hA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { hB[i] = __float2bfloat16(rnd()); fB[i] = __bfloat162float(hB[i]); }
    __nv_bfloat16 *dA = nullptr, *dB = nullptr;
    float* dD = nullptr;
    SDDM_CUDA_TRY(cudaMalloc(&dA, hA.size() * 2));
    SDDM_CUDA_TRY(cudaMalloc(&dB, hB.size() * 2));
    SDDM_CUDA_TRY(cudaMalloc(&dD, (size_t)128 * N * 4));
    SDDM_CUDA_TRY(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    SDDM_CUDA_TRY(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    SDDM_CUDA_TRY(cudaMemset(dD, 0xFF, (size_t)128 * N * 4));
    const size_t a_bytes = a_off + (size_t)(K / 8) * a_lbo + 16 * a_sbo + 256;
    const size_t smem = ((a_bytes + 127) & ~(size_t)127) + (size_t)(K / 8) * N * 16 + 256;
    SDDM_CUDA_TRY(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_probe_kernel<<<1, 128, smem>>>(dA, dB, dD, N, K, a_off, a_lbo, a_sbo, swap);
    count_launch();
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> hD((size_t)128 * N);
    if (e == cudaSuccess) e = cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    if (e != cudaSuccess) { set_error("probe: CUDA error %s: %s", cudaGetErrorName(e), cudaGetErrorString(e)); return SDDM_E_CUDA; }
    float worst = 0.f;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double ref = 0.0;
            for (int k = 0; k < K; ++k) ref += (double)fA[(size_t)m * K + k] * (double)fB[(size_t)n * K + k];
            const float d = hD[(size_t)m * N + n];
            const float err = (d == d) ? fabsf(d - (float)ref) : 1e30f;
            if (err > worst) worst = err;
        }
    *max_err_host = worst;
    return SDDM_OK;
}
