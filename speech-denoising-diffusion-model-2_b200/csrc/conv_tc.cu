// tcgen05 / TMEM implicit-GEMM 3x3 convolution of the UNetModified2 denoiser (sm_100a).
//
//   GEMM view : M = 128 output pixels (a 16 x 8 window of one sample), N = Cout (32..160), K = 9 * Cin (+ res_Cin)
//   operands  : bf16, fp32 accumulation in tensor memory (two accumulator stages: the MMAs of tile i+1 overlap the
//               epilogue of tile i)
//   raw load  : one TMA tensor load (cp.async.bulk.tensor.4d) per K slab brings the fp32 NHWC halo window of the source
//               (zero filled outside the image) into a shared-memory ring; stride-2 and nearest-x2 convolutions only
//               change the box (33x17 / 10x6 input pixels) - concat is a choice of tensor map.
//   A operand : transform warps turn a raw slab into the zero-padded, post-activation bf16 operand ONCE per slab, in the
//               no-swizzle K-major core-matrix layout [k8 plane][halo pixel][8 channels = 16 B]; the 9 taps are 9
//               shared-memory descriptors with shifted start addresses over the same slab (im2col-free).  The transform
//               fuses GroupNorm-apply + Swish + padding mask + nearest-x2 replication + stride-2 parity split + bf16
//               conversion (zero padding is applied AFTER the activation, UNetModified2.py:116-121).
//   B operand : weights pre-packed on the host as [Cin/16][tap][2][Cout][8] bf16, moved per chunk with cp.async.bulk and
//               kept resident in shared memory for the whole persistent CTA when they fit (else streamed through a ring).
//   extra K   : the ResnetBlock 1x1 res_conv over the raw block input accumulates into the same TMEM tile.
//   epilogue  : tcgen05.ld -> + bias (+ noise-level embedding, + res bias) (+ identity residual, prefetched by TMA) ->
//               128B-swizzled shared-memory tile -> TMA tensor store (NHWC fp32) + GroupNorm partial statistics
//               (column sums of the staged tile: sum, sum of squares per channel per 32-pixel quarter).
//
// Warp roles (640 threads, persistent CTA, static tile schedule):
//   warps 0-3 / 4-7 epilogue groups 0 / 1 (group e drains accumulator stage e; TMEM lane quarter = warp id % 4)
//   warps 8-11 / 12-15 transform groups 0 / 1 (alternate slabs, so two slabs are converted concurrently; ring depths are even)
//   warp 16 MMA issuer + TMEM owner | warp 17 weight loader | warp 18 raw-slab TMA issuer | warp 19 idle (round 1's second issuer)
//
// reference: Block / ResnetBlock / Downsample / Upsample, model/UNetModified2.py:93-142
#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "gn_fuse.cuh"
#ifdef SDDM_TC_HANG_NOTES   // debug build (SDDM_NVCC_EXTRA="-DSDDM_TC_HANG_NOTES=1"): waits of this file that time out leave a note (sddm_debug_hang)
#define SDDM_MBAR_TIMEOUT_NOTES 1
#endif
#include "tc_ptx.cuh"
#include "../../include/sddm_b200.h"

namespace sddm {
namespace {

constexpr int TH = 16, TW = 8;            // output window of one tile
constexpr int kEpiGroups = 2, kEpiGroupThreads = 128;         // epilogue group e owns accumulator stage e
// The single-thread roles get the HIGHEST warp ids of their scheduler partition (warp id % 4): the issue arbiter favours
// high warp ids, and a starved MMA issuer stalls the whole pipeline.
constexpr int kXfWarp0 = 8, kXfGroups = 2, kXfGroupThreads = 128;
constexpr int kMmaWarp = 16, kWldWarp = 17, kTmaWarp = 18, kMmaWarp2 = 19;
constexpr int kThreads = 20 * 32;   // 640
constexpr int kMaxRing = 6, kMaxW = 32;
template <bool A16> struct OutTile { static constexpr uint32_t kBytes = 128u * 32u * (A16 ? 2u : 4u); };   // 128 pixels x 32 channels staging tile

// geometry per mode ---------------------------------------------------------------------------------------
//   SLAB          input channels per A slab (one raw TMA box, SLAB/16 MMA K steps per tap)
//   RAW_H x RAW_W raw box in INPUT pixels;  PH x PW  operand halo grid;  PLANE  slots (16 B) between k8 planes
//   SBO           byte distance between consecutive 8-pixel groups of the M dimension (one output row)
template <int MODE> struct Geo;
//   EVEN_OFF      (stride 2 only) slot of the first even input column within an operand row (odd columns start at 0)
template <> struct Geo<CONV_S1> { static constexpr int SLAB = 32, RAW_H = 18, RAW_W = 10, PH = 18, PW = 10, PLANE = 182, SBO = 10 * 16, EVEN_OFF = 0; };
template <> struct Geo<CONV_UP> { static constexpr int SLAB = 32, RAW_H = 10, RAW_W = 6, PH = 18, PW = 10, PLANE = 182, SBO = 10 * 16, EVEN_OFF = 0; };
template <> struct Geo<CONV_S2> { static constexpr int SLAB = 16, RAW_H = 33, RAW_W = 17, PH = 33, PW = 20, PLANE = 662, SBO = 2 * 20 * 16, EVEN_OFF = 12; };
// PLANE = 6 (mod 8) slots: a quarter warp storing 2 consecutive pixels x 4 planes (or, stride 2, 4 pixels x 2 planes with the
// even columns 12 slots after the odd ones) hits 8 distinct 16-byte bank groups.

struct alignas(64) TcMaps {
    CUtensorMap src[4];   // main source 0 / 1, res_conv source 0 / 1 (32-channel boxes: 128B swizzle)
    CUtensorMap out;      // output store, box 32 ch x 8 x 16, 128B swizzle
    CUtensorMap res;      // identity residual load, same geometry
};

struct TcArgs {
    ConvP p;
    int tiles_x, tiles_y, ntiles;   // per-sample tile grid, total tiles (B * tiles_x * tiles_y)
    unsigned long long magic_per, magic_tx;   // ceil(2^32 / d) for d = tiles per sample, tiles_x
    int n_main, n_res;              // A slabs of the main conv (SLAB channels) / the 1x1 res_conv (32 channels)
    int n_main_chunks, n_res_chunks;   // weight chunks: 16 channels x 3 taps (one filter row) / 32 channels x 1 tap
    int NR, NA, NW;                 // ring depths: raw slabs, operand slabs, weight chunks
    int NRES, NOUT;                 // per epilogue group: residual tiles, out staging tiles
    int resident;                   // all weight chunks stay in shared memory
    int acc_stride, tmem_cols;
    int temb_per_row;
    uint32_t off_out, off_res, off_raw, off_a, off_w;   // byte offsets from the 1024-aligned shared-memory base
    uint32_t raw_stage, a_stage, w_stage;
    // split-bf16 high-precision mode (SDDM_PREC_BF16X3): every operand is a bf16 (hi, lo) pair, x = hi + lo to 2^-17, and
    // a . w = a_hi w_hi + a_hi w_lo + a_lo w_hi on the tensor cores (fp32 accumulate); the lo images sit a_half / w_half bytes
    // behind the hi images inside each operand / weight stage
    int x3;
    uint32_t a_half, w_half;
    int skip;                       // debug experiments: bit 0 no statistics pass, bit 1 no transform work, bit 2 no epilogue staging / store
    long long* trace;               // debug: per-role wait / busy cycle counters of CTA 0 (nullptr = off)
};

struct SmemHdr {
    uint64_t raw_full[kMaxRing], raw_empty[kMaxRing];
    uint64_t full_a[kMaxRing], empty_a[kMaxRing];
    uint64_t full_w[kMaxW], empty_w[kMaxW];
    uint64_t tmem_full[2], tmem_empty[2];
    uint64_t res_full[kEpiGroups][4];
    uint32_t tmem_base;
    uint32_t gn_last[kEpiGroups];
    uint32_t pad[13];
    float addv[kEpiGroups][256];
    alignas(16) float red[kEpiGroups][4][32][2];   // per-warp (pixel quarter) column sums of one 32-channel block, combined to one partial per tile
};
constexpr uint32_t kHdrBytes = 5120;
static_assert(sizeof(SmemHdr) <= kHdrBytes, "header too large");

struct TileCoord { int n, oy0, ox0, trem; };
__device__ __forceinline__ TileCoord decode_tile(const TcArgs& a, int tile) {
    TileCoord t;
    t.n = (int)(((unsigned long long)(unsigned)tile * a.magic_per) >> 32);
    t.trem = tile - t.n * (a.tiles_x * a.tiles_y);
    const int ty = (int)(((unsigned long long)(unsigned)t.trem * a.magic_tx) >> 32);
    t.oy0 = ty * TH;
    t.ox0 = (t.trem - ty * a.tiles_x) * TW;
    return t;
}
template <int MODE> __device__ __forceinline__ int org_of(int o0) {   // first input row / column of the raw box
    return MODE == CONV_S2 ? 2 * o0 - 1 : (MODE == CONV_UP ? (o0 >> 1) - 1 : o0 - 1);
}

// =====================================================================================================
template <int MODE, int TPC, bool A16, bool X3>
__global__ void __launch_bounds__(kThreads, 1) conv3x3_tc_kernel(const TcArgs a, const __grid_constant__ TcMaps maps) {
    using G = Geo<MODE>;
    constexpr uint32_t kOutTileBytes = OutTile<A16>::kBytes;
    constexpr uint32_t ESZ = A16 ? 2u : 4u;   // bytes per stored activation
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t dyn_u32 = smem_u32(smem_dyn);
    const uint32_t base_u32 = (dyn_u32 + 1023u) & ~1023u;
    SmemHdr* hdr = reinterpret_cast<SmemHdr*>(smem_dyn + (base_u32 - dyn_u32));
    const ConvP& p = a.p;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nA = a.n_main + a.n_res;
    // contiguous tile run of this CTA (near-equal split): a sample is touched by few CTAs, so the GroupNorm finalisation of one
    // sample (gn_fuse.cuh) cannot make the same CTA the last arriver of every following sample, and neighbouring tiles share L2 lines
    const int tq = a.ntiles / (int)gridDim.x, trm = a.ntiles - tq * (int)gridDim.x;
    const int tile0 = (int)blockIdx.x * tq + ((int)blockIdx.x < trm ? (int)blockIdx.x : trm);
    const int tend = tile0 + tq + ((int)blockIdx.x < trm ? 1 : 0);
    pdl_launch_dependents();   // PDL: the next kernel's CTAs may be scheduled as soon as SMs free up

    if (tid == 0) {
        for (int i = 0; i < kMaxRing; ++i) {
            mbar_init(smem_u32(&hdr->raw_full[i]), 1); mbar_init(smem_u32(&hdr->raw_empty[i]), kXfGroupThreads / 32);
            mbar_init(smem_u32(&hdr->full_a[i]), kXfGroupThreads / 32); mbar_init(smem_u32(&hdr->empty_a[i]), 1);
        }
        for (int i = 0; i < kMaxW; ++i) { mbar_init(smem_u32(&hdr->full_w[i]), 1); mbar_init(smem_u32(&hdr->empty_w[i]), 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&hdr->tmem_full[i]), 1); mbar_init(smem_u32(&hdr->tmem_empty[i]), kEpiGroupThreads / 32);
            for (int k = 0; k < 4; ++k) mbar_init(smem_u32(&hdr->res_full[i][k]), 1);
        }
        fence_barrier_init();
    }
    if (warp == kMmaWarp) tmem_alloc(smem_u32(&hdr->tmem_base), (uint32_t)a.tmem_cols);
#ifdef SDDM_TC_HANG_NOTES
    if (g_hang && tid == 0 && blockIdx.x == 0) {   // the plan of the launch that is running when a wait times out
        unsigned long long gid;
        asm volatile("mov.u64 %0, %%gridid;" : "=l"(gid));
        unsigned* o = g_hang + 60000 + 8 * (unsigned)(gid & 3ull);
        o[0] = (unsigned)gid; o[1] = base_u32;
        o[2] = ((unsigned)a.NR << 24) | ((unsigned)a.NA << 16) | ((unsigned)a.NW << 8) | (unsigned)(a.resident ? 1 : 0);
        o[3] = 0x80000000u | ((unsigned)a.n_main << 20) | ((unsigned)a.n_res << 16) | ((unsigned)(tend - tile0) << 8) | (unsigned)(MODE * 64 + TPC * 4 + (A16 ? 2 : 0) + (X3 ? 1 : 0));
        o[4] = (unsigned)p.Cin; o[5] = (unsigned)p.Cout; o[6] = (unsigned)p.Hout; o[7] = gridDim.x;
        g_hang[1] = base_u32; g_hang[2] = o[2]; g_hang[3] = o[3];
    }
#endif
    if (warp < 8 && !a.temb_per_row) {   // per-channel additive term of the epilogue (same for every row of the batch)
        for (int c = tid & 127; c < p.Cout; c += 128) {
            float v = __ldg(p.bias + c);
            if (p.temb) v += __ldg(p.temb + c);
            if (a.n_res) v += __ldg(p.res_bias + c);
            hdr->addv[warp >> 2][c] = v;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = hdr->tmem_base;
    const bool tr = a.trace != nullptr && blockIdx.x == 0;   // trace: cycles CTA 0 spends in each kind of wait
    long long tw[6] = {0, 0, 0, 0, 0, 0};
    const long long t_begin = tr ? clock64() : 0;

    if (warp < 8) {
        // ============================== epilogue (group e = accumulator stage e) ==========================
        pdl_wait();   // everything this role reads (residual) or writes (output, statistics) is ordered after the predecessor
        const int e = warp >> 2, w4 = warp & 3;
        const int m = tid & 127;
        const int py = m >> 3, px = m & 7;
        const int bar_id = 1 + e;
        const bool leader = m == 0;
        const int nblk = p.Cout >> 5;
        const bool has_res = p.res_identity != 0;
        const uint32_t obuf0 = base_u32 + a.off_out + (uint32_t)(e * a.NOUT) * kOutTileBytes;
        const uint32_t rbuf0 = base_u32 + a.off_res + (uint32_t)(e * a.NRES) * kOutTileBytes;
        const uint32_t addv_u32 = smem_u32(hdr->addv[e]);
        // tiles of this group: it = e, e + 2, ...
        int my_tiles = 0;
        for (int it = e, tile = tile0 + e; tile < tend; tile += 2, it += 2) ++my_tiles;
        const int total_blocks = my_tiles * nblk;
        auto issue_res = [&](int g) {   // leader: prefetch the identity-residual tile of this group's block g
            const int k = g / nblk, cb = g - k * nblk;
            const TileCoord t = decode_tile(a, tile0 + (e + 2 * k));
            const int buf = g % a.NRES;
            const uint32_t bar = smem_u32(&hdr->res_full[e][buf]);
            mbar_expect_tx(bar, kOutTileBytes);
            tma_load_4d(rbuf0 + (uint32_t)buf * kOutTileBytes, &maps.res, cb * 32, t.ox0, t.oy0, t.n, bar);
        };
        if (has_res && leader)
            for (int g = 0; g < a.NRES - 1 && g < total_blocks; ++g) issue_res(g);
        // swizzled column offsets for the statistics pass: row r of the staged tile keeps channel c at
        // r * 128 + (((c >> 2) ^ (r & 7)) << 4) + (c & 3) * 4
        uint32_t coloff[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) coloff[k] = (uint32_t)((((lane >> 2) ^ k) << 4) + (lane & 3) * 4);
        // bf16 staging tile: 64-byte rows, 64B swizzle (16-byte chunk q of row r sits at chunk q ^ ((r >> 1) & 3)); in the
        // statistics pass a lane owns one 32-bit word (2 channels) and the two half-warps take even / odd rows
        const int c2 = lane & 15, hrow = lane >> 4;
        int g = 0;
        uint32_t aph = 0;
        // GroupNorm finalisation of the consumer (gn_fuse.cuh): the tiles of a CTA are contiguous, so this group counts its tiles of
        // the current sample and arrives ONCE when it moves on (an atomic round trip per tile would cost ~1.5 us each); the group
        // whose arrival completes the sample's tile count finalises it
        int gn_n = -1;
        unsigned gn_cnt = 0;
        auto gn_flush = [&]() {
            if (!p.gn_on || gn_cnt == 0) return;
            __threadfence();
            group_bar(bar_id);
            if (leader) hdr->gn_last[e] = atomicAdd(p.gn.counter + gn_n, gn_cnt) + gn_cnt == (unsigned)p.gn.expect ? 1u : 0u;
            group_bar(bar_id);
            if (hdr->gn_last[e]) {
                __threadfence();
                gn_fused_finalize(p.gn, gn_n, m, kEpiGroupThreads);
                if (leader) p.gn.counter[gn_n] = 0u;
            }
            gn_cnt = 0;
        };
        for (int tile = tile0 + e; tile < tend; tile += 2, aph ^= 1u) {
            const TileCoord t = decode_tile(a, tile);
            if (t.n != gn_n) { gn_flush(); gn_n = t.n; }
            const bool valid = (t.oy0 + py) < p.Hout && (t.ox0 + px) < p.Wout;
            if (a.temb_per_row) {   // explicit per-row noise levels (sddm_eps with a noise_level vector)
                group_bar(bar_id);
                for (int c = m; c < p.Cout; c += 128) {
                    float v = __ldg(p.bias + c) + __ldg(p.temb + (int64_t)t.n * p.temb_stride + c);
                    if (a.n_res) v += __ldg(p.res_bias + c);
                    hdr->addv[e][c] = v;
                }
                group_bar(bar_id);
            }
            mbar_wait_t(smem_u32(&hdr->tmem_full[e]), aph, tr, tw[0]);
            tc_fence_after();
            const uint32_t tacc = tmem_base + ((uint32_t)(w4 * 32) << 16) + (uint32_t)(e * a.acc_stride);
            for (int cb = 0; cb < nblk; ++cb, ++g) {
                const uint32_t obuf = obuf0 + (uint32_t)(a.NOUT == 2 ? (g & 1) : 0) * kOutTileBytes;
                const long long tb0 = tr ? clock64() : 0;
                if (leader) {   // the TMA store that last read this staging buffer is done with it
                    if (a.NOUT == 2) bulk_wait_read_1(); else bulk_wait_read_0();
                }
                if (tr) tw[2] += clock64() - tb0;
                group_bar(bar_id);
                if (tr) tw[3] += clock64() - tb0;
                if (has_res && leader && g + a.NRES - 1 < total_blocks) issue_res(g + a.NRES - 1);
                const long long te0 = tr ? clock64() : 0;
                float v[32];
                tmem_ld32(tacc + (uint32_t)(cb * 32), v);
                if (tr) tw[4] += clock64() - te0;
                if (cb == nblk - 1) {   // accumulator fully drained: hand the TMEM stage back to the MMA warp early
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&hdr->tmem_empty[e]));   // one arrival per warp: same-address arrivals serialise
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint4 u = lds128(addv_u32 + (uint32_t)(cb * 32 + q * 4) * 4u);
                    v[4 * q + 0] += __uint_as_float(u.x); v[4 * q + 1] += __uint_as_float(u.y);
                    v[4 * q + 2] += __uint_as_float(u.z); v[4 * q + 3] += __uint_as_float(u.w);
                }
                if (has_res) {
                    const int buf = g % a.NRES;
                    mbar_wait_t(smem_u32(&hdr->res_full[e][buf]), (uint32_t)(g / a.NRES) & 1u, tr, tw[1]);
                    if (A16) {
                        const uint32_t rrow = rbuf0 + (uint32_t)buf * kOutTileBytes + (uint32_t)m * 64u;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint4 u = lds128(rrow + (uint32_t)((q ^ ((m >> 1) & 3)) << 4));
                            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
                                v[8 * q + 2 * k] += f.x; v[8 * q + 2 * k + 1] += f.y;
                            }
                        }
                    } else {
                        const uint32_t rrow = rbuf0 + (uint32_t)buf * kOutTileBytes + (uint32_t)m * 128u;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const uint4 u = lds128(rrow + (uint32_t)((q ^ (m & 7)) << 4));
                            v[4 * q + 0] += __uint_as_float(u.x); v[4 * q + 1] += __uint_as_float(u.y);
                            v[4 * q + 2] += __uint_as_float(u.z); v[4 * q + 3] += __uint_as_float(u.w);
                        }
                    }
                }
                if (!valid) {
#pragma unroll
                    for (int k = 0; k < 32; ++k) v[k] = 0.f;
                }
                if (a.skip & 4) {
                } else if (A16) {
                    const uint32_t orow = obuf + (uint32_t)m * 64u;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 u;
                        u.x = pack_bf16(v[8 * q + 0], v[8 * q + 1]); u.y = pack_bf16(v[8 * q + 2], v[8 * q + 3]);
                        u.z = pack_bf16(v[8 * q + 4], v[8 * q + 5]); u.w = pack_bf16(v[8 * q + 6], v[8 * q + 7]);
                        sts128(orow + (uint32_t)((q ^ ((m >> 1) & 3)) << 4), u);
                    }
                } else {
                    const uint32_t orow = obuf + (uint32_t)m * 128u;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        uint4 u;
                        u.x = __float_as_uint(v[4 * q + 0]); u.y = __float_as_uint(v[4 * q + 1]);
                        u.z = __float_as_uint(v[4 * q + 2]); u.w = __float_as_uint(v[4 * q + 3]);
                        sts128(orow + (uint32_t)((q ^ (m & 7)) << 4), u);
                    }
                }
                fence_async_smem();
                group_bar(bar_id);
                if (tr) tw[5] += clock64() - te0;
                if (leader && !(a.skip & 4)) {
                    tma_store_4d(&maps.out, obuf, cb * 32, t.ox0, t.oy0, t.n);
                    bulk_commit();
                }
                if (a.skip & 1) {
                } else if (p.parts && A16) {   // column sums of the staged (rounded) tile: 2 channels per lane, even / odd rows per half-warp
                    float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
                    const uint32_t qbase = obuf + (uint32_t)(w4 * 32 + hrow) * 64u + (uint32_t)(c2 & 3) * 4u;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {   // row = w4 * 32 + 2 i + hrow; (row >> 1) & 3 = i & 3
                        uint32_t wv;
                        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(wv) : "r"(qbase + (uint32_t)i * 128u + (uint32_t)(((c2 >> 2) ^ (i & 3)) << 4)));
                        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wv));
                        s1a += f.x; s1b += f.y;
                        s2a = fmaf(f.x, f.x, s2a); s2b = fmaf(f.y, f.y, s2b);
                    }
                    s1a += __shfl_xor_sync(0xffffffffu, s1a, 16); s1b += __shfl_xor_sync(0xffffffffu, s1b, 16);
                    s2a += __shfl_xor_sync(0xffffffffu, s2a, 16); s2b += __shfl_xor_sync(0xffffffffu, s2b, 16);
                    if (hrow == 0) *reinterpret_cast<float4*>(&hdr->red[e][w4][2 * c2][0]) = make_float4(s1a, s2a, s1b, s2b);
                } else if (p.parts) {   // column sums of the staged tile: channel = lane, pixel quarter = warp
                    float s1 = 0.f, s2 = 0.f;
                    const uint32_t qbase = obuf + (uint32_t)(w4 * 32) * 128u;
#pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        const float x = lds32(qbase + (uint32_t)r * 128u + coloff[r & 7]);
                        s1 += x;
                        s2 = fmaf(x, x, s2);
                    }
                    *reinterpret_cast<float2*>(&hdr->red[e][w4][lane][0]) = make_float2(s1, s2);
                }
                if (p.parts && !(a.skip & 1)) {   // one partial per tile: the four pixel quarters are added in a fixed order
                    group_bar(bar_id);
                    if (m < 32) {
                        float2 acc = *reinterpret_cast<const float2*>(&hdr->red[e][0][m][0]);
#pragma unroll
                        for (int qq = 1; qq < 4; ++qq) {
                            const float2 v2 = *reinterpret_cast<const float2*>(&hdr->red[e][qq][m][0]);
                            acc.x += v2.x; acc.y += v2.y;
                        }
                        reinterpret_cast<float2*>(p.parts)[((int64_t)t.n * p.nparts + t.trem) * p.Cout + cb * 32 + m] = acc;
                    }
                }
            }
            ++gn_cnt;
        }
        gn_flush();
        if (leader) bulk_wait_all();
        if (tr && leader) {
            long long* o = a.trace + (e == 0 ? 0 : 8);
            o[0] = clock64() - t_begin; o[1] = tw[0]; o[2] = tw[1]; o[3] = tw[2]; o[4] = tw[3]; o[5] = my_tiles; o[6] = tw[4]; o[7] = tw[5];
        }
    } else if (warp == kMmaWarp || warp == kMmaWarp2) {
        // ============================== MMA issuers =======================================================
        // With resident weights two warps issue: warp mw owns the tiles it = mw, mw + 2, ... (= accumulator stage mw), so one
        // warp's barrier round trips overlap the other's MMAs and the tensor pipe stays fed.  With streamed weights (one
        // ordered chunk ring) only the first warp works.
        // The whole warp runs the (warp-uniform) control flow so that descriptors live in uniform registers; one elected
        // lane issues the tcgen05.mma / tcgen05.commit instructions.
        constexpr int KS = G::SLAB / 16;
        constexpr int CPK = 9 / TPC;                         // weight chunks per 16-channel K step
        const uint32_t idesc = make_idesc(p.Cout);
        const uint32_t b_lbo = (uint32_t)p.Cout * 16u, b_sbo = 128u;
        const uint32_t a_lbo = (uint32_t)G::PLANE * 16u, a_sbo = (uint32_t)G::SBO;
        const uint64_t a_desc0 = make_desc_nosw(base_u32 + a.off_a, a_lbo, a_sbo);
        const uint64_t w_desc0 = make_desc_nosw(base_u32 + a.off_w, b_lbo, b_sbo);
        const uint32_t a_step = a.a_stage >> 4, w_step = a.w_stage >> 4, tap_step = 2u * (uint32_t)p.Cout;   // in 16-byte units
        // ONE issuing warp.  Two warps on alternate tiles (the round-1 design, enabled when the weights were resident and NA >= nA + 1)
        // are not safe: mbarrier waits only see the phase PARITY, and a warp that waits on operand stage s for use k needs use k - 1
        // of s to be filled already.  The other warp's slabs do not give it that: the two transform groups fill alternate slabs
        // independently, so slab 3 (group 1) can be ready while slab 0 (group 0, same stage as slab 4) is not - the wait for slab 4
        // then passes on the phase before slab 0, the warp multiplies an unfilled stage and releases it, and the pipeline wedges
        // (seen as a timed-out wait, once in ~10 full-size runs of the fp32-activation modes; tools/diag_hang.py --precision bf16
        // --infer on a build with -DSDDM_TC_HANG_NOTES=1).  A single warp consumes in order, so its own previous wait on the stage
        // orders the phases.  (To get the second warp back: one full_a barrier set per issuing warp.)
        const int mw = warp == kMmaWarp ? 0 : 1, nmma = 1;
        int sa = 0, sw = 0;
        uint32_t pa = 0, pw = 0;
        auto skip_slabs = [&](int n) { sa += n; while (sa >= a.NA) { sa -= a.NA; pa ^= 1u; } };
        bool w_ready = false;   // resident weights: all chunks have landed (after this warp's first tile)
        if (mw < nmma) skip_slabs(mw * nA);
        for (int it = mw, tile = tile0 + mw; mw < nmma && tile < tend; tile += nmma, it += nmma) {
            const int as = it & 1;
            const uint32_t aph = (uint32_t)(it >> 1) & 1u;
            mbar_wait_t(smem_u32(&hdr->tmem_empty[as]), aph ^ 1u, tr, tw[0]);
            const uint32_t d_tmem = tmem_base + (uint32_t)(as * a.acc_stride);
            uint32_t acc = 0;
            for (int ai = 0; ai < a.n_main; ++ai) {
                mbar_wait_t(smem_u32(&hdr->full_a[sa]), pa, tr, tw[1]);
                tc_fence_after();
                const uint64_t adesc = a_desc0 + (uint64_t)((uint32_t)sa * a_step);
#pragma unroll
                for (int h = 0; h < KS; ++h) {
#pragma unroll
                    for (int q = 0; q < CPK; ++q) {
                        const int c = (ai * KS + h) * CPK + q;
                        int wslot;
                        if (a.resident) {
                            wslot = c;
                            if (!w_ready) { mbar_wait_t(smem_u32(&hdr->full_w[c]), 0u, tr, tw[2]); tc_fence_after(); }
                        } else {
                            wslot = sw;
                            mbar_wait_t(smem_u32(&hdr->full_w[sw]), pw, tr, tw[2]);
                            tc_fence_after();
                        }
                        const uint64_t wdesc = w_desc0 + (uint64_t)((uint32_t)wslot * w_step);
                        const long long tm0 = tr ? clock64() : 0;
                        if (elect_one()) {
#pragma unroll
                            for (int tt = 0; tt < TPC; ++tt) {
                                const int tap = q * TPC + tt, ky = tap / 3, kx = tap - 3 * ky;
                                const int aslot = (MODE == CONV_S2) ? ky * G::PW + (kx == 1 ? G::EVEN_OFF : (kx >> 1)) : ky * G::PW + kx;
                                const uint64_t ad = adesc + (uint64_t)(uint32_t)(h * 2 * G::PLANE + aslot), wd = wdesc + (uint64_t)((uint32_t)tt * tap_step);
                                umma(d_tmem, ad, wd, idesc, acc);
                                acc = 1;
                                if (X3) {
                                    umma(d_tmem, ad, wd + (uint64_t)(a.w_half >> 4), idesc, 1u);
                                    umma(d_tmem, ad + (uint64_t)(a.a_half >> 4), wd, idesc, 1u);
                                }
                            }
                            if (!a.resident) umma_commit(smem_u32(&hdr->empty_w[sw]));
                        }
                        if (tr) tw[3] += clock64() - tm0;
                        acc = 1;
                        if (!a.resident && ++sw == a.NW) { sw = 0; pw ^= 1u; }
                    }
                }
                const long long tc0 = tr ? clock64() : 0;
                if (elect_one()) umma_commit(smem_u32(&hdr->empty_a[sa]));
                if (tr) tw[4] += clock64() - tc0;
                if (++sa == a.NA) { sa = 0; pa ^= 1u; }
            }
            // 1x1 res_conv over the raw block input: centre tap of a stride-1 halo slab, one weight chunk per 32-channel slab
            for (int ar = 0; ar < a.n_res; ++ar) {
                mbar_wait_t(smem_u32(&hdr->full_a[sa]), pa, tr, tw[1]);
                tc_fence_after();
                const uint64_t adesc = a_desc0 + (uint64_t)((uint32_t)sa * a_step + (uint32_t)(G::PW + 1));
                const int c = a.n_main_chunks + ar;
                int wslot;
                if (a.resident) {
                    wslot = c;
                    if (!w_ready) { mbar_wait_t(smem_u32(&hdr->full_w[c]), 0u, tr, tw[2]); tc_fence_after(); }
                } else {
                    wslot = sw;
                    mbar_wait_t(smem_u32(&hdr->full_w[sw]), pw, tr, tw[2]);
                    tc_fence_after();
                }
                const uint64_t wdesc = w_desc0 + (uint64_t)((uint32_t)wslot * w_step);
                if (elect_one()) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint64_t ad = adesc + (uint64_t)(uint32_t)(h * 2 * G::PLANE), wd = wdesc + (uint64_t)((uint32_t)h * tap_step);
                        umma(d_tmem, ad, wd, idesc, acc);
                        acc = 1;
                        if (X3) {
                            umma(d_tmem, ad, wd + (uint64_t)(a.w_half >> 4), idesc, 1u);
                            umma(d_tmem, ad + (uint64_t)(a.a_half >> 4), wd, idesc, 1u);
                        }
                    }
                    if (!a.resident) umma_commit(smem_u32(&hdr->empty_w[sw]));
                    umma_commit(smem_u32(&hdr->empty_a[sa]));
                }
                acc = 1;
                if (!a.resident && ++sw == a.NW) { sw = 0; pw ^= 1u; }
                if (++sa == a.NA) { sa = 0; pa ^= 1u; }
            }
            if (elect_one()) umma_commit(smem_u32(&hdr->tmem_full[as]));
            w_ready = true;
            skip_slabs((nmma - 1) * nA);   // the other warp's tile
        }
        if (tr && lane == 0 && mw == 0) { long long* o = a.trace + 16; o[0] = clock64() - t_begin; o[1] = tw[0]; o[2] = tw[1]; o[3] = tw[2]; o[4] = tw[3]; o[5] = tw[4]; }
    } else if (warp == kWldWarp) {
        // ============================== weight loader (one thread) ========================================
        if (lane == 0) {
            const int nchunks = a.n_main_chunks + a.n_res_chunks;
            const uint32_t w_ring = base_u32 + a.off_w;
            int sw = 0;
            uint32_t pw = 0;
            for (int it = 0, tile = tile0; tile < tend; ++tile, ++it) {
                if (a.resident && it > 0) break;
                for (int c = 0; c < nchunks; ++c) {
                    const bool is_res = c >= a.n_main_chunks;
                    const int wslot = a.resident ? c : sw;
                    if (!a.resident) mbar_wait(smem_u32(&hdr->empty_w[sw]), pw ^ 1u);
                    uint32_t bytes;
                    const void* src;
                    if (!is_res) {   // packed [Cin/16][tap][2][Cout][8]: filter row ky of chunk c16 is contiguous: TPC taps = 32*TPC*Cout bytes
                        bytes = 32u * TPC * (uint32_t)p.Cout;
                        src = p.w_tc + (size_t)c * 16 * TPC * p.Cout;
                    } else {
                        bytes = 64u * (uint32_t)p.Cout;
                        src = p.res_w_tc + (size_t)(c - a.n_main_chunks) * 32 * p.Cout;
                    }
                    const uint32_t bar = smem_u32(&hdr->full_w[wslot]);
                    mbar_expect_tx(bar, X3 ? 2u * bytes : bytes);
                    bulk_g2s(w_ring + (uint32_t)wslot * a.w_stage, src, bytes, bar);
                    if (X3) {   // the lo image of the same chunk
                        const void* src_lo = !is_res ? (const void*)(p.w_tc_lo + (size_t)c * 16 * TPC * p.Cout)
                                                     : (const void*)(p.res_w_tc_lo + (size_t)(c - a.n_main_chunks) * 32 * p.Cout);
                        bulk_g2s(w_ring + (uint32_t)wslot * a.w_stage + a.w_half, src_lo, bytes, bar);
                    }
                    if (!a.resident && ++sw == a.NW) { sw = 0; pw ^= 1u; }
                }
            }
        }
    } else if (warp == kTmaWarp) {
        // ============================== raw-slab TMA issuer (one thread) ==================================
        if (lane == 0) {
            pdl_wait();   // activations and GroupNorm scale / shift come from the preceding kernels
            constexpr uint32_t RAW_BYTES = (uint32_t)G::RAW_H * G::RAW_W * G::SLAB * ESZ;
            const uint32_t raw_ring = base_u32 + a.off_raw;
            int rs = 0;
            uint32_t pr = 0;
            for (int tile = tile0; tile < tend; ++tile) {
                const TileCoord t = decode_tile(a, tile);
                const int yo = org_of<MODE>(t.oy0), xo = org_of<MODE>(t.ox0);
                for (int ai = 0; ai < nA; ++ai) {
                    const bool is_res = ai >= a.n_main;
                    const int cbase = is_res ? (ai - a.n_main) * 32 : ai * G::SLAB;
                    const ConvSrc* srcs = is_res ? p.res_src : p.src;
                    const int s = (cbase < srcs[0].C) ? 0 : 1;
                    const int coff = cbase - (s ? srcs[0].C : 0);
                    const uint32_t bar = smem_u32(&hdr->raw_full[rs]);
                    mbar_wait_t(smem_u32(&hdr->raw_empty[rs]), pr ^ 1u, tr, tw[0]);
                    const bool affine = !is_res && srcs[s].scale != nullptr;
                    const uint32_t nch = is_res ? 32u : (uint32_t)G::SLAB;
                    mbar_expect_tx(bar, (is_res ? 18u * 10u * 32u * ESZ : RAW_BYTES) + (affine ? 8u * nch : 0u));
                    const uint32_t dst = raw_ring + (uint32_t)rs * a.raw_stage;
                    tma_load_4d(dst, &maps.src[(is_res ? 2 : 0) + s], coff, is_res ? t.ox0 - 1 : xo, is_res ? t.oy0 - 1 : yo, t.n, bar);
                    if (affine) {   // GroupNorm scale / shift of the slab's channels travel with the slab
                        bulk_g2s(dst + a.raw_stage - 256u, srcs[s].scale + (int64_t)t.n * p.Cin + cbase, 4u * nch, bar);
                        bulk_g2s(dst + a.raw_stage - 128u, srcs[s].shift + (int64_t)t.n * p.Cin + cbase, 4u * nch, bar);
                    }
                    if (++rs == a.NR) { rs = 0; pr ^= 1u; }
                }
            }
            if (tr) { long long* o = a.trace + 24; o[0] = clock64() - t_begin; o[1] = tw[0]; }
        }
    } else if (warp >= kXfWarp0 && warp < kXfWarp0 + kXfGroups * 4) {
        // ============================== transform: raw fp32 slab -> bf16 operand slab ====================
        // group gi handles the slabs gi, gi + kXfGroups, ... of the CTA's (tile, slab) sequence, so kXfGroups slabs are
        // in flight at once.  32-channel raw slabs are 128B-swizzled by the TMA (16-byte chunk c of pixel p sits at
        // chunk c ^ (p & 7)): reading chunks 2j and 2j+1 of consecutive pixels is bank-conflict free.
        constexpr int NPL = G::SLAB / 8;                       // k8 planes per slab
        const int gi = (warp - kXfWarp0) >> 2;
        const int gt = tid - (kXfWarp0 * 32 + gi * kXfGroupThreads);   // thread within the group
        const uint32_t raw_ring = base_u32 + a.off_raw, a_ring = base_u32 + a.off_a;

        // one 8-channel item: raw fp32 -> (affine, swish) -> masked -> packed bf16.   swish(y) = h + h * tanh(h), h = y / 2
        // raw loads: fp32 storage = two 16-byte chunks per item (u0, u1); bf16 storage = one chunk (u0) holding all 8 channels
        constexpr bool x3 = X3;
        uint4 o_lo = make_uint4(0u, 0u, 0u, 0u);   // lo halves of the last converted item (split-bf16 mode)
        auto convert = [&](auto aff_c, const uint4 u0, const uint4 u1, bool ok, const float (&sch)[8], const float (&shh)[8]) -> uint4 {
            constexpr bool AFF = decltype(aff_c)::value;
            float f[8];
            if (A16) {
                const uint32_t w[4] = {u0.x, u0.y, u0.z, u0.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 t2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
                    f[2 * k] = t2.x; f[2 * k + 1] = t2.y;
                }
            } else {
                f[0] = __uint_as_float(u0.x); f[1] = __uint_as_float(u0.y); f[2] = __uint_as_float(u0.z); f[3] = __uint_as_float(u0.w);
                f[4] = __uint_as_float(u1.x); f[5] = __uint_as_float(u1.y); f[6] = __uint_as_float(u1.z); f[7] = __uint_as_float(u1.w);
            }
            if (AFF) {
                if (x3) {   // high-precision Swish: y / (1 + 2^(-y log2 e)) with ex2.approx / rcp.approx (~1e-6), not tanh.approx (~5e-4)
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float y = 2.0f * fmaf(f[k], sch[k], shh[k]);
                        f[k] = __fdividef(y, 1.0f + __expf(-y));
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 h = ffma2(make_float2(f[2 * k], f[2 * k + 1]), make_float2(sch[2 * k], sch[2 * k + 1]), make_float2(shh[2 * k], shh[2 * k + 1]));
                        const float2 sw = ffma2(h, make_float2(tanh_approx(h.x), tanh_approx(h.y)), h);
                        f[2 * k] = sw.x; f[2 * k + 1] = sw.y;
                    }
                }
            }
            uint4 o;
            o.x = ok ? pack_bf16(f[0], f[1]) : 0u; o.y = ok ? pack_bf16(f[2], f[3]) : 0u;
            o.z = ok ? pack_bf16(f[4], f[5]) : 0u; o.w = ok ? pack_bf16(f[6], f[7]) : 0u;
            if (x3) {   // lo = bf16(x - hi): x = hi + lo to 2^-17
                const uint32_t hw[4] = {o.x, o.y, o.z, o.w};
                uint32_t lw[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 h2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hw[k]));
                    lw[k] = ok ? pack_bf16(f[2 * k] - h2.x, f[2 * k + 1] - h2.y) : 0u;
                }
                o_lo = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
            return o;
        };
        // operand store: the hi image, and in split-bf16 mode the lo image a_half bytes behind it
        auto put = [&](uint32_t addr, const uint4 o) {
            sts128(addr, o);
            if (x3) sts128(addr + a.a_half, o_lo);
        };

        // this group's position in the CTA-wide (tile, slab) sequence and in the rings; it advances kXfGroups slabs at a time
        int tile = tile0, ai = gi;
        while (ai >= nA) { ai -= nA; tile += 1; }
        int rs = gi % a.NR, sa = gi % a.NA;
        uint32_t pr = (uint32_t)(gi / a.NR) & 1u, pa = (uint32_t)(gi / a.NA) & 1u;
        int cur_tile = -1;
        TileCoord t{0, 0, 0, 0};
        while (tile < tend) {
            const long long ts0 = tr ? clock64() : 0;
            if (tile != cur_tile) { t = decode_tile(a, tile); cur_tile = tile; }
            const bool is_res = ai >= a.n_main;
            const bool s1_like = is_res || MODE == CONV_S1;
            const int npl = s1_like ? 4 : NPL;
            const int j = gt % npl, pix0 = gt / npl;
            const int cbase = is_res ? (ai - a.n_main) * 32 : ai * G::SLAB;
            const ConvSrc* srcs = is_res ? p.res_src : p.src;
            const int s = (cbase < srcs[0].C) ? 0 : 1;
            const bool affine = !is_res && srcs[s].scale != nullptr;
            // valid window of the raw box (in box coordinates): everything else is zero padding
            int yo, xo, rh, rw;
            if (s1_like) { yo = t.oy0 - 1; xo = t.ox0 - 1; rh = 18; rw = 10; }
            else { yo = org_of<MODE>(t.oy0); xo = org_of<MODE>(t.ox0); rh = G::RAW_H; rw = G::RAW_W; }
            const int ylo = yo < 0 ? -yo : 0, xlo = xo < 0 ? -xo : 0;
            const int yhi = (p.Hin - yo) < rh ? (p.Hin - yo) : rh, xhi = (p.Win - xo) < rw ? (p.Win - xo) : rw;
            const uint32_t raw = raw_ring + (uint32_t)rs * a.raw_stage;
            const uint32_t opd = a_ring + (uint32_t)sa * a.a_stage + (uint32_t)j * (uint32_t)G::PLANE * 16u;
            if (tr) tw[2] += clock64() - ts0;
            mbar_wait_t(smem_u32(&hdr->raw_full[rs]), pr, tr, tw[0]);
            mbar_wait_t(smem_u32(&hdr->empty_a[sa]), pa ^ 1u, tr, tw[1]);
            const long long ts1 = tr ? clock64() : 0;
            // per-channel affine of this thread's 8 channels, halved: swish(y) = h + h * tanh(h) with h = y / 2.
            // (scale, shift) of the slab were bulk-copied behind the raw box by the TMA warp.
            float sch[8], shh[8];
            if (affine) {
                const uint32_t ss = raw + a.raw_stage - 256u + (uint32_t)j * 32u;
                const uint4 s0 = lds128(ss), s1 = lds128(ss + 16u), h0 = lds128(ss + 128u), h1 = lds128(ss + 144u);
                sch[0] = 0.5f * __uint_as_float(s0.x); sch[1] = 0.5f * __uint_as_float(s0.y); sch[2] = 0.5f * __uint_as_float(s0.z); sch[3] = 0.5f * __uint_as_float(s0.w);
                sch[4] = 0.5f * __uint_as_float(s1.x); sch[5] = 0.5f * __uint_as_float(s1.y); sch[6] = 0.5f * __uint_as_float(s1.z); sch[7] = 0.5f * __uint_as_float(s1.w);
                shh[0] = 0.5f * __uint_as_float(h0.x); shh[1] = 0.5f * __uint_as_float(h0.y); shh[2] = 0.5f * __uint_as_float(h0.z); shh[3] = 0.5f * __uint_as_float(h0.w);
                shh[4] = 0.5f * __uint_as_float(h1.x); shh[5] = 0.5f * __uint_as_float(h1.y); shh[6] = 0.5f * __uint_as_float(h1.z); shh[7] = 0.5f * __uint_as_float(h1.w);
            }
            // Branch-free batches: every thread handles a (clamped) pixel in every round - the surplus lanes of the last round
            // redo the last pixel and store identical values - so the shared-memory loads of a batch issue back to back
            // and the per-element chains of its items interleave.
            auto run_slab = [&](auto aff_c) {
                if (MODE == CONV_UP && !is_res) {
                    constexpr int NPIX = G::RAW_H * G::RAW_W, PST = kXfGroupThreads / 4, ROUNDS = (NPIX + PST - 1) / PST;
                    uint4 rv[ROUNDS][2];
                    int pv[ROUNDS];
#pragma unroll
                    for (int r = 0; r < ROUNDS; ++r) {
                        const int pix = min(pix0 + r * PST, NPIX - 1);
                        pv[r] = pix;
                        if (A16) {
                            rv[r][0] = lds128(raw + (uint32_t)pix * 64u + (uint32_t)j * 16u); rv[r][1] = make_uint4(0u, 0u, 0u, 0u);
                        } else {
                            const uint32_t ra = raw + (uint32_t)pix * 128u + (uint32_t)(((2 * j) ^ (pix & 7)) << 4);
                            rv[r][0] = lds128(ra); rv[r][1] = lds128(ra ^ 16u);
                        }
                    }
#pragma unroll
                    for (int r = 0; r < ROUNDS; ++r) {
                        const int pix = pv[r];
                        const int ry = pix / G::RAW_W, rx = pix - ry * G::RAW_W;
                        const bool ok = ry >= ylo && ry < yhi && rx >= xlo && rx < xhi;
                        const uint4 o = convert(aff_c, rv[r][0], rv[r][1], ok, sch, shh);
#pragma unroll
                        for (int dy = 0; dy < 2; ++dy) {
                            const int hy = 2 * ry - 1 + dy;
                            if (hy < 0 || hy >= G::PH) continue;
#pragma unroll
                            for (int dx = 0; dx < 2; ++dx) {
                                const int hx = 2 * rx - 1 + dx;
                                if (hx < 0 || hx >= G::PW) continue;
                                put(opd + (uint32_t)(hy * G::PW + hx) * 16u, o);
                            }
                        }
                    }
                } else if (MODE == CONV_S2) {
                    constexpr int NPIX = G::RAW_H * G::RAW_W, PST = kXfGroupThreads / NPL, ROUNDS = (NPIX + PST - 1) / PST, RB = 3;
                    static_assert(MODE != CONV_S2 || ROUNDS % RB == 0, "stride-2 rounds come in batches");
#pragma unroll 1
                    for (int rb = 0; rb < ROUNDS / RB; ++rb) {
                        uint4 rv[RB][2];
                        int pv[RB];
#pragma unroll
                        for (int q = 0; q < RB; ++q) {
                            const int pix = min(pix0 + (rb * RB + q) * PST, NPIX - 1);
                            pv[q] = pix;
                            if (A16) {
                                rv[q][0] = lds128(raw + (uint32_t)pix * 32u + (uint32_t)j * 16u); rv[q][1] = make_uint4(0u, 0u, 0u, 0u);
                            } else {
                                const uint32_t ra = raw + (uint32_t)pix * 64u + (uint32_t)j * 32u;   // unswizzled 64-byte pixels
                                const bool swap = (pix >> 1) & 1;
                                const uint4 lo = lds128(ra + (swap ? 16u : 0u)), hi = lds128(ra + (swap ? 0u : 16u));
                                rv[q][0] = swap ? hi : lo; rv[q][1] = swap ? lo : hi;
                            }
                        }
#pragma unroll
                        for (int q = 0; q < RB; ++q) {
                            const int pix = pv[q];
                            const int hy = pix / G::RAW_W, hx = pix - hy * G::RAW_W;
                            const bool ok = hy >= ylo && hy < yhi && hx >= xlo && hx < xhi;
                            const uint4 o = convert(aff_c, rv[q][0], rv[q][1], ok, sch, shh);
                            const int slot = hy * G::PW + ((hx & 1) ? G::EVEN_OFF + (hx >> 1) : (hx >> 1));
                            put(opd + (uint32_t)slot * 16u, o);
                        }
                    }
                } else {   // stride-1 halo (main conv of CONV_S1, res_conv slabs)
                    constexpr int NPIX = 18 * 10, PST = kXfGroupThreads / 4, ROUNDS = (NPIX + PST - 1) / PST, RB = 3;
                    static_assert(ROUNDS % RB == 0, "stride-1 rounds come in batches");
#pragma unroll
                    for (int rb = 0; rb < ROUNDS / RB; ++rb) {
                        uint4 rv[RB][2];
                        int pv[RB];
#pragma unroll
                        for (int q = 0; q < RB; ++q) {
                            const int pix = min(pix0 + (rb * RB + q) * PST, NPIX - 1);
                            pv[q] = pix;
                            if (A16) {
                                rv[q][0] = lds128(raw + (uint32_t)pix * 64u + (uint32_t)j * 16u); rv[q][1] = make_uint4(0u, 0u, 0u, 0u);
                            } else {
                                const uint32_t ra = raw + (uint32_t)pix * 128u + (uint32_t)(((2 * j) ^ (pix & 7)) << 4);
                                rv[q][0] = lds128(ra); rv[q][1] = lds128(ra ^ 16u);
                            }
                        }
#pragma unroll
                        for (int q = 0; q < RB; ++q) {
                            const int pix = pv[q];
                            const int hy = pix / 10, hx = pix - hy * 10;
                            const bool ok = hy >= ylo && hy < yhi && hx >= xlo && hx < xhi;
                            const uint4 o = convert(aff_c, rv[q][0], rv[q][1], ok, sch, shh);
                            put(opd + (uint32_t)pix * 16u, o);
                        }
                    }
                }
            };
            if (!(a.skip & 2)) {
                if (affine) run_slab(std::true_type{}); else run_slab(std::false_type{});
            }
            const long long ts2 = tr ? clock64() : 0;
            fence_async_smem();
            __syncwarp();   // one arrival per warp (every lane has fenced its own operand stores above)
            if (lane == 0) {
                mbar_arrive(smem_u32(&hdr->full_a[sa]));
                mbar_arrive(smem_u32(&hdr->raw_empty[rs]));
            }
            if (tr) { tw[3] += ts2 - ts1; tw[4] += clock64() - ts2; }
            // advance by kXfGroups slabs
            ai += kXfGroups;
            while (ai >= nA) { ai -= nA; tile += 1; }
            rs += kXfGroups; if (rs >= a.NR) { rs -= a.NR; pr ^= 1u; }
            sa += kXfGroups; if (sa >= a.NA) { sa -= a.NA; pa ^= 1u; }
        }
        if (tr && gt == 0) { long long* o = a.trace + 32 + 8 * gi; o[0] = clock64() - t_begin; o[1] = tw[0]; o[2] = tw[1]; o[3] = tw[2]; o[4] = tw[3]; o[5] = tw[4]; }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
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
            const uint64_t da = swap_fields ? make_desc_nosw(aa, a_sbo, a_lbo) : make_desc_nosw(aa, a_lbo, a_sbo);
            const uint64_t db = swap_fields ? make_desc_nosw(bb, b_sbo, b_lbo) : make_desc_nosw(bb, b_lbo, b_sbo);
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

// issue-rate microbenchmark: `reps` back-to-back 128 x N x 16 MMAs, one CTA.  geo 0: canonical A tiles (SBO 128 B, 128-byte
// aligned core matrices), cycling over nA tiles; geo 1: the stride-1 halo geometry of the conv kernel (SBO 160 B, LBO 182*16 B),
// cycling over the 9 tap start offsets; geo 2: x-shifted dense copies (SBO 128 B, aligned), cycling over 9 (copy, row) offsets.
__global__ void __launch_bounds__(128) umma_rate_kernel(int N, int reps, int nA, int geo, long long* cycles) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t sA = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_region = 64u * 1024u, b_off = a_region;
    for (uint32_t i = tid; i < (a_region + (uint32_t)N * 32u + 1024u) / 4u; i += 128) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&tbase), 256);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tbase;
    if (warp == 0) {
        const uint32_t idesc = make_idesc(N);
        const uint64_t db = make_desc_nosw(sA + b_off, (uint32_t)N * 16u, 128);
        const uint64_t da0 = geo == 1 ? make_desc_nosw(sA, 182u * 16u, 160u) : (geo == 2 ? make_desc_nosw(sA, 2304u + 128u, 128u) : make_desc_nosw(sA, 2048, 128));
        long long t0 = 0;
        if (elect_one()) {
            t0 = clock64();
            // 9 MMAs per iteration with compile-time start offsets (16-byte units): the issue loop itself must not be the limiter
            uint32_t acc = 0;
            for (int r = 0; r < reps / 9; ++r) {
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const uint32_t o1 = (uint32_t)((tap / 3) * 10 + tap % 3), o2 = (uint32_t)((tap % 3) * 640 + (tap / 3) * 8), o0 = (uint32_t)tap * 256u;
                    umma(tmem, da0 + (uint64_t)(geo == 1 ? o1 : (geo == 2 ? o2 : (nA > 1 ? o0 : 0u))), db, idesc, acc);
                    acc = 1;
                }
            }
            umma_commit(smem_u32(&bar));
        }
        mbar_wait(smem_u32(&bar), 0);
        if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { __syncwarp(); tmem_dealloc(tmem, 256); }
}


long long* g_trace = nullptr;   // device buffer [64 launches][48 counters], set by sddm_debug_tc_trace
int g_trace_launch = 0;

constexpr size_t kSmemMax = 232448;   // 227 KB opt-in limit per CTA on sm_100

// NHWC tensor [B][H][W][C] (fp32, or bf16 when a16) as a 4-D tensor map (C innermost) with box (bc, bw, bh, 1);
// swz: 0 none, 64 / 128 = CU_TENSOR_MAP_SWIZZLE_64B / 128B
int encode_nhwc(CUtensorMap* m, const float* base, int B, int H, int W, int C, int bc, int bw, int bh, int swz, bool a16) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("conv tc: cuTensorMapEncodeTiled is unavailable"); return SDDM_E_CUDA; }
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t es = a16 ? 2 : 4;
    const cuuint64_t strides[3] = {(cuuint64_t)C * es, (cuuint64_t)W * C * es, (cuuint64_t)H * W * C * es};
    const cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = fn(m, a16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv tc: cuTensorMapEncodeTiled failed (%d) for [%d,%d,%d,%d] box %dx%dx%d", (int)r, B, H, W, C, bc, bw, bh); return SDDM_E_CUDA; }
    return SDDM_OK;
}

inline size_t align_up_sz(size_t v, size_t a) { return (v + a - 1) / a * a; }

// returns 1 when the shared-memory plan does not fit with TPC taps per weight chunk (the caller retries with smaller chunks)
template <int MODE, int TPC, bool A16, bool X3>
int launch_mode(TcArgs a, cudaStream_t st) {
    using G = Geo<MODE>;
    constexpr uint32_t kOutTileBytes = OutTile<A16>::kBytes;
    constexpr size_t ESZ = A16 ? 2 : 4;
    const ConvP& p = a.p;
    a.n_main = p.Cin / G::SLAB;
    a.n_main_chunks = (9 / TPC) * (p.Cin / 16);
    a.raw_stage = (uint32_t)align_up_sz((size_t)G::RAW_H * G::RAW_W * G::SLAB * ESZ + 256, 1024);   // + scale / shift tail
    if (a.n_res && a.raw_stage < 18u * 10u * 32u * ESZ + 256u) a.raw_stage = (uint32_t)align_up_sz(18 * 10 * 32 * ESZ + 256, 1024);
    a.a_stage = (uint32_t)align_up_sz((size_t)(G::SLAB / 8) * G::PLANE * 16, 128);
    if (a.n_res && a.a_stage < 4u * 182u * 16u) a.a_stage = (uint32_t)align_up_sz(4 * 182 * 16, 128);
    a.w_stage = 32u * TPC * (uint32_t)p.Cout;
    a.x3 = X3 ? 1 : 0;
    a.a_half = a.a_stage;
    a.w_half = a.w_stage;
    if (a.x3) { a.a_stage *= 2; a.w_stage *= 2; }   // (hi, lo) image pairs
    const int nchunks = a.n_main_chunks + a.n_res_chunks;
    const int nblk = p.Cout / 32;
    // shared-memory plan: header | out staging (per epilogue group) | residual ring (per group) | raw ring | operand ring |
    // weights.  Ring depth is worth more than epilogue double buffering: try the staging variants and keep the deepest rings.
    const size_t cap = kSmemMax - 1024;   // slack for the 1024-byte alignment of the base
    auto fixed = [&]() { return (size_t)kHdrBytes + (size_t)kEpiGroups * (a.NOUT + a.NRES) * kOutTileBytes + (size_t)a.NR * a.raw_stage + (size_t)a.NA * a.a_stage; };
    auto total = [&]() { return fixed() + (size_t)a.NW * a.w_stage; };
    const int want_out = nblk > 1 ? 2 : 1, want_res = p.res_identity ? 2 : 0;
    const int try_out[3] = {want_out, 1, 1}, try_res[3] = {want_res, want_res, want_res ? 1 : 0};
    int best_score = -1;
    TcArgs best = a;
    for (int k = 0; k < 3; ++k) {
        a.NOUT = try_out[k]; a.NRES = try_res[k];
        a.NR = 2; a.NA = 2;
        a.resident = nchunks <= kMaxW && fixed() + (size_t)nchunks * a.w_stage <= cap;
        a.NW = a.resident ? nchunks : (TPC == 9 ? 2 : 4);
        if (!a.resident && total() > cap) a.NW = 2;   // last resort: a two-deep chunk ring
        if (total() > cap) continue;
        // Ring depths stay EVEN: the two transform groups take alternate slabs, so with an even depth a stage always belongs to the
        // same group.  With an odd depth the groups alternate on a stage, and a group that runs ahead can poll raw_full[s] for use
        // k + 1 while the load of use k (the other group's, TMA loads complete out of order) is still in flight: mbarrier waits only
        // see the phase parity, the wait passes on the completed use k - 1, the group transforms the wrong slab and releases the
        // stage early - silent corruption, then a second arrive.expect_tx on an unfinished phase (Warp Illegal Instruction).
        for (bool grew = true; grew;) {   // spend what is left on deeper rings
            grew = false;
            if (a.NR + 2 <= kMaxRing) { a.NR += 2; if (total() <= cap) grew = true; else a.NR -= 2; }
            if (a.NA + 2 <= kMaxRing && a.NA + 2 <= a.NR) { a.NA += 2; if (total() <= cap) grew = true; else a.NA -= 2; }
            if (!a.resident && a.NW < 12 && a.NW < nchunks) { ++a.NW; if (total() <= cap) grew = true; else --a.NW; }
        }
        const int score = (a.NR < 4 ? a.NR : 4) * 100 + (a.NA < 4 ? a.NA : 4) * 10 + (2 - k);
        if (score > best_score) { best_score = score; best = a; }
    }
    if (best_score < 0) return 1;
    a = best;
    if ((a.NR | a.NA) & 1) { set_error("conv tc: internal: odd ring depth %d / %d (two transform groups need even rings)", a.NR, a.NA); return SDDM_E_INVALID; }
    a.off_out = kHdrBytes;
    a.off_res = a.off_out + (uint32_t)(kEpiGroups * a.NOUT) * kOutTileBytes;
    a.off_raw = a.off_res + (uint32_t)(kEpiGroups * a.NRES) * kOutTileBytes;
    a.off_a = a.off_raw + (uint32_t)a.NR * a.raw_stage;
    a.off_w = a.off_a + (uint32_t)a.NA * a.a_stage;
    const size_t smem = total() + 1024;

    TcMaps maps;
    memset(&maps, 0, sizeof(maps));
    int rc;
    // raw slabs: fp32 pixels of 32 channels are 128 bytes -> 128B swizzle; bf16 pixels (64 / 32 bytes) are read chunk-contiguously
    const int swz_raw = (!A16 && G::SLAB == 32) ? 128 : 0, swz_raw_res = A16 ? 0 : 128, swz_tile = A16 ? 64 : 128;
    for (int i = 0; i < p.nsrc; ++i)
        if ((rc = encode_nhwc(&maps.src[i], p.src[i].x, p.B, p.Hin, p.Win, p.src[i].C, G::SLAB, G::RAW_W, G::RAW_H, swz_raw, A16))) return rc;
    if (a.n_res)
        for (int i = 0; i < p.res_nsrc; ++i)
            if ((rc = encode_nhwc(&maps.src[2 + i], p.res_src[i].x, p.B, p.Hin, p.Win, p.res_src[i].C, 32, 10, 18, swz_raw_res, A16))) return rc;
    if ((rc = encode_nhwc(&maps.out, p.out, p.B, p.Hout, p.Wout, p.Cout, 32, TW, TH, swz_tile, A16))) return rc;
    if (p.res_identity && (rc = encode_nhwc(&maps.res, p.res_src[0].x, p.B, p.Hout, p.Wout, p.Cout, 32, TW, TH, swz_tile, A16))) return rc;

    SDDM_SET_MAX_SMEM((conv3x3_tc_kernel<MODE, TPC, A16, X3>), kSmemMax);
    { static int skip = -1; if (skip < 0) { const char* e = getenv("SDDM_TC_SKIP"); skip = e ? atoi(e) : 0; } a.skip = skip; }
    a.trace = g_trace ? g_trace + (size_t)(g_trace_launch++ % 64) * 48 : nullptr;
    const int grid = a.ntiles < num_sms() ? a.ntiles : num_sms();
    SDDM_CUDA_TRY(launch_pdl(conv3x3_tc_kernel<MODE, TPC, A16, X3>, dim3(grid), dim3(kThreads), smem, st, a, maps));
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

// debug: timed-out waits of conv3x3_tc_kernel leave their notes in the same buffer as conv_row.cu's (no-op unless built with SDDM_TC_HANG_NOTES)
cudaError_t conv_tc_set_hang_buffer(unsigned* dev_ptr) {
#ifdef SDDM_TC_HANG_NOTES
    return set_hang_buffer(dev_ptr);
#else
    (void)dev_ptr;
    return cudaSuccess;
#endif
}

int conv_tc_nparts(int Hout, int Wout) { return ((Hout + TH - 1) / TH) * ((Wout + TW - 1) / TW); }
int conv_tc_tiles(int Hout, int Wout) { return ((Hout + TH - 1) / TH) * ((Wout + TW - 1) / TW); }

int launch_conv_tc(const ConvP& p, cudaStream_t st) {
    if (!conv_tc_supported(p) || !p.w_tc) { set_error("conv tc: unsupported shape Cin=%d Cout=%d mode=%d", p.Cin, p.Cout, p.mode); return SDDM_E_INVALID; }
    if (p.x3 && (p.act16 || !p.w_tc_lo)) { set_error("conv tc: the split-bf16 mode needs fp32 activations and the lo weight image"); return SDDM_E_INVALID; }
    const int ein_h = p.mode == CONV_S2 ? p.Hout * 2 : (p.mode == CONV_UP ? p.Hout / 2 : p.Hout);
    const int ein_w = p.mode == CONV_S2 ? p.Wout * 2 : (p.mode == CONV_UP ? p.Wout / 2 : p.Wout);
    if (ein_h != p.Hin || ein_w != p.Win) { set_error("conv tc: inconsistent spatial sizes"); return SDDM_E_INVALID; }
    const bool has_res_conv = p.res_Cin && !p.res_identity;
    if (has_res_conv && p.x3 && !p.res_w_tc_lo) { set_error("conv tc: res_conv lo weight image missing"); return SDDM_E_INVALID; }
    if (has_res_conv && (!p.res_w_tc || p.mode != CONV_S1)) { set_error("conv tc: res_conv needs packed weights and stride 1"); return SDDM_E_INVALID; }
    if (p.mode == CONV_UP && ((p.Hout % TH) || (p.Wout % TW))) { set_error("conv tc: upsampled output must tile by 16x8"); return SDDM_E_INVALID; }
    TcArgs a{};
    a.p = p;
    a.tiles_x = (p.Wout + TW - 1) / TW;
    a.tiles_y = (p.Hout + TH - 1) / TH;
    a.ntiles = p.B * a.tiles_x * a.tiles_y;
    a.magic_per = (0x100000000ull + (unsigned long long)(a.tiles_x * a.tiles_y) - 1) / (unsigned long long)(a.tiles_x * a.tiles_y);
    a.magic_tx = (0x100000000ull + (unsigned long long)a.tiles_x - 1) / (unsigned long long)a.tiles_x;
    if (p.parts && p.nparts != a.tiles_x * a.tiles_y) { set_error("conv tc: nparts mismatch"); return SDDM_E_INVALID; }
    a.n_res = has_res_conv ? p.res_Cin / 32 : 0;
    a.n_res_chunks = has_res_conv ? p.res_Cin / 32 : 0;
    a.acc_stride = p.Cout <= 32 ? 32 : (p.Cout <= 64 ? 64 : (p.Cout <= 128 ? 128 : 256));
    a.tmem_cols = 2 * a.acc_stride;
    a.temb_per_row = (p.temb && p.temb_stride != 0) ? 1 : 0;
    int rc;
#define SDDM_TC_DISPATCH(M)                                                                                   \
    do {                                                                                                     \
        if (p.act16) { rc = launch_mode<M, 9, true, false>(a, st); if (rc == 1) rc = launch_mode<M, 3, true, false>(a, st); } \
        else if (!p.x3) { rc = launch_mode<M, 9, false, false>(a, st); if (rc == 1) rc = launch_mode<M, 3, false, false>(a, st); } \
        else {                                                                                               \
            rc = launch_mode<M, 9, false, true>(a, st);                                                      \
            if (rc == 1) rc = launch_mode<M, 3, false, true>(a, st);                                         \
            if (rc == 1) rc = launch_mode<M, 1, false, true>(a, st);   /* split-bf16 stride-2 layers: one tap per weight chunk */ \
        }                                                                                                    \
    } while (0)
    switch (p.mode) {
        case CONV_S1: SDDM_TC_DISPATCH(CONV_S1); break;
        case CONV_S2: SDDM_TC_DISPATCH(CONV_S2); break;
        default: SDDM_TC_DISPATCH(CONV_UP); break;
    }
#undef SDDM_TC_DISPATCH
    if (rc == 1) { set_error("conv tc: Cout=%d does not fit the shared-memory plan", p.Cout); return SDDM_E_INVALID; }
    return rc;
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
    if (geo == 1) { a_off = 11 * 16; a_lbo = 182 * 16; a_sbo = 160; }
    if (geo == 2) { a_off = (20 + 12) * 16; a_lbo = 662 * 16; a_sbo = 640; }
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

// debug: enable != 0 -> (re)start tracing: the next 64 conv_tc launches record per-role wait cycles of CTA 0;
// enable == 0 -> copy the [64][48] counters to host_out (may be null) and stop tracing.
extern "C" SDDM_API int sddm_debug_tc_trace(int enable, long long* host_out) {
    using namespace sddm;
    if (enable) {
        if (!g_trace) SDDM_CUDA_TRY(cudaMalloc(&g_trace, 64 * 48 * sizeof(long long)));
        SDDM_CUDA_TRY(cudaMemset(g_trace, 0, 64 * 48 * sizeof(long long)));
        g_trace_launch = 0;
        return SDDM_OK;
    }
    if (!g_trace) { set_error("tracing was not enabled"); return SDDM_E_STATE; }
    SDDM_CUDA_TRY(cudaDeviceSynchronize());
    if (host_out) SDDM_CUDA_TRY(cudaMemcpy(host_out, g_trace, 64 * 48 * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(g_trace);
    g_trace = nullptr;
    return SDDM_OK;
}

// debug: average cycles per back-to-back tcgen05.mma (M 128, K 16, bf16) for a given N and A-operand geometry (see umma_rate_kernel)
extern "C" SDDM_API int sddm_debug_umma_rate(int N, int reps, int nA, int geo, float* cycles_per_mma) {
    using namespace sddm;
    if (!cycles_per_mma || N % 16 || N < 16 || N > 256 || reps < 1 || nA < 1 || nA > 16 || geo < 0 || geo > 2) { set_error("umma_rate: bad arguments"); return SDDM_E_INVALID; }
    const int nblocks = nA > 9 ? 148 : 1;   // nA > 9: one CTA per SM, report the slowest
    long long* d = nullptr;
    SDDM_CUDA_TRY(cudaMalloc(&d, 148 * sizeof(long long)));
    SDDM_CUDA_TRY(cudaMemset(d, 0, 148 * sizeof(long long)));
    const size_t smem = 64 * 1024 + (size_t)N * 32 + 2048;
    SDDM_CUDA_TRY(cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_rate_kernel<<<nblocks, 128, smem>>>(N, reps, nA, geo, d);
    count_launch();
    long long hv[148], h = 0;
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(hv, d, sizeof(hv), cudaMemcpyDeviceToHost);
    for (int i = 0; i < nblocks; ++i) h = hv[i] > h ? hv[i] : h;
    cudaFree(d);
    if (e != cudaSuccess) { set_error("umma_rate: CUDA error %s", cudaGetErrorName(e)); return SDDM_E_CUDA; }
    *cycles_per_mma = (float)h / (float)((reps / 9) * 9);
    return SDDM_OK;
}

