// tcgen05 / TMEM / TMA kernels of the DiffWave denoiser (cfg 5; reference model/diffwave.py:64-108), sm_100a.
//
// dw_layer_tc_kernel — one ResidualBlock over all time tiles (persistent CTAs, one per SM, tile = 128 samples of one utterance):
//   operands   the residual stream is time-major bf16 [B][T][64]: a row is exactly one 128-byte swizzle atom, so ONE TMA box
//              [128 t][64 c] (SWIZZLE_128B) IS a K-major UMMA operand.  The three taps of the dilated conv are three boxes at
//              t0 - d, t0, t0 + d; rows outside [0, T) are zero filled by TMA (the conv's zero padding).
//   MMA 1      acc1[128 x 128] = sum_tap X_tap[128 x 64] . Wd_tap^T          (12 tcgen05.mma, fp32 in TMEM)
//   epilogue 1 + cached conditioner row (bf16, read straight from HBM) + diffusion-step bias -> sigmoid(gate) * tanh(filter)
//              -> bf16 z tile written into shared memory in the same swizzled K-major layout = operand of MMA 2
//   MMA 2      acc2[128 x 128] = z[128 x 64] . [W_res | W_skip]^T            (4 tcgen05.mma)
//   epilogue 2 x_out = (x + acc2[:, :64] + b_res) / sqrt(2)  (x re-read from the centre box in shared memory, bf16 store);
//              skip += acc2[:, 64:] + b_skip  (fp32 read-modify-write by the owning thread; the loads are issued while MMA 2 runs)
//   pipeline   warps 0-3 / 4-7: epilogue groups 0 / 1 (alternate tiles; group g owns operand stage g and TMEM columns
//              [256 g, 256 g + 256)), warp 8: TMA producer, warp 9: MMA issuer.  MMA 2 of tile i-1 is issued after MMA 1 of
//              tile i, so the tensor pipe works on one group's tile while the other group runs its epilogue.
//   weights    Wd (48 KB) and [W_res | W_skip] (16 KB) stay resident in shared memory for the CTA's lifetime.
//
// dw_cond_tc_kernel — conditioner_projection of one layer for one utterance: [128 t x KP] . [128 n x KP]^T, 64-wide K chunks
//   through a 4-stage TMA ring, + bias, bf16 output row [128] per time step (the cache the layer kernel reads).
#include <cstring>
#include <map>
#include <tuple>

#include "diffwave.cuh"
#include "tc_ptx.cuh"

namespace sddm {
namespace {

// ---- layer kernel ------------------------------------------------------------------------------------
constexpr int kTile = 128;
constexpr uint32_t kBox = kTile * DW_C * 2;             // 16 KB: one [128][64] bf16 box (operand / z / cond half / x_out tile)
constexpr uint32_t kOffW1 = 0;                          // 3 boxes
constexpr uint32_t kOffW2 = 3 * kBox;                   // [64 n][64 k]: half a box
constexpr uint32_t kOffA = kOffW2 + kBox / 2;           // 2 stages x 3 boxes  (centre, left, right)
constexpr uint32_t kOffZ = kOffA + 6 * kBox;            // 2 boxes (one per epilogue group): z, then reused for x_out
constexpr uint32_t kOffCond = kOffZ + 2 * kBox;         // 2 boxes: conditioner columns 0-63 (gate) | 64-127 (filter), one tile
constexpr uint32_t kOffB2 = kOffCond + 2 * kBox;        // 64 floats
constexpr uint32_t kOffBar = kOffB2 + 256;
constexpr uint32_t kLayerSmem = kOffBar + 256 + 1024;   // + alignment slack
constexpr int kLayerThreads = 320;
enum { B_AFULL = 0, B_AEMPTY = 2, B_ACC1F = 4, B_ACC1E = 6, B_ZFULL = 8, B_ACC2F = 10, B_CFULL = 12, B_WFULL = 14, B_CEMPTY = 15, B_COUNT = 16 };

struct alignas(64) LayerMaps {
    CUtensorMap x, xo, cond, zc, w1, w2;
};

__global__ void __launch_bounds__(kLayerThreads, 1) dw_layer_tc_kernel(const __grid_constant__ LayerMaps maps, DwLayerTc p, int ntiles,
                                                                       int tiles_per_row) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* gbase = smem_raw + (base - smem_u32(smem_raw));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bars = base + kOffBar;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + kOffBar + 192);
    float* b2s = reinterpret_cast<float*>(gbase + kOffB2);
    auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar(B_AFULL + s), 1);
            mbar_init(bar(B_AEMPTY + s), 128);
            mbar_init(bar(B_ACC1F + s), 1);
            mbar_init(bar(B_ACC1E + s), 128);
            mbar_init(bar(B_ZFULL + s), 1);
            mbar_init(bar(B_ACC2F + s), 1);
            mbar_init(bar(B_CFULL + s), 1);   // one per group although the buffer is shared: a waiter is never two phases behind
        }
        mbar_init(bar(B_WFULL), 1);
        mbar_init(bar(B_CEMPTY), 128);
        fence_barrier_init();
    }
    if (tid < DW_C) b2s[tid] = p.b2[tid];
    if (warp == 9) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int my_tiles = ntiles > (int)blockIdx.x ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (warp == 8) {
        // ================= TMA producer: operand boxes + conditioner tile =================
        if (lane == 0) {
            mbar_expect_tx(bar(B_WFULL), 3 * kBox + kBox / 2);   // weights are not produced by the previous kernel: load before the PDL wait
            for (int tap = 0; tap < 3; ++tap) tma_load_2d(base + kOffW1 + tap * kBox, &maps.w1, 0, tap * DW_N, bar(B_WFULL));
            tma_load_2d(base + kOffW2, &maps.w2, 0, 0, bar(B_WFULL));
        }
        pdl_wait();
        if (lane == 0) {
            for (int i = 0; i < my_tiles; ++i) {
                const int tile = blockIdx.x + i * gridDim.x;
                const int b = tile / tiles_per_row, t0 = (tile - b * tiles_per_row) * kTile;
                const int s = i & 1, n = i >> 1;
                mbar_wait(bar(B_AEMPTY + s), (n & 1) ^ 1);
                mbar_expect_tx(bar(B_AFULL + s), 3 * kBox);
                const uint32_t dst = base + kOffA + s * 3 * kBox;
                tma_load_3d(dst, &maps.x, 0, t0, b, bar(B_AFULL + s));
                tma_load_3d(dst + kBox, &maps.x, 0, t0 - p.dil, b, bar(B_AFULL + s));
                tma_load_3d(dst + 2 * kBox, &maps.x, 0, t0 + p.dil, b, bar(B_AFULL + s));
                mbar_wait(bar(B_CEMPTY), (i & 1) ^ 1);
                mbar_expect_tx(bar(B_CFULL + s), 2 * kBox);
                tma_load_3d(base + kOffCond, &maps.cond, 0, t0, b, bar(B_CFULL + s));
                tma_load_3d(base + kOffCond + kBox, &maps.cond, 64, t0, b, bar(B_CFULL + s));
            }
        }
    } else if (warp == 9) {
        // ================= MMA issuer =================
        const uint32_t idesc1 = make_idesc(DW_N), idesc2 = make_idesc(DW_C);
        mbar_wait(bar(B_WFULL), 0);
        for (int i = 0; i <= my_tiles; ++i) {
            if (i < my_tiles) {
                const int s = i & 1, n = i >> 1;
                mbar_wait(bar(B_AFULL + s), n & 1);
                mbar_wait(bar(B_ACC1E + s), (n & 1) ^ 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a0 = base + kOffA + s * 3 * kBox, acc1 = tmem + 256u * s;
#pragma unroll
                    for (int bx = 0; bx < 3; ++bx) {   // box 0 / 1 / 2 = centre / left / right = filter tap 1 / 0 / 2
                        const int tap = bx == 0 ? 1 : (bx == 1 ? 0 : 2);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma(acc1, make_desc_sw128(a0 + bx * kBox + k * 32), make_desc_sw128(base + kOffW1 + tap * kBox + k * 32), idesc1,
                                 (bx | k) != 0);
                    }
                    umma_commit(bar(B_ACC1F + s));
                }
                __syncwarp();
            }
            if (i >= 1) {   // second GEMM of the previous tile (the other epilogue group's)
                const int j = i - 1, s = j & 1, n = j >> 1;
                mbar_wait(bar(B_ZFULL + s), n & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t z0 = base + kOffZ + s * kBox, acc2 = tmem + 256u * s + 128u;
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma(acc2, make_desc_sw128(z0 + k * 32), make_desc_sw128(base + kOffW2 + k * 32), idesc2, k != 0);
                    umma_commit(bar(B_ACC2F + s));
                }
                __syncwarp();
            }
        }
    } else {
        // ================= epilogue groups =================
        pdl_wait();
        const int g = warp >> 2, row = (warp & 3) * 32 + lane;
        const bool leader = (tid & 127) == 0;
        const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t acc1 = tmem + 256u * g + lane_off, acc2 = acc1 + 128u;
        const uint32_t zbox = base + kOffZ + g * kBox, zrow = zbox + row * 128, xrow = base + kOffA + g * 3 * kBox + row * 128;
        const uint32_t crow = base + kOffCond + row * 128;
        const uint32_t sw = (uint32_t)(row & 7);
        const uint32_t zmask = (uint32_t)p.opaque_zero;   // always 0, but only the host knows
        for (int i = g, n = 0; i < my_tiles; i += 2, ++n) {
            const int tile = blockIdx.x + i * gridDim.x;
            const int b = tile / tiles_per_row, t0 = (tile - b * tiles_per_row) * kTile, t = t0 + row;
            // ---- epilogue 1: gate
            uint4 cv[16];                          // this row of the cached conditioner: gate chunks 0-7, filter chunks 8-15
            mbar_wait(bar(B_CFULL + g), n & 1);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                cv[q] = lds128(crow + ((((uint32_t)q) ^ sw) << 4));
                cv[8 + q] = lds128(crow + kBox + ((((uint32_t)q) ^ sw) << 4));
            }
            // Release only once the loads have RETURNED: ld.shared is asynchronous and mbarrier.arrive does not wait for it, so an
            // arrive issued right behind the loads lets the producer's TMA refill overtake loads still queued in the shared-memory
            // pipe (seen as rare corrupted tiles at > 8 tiles per CTA).  Folding one word of every load into the barrier address
            // (masked by a zero the compiler cannot see) makes the arrive data-dependent on all of them.
            uint32_t dep = 0;
#pragma unroll
            for (int q = 0; q < 16; ++q) dep ^= cv[q].x;
            mbar_arrive(bar(B_CEMPTY) + (dep & zmask));   // the shared conditioner buffer may be refilled for the next tile
            const int v = (t < p.dil ? 1 : 0) | (t + p.dil >= p.T ? 2 : 0);
            const float4* b1 = reinterpret_cast<const float4*>(p.bias1 + (size_t)b * p.bias1_row_stride + v * DW_N);
            mbar_wait(bar(B_AFULL + g), n & 1);    // acquires the TMA-written centre box for this thread's reads
            uint4 xv[8];                           // this row of the layer input, for the residual connection of epilogue 2
#pragma unroll
            for (int q = 0; q < 8; ++q) xv[q] = lds128(xrow + ((((uint32_t)q) ^ sw) << 4));
            dep = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) dep ^= xv[q].x;
            mbar_wait(bar(B_ACC1F + g), n & 1);
            mbar_arrive(bar(B_AEMPTY + g) + (dep & zmask));   // MMA 1 has consumed the three boxes and x IS in registers: refill the stage
            tc_fence_after();
            if (leader) bulk_wait_read_0();        // the x_out store of this group's previous tile has finished reading the z box
            group_bar(1 + g);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float ga[32], fa[32];
                tmem_ld32(acc1 + 32 * h, ga);
                tmem_ld32(acc1 + 64 + 32 * h, fa);
                if (h == 1) {   // every TMEM read of acc1 is done: MMA 1 of this group's next tile may overwrite it
                    tc_fence_before();
                    mbar_arrive(bar(B_ACC1E + g));
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {   // 8 channels per 16-byte chunk
                    const uint4 cg = cv[4 * h + q], cf = cv[8 + 4 * h + q];
                    const float4 bg0 = __ldg(b1 + 8 * h + 2 * q), bg1 = __ldg(b1 + 8 * h + 2 * q + 1);
                    const float4 bf0 = __ldg(b1 + 16 + 8 * h + 2 * q), bf1 = __ldg(b1 + 16 + 8 * h + 2 * q + 1);
                    const float cgs[8] = {bf16_lo(cg.x), bf16_hi(cg.x), bf16_lo(cg.y), bf16_hi(cg.y), bf16_lo(cg.z), bf16_hi(cg.z), bf16_lo(cg.w), bf16_hi(cg.w)};
                    const float cfs[8] = {bf16_lo(cf.x), bf16_hi(cf.x), bf16_lo(cf.y), bf16_hi(cf.y), bf16_lo(cf.z), bf16_hi(cf.z), bf16_lo(cf.w), bf16_hi(cf.w)};
                    const float bgs[8] = {bg0.x, bg0.y, bg0.z, bg0.w, bg1.x, bg1.y, bg1.z, bg1.w};
                    const float bfs[8] = {bf0.x, bf0.y, bf0.z, bf0.w, bf1.x, bf1.y, bf1.z, bf1.w};
                    float z[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float gt = ga[8 * q + e] + cgs[e] + bgs[e], ft = fa[8 * q + e] + cfs[e] + bfs[e];
                        z[e] = fmaf(0.5f, tanh_approx(0.5f * gt), 0.5f) * tanh_approx(ft);     // sigmoid(g) * tanh(f)
                    }
                    uint4 o;
                    o.x = pack_bf16(z[0], z[1]); o.y = pack_bf16(z[2], z[3]); o.z = pack_bf16(z[4], z[5]); o.w = pack_bf16(z[6], z[7]);
                    sts128(zrow + ((((uint32_t)(4 * h + q)) ^ sw) << 4), o);
                }
            }
            fence_async_smem();
            group_bar(1 + g);
            if (leader) {
                mbar_arrive(bar(B_ZFULL + g));                 // MMA 2 may read z
                tma_store_3d(&maps.zc, zbox, 0, t0, b);        // ... and so does the store into the z cache
                bulk_commit();
            }
            // ---- epilogue 2: residual
            mbar_wait(bar(B_ACC2F + g), n & 1);
            tc_fence_after();
            uint4 xo[8];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float r[32];
                tmem_ld32(acc2 + 32 * h, r);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint4 xq = xv[4 * h + q];
                    const float xs[8] = {bf16_lo(xq.x), bf16_hi(xq.x), bf16_lo(xq.y), bf16_hi(xq.y), bf16_lo(xq.z), bf16_hi(xq.z), bf16_lo(xq.w), bf16_hi(xq.w)};
                    float o[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] = (xs[e] + r[8 * q + e] + b2s[32 * h + 8 * q + e]) * 0.70710678118654752f;
                    xo[4 * h + q].x = pack_bf16(o[0], o[1]); xo[4 * h + q].y = pack_bf16(o[2], o[3]);
                    xo[4 * h + q].z = pack_bf16(o[4], o[5]); xo[4 * h + q].w = pack_bf16(o[6], o[7]);
                }
            }
            tc_fence_before();
            if (leader) bulk_wait_read_0();        // the z-cache store has finished reading the z box (MMA 2 has: acc2 is complete)
            group_bar(1 + g);
#pragma unroll
            for (int q = 0; q < 8; ++q) sts128(zrow + ((((uint32_t)q) ^ sw) << 4), xo[q]);
            fence_async_smem();
            group_bar(1 + g);
            if (leader) {
                tma_store_3d(&maps.xo, zbox, 0, t0, b);
                bulk_commit();
            }
        }
        if (leader) bulk_wait_all();               // global writes of this CTA are complete before the kernel's end is signalled
    }
    __syncthreads();
    pdl_launch_dependents();
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem, 512);
}

// ---- skip / output head --------------------------------------------------------------------------------
constexpr int kFinStages = 6;
constexpr uint32_t kFinStage = kBox + kBox / 2;                  // z box + [64][64] weight box
constexpr uint32_t kFinOffS = kFinStages * kFinStage;            // s tile [128][64] bf16
constexpr uint32_t kFinOffWsp = kFinOffS + kBox;                 // [64][64]
constexpr uint32_t kFinOffBar = kFinOffWsp + kBox / 2;
constexpr uint32_t kFinSmem = kFinOffBar + 256 + 1024;

struct alignas(64) FinalMaps {
    CUtensorMap zc, ws, wsp;
};

__global__ void __launch_bounds__(192, 1) dw_final_tc_kernel(const __grid_constant__ FinalMaps maps, DwFinalTc p, int tiles_per_row) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* gbase = smem_raw + (base - smem_u32(smem_raw));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bars = base + kFinOffBar;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + kFinOffBar + 192);
    auto full = [&](int s) { return bars + 8u * (uint32_t)s; };
    auto empty = [&](int s) { return bars + 8u * (uint32_t)(kFinStages + s); };
    const uint32_t accf = bars + 8u * (2 * kFinStages), sfull = accf + 8, acc2f = accf + 16, wfull = accf + 24;
    if (tid == 0) {
        for (int s = 0; s < kFinStages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        mbar_init(accf, 1);
        mbar_init(sfull, 128);
        mbar_init(acc2f, 1);
        mbar_init(wfull, 1);
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(smem_u32(tmem_slot), 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int b = blockIdx.x / tiles_per_row, t0 = (blockIdx.x - b * tiles_per_row) * kTile;
    if (warp == 4) {
        if (lane == 0) {
            mbar_expect_tx(wfull, kBox / 2);
            tma_load_2d(base + kFinOffWsp, &maps.wsp, 0, 0, wfull);
        }
        pdl_wait();
        if (lane == 0) {
            for (int l = 0; l < p.L; ++l) {
                const int s = l % kFinStages, n = l / kFinStages;
                mbar_wait(empty(s), (n & 1) ^ 1);
                mbar_expect_tx(full(s), kFinStage);
                tma_load_3d(base + s * kFinStage, &maps.zc, 0, b * p.T + t0, l, full(s));
                tma_load_2d(base + s * kFinStage + kBox, &maps.ws, 0, l * DW_C, full(s));
            }
        }
    } else if (warp == 5) {
        const uint32_t idesc = make_idesc(DW_C);
        for (int l = 0; l < p.L; ++l) {
            const int s = l % kFinStages, n = l / kFinStages;
            mbar_wait(full(s), n & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a0 = base + s * kFinStage;
#pragma unroll
                for (int k = 0; k < 4; ++k) umma(tmem, make_desc_sw128(a0 + k * 32), make_desc_sw128(a0 + kBox + k * 32), idesc, (l | k) != 0);
                umma_commit(empty(s));
                if (l == p.L - 1) umma_commit(accf);
            }
            __syncwarp();
        }
        mbar_wait(wfull, 0);
        mbar_wait(sfull, 0);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma(tmem + 64, make_desc_sw128(base + kFinOffS + k * 32), make_desc_sw128(base + kFinOffWsp + k * 32), idesc, k != 0);
            umma_commit(acc2f);
        }
        __syncwarp();
    } else {
        const int row = warp * 32 + lane;
        const uint32_t acc = tmem + ((uint32_t)(warp * 32) << 16);
        const uint32_t srow = base + kFinOffS + row * 128, sw = (uint32_t)(row & 7);
        mbar_wait(accf, 0);
        tc_fence_after();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float r[32];
            tmem_ld32(acc + 32 * h, r);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = (r[8 * q + e] + __ldg(p.bsum + 32 * h + 8 * q + e)) * p.inv_sqrt_layers;
                uint4 o;
                o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
                sts128(srow + ((((uint32_t)(4 * h + q)) ^ sw) << 4), o);
            }
        }
        tc_fence_before();
        fence_async_smem();
        mbar_arrive(sfull);
        mbar_wait(acc2f, 0);
        tc_fence_after();
        float e = p.bo;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float r[32];
            tmem_ld32(acc + 64 + 32 * h, r);
#pragma unroll
            for (int c = 0; c < 32; ++c) e = fmaf(fmaxf(r[c] + __ldg(p.bsp + 32 * h + c), 0.f), __ldg(p.wo + 32 * h + c), e);
        }
        p.eps[(size_t)b * p.T + t0 + row] = e;
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem, 128);
}

// ---- conditioner GEMM --------------------------------------------------------------------------------
constexpr int kCondStages = 4;
constexpr uint32_t kCondStage = 2 * kBox;                               // A box + B box
constexpr uint32_t kCondOffBar = kCondStages * kCondStage;
constexpr uint32_t kCondSmem = kCondOffBar + 256 + 1024;

struct alignas(64) CondMaps {
    CUtensorMap up, w;
};

__global__ void __launch_bounds__(192, 1) dw_cond_tc_kernel(const __grid_constant__ CondMaps maps, DwCondTc p, int nchunks) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* gbase = smem_raw + (base - smem_u32(smem_raw));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bars = base + kCondOffBar;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + kCondOffBar + 128);
    auto full = [&](int s) { return bars + 8u * (uint32_t)s; };
    auto empty = [&](int s) { return bars + 8u * (uint32_t)(kCondStages + s); };
    const uint32_t accf = bars + 8u * 2 * kCondStages;
    if (tid == 0) {
        for (int s = 0; s < kCondStages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        mbar_init(accf, 1);
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(smem_u32(tmem_slot), 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int t0 = blockIdx.x * kTile;
    if (warp == 4) {
        if (lane == 0) {
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % kCondStages, n = c / kCondStages;
                mbar_wait(empty(s), (n & 1) ^ 1);
                mbar_expect_tx(full(s), kCondStage);
                tma_load_2d(base + s * kCondStage, &maps.up, c * 64, t0, full(s));
                tma_load_2d(base + s * kCondStage + kBox, &maps.w, c * 64, 0, full(s));
            }
        }
    } else if (warp == 5) {
        const uint32_t idesc = make_idesc(DW_N);
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % kCondStages, n = c / kCondStages;
            mbar_wait(full(s), n & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a0 = base + s * kCondStage;
#pragma unroll
                for (int k = 0; k < 4; ++k) umma(tmem, make_desc_sw128(a0 + k * 32), make_desc_sw128(a0 + kBox + k * 32), idesc, (c | k) != 0);
                umma_commit(empty(s));
                if (c == nchunks - 1) umma_commit(accf);
            }
            __syncwarp();
        }
    } else {
        const int row = warp * 32 + lane;
        mbar_wait(accf, 0);
        tc_fence_after();
        __nv_bfloat16* o = p.out + (size_t)(t0 + row) * DW_N;
        const uint32_t acc = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            float r[32];
            tmem_ld32(acc + 32 * h, r);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = r[8 * q + e] + __ldg(p.bias + 32 * h + 8 * q + e);
                uint4 pk;
                pk.x = pack_bf16(v[0], v[1]); pk.y = pack_bf16(v[2], v[3]); pk.z = pack_bf16(v[4], v[5]); pk.w = pack_bf16(v[6], v[7]);
                *reinterpret_cast<uint4*>(o + 32 * h + 8 * q) = pk;
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem, 128);
}

// ---- host ----------------------------------------------------------------------------------------------
}  // namespace

int launch_dw_layer_tc(const DwLayerTc& p, cudaStream_t st) {
    if (p.T % kTile) { set_error("diffwave tc: T must be a multiple of %d", kTile); return SDDM_E_INVALID; }
    SDDM_SET_MAX_SMEM(dw_layer_tc_kernel, kLayerSmem);
    // tensor maps depend only on (buffers, shape): cache them (one sampling run re-launches the same 30 layers T_steps times)
    static std::map<std::tuple<const void*, const void*, const void*, const void*, const void*, int, int>, LayerMaps> cache;
    const auto key = std::make_tuple((const void*)p.x_in, (const void*)p.x_out, (const void*)p.cond, (const void*)p.zc, (const void*)p.w1, p.B, p.T);
    auto it = cache.find(key);
    if (it == cache.end()) {
        if (cache.size() > 4096) cache.clear();
        LayerMaps m;
        int rc = encode_bf16(&m.x, p.x_in, DW_C, DW_C, p.T, p.B, 128);
        if (rc) return rc;
        if ((rc = encode_bf16(&m.xo, p.x_out, DW_C, DW_C, p.T, p.B, 128))) return rc;
        if ((rc = encode_bf16(&m.cond, p.cond, DW_N, DW_N, p.T, p.B, 128))) return rc;
        if ((rc = encode_bf16(&m.zc, p.zc, DW_C, DW_C, p.T, p.B, 128))) return rc;
        if ((rc = encode_bf16(&m.w1, p.w1, DW_C, DW_C, 3 * DW_N, 0, 128))) return rc;
        if ((rc = encode_bf16(&m.w2, p.w2, DW_C, DW_C, DW_C, 0, 64))) return rc;
        it = cache.emplace(key, m).first;
    }
    const int tiles_per_row = p.T / kTile, ntiles = p.B * tiles_per_row;
    const int grid = ntiles < num_sms() ? ntiles : num_sms();
    SDDM_CUDA_TRY(launch_pdl(dw_layer_tc_kernel, dim3(grid), dim3(kLayerThreads), kLayerSmem, st, it->second, p, ntiles, tiles_per_row));
    count_launch();
    return SDDM_OK;
}

int launch_dw_final_tc(const DwFinalTc& p, cudaStream_t st) {
    if (p.T % kTile) { set_error("diffwave tc: T must be a multiple of %d", kTile); return SDDM_E_INVALID; }
    SDDM_SET_MAX_SMEM(dw_final_tc_kernel, kFinSmem);
    static std::map<std::tuple<const void*, const void*, int, int, int>, FinalMaps> cache;
    const auto key = std::make_tuple((const void*)p.zc, (const void*)p.ws, p.L, p.B, p.T);
    auto it = cache.find(key);
    if (it == cache.end()) {
        if (cache.size() > 256) cache.clear();
        FinalMaps m;
        int rc = encode_bf16(&m.zc, p.zc, DW_C, DW_C, (long long)p.B * p.T, p.L, 128);
        if (rc) return rc;
        if ((rc = encode_bf16(&m.ws, p.ws, DW_C, DW_C, (long long)p.L * DW_C, 0, 64))) return rc;
        if ((rc = encode_bf16(&m.wsp, p.wsp, DW_C, DW_C, DW_C, 0, 64))) return rc;
        it = cache.emplace(key, m).first;
    }
    const int tiles_per_row = p.T / kTile;
    SDDM_CUDA_TRY(launch_pdl(dw_final_tc_kernel, dim3(p.B * tiles_per_row), dim3(192), kFinSmem, st, it->second, p, tiles_per_row));
    count_launch();
    return SDDM_OK;
}

int launch_dw_cond_tc(const DwCondTc& p, cudaStream_t st) {
    if (p.T % kTile || p.KP % 64) { set_error("diffwave tc: T %% 128 and KP %% 64 must be 0"); return SDDM_E_INVALID; }
    SDDM_SET_MAX_SMEM(dw_cond_tc_kernel, kCondSmem);
    CondMaps m;
    int rc = encode_bf16(&m.up, p.up, p.KP, p.KP, p.T, 0, 128);
    if (rc) return rc;
    if ((rc = encode_bf16(&m.w, p.w, p.KP, p.KP, DW_N, 0, 128))) return rc;
    dw_cond_tc_kernel<<<p.T / kTile, 192, kCondSmem, st>>>(m, p, p.KP / 64);
    SDDM_LAUNCH_CHECK();
    return SDDM_OK;
}

}  // namespace sddm
