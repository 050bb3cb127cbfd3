// tcgen05 / TMEM / TMA convolution of the WaveGrad denoiser (cfg 4; reference model/wavegrad.py:52-137), sm_100a.
//
// Every Conv1d is a GEMM tile [128 time rows] x [128 output columns] accumulated in TMEM over (tap, 64-channel chunk) steps.
// Both operands arrive by TMA in the 128-byte-swizzled K-major layout (5-stage ring, 32 KB per stage): the activation box is
// a window of a time-major bf16 tensor shifted by the tap offset (zero filled outside = the padding), the weight box a slice
// of the pre-packed [N][K] matrix.  Everything the reference applies between two convs (leaky_relu, the FiLM affine,
// nearest-neighbour resampling) is moved to the producer's epilogue or into the view / the polyphase weight packing, so
// the main loop is pure TMA -> tcgen05.mma.  Persistent CTAs, two TMEM accumulator stages, two epilogue groups.
#include <cstring>
#include <map>
#include <tuple>

#include "tc_ptx.cuh"
#include "wavegrad.cuh"

namespace sddm {
namespace {

constexpr int kTileM = 128, kTileN = 128;
constexpr uint32_t kBoxBytes = 128 * 64 * 2;          // 16 KB
constexpr int kStages = 5;
constexpr uint32_t kStageBytes = 2 * kBoxBytes;
constexpr int kGroups = 3;                                     // epilogue groups = TMEM accumulator stages
constexpr uint32_t kOffTile = kStages * kStageBytes;          // one transposition tile per epilogue warp, 4608 B each
constexpr uint32_t kOffBars = kOffTile + 4 * kGroups * 4608;
constexpr uint32_t kSmem = kOffBars + 256 + 1024;
constexpr int kThreads = (4 * kGroups + 2) * 32;
constexpr int kProdWarp = 4 * kGroups, kMmaWarp = 4 * kGroups + 1;

struct alignas(64) WgMaps {
    CUtensorMap a, w;
};

__device__ __forceinline__ float lrelu02(float v) { return v > 0.f ? v : 0.2f * v; }

// Persistent CTA (one per SM): warps 4g .. 4g+3 = epilogue group g of kGroups (tiles round robin, TMEM columns [128 g, 128 g + 128)),
// then one TMA producer warp and one MMA issuer warp.  The operand ring runs ahead across tile boundaries, so the loads and MMAs of the
// next tiles overlap the (global-memory-latency-bound) epilogues of the previous two.
__global__ void __launch_bounds__(kThreads, 1) wg_conv_tc_kernel(const __grid_constant__ WgMaps maps, WgTcConv p, int tiles_n, int tiles_m, int ntiles) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* gbase = smem_raw + (base - smem_u32(smem_raw));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bars = base + kOffBars;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + kOffBars + 192);
    auto full = [&](int s) { return bars + 8u * (uint32_t)s; };
    auto empty = [&](int s) { return bars + 8u * (uint32_t)(kStages + s); };
    auto accf = [&](int g) { return bars + 8u * (uint32_t)(2 * kStages + g); };
    auto acce = [&](int g) { return bars + 8u * (uint32_t)(2 * kStages + kGroups + g); };
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        for (int g = 0; g < kGroups; ++g) { mbar_init(accf(g), 1); mbar_init(acce(g), 128); }
        fence_barrier_init();
    }
    if (warp == kMmaWarp) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int my_tiles = ntiles > (int)blockIdx.x ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int cchunks = p.Cin / 64, nchunks = p.ntaps * cchunks;
    // tile index -> (n tile fastest, m tile, batch): CTAs that run concurrently share activation rows in L2
    auto tile_of = [&](int i, int& n0, int& s0, int& b) {
        const int tile = blockIdx.x + i * gridDim.x;
        const int tn = tile % tiles_n, r = tile / tiles_n;
        n0 = tn * kTileN; s0 = (r % tiles_m) * kTileM; b = r / tiles_m;
    };
    if (warp == kProdWarp) {
        pdl_wait();
        if (lane == 0) {
            int c = 0;
            for (int i = 0; i < my_tiles; ++i) {
                int n0, s0, b;
                tile_of(i, n0, s0, b);
                for (int k = 0; k < nchunks; ++k, ++c) {
                    const int st = c % kStages, n = c / kStages;
                    const int tap = k / cchunks, cc = k - tap * cchunks;
                    mbar_wait(empty(st), (n & 1) ^ 1);
                    mbar_expect_tx(full(st), kStageBytes);
                    tma_load_3d(base + st * kStageBytes, &maps.a, cc * 64, s0 + p.toff[tap], b, full(st));
                    tma_load_2d(base + st * kStageBytes + kBoxBytes, &maps.w, tap * p.Cin + cc * 64, n0, full(st));
                }
            }
        }
    } else if (warp == kMmaWarp) {
        const uint32_t idesc = make_idesc(kTileN);
        int c = 0;
        for (int i = 0; i < my_tiles; ++i) {
            const int g = i % kGroups, ng = i / kGroups;
            mbar_wait(acce(g), (ng & 1) ^ 1);
            tc_fence_after();
            for (int k = 0; k < nchunks; ++k, ++c) {
                const int st = c % kStages, n = c / kStages;
                mbar_wait(full(st), n & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a0 = base + st * kStageBytes;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma(tmem + 128u * g, make_desc_sw128(a0 + kk * 32), make_desc_sw128(a0 + kBoxBytes + kk * 32), idesc, (k | kk) != 0);
                    umma_commit(empty(st));
                    if (k == nchunks - 1) umma_commit(accf(g));
                }
                __syncwarp();
            }
        }
    } else {
        pdl_wait();
        const int g = warp >> 2, wq = warp & 3;
        const uint32_t acc = tmem + 128u * g + ((uint32_t)(wq * 32) << 16);
        // Transposition tile of this warp (32 rows x 32 columns, pitch 36 floats): TMEM hands every thread one ROW, but global
        // memory wants a warp access to cover whole 128-byte row segments (8 lanes x 16 B per row, 4 rows per instruction) —
        // row-per-thread accesses cost 32 L1 wavefronts per instruction.
        const uint32_t tile = base + kOffTile + (uint32_t)warp * 4608u;
        const int rsub = lane >> 3, cl = 4 * (lane & 7);
        for (int i = g, ng = 0; i < my_tiles; i += kGroups, ++ng) {
        int n0, s0t, b;
        tile_of(i, n0, s0t, b);
        const int s0 = s0t + wq * 32 - warp * 32;   // so that s0 + warp * 32 is this warp's first row
        mbar_wait(accf(g), ng & 1);
        tc_fence_after();
#pragma unroll 1
        for (int h = 0; h < 4; ++h) {
            float v[32];
            __syncwarp();                    // the previous chunk's tile reads are done (and the warp is converged for the collective)
            tmem_ld32(acc + 32 * h, v);
            if (h == 3) {                    // the accumulator stage has been read completely: the MMAs of this group's next tile may start
                tc_fence_before();
                mbar_arrive(acce(g));
            }
#pragma unroll
            for (int q = 0; q < 8; ++q)
                sts128(tile + lane * 144 + q * 16, make_uint4(__float_as_uint(v[4 * q]), __float_as_uint(v[4 * q + 1]), __float_as_uint(v[4 * q + 2]),
                                                                  __float_as_uint(v[4 * q + 3])));
            __syncwarp();
            const int j0 = n0 + 32 * h;
            if (j0 >= p.Ntot) {              // beyond the GEMM's N: only the zero padding of a narrower-than-ld16 bf16 tensor remains
                if (p.phases == 1 && j0 < p.ld16) {
                    for (int it = 0; it < 8; ++it) {
                        const int sr = s0 + warp * 32 + it * 4 + rsub;
                        if (sr >= p.L_out || sr >= p.rows) continue;
                        const size_t row = (size_t)b * p.L_out + sr;
                        if (p.raw16) *reinterpret_cast<uint2*>(p.raw16 + row * p.ld16 + j0 + cl) = make_uint2(0, 0);
                        if (p.act16) *reinterpret_cast<uint2*>(p.act16 + row * p.ld16 + j0 + cl) = make_uint2(0, 0);
                    }
                }
                continue;
            }
            const int phase = j0 / p.H, c0 = j0 - phase * p.H + cl;
            const float4 bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + c0));
            float4 pe4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.post_lrelu) pe4 = __ldg(reinterpret_cast<const float4*>(p.pe + (size_t)b * p.pe_stride + c0));
            // all global loads of the chunk are issued before the first use (8 row groups x up to 3 tensors in flight per thread)
            float4 ad[8];
            uint2 sh[8], sc[8];
            bool ok[8];
            size_t rowi[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int sr = s0 + warp * 32 + it * 4 + rsub;
                const int t = p.phases * sr + phase;
                ok[it] = sr < p.rows && t < p.L_out;
                rowi[it] = (size_t)b * p.L_out + t;
                ad[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                sh[it] = sc[it] = make_uint2(0u, 0u);
                if (ok[it]) {
                    if (p.add) ad[it] = __ldg(reinterpret_cast<const float4*>(p.add + ((size_t)b * p.add_rows + t / p.add_div) * p.H + c0));
                    if (p.act_mode == 2) {
                        sh[it] = __ldg(reinterpret_cast<const uint2*>(p.film + rowi[it] * 2 * p.H + c0));
                        sc[it] = __ldg(reinterpret_cast<const uint2*>(p.film + rowi[it] * 2 * p.H + p.H + c0));
                    }
                }
            }
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                if (!ok[it]) continue;
                const size_t row = rowi[it];
                const uint4 u = lds128(tile + (it * 4 + rsub) * 144 + (lane & 7) * 16);
                float4 x = make_float4(__uint_as_float(u.x) + bias4.x, __uint_as_float(u.y) + bias4.y, __uint_as_float(u.z) + bias4.z,
                                       __uint_as_float(u.w) + bias4.w);
                if (p.post_lrelu) x = make_float4(lrelu02(x.x) + pe4.x, lrelu02(x.y) + pe4.y, lrelu02(x.z) + pe4.z, lrelu02(x.w) + pe4.w);
                x.x += ad[it].x; x.y += ad[it].y; x.z += ad[it].z; x.w += ad[it].w;
                if (p.raw32) *reinterpret_cast<float4*>(p.raw32 + row * p.H + c0) = x;
                if (p.raw16) *reinterpret_cast<uint2*>(p.raw16 + row * p.ld16 + c0) = make_uint2(pack_bf16(x.x, x.y), pack_bf16(x.z, x.w));
                if (p.act_mode) {
                    if (p.act_mode == 2)
                        x = make_float4(fmaf(bf16_lo(sc[it].x), x.x, bf16_lo(sh[it].x)), fmaf(bf16_hi(sc[it].x), x.y, bf16_hi(sh[it].x)),
                                        fmaf(bf16_lo(sc[it].y), x.z, bf16_lo(sh[it].y)), fmaf(bf16_hi(sc[it].y), x.w, bf16_hi(sh[it].y)));
                    *reinterpret_cast<uint2*>(p.act16 + row * p.ld16 + c0) =
                        make_uint2(pack_bf16(lrelu02(x.x), lrelu02(x.y)), pack_bf16(lrelu02(x.z), lrelu02(x.w)));
                }
            }
        }
        }
        tc_fence_before();
    }
    __syncthreads();
    pdl_launch_dependents();
    if (warp == kMmaWarp) tmem_dealloc(tmem, 512);
}

}  // namespace

int launch_wg_conv_tc(const WgTcConv& p, cudaStream_t st) {
    if (p.Cin % 64 || p.H % 32 || p.Ntot != p.phases * p.H || p.ntaps < 1 || p.ntaps > 3) { set_error("wavegrad tc: unsupported conv shape (Cin=%d H=%d N=%d taps=%d)", p.Cin, p.H, p.Ntot, p.ntaps); return SDDM_E_INVALID; }
    SDDM_SET_MAX_SMEM(wg_conv_tc_kernel, kSmem);
    static std::map<std::tuple<const void*, const void*, long long, long long, int, int, int, int>, WgMaps> cache;
    const auto key = std::make_tuple((const void*)p.a, (const void*)p.w, p.a_row_pitch, p.a_batch_pitch, p.rows, p.Cin, p.Ntot, p.B * 4 + p.ntaps);
    auto it = cache.find(key);
    if (it == cache.end()) {
        if (cache.size() > 4096) cache.clear();
        WgMaps m;
        int rc = encode_bf16_view(&m.a, p.a, p.Cin, p.a_row_pitch, p.rows, p.a_batch_pitch, p.B, 128);
        if (rc) return rc;
        if ((rc = encode_bf16_view(&m.w, p.w, p.ntaps * p.Cin, (long long)p.ntaps * p.Cin, p.Ntot, 0, 0, 128))) return rc;
        it = cache.emplace(key, m).first;
    }
    const int tiles_n = (p.Ntot + kTileN - 1) / kTileN, tiles_m = (p.rows + kTileM - 1) / kTileM, ntiles = tiles_n * tiles_m * p.B;
    const int grid = ntiles < num_sms() ? ntiles : num_sms();
    SDDM_CUDA_TRY(launch_pdl(wg_conv_tc_kernel, dim3(grid), dim3(kThreads), kSmem, st, it->second, p, tiles_n, tiles_m, ntiles));
    count_launch();
    return SDDM_OK;
}

}  // namespace sddm
