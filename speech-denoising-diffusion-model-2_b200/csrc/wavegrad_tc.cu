// tcgen05 / TMEM / TMA convolution of the WaveGrad denoiser (cfg 4; reference model/wavegrad.py:52-137), sm_100a.
//
// Every Conv1d is a GEMM tile [128 time rows] x [128 output columns] accumulated in TMEM over (tap, 64-channel chunk) steps.
// Both operands arrive by TMA in the 128-byte-swizzled K-major layout (3-stage ring, 32 KB per stage): the activation box is
// a window of a time-major bf16 tensor shifted by the tap offset (zero filled outside = the padding), the weight box a slice
// of the pre-packed [N][K] matrix.  Everything the reference applies between two convs (leaky_relu, the FiLM affine,
// nearest-neighbour resampling) is moved to the producer's epilogue or into the view / the polyphase weight packing, so
// the main loop is pure TMA -> tcgen05.mma.  Two CTAs per SM: one runs its epilogue while the other streams its K loop.
#include <cstring>
#include <map>
#include <tuple>

#include "tc_ptx.cuh"
#include "wavegrad.cuh"

namespace sddm {
namespace {

constexpr int kTileM = 128, kTileN = 128;
constexpr uint32_t kBoxBytes = 128 * 64 * 2;          // 16 KB
constexpr int kStages = 3;
constexpr uint32_t kStageBytes = 2 * kBoxBytes;
constexpr uint32_t kOffBars = kStages * kStageBytes;
constexpr uint32_t kSmem = kOffBars + 128 + 1024;

struct alignas(64) WgMaps {
    CUtensorMap a, w;
};

__device__ __forceinline__ float lrelu02(float v) { return v > 0.f ? v : 0.2f * v; }

__global__ void __launch_bounds__(192, 2) wg_conv_tc_kernel(const __grid_constant__ WgMaps maps, WgTcConv p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* gbase = smem_raw + (base - smem_u32(smem_raw));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bars = base + kOffBars;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + kOffBars + 64);
    auto full = [&](int s) { return bars + 8u * (uint32_t)s; };
    auto empty = [&](int s) { return bars + 8u * (uint32_t)(kStages + s); };
    const uint32_t accf = bars + 8u * 2 * kStages;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        mbar_init(accf, 1);
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(smem_u32(tmem_slot), 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int n0 = blockIdx.x * kTileN, s0 = blockIdx.y * kTileM, b = blockIdx.z;
    const int cchunks = p.Cin / 64, nchunks = p.ntaps * cchunks;
    if (warp == 4) {
        pdl_wait();
        if (lane == 0) {
            for (int c = 0; c < nchunks; ++c) {
                const int st = c % kStages, n = c / kStages;
                const int tap = c / cchunks, cc = c - tap * cchunks;
                mbar_wait(empty(st), (n & 1) ^ 1);
                mbar_expect_tx(full(st), kStageBytes);
                tma_load_3d(base + st * kStageBytes, &maps.a, cc * 64, s0 + p.toff[tap], b, full(st));
                tma_load_2d(base + st * kStageBytes + kBoxBytes, &maps.w, tap * p.Cin + cc * 64, n0, full(st));
            }
        }
    } else if (warp == 5) {
        const uint32_t idesc = make_idesc(kTileN);
        for (int c = 0; c < nchunks; ++c) {
            const int st = c % kStages, n = c / kStages;
            mbar_wait(full(st), n & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a0 = base + st * kStageBytes;
#pragma unroll
                for (int k = 0; k < 4; ++k) umma(tmem, make_desc_sw128(a0 + k * 32), make_desc_sw128(a0 + kBoxBytes + k * 32), idesc, (c | k) != 0);
                umma_commit(empty(st));
                if (c == nchunks - 1) umma_commit(accf);
            }
            __syncwarp();
        }
    } else {
        pdl_wait();
        const int s = s0 + warp * 32 + lane;
        const uint32_t acc = tmem + ((uint32_t)(warp * 32) << 16);
        mbar_wait(accf, 0);
        tc_fence_after();
#pragma unroll 1
        for (int h = 0; h < 4; ++h) {
            float v[32];
            __syncwarp();                    // reconverge after the guarded stores of the previous chunk
            tmem_ld32(acc + 32 * h, v);      // warp-collective: every lane takes part, stores are guarded below
            const int j0 = n0 + 32 * h;
            if (j0 >= p.Ntot) {              // beyond the GEMM's N: only the zero padding of a narrower-than-ld16 bf16 tensor remains
                if (p.phases == 1 && j0 < p.ld16 && s < p.L_out) {
                    const size_t row = (size_t)b * p.L_out + s;
                    const uint4 z = make_uint4(0, 0, 0, 0);
                    for (int q = 0; q < 4; ++q) {
                        if (p.raw16) *reinterpret_cast<uint4*>(p.raw16 + row * p.ld16 + j0 + 8 * q) = z;
                        if (p.act16) *reinterpret_cast<uint4*>(p.act16 + row * p.ld16 + j0 + 8 * q) = z;
                    }
                }
                continue;
            }
            const int phase = j0 / p.H, c0 = j0 - phase * p.H;
            const int t = p.phases * s + phase;
            if (t >= p.L_out || s >= p.rows) continue;
            const size_t row = (size_t)b * p.L_out + t;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += __ldg(p.bias + c0 + i);
            if (p.post_lrelu) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = lrelu02(v[i]) + __ldg(p.pe + (size_t)b * p.pe_stride + c0 + i);
            }
            if (p.add) {
                const float4* ap = reinterpret_cast<const float4*>(p.add + ((size_t)b * p.add_rows + t / p.add_div) * p.H + c0);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 a = __ldg(ap + q);
                    v[4 * q] += a.x; v[4 * q + 1] += a.y; v[4 * q + 2] += a.z; v[4 * q + 3] += a.w;
                }
            }
            if (p.raw32) {
                float4* o = reinterpret_cast<float4*>(p.raw32 + row * p.H + c0);
#pragma unroll
                for (int q = 0; q < 8; ++q) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
            if (p.raw16) {
                uint4* o = reinterpret_cast<uint4*>(p.raw16 + row * p.ld16 + c0);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    o[q] = make_uint4(pack_bf16(v[8 * q], v[8 * q + 1]), pack_bf16(v[8 * q + 2], v[8 * q + 3]), pack_bf16(v[8 * q + 4], v[8 * q + 5]),
                                      pack_bf16(v[8 * q + 6], v[8 * q + 7]));
            }
            if (p.act_mode) {
                if (p.act_mode == 2) {
                    const float4* sh = reinterpret_cast<const float4*>(p.film + row * 2 * p.H + c0);
                    const float4* sc = reinterpret_cast<const float4*>(p.film + row * 2 * p.H + p.H + c0);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 a = __ldg(sh + q), m = __ldg(sc + q);
                        v[4 * q] = fmaf(m.x, v[4 * q], a.x); v[4 * q + 1] = fmaf(m.y, v[4 * q + 1], a.y);
                        v[4 * q + 2] = fmaf(m.z, v[4 * q + 2], a.z); v[4 * q + 3] = fmaf(m.w, v[4 * q + 3], a.w);
                    }
                }
                uint4* o = reinterpret_cast<uint4*>(p.act16 + row * p.ld16 + c0);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    o[q] = make_uint4(pack_bf16(lrelu02(v[8 * q]), lrelu02(v[8 * q + 1])), pack_bf16(lrelu02(v[8 * q + 2]), lrelu02(v[8 * q + 3])),
                                      pack_bf16(lrelu02(v[8 * q + 4]), lrelu02(v[8 * q + 5])), pack_bf16(lrelu02(v[8 * q + 6]), lrelu02(v[8 * q + 7])));
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    pdl_launch_dependents();
    if (warp == 5) tmem_dealloc(tmem, 128);
}

}  // namespace

int launch_wg_conv_tc(const WgTcConv& p, cudaStream_t st) {
    if (p.Cin % 64 || p.H % 32 || p.Ntot != p.phases * p.H || p.ntaps < 1 || p.ntaps > 3) { set_error("wavegrad tc: unsupported conv shape (Cin=%d H=%d N=%d taps=%d)", p.Cin, p.H, p.Ntot, p.ntaps); return SDDM_E_INVALID; }
    static bool attr = false;
    if (!attr) {
        SDDM_CUDA_TRY(cudaFuncSetAttribute(wg_conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
        attr = true;
    }
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
    dim3 grid((p.Ntot + kTileN - 1) / kTileN, (p.rows + kTileM - 1) / kTileM, p.B);
    // a narrower-than-ld16 output also needs its padding columns written: they live in the same (only) N tile
    SDDM_CUDA_TRY(launch_pdl(wg_conv_tc_kernel, grid, dim3(192), kSmem, st, it->second, p));
    count_launch();
    return SDDM_OK;
}

}  // namespace sddm
