// Row-streaming tcgen05 / TMEM 3x3 convolution for the full-resolution level of UNetModified2 (frame grid H x 128, bf16
// activations): stem (2 -> 32), the 32 / 64 -> 32 ResnetBlock convolutions (+ 1x1 res_conv, + identity residual), the
// nearest-x2 Upsample conv and the final Block (32 -> 1).  These layers hold 35 % of the network's FLOPs in N = Cout = 32
// GEMMs; with pixels as M and the 9 taps as K (conv_tc.cu) every tap re-fetches the 4 KB pixel operand from shared memory for
// a 16-cycle MMA, which caps the tensor pipe at 40 % and saturates shared memory (profiles/r1_umma_rate.txt).
//
// Here the vertical taps move to the N side and are summed from TMEM by the epilogue:
//
//   M tile   = one image row (128 pixels) of one sample                     -> A operand: [k8 plane][1 + 128 + 1 pixels][16 B]
//   N        = (ky, cout) = 3 x 32 = 96                                     -> B operand: resident weights [k16][kx][2][96][8]
//   K        = Cin, horizontal taps kx = shifted start address of the same row operand (x halo = two zero slots)
//   D_r[x][ky][co] = sum_{kx, ci} A_r[x + kx - 1][ci] W[co][ci][ky][kx]      (one accumulator slot of 96 TMEM columns per INPUT row r)
//   out[y][x][co]  = D_{y-1}[x][0][co] + D_y[x][1][co] + D_{y+1}[x][2][co]   (three column-shifted TMEM reads, same lane = same x)
//
// so a pixel operand byte is fetched 3 x (kx) instead of 9 x, each MMA is M128 x N96 x K16 (fetch 56 cycles vs 48 of math,
// against 40 vs 16), every input row is loaded / transformed exactly once per CTA (a CTA owns a contiguous run of 16-row blocks;
// the only halo is one row above / below the run), and there is no im2col anywhere.
//
//   raw load  : TMA (cp.async.bulk.tensor.4d) of one bf16 NHWC row slab [128 px][32 ch] (nearest x2: the [64 px] row y / 2;
//               concat = choice of tensor map), GroupNorm scale / shift of the sample travel with it; stem: two 512-byte bulk
//               copies of the waveform windows of frame r (framing = addressing: frame r = samples [hop r, hop r + 128))
//   transform : GroupNorm-apply + Swish (one tanh.approx) + bf16 -> operand row; stem: fp32 sample -> bf16 hi + lo pair
//               (two K slots each, so the waveform enters the tensor core at 2^-17 relative precision)
//   MMA       : one warp, 3 (kx) x Cin/16 tcgen05.mma per input row, + the 1x1 res_conv over the raw block input as extra K
//               into the centre (ky = 1) columns
//   epilogue  : two groups (even / odd output rows): 3 x tcgen05.ld -> + bias (+ noise embedding + res bias) (+ identity
//               residual row, TMA prefetched) -> bf16 -> 64B-swizzled staging row -> TMA store; GroupNorm partial statistics
//               accumulate in registers over the 8 rows a group owns of a 16-row block; the last (block, group) of a sample
//               finalises the consumer's GroupNorm (gn_fuse.cuh).  Final Block: 3 scalars per pixel -> frames.
//
// reference: Block / ResnetBlock / Upsample / stem / final_conv, model/UNetModified2.py:93-142,177-178,235
#include <cstdlib>
#include <cstring>

#include "gn_fuse.cuh"
#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace sddm {
namespace {

constexpr int RW = 128;                      // row width = MMA M
constexpr int kRowThreads = 20 * 32;         // warps 0-3 / 4-7 epilogue groups, 8-15 transform groups, 16 MMA, 17 weights, 18 raw TMA
constexpr int kEpi = 2, kGrp = 128;
constexpr int kXf0 = 8, kMma = 16, kWld = 17, kTma = 18;
constexpr int PLANE = 134;                   // operand slots (16 B) per k8 plane: 1 + 128 + 1 used; = 6 (mod 8): conflict-free stores
constexpr uint32_t kAStage = 4u * PLANE * 16u;            // 8576 B: one 32-channel operand row
constexpr uint32_t kRawStage = 8192u + 256u;              // bf16 row slab + scale / shift tail
constexpr uint32_t kOutTile = 8192u;                      // 128 px x 32 ch bf16 staging row (64B swizzle)
constexpr int kMaxRing = 8, kMaxSlots = 5;
constexpr size_t kSmemCap = 232448 - 1024;

struct alignas(64) RowMaps {
    CUtensorMap src[2];    // main sources (concat order), box 32 ch x (128 | 64) px
    CUtensorMap rsrc[2];   // raw block input of the 1x1 res_conv (or [0] = identity residual, 64B swizzle)
    CUtensorMap out;       // output rows, 64B swizzle
};

struct RowArgs {
    int B, H, nblocks;          // 16-row blocks in the whole batch
    int stem, final_out, up;
    int n_main, n_res, ksteps;  // 32-channel slabs per input row: main conv / 1x1 res_conv; K16 steps per main slab
    int C0, rC0;                // channels of source 0 (concat boundary) of the main / res inputs
    int Cin, affine;
    const float* scale; const float* shift;     // [B][Cin] GroupNorm of the input
    const __nv_bfloat16* w;                      // packed: main chunks, then res chunks
    uint32_t w_bytes;
    int n_cols;                 // N of the main MMAs: 96 (or 16 for the final Block)
    const float* bias; const float* temb; int temb_stride; const float* res_bias;
    int res_identity;
    const float* cond; const float* x_t; int L, hop;   // stem
    float* parts; int nparts;   // [B][nparts][32][2]
    float* frames; float final_bias;
    int gn_on; GnFuse gn;
    int NR, NA, NS, NOUT, NRES, slot_cols, tmem_cols;
    uint32_t off_w, off_raw, off_a, off_out, off_res;
};

struct RowHdr {
    uint64_t raw_full[kMaxRing], raw_empty[kMaxRing];
    uint64_t full_a[kMaxRing], empty_a[kMaxRing];
    uint64_t acc_full[kMaxSlots], acc_empty[kMaxSlots];
    uint64_t res_full[kEpi][2];
    uint64_t w_full;
    uint32_t tmem_base;
    uint32_t gn_last[kEpi];
    uint32_t pad[5];
    float addv[kEpi][32];
};
constexpr uint32_t kHdr = 1024;
static_assert(sizeof(RowHdr) <= kHdr, "header too large");

// the k-th sample segment of a CTA that owns the 16-row blocks [b0, b1): output rows [ya, yb) of sample n need input rows [r0, r1]
struct Seg { int n, ya, yb, r0, r1; };
__device__ __forceinline__ bool seg_at(int H, int b0, int b1, int k, Seg& s) {
    const int bps = H >> 4, n0 = b0 / bps;
    s.n = n0 + k;
    if (s.n * bps >= b1) return false;
    s.ya = k == 0 ? (b0 - n0 * bps) << 4 : 0;
    const int e = (b1 - s.n * bps) << 4;
    s.yb = e < H ? e : H;
    s.r0 = s.ya > 0 ? s.ya - 1 : 0;
    s.r1 = s.yb < H ? s.yb : H - 1;
    return true;
}

__global__ void __launch_bounds__(kRowThreads, 1) conv_row_kernel(const RowArgs a, const __grid_constant__ RowMaps maps) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t dyn = smem_u32(smem_dyn), base = (dyn + 1023u) & ~1023u;
    RowHdr* hdr = reinterpret_cast<RowHdr*>(smem_dyn + (base - dyn));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nslab = a.n_main + a.n_res;
    pdl_launch_dependents();

    // contiguous run of blocks of this CTA
    const int q = a.nblocks / (int)gridDim.x, rem = a.nblocks - q * (int)gridDim.x;
    const int b0 = (int)blockIdx.x * q + ((int)blockIdx.x < rem ? (int)blockIdx.x : rem);
    const int b1 = b0 + q + ((int)blockIdx.x < rem ? 1 : 0);

    if (tid == 0) {
        for (int i = 0; i < kMaxRing; ++i) {
            mbar_init(smem_u32(&hdr->raw_full[i]), 1); mbar_init(smem_u32(&hdr->raw_empty[i]), kGrp);
            mbar_init(smem_u32(&hdr->full_a[i]), kGrp); mbar_init(smem_u32(&hdr->empty_a[i]), 1);
        }
        for (int i = 0; i < kMaxSlots; ++i) { mbar_init(smem_u32(&hdr->acc_full[i]), 1); mbar_init(smem_u32(&hdr->acc_empty[i]), 12); }
        for (int e = 0; e < kEpi; ++e) for (int k = 0; k < 2; ++k) mbar_init(smem_u32(&hdr->res_full[e][k]), 1);
        mbar_init(smem_u32(&hdr->w_full), 1);
        fence_barrier_init();
    }
    if (warp == kMma) tmem_alloc(smem_u32(&hdr->tmem_base), (uint32_t)a.tmem_cols);
    // operand ring: zero once - the x halo slots (0 and 129) of every plane are never written again (stem: plane 1 stays zero too)
    for (uint32_t i = (uint32_t)tid; i < (uint32_t)a.NA * kAStage / 16u; i += kRowThreads)
        sts128(base + a.off_a + i * 16u, make_uint4(0u, 0u, 0u, 0u));
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = hdr->tmem_base;

    if (warp < 8) {
        // ============================== epilogue: group e owns the output rows with y & 1 == e ===========================
        pdl_wait();
        const int e = warp >> 2, w4 = warp & 3, m = tid & 127;
        const int bar_id = 1 + e;
        const bool leader = m == 0;
        const bool has_res = a.res_identity != 0;
        const uint32_t obuf0 = base + a.off_out + (uint32_t)(e * a.NOUT) * kOutTile;
        const uint32_t rbuf0 = base + a.off_res + (uint32_t)(e * 2) * kOutTile;
        const uint32_t addv_u32 = smem_u32(hdr->addv[e]);
        const uint32_t lane_tm = tmem_base + ((uint32_t)(w4 * 32) << 16);
        const int c2 = lane & 15, hrow = lane >> 4;
        float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;   // statistics of this group's rows of the current block (2 channels per lane)
        int ob = 0;                                          // rows stored so far (staging buffer parity)
        int rq = 0;                                          // residual rows consumed so far
        // residual prefetch runs one row ahead: the (n, y) sequence of this group
        int pk = 0, py = -1, pn = 0, pend = 0, pissued = 0;
        Seg ps{};
        auto res_next = [&]() -> bool {   // advance (pk, py) to this group's next output row; false when exhausted
            for (;;) {
                if (py < 0) {
                    if (!seg_at(a.H, b0, b1, pk, ps)) return false;
                    py = ps.ya + ((ps.ya & 1) == e ? 0 : 1);
                    pend = ps.yb; pn = ps.n;
                } else {
                    py += 2;
                }
                if (py < pend) return true;
                py = -1; ++pk;
            }
        };
        auto res_issue = [&]() {
            if (!res_next()) return;
            const uint32_t bar = smem_u32(&hdr->res_full[e][pissued & 1]);
            mbar_expect_tx(bar, kOutTile);
            tma_load_4d(rbuf0 + (uint32_t)(pissued & 1) * kOutTile, &maps.rsrc[0], 0, 0, py, pn, bar);
            ++pissued;
        };
        if (has_res && leader) res_issue();
        int item_base = 0;
        Seg s;
        for (int k = 0; seg_at(a.H, b0, b1, k, s); item_base += s.r1 - s.r0 + 1, ++k) {
            // per-channel additive term of this sample: bias (+ noise-level embedding row) (+ res_conv bias)
            group_bar(bar_id);
            if (m < 32 && !a.final_out) {
                float v = __ldg(a.bias + m);
                if (a.temb) v += __ldg(a.temb + (int64_t)s.n * a.temb_stride + m);
                if (a.n_res) v += __ldg(a.res_bias + m);
                hdr->addv[e][m] = v;
            }
            group_bar(bar_id);
            for (int y = s.ya + ((s.ya & 1) == e ? 0 : 1); y < s.yb; y += 2) {
                const bool vA = y - 1 >= s.r0, vC = y + 1 <= s.r1;
                const int iB = item_base + (y - s.r0), iA = iB - 1, iC = iB + 1;
                const int last = vC ? iC : iB;
                mbar_wait(smem_u32(&hdr->acc_full[last % a.NS]), (uint32_t)(last / a.NS) & 1u);
                tc_fence_after();
                // arrivals owed to the three accumulator slots (see header: rows outside [ya, yb) never arrive)
                const int first = y == s.ya, lastrow = y == s.yb - 1;
                // (an input row r is read by the output rows r - 1, r, r + 1; those outside [ya, yb) never come, so the first / last
                //  row of the segment arrives in their place: every slot sees exactly 3 arrivals x 4 warps)
                const uint32_t nA = first ? 3u : 1u, nB = 1u + (first ? 1u : 0u) + (lastrow ? 1u : 0u), nC = lastrow ? 3u : 1u;
                if (a.final_out) {
                    uint32_t ra[4] = {0u, 0u, 0u, 0u}, rb[4], rc[4] = {0u, 0u, 0u, 0u};
                    if (vA) tmem_ld4_nowait(lane_tm + (uint32_t)((iA % a.NS) * a.slot_cols), ra);
                    tmem_ld4_nowait(lane_tm + (uint32_t)((iB % a.NS) * a.slot_cols), rb);
                    if (vC) tmem_ld4_nowait(lane_tm + (uint32_t)((iC % a.NS) * a.slot_cols), rc);
                    tmem_wait_ld();
                    tc_fence_before();
                    if (lane == 0) {
                        if (vA) mbar_arrive_n(smem_u32(&hdr->acc_empty[iA % a.NS]), nA);
                        mbar_arrive_n(smem_u32(&hdr->acc_empty[iB % a.NS]), nB);
                        if (vC) mbar_arrive_n(smem_u32(&hdr->acc_empty[iC % a.NS]), nC);
                    }
                    const float f = (__uint_as_float(ra[0]) + __uint_as_float(rb[1])) + __uint_as_float(rc[2]) + a.final_bias;
                    a.frames[((int64_t)s.n * a.H + y) * RW + m] = f;
                    continue;
                }
                float v[32];
                {
                    uint32_t r0[32], r1[32];
                    tmem_ld32_nowait(lane_tm + (uint32_t)((iB % a.NS) * a.slot_cols + 32), r0);
                    if (vA) tmem_ld32_nowait(lane_tm + (uint32_t)((iA % a.NS) * a.slot_cols), r1);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r0[i]) + (vA ? __uint_as_float(r1[i]) : 0.f);
                    if (vC) {
                        tmem_ld32_nowait(lane_tm + (uint32_t)((iC % a.NS) * a.slot_cols + 64), r0);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] += __uint_as_float(r0[i]);
                    }
                }
                tc_fence_before();
                if (lane == 0) {
                    if (vA) mbar_arrive_n(smem_u32(&hdr->acc_empty[iA % a.NS]), nA);
                    mbar_arrive_n(smem_u32(&hdr->acc_empty[iB % a.NS]), nB);
                    if (vC) mbar_arrive_n(smem_u32(&hdr->acc_empty[iC % a.NS]), nC);
                }
#pragma unroll
                for (int qd = 0; qd < 8; ++qd) {
                    const uint4 u = lds128(addv_u32 + (uint32_t)qd * 16u);
                    v[4 * qd + 0] += __uint_as_float(u.x); v[4 * qd + 1] += __uint_as_float(u.y);
                    v[4 * qd + 2] += __uint_as_float(u.z); v[4 * qd + 3] += __uint_as_float(u.w);
                }
                const uint32_t obuf = obuf0 + (uint32_t)(a.NOUT == 2 ? (ob & 1) : 0) * kOutTile;
                if (leader) {   // the TMA store that last read this staging buffer is done with it
                    if (a.NOUT == 2) bulk_wait_read_1(); else bulk_wait_read_0();
                }
                if (has_res) {
                    mbar_wait(smem_u32(&hdr->res_full[e][rq & 1]), (uint32_t)(rq >> 1) & 1u);
                    const uint32_t rrow = rbuf0 + (uint32_t)(rq & 1) * kOutTile + (uint32_t)m * 64u;
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd) {
                        const uint4 u = lds128(rrow + (uint32_t)((qd ^ ((m >> 1) & 3)) << 4));
                        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) { v[8 * qd + 2 * kk] += bf16_lo(w[kk]); v[8 * qd + 2 * kk + 1] += bf16_hi(w[kk]); }
                    }
                    ++rq;
                }
                group_bar(bar_id);   // staging buffer free (leader waited), residual buffer fully read by the whole group
                if (has_res && leader) res_issue();
                const uint32_t orow = obuf + (uint32_t)m * 64u;
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    uint4 u;
                    u.x = pack_bf16(v[8 * qd + 0], v[8 * qd + 1]); u.y = pack_bf16(v[8 * qd + 2], v[8 * qd + 3]);
                    u.z = pack_bf16(v[8 * qd + 4], v[8 * qd + 5]); u.w = pack_bf16(v[8 * qd + 6], v[8 * qd + 7]);
                    sts128(orow + (uint32_t)((qd ^ ((m >> 1) & 3)) << 4), u);
                }
                fence_async_smem();
                group_bar(bar_id);
                if (leader) {
                    tma_store_4d(&maps.out, obuf, 0, 0, y, s.n);
                    bulk_commit();
                }
                ++ob;
                // column sums of the staged (rounded) row: 2 channels per lane, even / odd pixels per half-warp
                {
                    const uint32_t qbase = obuf + (uint32_t)(w4 * 32 + hrow) * 64u + (uint32_t)(c2 & 3) * 4u;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        uint32_t wv;
                        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(wv) : "r"(qbase + (uint32_t)i * 128u + (uint32_t)(((c2 >> 2) ^ (i & 3)) << 4)));
                        const float fx = bf16_lo(wv), fy = bf16_hi(wv);
                        s1a += fx; s1b += fy;
                        s2a = fmaf(fx, fx, s2a); s2b = fmaf(fy, fy, s2b);
                    }
                }
                if ((y & 15) >= 14) {   // this group's last row of the 16-row block: publish its partial, maybe finalise the sample
                    s1a += __shfl_xor_sync(0xffffffffu, s1a, 16); s1b += __shfl_xor_sync(0xffffffffu, s1b, 16);
                    s2a += __shfl_xor_sync(0xffffffffu, s2a, 16); s2b += __shfl_xor_sync(0xffffffffu, s2b, 16);
                    if (hrow == 0) {
                        float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float2*>(a.parts) +
                                                                ((int64_t)s.n * a.nparts + (y >> 4) * 8 + e * 4 + w4) * 32 + 2 * c2);
                        *dst = make_float4(s1a, s2a, s1b, s2b);
                    }
                    s1a = s1b = s2a = s2b = 0.f;
                    if (a.gn_on) {
                        __threadfence();
                        group_bar(bar_id);
                        if (leader) hdr->gn_last[e] = atomicAdd(a.gn.counter + s.n, 1u) == (unsigned)(a.gn.expect - 1) ? 1u : 0u;
                        group_bar(bar_id);
                        if (hdr->gn_last[e]) {
                            __threadfence();
                            gn_fused_finalize(a.gn, s.n, m, kGrp);
                            if (leader) a.gn.counter[s.n] = 0u;
                        }
                    }
                }
            }
        }
        if (leader) bulk_wait_all();
    } else if (warp == kMma) {
        // ============================== MMA issuer ==========================================================================
        const uint32_t idesc = make_idesc(a.n_cols), idesc_res = make_idesc(32);
        const uint32_t w_chunk = (uint32_t)a.n_cols * 32u;                    // bytes per (k16, kx) weight chunk: 2 halves x N rows x 16 B
        const uint64_t a_desc0 = make_desc_nosw(base + a.off_a, (uint32_t)PLANE * 16u, 128u);
        const uint64_t w_desc0 = make_desc_nosw(base + a.off_w, (uint32_t)a.n_cols * 16u, 128u);
        const uint64_t wr_desc0 = make_desc_nosw(base + a.off_w + (uint32_t)(a.n_main * a.ksteps * 3) * w_chunk, 32u * 16u, 128u);
        mbar_wait(smem_u32(&hdr->w_full), 0u);
        tc_fence_after();
        int it = 0, sa = 0;
        uint32_t pa = 0;
        Seg s;
        for (int k = 0; seg_at(a.H, b0, b1, k, s); ++k) {
            for (int r = s.r0; r <= s.r1; ++r, ++it) {
                const int slot = it % a.NS;
                mbar_wait(smem_u32(&hdr->acc_empty[slot]), ((uint32_t)(it / a.NS) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(slot * a.slot_cols);
                uint32_t acc = 0;
                for (int sl = 0; sl < nslab; ++sl) {
                    mbar_wait(smem_u32(&hdr->full_a[sa]), pa);
                    tc_fence_after();
                    const uint64_t adesc = a_desc0 + (uint64_t)((uint32_t)sa * (kAStage >> 4));
                    if (elect_one()) {
                        if (sl < a.n_main) {
                            for (int h = 0; h < a.ksteps; ++h)
#pragma unroll
                                for (int kx = 0; kx < 3; ++kx) {
                                    umma(d_tmem, adesc + (uint64_t)(uint32_t)(h * 2 * PLANE + kx),
                                         w_desc0 + (uint64_t)((uint32_t)((sl * a.ksteps + h) * 3 + kx) * (w_chunk >> 4)), idesc, acc);
                                    acc = 1;
                                }
                        } else {   // 1x1 res_conv over the raw block input: centre tap, centre (ky = 1) columns
#pragma unroll
                            for (int h = 0; h < 2; ++h)
                                umma(d_tmem + 32u, adesc + (uint64_t)(uint32_t)(h * 2 * PLANE + 1),
                                     wr_desc0 + (uint64_t)((uint32_t)((sl - a.n_main) * 2 + h) * (1024u >> 4)), idesc_res, 1u);
                        }
                        umma_commit(smem_u32(&hdr->empty_a[sa]));
                    }
                    acc = 1;
                    if (++sa == a.NA) { sa = 0; pa ^= 1u; }
                }
                if (elect_one()) umma_commit(smem_u32(&hdr->acc_full[slot]));
            }
        }
    } else if (warp == kWld) {
        if (lane == 0) {   // weights: one bulk copy, resident for the whole CTA
            const uint32_t bar = smem_u32(&hdr->w_full);
            mbar_expect_tx(bar, a.w_bytes);
            bulk_g2s(base + a.off_w, a.w, a.w_bytes, bar);
        }
    } else if (warp == kTma) {
        // ============================== raw row loader (one thread) =========================================================
        if (lane == 0) {
            pdl_wait();
            int rs = 0;
            uint32_t pr = 0;
            Seg s;
            for (int k = 0; seg_at(a.H, b0, b1, k, s); ++k)
                for (int r = s.r0; r <= s.r1; ++r)
                    for (int sl = 0; sl < nslab; ++sl) {
                        const uint32_t bar = smem_u32(&hdr->raw_full[rs]);
                        mbar_wait(smem_u32(&hdr->raw_empty[rs]), pr ^ 1u);
                        const uint32_t dst = base + a.off_raw + (uint32_t)rs * kRawStage;
                        if (a.stem) {
                            mbar_expect_tx(bar, 1024u);
                            bulk_g2s(dst, a.cond + (int64_t)s.n * a.L + (int64_t)r * a.hop, 512u, bar);
                            bulk_g2s(dst + 512u, a.x_t + (int64_t)s.n * a.L + (int64_t)r * a.hop, 512u, bar);
                        } else if (sl < a.n_main) {
                            const int cb = sl * 32, si = cb < a.C0 ? 0 : 1;
                            const bool aff = a.affine != 0;
                            mbar_expect_tx(bar, (a.up ? 4096u : 8192u) + (aff ? 256u : 0u));
                            tma_load_4d(dst, &maps.src[si], cb - (si ? a.C0 : 0), 0, a.up ? (r >> 1) : r, s.n, bar);
                            if (aff) {
                                bulk_g2s(dst + 8192u, a.scale + (int64_t)s.n * a.Cin + cb, 128u, bar);
                                bulk_g2s(dst + 8192u + 128u, a.shift + (int64_t)s.n * a.Cin + cb, 128u, bar);
                            }
                        } else {
                            const int cb = (sl - a.n_main) * 32, si = cb < a.rC0 ? 0 : 1;
                            mbar_expect_tx(bar, 8192u);
                            tma_load_4d(dst, &maps.rsrc[si], cb - (si ? a.rC0 : 0), 0, r, s.n, bar);
                        }
                        if (++rs == a.NR) { rs = 0; pr ^= 1u; }
                    }
        }
    } else if (warp >= kXf0 && warp < kXf0 + 8) {
        // ============================== transform: raw row slab -> bf16 operand row ========================================
        // two groups of 128 threads take alternate slabs of the CTA's (row, slab) sequence
        const int gi = (warp - kXf0) >> 2, gt = tid - (kXf0 * 32 + gi * kGrp);
        const int j = gt & 3, px0 = gt >> 2;
        int qn = 0;   // global slab counter
        Seg s;
        for (int k = 0; seg_at(a.H, b0, b1, k, s); ++k)
            for (int r = s.r0; r <= s.r1; ++r)
                for (int sl = 0; sl < nslab; ++sl, ++qn) {
                    if ((qn & 1) != gi) continue;
                    const int rs = qn % a.NR, sa = qn % a.NA;
                    const uint32_t raw = base + a.off_raw + (uint32_t)rs * kRawStage;
                    const uint32_t opd = base + a.off_a + (uint32_t)sa * kAStage;
                    mbar_wait(smem_u32(&hdr->raw_full[rs]), (uint32_t)(qn / a.NR) & 1u);
                    mbar_wait(smem_u32(&hdr->empty_a[sa]), ((uint32_t)(qn / a.NA) & 1u) ^ 1u);
                    if (a.stem) {
                        // K slots of plane 0: [cond hi, x_t hi, cond lo, x_t lo, 0, 0, 0, 0]
                        const float c = lds32(raw + (uint32_t)gt * 4u), x = lds32(raw + 512u + (uint32_t)gt * 4u);
                        const __nv_bfloat16 ch = __float2bfloat16_rn(c), xh = __float2bfloat16_rn(x);
                        const float cl = c - __bfloat162float(ch), xl = x - __bfloat162float(xh);
                        uint4 o;
                        o.x = (uint32_t)__bfloat16_as_ushort(ch) | ((uint32_t)__bfloat16_as_ushort(xh) << 16);
                        o.y = pack_bf16(cl, xl);
                        o.z = 0u; o.w = 0u;
                        sts128(opd + (uint32_t)(gt + 1) * 16u, o);
                    } else {
                        const bool aff = a.affine != 0 && sl < a.n_main;
                        const bool up = a.up != 0 && sl < a.n_main;
                        float sch[8], shh[8];
                        if (aff) {   // halved: swish(y) = h + h tanh(h), h = y / 2
                            const uint32_t ss = raw + 8192u + (uint32_t)j * 32u;
                            const uint4 s0 = lds128(ss), s1 = lds128(ss + 16u), h0 = lds128(ss + 128u), h1 = lds128(ss + 144u);
                            sch[0] = 0.5f * __uint_as_float(s0.x); sch[1] = 0.5f * __uint_as_float(s0.y); sch[2] = 0.5f * __uint_as_float(s0.z); sch[3] = 0.5f * __uint_as_float(s0.w);
                            sch[4] = 0.5f * __uint_as_float(s1.x); sch[5] = 0.5f * __uint_as_float(s1.y); sch[6] = 0.5f * __uint_as_float(s1.z); sch[7] = 0.5f * __uint_as_float(s1.w);
                            shh[0] = 0.5f * __uint_as_float(h0.x); shh[1] = 0.5f * __uint_as_float(h0.y); shh[2] = 0.5f * __uint_as_float(h0.z); shh[3] = 0.5f * __uint_as_float(h0.w);
                            shh[4] = 0.5f * __uint_as_float(h1.x); shh[5] = 0.5f * __uint_as_float(h1.y); shh[6] = 0.5f * __uint_as_float(h1.z); shh[7] = 0.5f * __uint_as_float(h1.w);
                        }
                        const int rounds = up ? 2 : 4;
                        uint4 rv[4];
#pragma unroll
                        for (int rd = 0; rd < 4; ++rd)
                            if (rd < rounds) rv[rd] = lds128(raw + (uint32_t)(px0 + 32 * rd) * 64u + (uint32_t)j * 16u);
#pragma unroll
                        for (int rd = 0; rd < 4; ++rd) {
                            if (rd >= rounds) break;
                            uint4 o = rv[rd];
                            if (aff) {
                                const uint32_t w[4] = {o.x, o.y, o.z, o.w};
                                uint32_t ow[4];
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk) {
                                    const float h0 = fmaf(bf16_lo(w[kk]), sch[2 * kk], shh[2 * kk]);
                                    const float h1 = fmaf(bf16_hi(w[kk]), sch[2 * kk + 1], shh[2 * kk + 1]);
                                    ow[kk] = pack_bf16(fmaf(h0, tanh_approx(h0), h0), fmaf(h1, tanh_approx(h1), h1));
                                }
                                o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                            }
                            const int px = px0 + 32 * rd;
                            const uint32_t dstp = opd + (uint32_t)j * (uint32_t)PLANE * 16u;
                            if (up) {
                                sts128(dstp + (uint32_t)(2 * px + 1) * 16u, o);
                                sts128(dstp + (uint32_t)(2 * px + 2) * 16u, o);
                            } else {
                                sts128(dstp + (uint32_t)(px + 1) * 16u, o);
                            }
                        }
                    }
                    fence_async_smem();
                    mbar_arrive(smem_u32(&hdr->full_a[sa]));
                    mbar_arrive(smem_u32(&hdr->raw_empty[rs]));
                }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMma) {
        __syncwarp();
        tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
    }
}

// NHWC bf16 tensor [B][H][W][C] as a 4-D tensor map (C innermost) with box (32, bw, 1, 1); swz 0 / 64
int encode_rows(CUtensorMap* m, const void* basep, int B, int H, int W, int C, int bw, int swz) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("conv row: cuTensorMapEncodeTiled is unavailable"); return SDDM_E_CUDA; }
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    const cuuint32_t box[4] = {32, (cuuint32_t)bw, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(basep), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv row: cuTensorMapEncodeTiled failed (%d) for [%d,%d,%d,%d]", (int)r, B, H, W, C); return SDDM_E_CUDA; }
    return SDDM_OK;
}

int device_sms() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        n = 148;
    }
    return n;
}

}  // namespace

int conv_row_nparts(int H) { return (H / 16) * 8; }
int conv_row_arrivals(int H) { return (H / 16) * 2; }

bool conv_row_supported(const ConvP& p) {
    if (!p.act16 || p.Wout != RW || p.Hout % 16 || p.Hout < 16) return false;
    if (p.Cout != 32 && !(p.Cout == 1)) return false;
    if (p.mode != CONV_S1 && p.mode != CONV_UP) return false;
    if (p.Cin != 32 && p.Cin != 64) return false;
    if (p.Cout == 1 && (p.Cin != 32 || p.mode != CONV_S1 || p.res_Cin || p.res_identity)) return false;
    for (int i = 0; i < p.nsrc; ++i)
        if (p.src[i].C % 32) return false;
    if (p.res_Cin && !p.res_identity && (p.res_Cin != 64 || p.mode != CONV_S1)) return false;
    if (p.res_identity && (p.Cin != 32 || p.Cout != 32 || p.mode != CONV_S1)) return false;
    return true;
}

// shared-memory plan + tensor maps + launch
static int launch_row(RowArgs a, RowMaps& maps, cudaStream_t st) {
    a.nblocks = a.B * (a.H / 16);
    a.slot_cols = a.final_out ? 32 : 96;
    a.NS = 5;
    a.tmem_cols = a.final_out ? 256 : 512;
    a.NOUT = a.final_out ? 0 : 2;
    a.NRES = a.res_identity ? 2 : 0;
    const size_t w_al = ((size_t)a.w_bytes + 1023) & ~(size_t)1023;
    const size_t fixed = kHdr + w_al + (size_t)kEpi * (a.NOUT + a.NRES) * kOutTile;
    a.NR = 4; a.NA = 4;
    for (bool grew = true; grew;) {
        grew = false;
        if (a.NR < kMaxRing && fixed + (size_t)(a.NR + 1) * kRawStage + (size_t)a.NA * kAStage <= kSmemCap) { ++a.NR; grew = true; }
        if (a.NA < kMaxRing && fixed + (size_t)a.NR * kRawStage + (size_t)(a.NA + 1) * kAStage <= kSmemCap) { ++a.NA; grew = true; }
    }
    if (fixed + (size_t)a.NR * kRawStage + (size_t)a.NA * kAStage > kSmemCap) { set_error("conv row: shared-memory plan does not fit"); return SDDM_E_INVALID; }
    a.off_w = kHdr;
    a.off_out = a.off_w + (uint32_t)w_al;                                     // 1024-aligned (swizzled TMA tiles)
    a.off_res = a.off_out + (uint32_t)(kEpi * a.NOUT) * kOutTile;
    a.off_raw = a.off_res + (uint32_t)(kEpi * a.NRES) * kOutTile;
    a.off_a = a.off_raw + (uint32_t)a.NR * kRawStage;
    const size_t smem = a.off_a + (size_t)a.NA * kAStage + 1024;
    SDDM_CUDA_TRY(cudaFuncSetAttribute(conv_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemCap + 1024)));   // per device: set every time
    const int sms = device_sms();
    const int grid = a.nblocks < sms ? a.nblocks : sms;
    SDDM_CUDA_TRY(launch_pdl(conv_row_kernel, dim3(grid), dim3(kRowThreads), smem, st, a, maps));
    SDDM_LAUNCH_CHECK();
    return SDDM_OK;
}

// ResnetBlock / Upsample convolutions (Cout = 32) and the final Block (Cout = 1, frames out) on a 128-wide level
int launch_conv_row(const ConvP& p, const __nv_bfloat16* w_row, uint32_t w_bytes, float* frames, float final_bias, cudaStream_t st) {
    if (!conv_row_supported(p) || !w_row) { set_error("conv row: unsupported shape Cin=%d Cout=%d mode=%d W=%d", p.Cin, p.Cout, p.mode, p.Wout); return SDDM_E_INVALID; }
    RowArgs a{};
    RowMaps maps;
    memset(&maps, 0, sizeof(maps));
    a.B = p.B; a.H = p.Hout;
    a.final_out = p.Cout == 1;
    a.up = p.mode == CONV_UP;
    a.n_main = p.Cin / 32;
    a.ksteps = 2;
    const bool res_conv = p.res_Cin && !p.res_identity;
    a.n_res = res_conv ? p.res_Cin / 32 : 0;
    a.C0 = p.src[0].C; a.rC0 = p.res_src[0].C;
    a.Cin = p.Cin;
    a.affine = p.src[0].scale != nullptr;
    a.scale = p.src[0].scale; a.shift = p.src[0].shift;
    a.w = w_row; a.w_bytes = w_bytes;
    a.n_cols = a.final_out ? 16 : 96;
    a.bias = p.bias; a.temb = p.temb; a.temb_stride = p.temb_stride; a.res_bias = p.res_bias;
    a.res_identity = p.res_identity;
    a.parts = p.parts; a.nparts = p.nparts;
    a.frames = frames; a.final_bias = final_bias;
    a.gn_on = p.gn_on; a.gn = p.gn;
    if (!a.final_out && (!p.parts || p.nparts != conv_row_nparts(p.Hout))) { set_error("conv row: nparts mismatch"); return SDDM_E_INVALID; }
    int rc;
    for (int i = 0; i < p.nsrc; ++i)
        if ((rc = encode_rows(&maps.src[i], p.src[i].x, p.B, p.Hin, p.Win, p.src[i].C, a.up ? 64 : 128, 0))) return rc;
    if (res_conv)
        for (int i = 0; i < p.res_nsrc; ++i)
            if ((rc = encode_rows(&maps.rsrc[i], p.res_src[i].x, p.B, p.Hout, RW, p.res_src[i].C, 128, 0))) return rc;
    if (p.res_identity && (rc = encode_rows(&maps.rsrc[0], p.res_src[0].x, p.B, p.Hout, RW, 32, 128, 64))) return rc;
    if (!a.final_out && (rc = encode_rows(&maps.out, p.out, p.B, p.Hout, RW, 32, 128, 64))) return rc;
    return launch_row(a, maps, st);
}

// stem: SignalToFrames x 2 + cat + conv3x3(2 -> 32)
int launch_stem_row(const StemP& p, const __nv_bfloat16* w_row, uint32_t w_bytes, cudaStream_t st) {
    if (!p.act16 || p.W != RW || p.H % 16 || p.CO != 32 || !w_row) { set_error("stem row: unsupported shape"); return SDDM_E_INVALID; }
    if (p.nparts != conv_row_nparts(p.H)) { set_error("stem row: nparts mismatch"); return SDDM_E_INVALID; }
    RowArgs a{};
    RowMaps maps;
    memset(&maps, 0, sizeof(maps));
    a.B = p.B; a.H = p.H;
    a.stem = 1;
    a.n_main = 1; a.ksteps = 1; a.n_res = 0;
    a.Cin = 2;
    a.w = w_row; a.w_bytes = w_bytes;
    a.n_cols = 96;
    a.bias = p.bias;
    a.cond = p.cond; a.x_t = p.x_t; a.L = p.L; a.hop = p.hop;
    a.parts = p.parts; a.nparts = p.nparts;
    a.gn_on = p.gn_on; a.gn = p.gn;
    int rc;
    if ((rc = encode_rows(&maps.out, p.out, p.B, p.H, RW, 32, 128, 64))) return rc;
    return launch_row(a, maps, st);
}

}  // namespace sddm
