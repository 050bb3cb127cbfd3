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
// PAIR variant (64-wide level, Cout = 32 ResnetBlock convolutions): the M tile is one image row of TWO samples; see RowRings and the
// transform for the three kx-shifted operand copies that keep the seam between the samples zero.
//
// Ring rule: the two transform groups take alternate slabs, so the raw ring depth is EVEN (a stage always belongs to one group); with an
// odd depth a group could pass a parity wait on a stage whose current use - the other group's - was still loading.
//
// reference: Block / ResnetBlock / Upsample / stem / final_conv, model/UNetModified2.py:93-142,177-178,235
#include <cstdlib>
#include <cstring>

#include "gn_fuse.cuh"
#include "kernels.cuh"

#define SDDM_MBAR_TIMEOUT_NOTES 1   // waits of this file that time out leave a note (sddm_debug_hang)
#include "tc_ptx.cuh"

namespace sddm {
namespace {

#ifndef SDDM_ROW_TRACE
#define SDDM_ROW_TRACE 0
#endif

constexpr int RW = 128;                      // row width = MMA M
constexpr int kRowThreads = 20 * 32;         // warps 0-3 / 4-7 epilogue groups, 8-15 transform groups, 16 MMA, 17 weights, 18 raw TMA
constexpr int kEpi = 2, kGrp = 128;
constexpr int kXf0 = 8, kMma = 16, kWld = 17, kTma = 18;
constexpr int PLANE = 134;                   // operand slots (16 B) per k8 plane: 1 + 128 + 1 used; = 6 (mod 8): conflict-free stores
constexpr uint32_t kAStage = 4u * PLANE * 16u;            // 8576 B: one 32-channel operand row
constexpr uint32_t kRawStage = 8192u + 256u;              // bf16 row slab + scale / shift tail
constexpr uint32_t kOutTile = 8192u;                      // 128 px x 32 ch bf16 staging row (64B swizzle)
constexpr int kNR = 6, kNA = 6, kNS = 5;                  // ring depths: raw slabs, operand slabs, accumulator slots (compile time: cheap % and /)
// PAIR variant (64-wide level): the M tile is one image row of TWO samples (64 + 64 pixels).  A shifted start address would pull the
// neighbour sample's edge pixel across the seam, so the transform writes three operand copies per slab instead (kx = 0, 1, 2: shifted
// by +1 / 0 / -1 slots, the seam slots stay zero); blocks are 4 rows (H = 128 rows give 148 CTAs ~7 blocks each)
constexpr uint32_t kPairRawStage = 8192u + 512u;          // two samples' scale / shift behind the slab
constexpr uint32_t kPairAStage = 3u * kAStage;
constexpr size_t kSmemCap = 232448 - 1024;
enum RowKind { ROW_ACT = 0, ROW_STEM = 1, ROW_FINAL = 2 };

struct alignas(64) RowMaps {
    CUtensorMap src[2];    // main sources (concat order), box 32 ch x (128 | 64) px
    CUtensorMap rsrc[2];   // raw block input of the 1x1 res_conv (or [0] = identity residual, 64B swizzle)
    CUtensorMap out;       // output rows, 64B swizzle
};

struct RowArgs {
    int B, H, nblocks;          // 16-row blocks in the whole batch (PAIR: 4-row blocks of sample pairs)
    int C0, rC0;                // channels of source 0 (concat boundary) of the main / res inputs
    int Cin;
    const float* scale; const float* shift;     // [B][Cin] GroupNorm of the input
    const __nv_bfloat16* w;                      // packed: main chunks, then res chunks
    uint32_t w_bytes;
    const float* bias; const float* temb; int temb_stride; const float* res_bias;
    const float* cond; const float* x_t; int L, hop;   // stem
    float* parts; int nparts;   // [B][nparts][32][2]
    float* frames; float final_bias;   // final Block: frames are written only when non-null (sddm_eps / debug fetch)
    int post_on;                       // final Block: overlap-add of the frames + posterior update fused into the epilogue
    PostP post; PostCoef pk;
    int gn_on; GnFuse gn;
    uint32_t off_w, off_raw, off_a, off_out, off_res;
    long long* trace;           // debug: per-role wait / busy cycle counters of CTA 0 (nullptr = off), 32 counters per launch
};

struct RowHdr {
    uint64_t raw_full[kNR], raw_empty[kNR];
    uint64_t full_a[kNA], empty_a[kNA];
    uint64_t acc_full[kNS], acc_empty[kNS];
    uint64_t res_full[kEpi][2];
    uint64_t w_full;
    uint32_t tmem_base;
    uint32_t gn_last[kEpi][2];
    alignas(16) float addv[kEpi][2][32];                  // [1] = second sample of a PAIR tile
    // final Block: positions 64..127 of frame row y wait here for the group that owns row y + 1 (overlap-add partner)
    uint64_t f_full[4], f_empty[4];
    alignas(16) float fup[4][64];
};
constexpr uint32_t kHdr = 2048;
static_assert(sizeof(RowHdr) <= kHdr, "header too large");

// the k-th sample segment of a CTA that owns the 16-row blocks [b0, b1): output rows [ya, yb) of sample n need input rows [r0, r1]
// yo = first row whose results this CTA publishes; ext (final Block with the fused overlap-add): the run also recomputes frame row
// ya - 1, whose upper half the samples of row ya need (it belongs to the CTA above, which publishes it)
// (LG = log2 of the rows per block; PAIR: n counts sample pairs)
struct Seg { int n, ya, yb, r0, r1, yo; };
template <int LG = 4>
__device__ __forceinline__ bool seg_at(int H, int b0, int b1, int k, Seg& s, bool ext) {
    const int bps = H >> LG, n0 = b0 / bps;
    s.n = n0 + k;
    if (s.n * bps >= b1) return false;
    s.ya = k == 0 ? (b0 - n0 * bps) << LG : 0;
    s.yo = s.ya;
    if (ext && s.ya > 0) s.ya -= 1;
    const int e = (b1 - s.n * bps) << LG;
    s.yb = e < H ? e : H;
    s.r0 = s.ya > 0 ? s.ya - 1 : 0;
    s.r1 = s.yb < H ? s.yb : H - 1;
    return true;
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// KIND: ROW_ACT (bf16 activation rows in, 32 channels out), ROW_STEM (waveform windows in), ROW_FINAL (1 channel out -> frames)
// NMAIN / NRES: 32-channel slabs per input row of the 3x3 conv / the 1x1 res_conv; UP: nearest x2 of a [H/2][64] input;
// AFF: GroupNorm-apply + Swish on the main slabs; RESID: identity residual added in the epilogue; PAIR: 64-wide level, two samples per tile
template <int NMAIN, bool PAIR>
struct RowRings {
    // The raw ring depth must be EVEN: the two transform groups take alternate slabs, so a stage of an even ring always belongs to the
    // same group.  On an odd ring the groups alternate on a stage; a group that runs ahead then polls raw_full[s] for use k + 1 while
    // the other group's load of use k is still in flight (TMA loads complete out of order), the parity wait passes on the completed
    // use k - 1, the group transforms the wrong slab and releases the stage early (seen as a Warp Illegal Instruction at the loader's
    // next arrive.expect_tx on the unfinished phase).  The operand ring may be odd: the MMA warp consumes slabs in order, and a group
    // reaches use k of a stage only after the MMA consumed the slab before its previous one, which is past use k - 2 of that stage.
    static constexpr int NR = PAIR ? (NMAIN >= 4 ? 4 : 6) : kNR;
    static constexpr int NA = PAIR ? (NMAIN >= 4 ? 3 : 4) : kNA;
    static_assert(NR % 2 == 0, "raw ring depth must be even (two transform groups)");
    static constexpr uint32_t RAW = PAIR ? kPairRawStage : kRawStage;
    static constexpr uint32_t AST = PAIR ? kPairAStage : kAStage;
    static constexpr int LG = PAIR ? 2 : 4;
};
template <int KIND, int NMAIN, int NRES, bool UP, bool AFF, bool RESID, bool PAIR = false>
__global__ void __launch_bounds__(kRowThreads, 1) conv_row_kernel(const RowArgs a, const __grid_constant__ RowMaps maps) {
    static_assert(!PAIR || (KIND == ROW_ACT && !UP && !RESID), "the pair tile serves plain / res_conv ResnetBlock convolutions only");
    using RR = RowRings<NMAIN, PAIR>;
    constexpr int NR = RR::NR, NA = RR::NA, LG = RR::LG;
    constexpr uint32_t RAWST = RR::RAW, AST = RR::AST;
    constexpr int NSLAB = NMAIN + NRES;
    constexpr int KSTEPS = KIND == ROW_STEM ? 1 : 2;                 // K16 steps per main slab
    constexpr int NCOLS = KIND == ROW_FINAL ? 16 : 96;               // N of the main MMAs
    constexpr int SLOT = KIND == ROW_FINAL ? 32 : 96;                // TMEM columns per accumulator slot
    constexpr uint32_t TMEM_COLS = KIND == ROW_FINAL ? 256u : 512u;
    constexpr uint32_t W_CHUNK = (uint32_t)NCOLS * 32u;              // bytes per (k16, kx) weight chunk: 2 halves x N rows x 16 B
    constexpr bool EXT = KIND == ROW_FINAL;                          // final Block: runs start one frame row early (overlap-add partner)
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t dyn = smem_u32(smem_dyn), base = (dyn + 1023u) & ~1023u;
    RowHdr* hdr = reinterpret_cast<RowHdr*>(smem_dyn + (base - dyn));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pdl_launch_dependents();

    // contiguous run of blocks of this CTA, and its number of input rows (= accumulator items)
    const int q = a.nblocks / (int)gridDim.x, rem = a.nblocks - q * (int)gridDim.x;
    const int b0 = (int)blockIdx.x * q + ((int)blockIdx.x < rem ? (int)blockIdx.x : rem);
    const int b1 = b0 + q + ((int)blockIdx.x < rem ? 1 : 0);
    int nitems = 0;
    {
        Seg s;
        for (int k = 0; seg_at<LG>(a.H, b0, b1, k, s, EXT); ++k) nitems += s.r1 - s.r0 + 1;
    }

    if (tid == 0) {
        for (int i = 0; i < NR; ++i) { mbar_init(smem_u32(&hdr->raw_full[i]), 1); mbar_init(smem_u32(&hdr->raw_empty[i]), kGrp / 32); }
        for (int i = 0; i < NA; ++i) { mbar_init(smem_u32(&hdr->full_a[i]), kGrp / 32); mbar_init(smem_u32(&hdr->empty_a[i]), 1); }
        for (int i = 0; i < kNS; ++i) { mbar_init(smem_u32(&hdr->acc_full[i]), 1); mbar_init(smem_u32(&hdr->acc_empty[i]), 12); }
        for (int e = 0; e < kEpi; ++e) for (int k = 0; k < 2; ++k) mbar_init(smem_u32(&hdr->res_full[e][k]), 1);
        mbar_init(smem_u32(&hdr->w_full), 1);
        for (int i = 0; i < 4; ++i) { mbar_init(smem_u32(&hdr->f_full[i]), 2); mbar_init(smem_u32(&hdr->f_empty[i]), 2); }
        fence_barrier_init();
    }
    if (warp == kMma) tmem_alloc(smem_u32(&hdr->tmem_base), TMEM_COLS);
    // operand ring: zero once - the x halo slots (0 and 129; PAIR: the seam slots of the shifted copies) of every plane are never
    // written again (stem: plane 1 stays zero too)
    for (uint32_t i = (uint32_t)tid; i < (uint32_t)NA * AST / 16u; i += kRowThreads)
        sts128(base + a.off_a + i * 16u, make_uint4(0u, 0u, 0u, 0u));
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = hdr->tmem_base;
    if (g_hang && tid == 0 && blockIdx.x == 0) { g_hang[1] = base; g_hang[2] = (unsigned)nitems; g_hang[3] = (unsigned)(NSLAB * 100 + NA * 10 + NR); }
    const bool tr = SDDM_ROW_TRACE && a.trace != nullptr && blockIdx.x == 0;   // compiled out unless built with -DSDDM_ROW_TRACE=1
    long long tw[6] = {0, 0, 0, 0, 0, 0};
    const long long t_begin = tr ? clock64() : 0;

    if (warp < 8) {
        // ============================== epilogue: group e owns the output rows with y & 1 == e ===========================
        pdl_wait();
        const int e = warp >> 2, w4 = warp & 3, m = tid & 127;
        const int bar_id = 1 + e;
        // single-thread duties sit on the warps of scheduler partitions 3 and 2 (warp id % 4): partition 0 already hosts the MMA
        // issuer and one warp of every other group, and the group advances at the pace of its slowest warp
        const bool leader = m == 96;       // TMA stores (+ their bulk-group waits), GroupNorm arrival
        const bool res_leader = m == 64;   // identity-residual prefetch (mbarrier based: any thread may wait on it)
        const uint32_t obuf0 = base + a.off_out + (uint32_t)(e * 2) * kOutTile;
        const uint32_t rbuf0 = base + a.off_res + (uint32_t)(e * 2) * kOutTile;
        const uint32_t addv_u32 = smem_u32(hdr->addv[e][PAIR ? (m >> 6) : 0]);
        const uint32_t lane_tm = tmem_base + ((uint32_t)(w4 * 32) << 16);
        const uint32_t swz = (uint32_t)((m >> 1) & 3);       // 64B swizzle of this thread's staging / residual row
        const int c2 = lane & 15, hrow = lane >> 4;
        float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;   // statistics of this group's rows of the current block (2 channels per lane)
        int ob = 0;                                          // rows stored so far (staging / residual buffer parity)
        uint32_t fwc[2] = {0u, 0u}, frc[2] = {0u, 0u};       // final Block: fills / reads so far of the two fup slots this group writes / reads
        // identity residual: TMA prefetch two rows ahead along this group's (n, y) sequence
        int pk = 0, py = -1, pn = 0, pend = 0, pissued = 0;
        Seg ps{};
        auto res_issue = [&]() {
            for (;;) {   // advance (pk, py) to this group's next output row
                if (py < 0) {
                    if (!seg_at<LG>(a.H, b0, b1, pk, ps, EXT)) return;
                    py = ps.ya + ((ps.ya & 1) == e ? 0 : 1);
                    pend = ps.yb; pn = ps.n;
                } else {
                    py += 2;
                }
                if (py < pend) break;
                py = -1; ++pk;
            }
            const uint32_t bar = smem_u32(&hdr->res_full[e][pissued & 1]);
            mbar_expect_tx(bar, kOutTile);
            tma_load_4d(rbuf0 + (uint32_t)(pissued & 1) * kOutTile, &maps.rsrc[0], 0, 0, py, pn, bar);
            ++pissued;
        };
        if (RESID && res_leader) { res_issue(); res_issue(); }
        int item_base = 0;
        Seg s;
        for (int k = 0; seg_at<LG>(a.H, b0, b1, k, s, EXT); item_base += s.r1 - s.r0 + 1, ++k) {
            if (KIND != ROW_FINAL) {
                // per-channel additive term of this sample: bias (+ noise-level embedding row) (+ res_conv bias)
                group_bar(bar_id);
                if (m < (PAIR ? 64 : 32)) {
                    const int c = m & 31;
                    int nn = s.n;
                    if (PAIR) { nn = 2 * s.n + (m >> 5); if (nn >= a.B) nn = a.B - 1; }
                    float v = __ldg(a.bias + c);
                    if (a.temb) v += __ldg(a.temb + (int64_t)nn * a.temb_stride + c);
                    if (NRES) v += __ldg(a.res_bias + c);
                    hdr->addv[e][m >> 5][c] = v;
                }
                group_bar(bar_id);
            }
            for (int y = s.ya + ((s.ya & 1) == e ? 0 : 1); y < s.yb; y += 2) {
                const bool vA = y - 1 >= s.r0, vC = y + 1 <= s.r1;
                const int iB = item_base + (y - s.r0), last = vC ? iB + 1 : iB;
                const int slB = iB % kNS, slA = (iB + kNS - 1) % kNS, slC = (iB + 1) % kNS;
                // final Block: the posterior update of this row's samples reads x_t (+ injected noise, + the condition) from HBM -
                // issue those loads now, so their latency hides behind the accumulator wait
                int p_sidx = -1;
                float p_x = 0.f, p_z = 0.f, p_c = 0.f;
                if (KIND == ROW_FINAL && a.post_on) {
                    const PostP& q = a.post;
                    if (w4 < 2) { if (y >= s.yo) p_sidx = y * (RW / 2) + m; }
                    else if (y + 1 == a.H) p_sidx = a.H * (RW / 2) + (m - 64);
                    if (p_sidx >= 0 && q.do_update) {
                        const int64_t gi = (int64_t)s.n * q.L + p_sidx;
                        p_x = q.x_in[gi];
                        if (q.z && q.t > 1) p_z = q.z[gi];
                        if (q.variant == SDDM_VAR_SUPPORTIVE || q.variant == SDDM_VAR_CONDITIONAL) p_c = q.cond[gi];
                    }
                }
                mbar_wait_t(smem_u32(&hdr->acc_full[last % kNS]), (uint32_t)(last / kNS) & 1u, tr, tw[0]);
                tc_fence_after();
                // arrivals owed to the three accumulator slots: an input row r is read by the output rows r - 1, r, r + 1; those
                // outside [ya, yb) never come, so the first / last row of the segment arrives in their place (3 x 4 warps per slot)
                const bool first = y == s.ya, lastrow = y == s.yb - 1;
                const uint32_t nA = first ? 3u : 1u, nB = 1u + (first ? 1u : 0u) + (lastrow ? 1u : 0u), nC = lastrow ? 3u : 1u;
                const long long te0 = tr ? clock64() : 0;
                if (KIND == ROW_FINAL) {
                    uint32_t ra[4] = {0u, 0u, 0u, 0u}, rb[4], rc[4] = {0u, 0u, 0u, 0u};
                    if (vA) tmem_ld4_nowait(lane_tm + (uint32_t)(slA * SLOT), ra);
                    tmem_ld4_nowait(lane_tm + (uint32_t)(slB * SLOT), rb);
                    if (vC) tmem_ld4_nowait(lane_tm + (uint32_t)(slC * SLOT), rc);
                    tmem_wait_ld();
                    tc_fence_before();
                    if (lane == 0) {
                        if (vA) mbar_arrive_n(smem_u32(&hdr->acc_empty[slA]), nA);
                        mbar_arrive_n(smem_u32(&hdr->acc_empty[slB]), nB);
                        if (vC) mbar_arrive_n(smem_u32(&hdr->acc_empty[slC]), nC);
                    }
                    if (tr) tw[1] += clock64() - te0;
                    const float f = (__uint_as_float(ra[0]) + __uint_as_float(rb[1])) + __uint_as_float(rc[2]) + a.final_bias;
                    if (a.frames && y >= s.yo) a.frames[((int64_t)s.n * a.H + y) * RW + m] = f;
                    if (!a.post_on) continue;
                    // overlap-add (UNetModified2.py:30-41, hop = W / 2): eps[hop a + j] = frame[a - 1][j + hop] + frame[a][j], then the
                    // posterior update of that sample (post_one, bit-exact) - frames / eps_hat never touch HBM in the sampling loop
                    auto emit = [&](float ev) {
                        const PostP& q = a.post;
                        const int64_t gi = (int64_t)s.n * q.L + p_sidx;
                        if (q.eps_out) q.eps_out[gi] = ev;
                        if (!q.do_update) return;
                        const bool add_noise = q.t > 1;
                        float zv = p_z;
                        if (add_noise && !q.z)
                        {
                            const uint64_t sd = q.seed_dev ? q.seed_dev[0] : q.seed;
                            const int64_t r0 = q.seed_dev ? (int64_t)q.seed_dev[1] : q.row0;
                            zv = philox_normal1(sd, (uint32_t)(p_sidx >> 2), (uint64_t)(r0 + s.n), (uint32_t)(q.T + 1 - q.t), p_sidx & 3);
                        }
                        const float o = post_one(q.variant, a.pk, p_x, ev, p_c, zv, add_noise);
                        q.x_out[gi] = o;
                        if (q.x_trace) q.x_trace[gi] = o;
                    };
                    if (w4 >= 2) {   // positions 64..127 belong to sample row y + 1
                        if (y + 1 < s.yb) {
                            const int slot = y & 3, ci = (y >> 1) & 1;
                            mbar_wait(smem_u32(&hdr->f_empty[slot]), (fwc[ci] & 1u) ^ 1u);
                            hdr->fup[slot][m - 64] = f;
                            __syncwarp();
                            if (lane == 0) mbar_arrive(smem_u32(&hdr->f_full[slot]));
                            ++fwc[ci];
                        } else if (y + 1 == a.H) {
                            emit(f);                    // the last half frame of the sample has no partner
                        }
                    } else if (y >= s.yo) {   // positions 0..63: sample row y
                        float ev = f;
                        if (y > 0) {
                            const int slot = (y - 1) & 3, ci = ((y - 1) >> 1) & 1;
                            mbar_wait(smem_u32(&hdr->f_full[slot]), frc[ci] & 1u);
                            ev = hdr->fup[slot][m] + f;      // ascending frame order, as the reference accumulates
                            __syncwarp();
                            if (lane == 0) mbar_arrive(smem_u32(&hdr->f_empty[slot]));
                            ++frc[ci];
                        }
                        emit(ev);
                    }
                    continue;
                }
                const uint32_t obuf = obuf0 + (uint32_t)(ob & 1) * kOutTile;
                const uint32_t orow = obuf + (uint32_t)m * 64u;
                const uint32_t rrow = rbuf0 + (uint32_t)(ob & 1) * kOutTile + (uint32_t)m * 64u;
                if (RESID) mbar_wait(smem_u32(&hdr->res_full[e][ob & 1]), (uint32_t)(ob >> 1) & 1u);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t ra[16], rb[16], rc[16];
                    tmem_ld16_nowait(lane_tm + (uint32_t)(slB * SLOT + 32 + half * 16), rb);
                    if (vA) tmem_ld16_nowait(lane_tm + (uint32_t)(slA * SLOT + half * 16), ra);
                    if (vC) tmem_ld16_nowait(lane_tm + (uint32_t)(slC * SLOT + 64 + half * 16), rc);
                    tmem_wait_ld();
                    if (half == 1) {   // all three slots fully read
                        tc_fence_before();
                        if (lane == 0) {
                            if (vA) mbar_arrive_n(smem_u32(&hdr->acc_empty[slA]), nA);
                            mbar_arrive_n(smem_u32(&hdr->acc_empty[slB]), nB);
                            if (vC) mbar_arrive_n(smem_u32(&hdr->acc_empty[slC]), nC);
                        }
                    }
                    float v[16];
                    {   // tap sum + additive term on fp32 pairs (FADD2): same per-element order B + A + C + addv as scalar code
                        float2 p[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) p[i] = make_float2(__uint_as_float(rb[2 * i]), __uint_as_float(rb[2 * i + 1]));
                        if (vA) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) p[i] = fadd2(p[i], make_float2(__uint_as_float(ra[2 * i]), __uint_as_float(ra[2 * i + 1])));
                        }
                        if (vC) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) p[i] = fadd2(p[i], make_float2(__uint_as_float(rc[2 * i]), __uint_as_float(rc[2 * i + 1])));
                        }
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd) {
                            const uint4 u = lds128(addv_u32 + (uint32_t)(half * 4 + qd) * 16u);
                            p[2 * qd] = fadd2(p[2 * qd], make_float2(__uint_as_float(u.x), __uint_as_float(u.y)));
                            p[2 * qd + 1] = fadd2(p[2 * qd + 1], make_float2(__uint_as_float(u.z), __uint_as_float(u.w)));
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) { v[2 * i] = p[i].x; v[2 * i + 1] = p[i].y; }
                    }
                    if (RESID) {
#pragma unroll
                        for (int qd = 0; qd < 2; ++qd) {
                            const uint4 u = lds128(rrow + ((((uint32_t)(half * 2 + qd)) ^ swz) << 4));
                            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) { v[8 * qd + 2 * kk] += bf16_lo(w[kk]); v[8 * qd + 2 * kk + 1] += bf16_hi(w[kk]); }
                        }
                    }
#pragma unroll
                    for (int qd = 0; qd < 2; ++qd) {
                        uint4 u;
                        u.x = pack_bf16(v[8 * qd + 0], v[8 * qd + 1]); u.y = pack_bf16(v[8 * qd + 2], v[8 * qd + 3]);
                        u.z = pack_bf16(v[8 * qd + 4], v[8 * qd + 5]); u.w = pack_bf16(v[8 * qd + 6], v[8 * qd + 7]);
                        sts128(orow + ((((uint32_t)(half * 2 + qd)) ^ swz) << 4), u);
                    }
                }
                if (tr) tw[1] += clock64() - te0;
                fence_async_smem();
                // every TMA store issued so far has finished reading its staging row (the newest was issued a whole row ago), so
                // after the barrier the OTHER buffer is free for the next row: one barrier per row
                const long long tw0 = tr ? clock64() : 0;
                if (leader) bulk_wait_read_0();
                const long long tg0 = tr ? clock64() : 0;
                group_bar(bar_id);
                if (tr) { tw[3] += tg0 - tw0; tw[2] += clock64() - tg0; }
                const long long ts0 = tr ? clock64() : 0;
                if (leader) {
                    tma_store_4d(&maps.out, obuf, 0, 0, y, PAIR ? 2 * s.n : s.n);   // PAIR: box of two samples, the odd tail is clipped
                    bulk_commit();
                }
                if (RESID && res_leader) res_issue();   // the residual buffer of this row is free: refill it for this group's row after next
                if (tr) tw[4] += clock64() - ts0;
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
                if (tr) tw[5] += clock64() - ts0;
                if ((y & ((1 << LG) - 1)) >= (1 << LG) - 2) {   // this group's last row of the block: publish its partial
                    s1a += __shfl_xor_sync(0xffffffffu, s1a, 16); s1b += __shfl_xor_sync(0xffffffffu, s1b, 16);
                    s2a += __shfl_xor_sync(0xffffffffu, s2a, 16); s2b += __shfl_xor_sync(0xffffffffu, s2b, 16);
                    // PAIR: warps 0-1 hold the first sample of the pair, warps 2-3 the second (absent behind an odd batch)
                    const int pn_ = PAIR ? 2 * s.n + (w4 >> 1) : s.n;
                    const int slot = PAIR ? (y >> LG) * 4 + e * 2 + (w4 & 1) : (y >> LG) * 8 + e * 4 + w4;
                    if (hrow == 0 && pn_ < a.B) {
                        float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float2*>(a.parts) + ((int64_t)pn_ * a.nparts + slot) * 32 + 2 * c2);
                        *dst = make_float4(s1a, s2a, s1b, s2b);
                    }
                    s1a = s1b = s2a = s2b = 0.f;
                }
            }
            // one arrival per (segment, group) - an atomic round trip per block would cost ~1.5 us each: this group's rows of sample n
            // are done and their partials published; the group that completes the sample's H rows finalises the consumer's GroupNorm
            if (KIND != ROW_FINAL && a.gn_on) {
                __threadfence();
                group_bar(bar_id);
                if (leader) {
                    const unsigned add = (unsigned)((s.yb - s.ya) >> 1);
#pragma unroll
                    for (int hh = 0; hh < (PAIR ? 2 : 1); ++hh) {
                        const int nn = PAIR ? 2 * s.n + hh : s.n;
                        hdr->gn_last[e][hh] = nn < a.B && atomicAdd(a.gn.counter + nn, add) + add == (unsigned)a.gn.expect ? 1u : 0u;
                    }
                }
                group_bar(bar_id);
#pragma unroll
                for (int hh = 0; hh < (PAIR ? 2 : 1); ++hh)
                    if (hdr->gn_last[e][hh]) {
                        const int nn = PAIR ? 2 * s.n + hh : s.n;
                        __threadfence();
                        gn_fused_finalize(a.gn, nn, m, kGrp);
                        if (leader) a.gn.counter[nn] = 0u;
                    }
            }
        }
        if (leader) bulk_wait_all();
        if (tr && leader) {
            long long* o = a.trace + e * 8;
            o[0] = clock64() - t_begin; o[1] = tw[0]; o[2] = tw[1]; o[3] = tw[2]; o[4] = tw[3]; o[5] = tw[4]; o[6] = tw[5]; o[7] = ob;
        }
    } else if (warp == kMma) {
        // ============================== MMA issuer ==========================================================================
        const uint32_t idesc = make_idesc(NCOLS), idesc_res = make_idesc(32);
        const uint64_t a_desc0 = make_desc_nosw(base + a.off_a, (uint32_t)PLANE * 16u, 128u);
        const uint64_t w_desc0 = make_desc_nosw(base + a.off_w, (uint32_t)NCOLS * 16u, 128u);
        const uint64_t wr_desc0 = make_desc_nosw(base + a.off_w + (uint32_t)(NMAIN * KSTEPS * 3) * W_CHUNK, 32u * 16u, 128u);
        mbar_wait_t(smem_u32(&hdr->w_full), 0u, tr, tw[0]);
        tc_fence_after();
        int sa = 0;
        uint32_t pa = 0;
        for (int it = 0; it < nitems; ++it) {
            const int slot = it % kNS;
            mbar_wait_t(smem_u32(&hdr->acc_empty[slot]), ((uint32_t)(it / kNS) & 1u) ^ 1u, tr, tw[1]);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(slot * SLOT);
#pragma unroll
            for (int sl = 0; sl < NSLAB; ++sl) {
                mbar_wait_t(smem_u32(&hdr->full_a[sa]), pa, tr, tw[2]);
                tc_fence_after();
                const uint64_t adesc = a_desc0 + (uint64_t)((uint32_t)sa * (AST >> 4));
                if (elect_one()) {
                    if (sl < NMAIN) {
#pragma unroll
                        for (int h = 0; h < KSTEPS; ++h)
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx)
                                umma(d_tmem, adesc + (uint64_t)(PAIR ? (uint32_t)kx * (kAStage >> 4) + (uint32_t)(h * 2 * PLANE) : (uint32_t)(h * 2 * PLANE + kx)),
                                     w_desc0 + (uint64_t)((uint32_t)((sl * KSTEPS + h) * 3 + kx) * (W_CHUNK >> 4)), idesc, (sl | h | kx) ? 1u : 0u);
                    } else {   // 1x1 res_conv over the raw block input: centre tap, centre (ky = 1) columns
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                            umma(d_tmem + 32u, adesc + (uint64_t)(PAIR ? (kAStage >> 4) + (uint32_t)(h * 2 * PLANE) : (uint32_t)(h * 2 * PLANE + 1)),
                                 wr_desc0 + (uint64_t)((uint32_t)((sl - NMAIN) * 2 + h) * (1024u >> 4)), idesc_res, 1u);
                    }
                    umma_commit(smem_u32(&hdr->empty_a[sa]));
                    if (sl == NSLAB - 1) umma_commit(smem_u32(&hdr->acc_full[slot]));
                }
                if (++sa == NA) { sa = 0; pa ^= 1u; }
            }
        }
        if (tr && lane == 0) { long long* o = a.trace + 16; o[0] = clock64() - t_begin; o[1] = tw[0]; o[2] = tw[1]; o[3] = tw[2]; o[4] = nitems; }
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
            for (int k = 0; seg_at<LG>(a.H, b0, b1, k, s, EXT); ++k)
                for (int r = s.r0; r <= s.r1; ++r) {
#pragma unroll
                    for (int sl = 0; sl < NSLAB; ++sl) {
                        const uint32_t bar = smem_u32(&hdr->raw_full[rs]);
                        mbar_wait_t(smem_u32(&hdr->raw_empty[rs]), pr ^ 1u, tr, tw[0]);
                        const uint32_t dst = base + a.off_raw + (uint32_t)rs * RAWST;
                        if (KIND == ROW_STEM) {
                            mbar_expect_tx(bar, 1024u);
                            bulk_g2s(dst, a.cond + (int64_t)s.n * a.L + (int64_t)r * a.hop, 512u, bar);
                            bulk_g2s(dst + 512u, a.x_t + (int64_t)s.n * a.L + (int64_t)r * a.hop, 512u, bar);
                        } else if (sl < NMAIN) {
                            const int cb = sl * 32, si = cb < a.C0 ? 0 : 1;
                            mbar_expect_tx(bar, (UP ? 4096u : 8192u) + (AFF ? (PAIR ? 512u : 256u) : 0u));
                            tma_load_4d(dst, &maps.src[si], cb - (si ? a.C0 : 0), 0, UP ? (r >> 1) : r, PAIR ? 2 * s.n : s.n, bar);   // PAIR: box of two samples, zero fill behind an odd batch
                            if (AFF) {
                                const int n0 = PAIR ? 2 * s.n : s.n;
                                bulk_g2s(dst + 8192u, a.scale + (int64_t)n0 * a.Cin + cb, 128u, bar);
                                bulk_g2s(dst + 8192u + 128u, a.shift + (int64_t)n0 * a.Cin + cb, 128u, bar);
                                if (PAIR) {
                                    const int n1 = n0 + 1 < a.B ? n0 + 1 : n0;
                                    bulk_g2s(dst + 8192u + 256u, a.scale + (int64_t)n1 * a.Cin + cb, 128u, bar);
                                    bulk_g2s(dst + 8192u + 384u, a.shift + (int64_t)n1 * a.Cin + cb, 128u, bar);
                                }
                            }
                        } else {
                            const int cb = (sl - NMAIN) * 32, si = cb < a.rC0 ? 0 : 1;
                            mbar_expect_tx(bar, 8192u);
                            tma_load_4d(dst, &maps.rsrc[si], cb - (si ? a.rC0 : 0), 0, r, PAIR ? 2 * s.n : s.n, bar);
                        }
                        if (++rs == NR) { rs = 0; pr ^= 1u; }
                    }
                }
            if (tr) { long long* o = a.trace + 21; o[0] = clock64() - t_begin; o[1] = tw[0]; }
        }
    } else if (warp >= kXf0 && warp < kXf0 + 8) {
        // ============================== transform: raw row slab -> bf16 operand row ========================================
        // two groups of 128 threads take alternate slabs of the CTA's (row, slab) sequence; no coordinates needed here
        const int gi = (warp - kXf0) >> 2, gt = tid - (kXf0 * 32 + gi * kGrp);
        const int j = gt & 3, px0 = gt >> 2;
        const int nslabs = nitems * NSLAB;
        for (int qn = gi; qn < nslabs; qn += 2) {
            const int rs = qn % NR, sa = qn % NA, sl = NSLAB == 1 ? 0 : qn % NSLAB;
            const uint32_t raw = base + a.off_raw + (uint32_t)rs * RAWST;
            const uint32_t opd = base + a.off_a + (uint32_t)sa * AST;
            mbar_wait_t(smem_u32(&hdr->raw_full[rs]), (uint32_t)(qn / NR) & 1u, tr, tw[0]);
            mbar_wait_t(smem_u32(&hdr->empty_a[sa]), ((uint32_t)(qn / NA) & 1u) ^ 1u, tr, tw[1]);
            const long long tx0 = tr ? clock64() : 0;
            if (KIND == ROW_STEM) {
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
                const bool is_main = sl < NMAIN;
                const bool aff = AFF && is_main, up = UP && is_main;
                float2 sc2[4], sh2[4];   // (scale, shift) / 2 of this thread's 8 channels as fp32 pairs: swish(y) = h + h tanh(h), h = y / 2
                auto load_aff = [&](uint32_t ss) {
                    const uint4 s0 = lds128(ss), s1 = lds128(ss + 16u), h0 = lds128(ss + 128u), h1 = lds128(ss + 144u);
                    sc2[0] = make_float2(0.5f * __uint_as_float(s0.x), 0.5f * __uint_as_float(s0.y)); sc2[1] = make_float2(0.5f * __uint_as_float(s0.z), 0.5f * __uint_as_float(s0.w));
                    sc2[2] = make_float2(0.5f * __uint_as_float(s1.x), 0.5f * __uint_as_float(s1.y)); sc2[3] = make_float2(0.5f * __uint_as_float(s1.z), 0.5f * __uint_as_float(s1.w));
                    sh2[0] = make_float2(0.5f * __uint_as_float(h0.x), 0.5f * __uint_as_float(h0.y)); sh2[1] = make_float2(0.5f * __uint_as_float(h0.z), 0.5f * __uint_as_float(h0.w));
                    sh2[2] = make_float2(0.5f * __uint_as_float(h1.x), 0.5f * __uint_as_float(h1.y)); sh2[3] = make_float2(0.5f * __uint_as_float(h1.z), 0.5f * __uint_as_float(h1.w));
                };
                if (aff) load_aff(raw + 8192u + (uint32_t)j * 32u);
                const uint32_t dstp = opd + (uint32_t)j * (uint32_t)PLANE * 16u;
                uint4 rv[4];
#pragma unroll
                for (int rd = 0; rd < 4; ++rd)
                    if (!up || rd < 2) rv[rd] = lds128(raw + (uint32_t)(px0 + 32 * rd) * 64u + (uint32_t)j * 16u);
#pragma unroll
                for (int rd = 0; rd < 4; ++rd) {
                    if (up && rd >= 2) break;
                    if (PAIR && aff && rd == 2) load_aff(raw + 8192u + 256u + (uint32_t)j * 32u);   // pixels 64..127: the pair's second sample
                    uint4 o = rv[rd];
                    if (aff) {
                        const uint32_t w[4] = {o.x, o.y, o.z, o.w};
                        uint32_t ow[4];
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const float2 h = ffma2(bf16x2_to_f32x2(w[kk]), sc2[kk], sh2[kk]);
                            const float2 sw = ffma2(h, make_float2(tanh_approx(h.x), tanh_approx(h.y)), h);
                            ow[kk] = pack_bf16(sw.x, sw.y);
                        }
                        o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                    }
                    const int px = px0 + 32 * rd;
                    if (PAIR) {   // copy kx holds pixel px + kx - 1 of the same sample at tile row px (seam slots stay zero)
                        const uint32_t d1 = dstp + kAStage + (uint32_t)px * 16u;
                        sts128(d1, o);
                        if (is_main) {
                            if ((px & 63) != 63) sts128(d1 - kAStage + 16u, o);
                            if ((px & 63) != 0) sts128(d1 + kAStage - 16u, o);
                        }
                    } else if (up) {
                        sts128(dstp + (uint32_t)(2 * px + 1) * 16u, o);
                        sts128(dstp + (uint32_t)(2 * px + 2) * 16u, o);
                    } else {
                        sts128(dstp + (uint32_t)(px + 1) * 16u, o);
                    }
                }
            }
            // one arrival per warp (32 same-address mbarrier arrivals serialise in the shared-memory pipe): every lane fences its
            // own operand stores towards the async proxy, the warp converges, lane 0 publishes
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(smem_u32(&hdr->full_a[sa]));
                mbar_arrive(smem_u32(&hdr->raw_empty[rs]));
            }
            if (tr) tw[2] += clock64() - tx0;
        }
        if (tr && gt == 0) { long long* o = a.trace + 24 + 4 * gi; o[0] = clock64() - t_begin; o[1] = tw[0]; o[2] = tw[1]; o[3] = tw[2]; }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMma) {
        __syncwarp();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// NHWC bf16 tensor [B][H][W][C] as a 4-D tensor map (C innermost) with box (32, bw, 1, bn); swz 0 / 64
int encode_rows(CUtensorMap* m, const void* basep, int B, int H, int W, int C, int bw, int swz, int bn = 1) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("conv row: cuTensorMapEncodeTiled is unavailable"); return SDDM_E_CUDA; }
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    const cuuint32_t box[4] = {32, (cuuint32_t)bw, 1, (cuuint32_t)bn};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(basep), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv row: cuTensorMapEncodeTiled failed (%d) for [%d,%d,%d,%d]", (int)r, B, H, W, C); return SDDM_E_CUDA; }
    return SDDM_OK;
}

long long* g_row_trace = nullptr;   // device buffer [64 launches][32 counters], set by sddm_debug_row_trace
int g_row_trace_launch = 0;

int device_sms() { return device_sm_count(); }

}  // namespace

int conv_row_nparts(int H, int W) { return W == RW ? (H / 16) * 8 : (H / 4) * 4; }   // (block, group, warp [, sample of the pair])
int conv_row_arrivals(int H) { return H; }   // the counter adds up completed rows

// 64-wide level: ResnetBlock convolutions with Cout = 32 on the pair tile (GroupNorm input, optional 1x1 res_conv)
static bool pair_supported(const ConvP& p) {
    static const bool off = [] { const char* e = getenv("SDDM_NO_ROW_PAIR"); return e && e[0] == '1'; }();   // A/B switch: 64-wide level stays on conv_tc.cu
    if (off) return false;
    if (!p.act16 || p.Wout != RW / 2 || p.Hout % 4 || p.Hout < 4 || p.Cout != 32 || p.mode != CONV_S1 || p.res_identity) return false;
    if (!p.src[0].scale) return false;
    for (int i = 0; i < p.nsrc; ++i)
        if (p.src[i].C % 32) return false;
    for (int i = 0; i < p.res_nsrc; ++i)
        if (p.res_Cin && p.res_src[i].C % 32) return false;
    const int cin = p.Cin, rc = p.res_Cin;
    return (cin == 128 && rc == 0) || (cin == 64 && rc == 0) || (cin == 32 && rc == 0) || (cin == 32 && rc == 128) || (cin == 32 && rc == 64);
}

bool conv_row_supported(const ConvP& p) {
    if (pair_supported(p)) return true;
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

// shared-memory plan + launch of one instantiation
template <int KIND, int NMAIN, int NRES, bool UP, bool AFF, bool RESID, bool PAIR = false>
static int launch_row_t(RowArgs a, const RowMaps& maps, cudaStream_t st) {
    using RR = RowRings<NMAIN, PAIR>;
    a.nblocks = PAIR ? ((a.B + 1) / 2) * (a.H >> RR::LG) : a.B * (a.H >> RR::LG);
    const size_t w_al = ((size_t)a.w_bytes + 1023) & ~(size_t)1023;
    const int nout = KIND == ROW_FINAL ? 0 : 2, nres = RESID ? 2 : 0;
    a.off_w = kHdr;
    a.off_out = a.off_w + (uint32_t)w_al;                                     // 1024-aligned (swizzled TMA tiles)
    a.off_res = a.off_out + (uint32_t)(kEpi * nout) * kOutTile;
    a.off_raw = a.off_res + (uint32_t)(kEpi * nres) * kOutTile;
    a.off_a = a.off_raw + (uint32_t)RR::NR * RR::RAW;
    const size_t smem = a.off_a + (size_t)RR::NA * RR::AST + 1024;
    if (smem > kSmemCap + 1024) { set_error("conv row: shared-memory plan does not fit (%zu bytes)", smem); return SDDM_E_INVALID; }
    auto kern = conv_row_kernel<KIND, NMAIN, NRES, UP, AFF, RESID, PAIR>;
    SDDM_SET_MAX_SMEM(kern, kSmemCap + 1024);
    a.trace = g_row_trace ? g_row_trace + (size_t)(g_row_trace_launch++ % 64) * 32 : nullptr;
    const int sms = device_sms();
    const int grid = a.nblocks < sms ? a.nblocks : sms;
    SDDM_CUDA_TRY(launch_pdl(kern, dim3(grid), dim3(kRowThreads), smem, st, a, maps));
    SDDM_LAUNCH_CHECK();
    return SDDM_OK;
}

// ResnetBlock / Upsample convolutions (Cout = 32) and the final Block (Cout = 1, frames out) on a 128-wide level
int launch_conv_row(const ConvP& p, const __nv_bfloat16* w_row, uint32_t w_bytes, float* frames, float final_bias, cudaStream_t st,
                    const PostP* post, const float* k8) {
    if (!conv_row_supported(p) || !w_row) { set_error("conv row: unsupported shape Cin=%d Cout=%d mode=%d W=%d", p.Cin, p.Cout, p.mode, p.Wout); return SDDM_E_INVALID; }
    RowArgs a{};
    RowMaps maps;
    memset(&maps, 0, sizeof(maps));
    a.B = p.B; a.H = p.Hout;
    const bool final_out = p.Cout == 1, up = p.mode == CONV_UP, aff = p.src[0].scale != nullptr;
    const bool res_conv = p.res_Cin && !p.res_identity;
    a.C0 = p.src[0].C; a.rC0 = p.res_src[0].C;
    a.Cin = p.Cin;
    a.scale = p.src[0].scale; a.shift = p.src[0].shift;
    a.w = w_row; a.w_bytes = w_bytes;
    a.bias = p.bias; a.temb = p.temb; a.temb_stride = p.temb_stride; a.res_bias = p.res_bias;
    a.parts = p.parts; a.nparts = p.nparts;
    a.frames = frames; a.final_bias = final_bias;
    if (final_out && post) {
        if (!k8 || post->L != (p.Hout + 1) * (RW / 2)) { set_error("conv row: fused overlap-add needs hop = W / 2 and L = (H + 1) hop"); return SDDM_E_INVALID; }
        a.post_on = 1;
        a.post = *post;
        a.pk = PostCoef{k8[0], k8[1], k8[2], k8[3], k8[4], k8[5], k8[6], k8[7]};
    }
    a.gn_on = p.gn_on; a.gn = p.gn;
    if (!final_out && (!p.parts || p.nparts != conv_row_nparts(p.Hout, p.Wout))) { set_error("conv row: nparts mismatch"); return SDDM_E_INVALID; }
    int rc;
    if (p.Wout == RW / 2) {   // 64-wide level: tiles of two samples (ups.11 / ups.12 of config_unet.json)
        for (int i = 0; i < p.nsrc; ++i)
            if ((rc = encode_rows(&maps.src[i], p.src[i].x, p.B, p.Hin, p.Win, p.src[i].C, 64, 0, 2))) return rc;
        if (res_conv)
            for (int i = 0; i < p.res_nsrc; ++i)
                if ((rc = encode_rows(&maps.rsrc[i], p.res_src[i].x, p.B, p.Hout, p.Wout, p.res_src[i].C, 64, 0, 2))) return rc;
        if ((rc = encode_rows(&maps.out, p.out, p.B, p.Hout, p.Wout, 32, 64, 64, 2))) return rc;
        const int rcin = res_conv ? p.res_Cin : 0;
        if (p.Cin == 128 && !rcin) return launch_row_t<ROW_ACT, 4, 0, false, true, false, true>(a, maps, st);
        if (p.Cin == 64 && !rcin) return launch_row_t<ROW_ACT, 2, 0, false, true, false, true>(a, maps, st);
        if (p.Cin == 32 && !rcin) return launch_row_t<ROW_ACT, 1, 0, false, true, false, true>(a, maps, st);
        if (p.Cin == 32 && rcin == 128) return launch_row_t<ROW_ACT, 1, 4, false, true, false, true>(a, maps, st);
        if (p.Cin == 32 && rcin == 64) return launch_row_t<ROW_ACT, 1, 2, false, true, false, true>(a, maps, st);
        set_error("conv row: no pair instantiation for Cin=%d res_conv=%d", p.Cin, rcin);
        return SDDM_E_INVALID;
    }
    for (int i = 0; i < p.nsrc; ++i)
        if ((rc = encode_rows(&maps.src[i], p.src[i].x, p.B, p.Hin, p.Win, p.src[i].C, up ? 64 : 128, 0))) return rc;
    if (res_conv)
        for (int i = 0; i < p.res_nsrc; ++i)
            if ((rc = encode_rows(&maps.rsrc[i], p.res_src[i].x, p.B, p.Hout, RW, p.res_src[i].C, 128, 0))) return rc;
    if (p.res_identity && (rc = encode_rows(&maps.rsrc[0], p.res_src[0].x, p.B, p.Hout, RW, 32, 128, 64))) return rc;
    if (!final_out && (rc = encode_rows(&maps.out, p.out, p.B, p.Hout, RW, 32, 128, 64))) return rc;
    // the seven layer shapes of a 128-wide level (config_unet.json: final Block, downs.1 block1 / block2, ups.13, ups.14 block1 / block2)
    if (final_out && aff) return launch_row_t<ROW_FINAL, 1, 0, false, true, false>(a, maps, st);
    if (up && p.Cin == 32 && !aff && !res_conv && !p.res_identity) return launch_row_t<ROW_ACT, 1, 0, true, false, false>(a, maps, st);
    if (up && p.Cin == 32 && aff && !res_conv && !p.res_identity) return launch_row_t<ROW_ACT, 1, 0, true, true, false>(a, maps, st);
    if (!up && p.Cin == 32 && !aff && !res_conv && !p.res_identity) return launch_row_t<ROW_ACT, 1, 0, false, false, false>(a, maps, st);
    if (!up && p.Cin == 32 && aff && !res_conv && !p.res_identity) return launch_row_t<ROW_ACT, 1, 0, false, true, false>(a, maps, st);
    if (!up && p.Cin == 32 && aff && p.res_identity) return launch_row_t<ROW_ACT, 1, 0, false, true, true>(a, maps, st);
    if (!up && p.Cin == 32 && aff && res_conv) return launch_row_t<ROW_ACT, 1, 2, false, true, false>(a, maps, st);
    if (!up && p.Cin == 64 && aff && !res_conv && !p.res_identity) return launch_row_t<ROW_ACT, 2, 0, false, true, false>(a, maps, st);
    set_error("conv row: no instantiation for Cin=%d up=%d affine=%d res_conv=%d identity=%d final=%d", p.Cin, (int)up, (int)aff, (int)res_conv, p.res_identity, (int)final_out);
    return SDDM_E_INVALID;
}

// stem: SignalToFrames x 2 + cat + conv3x3(2 -> 32)
int launch_stem_row(const StemP& p, const __nv_bfloat16* w_row, uint32_t w_bytes, cudaStream_t st) {
    if (!p.act16 || p.W != RW || p.H % 16 || p.CO != 32 || !w_row) { set_error("stem row: unsupported shape"); return SDDM_E_INVALID; }
    if (p.nparts != conv_row_nparts(p.H, RW)) { set_error("stem row: nparts mismatch"); return SDDM_E_INVALID; }
    RowArgs a{};
    RowMaps maps;
    memset(&maps, 0, sizeof(maps));
    a.B = p.B; a.H = p.H;
    a.Cin = 2;
    a.w = w_row; a.w_bytes = w_bytes;
    a.bias = p.bias;
    a.cond = p.cond; a.x_t = p.x_t; a.L = p.L; a.hop = p.hop;
    a.parts = p.parts; a.nparts = p.nparts;
    a.gn_on = p.gn_on; a.gn = p.gn;
    int rc;
    if ((rc = encode_rows(&maps.out, p.out, p.B, p.H, RW, 32, 128, 64))) return rc;
    return launch_row_t<ROW_STEM, 1, 0, false, false, false>(a, maps, st);
}

}  // namespace sddm

// debug: enable != 0 -> (re)start tracing: the next 64 conv_row launches record per-role wait cycles of CTA 0;
// enable == 0 -> copy the [64][32] counters to host_out (may be null) and stop tracing.
extern "C" SDDM_API int sddm_debug_row_trace(int enable, long long* host_out) {
    using namespace sddm;
    if (enable) {
        if (!g_row_trace) SDDM_CUDA_TRY(cudaMalloc(&g_row_trace, 64 * 32 * sizeof(long long)));
        SDDM_CUDA_TRY(cudaMemset(g_row_trace, 0, 64 * 32 * sizeof(long long)));
        g_row_trace_launch = 0;
        return SDDM_OK;
    }
    if (!g_row_trace) { set_error("tracing was not enabled"); return SDDM_E_STATE; }
    SDDM_CUDA_TRY(cudaDeviceSynchronize());
    if (host_out) SDDM_CUDA_TRY(cudaMemcpy(host_out, g_row_trace, 64 * 32 * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(g_row_trace);
    g_row_trace = nullptr;
    return SDDM_OK;
}

// debug: enable != 0 -> waits of the row kernel that time out leave a note (CTA, thread, barrier, parity) in a mapped host buffer;
// enable == 0 -> copy the 65536 words to host_out (valid even after the trap killed the context)
extern "C" SDDM_API int sddm_debug_hang(int enable, unsigned* host_out) {
    using namespace sddm;
    static unsigned* h_buf = nullptr;
    if (enable) {
        if (!h_buf) {
            SDDM_CUDA_TRY(cudaHostAlloc(&h_buf, 65536 * sizeof(unsigned), cudaHostAllocMapped));
            unsigned* d = nullptr;
            SDDM_CUDA_TRY(cudaHostGetDevicePointer(&d, h_buf, 0));
            SDDM_CUDA_TRY(set_hang_buffer(d));
            SDDM_CUDA_TRY(conv_tc_set_hang_buffer(d));
        }
        memset(h_buf, 0, 65536 * sizeof(unsigned));
        return SDDM_OK;
    }
    if (!h_buf) { set_error("hang notes were not enabled"); return SDDM_E_STATE; }
    if (host_out) memcpy(host_out, h_buf, 65536 * sizeof(unsigned));
    return SDDM_OK;
}
