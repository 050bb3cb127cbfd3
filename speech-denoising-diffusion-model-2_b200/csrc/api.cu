// C ABI (include/sddm_b200.h): plan lifecycle, weight packing, the static op program of one UNetModified2
// forward, and the reverse-diffusion driver.  Host-side only; kernels live in the other translation units.
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/sddm_b200.h"
#include "kernels.cuh"

namespace sddm {

constexpr int kMaxRows = 65536;   // batch rows per call (size of the plan-owned GroupNorm arrival counters)

static thread_local char g_err[1024] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------------
// network description derived from the config (mirrors UNetModified2.__init__, UNetModified2.py:146-235)
// ---------------------------------------------------------------------------------------------------
struct NodeDesc {
    enum Kind { STEM, RES, DOWN, UP } kind;
    std::string key;   // "downs.3", "mid.0", "ups.7"
    int cin, cout;     // RES: cin is the (concatenated) input channel count
    bool cat;          // RES in `ups`: input = cat(current, skip)
};

struct TensorInfo {
    std::string name;
    int C, H, W, nparts;
    size_t data_off, parts_off;   // float offsets PER SAMPLE (section base = off * B)
};

struct Op {
    enum Kind { STEM, GN, CONV, FINAL } kind;
    // GN
    int gn_src[2] = {-1, -1};
    int gn_nsrc = 0;
    size_t gamma_off = 0, beta_off = 0;   // offsets into the packed fp32 arena
    size_t ss_off = 0;                    // per-sample float offset of [scale(Ctot), shift(Ctot)]
    int Ctot = 0;
    bool gn_fused = false;                // finalised by the preceding producer kernel (gn_fuse.cuh): no launch of its own
    // CONV / FINAL
    int src[2] = {-1, -1};
    int nsrc = 0;
    bool in_gn = false;
    size_t in_ss_off = 0;
    int mode = CONV_S1;
    int out = -1;
    size_t w_off = 0, wtc_off = 0, bias_off = 0;
    size_t wtc_lo_off = 0, reswtc_lo_off = 0;   // split-bf16 mode: lo weight images
    int temb_off = -1;
    int res_kind = 0;   // 0 none, 1 identity, 2 1x1 conv
    size_t resw_off = 0, reswtc_off = 0, resb_off = 0;
    bool use_tc = false;
    bool use_row = false;                 // conv_row.cu (128-wide level, bf16 activations); STEM / CONV / FINAL
    size_t wrow_off = 0;                  // bf16 arena offset of the row-kernel weight image (main chunks, then res_conv chunks)
    uint32_t wrow_bytes = 0;
    // introspection (bench roofline): algorithmic work per batch row
    std::string label;
    double flops = 0.0, bytes = 0.0;
};

struct Prof {   // CUDA-event timing of every launch of the op program (bench.py roofline; off by default)
    bool on = false;
    std::vector<cudaEvent_t> ev;
    std::vector<int> op_of;
    size_t used = 0;
    std::vector<double> ms;
    std::vector<long long> cnt;
};

}  // namespace sddm

using namespace sddm;

struct sddm_plan {
    sddm_config cfg{};
    int H = 0, W = 0;   // frame grid: n_frames x segment_len
    int E = 0;          // concatenated noise-embedding width
    std::vector<NodeDesc> nodes;
    std::map<std::string, std::vector<int64_t>> expect;   // weight name -> shape
    std::map<std::string, std::vector<float>> host_w;
    bool have_sched = false, finalized = false;
    std::vector<float> sch[12];   // order of sddm_schedule fields
    // device-resident packed parameters
    float* d_f32 = nullptr;
    __nv_bfloat16* d_bf16 = nullptr;
    float* d_temb_table = nullptr;   // [T+1][E]
    unsigned int* d_gn_counters = nullptr;   // [kMaxRows] arrival counters of the fused GroupNorm finalisation (all zero between kernels)
    size_t off_freq = 0, off_w1 = 0, off_b1 = 0, off_w2 = 0, off_b2 = 0, off_wn = 0, off_bn = 0;
    size_t off_stem_w = 0, off_stem_b = 0, off_final_w = 0;
    float final_bias = 0.f;
    // program
    std::vector<TensorInfo> tensors;
    std::vector<Op> ops;
    size_t off_x = 0, off_frames = 0, off_temb_rows = 0, off_nl = 0;
    size_t floats_per_sample = 0;
    int launches_per_eps = 0;
    bool post_fused = false;   // the final Block's row kernel also does the overlap-add + posterior update (no post_kernel launch)
    // internal arena for sddm_enhance_host
    cudaStream_t own_stream = nullptr;
    float* d_cond = nullptr;
    float* d_out = nullptr;
    void* d_ws = nullptr;
    int arena_rows = 0;
    unsigned long long* d_seed = nullptr;                 // {seed, row0} read by the graph-captured sampler
    std::map<std::pair<int, int>, cudaGraphExec_t> graphs;   // (rows, variant) -> captured sddm_sample on the arena buffers
    std::map<std::pair<int, int>, int> graph_calls;
    Prof prof;
};

namespace sddm {

static void add_expect(sddm_plan* p, const std::string& k, std::vector<int64_t> shape) { p->expect[k] = std::move(shape); }

static void expect_res(sddm_plan* p, const std::string& key, int cin, int cout, int emb) {
    add_expect(p, key + ".noise_func.noise_func.0.weight", {cout, emb});
    add_expect(p, key + ".noise_func.noise_func.0.bias", {cout});
    add_expect(p, key + ".block1.block.0.weight", {cin});
    add_expect(p, key + ".block1.block.0.bias", {cin});
    add_expect(p, key + ".block1.block.3.weight", {cout, cin, 3, 3});
    add_expect(p, key + ".block1.block.3.bias", {cout});
    add_expect(p, key + ".block2.block.0.weight", {cout});
    add_expect(p, key + ".block2.block.0.bias", {cout});
    add_expect(p, key + ".block2.block.3.weight", {cout, cout, 3, 3});
    add_expect(p, key + ".block2.block.3.bias", {cout});
    if (cin != cout) {
        add_expect(p, key + ".res_conv.weight", {cout, cin, 1, 1});
        add_expect(p, key + ".res_conv.bias", {cout});
    }
}

static int build_nodes(sddm_plan* p) {
    const sddm_config& c = p->cfg;
    const int inner = c.inner_channel, rb = c.res_blocks;
    add_expect(p, "noise_level_mlp.1.weight", {4 * inner, inner});
    add_expect(p, "noise_level_mlp.1.bias", {4 * inner});
    add_expect(p, "noise_level_mlp.3.weight", {inner, 4 * inner});
    add_expect(p, "noise_level_mlp.3.bias", {inner});
    add_expect(p, "downs.0.weight", {inner, c.in_channel, 3, 3});
    add_expect(p, "downs.0.bias", {inner});
    p->nodes.push_back({NodeDesc::STEM, "downs.0", c.in_channel, inner, false});
    std::vector<int> feat{inner};
    int cin = inner, idx = 1;
    for (int l = 0; l < c.n_mults; ++l) {
        const int cout = inner * c.channel_mults[l];
        for (int r = 0; r < rb; ++r) {
            const std::string key = "downs." + std::to_string(idx++);
            p->nodes.push_back({NodeDesc::RES, key, cin, cout, false});
            expect_res(p, key, cin, cout, inner);
            feat.push_back(cout);
            cin = cout;
        }
        const std::string key = "downs." + std::to_string(idx++);
        p->nodes.push_back({NodeDesc::DOWN, key, cout, cout, false});
        add_expect(p, key + ".conv.weight", {cout, cout, 3, 3});
        add_expect(p, key + ".conv.bias", {cout});
        feat.push_back(cout);
    }
    p->nodes.push_back({NodeDesc::RES, "mid.0", cin, cin, false});
    expect_res(p, "mid.0", cin, cin, inner);
    idx = 0;
    int cout = cin;
    for (int l = c.n_mults - 1; l >= 0; --l) {
        cin = inner * c.channel_mults[l];
        cout = cin;
        {
            const std::string key = "ups." + std::to_string(idx++);
            const int ctot = cin + feat.back();
            feat.pop_back();
            p->nodes.push_back({NodeDesc::RES, key, ctot, cout, true});
            expect_res(p, key, ctot, cout, inner);
        }
        {
            const std::string key = "ups." + std::to_string(idx++);
            p->nodes.push_back({NodeDesc::UP, key, cout, cout, false});
            add_expect(p, key + ".conv.weight", {cout, cout, 3, 3});
            add_expect(p, key + ".conv.bias", {cout});
        }
        cout = (l == 0) ? inner : inner * c.channel_mults[l - 1];
        for (int r = 0; r < rb; ++r) {
            const std::string key = "ups." + std::to_string(idx++);
            const int ctot = cin + feat.back();
            feat.pop_back();
            p->nodes.push_back({NodeDesc::RES, key, ctot, cout, true});
            expect_res(p, key, ctot, cout, inner);
            cin = cout;
        }
    }
    add_expect(p, "final_conv.block.0.weight", {cout});
    add_expect(p, "final_conv.block.0.bias", {cout});
    add_expect(p, "final_conv.block.3.weight", {c.out_channel, cout, 3, 3});
    add_expect(p, "final_conv.block.3.bias", {c.out_channel});
    return SDDM_OK;
}

// ---------------------------------------------------------------------------------------------------
// packing
// ---------------------------------------------------------------------------------------------------
struct Arena {
    std::vector<float> f;
    std::vector<__nv_bfloat16> h;
    size_t put(const std::vector<float>& v) {
        size_t o = align_up(f.size(), 32);
        f.resize(o);
        f.insert(f.end(), v.begin(), v.end());
        return o;
    }
    size_t put_h(const std::vector<__nv_bfloat16>& v) {
        size_t o = align_up(h.size(), 64);
        h.resize(o);
        h.insert(h.end(), v.begin(), v.end());
        return o;
    }
};

// [Cout][Cin][k][k] -> fp32 [Cin/8][k*k][8][Cout]
static std::vector<float> pack_conv_f32(const std::vector<float>& w, int cout, int cin, int kk) {
    std::vector<float> o((size_t)cin * kk * cout);
    for (int ci = 0; ci < cin; ++ci)
        for (int t = 0; t < kk; ++t)
            for (int co = 0; co < cout; ++co)
                o[((size_t)((ci / 8) * kk + t) * 8 + (ci % 8)) * cout + co] = w[((size_t)co * cin + ci) * kk + t];
    return o;
}

// [Cout][Cin][k][k] -> bf16 [Cin/16][k*k][2][Cout][8]: per (16-channel chunk, tap) a K-major UMMA operand image made of
// 8x8 core matrices (8 couts x 8 cin = 128 contiguous bytes); LBO (K step) = Cout*16 B, SBO (N step) = 128 B.  One chunk
// (all taps) is one contiguous block of 288*Cout bytes (k=3) / 32*Cout bytes (k=1): a single bulk copy in conv_tc.cu.
static std::vector<__nv_bfloat16> pack_conv_tc(const std::vector<float>& w, int cout, int cin, int kk) {
    std::vector<__nv_bfloat16> o((size_t)cin * kk * cout);
    for (int ci = 0; ci < cin; ++ci)
        for (int t = 0; t < kk; ++t)
            for (int co = 0; co < cout; ++co) {
                const int ch = ci / 16, j = (ci % 16) / 8, e = ci % 8;
                o[((((size_t)ch * kk + t) * 2 + j) * cout + co) * 8 + e] = __float2bfloat16(w[((size_t)co * cin + ci) * kk + t]);
            }
    return o;
}

// row kernel (conv_row.cu): [Cout][Cin][3][3] -> bf16 [Cin/16][kx][2][n][8] with n = ky * 32 + cout (Cout = 32, 96 rows) or
// n = ky (Cout = 1, padded to 16 rows): per (16-channel chunk, kx) a K-major no-swizzle UMMA B operand (LBO = rows * 16 B, SBO = 128 B)
static std::vector<__nv_bfloat16> pack_conv_row(const std::vector<float>& w, int cout, int cin) {
    const int rows = cout == 1 ? 16 : 96;
    std::vector<__nv_bfloat16> o((size_t)(cin / 16) * 3 * 2 * rows * 8, __float2bfloat16(0.f));
    for (int ci = 0; ci < cin; ++ci)
        for (int ky = 0; ky < 3; ++ky)
            for (int kx = 0; kx < 3; ++kx)
                for (int co = 0; co < cout; ++co) {
                    const int c16 = ci / 16, half = (ci % 16) / 8, e = ci % 8, n = cout == 1 ? ky : ky * 32 + co;
                    o[((((size_t)c16 * 3 + kx) * 2 + half) * rows + n) * 8 + e] = __float2bfloat16(w[((size_t)co * cin + ci) * 9 + ky * 3 + kx]);
                }
    return o;
}

// stem on the row kernel: one K16 chunk whose slots are [cond hi, x_t hi, cond lo, x_t lo, 0...]: the bf16 weight of input
// channel c multiplies both halves of the split waveform sample
static std::vector<__nv_bfloat16> pack_stem_row(const std::vector<float>& w, int cout) {
    std::vector<__nv_bfloat16> o((size_t)3 * 2 * 96 * 8, __float2bfloat16(0.f));
    for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx)
            for (int co = 0; co < cout; ++co)
                for (int e = 0; e < 4; ++e)
                    o[(((size_t)kx * 2 + 0) * 96 + ky * 32 + co) * 8 + e] = __float2bfloat16(w[((size_t)co * 2 + (e & 1)) * 9 + ky * 3 + kx]);
    return o;
}

// w - float(bf16(w)): the part of the weights a bf16 image drops (its own bf16 image is the "lo" operand of the split-bf16 mode)
static std::vector<float> bf16_residual(const std::vector<float>& w) {
    std::vector<float> r(w.size());
    for (size_t i = 0; i < w.size(); ++i) r[i] = w[i] - __bfloat162float(__float2bfloat16(w[i]));
    return r;
}

static int new_tensor(sddm_plan* p, const std::string& name, int C, int H, int W) {
    TensorInfo t{name, C, H, W, 0, 0, 0};
    p->tensors.push_back(t);
    return (int)p->tensors.size() - 1;
}

static const std::vector<float>& W_(sddm_plan* p, const std::string& k) { return p->host_w.at(k); }

static void fill_conv_op(sddm_plan* p, Arena& a, Op& op, const std::string& wkey, int cout, int cin, bool tc_ok) {
    op.w_off = a.put(pack_conv_f32(W_(p, wkey + ".weight"), cout, cin, 9));
    op.bias_off = a.put(W_(p, wkey + ".bias"));
    if (tc_ok && cin % 32 == 0) {
        op.wtc_off = a.put_h(pack_conv_tc(W_(p, wkey + ".weight"), cout, cin, 9));
        if (p->cfg.precision == SDDM_PREC_BF16X3) op.wtc_lo_off = a.put_h(pack_conv_tc(bf16_residual(W_(p, wkey + ".weight")), cout, cin, 9));
    }
}

static int emit_gn(sddm_plan* p, Arena& a, std::vector<int> srcs, const std::string& gkey, size_t* ss_off_cursor) {
    Op op;
    op.kind = Op::GN;
    op.gn_nsrc = (int)srcs.size();
    int ctot = 0;
    for (size_t i = 0; i < srcs.size(); ++i) {
        op.gn_src[i] = srcs[i];
        ctot += p->tensors[srcs[i]].C;
    }
    op.Ctot = ctot;
    op.gamma_off = a.put(W_(p, gkey + ".weight"));
    op.beta_off = a.put(W_(p, gkey + ".bias"));
    op.ss_off = *ss_off_cursor;
    *ss_off_cursor += align_up((size_t)2 * ctot, 32);
    op.label = "gn:" + gkey;
    for (size_t i = 0; i < srcs.size(); ++i) op.bytes += 8.0 * p->tensors[srcs[i]].nparts * p->tensors[srcs[i]].C;
    op.bytes += 8.0 * ctot;
    p->ops.push_back(op);
    return (int)p->ops.size() - 1;
}

static ConvP shape_probe(const sddm_plan* p, const Op& op) {
    ConvP c{};
    const TensorInfo& o = p->tensors[op.out];
    c.Hout = o.H; c.Wout = o.W; c.Cout = o.C;
    c.Hin = p->tensors[op.src[0]].H; c.Win = p->tensors[op.src[0]].W;
    c.nsrc = op.nsrc; c.mode = op.mode;
    c.Cin = 0;
    for (int i = 0; i < op.nsrc; ++i) { c.src[i].C = p->tensors[op.src[i]].C; c.Cin += c.src[i].C; }
    c.res_identity = op.res_kind == 1;
    c.res_Cin = 0;   // channels of the raw block input read by the 1x1 res_conv
    if (op.res_kind == 2) {
        c.res_nsrc = op.gn_nsrc;
        for (int i = 0; i < op.gn_nsrc; ++i) { c.res_src[i].C = p->tensors[op.gn_src[i]].C; c.res_Cin += c.res_src[i].C; }
    }
    static const float probe_affine = 0.f;   // shape probes carry no buffers: mark "input goes through GroupNorm + Swish" for conv_row_supported
    if (op.in_gn) c.src[0].scale = &probe_affine;
    c.act16 = p->cfg.precision == SDDM_PREC_BF16_ACT;
    return c;
}

static int build_program(sddm_plan* p, Arena& a) {
    const sddm_config& c = p->cfg;
    const bool want_tc = c.precision != SDDM_PREC_FP32;
    const bool act16 = c.precision == SDDM_PREC_BF16_ACT;
    const double esz = act16 ? 2.0 : 4.0;
    bool act16_ok = true;
    const char* no_row = getenv("SDDM_NO_ROW_KERNEL");   // A/B switch: keep every layer on conv_tc.cu / the CUDA-core stem and final conv
    const bool use_row_kernels = act16 && !(no_row && no_row[0] == '1');
    size_t ss_cursor = 0;   // relative; rebased after tensors are laid out
    int temb_cursor = 0;
    std::vector<int> feats;
    int cur = -1, H = p->H, W = p->W;
    auto set_tc = [&](Op& op, const std::string& label) {
        ConvP probe = shape_probe(p, op);
        op.label = "conv:" + label;
        op.flops = 2.0 * 9.0 * probe.Cin * probe.Cout * probe.Hout * probe.Wout + 2.0 * probe.res_Cin * probe.Cout * probe.Hout * probe.Wout;
        op.bytes = esz * ((double)probe.Cin * probe.Hin * probe.Win + (double)probe.Cout * probe.Hout * probe.Wout +
                          (probe.res_identity ? (double)probe.Cout * probe.Hout * probe.Wout : 0.0) +
                          (double)probe.res_Cin * probe.Hin * probe.Win);   // the 1x1 res_conv re-reads the raw block input
        op.use_tc = want_tc && conv_tc_supported(probe);
        op.use_row = want_tc && use_row_kernels && conv_row_supported(probe);
        if (act16 && !op.use_tc) act16_ok = false;
        p->tensors[op.out].nparts = op.use_row ? conv_row_nparts(probe.Hout, probe.Wout)
                                               : (op.use_tc ? conv_tc_nparts(probe.Hout, probe.Wout) : conv_fp32_nparts(probe.Hout, probe.Wout));
    };
    // row-kernel weight image of a convolution op: main chunks, then the 1x1 res_conv chunks ([Cin/16][1][2][32][8] = pack_conv_tc with k = 1)
    auto fill_row_w = [&](Op& op, const std::string& wkey, int cout, int cin, const std::string& reskey, int res_cin) {
        std::vector<__nv_bfloat16> img = pack_conv_row(W_(p, wkey + ".weight"), cout, cin);
        if (res_cin) {
            const std::vector<__nv_bfloat16> r = pack_conv_tc(W_(p, reskey + ".weight"), cout, res_cin, 1);
            img.insert(img.end(), r.begin(), r.end());
        }
        op.wrow_off = a.put_h(img);
        op.wrow_bytes = (uint32_t)(img.size() * sizeof(__nv_bfloat16));
    };
    auto emit_res = [&](const NodeDesc& nd, std::vector<int> srcs) {
        const int gn1 = emit_gn(p, a, srcs, nd.key + ".block1.block.0", &ss_cursor);
        const int h = new_tensor(p, nd.key + ".h", nd.cout, H, W);
        Op c1;
        c1.kind = Op::CONV;
        c1.nsrc = (int)srcs.size();
        for (size_t i = 0; i < srcs.size(); ++i) c1.src[i] = srcs[i];
        c1.in_gn = true;
        c1.in_ss_off = p->ops[gn1].ss_off;
        c1.out = h;
        fill_conv_op(p, a, c1, nd.key + ".block1.block.3", nd.cout, nd.cin, want_tc);
        c1.temb_off = temb_cursor;
        temb_cursor += nd.cout;
        set_tc(c1, nd.key + ".block1");
        if (c1.use_row) fill_row_w(c1, nd.key + ".block1.block.3", nd.cout, nd.cin, "", 0);
        p->ops.push_back(c1);
        const int gn2 = emit_gn(p, a, {h}, nd.key + ".block2.block.0", &ss_cursor);
        const int out = new_tensor(p, nd.key, nd.cout, H, W);
        Op c2;
        c2.kind = Op::CONV;
        c2.nsrc = 1;
        c2.src[0] = h;
        c2.in_gn = true;
        c2.in_ss_off = p->ops[gn2].ss_off;
        c2.out = out;
        fill_conv_op(p, a, c2, nd.key + ".block2.block.3", nd.cout, nd.cout, want_tc);
        if (nd.cin != nd.cout) {
            c2.res_kind = 2;
            c2.resw_off = a.put(pack_conv_f32(W_(p, nd.key + ".res_conv.weight"), nd.cout, nd.cin, 1));
            c2.resb_off = a.put(W_(p, nd.key + ".res_conv.bias"));
            if (want_tc && nd.cin % 32 == 0) {
                c2.reswtc_off = a.put_h(pack_conv_tc(W_(p, nd.key + ".res_conv.weight"), nd.cout, nd.cin, 1));
                if (c.precision == SDDM_PREC_BF16X3) c2.reswtc_lo_off = a.put_h(pack_conv_tc(bf16_residual(W_(p, nd.key + ".res_conv.weight")), nd.cout, nd.cin, 1));
            }
            // the residual reads the raw block input; record its sources in gn_src (unused by CONV otherwise)
            c2.gn_nsrc = (int)srcs.size();
            for (size_t i = 0; i < srcs.size(); ++i) c2.gn_src[i] = srcs[i];
        } else {
            c2.res_kind = 1;
            c2.gn_nsrc = 1;
            c2.gn_src[0] = srcs[0];
        }
        set_tc(c2, nd.key + ".block2");
        if (c2.use_row) fill_row_w(c2, nd.key + ".block2.block.3", nd.cout, nd.cout, nd.key + ".res_conv", c2.res_kind == 2 ? nd.cin : 0);
        p->ops.push_back(c2);
        return out;
    };
    auto emit_plain = [&](const NodeDesc& nd, int src, int mode, int Ho, int Wo) {
        const int out = new_tensor(p, nd.key, nd.cout, Ho, Wo);
        Op cv;
        cv.kind = Op::CONV;
        cv.nsrc = 1;
        cv.src[0] = src;
        cv.mode = mode;
        cv.out = out;
        fill_conv_op(p, a, cv, nd.key + ".conv", nd.cout, nd.cin, want_tc);
        set_tc(cv, nd.key + ".conv");
        if (cv.use_row) fill_row_w(cv, nd.key + ".conv", nd.cout, nd.cin, "", 0);
        p->ops.push_back(cv);
        return out;
    };

    // weights of the embedding MLP and the concatenated per-block Linear layers
    {
        const int inner = c.inner_channel;
        std::vector<float> freq;
        auto it = p->host_w.find("noise_level_mlp.0.embedding_vector");
        if (it != p->host_w.end()) {
            freq = it->second;
        } else {   // UNetModified2.py:55: 1e4 * 10^(-4k/half)
            for (int k = 0; k < inner / 2; ++k) freq.push_back((float)(1e4 * std::pow(10.0, -(double)k * 4.0 / (inner / 2))));
        }
        p->off_freq = a.put(freq);
        p->off_w1 = a.put(W_(p, "noise_level_mlp.1.weight"));
        p->off_b1 = a.put(W_(p, "noise_level_mlp.1.bias"));
        p->off_w2 = a.put(W_(p, "noise_level_mlp.3.weight"));
        p->off_b2 = a.put(W_(p, "noise_level_mlp.3.bias"));
        std::vector<float> wn, bn;
        for (const NodeDesc& nd : p->nodes)
            if (nd.kind == NodeDesc::RES) {
                const auto& w = W_(p, nd.key + ".noise_func.noise_func.0.weight");
                const auto& b = W_(p, nd.key + ".noise_func.noise_func.0.bias");
                wn.insert(wn.end(), w.begin(), w.end());
                bn.insert(bn.end(), b.begin(), b.end());
            }
        p->E = (int)bn.size();
        p->off_wn = a.put(wn);
        p->off_bn = a.put(bn);
    }

    for (const NodeDesc& nd : p->nodes) {
        switch (nd.kind) {
            case NodeDesc::STEM: {
                const auto& w = W_(p, "downs.0.weight");   // [CO][2][3][3] -> [2][9][CO]
                std::vector<float> pk((size_t)18 * nd.cout);
                for (int co = 0; co < nd.cout; ++co)
                    for (int ci = 0; ci < 2; ++ci)
                        for (int t = 0; t < 9; ++t) pk[((size_t)ci * 9 + t) * nd.cout + co] = w[((size_t)co * 2 + ci) * 9 + t];
                p->off_stem_w = a.put(pk);
                p->off_stem_b = a.put(W_(p, "downs.0.bias"));
                cur = new_tensor(p, nd.key, nd.cout, H, W);
                Op op;
                op.kind = Op::STEM;
                op.use_row = use_row_kernels && W == 128 && H % 16 == 0 && nd.cout == 32;
                p->tensors[cur].nparts = op.use_row ? conv_row_nparts(H, W) : stem_nparts(H, W);
                if (op.use_row) {
                    const std::vector<__nv_bfloat16> img = pack_stem_row(w, nd.cout);
                    op.wrow_off = a.put_h(img);
                    op.wrow_bytes = (uint32_t)(img.size() * sizeof(__nv_bfloat16));
                }
                op.out = cur;
                op.label = "stem:downs.0";
                op.flops = 2.0 * 18.0 * nd.cout * H * W;
                op.bytes = 8.0 * c.num_samples + esz * (double)nd.cout * H * W;
                p->ops.push_back(op);
                feats.push_back(cur);
                break;
            }
            case NodeDesc::RES: {
                std::vector<int> srcs{cur};
                if (nd.cat) {
                    srcs.push_back(feats.back());
                    feats.pop_back();
                }
                cur = emit_res(nd, srcs);
                if (nd.key.rfind("downs.", 0) == 0) feats.push_back(cur);
                break;
            }
            case NodeDesc::DOWN:
                if (H % 2 || W % 2) { set_error("Downsample of odd size %dx%d", H, W); return SDDM_E_INVALID; }
                H /= 2; W /= 2;
                cur = emit_plain(nd, cur, CONV_S2, H, W);
                feats.push_back(cur);
                break;
            case NodeDesc::UP:
                H *= 2; W *= 2;
                cur = emit_plain(nd, cur, CONV_UP, H, W);
                break;
        }
    }
    // final Block
    {
        const int gnf = emit_gn(p, a, {cur}, "final_conv.block.0", &ss_cursor);
        const auto& w = W_(p, "final_conv.block.3.weight");   // [1][C][3][3] -> [9][C]
        const int C = p->tensors[cur].C;
        std::vector<float> pk((size_t)9 * C);
        for (int ci = 0; ci < C; ++ci)
            for (int t = 0; t < 9; ++t) pk[(size_t)t * C + ci] = w[(size_t)ci * 9 + t];
        p->off_final_w = a.put(pk);
        p->final_bias = W_(p, "final_conv.block.3.bias")[0];
        Op op;
        op.kind = Op::FINAL;
        op.nsrc = 1;
        op.src[0] = cur;
        op.in_gn = true;
        op.in_ss_off = p->ops[gnf].ss_off;
        op.label = "final_conv";
        op.use_row = use_row_kernels && p->W == 128 && p->H % 16 == 0 && C == 32;
        if (op.use_row) {
            const std::vector<__nv_bfloat16> img = pack_conv_row(w, 1, C);
            op.wrow_off = a.put_h(img);
            op.wrow_bytes = (uint32_t)(img.size() * sizeof(__nv_bfloat16));
        }
        op.flops = 2.0 * 9.0 * C * p->H * p->W;
        op.bytes = esz * (double)C * p->H * p->W + 4.0 * (double)p->H * p->W;   // activation in, frames (or, fused: x_t in + x_{t-1} out) 
        p->ops.push_back(op);
    }
    if (temb_cursor != p->E) { set_error("internal: embedding width mismatch"); return SDDM_E_INVALID; }
    if (!act16_ok) { set_error("bf16 activation storage needs every convolution on the tcgen05 path (channel counts multiples of 32)"); return SDDM_E_INVALID; }

    // workspace layout (per-sample float offsets)
    size_t off = 0;
    auto take = [&](size_t n) { size_t o = off; off += align_up(n, 32); return o; };
    p->off_x = take((size_t)c.num_samples);
    p->off_frames = take((size_t)p->H * p->W);
    p->off_temb_rows = take((size_t)p->E);
    p->off_nl = take(32);
    for (TensorInfo& t : p->tensors) {
        t.data_off = take(act16 ? ((size_t)t.C * t.H * t.W + 1) / 2 : (size_t)t.C * t.H * t.W);
        t.parts_off = take((size_t)t.nparts * t.C * 2);
    }
    const size_t ss_base = take(ss_cursor);
    for (Op& op : p->ops) {
        if (op.kind == Op::GN) op.ss_off += ss_base;
        if (op.in_gn) op.in_ss_off += ss_base;
    }
    p->floats_per_sample = off;
    // a GroupNorm whose newest source comes out of a tcgen05 convolution or the stem is finalised inside that producer
    int fused = 0;
    for (size_t i = 1; i < p->ops.size(); ++i) {
        Op& g = p->ops[i];
        const Op& prev = p->ops[i - 1];
        if (g.kind != Op::GN || !want_tc) continue;
        // (conv_tc.cu and conv_row.cu give every CTA a contiguous run of tiles / rows: with a round-robin schedule the CTA that
        //  finalises sample n falls behind and is then the last arriver for every following sample - measured 57 -> 606 us)
        const bool producer_ok = prev.kind == Op::STEM ? prev.use_row : (prev.kind == Op::CONV && prev.use_tc);
        bool small = true;   // the last-arriving CTA walks every partial of the sample: keep that walk short
        for (int k = 0; k < g.gn_nsrc; ++k) small = small && p->tensors[g.gn_src[k]].nparts <= 256;
        static const bool no_fuse = [] { const char* e = getenv("SDDM_NO_GN_FUSE"); return e && e[0] == '1'; }();   // A/B switch: separate GroupNorm finalisation kernels
        if (producer_ok && small && prev.out == g.gn_src[0] && !no_fuse) { g.gn_fused = true; ++fused; }
    }
    const char* no_pf = getenv("SDDM_NO_POST_FUSE");   // A/B switch: keep the separate overlap-add / posterior kernel
    p->post_fused = p->ops.back().kind == Op::FINAL && p->ops.back().use_row && 2 * c.segment_stride == c.segment_len && !(no_pf && no_pf[0] == '1');
    p->launches_per_eps = (int)p->ops.size() - fused + (p->post_fused ? 0 : 1);   // + the overlap-add / posterior kernel unless the final Block does it
    return SDDM_OK;
}

// ---------------------------------------------------------------------------------------------------
// execution
// ---------------------------------------------------------------------------------------------------
static inline float* sect(const sddm_plan*, void* ws, size_t off, int B) { return reinterpret_cast<float*>(ws) + off * (size_t)B; }

static int prof_flush(sddm_plan* p) {
    Prof& pr = p->prof;
    if (pr.used == 0) return SDDM_OK;
    SDDM_CUDA_TRY(cudaEventSynchronize(pr.ev[2 * pr.used - 1]));
    for (size_t k = 0; k < pr.used; ++k) {
        float ms = 0.f;
        SDDM_CUDA_TRY(cudaEventElapsedTime(&ms, pr.ev[2 * k], pr.ev[2 * k + 1]));
        pr.ms[pr.op_of[k]] += ms;
        pr.cnt[pr.op_of[k]] += 1;
    }
    pr.used = 0;
    return SDDM_OK;
}

static int prof_mark(sddm_plan* p, int op, bool begin, cudaStream_t st) {
    Prof& pr = p->prof;
    if (!pr.on) return SDDM_OK;
    if (begin) {
        if (2 * pr.used + 2 > pr.ev.size()) {
            if (pr.ev.size() >= 65536) {
                int rc = prof_flush(p);
                if (rc) return rc;
            } else {
                for (int i = 0; i < 1024; ++i) {
                    cudaEvent_t e;
                    SDDM_CUDA_TRY(cudaEventCreate(&e));
                    pr.ev.push_back(e);
                }
                pr.op_of.resize(pr.ev.size() / 2);
            }
        }
        pr.op_of[pr.used] = op;
        SDDM_CUDA_TRY(cudaEventRecord(pr.ev[2 * pr.used], st));
    } else {
        SDDM_CUDA_TRY(cudaEventRecord(pr.ev[2 * pr.used + 1], st));
        pr.used += 1;
    }
    return SDDM_OK;
}

// descriptor of a GroupNorm op for the producer-side finalisation (the producer of gn_src[0] runs it)
static void fill_gn_fuse(const sddm_plan* p, const Op& op, int B, void* ws, int expect, GnFuse* g) {
    g->nsrc = op.gn_nsrc;
    for (int i = 0; i < op.gn_nsrc; ++i) {
        const TensorInfo& t = p->tensors[op.gn_src[i]];
        g->parts[i] = sect(p, ws, t.parts_off, B);
        g->C[i] = t.C;
        g->nparts[i] = t.nparts;
    }
    const TensorInfo& t0 = p->tensors[op.gn_src[0]];
    g->gamma = p->d_f32 + op.gamma_off;
    g->beta = p->d_f32 + op.beta_off;
    g->scale = sect(p, ws, op.ss_off, B);
    g->shift = g->scale + (size_t)op.Ctot * B;
    g->Ctot = op.Ctot; g->groups = p->cfg.norm_groups; g->HW = t0.H * t0.W; g->eps = 1e-5f;
    g->counter = p->d_gn_counters;
    g->expect = expect;
}

// one UNetModified2 forward up to the final conv frames (ws.frames); temb: device pointer, row stride
static int run_unet(sddm_plan* p, const float* cond, const float* x_t, const float* temb, int temb_stride, int B, void* ws,
                    cudaStream_t st, const PostP* post = nullptr, const float* post_k8 = nullptr, bool keep_frames = true) {
    const sddm_config& c = p->cfg;
    static const bool sync_each = [] { const char* e = getenv("SDDM_SYNC_EACH_OP"); return e && e[0] == '1'; }();
    for (size_t oi = 0; oi < p->ops.size(); ++oi) {
        const Op& op = p->ops[oi];
        if (op.kind == Op::GN && op.gn_fused) continue;
        const Op* fuse = (oi + 1 < p->ops.size() && p->ops[oi + 1].kind == Op::GN && p->ops[oi + 1].gn_fused) ? &p->ops[oi + 1] : nullptr;
        int rc = prof_mark(p, (int)oi, true, st);
        if (rc != SDDM_OK) return rc;
        switch (op.kind) {
            case Op::STEM: {
                const TensorInfo& o = p->tensors[op.out];
                StemP sp{cond, x_t, p->d_f32 + p->off_stem_w, p->d_f32 + p->off_stem_b, sect(p, ws, o.data_off, B),
                         sect(p, ws, o.parts_off, B), B, c.num_samples, o.H, o.W, c.segment_stride, o.C, o.nparts};
                sp.act16 = c.precision == SDDM_PREC_BF16_ACT;
                if (fuse) { sp.gn_on = 1; fill_gn_fuse(p, *fuse, B, ws, op.use_row ? conv_row_arrivals(o.H) : o.nparts, &sp.gn); }
                rc = op.use_row ? launch_stem_row(sp, p->d_bf16 + op.wrow_off, op.wrow_bytes, st) : launch_stem(sp, st);
                break;
            }
            case Op::GN: {
                GnP g{};
                g.nsrc = op.gn_nsrc;
                for (int i = 0; i < op.gn_nsrc; ++i) {
                    const TensorInfo& t = p->tensors[op.gn_src[i]];
                    g.parts[i] = sect(p, ws, t.parts_off, B);
                    g.C[i] = t.C;
                    g.nparts[i] = t.nparts;
                }
                const TensorInfo& t0 = p->tensors[op.gn_src[0]];
                g.gamma = p->d_f32 + op.gamma_off;
                g.beta = p->d_f32 + op.beta_off;
                g.scale = sect(p, ws, op.ss_off, B);
                g.shift = g.scale + (size_t)op.Ctot * B;
                g.B = B; g.Ctot = op.Ctot; g.groups = c.norm_groups; g.HW = t0.H * t0.W; g.eps = 1e-5f;
                rc = launch_gn_finalize(g, st);
                break;
            }
            case Op::CONV: {
                ConvP cp{};
                const TensorInfo& o = p->tensors[op.out];
                cp.nsrc = op.nsrc;
                cp.Cin = 0;
                for (int i = 0; i < op.nsrc; ++i) cp.Cin += p->tensors[op.src[i]].C;
                for (int i = 0; i < op.nsrc; ++i) {
                    const TensorInfo& t = p->tensors[op.src[i]];
                    cp.src[i].x = sect(p, ws, t.data_off, B);
                    cp.src[i].C = t.C;
                    if (op.in_gn) {
                        cp.src[i].scale = sect(p, ws, op.in_ss_off, B);
                        cp.src[i].shift = cp.src[i].scale + (size_t)cp.Cin * B;
                    }
                }
                cp.Hin = p->tensors[op.src[0]].H; cp.Win = p->tensors[op.src[0]].W;
                cp.Hout = o.H; cp.Wout = o.W; cp.Cout = o.C; cp.mode = op.mode;
                cp.w = p->d_f32 + op.w_off;
                cp.w_tc = p->d_bf16 ? p->d_bf16 + op.wtc_off : nullptr;
                cp.x3 = c.precision == SDDM_PREC_BF16X3;
                cp.w_tc_lo = (cp.x3 && p->d_bf16) ? p->d_bf16 + op.wtc_lo_off : nullptr;
                cp.bias = p->d_f32 + op.bias_off;
                if (op.temb_off >= 0) { cp.temb = temb + op.temb_off; cp.temb_stride = temb_stride; }
                if (op.res_kind) {
                    cp.res_nsrc = op.gn_nsrc;
                    cp.res_Cin = 0;
                    for (int i = 0; i < op.gn_nsrc; ++i) {
                        const TensorInfo& t = p->tensors[op.gn_src[i]];
                        cp.res_src[i].x = sect(p, ws, t.data_off, B);
                        cp.res_src[i].C = t.C;
                        cp.res_Cin += t.C;
                    }
                    if (op.res_kind == 1) {
                        cp.res_identity = 1;
                    } else {
                        cp.res_w = p->d_f32 + op.resw_off;
                        cp.res_w_tc = p->d_bf16 ? p->d_bf16 + op.reswtc_off : nullptr;
                        cp.res_w_tc_lo = (cp.x3 && p->d_bf16) ? p->d_bf16 + op.reswtc_lo_off : nullptr;
                        cp.res_bias = p->d_f32 + op.resb_off;
                    }
                }
                cp.out = sect(p, ws, o.data_off, B);
                cp.parts = sect(p, ws, o.parts_off, B);
                cp.nparts = o.nparts;
                cp.B = B;
                cp.act16 = c.precision == SDDM_PREC_BF16_ACT;
                if (fuse) { cp.gn_on = 1; fill_gn_fuse(p, *fuse, B, ws, op.use_row ? conv_row_arrivals(o.H) : conv_tc_tiles(o.H, o.W), &cp.gn); }
                if (op.use_row) rc = launch_conv_row(cp, p->d_bf16 + op.wrow_off, op.wrow_bytes, nullptr, 0.f, st, nullptr, nullptr);
                else rc = op.use_tc ? launch_conv_tc(cp, st) : launch_conv_fp32(cp, st);
                break;
            }
            case Op::FINAL: {
                const TensorInfo& t = p->tensors[op.src[0]];
                if (op.use_row) {
                    ConvP cp{};
                    cp.nsrc = 1; cp.Cin = t.C; cp.Cout = 1;
                    cp.src[0].x = sect(p, ws, t.data_off, B); cp.src[0].C = t.C;
                    cp.src[0].scale = sect(p, ws, op.in_ss_off, B);
                    cp.src[0].shift = cp.src[0].scale + (size_t)t.C * B;
                    cp.Hin = cp.Hout = t.H; cp.Win = cp.Wout = t.W; cp.mode = CONV_S1; cp.B = B; cp.act16 = 1;
                    // fused tail: overlap-add + posterior (or just eps_hat) in the epilogue; frames only for sddm_eps / debug fetch
                    rc = launch_conv_row(cp, p->d_bf16 + op.wrow_off, op.wrow_bytes, (keep_frames || !p->post_fused) ? sect(p, ws, p->off_frames, B) : nullptr,
                                         p->final_bias, st, p->post_fused ? post : nullptr, post_k8);
                    break;
                }
                FinalP f{};
                f.x = sect(p, ws, t.data_off, B);
                f.scale = sect(p, ws, op.in_ss_off, B);
                f.shift = f.scale + (size_t)t.C * B;
                f.w = p->d_f32 + p->off_final_w;
                f.bias = p->final_bias;
                f.frames = sect(p, ws, p->off_frames, B);
                f.B = B; f.H = t.H; f.W = t.W; f.C = t.C;
                f.fast_math = c.precision != SDDM_PREC_FP32;
                f.act16 = c.precision == SDDM_PREC_BF16_ACT;
                rc = launch_final_conv(f, st);
                break;
            }
        }
        if (rc != SDDM_OK) return rc;
        if ((rc = prof_mark(p, (int)oi, false, st))) return rc;
        if (sync_each) {   // debug: name the op whose kernel faults
            const cudaError_t e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { set_error("op %zu (%s) failed: %s", oi, op.label.c_str(), cudaGetErrorString(e)); return SDDM_E_CUDA; }
        }
    }
    return SDDM_OK;
}

static void step_coefs(const sddm_plan* p, int variant, int t, float* k8) {
    enum { BETAS, ALPHAS, SAB, PNC, SIGMA, SGAM, SSIG, SQD, CXT, CYT, CEPS, SDE };
    volatile float gam = p->sch[SGAM][t];
    volatile float omg = 1.0f - gam;
    float sig;
    switch (variant) {
        case SDDM_VAR_SR3: sig = sqrtf(p->sch[BETAS][t]); break;
        case SDDM_VAR_SUPPORTIVE: sig = fmaxf(0.0f, p->sch[SSIG][t]); break;
        case SDDM_VAR_CONDITIONAL: sig = p->sch[SDE][t]; break;
        default: sig = p->sch[SIGMA][t];
    }
    k8[0] = p->sch[PNC][t];
    k8[1] = sqrtf(p->sch[ALPHAS][t]);   // alphas[t] ** 0.5 (correctly rounded, as torch's pow(.,0.5) -> sqrt)
    k8[2] = sig;
    k8[3] = gam;
    k8[4] = omg;
    k8[5] = p->sch[CXT][t];
    k8[6] = p->sch[CYT][t];
    k8[7] = p->sch[CEPS][t];
}

static int check_ready(const sddm_plan* p) {
    if (!p) { set_error("null plan"); return SDDM_E_INVALID; }
    if (!p->finalized) { set_error("plan not finalised (load weights, set schedule, call sddm_plan_finalize)"); return SDDM_E_STATE; }
    return SDDM_OK;
}

static int check_variant(int v) {
    if (v < SDDM_VAR_ORIGINAL || v > SDDM_VAR_CONDITIONAL) { set_error("unknown p_transition variant %d", v); return SDDM_E_INVALID; }
    return SDDM_OK;
}

}  // namespace sddm

// ===================================================================================================
// C ABI
// ===================================================================================================
extern "C" {

const char* sddm_last_error(void) { return g_err; }
int sddm_version(void) { return 100; }
uint64_t sddm_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int sddm_plan_create(const sddm_config* cfg, sddm_plan** out) {
    if (!cfg || !out) { set_error("null argument"); return SDDM_E_INVALID; }
    const sddm_config& c = *cfg;
    if (c.in_channel != 2 || c.out_channel != 1) { set_error("in_channel must be 2 and out_channel 1 (got %d, %d)", c.in_channel, c.out_channel); return SDDM_E_INVALID; }
    if (c.n_mults < 1 || c.n_mults > 8 || c.res_blocks < 1 || c.n_timestep < 1) { set_error("bad n_mults / res_blocks / n_timestep"); return SDDM_E_INVALID; }
    if (c.segment_len <= 0 || c.segment_stride <= 0 || c.num_samples < c.segment_len ||
        (c.num_samples - c.segment_len) % c.segment_stride != 0) {   // UNetModified2.py:13
        set_error("(num_samples - segment_len) %% segment_stride must be 0");
        return SDDM_E_INVALID;
    }
    if (c.num_samples % 4 || c.segment_len % 4 || c.segment_stride % 4) { set_error("num_samples, segment_len, segment_stride must be multiples of 4"); return SDDM_E_INVALID; }
    if (c.inner_channel != 32 && c.inner_channel != 64) { set_error("inner_channel must be 32 or 64"); return SDDM_E_INVALID; }
    if (c.norm_groups < 1 || c.inner_channel % c.norm_groups) { set_error("inner_channel must be divisible by norm_groups"); return SDDM_E_INVALID; }
    if (c.precision != SDDM_PREC_FP32 && c.precision != SDDM_PREC_BF16 && c.precision != SDDM_PREC_BF16_ACT && c.precision != SDDM_PREC_BF16X3) { set_error("unknown precision %d", c.precision); return SDDM_E_INVALID; }
    const int H = (c.num_samples - c.segment_len) / c.segment_stride + 1, W = c.segment_len;
    // every level must tile: the deepest (H/2^n x W/2^n) by 8x4, all others by 16x8
    if (H % (1 << c.n_mults) || W % (1 << c.n_mults) || (H >> c.n_mults) % 8 || (W >> c.n_mults) % 4) {
        set_error("frame grid %dx%d does not tile through %d levels", H, W, c.n_mults);
        return SDDM_E_INVALID;
    }
    for (int i = 0; i < c.n_mults; ++i)
        if (c.channel_mults[i] < 1) { set_error("channel_mults must be positive"); return SDDM_E_INVALID; }
    sddm_plan* p = new sddm_plan();
    p->cfg = c;
    p->H = H;
    p->W = W;
    build_nodes(p);
    *out = p;
    return SDDM_OK;
}

void sddm_plan_destroy(sddm_plan* p) {
    if (!p) return;
    cudaFree(p->d_f32);
    cudaFree(p->d_bf16);
    cudaFree(p->d_temb_table);
    cudaFree(p->d_gn_counters);
    cudaFree(p->d_cond);
    cudaFree(p->d_out);
    cudaFree(p->d_ws);
    cudaFree(p->d_seed);
    for (auto& kv : p->graphs) cudaGraphExecDestroy(kv.second);
    if (p->own_stream) cudaStreamDestroy(p->own_stream);
    for (cudaEvent_t e : p->prof.ev) cudaEventDestroy(e);
    delete p;
}

int sddm_plan_load_weight(sddm_plan* p, const char* name, const void* data, const int64_t* shape, int ndim) {
    if (!p || !name || !data || (!shape && ndim > 0)) { set_error("null argument"); return SDDM_E_INVALID; }
    if (p->finalized) { set_error("plan already finalised"); return SDDM_E_STATE; }
    const std::string key(name);
    size_t numel = 1;
    for (int i = 0; i < ndim; ++i) numel *= (size_t)shape[i];
    if (key == "noise_level_mlp.0.embedding_vector") {
        if (numel != (size_t)p->cfg.inner_channel / 2) { set_error("embedding_vector must have inner_channel/2 entries"); return SDDM_E_INVALID; }
    } else {
        auto it = p->expect.find(key);
        if (it == p->expect.end()) { set_error("unexpected weight '%s'", name); return SDDM_E_INVALID; }
        const std::vector<int64_t>& e = it->second;
        bool ok = (int)e.size() == ndim;
        for (int i = 0; ok && i < ndim; ++i) ok = e[i] == shape[i];
        if (!ok) { set_error("weight '%s': shape mismatch", name); return SDDM_E_INVALID; }
    }
    std::vector<float> v(numel);
    cudaPointerAttributes attr{};
    const cudaError_t qe = cudaPointerGetAttributes(&attr, data);
    if (qe == cudaSuccess && (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged)) {
        SDDM_CUDA_TRY(cudaMemcpy(v.data(), data, numel * sizeof(float), cudaMemcpyDeviceToHost));
    } else {   // plain host memory (also the only possibility when no device is present)
        if (qe != cudaSuccess) cudaGetLastError();
        std::memcpy(v.data(), data, numel * sizeof(float));
    }
    p->host_w[key] = std::move(v);
    return SDDM_OK;
}

int sddm_plan_set_schedule(sddm_plan* p, const sddm_schedule* s, int n) {
    if (!p || !s) { set_error("null argument"); return SDDM_E_INVALID; }
    if (n != p->cfg.n_timestep + 1) { set_error("schedule length %d != n_timestep + 1 = %d", n, p->cfg.n_timestep + 1); return SDDM_E_INVALID; }
    const float* ptrs[12] = {s->betas, s->alphas, s->sqrt_alpha_bar, s->predicted_noise_coeff, s->sigma, s->supportive_gamma,
                             s->supportive_sigma_hat, s->sqrt_delta, s->c_xt, s->c_yt, s->c_epst, s->sqrt_delta_estimated};
    for (int i = 0; i < 12; ++i) {
        if (!ptrs[i]) { set_error("schedule table %d is null", i); return SDDM_E_INVALID; }
        p->sch[i].assign(ptrs[i], ptrs[i] + n);
    }
    p->have_sched = true;
    return SDDM_OK;
}

int sddm_plan_finalize(sddm_plan* p) {
    if (!p) { set_error("null plan"); return SDDM_E_INVALID; }
    if (p->finalized) return SDDM_OK;
    if (!p->have_sched) { set_error("schedule not set"); return SDDM_E_STATE; }
    for (const auto& kv : p->expect)
        if (!p->host_w.count(kv.first)) { set_error("missing weight '%s'", kv.first.c_str()); return SDDM_E_STATE; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device: this library has no CPU fallback");
        return SDDM_E_CUDA;
    }
    Arena a;
    int rc = build_program(p, a);
    if (rc != SDDM_OK) return rc;
    SDDM_CUDA_TRY(cudaMalloc(&p->d_f32, a.f.size() * sizeof(float)));
    SDDM_CUDA_TRY(cudaMemcpy(p->d_f32, a.f.data(), a.f.size() * sizeof(float), cudaMemcpyHostToDevice));
    if (!a.h.empty()) {
        SDDM_CUDA_TRY(cudaMalloc(&p->d_bf16, a.h.size() * sizeof(__nv_bfloat16)));
        SDDM_CUDA_TRY(cudaMemcpy(p->d_bf16, a.h.data(), a.h.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
    }
    // per-timestep embedding table: row t = E(noise level sqrt_alpha_bar[t])
    const int T1 = p->cfg.n_timestep + 1;
    float* d_nl = nullptr;
    SDDM_CUDA_TRY(cudaMalloc(&p->d_temb_table, (size_t)T1 * p->E * sizeof(float)));
    SDDM_CUDA_TRY(cudaMalloc(&d_nl, T1 * sizeof(float)));
    SDDM_CUDA_TRY(cudaMemcpy(d_nl, p->sch[2].data(), T1 * sizeof(float), cudaMemcpyHostToDevice));
    TembP tp{d_nl, p->d_f32 + p->off_freq, p->d_f32 + p->off_w1, p->d_f32 + p->off_b1, p->d_f32 + p->off_w2, p->d_f32 + p->off_b2,
             p->d_f32 + p->off_wn, p->d_f32 + p->off_bn, p->d_temb_table, T1, p->cfg.inner_channel, p->E};
    rc = launch_temb(tp, nullptr);
    if (rc != SDDM_OK) { cudaFree(d_nl); return rc; }
    SDDM_CUDA_TRY(cudaMalloc(&p->d_gn_counters, kMaxRows * sizeof(unsigned int)));
    SDDM_CUDA_TRY(cudaMemset(p->d_gn_counters, 0, kMaxRows * sizeof(unsigned int)));
    SDDM_CUDA_TRY(cudaDeviceSynchronize());
    cudaFree(d_nl);
    p->host_w.clear();
    p->finalized = true;
    return SDDM_OK;
}

size_t sddm_workspace_bytes(const sddm_plan* p, int B) {
    if (!p || !p->finalized || B <= 0) return 0;
    return p->floats_per_sample * (size_t)B * sizeof(float) + 256;
}

int sddm_plan_launches_per_eps(const sddm_plan* p) { return (p && p->finalized) ? p->launches_per_eps : 0; }

static int check_ws(const sddm_plan* p, int B, const void* ws, size_t ws_bytes) {
    if (B <= 0 || B > kMaxRows) { set_error("batch must be in [1, %d]", kMaxRows); return SDDM_E_INVALID; }
    if (!ws || ws_bytes < sddm_workspace_bytes(p, B)) { set_error("workspace too small: %zu < %zu", ws_bytes, sddm_workspace_bytes(p, B)); return SDDM_E_WORKSPACE; }
    if ((uintptr_t)ws % 256) { set_error("workspace must be 256-byte aligned"); return SDDM_E_INVALID; }
    return SDDM_OK;
}

int sddm_eps(sddm_plan* p, const float* cond, const float* x_t, const float* noise_level, int t, float* eps_out, int B, void* ws,
             size_t ws_bytes, void* stream) {
    int rc = check_ready(p);
    if (rc) return rc;
    if ((rc = check_ws(p, B, ws, ws_bytes))) return rc;
    if (!cond || !x_t || !eps_out) { set_error("null buffer"); return SDDM_E_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const float* temb;
    int stride;
    if (noise_level) {
        float* rows = sect(p, ws, p->off_temb_rows, B);
        TembP tp{noise_level, p->d_f32 + p->off_freq, p->d_f32 + p->off_w1, p->d_f32 + p->off_b1, p->d_f32 + p->off_w2,
                 p->d_f32 + p->off_b2, p->d_f32 + p->off_wn, p->d_f32 + p->off_bn, rows, B, p->cfg.inner_channel, p->E};
        if ((rc = launch_temb(tp, st))) return rc;
        temb = rows;
        stride = p->E;
    } else {
        if (t < 0 || t > p->cfg.n_timestep) { set_error("t=%d out of range", t); return SDDM_E_INVALID; }
        temb = p->d_temb_table + (size_t)t * p->E;
        stride = 0;
    }
    PostP pp{};
    pp.frames = sect(p, ws, p->off_frames, B);
    pp.eps_out = eps_out;
    pp.do_update = 0;
    pp.B = B; pp.L = p->cfg.num_samples; pp.F = p->cfg.segment_len; pp.hop = p->cfg.segment_stride; pp.n_frames = p->H;
    pp.t = 1; pp.T = p->cfg.n_timestep;
    float k8[8] = {0};
    if ((rc = run_unet(p, cond, x_t, temb, stride, B, ws, st, &pp, k8, true))) return rc;
    if (p->post_fused) return SDDM_OK;   // the final Block's kernel wrote eps_hat (and the frames, for sddm_debug_fetch)
    return launch_post_coef(pp, k8, st);
}

static int x_T_impl(sddm_plan* p, int variant, const float* cond, const float* z, uint64_t seed, int64_t row0, float* x_out, int B,
                    void* stream, const unsigned long long* seed_dev);

int sddm_x_T(sddm_plan* p, int variant, const float* cond, const float* z, uint64_t seed, int64_t row0, float* x_out, int B,
             void* stream) {
    return x_T_impl(p, variant, cond, z, seed, row0, x_out, B, stream, nullptr);
}

static int x_T_impl(sddm_plan* p, int variant, const float* cond, const float* z, uint64_t seed, int64_t row0, float* x_out, int B,
                    void* stream, const unsigned long long* seed_dev) {
    int rc = check_ready(p);
    if (rc) return rc;
    if ((rc = check_variant(variant))) return rc;
    if (!x_out || B <= 0 || (!cond && variant != SDDM_VAR_ORIGINAL && variant != SDDM_VAR_SR3)) { set_error("null buffer / bad batch"); return SDDM_E_INVALID; }
    const int T = p->cfg.n_timestep;
    volatile float a = p->sch[2][T];
    volatile float sq = a * a;
    volatile float om = 1.0f - sq;
    float b = sqrtf(om);                              // diffusion.py:297
    if (variant == SDDM_VAR_CONDITIONAL) b = p->sch[7][T];   // sqrt_delta[T], diffusion.py:317
    return launch_x_T_coef(variant, a, b, cond, z, seed, row0, x_out, B, p->cfg.num_samples, (cudaStream_t)stream, seed_dev);
}

int sddm_p_step(sddm_plan* p, int variant, float* x_t, const float* eps, const float* cond, const float* z, uint64_t seed,
                int64_t row0, int t, int B, void* stream) {
    int rc = check_ready(p);
    if (rc) return rc;
    if ((rc = check_variant(variant))) return rc;
    if (!x_t || !eps || B <= 0) { set_error("null buffer / bad batch"); return SDDM_E_INVALID; }
    if (t < 1 || t > p->cfg.n_timestep) { set_error("t=%d out of range [1, %d]", t, p->cfg.n_timestep); return SDDM_E_INVALID; }
    if ((variant == SDDM_VAR_SUPPORTIVE || variant == SDDM_VAR_CONDITIONAL) && !cond) { set_error("variant needs the condition"); return SDDM_E_INVALID; }
    PostP pp{};
    pp.eps_in = eps;
    pp.x_in = x_t; pp.x_out = x_t;
    pp.cond = cond; pp.z = z; pp.seed = seed; pp.row0 = row0;
    pp.variant = variant; pp.t = t; pp.T = p->cfg.n_timestep; pp.do_update = 1;
    pp.B = B; pp.L = p->cfg.num_samples; pp.F = p->cfg.segment_len; pp.hop = p->cfg.segment_stride; pp.n_frames = p->H;
    float k8[8];
    step_coefs(p, variant, t, k8);
    return launch_post_coef(pp, k8, (cudaStream_t)stream);
}

int sddm_x_T_raw(int variant, float a, float b, const float* cond, const float* z, uint64_t seed, int64_t row0, float* x_out,
                 int B, int L, void* stream) {
    int rc = check_variant(variant);
    if (rc) return rc;
    if (!x_out || B <= 0 || L <= 0 || L % 4) { set_error("bad buffer / batch / length (L must be a multiple of 4)"); return SDDM_E_INVALID; }
    if (!cond && variant != SDDM_VAR_ORIGINAL && variant != SDDM_VAR_SR3) { set_error("variant needs the condition"); return SDDM_E_INVALID; }
    return launch_x_T_coef(variant, a, b, cond, z, seed, row0, x_out, B, L, (cudaStream_t)stream);
}

int sddm_p_step_raw(int variant, const float* k8, float* x_t, const float* eps, const float* cond, const float* z, uint64_t seed,
                    int64_t row0, int t, int T, int B, int L, void* stream) {
    int rc = check_variant(variant);
    if (rc) return rc;
    if (!k8 || !x_t || !eps || B <= 0 || L <= 0 || L % 4) { set_error("bad buffer / batch / length (L must be a multiple of 4)"); return SDDM_E_INVALID; }
    if (t < 1 || t > T) { set_error("t=%d out of range [1, %d]", t, T); return SDDM_E_INVALID; }
    if ((variant == SDDM_VAR_SUPPORTIVE || variant == SDDM_VAR_CONDITIONAL) && !cond) { set_error("variant needs the condition"); return SDDM_E_INVALID; }
    PostP pp{};
    pp.eps_in = eps;
    pp.x_in = x_t; pp.x_out = x_t;
    pp.cond = cond; pp.z = z; pp.seed = seed; pp.row0 = row0;
    pp.variant = variant; pp.t = t; pp.T = T; pp.do_update = 1;
    pp.B = B; pp.L = L; pp.F = 4; pp.hop = 4; pp.n_frames = 0;
    return launch_post_coef(pp, k8, (cudaStream_t)stream);
}

int sddm_q_sample_raw(int mode, const float* coef, const float* x0, const float* y, const float* noise, uint64_t seed, int64_t row0,
                      float* x_t, float* combined_noise, float* noise_out, int B, int L, void* stream) {
    if (mode != 0 && mode != 1) { set_error("unknown q_transition mode %d", mode); return SDDM_E_INVALID; }
    if (!coef || !x0 || !x_t || B <= 0 || L <= 0 || L % 4) { set_error("bad buffer / batch / length (L must be a multiple of 4)"); return SDDM_E_INVALID; }
    if (mode == 1 && (!y || !combined_noise)) { set_error("the conditional q_transition needs the condition and a combined-noise output"); return SDDM_E_INVALID; }
    return launch_q_sample(mode, coef, x0, y, noise, seed, row0, x_t, combined_noise, noise_out, B, L, (cudaStream_t)stream);
}

static int sample_impl(sddm_plan* p, int variant, const float* cond, const float* noises, uint64_t seed, int64_t row0, float* out,
                       float* eps_trace, float* x_trace, int B, void* ws, size_t ws_bytes, void* stream, const unsigned long long* seed_dev);

int sddm_sample(sddm_plan* p, int variant, const float* cond, const float* noises, uint64_t seed, int64_t row0, float* out,
                float* eps_trace, float* x_trace, int B, void* ws, size_t ws_bytes, void* stream) {
    return sample_impl(p, variant, cond, noises, seed, row0, out, eps_trace, x_trace, B, ws, ws_bytes, stream, nullptr);
}

static int sample_impl(sddm_plan* p, int variant, const float* cond, const float* noises, uint64_t seed, int64_t row0, float* out,
                       float* eps_trace, float* x_trace, int B, void* ws, size_t ws_bytes, void* stream, const unsigned long long* seed_dev) {
    int rc = check_ready(p);
    if (rc) return rc;
    if ((rc = check_variant(variant))) return rc;
    if ((rc = check_ws(p, B, ws, ws_bytes))) return rc;
    if (!cond || !out) { set_error("null buffer"); return SDDM_E_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const int T = p->cfg.n_timestep, L = p->cfg.num_samples;
    const size_t BL = (size_t)B * L;
    float* x = sect(p, ws, p->off_x, B);
    if ((rc = x_T_impl(p, variant, cond, noises, seed, row0, x, B, stream, seed_dev))) return rc;
    for (int t = T; t >= 1; --t) {
        PostP pp{};
        pp.frames = sect(p, ws, p->off_frames, B);
        pp.eps_out = eps_trace ? eps_trace + (size_t)(T - t) * BL : nullptr;
        pp.x_in = x;
        pp.x_out = (t == 1) ? out : x;
        pp.x_trace = x_trace ? x_trace + (size_t)(T - t) * BL : nullptr;
        pp.cond = cond;
        pp.z = (noises && t > 1) ? noises + (size_t)(T + 1 - t) * BL : nullptr;
        pp.seed = seed; pp.row0 = row0; pp.seed_dev = seed_dev;
        pp.variant = variant; pp.t = t; pp.T = T; pp.do_update = 1;
        pp.B = B; pp.L = L; pp.F = p->cfg.segment_len; pp.hop = p->cfg.segment_stride; pp.n_frames = p->H;
        float k8[8];
        step_coefs(p, variant, t, k8);
        // one forward; with the row kernels its last launch also does the overlap-add + posterior update (frames / eps_hat stay on chip)
        if ((rc = run_unet(p, cond, x, p->d_temb_table + (size_t)t * p->E, 0, B, ws, st, &pp, k8, false))) return rc;
        if (p->post_fused) continue;
        if ((rc = prof_mark(p, (int)p->ops.size(), true, st))) return rc;
        if ((rc = launch_post_coef(pp, k8, st))) return rc;
        if ((rc = prof_mark(p, (int)p->ops.size(), false, st))) return rc;
    }
    return SDDM_OK;
}

int sddm_enhance_host(sddm_plan* p, int variant, const float* cond_host, float* out_host, int B, uint64_t seed, int64_t row0,
                      int max_rows) {
    int rc = check_ready(p);
    if (rc) return rc;
    if (!cond_host || !out_host || B <= 0) { set_error("null buffer / bad batch"); return SDDM_E_INVALID; }
    if (max_rows <= 0) max_rows = 64;
    const int R = B < max_rows ? B : max_rows;
    const size_t L = (size_t)p->cfg.num_samples;
    if (!p->own_stream) SDDM_CUDA_TRY(cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking));
    if (p->arena_rows < R) {
        cudaFree(p->d_cond); cudaFree(p->d_out); cudaFree(p->d_ws);
        p->d_cond = p->d_out = nullptr; p->d_ws = nullptr; p->arena_rows = 0;
        for (auto& kv : p->graphs) cudaGraphExecDestroy(kv.second);   // captured on the old buffers
        p->graphs.clear();
        p->graph_calls.clear();
        SDDM_CUDA_TRY(cudaMalloc(&p->d_cond, R * L * sizeof(float)));
        SDDM_CUDA_TRY(cudaMalloc(&p->d_out, R * L * sizeof(float)));
        SDDM_CUDA_TRY(cudaMalloc(&p->d_ws, sddm_workspace_bytes(p, R)));
        p->arena_rows = R;
    }
    if (!p->d_seed) SDDM_CUDA_TRY(cudaMalloc(&p->d_seed, 2 * sizeof(unsigned long long)));
    static int no_graph = -1;   // SDDM_NO_GRAPH=1: always enqueue the ~4400 launches of a sampling run one by one
    if (no_graph < 0) { const char* ev = getenv("SDDM_NO_GRAPH"); no_graph = (ev && ev[0] == '1') ? 1 : 0; }
    for (int r0 = 0; r0 < B; r0 += R) {
        const int nb = (B - r0) < R ? (B - r0) : R;
        SDDM_CUDA_TRY(cudaMemcpyAsync(p->d_cond, cond_host + (size_t)r0 * L, nb * L * sizeof(float), cudaMemcpyHostToDevice, p->own_stream));
        // The whole sampling run of a sub-batch (x_T + T x 44 launches) on the plan's own buffers is captured ONCE per (rows, variant)
        // into a CUDA graph and replayed afterwards: one launch instead of thousands (what a single-utterance caller is bound by);
        // the Philox seed / first global row are read from device memory so that the same graph serves every call.
        const std::pair<int, int> key(nb, variant);
        // Small sub-batches only: measured on B200, replay beats stream launches at 2 rows (54.4 vs 57.2 ms per 2 s clip) but loses at 64
        // rows (200.6 vs 184.7 ms), where the stream path's programmatic dependent launches overlap each kernel's prologue with its
        // predecessor's tail and the graph's kernel-to-kernel edges do not.
        const bool use_graph = !no_graph && nb <= 8 && !p->prof.on && p->graph_calls[key]++ >= 1;   // the first call of a shape runs eagerly (warm-up)
        if (use_graph) {
            const unsigned long long sr[2] = {(unsigned long long)seed, (unsigned long long)(row0 + r0)};
            SDDM_CUDA_TRY(cudaMemcpyAsync(p->d_seed, sr, sizeof(sr), cudaMemcpyHostToDevice, p->own_stream));
            auto it = p->graphs.find(key);
            if (it == p->graphs.end()) {
                cudaGraph_t g = nullptr;
                SDDM_CUDA_TRY(cudaStreamBeginCapture(p->own_stream, cudaStreamCaptureModeThreadLocal));
                rc = sample_impl(p, variant, p->d_cond, nullptr, 0, 0, p->d_out, nullptr, nullptr, nb, p->d_ws,
                                 sddm_workspace_bytes(p, p->arena_rows), p->own_stream, p->d_seed);
                const cudaError_t ce = cudaStreamEndCapture(p->own_stream, &g);
                if (rc) { if (g) cudaGraphDestroy(g); return rc; }
                if (ce != cudaSuccess) { set_error("graph capture failed: %s", cudaGetErrorString(ce)); cudaGetLastError(); return SDDM_E_CUDA; }
                cudaGraphExec_t ge = nullptr;
                const cudaError_t ie = cudaGraphInstantiate(&ge, g, 0);
                cudaGraphDestroy(g);
                if (ie != cudaSuccess) { set_error("graph instantiation failed: %s", cudaGetErrorString(ie)); cudaGetLastError(); return SDDM_E_CUDA; }
                it = p->graphs.emplace(key, ge).first;
            }
            SDDM_CUDA_TRY(cudaGraphLaunch(it->second, p->own_stream));
            count_launch(1 + p->cfg.n_timestep * p->launches_per_eps);   // kernels inside the replayed graph
        } else {
            rc = sample_impl(p, variant, p->d_cond, nullptr, seed, row0 + r0, p->d_out, nullptr, nullptr, nb, p->d_ws,
                             sddm_workspace_bytes(p, p->arena_rows), p->own_stream, nullptr);
            if (rc) return rc;
        }
        SDDM_CUDA_TRY(cudaMemcpyAsync(out_host + (size_t)r0 * L, p->d_out, nb * L * sizeof(float), cudaMemcpyDeviceToHost, p->own_stream));
    }
    SDDM_CUDA_TRY(cudaStreamSynchronize(p->own_stream));
    return SDDM_OK;
}

int sddm_frames(const float* sig, float* frames, int B, int n, int F, int hop, void* stream) {
    if (!sig || !frames || B <= 0 || F <= 0 || hop <= 0 || n < F) { set_error("bad argument"); return SDDM_E_INVALID; }
    if ((n - F) % hop) { set_error("(n_samples - F) %% stride must be 0"); return SDDM_E_INVALID; }
    return launch_frames(sig, frames, B, n, F, hop, (cudaStream_t)stream);
}

int sddm_overlap_add(const float* frames, float* sig, int B, int n, int F, int hop, void* stream) {
    if (!sig || !frames || B <= 0 || F <= 0 || hop <= 0 || n < F) { set_error("bad argument"); return SDDM_E_INVALID; }
    if ((n - F) % hop) { set_error("(n_samples - F) %% stride must be 0"); return SDDM_E_INVALID; }
    return launch_overlap_add(frames, sig, B, n, F, hop, (cudaStream_t)stream);
}

static int check_rows_args(const void* a, const void* b, const int64_t* so, const int64_t* ro, int n_utt, int T, int64_t lo, int64_t hi) {
    if (!a || !b || !so || !ro || n_utt <= 0 || T <= 0 || lo < 0 || hi < lo) { set_error("bad argument"); return SDDM_E_INVALID; }
    if (hi - lo > 0x7fffffffLL) { set_error("too many rows in one call"); return SDDM_E_INVALID; }
    return SDDM_OK;
}

int sddm_chunk_rows(const float* flat, const int64_t* sample_off, const int64_t* row_off, int n_utt, int T, int64_t row_lo, int64_t row_hi,
                    float* rows, void* stream) {
    int rc = check_rows_args(flat, rows, sample_off, row_off, n_utt, T, row_lo, row_hi);
    if (rc) return rc;
    return launch_chunk_rows(flat, sample_off, row_off, n_utt, T, row_lo, row_hi, rows, (cudaStream_t)stream);
}

int sddm_regroup_rows(const float* rows, const int64_t* sample_off, const int64_t* row_off, int n_utt, int T, int64_t row_lo, int64_t row_hi,
                      float* flat_out, void* stream) {
    int rc = check_rows_args(rows, flat_out, sample_off, row_off, n_utt, T, row_lo, row_hi);
    if (rc) return rc;
    return launch_regroup_rows(rows, sample_off, row_off, n_utt, T, row_lo, row_hi, flat_out, (cudaStream_t)stream);
}

int sddm_plan_num_ops(const sddm_plan* p) { return (p && p->finalized) ? (int)p->ops.size() + 1 : 0; }

int sddm_profile_enable(sddm_plan* p, int on) {
    int rc = check_ready(p);
    if (rc) return rc;
    if ((rc = prof_flush(p))) return rc;
    p->prof.on = on != 0;
    p->prof.ms.assign(p->ops.size() + 1, 0.0);
    p->prof.cnt.assign(p->ops.size() + 1, 0);
    return SDDM_OK;
}

int sddm_profile_read(sddm_plan* p, int op, double* total_ms, int64_t* launches, double* flops_per_row, double* bytes_per_row,
                      int* uses_tensor_cores, char* label, int label_cap) {
    int rc = check_ready(p);
    if (rc) return rc;
    if (op < 0 || op > (int)p->ops.size()) { set_error("op index %d out of range", op); return SDDM_E_INVALID; }
    if ((rc = prof_flush(p))) return rc;
    const bool is_post = op == (int)p->ops.size();
    const double L = p->cfg.num_samples;
    if (total_ms) *total_ms = p->prof.ms.empty() ? 0.0 : p->prof.ms[op];
    if (launches) *launches = p->prof.cnt.empty() ? 0 : (int64_t)p->prof.cnt[op];
    if (flops_per_row) *flops_per_row = is_post ? 8.0 * L : p->ops[op].flops;
    // ola + posterior with in-kernel Philox: read frames (2 per sample) + x_t, write x_{t-1}
    if (bytes_per_row) *bytes_per_row = is_post ? 4.0 * (2.0 * L + 2.0 * L) : p->ops[op].bytes;
    if (uses_tensor_cores) *uses_tensor_cores = is_post ? 0 : (p->ops[op].use_row ? 2 : (p->ops[op].use_tc ? 1 : 0));   // 2 = conv_row.cu, 1 = conv_tc.cu
    if (label && label_cap > 0) snprintf(label, (size_t)label_cap, "%s", is_post ? "ola_posterior" : p->ops[op].label.c_str());
    return SDDM_OK;
}

int sddm_debug_fetch(sddm_plan* p, const char* node, void* ws, int B, float* out, int64_t* chw, void* stream) {
    int rc = check_ready(p);
    if (rc) return rc;
    if (!node || !ws || B <= 0) { set_error("bad argument"); return SDDM_E_INVALID; }
    for (const TensorInfo& t : p->tensors)
        if (t.name == node) {
            const size_t n = (size_t)t.C * t.H * t.W;
            if (chw) { chw[0] = t.C; chw[1] = t.H; chw[2] = t.W; }
            if (out && p->cfg.precision == SDDM_PREC_BF16_ACT) return launch_bf16_to_f32(sect(p, ws, t.data_off, B), out, n * B, (cudaStream_t)stream);
            if (out) SDDM_CUDA_TRY(cudaMemcpyAsync(out, sect(p, ws, t.data_off, B), n * B * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
            return SDDM_OK;
        }
    if (std::string(node) == "frames") {
        if (chw) { chw[0] = 1; chw[1] = p->H; chw[2] = p->W; }
        if (out) SDDM_CUDA_TRY(cudaMemcpyAsync(out, sect(p, ws, p->off_frames, B), (size_t)p->H * p->W * B * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        return SDDM_OK;
    }
    set_error("unknown node '%s'", node);
    return SDDM_E_INVALID;
}

}  // extern "C"
