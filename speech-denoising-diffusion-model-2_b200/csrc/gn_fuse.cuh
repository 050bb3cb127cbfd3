// GroupNorm finalisation fused into the PRODUCER of the normalised tensor (replaces one gn_finalize_kernel launch per GroupNorm).
//
// Every CTA that completes a tile of sample n publishes its partial statistics, fences, and bumps counter[n]; the CTA that
// observes the last arrival reduces ALL partials of that sample (of every source of the consumer's GroupNorm: a concatenated
// input has two) in fp64, in an order that depends only on (sample, group) - so results stay deterministic and independent of
// the batch size and of which CTA happens to finish last - and writes scale = gamma * rstd, shift = beta - mean * scale.
// The finalising CTA resets the counter, so the buffer is all-zero again when the kernel ends (graph replay safe).
//
// reference: nn.GroupNorm(groups, C, eps=1e-5) inside Block, model/UNetModified2.py:113-124
#pragma once
#include "common.cuh"

namespace sddm {

// executed by `nthreads` (multiple of 32, >= 32) threads of one CTA that all pass the same arguments; tid in [0, nthreads).
// The caller guarantees that all partials of sample n are visible (fence + counter protocol above).
__device__ __forceinline__ void gn_fused_finalize(const GnFuse& f, int n, int tid, int nthreads) {
    const int cpg = f.Ctot / f.groups;
    const int lanes = 4;                                   // threads cooperating on one group
    const int sub = tid & (lanes - 1);
    for (int g = tid / lanes; g < f.groups; g += nthreads / lanes) {
        const int c_lo = g * cpg;
        double sum = 0.0, sq = 0.0;
        int off = 0;
        for (int s = 0; s < f.nsrc; ++s) {
            const int C = f.C[s], np = f.nparts[s];
            const int lo = c_lo > off ? c_lo : off, hi = (c_lo + cpg) < (off + C) ? (c_lo + cpg) : (off + C);
            const int w = hi - lo;                         // channels of this group inside source s
            if (w > 0) {
                // flat (part, channel) walk with independent loads: 16 are in flight per thread (the walk is latency-bound)
                const float2* basep = reinterpret_cast<const float2*>(f.parts[s]) + (int64_t)n * np * C + (lo - off);
                const int total = np * w;
#pragma unroll 1
                for (int i0 = sub; i0 < total; i0 += lanes * 16) {
                    float2 v[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const int i = i0 + u * lanes;
                        const int part = w == 1 ? i : i / w, j = w == 1 ? 0 : i - part * w;
                        v[u] = i < total ? __ldcg(basep + (int64_t)part * C + j) : make_float2(0.f, 0.f);
                    }
#pragma unroll
                    for (int u = 0; u < 16; ++u) { sum += (double)v[u].x; sq += (double)v[u].y; }
                }
            }
            off += C;
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1); sq += __shfl_xor_sync(0xffffffffu, sq, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2); sq += __shfl_xor_sync(0xffffffffu, sq, 2);
        const double cnt = (double)cpg * (double)f.HW;
        const double mean = sum / cnt;
        double var = sq / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        const double rstd = 1.0 / sqrt(var + (double)f.eps);
        for (int j = sub; j < cpg; j += lanes) {
            const int c = c_lo + j;
            const double sc = (double)__ldg(f.gamma + c) * rstd;
            f.scale[(int64_t)n * f.Ctot + c] = (float)sc;
            f.shift[(int64_t)n * f.Ctot + c] = (float)((double)__ldg(f.beta + c) - mean * sc);
        }
    }
}

}  // namespace sddm
