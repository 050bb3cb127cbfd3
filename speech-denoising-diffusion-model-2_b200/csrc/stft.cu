// STFT feature front-end (cfg 5): magnitude spectrogram / mel spectrogram + log10 / clamp compression.
//
// reference: prepare_spectrogram.py:20-55 = torchaudio TT.Spectrogram(n_fft=1024, hop, window_fn, power=1, normalized=True)
// -> torch.stft(center=True, pad_mode="reflect", onesided) / sqrt(sum w^2) -> abs  [-> MelScale (HTK triangular filterbank)]
// -> clamp((log10(S) - 1 + 5) / 5, 0, 1).   Output layout [B][n_out][frames] as the reference's .npy files.
//
// HBM bound (reads 4 L bytes, writes ~8 L bytes per utterance of L samples, ~200 L flop).  One CTA = 32 consecutive frames of
// one utterance, one warp per frame (4 frames per warp):
//   * the 1024 windowed (reflect-padded) samples are packed into 512 complex points, 16 per lane: z[32 n1 + lane];
//   * 512-point FFT as a four-step 16 x 32 decomposition: a 16-point radix-2 DIF in registers (over n1), the twiddle
//     W512^(lane k1), then a 32-point radix-2 DIF ACROSS LANES with warp shuffles (over lane);
//   * the spectrum is staged in shared memory (padded, conflict free) for the real-FFT post-processing
//     X[k] = (Z[k] + conj Z[512-k]) / 2 - i/2 W1024^k (Z[k] - conj Z[512-k]),  k = 0..512;
//   * magnitudes of the CTA's 32 frames are staged as a [513][32] tile so that global stores (and the mel filterbank
//     contraction) run along the contiguous frame dimension.
#include "kernels.cuh"
#include "../../include/sddm_b200.h"

namespace sddm {
namespace {

constexpr int NFFT = 1024, NBIN = NFFT / 2 + 1, FPB = 32;   // frames per CTA
constexpr int ZPAD = 17 * 32;                              // padded 512-point spectrum: index (k & 15) + 17 * (k >> 4)
constexpr int TILE_LD = FPB + 1;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ int zidx(int k) { return (k & 15) + 17 * (k >> 4); }

struct StftP {
    const float* wav;      // [B][L]
    const float* window;   // [1024]
    const float* melfb;    // [513][n_mels] or nullptr
    const int* mel_lo;     // [n_mels] first bin with a non-zero weight (nullptr: 0)
    const int* mel_hi;     // [n_mels] one past the last non-zero bin (nullptr: 513)
    float* out;            // [B][n_out][frames]
    int B, L, hop, frames, n_mels, log_clamp;
    float inv_norm;        // 1 / sqrt(sum w^2)
};

// W16^e = exp(-2 pi i e / 16), e = 0..7: the twiddles of the in-register 16-point DIF (compile-time constants)
__device__ __forceinline__ float2 w16(int e) {
    constexpr float c[8] = {1.0f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f, 0.0f, -0.38268343236508977f,
                            -0.70710678118654752f, -0.92387953251128674f};
    constexpr float sn[8] = {0.0f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f, -1.0f, -0.92387953251128674f,
                             -0.70710678118654752f, -0.38268343236508977f};
    return make_float2(c[e], sn[e]);
}

__global__ void __launch_bounds__(256, 2) stft_kernel(StftP p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* w1024 = reinterpret_cast<float2*>(smem_raw);                // [513]  exp(-2 pi i k / 1024)   (real-FFT post-processing)
    float* win = reinterpret_cast<float*>(w1024 + 520);                 // [1024] window
    float2* zbuf = reinterpret_cast<float2*>(win + NFFT);               // [8 warps][ZPAD]
    float* tile = reinterpret_cast<float*>(zbuf + 8 * ZPAD);            // [513][TILE_LD] magnitudes of the CTA's frames
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y, m0 = blockIdx.x * FPB;
    for (int t = tid; t < NBIN; t += 256) {
        float sv, cv;
        sincospif(-2.0f * (float)t / 1024.0f, &sv, &cv);
        w1024[t] = make_float2(cv, sv);
    }
    for (int t = tid; t < NFFT; t += 256) win[t] = __ldg(p.window + t);
    // per-lane twiddles, computed once: W512^(lane k1(q)) between the two FFT stages, and for stage s (half = 16 >> s) of the
    // cross-lane DIF the factor applied to the UPPER lane of a butterfly, W_(2 half)^(lane mod half) (1 for the lower lane)
    float2 tw1[16], tws[4];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int k1 = ((q & 1) << 3) | ((q & 2) << 1) | ((q & 4) >> 1) | ((q & 8) >> 3);
        float sv, cv;
        sincospif(-2.0f * (float)((lane * k1) & 511) / 512.0f, &sv, &cv);
        tw1[q] = make_float2(cv, sv);
    }
#pragma unroll
    for (int st = 0; st < 4; ++st) {
        const int half = 16 >> st;
        float sv, cv;
        sincospif(-2.0f * (float)((lane & (half - 1)) * (16 / half) * 16) / 512.0f, &sv, &cv);
        tws[st] = (lane & half) ? make_float2(cv, sv) : make_float2(1.0f, 0.0f);
    }
    __syncthreads();
    const float* x = p.wav + (int64_t)b * p.L;
    float2* zb = zbuf + warp * ZPAD;
    const int k2 = (int)(__brev((unsigned)lane) >> 27);
    const int zlo = (lane & 15) + 17 * (lane >> 4);   // zidx(lane + 32 r) = zlo + 34 r
    for (int f = 0; f < FPB / 8; ++f) {
        const int ml = warp * (FPB / 8) + f, m = m0 + ml;
        if (m < p.frames) {   // warp-uniform
            // ---- load + window + pack: register n1 holds z[32 n1 + lane] = (x[64 n1 + 2 lane], x[64 n1 + 2 lane + 1])
            float2 v[16];
            const int i0 = p.hop * m - NFFT / 2;
            if (i0 >= 0 && i0 + NFFT <= p.L && ((i0 | (int)((uintptr_t)x >> 2)) & 1) == 0) {   // interior frame, 8-byte aligned pairs
                const float2* xp = reinterpret_cast<const float2*>(x + i0) + lane;
                const float2* wp = reinterpret_cast<const float2*>(win) + lane;
#pragma unroll
                for (int n1 = 0; n1 < 16; ++n1) {
                    const float2 xv = __ldg(xp + 32 * n1), wv = wp[32 * n1];
                    v[n1] = make_float2(xv.x * wv.x, xv.y * wv.y);
                }
            } else {
#pragma unroll
                for (int n1 = 0; n1 < 16; ++n1) {
                    const int s0 = 64 * n1 + 2 * lane;
                    float xv[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        int i = i0 + s0 + q;   // centre = True, reflect padding
                        if (i < 0) i = -i;
                        if (i >= p.L) i = 2 * (p.L - 1) - i;
                        xv[q] = __ldg(x + i) * win[s0 + q];
                    }
                    v[n1] = make_float2(xv[0], xv[1]);
                }
            }
            // ---- 16-point radix-2 DIF over the registers: afterwards register q holds Y[k1 = bitrev4(q)]
#pragma unroll
            for (int half = 8; half >= 1; half >>= 1) {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    if (q & half) continue;
                    const float2 a = v[q], c = v[q + half];
                    v[q] = make_float2(a.x + c.x, a.y + c.y);
                    const float2 d = make_float2(a.x - c.x, a.y - c.y);
                    const int e = (q & (half - 1)) * (8 / half);      // W16^e
                    v[q + half] = e == 0 ? d : (e == 4 ? make_float2(d.y, -d.x) : cmul(d, w16(e)));
                }
            }
            // ---- twiddle W512^(lane * k1), then the 32-point radix-2 DIF ACROSS LANES: lane ends with k2 = bitrev5(lane)
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = cmul(v[q], tw1[q]);
#pragma unroll
            for (int st = 0; st < 5; ++st) {
                const int half = 16 >> st;
                const float sgn = (lane & half) ? -1.0f : 1.0f;      // lower lane: v + o; upper lane: (o - v) * twiddle
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const float ox = __shfl_xor_sync(0xffffffffu, v[q].x, half), oy = __shfl_xor_sync(0xffffffffu, v[q].y, half);
                    const float2 a = make_float2(fmaf(v[q].x, sgn, ox), fmaf(v[q].y, sgn, oy));
                    v[q] = st < 4 ? cmul(a, tws[st < 4 ? st : 0]) : a;   // the last stage's twiddle is 1
                }
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int k1 = ((q & 1) << 3) | ((q & 2) << 1) | ((q & 4) >> 1) | ((q & 8) >> 3);
                zb[zidx(k1 + 16 * k2)] = v[q];
            }
            __syncwarp();
            // ---- real-FFT post-processing + magnitude: bins k = lane + 32 r
#pragma unroll 4
            for (int r = 0; r < 17; ++r) {
                const int k = lane + 32 * r;
                if (k >= NBIN) break;
                const float2 zk = zb[k < 512 ? zlo + 34 * r : 0], zr = zb[zidx((512 - k) & 511)];
                const float2 e = make_float2(0.5f * (zk.x + zr.x), 0.5f * (zk.y - zr.y));       // (Z[k] + conj Z[512-k]) / 2
                const float2 d = make_float2(zk.x - zr.x, zk.y + zr.y);                         //  Z[k] - conj Z[512-k]
                const float2 wd = cmul(w1024[k], d);
                const float re = e.x + 0.5f * wd.y, im = e.y - 0.5f * wd.x;                     // e - i/2 * wd
                tile[k * TILE_LD + ml] = sqrtf(re * re + im * im) * p.inv_norm;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    const int nvalid = (p.frames - m0) < FPB ? (p.frames - m0) : FPB;
    if (p.melfb == nullptr) {
        float* o = p.out + (int64_t)b * NBIN * p.frames;
        const int ml = tid & 31;
        if (ml < nvalid) {
            for (int k = tid >> 5; k < NBIN; k += 8) {   // a warp writes one 128-byte row segment of bin k
                float sv = tile[k * TILE_LD + ml];
                if (p.log_clamp) sv = fminf(fmaxf(fmaf(__log2f(sv), 0.0602059991327962f, 0.8f), 0.0f), 1.0f);   // (log10(s) - 1 + 5) / 5
                o[(int64_t)k * p.frames + m0 + ml] = sv;
            }
        }
    } else {   // MelScale: mel[j][m] = sum_k S[k][m] * fb[k][j]
        float* o = p.out + (int64_t)b * p.n_mels * p.frames;
        for (int i = tid; i < p.n_mels * FPB; i += 256) {
            const int j = i / FPB, ml = i - j * FPB;
            if (ml < nvalid) {
                const int klo = p.mel_lo ? __ldg(p.mel_lo + j) : 0, khi = p.mel_hi ? __ldg(p.mel_hi + j) : NBIN;
                // four independent partial sums: the filterbank loads (L1 hits, ~40 cycles each) would otherwise serialise the loop
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                int k = klo;
                const float* fb = p.melfb + j;
                for (; k + 3 < khi; k += 4) {
                    a0 = fmaf(tile[k * TILE_LD + ml], __ldg(fb + (int64_t)k * p.n_mels), a0);
                    a1 = fmaf(tile[(k + 1) * TILE_LD + ml], __ldg(fb + (int64_t)(k + 1) * p.n_mels), a1);
                    a2 = fmaf(tile[(k + 2) * TILE_LD + ml], __ldg(fb + (int64_t)(k + 2) * p.n_mels), a2);
                    a3 = fmaf(tile[(k + 3) * TILE_LD + ml], __ldg(fb + (int64_t)(k + 3) * p.n_mels), a3);
                }
                for (; k < khi; ++k) a0 = fmaf(tile[k * TILE_LD + ml], __ldg(fb + (int64_t)k * p.n_mels), a0);
                float acc = (a0 + a1) + (a2 + a3);
                if (p.log_clamp) acc = fminf(fmaxf((log10f(acc) - 1.0f + 5.0f) / 5.0f, 0.0f), 1.0f);
                o[(int64_t)j * p.frames + m0 + ml] = acc;
            }
        }
    }
}

}  // namespace
}  // namespace sddm

extern "C" SDDM_API int sddm_stft_features(const float* wav, int B, int L, int n_fft, int hop, const float* window, float inv_norm,
                                           const float* mel_fb, const int32_t* mel_lo, const int32_t* mel_hi, int n_mels, int log_clamp, float* out,
                                           void* stream) {
    using namespace sddm;
    if (!wav || !window || !out || B <= 0) { set_error("stft: null buffer / bad batch"); return SDDM_E_INVALID; }
    if (n_fft != NFFT) { set_error("stft: n_fft must be %d (config_diffwave.json window_length), got %d", NFFT, n_fft); return SDDM_E_INVALID; }
    if (hop <= 0 || L <= n_fft / 2) { set_error("stft: hop must be positive and L > n_fft/2 (reflect padding), got hop=%d L=%d", hop, L); return SDDM_E_INVALID; }
    if (mel_fb && n_mels <= 0) { set_error("stft: n_mels must be positive with a filterbank"); return SDDM_E_INVALID; }
    StftP p{wav, window, mel_fb, mel_fb ? mel_lo : nullptr, mel_fb ? mel_hi : nullptr, out, B, L, hop, 1 + L / hop, mel_fb ? n_mels : 0, log_clamp, inv_norm};
    const size_t smem = (520 + 8 * ZPAD) * sizeof(float2) + (size_t)NFFT * sizeof(float) + (size_t)NBIN * TILE_LD * sizeof(float);
    SDDM_SET_MAX_SMEM(stft_kernel, smem);
    dim3 grid((p.frames + FPB - 1) / FPB, B);
    stft_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(p);
    SDDM_LAUNCH_CHECK();
    return SDDM_OK;
}
