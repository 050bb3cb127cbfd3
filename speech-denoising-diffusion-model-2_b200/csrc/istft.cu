// Overlap-add inverse STFT (named by BASELINE.json's north_star next to the STFT front-end; the reference itself never inverts a
// spectrogram - SURVEY.md section 0.3 - so the semantics are those of torch.istft, the inverse of the torch.stft call inside
// torchaudio's Spectrogram that prepare_spectrogram.py:20-35 uses):
//
//   y_m[n]  = irfft(X[:, m])[n] * w[n]                       n = 0..1023           (istft_frames_kernel)
//   out[t]  = sum_m y_m[t + 512 - m hop] / sum_m w[t + 512 - m hop]^2              (istft_ola_kernel; centre = True trims n_fft / 2)
//
// Kernel 1: one CTA (256 threads) per frame.  The one-sided spectrum (513 bins) is extended to its Hermitian 1024-point form in
// shared memory at the bit-reversed positions, then a radix-2 decimation-in-time inverse FFT runs its 10 stages in shared memory
// (two butterflies per thread and stage, twiddles from a 512-entry table built once per CTA); the windowed frame goes to a
// workspace.  Kernel 2 gathers the (at most n_fft / hop) frames that cover an output sample in ascending frame order -
// deterministic, no atomics - and divides by the window envelope.
#include "kernels.cuh"
#include "../../include/sddm_b200.h"

namespace sddm {
namespace {

constexpr int NF = 1024, NB = NF / 2 + 1;

__device__ __forceinline__ int bitrev10(int v) { return (int)(__brev((unsigned)v) >> 22); }

__global__ void __launch_bounds__(256) istft_frames_kernel(const float2* __restrict__ spec, const float* __restrict__ window,
                                                           float* __restrict__ frames_out, int frames) {
    __shared__ float2 z[NF];
    __shared__ float2 tw[NF / 2];     // exp(+2 pi i k / 1024)
    const int tid = threadIdx.x, m = blockIdx.x, b = blockIdx.y;
    for (int k = tid; k < NF / 2; k += 256) {
        float sv, cv;
        sincospif(2.0f * (float)k / (float)NF, &sv, &cv);
        tw[k] = make_float2(cv, sv);
    }
    // Hermitian extension, stored at the bit-reversed position (input order of a decimation-in-time FFT); the imaginary parts of
    // the DC and Nyquist bins are ignored, as a real inverse transform does
    const float2* col = spec + (int64_t)b * NB * frames + m;          // [513][frames] complex, frame m
    for (int k = tid; k < NF; k += 256) {
        float2 v;
        if (k <= NF / 2) {
            v = __ldg(col + (int64_t)k * frames);
            if (k == 0 || k == NF / 2) v.y = 0.f;
        } else {
            v = __ldg(col + (int64_t)(NF - k) * frames);
            v.y = -v.y;
        }
        z[bitrev10(k)] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int span = 1; span < NF; span <<= 1) {   // 10 stages; butterfly (i, i + span), twiddle exp(+2 pi i j / (2 span)), j = i % span
        for (int q = tid; q < NF / 2; q += 256) {
            const int j = q & (span - 1), i = ((q - j) << 1) + j;
            const float2 w = tw[j * (NF / 2 / span)];
            const float2 a = z[i], c = z[i + span];
            const float2 t = make_float2(c.x * w.x - c.y * w.y, c.x * w.y + c.y * w.x);
            z[i] = make_float2(a.x + t.x, a.y + t.y);
            z[i + span] = make_float2(a.x - t.x, a.y - t.y);
        }
        __syncthreads();
    }
    float* dst = frames_out + ((int64_t)b * frames + m) * NF;
    const float inv = 1.0f / (float)NF;
    for (int n = tid; n < NF; n += 256) dst[n] = z[n].x * inv * __ldg(window + n);
}

__global__ void __launch_bounds__(256) istft_ola_kernel(const float* __restrict__ fr, const float* __restrict__ window, float* __restrict__ out,
                                                        int frames, int hop, int L) {
    const int b = blockIdx.y;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < L; t += gridDim.x * blockDim.x) {
        const int tt = t + NF / 2;                       // centre = True: the first n_fft / 2 samples are padding
        int m_lo = (tt - (NF - 1) + hop - 1) / hop;      // smallest m with m hop + 1023 >= tt
        if (tt - (NF - 1) < 0) m_lo = 0;
        int m_hi = tt / hop;
        if (m_hi > frames - 1) m_hi = frames - 1;
        float acc = 0.f, env = 0.f;
        for (int m = m_lo; m <= m_hi; ++m) {             // ascending frame order
            const int n = tt - m * hop;
            const float w = __ldg(window + n);
            acc += fr[((int64_t)b * frames + m) * NF + n];
            env = fmaf(w, w, env);
        }
        out[(int64_t)b * L + t] = env > 1e-11f ? acc / env : 0.f;
    }
}

}  // namespace
}  // namespace sddm

extern "C" SDDM_API size_t sddm_istft_workspace_bytes(int B, int frames) {
    if (B <= 0 || frames <= 0) return 0;
    return (size_t)B * frames * sddm::NF * sizeof(float);
}

extern "C" SDDM_API int sddm_istft(const float* spec_ri, int B, int frames, int n_fft, int hop, const float* window, int L, float* wav_out,
                                   void* ws, size_t ws_bytes, void* stream) {
    using namespace sddm;
    if (!spec_ri || !window || !wav_out || !ws || B <= 0 || frames <= 0) { set_error("istft: null buffer / bad batch"); return SDDM_E_INVALID; }
    if (n_fft != NF) { set_error("istft: n_fft must be %d, got %d", NF, n_fft); return SDDM_E_INVALID; }
    if (hop <= 0 || hop > NF) { set_error("istft: hop must be in [1, n_fft]"); return SDDM_E_INVALID; }
    if (L <= 0 || L > (frames - 1) * hop + NF - NF / 2) { set_error("istft: length %d exceeds what %d frames cover", L, frames); return SDDM_E_INVALID; }
    if (ws_bytes < sddm_istft_workspace_bytes(B, frames)) { set_error("istft: workspace too small"); return SDDM_E_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    float* fr = reinterpret_cast<float*>(ws);
    istft_frames_kernel<<<dim3(frames, B), 256, 0, st>>>(reinterpret_cast<const float2*>(spec_ri), window, fr, frames);
    SDDM_LAUNCH_CHECK();
    int gx = (L + 255) / 256;
    if (gx > 148 * 8) gx = 148 * 8;
    istft_ola_kernel<<<dim3(gx, B), 256, 0, st>>>(fr, window, wav_out, frames, hop, L);
    SDDM_LAUNCH_CHECK();
    return SDDM_OK;
}
