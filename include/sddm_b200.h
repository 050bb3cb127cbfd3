/*
 * sddm_b200.h — C ABI of the B200-native reverse-diffusion speech-enhancement hot path.
 *
 * The reference (yangye1098/Speech-Denoising-Diffusion-Model-2) is pure Python/PyTorch and has no FFI;
 * each entry point below names the reference function(s) (file:line under /root/reference) whose device
 * work it replaces.  Plain pointers and sizes only: no torch types cross this boundary.
 *
 * Conventions
 *   - every function returns 0 on success, a negative SDDM_E_* code otherwise; the message is available
 *     through sddm_last_error() (thread-local).  No exceptions cross the ABI, and there is NO CPU fallback:
 *     unsupported shapes / missing device => error.
 *   - all device pointers are fp32, contiguous, on the current CUDA device; work is enqueued on `stream`
 *     (a cudaStream_t passed as void*), no hidden synchronisation (CUDA-graph capturable) unless stated.
 *   - the caller owns every buffer including the workspace; the plan owns only packed weights + tables.
 *   - a plan is not thread-safe; use one plan per (process, device).
 *   - waveforms are [B, 1, L] (L = num_samples, 16448 for config_unet.json); one row = one chunk.
 */
#ifndef SDDM_B200_H
#define SDDM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(SDDM_BUILD) && defined(__GNUC__)
#define SDDM_API __attribute__((visibility("default")))
#else
#define SDDM_API
#endif

#define SDDM_OK 0
#define SDDM_E_INVALID (-1)     /* bad argument / unsupported configuration */
#define SDDM_E_STATE (-2)       /* call order: weights / schedule missing, plan not finalised */
#define SDDM_E_CUDA (-3)        /* a CUDA runtime call failed */
#define SDDM_E_WORKSPACE (-4)   /* workspace too small */

/* precision modes for the denoiser convolutions */
#define SDDM_PREC_FP32 0        /* CUDA-core fp32 FMA (parity mode: eps_hat error ~1e-6) */
#define SDDM_PREC_BF16 1        /* tcgen05 / TMEM implicit GEMM, bf16 operands, fp32 accumulate, fp32 activations in HBM */
#define SDDM_PREC_BF16_ACT 2    /* same, and the UNet's intermediate activations are stored as bf16 in HBM */
#define SDDM_PREC_BF16X3 3      /* tensor-core high-precision mode (north_star's "TF32 mode", stricter): fp32 activations, every conv
                                 * operand split into a bf16 (hi, lo) pair, a.w = a_hi w_hi + a_hi w_lo + a_lo w_hi on tcgen05 with
                                 * fp32 accumulation (operands carried to 2^-17); eps_hat error ~1e-5 vs the reference */

/* posterior-update variants: SDDM.p_transition argument, model/model.py:17-26 */
#define SDDM_VAR_ORIGINAL 0     /* diffusion.py:177-190, x_T = pure noise               */
#define SDDM_VAR_CONDITION_IN 1 /* diffusion.py:177-190, x_T = get_x_T (diffusion.py:281) */
#define SDDM_VAR_SR3 2          /* diffusion.py:164-175 */
#define SDDM_VAR_SUPPORTIVE 3   /* diffusion.py:192-209 */
#define SDDM_VAR_CONDITIONAL 4  /* diffusion.py:211-222, x_T = get_x_T_conditional (:302) */

typedef struct sddm_plan sddm_plan;

/* Mirrors the kwargs of UNetModified2.__init__ (model/UNetModified2.py:146-160) + GaussianDiffusion.n_timestep. */
typedef struct sddm_config {
    int32_t n_timestep;       /* T (config_unet.json: 100) */
    int32_t num_samples;      /* L, samples per chunk (16448) */
    int32_t segment_len;      /* frame length F (128) */
    int32_t segment_stride;   /* hop (64); (L - F) % hop must be 0 (UNetModified2.py:13) */
    int32_t in_channel;       /* must be 2 */
    int32_t out_channel;      /* must be 1 */
    int32_t inner_channel;    /* 32 */
    int32_t norm_groups;      /* 32 */
    int32_t n_mults;          /* len(channel_mults) <= 8 */
    int32_t channel_mults[8]; /* (1,2,3,4,5) */
    int32_t res_blocks;       /* 1 */
    int32_t precision;        /* SDDM_PREC_* */
    int32_t reserved[4];
} sddm_config;

/* The 12 per-timestep coefficient tables, each of length n_timestep + 1 (host pointers, fp32).
 * = the registered buffers of GaussianDiffusion (model/diffusion.py:87-161). */
typedef struct sddm_schedule {
    const float* betas;
    const float* alphas;
    const float* sqrt_alpha_bar;
    const float* predicted_noise_coeff;
    const float* sigma;
    const float* supportive_gamma;
    const float* supportive_sigma_hat;
    const float* sqrt_delta;
    const float* c_xt;
    const float* c_yt;
    const float* c_epst;
    const float* sqrt_delta_estimated;
} sddm_schedule;

SDDM_API const char* sddm_last_error(void);
SDDM_API int sddm_version(void);
/* number of kernels this library has launched in this process (monotonic). */
SDDM_API uint64_t sddm_launch_count(void);

/* ---- plan lifecycle ------------------------------------------------------------------------------- */
/* replaces: UNetModified2.__init__ (UNetModified2.py:146-235) as far as shapes go. */
SDDM_API int sddm_plan_create(const sddm_config* cfg, sddm_plan** out);
SDDM_API void sddm_plan_destroy(sddm_plan* plan);

/* replaces: model.load_state_dict (infer.py:51).  `name` is the reference state_dict key WITHOUT the
 * 'noise_estimate_model.' prefix (e.g. "downs.1.block1.block.3.weight"); data is fp32, host or device
 * (copied immediately, the pointer is not retained).  Shape is checked against the config. */
SDDM_API int sddm_plan_load_weight(sddm_plan* plan, const char* name, const void* data, const int64_t* shape, int ndim);

/* replaces: GaussianDiffusion.__init__ buffers (diffusion.py:87-161); n must equal n_timestep + 1. */
SDDM_API int sddm_plan_set_schedule(sddm_plan* plan, const sddm_schedule* sch, int n);

/* Packs weights into kernel layouts and precomputes the per-timestep noise-embedding table
 * E[t] = concat_i Linear_i(noise_level_mlp(sqrt_alpha_bar[t]))  (UNetModified2.py:49-89,168-174,249).
 * Must be called after all weights + the schedule are loaded; synchronises the device once. */
SDDM_API int sddm_plan_finalize(sddm_plan* plan);

/* bytes of device workspace needed by sddm_eps / sddm_sample for a batch of B chunks. */
SDDM_API size_t sddm_workspace_bytes(const sddm_plan* plan, int B);

/* ---- hot path ------------------------------------------------------------------------------------- */
/* replaces: UNetModified2.forward (UNetModified2.py:237-269) = framing + cat + UNet + overlap-add.
 * noise_level: device [B] (per-row, as the module API allows) or NULL, in which case row t of the
 * precomputed table is used for every row (what SDDM.infer does, model.py:108-110). */
SDDM_API int sddm_eps(sddm_plan* plan, const float* cond, const float* x_t, const float* noise_level, int t,
             float* eps_out, int B, void* ws, size_t ws_bytes, void* stream);

/* replaces: GaussianDiffusion.get_x_T / get_x_T_conditional (diffusion.py:281-320).
 * z: injected N(0,1) noise [B,1,L] or NULL => Philox4x32-10 keyed (seed, row0 + row, draw 0). */
SDDM_API int sddm_x_T(sddm_plan* plan, int variant, const float* cond, const float* z, uint64_t seed, int64_t row0,
             float* x_out, int B, void* stream);

/* replaces: GaussianDiffusion.p_transition{,_sr3,_supportive,_conditional} (diffusion.py:164-222),
 * in place on x_t, including the clamp to [-1,1].  z as above (draw index T + 1 - t); ignored for t == 1. */
SDDM_API int sddm_p_step(sddm_plan* plan, int variant, float* x_t, const float* eps, const float* cond, const float* z,
                uint64_t seed, int64_t row0, int t, int B, void* stream);

/* Plan-less forms of the two calls above for callers that hold only the GaussianDiffusion tables
 * (diffusion.p_transition / get_x_T invoked directly, as trainer code may): the step scalars are passed by value.
 *   x_T:   x = a * cond + b * z           (a = sqrt_alpha_bar[T]; b = sqrt(1 - a^2) or sqrt_delta[T])
 *   step:  k8 = {predicted_noise_coeff[t], sqrt(alphas[t]), noise std, gamma, 1 - gamma, c_xt, c_yt, c_epst}[t] */
SDDM_API int sddm_x_T_raw(int variant, float a, float b, const float* cond, const float* z, uint64_t seed, int64_t row0,
                 float* x_out, int B, int L, void* stream);
SDDM_API int sddm_p_step_raw(int variant, const float* k8, float* x_t, const float* eps, const float* cond, const float* z,
                    uint64_t seed, int64_t row0, int t, int T, int B, int L, void* stream);

/* replaces: GaussianDiffusion.q_stochastic / q_stochastic_conditional (diffusion.py:225-279), the forward-diffusion draw of the
 * training step SDDM.forward (model/model.py:29-48).  coef: device [B][4] per-row scalars
 *   mode 0: {sqrt_alpha_bar_sample, sqrt(1 - sample^2), -, -}            x_t = c0 x_0 + c1 z
 *   mode 1: {sqrt_alpha_bar[t], m[t] sqrt_alpha_bar[t], sqrt_delta[t], 1 / sqrt(1 - alpha_bar[t])}
 *           x_t = (c0 x_0 + c1 (y - x_0)) + c2 z;  combined_noise = c3 (c1 (y - x_0) + c2 z)
 * noise: injected N(0,1) [B,1,L] or NULL => Philox (seed, row0 + row, draw 0), written to noise_out when that is not NULL. */
SDDM_API int sddm_q_sample_raw(int mode, const float* coef, const float* x0, const float* y, const float* noise, uint64_t seed, int64_t row0,
                      float* x_t, float* combined_noise, float* noise_out, int B, int L, void* stream);

/* replaces: SDDM.infer (model/model.py:50-124, non-continuous branch): x_T init + T x (eps_hat, update).
 * noises: NULL (Philox) or injected [T, B, L] (noises[0] -> x_T, noises[k] -> step t = T + 1 - k).
 * eps_trace: NULL or [T, B, L] receiving eps_hat of step t at index T - t (verification mode).
 * x_trace:   NULL or [T, B, L] receiving x_{t-1} of step t at index T - t. */
SDDM_API int sddm_sample(sddm_plan* plan, int variant, const float* cond, const float* noises, uint64_t seed, int64_t row0,
                float* out, float* eps_trace, float* x_trace, int B, void* ws, size_t ws_bytes, void* stream);

/* replaces: infer.py:72-77 (H2D copy, model.infer, D2H copy) for HOST buffers: cond/out are host
 * [B,1,L] fp32 (pinned for full-speed copies).  The plan keeps an internal device arena; rows are
 * processed in sub-batches of at most `max_rows_per_pass` (<=0: library default).  Blocks until done. */
SDDM_API int sddm_enhance_host(sddm_plan* plan, int variant, const float* cond_host, float* out_host, int B,
                      uint64_t seed, int64_t row0, int max_rows_per_pass);

/* ---- dataset edge on the device ------------------------------------------------------------------- */
/* The utterances of a batch sit back to back in `flat` (device): utterance u = flat[sample_off[u] .. sample_off[u + 1]) and owns the rows
 * row_off[u] .. row_off[u + 1] of the [N, 1, T] batch (ceil(len / T) rows, last one zero padded).  sample_off / row_off: device int64
 * [n_utt + 1].  A rank of a sharded run converts only its own row range [row_lo, row_hi).
 * replaces: InferDataset.__getitem__ (F.pad + view) + infer_data_collate (torch.cat), data_loader/data_loaders.py:101-155 */
SDDM_API int sddm_chunk_rows(const float* flat, const int64_t* sample_off, const int64_t* row_off, int n_utt, int T, int64_t row_lo,
                    int64_t row_hi, float* rows /* [row_hi - row_lo][T] */, void* stream);
/* the inverse, trimmed to the utterance lengths (only samples of [row_lo, row_hi) are written).
 * replaces: the per-file regroup loop of infer.py:81-120 (output[batch_index_temp].reshape(1, -1)) */
SDDM_API int sddm_regroup_rows(const float* rows /* [row_hi - row_lo][T] */, const int64_t* sample_off, const int64_t* row_off, int n_utt, int T,
                      int64_t row_lo, int64_t row_hi, float* flat_out, void* stream);

/* ---- framing helpers (standalone; the sampler uses fused versions) -------------------------------- */
/* replaces: SignalToFrames.forward (UNetModified2.py:23-28): [B,1,n] -> [B,1,n_frames,F]. */
SDDM_API int sddm_frames(const float* sig, float* frames, int B, int n_samples, int frame_len, int stride, void* stream);
/* replaces: SignalToFrames.overlapAdd (UNetModified2.py:30-41): [B,1,n_frames,F] -> [B,1,n]. */
SDDM_API int sddm_overlap_add(const float* frames, float* sig, int B, int n_samples, int frame_len, int stride, void* stream);

/* ---- STFT feature front-end (cfg 5) --------------------------------------------------------------- */
/* replaces: prepare_spectrogram.py:20-55 = torchaudio Spectrogram / MelSpectrogram (torch.stft, centre = True, reflect padding,
 * onesided, power 1, normalized = "window") followed by clamp((log10(S) - 1 + 5) / 5, 0, 1).
 * wav: device [B, L] fp32; window: device [n_fft] (torch.hamming_window / hann_window, periodic); inv_norm = 1 / sqrt(sum w^2);
 * mel_fb: NULL (linear spectrogram, n_out = n_fft/2 + 1) or device [n_fft/2 + 1, n_mels] triangular filterbank (n_out = n_mels);
 * mel_lo / mel_hi: NULL or device int32 [n_mels], the band [lo, hi) of bins with non-zero weight per filter (skips the zeros);
 * out: device [B, n_out, 1 + L / hop] fp32.  log_clamp = 0 returns the magnitudes themselves.  n_fft must be 1024. */
SDDM_API int sddm_stft_features(const float* wav, int B, int L, int n_fft, int hop, const float* window, float inv_norm,
                                const float* mel_fb, const int32_t* mel_lo, const int32_t* mel_hi, int n_mels, int log_clamp,
                                float* out, void* stream);

/* overlap-add inverse STFT = torch.istft(spec, n_fft = 1024, hop, window = window, center = True, normalized = False, onesided = True,
 * length = L): the inverse of the torch.stft inside the front-end above (north_star names it; the reference never inverts a spectrogram).
 * spec_ri: device [B, 513, frames, 2] fp32 (torch.view_as_real of the complex spectrogram); window: device [1024]; wav_out: device [B, L];
 * ws: device workspace of sddm_istft_workspace_bytes(B, frames) bytes (the windowed frames). */
SDDM_API size_t sddm_istft_workspace_bytes(int B, int frames);
SDDM_API int sddm_istft(const float* spec_ri, int B, int frames, int n_fft, int hop, const float* window, int L, float* wav_out,
               void* ws, size_t ws_bytes, void* stream);

/* ---- cfg 5: DiffWave denoiser + spectrogram-conditioned sampling loop --------------------------------- */
/* replaces: DiffWave (model/diffwave.py:111-155) under SDDM_spectrogram.infer (model/model.py:206-257).
 * Device layout is time-major ([B][T][64] residual stream); the conditioner path (SpectrogramUpsampler + the 30
 * conditioner_projection 1x1 convs, all independent of the diffusion step) is evaluated ONCE per batch by
 * sddm_dw_condition and cached in the workspace, so one eps_hat evaluation only reads it. */
typedef struct sddm_dw_plan sddm_dw_plan;

#define SDDM_DW_COND_SQRT_ALPHA_BAR 0   /* SDDM(noise_condition='sqrt_alpha_bar'): step value = sqrt_alpha_bar[t] */
#define SDDM_DW_COND_TIME_STEP 1        /* config_diffwave.json: step value = t                                   */

/* Mirrors DiffWave.__init__ kwargs (diffwave.py:112-131) + what SDDM_spectrogram needs (model.py:208-210). */
typedef struct sddm_dw_config {
    int32_t n_timestep;            /* T of the diffusion (config_diffwave.json: 200) */
    int32_t freq_bins;             /* 513 */
    int32_t residual_channels;     /* must be 64 */
    int32_t residual_layers;       /* 30 */
    int32_t dilation_cycle_length; /* 10 (dilation of layer i = 2^(i mod cycle), <= 2048) */
    int32_t hop_samples;           /* must be 256 = the upsampler's 16 x 16 */
    int32_t noise_condition;       /* SDDM_DW_COND_* */
    int32_t precision;             /* SDDM_PREC_FP32 (CUDA cores) or SDDM_PREC_BF16 (tcgen05, bf16 residual stream + cache) */
    int32_t reserved[4];
} sddm_dw_config;

SDDM_API int sddm_dw_plan_create(const sddm_dw_config* cfg, sddm_dw_plan** out);
SDDM_API void sddm_dw_plan_destroy(sddm_dw_plan* plan);
/* name = DiffWave state_dict key ("residual_layers.7.dilated_conv.weight", ...), host fp32, shape checked. */
SDDM_API int sddm_dw_plan_load_weight(sddm_dw_plan* plan, const char* name, const void* data, const int64_t* shape, int ndim);
SDDM_API int sddm_dw_plan_set_schedule(sddm_dw_plan* plan, const sddm_schedule* sch, int n);
SDDM_API int sddm_dw_plan_finalize(sddm_dw_plan* plan);
/* workspace for B utterances of `frames` spectrogram frames (audio length hop_samples * frames). */
SDDM_API size_t sddm_dw_workspace_bytes(const sddm_dw_plan* plan, int B, int frames);

/* replaces: SpectrogramUpsampler.forward (diffwave.py:54-61) + every ResidualBlock's conditioner_projection (:87).
 * spec: device [B, freq_bins, frames] fp32.  Must precede sddm_dw_eps on the same workspace. */
SDDM_API int sddm_dw_condition(sddm_dw_plan* plan, const float* spec, int B, int frames, void* ws, size_t ws_bytes, void* stream);
/* replaces: DiffWave.forward (diffwave.py:133-155) for the conditioned workspace.  audio: device [B, 1, 256 * frames];
 * diffusion_step: device [B] step values (as the module API allows) or NULL => the value SDDM_spectrogram.infer
 * passes at step t (model.py:246-252). */
SDDM_API int sddm_dw_eps(sddm_dw_plan* plan, const float* audio, const float* diffusion_step, int t, float* eps_out, int B,
                         int frames, void* ws, size_t ws_bytes, void* stream);
/* replaces: SDDM_spectrogram.infer (model.py:212-257, non-continuous): x_T ~ N(0,1), T x (eps_hat, p_transition).
 * noises: NULL (Philox) or [T, B, L] (noises[0] -> x_T, noises[k] -> step t = T + 1 - k); eps_trace: NULL or [T, B, L]. */
SDDM_API int sddm_dw_sample(sddm_dw_plan* plan, const float* spec, const float* noises, uint64_t seed, int64_t row0, float* out,
                            float* eps_trace, int B, int frames, void* ws, size_t ws_bytes, void* stream);
/* per-launch CUDA-event timing of the tcgen05 path (bench roofline): enable resets the totals; kind 0 = residual-layer kernel,
 * kind 1 = skip / output head.  Both calls synchronise the device. */
SDDM_API int sddm_dw_profile_enable(sddm_dw_plan* plan, int on);
SDDM_API int sddm_dw_profile_read(sddm_dw_plan* plan, int kind, double* total_ms, int64_t* launches);
/* test hook: "upsampled" ([T, freq_bins] of utterance B-1), "x" (residual stream after the last eps call, [B, T, 64]),
 * "skip" (sum of skips, [B, T, 64]), "cond<i>" (cached conditioner of layer i, [B, T, 128]) -> fp32 out; *n = element count. */
SDDM_API int sddm_dw_debug_fetch(sddm_dw_plan* plan, const char* what, void* ws, int B, int frames, float* out, int64_t* n,
                                 void* stream);

/* ---- cfg 4: WaveGrad denoiser + spectrogram-conditioned sampling loop -------------------------------- */
/* replaces: WaveGrad (model/wavegrad.py:140-179; the architecture has no constructor arguments: 5 DBlocks / FiLMs / UBlocks,
 * 128 mel bins, 300 samples per frame) under SDDM_spectrogram.infer (model/model.py:206-257). */
typedef struct sddm_wg_plan sddm_wg_plan;

typedef struct sddm_wg_config {
    int32_t n_timestep;       /* T of the diffusion (config_wavegrad.json: 1000) */
    int32_t hop_samples;      /* must be 300 = 5 * 5 * 3 * 2 * 2 */
    int32_t noise_condition;  /* SDDM_DW_COND_* (config_wavegrad.json: sqrt_alpha_bar, the SDDM default) */
    int32_t precision;        /* SDDM_PREC_FP32 (CUDA cores) or SDDM_PREC_BF16 (tcgen05 convs, bf16 activations) */
    int32_t reserved[4];
} sddm_wg_config;

SDDM_API int sddm_wg_plan_create(const sddm_wg_config* cfg, sddm_wg_plan** out);
SDDM_API void sddm_wg_plan_destroy(sddm_wg_plan* plan);
/* name = WaveGrad state_dict key ("upsample.2.block3.1.weight", ...) or "film.<i>.encoding.frequencies" (the
 * PositionalEncoding frequency vector exp(-ln(1e4) k / (dim/2)), wavegrad.py:45-46); host fp32, shape checked. */
SDDM_API int sddm_wg_plan_load_weight(sddm_wg_plan* plan, const char* name, const void* data, const int64_t* shape, int ndim);
SDDM_API int sddm_wg_plan_set_schedule(sddm_wg_plan* plan, const sddm_schedule* sch, int n);
SDDM_API int sddm_wg_plan_finalize(sddm_wg_plan* plan);
SDDM_API size_t sddm_wg_workspace_bytes(const sddm_wg_plan* plan, int B, int frames);
/* replaces: WaveGrad.forward (wavegrad.py:167-179).  spec: device [B, 128, frames]; audio: device [B, 300 * frames];
 * noise_level: device [B] or NULL => the value SDDM_spectrogram.infer passes at step t; eps_out: device [B, 300 * frames]. */
SDDM_API int sddm_wg_eps(sddm_wg_plan* plan, const float* spec, const float* audio, const float* noise_level, int t, float* eps_out,
                         int B, int frames, void* ws, size_t ws_bytes, void* stream);
/* replaces: SDDM_spectrogram.infer (model.py:212-257, non-continuous) around WaveGrad; buffers as sddm_dw_sample. */
SDDM_API int sddm_wg_sample(sddm_wg_plan* plan, const float* spec, const float* noises, uint64_t seed, int64_t row0, float* out,
                            float* eps_trace, int B, int frames, void* ws, size_t ws_bytes, void* stream);
/* per-launch CUDA-event timing of the tcgen05 conv launches (bench roofline): enable resets the totals; read returns the summed
 * duration, the number of launches, the GEMM flops they executed and their algorithmic bytes.  Both calls synchronise the device. */
SDDM_API int sddm_wg_profile_enable(sddm_wg_plan* plan, int on);
SDDM_API int sddm_wg_profile_read(sddm_wg_plan* plan, double* total_ms, int64_t* launches, double* flops, double* bytes);
/* test hook: activation "d0".."d4" (downsample outputs) / "u0".."u4" (upsample outputs) of the last sddm_wg_eps call on this
 * workspace -> out [B, L, C] fp32 (time-major); shape2 receives {L, C}. */
SDDM_API int sddm_wg_debug_fetch(sddm_wg_plan* plan, const char* what, void* ws, int B, int frames, float* out, int64_t* shape2,
                                 void* stream);

/* ---- introspection / test hooks ------------------------------------------------------------------- */
/* number of kernel launches one sddm_eps call enqueues for this plan. */
SDDM_API int sddm_plan_launches_per_eps(const sddm_plan* plan);
/* per-launch CUDA-event timing of the op program (bench.py's live roofline measurement; off by default).
 * ops 0..num_ops-2 are the UNet program in launch order, the last index is the overlap-add + posterior kernel. */
SDDM_API int sddm_plan_num_ops(const sddm_plan* plan);
SDDM_API int sddm_profile_enable(sddm_plan* plan, int on);   /* also resets the accumulated totals */
SDDM_API int sddm_profile_read(sddm_plan* plan, int op, double* total_ms, int64_t* launches, double* flops_per_row,
                      double* bytes_per_row, int* uses_tensor_cores, char* label, int label_cap);
/* ---- SNR-adaptive diffusion (diffusion half) -------------------------------------------------------------------
 * reference: VariableGaussianDiffusion, model/diffusion.py:329-446.  Frames x [B][N][L] fp32 (the reference's [B, 1, N, L]),
 * snr [B][N] fp32 (dB, one estimate per frame), T = n_timestep, scale = snr_estimate_scale.  Every frame has its own linear beta
 * schedule (end value from its SNR); the kernels recompute the few schedule terms they need per frame - the reference rebuilds the
 * whole [B, 1, N, T + 1] schedule with numpy on the host in every call.  z == NULL: in-kernel Philox4x32-10 normals keyed by the global
 * frame id (row0 + b) * N + n.  The two networks of that variant (SNREstimator, UNetModified2_withVariableNoiseLevel) are not built. */
/* get_beta_schedule (:345-359): betas, alpha_bar [B][N][T + 1] */
SDDM_API int sddm_var_schedule(const float* snr, int B, int N, int T, float scale, float* betas, float* alpha_bar, void* stream);
/* get_noise_level (:438-444): out[b][n] = sqrt(alpha_bar_t) */
SDDM_API int sddm_var_noise_level(const float* snr, int B, int N, int T, float scale, int t, float* out, void* stream);
/* get_x_T (:417-435, t = T, x = condition) and q_stochastic with an integer step (:394-415, x = x_0, z = the noise):
 * out = s x + sqrt(1 - s^2) z with s = sqrt(alpha_bar_t) of the frame; noise_level (may be NULL) receives s, [B][N] */
SDDM_API int sddm_var_mix(const float* x, const float* snr, const float* z, uint64_t seed, int64_t row0, int B, int N, int L, int T, float scale,
                          int t, float* out, float* noise_level, void* stream);
/* p_transition (:373-391): out = clamp((x_t - beta_t / sqrt(1 - ab_t) eps) / sqrt(1 - beta_t) [+ sigma_t z if t > 1], -1, 1) */
SDDM_API int sddm_var_posterior(const float* x_t, const float* eps, const float* snr, const float* z, uint64_t seed, int64_t row0, int B, int N,
                                int L, int T, float scale, int t, float* out, void* stream);

/* copies the NHWC activation of a named UNet node ("downs.3", "mid.0", "ups.7", ...) produced by the last
 * sddm_eps call on this workspace into out (device, [B,H,W,C] fp32); returns C*H*W via *chw. */
SDDM_API int sddm_debug_fetch(sddm_plan* plan, const char* node, void* ws, int B, float* out, int64_t* chw, void* stream);
/* single-tile tcgen05 descriptor self-test: returns max |err| of a 128 x N x K bf16 MMA vs an fp32 reference
 * computed on the device with CUDA cores, through *max_err. variant selects the descriptor convention. */
SDDM_API int sddm_debug_umma_probe(int variant, int N, int K, float* max_err_host);
/* pipeline trace of the tcgen05 conv kernel: enable != 0 starts recording (next 64 conv launches, CTA 0 of each: cycles per
 * role spent waiting on each barrier); enable == 0 copies the [64][48] int64 counters to host_out and stops. */
SDDM_API int sddm_debug_tc_trace(int enable, long long* host_out);
/* the same for the row-streaming kernel of the 128-wide level (conv_row.cu): [64][32] int64 counters per launch:
 * [0..7] / [8..15] epilogue group 0 / 1: total, wait acc_full, TMEM loads, store-buffer wait, residual wait, barrier 1, barrier 2, rows;
 * [16..20] MMA warp: total, wait weights, wait acc_empty, wait full_a, rows; [21..22] raw loader: total, wait raw_empty;
 * [24..27] / [28..31] transform group 0 / 1: total, wait raw_full, wait empty_a, work */
SDDM_API int sddm_debug_row_trace(int enable, long long* host_out);
/* debug: enable != 0 -> a wait of the row kernel (and, in a build with -DSDDM_TC_HANG_NOTES=1, of conv3x3_tc_kernel) that times out
 * (pipeline bug; the kernel then traps) first notes who waits on what in mapped host memory; enable == 0 -> copy the 65536 words out:
 * [0] 1 = some wait timed out, [1] shared-memory base of the CTA header, [2], [3] plan of the last launch that started,
 * [4 + 4 (3000 (launch id mod 4) + 20 cta + warp) ..] = (cta + 1, thread, barrier address, parity | launch id << 1) of a warp that timed out,
 * [60000 + 8 (launch id mod 4) ..] = (launch id, header base, ring depths, slab counts, Cin, Cout, Hout, grid) of the conv3x3_tc launches */
SDDM_API int sddm_debug_hang(int enable, unsigned* host_out);
/* issue-rate microbenchmark: average cycles per back-to-back tcgen05.mma (M 128, K 16, bf16, shared-memory operands) */
SDDM_API int sddm_debug_umma_rate(int N, int reps, int nA, int geo, float* cycles_per_mma);

#ifdef __cplusplus
}
#endif
#endif /* SDDM_B200_H */
