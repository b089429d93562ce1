/* flowtron_b200 — C ABI of the B200-native Flowtron hot path (libflowtron_b200.so).
 *
 * The reference (NVIDIA/flowtron) has no FFI: the hot path sits behind Python nn.Modules.  These entry
 * points are what a binding for that path replaces; each cites the reference interface it stands for.
 * Conventions: all pointers are DEVICE pointers owned by the caller (torch's allocator in the shipped
 * host layer); no entry point allocates, synchronises the stream, or touches the default stream unless
 * `stream` is 0; return 0 on success, negative on error (message in ft_last_error()).  Weights are
 * passed in PyTorch's native layouts ([4H, in], gate order i,f,g,o) so checkpoints need no conversion.
 */
#ifndef FLOWTRON_B200_H
#define FLOWTRON_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* element formats */
#define FT_F16 0
#define FT_BF16 1
#define FT_TF32 2   /* fp32 storage, tf32 tensor-core math */

int ft_version(void);
const char* ft_last_error(void);
/* 0 = ok; non-zero = code written by a device-side watchdog (a bounded spin-wait expired). Synchronises. */
int ft_device_status(void);
/* Optional per-kernel device timing with CUDA events on the launching stream (bench.py roofline leg). */
void ft_timing_enable(int on);
void ft_timing_reset(void);
int ft_timing_report(char* buf, int cap);   /* synchronises; "name m n k count total_ms" per line */
/* number of kernels this library has launched (bench.py's gpu_launches) */
long long ft_launch_count(void);
void ft_reset_launch_count(void);

/* C[M,N] = act(alpha * A[M,K] B[N,K]^T + bias + bias2) (+ C32 if beta).  tcgen05/TMA GEMM.
 * a_mn / b_mn = 1: operand is given MN-major, i.e. A points at a row-major [K, M] tensor.
 * Replaces: nn.Linear / Conv1d(k=1) / LSTM input projections (flowtron.py:278-288, 651, 654-655) and
 * their autograd dgrad / wgrad. */
int ft_gemm(int M, int N, int K, const void* A, long long lda, int a_fmt, int a_mn, const void* B, long long ldb,
            int b_fmt, int b_mn, const float* bias, const float* bias2, int act, int beta, float alpha, float* C32,
            long long ldc32, void* C16, long long ldc16, int c16_fmt, void* stream);

/* One nn.LSTM layer recurrence, hidden size 1024 (flowtron.py:654-655 `lstm` / `attention_lstm`, run via
 * run_padded_sequence :671-695).  xproj = W_ih x + b_ih + b_hh for all T*B rows (row = t*B + b), fp32
 * [T*B, 4096], gate order i,f,g,o.  whh16: fp16 copy of weight_hh [4096,1024].  Writes h (fp16, zero at
 * t >= lens[b]) to hseq16 [T*B, ldh] and, when non-null, the post-activation gates (fp16 [T*B,4096]) and
 * cell state (fp32 [T*B,1024]) needed by ft_lstm_bwd.  flags: int scratch [T*16].  B <= 128. */
int ft_lstm_fwd(int T, int B, const float* xproj, const void* whh16, const int* lens, void* hseq16, long long ldh,
                void* gates16, float* cstate, float* h32, long long ldh32, int* flags, void* stream);

/* debug/profiling: device buffer [T][16] that receives clock64 stamps of CTA 0 of later ft_lstm_fwd launches */
void ft_debug_set_lstm_trace(long long* buf);
/* same for ft_ar_step_infer: [T][32] stamps of CTA 0 (phase boundaries, tools/trace_infer.py) */
void ft_debug_set_infer_trace(long long* buf);

/* Process-wide switch: 1 = ft_lstm_fwd (and ft_ar_step_fwd) use the 64-CTA recurrence kernel so two half-batch
 * launches on two streams run concurrently (the BPTT kernel always uses 64 CTAs for B <= 32). */
void ft_set_lstm_half_sm(int on);
/* ft_gemm tile policy: 0 = one CTA per 128 x 256 tile always; 1 (default) = CTA pairs (tcgen05 cta_group::2, 256 x 256 tiles) for
 * contractions >= 1e11 FLOP; 2 = CTA pairs whenever the output is wide and has >= SMs/2 such tiles. */
void ft_set_gemm_pair_mode(int mode);

/* BPTT of the same layer (autograd of nn.LSTM).  dh_ext: gradient w.r.t. the layer outputs, fp32
 * [T*B, ldd] (ignored at t >= lens[b]; the caller may pre-multiply it by a power-of-two loss scale).  whhT16:
 * fp16 copy of weight_hh^T [1024,4096].  Writes dG (fp16 [T*B,4096], saturating), the gradient w.r.t. the gate
 * pre-activations; dW_ih, dW_hh, db and dx are GEMMs / column sums over dG done by the caller.
 * flags: int scratch [T*64]. */
int ft_lstm_bwd(int T, int B, const float* dh_ext, long long ldd, const void* whhT16, const void* gates16,
                const float* cstate, const int* lens, void* dG16, int* flags, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * AR_Step / AR_Back_Step (flowtron.py:595-828): one autoregressive flow step, training direction.
 * ---------------------------------------------------------------------------------------------------- */
typedef struct {
    int T, B, L;        /* mel frames (padded max), batch, text tokens (padded max) */
    int n_mel;          /* 80  (n_mel_channels)                                     */
    int n_hidden;       /* 1024 (must be 1024 in this build)                        */
    int n_attn;         /* 640 (n_attn_channels)                                    */
    int n_text;         /* 640 = n_text_dim + n_speaker_dim                         */
    int reversed;       /* 1 = AR_Back_Step semantics (flowtron.py:605-627)         */
    int has_gate;       /* flow owns gate_layer (last flow only, flowtron.py:854)   */
    int has_prior;      /* attn_prior given                                         */
    float temperature;  /* Attention.temperature (flowtron.py:532, 573)             */
} FtArStepDesc;

/* fp32 device pointers in PyTorch's native layouts == the state_dict entries of one AR_Step
 * (SURVEY.md 8a row 15).  The same struct type carries the parameter gradients in ft_ar_step_bwd. */
typedef struct {
    float *attn_lstm_w_ih, *attn_lstm_w_hh, *attn_lstm_b_ih, *attn_lstm_b_hh;   /* attention_lstm.*_l0  [4H,M] [4H,H] [4H] [4H] */
    float *lstm_w_ih0, *lstm_w_hh0, *lstm_b_ih0, *lstm_b_hh0;                   /* lstm.*_l0            [4H,H+A] [4H,H] ...      */
    float *lstm_w_ih1, *lstm_w_hh1, *lstm_b_ih1, *lstm_b_hh1;                   /* lstm.*_l1            [4H,H] [4H,H] ...        */
    float *att_query, *att_key, *att_value, *att_v;                            /* attention_layer.{query,key,value,v}.linear_layer.weight */
    float *dense_w0, *dense_b0, *dense_w1, *dense_b1;                          /* dense_layer.layers.{0,1}.linear_layer.{weight,bias}     */
    float *conv_w, *conv_b;                                                     /* conv.weight [2M,H,1], conv.bias [2M]                    */
    float *gate_w, *gate_b;                                                     /* gate_layer.linear_layer.{weight,bias} or NULL           */
} FtArStepWeights;

/* Bytes of the two caller-owned work areas: `saved` persists from ft_ar_step_fwd to ft_ar_step_bwd of the
 * same flow; `scratch` is transient and may be shared by all flows on a stream. */
size_t ft_ar_step_saved_bytes(const FtArStepDesc* d);
size_t ft_ar_step_scratch_bytes(const FtArStepDesc* d);
/* Byte offset / size of a named intermediate inside `saved` (debug + tests): returns 0 on success. */
int ft_ar_step_saved_lookup(const FtArStepDesc* d, const char* name, size_t* offset, size_t* bytes);

/* AR_Step.forward / AR_Back_Step.forward (flowtron.py:725-773, 605-627).
 *   mel [T,B,M] f32, text [L,B,E] f32 (encoder outputs), in_lens/out_lens int32 [B] (NULL = no padding),
 *   attn_prior [B,T,L] f32 or NULL  ->  mel_out [T,B,M], log_s [T,B,M], gates [T,B] (NULL if !has_gate),
 *   attn [B,T,L], attn_logprob [B,T,L].  For reversed flows log_s/gates/attn are in flipped time exactly as
 *   the reference returns them. */
int ft_ar_step_fwd(const FtArStepDesc* d, const FtArStepWeights* w, const float* mel, const float* text,
                   const int* in_lens, const int* out_lens, const float* attn_prior, float* mel_out, float* log_s,
                   float* gates, float* attn, float* attn_logprob, void* saved, void* scratch, void* stream);

/* Optional, one-shot, per host thread: the NEXT ft_ar_step_fwd call on this thread waits for `cuda_event` (a
 * cudaEvent_t recorded on whatever stream produced `text`) on its own stream right before it first reads `text` --
 * i.e. after the attention LSTM, which depends on mel only -- so the caller can run the text encoder
 * (flowtron.py:871-880) concurrently with the first third of the flow instead of in front of it.  NULL clears it. */
void ft_ar_step_set_text_ready_event(void* cuda_event);

/* Autograd of the above.  Incoming gradients (any may be NULL = zero): d_mel_out, d_log_s [T,B,M], d_gates [T,B],
 * d_attn, d_attn_logprob [B,T,L].  Outputs: d_mel [T,B,M], d_text [L,B,E] (overwritten), parameter gradients
 * into `g` (overwritten; same layouts as the weights). */
int ft_ar_step_bwd(const FtArStepDesc* d, const FtArStepWeights* w, const float* mel, const int* in_lens,
                   const int* out_lens, const float* attn, const float* d_mel_out, const float* d_log_s,
                   const float* d_gates, const float* d_attn, const float* d_attn_logprob, float* d_mel, float* d_text,
                   const FtArStepWeights* g, void* saved, void* scratch, void* stream);

/* The same backward as two calls, so that the caller's autograd can overlap work that depends only on d_text (the text
 * encoder's backward) with the attention LSTM's BPTT (a 64-SM kernel):
 *   ft_ar_step_bwd_main      : everything but the attention LSTM.  Writes d_text and the gradients of conv, dense_layer,
 *                              lstm layer 1 and the gate; leaves dhA / the coupling-path input gradient / the loss scale in
 *                              `carry` (ft_ar_step_bwd_carry_bytes, caller-owned, must survive until the second call) and
 *                              dG0 / dQ / dK / dV in `scratch` (nothing else may use `scratch` in between).
 *   ft_ar_step_bwd_attn_lstm : attention-LSTM BPTT; writes d_mel and the gradients of attention_lstm, lstm layer 0 and the
 *                              attention query / key / value projections (their GEMMs run underneath the BPTT kernel).
 * ft_ar_step_bwd == the two back to back. */
size_t ft_ar_step_bwd_carry_bytes(const FtArStepDesc* d);
int ft_ar_step_bwd_main(const FtArStepDesc* d, const FtArStepWeights* w, const float* mel, const int* in_lens,
                        const int* out_lens, const float* attn, const float* d_mel_out, const float* d_log_s,
                        const float* d_gates, const float* d_attn, const float* d_attn_logprob, float* d_text,
                        const FtArStepWeights* g, void* saved, void* scratch, void* carry, void* stream);
int ft_ar_step_bwd_attn_lstm(const FtArStepDesc* d, const FtArStepWeights* w, const int* out_lens, float* d_mel,
                             const FtArStepWeights* g, void* saved, void* scratch, void* carry, void* stream);

/* FlowtronLoss default branch (flowtron.py:205-243): sums[0]=sum (z m)^2, [1]=sum_flows sum log_s m,
 * [2]=sum m BCEWithLogits(gate m, target), [3]=n=sum m.  log_s_list: HOST array of n_flows (<= 16) device pointers (passed
 * to the kernel by value: no upload, so the call can be captured in a CUDA graph). */
int ft_nll_reduce(const float* z, const float* const* log_s_list, int n_flows, const float* gate,
                  const float* gate_target, const int* out_lens, int T, int B, int M, float* sums, void* stream);
/* Gradients of g_nll*nll + g_gate*gate_loss w.r.t. z, every log_s (identical tensor), gate logits. */
int ft_nll_grad(const float* z, const float* gate, const float* gate_target, const int* out_lens, int T, int B, int M,
                float sigma, const float* sums, const float* g_nll, const float* g_gate, float* dz, float* dlog_s,
                float* dgate, void* stream);

/* Optimizer step of the training loop around the hot path (SURVEY.md 8f "next"): torch.nn.utils.clip_grad_norm_
 * (train.py:324-329) followed by RAdam.step (radam.py:44-122), on flat fp32 device buffers, no host sync.
 *   ft_sumsq_partials: partials[0..1023] = per-block sums of x[i]^2 over x[0..n) (one call per contiguous gradient
 *     segment, each into its own 1024 floats);
 *   ft_clip_coef: norm_coef[0] = sqrt(sum partials) (the total gradient norm clip_grad_norm_ returns),
 *     norm_coef[1] = min(1, max_norm / (norm + 1e-6)) (the factor it multiplies every gradient by);
 *   ft_radam_step: p, m (exp_avg), v (exp_avg_sq) updated in place from g * grad_coef[0] (grad_coef NULL = 1):
 *     v = beta2 v + (1-beta2) g^2; m = beta1 m + (1-beta1) g; p -= weight_decay_lr p (weight_decay*lr, radam.py:109);
 *     use_denom (N_sma >= 5, radam.py:115): p -= step_size m / (sqrt(v) + eps), else p -= step_size m.
 *     N_sma / step_size are the host scalars of radam.py:87-107. */
int ft_sumsq_partials(const float* x, long long n, float* partials, void* stream);
int ft_clip_coef(const float* partials, int n_partials, float max_norm, float* norm_coef, void* stream);
int ft_radam_step(float* p, const float* g, float* m, float* v, long long n, double beta1, double beta2, double eps,
                  double weight_decay_lr, double step_size, int use_denom, const float* grad_coef, void* stream);
/* CUDA-graph-capturable form: the step count lives on the device (step_dev[0] = updates done so far); the kernel derives
 * N_sma / step_size from step_dev[0] + 1 itself (fp64), so a captured step replays with the right schedule.  Call
 * ft_step_increment once after all segments of an optimizer step. */
int ft_radam_step_dev(float* p, const float* g, float* m, float* v, long long n, double beta1, double beta2, double eps,
                      double weight_decay, double lr, const int* step_dev, const float* grad_coef, void* stream);
int ft_step_increment(int* step_dev, void* stream);

/* AR_Step.infer (flowtron.py:775-828): sequential inverse of one flow for all T frames in ONE persistent launch.
 *   residual [T,B,M] f32 in flow-time order (the caller flips for AR_Back_Step, :629-642), text [L,B,E] f32,
 *   attn_prior [B,T,L] or NULL (row i is used at frame i); attn_forced [T,B,L] or NULL: forced alignments (the `attns`
 *   argument, flowtron.py:585-588, 797: scoring, softmax and the prior are skipped, context = attn_forced[i] . V).
 *   Uses d->T/B/L/n_*, has_gate, has_prior, temperature.
 *   out [T,B,M]: generated frames (zeros after a sample's gate fired; the frame that trips the gate is emitted),
 *   attn_out [T,B,L]: per-frame attention weights, n_frames [B] int32: frames emitted per sample.
 *   With B == 1 this is exactly the reference's loop-with-break; B > 1 with a gate is the per-sample-stop
 *   extension (the reference raises there). */
size_t ft_ar_step_infer_scratch_bytes(const FtArStepDesc* d);
int ft_ar_step_infer(const FtArStepDesc* d, const FtArStepWeights* w, const float* residual, const float* text,
                     const float* attn_prior, const float* attn_forced, float gate_threshold, float* out, float* attn_out,
                     int* n_frames, void* scratch, void* stream);

/* TacotronSTFT.mel_spectrogram (audio_processing.py:117-134) for a ragged batch.  wav: concatenated utterances
 * (f32 in [-1,1]); sample_offsets / frame_offsets: device int64 [n_utt+1] prefix sums (frames of utterance u =
 * 1 + N_u / hop); window [n_fft] (periodic Hann, centre-padded); mel_basis [n_mel, n_fft/2+1] with its non-zero
 * column range [band_lo[m], band_hi[m]) per row.  mel_out: per-utterance [n_mel, F_u] blocks, concatenated
 * (== [B, n_mel, F] for equal lengths).  scratch: ft_mel_scratch_bytes(n_fft, chunk_frames). */
size_t ft_mel_scratch_bytes(int n_fft, long long chunk_frames);
int ft_mel_spectrogram(const float* wav, const long long* sample_offsets, const long long* frame_offsets, int n_utt,
                       long long total_frames, const float* window, const float* mel_basis, const int* band_lo,
                       const int* band_hi, int n_mel, int n_fft, int hop, float clip, float* mel_out, void* scratch,
                       long long chunk_frames, void* stream);

/* Fused form for n_fft == 1024 (the shipped configuration, config.json:31-33): ONE kernel -- reflect pad + window on the
 * load, a 1024-point real FFT per warp in shared memory, |X|, sparse filterbank, log -- reads the waveform once and writes
 * the mel once (no frames / spectrum round trip, no scratch).  wav_fmt: 0 = f32 in [-1,1], 1 = int16 PCM scaled by
 * 1/32768 on load (data.py:150 `audio / max_wav_value`).  Same output layout as ft_mel_spectrogram. */
int ft_mel_spectrogram_fused(const void* wav, int wav_fmt, const long long* sample_offsets, const long long* frame_offsets,
                             int n_utt, long long total_frames, const float* window, const float* mel_basis, const int* band_lo,
                             const int* band_hi, int n_mel, int n_fft, int hop, float clip, float* mel_out, void* stream);

/* STFT.transform (audio_processing.py:207-235), n_fft == 1024: magnitude and phase (atan2(imag, real)) as per-utterance
 * [n_fft/2+1, F_u] blocks (== [B, 513, F] for equal lengths), same kernel as above without the filterbank. */
int ft_stft_transform(const float* wav, const long long* sample_offsets, const long long* frame_offsets, int n_utt,
                      long long total_frames, const float* window, int n_fft, int hop, float* magnitude, float* phase,
                      void* stream);

/* Data layer on the device.
 * ft_attn_prior: beta_binomial_prior_distribution (data.py:31-41) for a whole batch directly in DataCollate's zero-padded
 *   [B,T,L] layout (data.py:222-243): prior[b,i-1,k] = BetaBinomial(k; n=in_lens[b]-1, a=s*i, b=s*(out_lens[b]+1-i)) for
 *   i <= out_lens[b], k < in_lens[b], else 0; values < threshold are zeroed when threshold > 0 (data.py:137-139).
 *   fp64 log-gamma evaluation, fp32 output.
 * ft_collate_mel: DataCollate's mel / gate padding (data.py:214-236) from ft_mel_spectrogram's packed output: row i of
 *   mel_padded [B,n_mel,T] is utterance order[i]; gate_padded [B,T] = 1 from its last frame on; out_lens[i] = its frames. */
int ft_attn_prior(const int* in_lens, const int* out_lens, int B, int T, int L, float scaling, float threshold, float* prior,
                  void* stream);
int ft_collate_mel(const float* mel_packed, const long long* frame_offsets, const int* order, int B, int n_mel, int T,
                   float* mel_padded, float* gate_padded, int* out_lens, void* stream);

/* AttentionCTCLoss (flowtron.py:155-182) for a whole batch, with its gradient, in one launch (the
 * reference loops over utterances in Python).  attn_logprob [B,T,L] f32 of ONE flow; classes = blank (constant logit
 * blank_logprob) + the utterance's in_lens[b] tokens, log_softmax over them, CTC against the target 1..in_lens[b] over
 * out_lens[b] frames, reduction 'mean' (nll / in_lens[b]), zero_infinity.  time_reversed = 1: the tensor is in an
 * AR_Back_Step's flipped time (flowtron.py:250-256).  cost [B]; d_attn_logprob [B,T,L] = d cost[b] / d attn_logprob (zeros
 * outside the utterance's frames / tokens).  scratch: ft_attn_ctc_scratch_bytes(B, T, L). */
size_t ft_attn_ctc_scratch_bytes(int B, int T, int L);
int ft_attn_ctc_loss(const float* attn_logprob, const int* in_lens, const int* out_lens, int B, int T, int L, int time_reversed,
                     float blank_logprob, float* cost, float* d_attn_logprob, void* scratch, void* stream);


/* ------------------------------------------------------------------------------------------------------
 * Text Encoder (flowtron.py:467-525): 3 x [Conv1d(512, 512, k=5) -> MaskedInstanceNorm1d(affine) -> relu -> dropout(0.5)],
 * then a packed bidirectional LSTM(512 -> 2 x 256); forward + backward.
 * Replaces Encoder.forward (masked = 1: tokens at l >= in_lens[b] are zeroed before every convolution, statistics run over
 * the utterance's own tokens, the BiLSTM behaves as on a packed sequence, outputs are zero at l >= in_lens[b]) and
 * Encoder.infer / the B == 1 case (masked = 0: no lengths).  Weights in PyTorch layouts; index 1 of the LSTM arrays is the
 * `_reverse` direction.
 * ---------------------------------------------------------------------------------------------------- */
typedef struct {
    int B, L;           /* utterances (1..64), padded text length                               */
    int C;              /* 512 (encoder_embedding_dim; must be 512 in this build)               */
    int n_convs, ksize; /* 3, 5                                                                 */
    int masked;         /* 1: use in_lens (Encoder.forward, B > 1); 0: every token valid        */
    float dropout_p;    /* 0 (eval) or the training probability (flowtron.py:502 hard-codes 0.5) */
    float eps;          /* instance-norm eps (1e-5)                                             */
} FtEncoderDesc;

typedef struct {
    const float* conv_w[3]; const float* conv_b[3];    /* [512,512,5], [512] */
    const float* norm_w[3]; const float* norm_b[3];    /* [512]              */
    const float* w_ih[2]; const float* w_hh[2];        /* [1024,512], [1024,256] */
    const float* b_ih[2]; const float* b_hh[2];        /* [1024]             */
} FtEncoderWeights;

typedef struct {                                       /* gradients, same shapes, OVERWRITTEN */
    float* d_conv_w[3]; float* d_conv_b[3];
    float* d_norm_w[3]; float* d_norm_b[3];
    float* d_w_ih[2]; float* d_w_hh[2];
    float* d_b_ih[2]; float* d_b_hh[2];
} FtEncoderGrads;

size_t ft_encoder_saved_bytes(const FtEncoderDesc* d);          /* forward -> backward tensors */
size_t ft_encoder_fwd_scratch_bytes(const FtEncoderDesc* d);
size_t ft_encoder_bwd_scratch_bytes(const FtEncoderDesc* d);
/* x [B, 512, L] fp32 (embedding output, channels first, as Encoder.forward receives it); in_lens int32 [B] (masked only);
 * rng_state: device uint64[2] = {seed, call counter} for the dropout masks (advanced by the call; NULL when dropout_p == 0);
 * out: fp32, element (b, l, c) at out[b * out_stride_b + l * out_stride_l + c], c < 512 (forward direction first) -- lets the
 * caller write straight into the [L, B, 640] text tensor the flows read. */
int ft_encoder_fwd(const FtEncoderDesc* d, const FtEncoderWeights* w, const float* x, const int* in_lens,
                   unsigned long long* rng_state, float* out, long long out_stride_b, long long out_stride_l, void* saved,
                   void* scratch, void* stream);
/* Token form: the nn.Embedding lookup in front of the encoder (flowtron.py:873 `self.embedding(text).transpose(1, 2)`) fused into
 * the first kernel (tokens int64 [B, L], emb_weight fp32 [n_text, 512]); the backward scatters into d_emb_weight (overwritten). */
int ft_encoder_fwd_tokens(const FtEncoderDesc* d, const FtEncoderWeights* w, const long long* tokens, const float* emb_weight,
                          int n_text, const int* in_lens, unsigned long long* rng_state, float* out, long long out_stride_b,
                          long long out_stride_l, void* saved, void* scratch, void* stream);
int ft_encoder_bwd_tokens(const FtEncoderDesc* d, const FtEncoderWeights* w, const float* d_out, long long d_stride_b,
                          long long d_stride_l, const void* saved, const long long* tokens, int n_text, float* d_emb_weight,
                          const FtEncoderGrads* grads, void* scratch, void* stream);
/* d_out: dense [B, L, 512] fp32; d_x [B, 512, L] fp32 (may be NULL). */
int ft_encoder_bwd(const FtEncoderDesc* d, const FtEncoderWeights* w, const float* d_out, long long d_stride_b,
                   long long d_stride_l, const void* saved, float* d_x, const FtEncoderGrads* grads, void* scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif
