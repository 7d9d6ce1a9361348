/*
 * rnnt_b200.h -- C ABI of the B200-native streaming RNN-Transducer inference path.
 *
 * This is the drop-in boundary for the ONE hot path of iceychris/LibreASR
 * (audio -> log-mel/stack -> stacked-LSTM encoder -> GRU predictor -> joint ->
 * greedy RNN-T decode).  The reference has no FFI layer of its own: its
 * "operator API" for this path is the Python module surface of
 * libreasr/lib/models.py and the YAML-named transforms of libreasr/lib/transforms.py
 * (SURVEY.md section 8b).  Each entry point below therefore names the reference
 * function(s) it replaces (paths relative to the reference root); the Python package
 * `libreasr_b200` binds these symbols with ctypes and re-exposes the reference's
 * module surface on top of them (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - plain C types only; every pointer marked "dev" is a CUDA device pointer owned
 *     by the caller, "host" is host memory; no allocation happens inside hot calls
 *     once rnnt_b200_reserve() covered the shapes (workspaces grow on demand otherwise)
 *   - all kernels are enqueued on `stream` (a cudaStream_t passed as void*); calls are
 *     asynchronous unless stated otherwise
 *   - every function returns 0 on success, a negative rnnt_b200_status otherwise, and
 *     never throws; rnnt_b200_last_error() gives the message
 *   - tensors are dense row-major fp32 unless stated; B = utterances/streams,
 *     n = samples, T = encoder steps, X = n_mels*n_stack, H = hidden, J = joint, V = vocab
 *   - a handle may be used from several host threads as long as concurrent calls use
 *     distinct CUDA streams AND distinct caller-owned buffers; workspace-using calls
 *     (transcribe*, encode, decode) on one handle must be serialised by the caller
 */
#ifndef RNNT_B200_H_
#define RNNT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RNNT_B200_ABI_VERSION 2

typedef struct rnnt_b200_handle_s* rnnt_b200_handle;

typedef enum {
  RNNT_B200_OK = 0,
  RNNT_B200_ERR_INVALID = -1,   /* bad argument / shape (reference: ValueError, base_rnn.py:81-117) */
  RNNT_B200_ERR_STATE = -2,     /* call order (weights missing, not finalized) */
  RNNT_B200_ERR_CUDA = -3,      /* CUDA runtime error; message holds cudaGetErrorString */
  RNNT_B200_ERR_NOMEM = -4,
  RNNT_B200_ERR_UNSUPPORTED = -5
} rnnt_b200_status;

/* Arithmetic used for the dense contractions (LSTM gate GEMMs, joint projections). */
typedef enum {
  RNNT_B200_GEMM_FP32_SIMT = 0,   /* fp32 FMA on CUDA cores (bit-level closest to the reference) */
  RNNT_B200_GEMM_TC_FP16X3 = 1,   /* tcgen05, operands split hi+lo fp16, 3 MMAs, fp32 accumulate (~fp32 accuracy) */
  RNNT_B200_GEMM_TC_FP16 = 2      /* tcgen05, single fp16 pass (throughput mode, not token-exact; reserved) */
} rnnt_b200_gemm_mode;

/* Shape/config of the path.  Field names follow config/testing.yaml:120-229. */
typedef struct rnnt_b200_config {
  int32_t sample_rate;   /* sr: 16000 */
  int32_t n_fft;         /* melkwargs.n_fft: 1024 (only 1024 is implemented) */
  int32_t win_length;    /* int(win_length * sr) = 400, transforms.py:288 */
  int32_t hop_length;    /* int(hop_length * sr) = 160, transforms.py:289 */
  int32_t n_mels;        /* melkwargs.n_mels */
  int32_t n_stack;       /* StackDownsample.n_stack = 10 */
  int32_t downsample;    /* StackDownsample.downsample = 8 */
  int32_t enc_layers;    /* model.encoder.num_layers (LSTM) */
  int32_t pred_layers;   /* model.predictor.num_layers (NBRC == GRU) */
  int32_t hidden_sz;     /* model.hidden_sz == model.out_sz */
  int32_t embed_sz;      /* model.embed_sz */
  int32_t joint_sz;      /* model.joint_sz */
  int32_t vocab_sz;      /* model.vocab_sz */
  int32_t blank;         /* models.py:203,225 -> 0 */
  int32_t bos;           /* models.py:226-227 -> 2 */
  int32_t device;        /* CUDA device ordinal */
  int32_t gemm_mode;     /* rnnt_b200_gemm_mode */
  int32_t lm_layers;     /* lm.num_layers (config/testing.yaml:293-299,306-313); 0 = no fused LM (m.lm is None) */
  int32_t lm_hidden_sz;  /* lm.hidden_sz */
  int32_t lm_embed_sz;   /* lm.embed_sz (== lm_hidden_sz ties the output projection to the embedding, lm.py:27-29) */
  float log_offset;      /* transforms.py:312 -> 1e-6 */
  float ln_eps;          /* nn.LayerNorm default 1e-5 (models.py:84) */
  float bn_eps;          /* nn.BatchNorm1d default 1e-5 (custom_rnn.py:124) */
  float lm_alpha;        /* lm.py:13 ALPHA = 0.1 (the decode loops call fuse() with its defaults, models.py:431,558) */
  float lm_theta;        /* lm.py:14 THETA = 1.0 */
} rnnt_b200_config;

/* ---- lifecycle ---------------------------------------------------------------- */

int32_t rnnt_b200_abi_version(void);

/* Fills *cfg with the reference's shipped defaults (config/testing.yaml). */
int32_t rnnt_b200_default_config(rnnt_b200_config* cfg);

/* Replaces Transducer.__init__/from_config (models.py:190-259) as far as device
 * resources go.  Fails (ERR_CUDA) when no sm_100 device is present: there is no CPU
 * fallback. */
int32_t rnnt_b200_create(const rnnt_b200_config* cfg, rnnt_b200_handle* out);
int32_t rnnt_b200_destroy(rnnt_b200_handle h);

/* Message of the last failing call on this handle (never NULL).  h may be NULL for
 * create() failures. */
const char* rnnt_b200_last_error(rnnt_b200_handle h);

/* Replaces load_asr_model / nn.Module.load_state_dict (model_utils.py:61-95): one call
 * per tensor, `name` is the reference state_dict key verbatim (SURVEY.md section 8b:
 * "encoder.rnn_stack.rnns.0.weight_ih_l0", "predictor.rnn_stack.rnns.1.recurrent_kernel",
 * "joint.joint.2.bias", ...) plus two front-end tensors that the reference builds inside
 * torchaudio (transforms.py:290-296): "frontend.window" [win_length] and
 * "frontend.mel_fb" [n_fft/2+1, n_mels].  With cfg.lm_layers > 0 the fused language model's
 * state_dict (lm.py:20-29, loaded by load_lm, lm.py:86-101) is passed under the prefix "lm.":
 * "lm.embed.weight", "lm.rnn.weight_ih_l{k}", "lm.rnn.weight_hh_l{k}", "lm.rnn.bias_ih_l{k}",
 * "lm.rnn.bias_hh_l{k}", "lm.linear.weight", "lm.linear.bias".  `data` is HOST fp32, copied
 * before return.  Unknown names and wrong element counts return ERR_INVALID. */
int32_t rnnt_b200_set_weight(rnnt_b200_handle h, const char* name, const float* data_host, int64_t numel);

/* Repacks the weights into kernel layouts (gate interleave, transposes, BatchNorm
 * folded to scale/shift, embedding*ffn*kernel_0 table).  Synchronises `stream`.
 * ERR_STATE lists the first missing tensor. */
int32_t rnnt_b200_finalize(rnnt_b200_handle h, void* stream);

/* Pre-allocates workspaces for up to max_batch utterances of max_samples samples. */
int32_t rnnt_b200_reserve(rnnt_b200_handle h, int32_t max_batch, int64_t max_samples);

/* ---- shape helpers -------------------------------------------------------------- */

/* F = n/hop + 1 (torch.stft center=True), T = (F - n_stack)/downsample + 1 (unfold), 0 if F < n_stack. */
int64_t rnnt_b200_num_frames(rnnt_b200_handle h, int64_t n_samples);
int64_t rnnt_b200_num_steps(rnnt_b200_handle h, int64_t n_samples);

/* ---- features: a3 + a5 (+ a4 for streams) ---------------------------------------- */

/* Replaces TransformTime.encodes -> StackDownsample.encodes (transforms.py:301-323,
 * 436-441) for 16 kHz mono input.  audio_dev [B, n]; lens_dev [B] int32 valid samples
 * per utterance or NULL (= n); feats_dev [B, T, X] with T = num_steps(n); rows
 * t >= num_steps(lens[b]) are zero-filled. */
int32_t rnnt_b200_features(rnnt_b200_handle h, const float* audio_dev, const int32_t* lens_dev,
                           int32_t B, int64_t n, float* feats_dev, void* stream);

/* Replaces TransformTime.encodes alone (transforms.py:301-323): un-stacked log-mel frames
 * logmel_dev [B, F, n_mels], F = num_frames(n).  For callers that keep the reference's
 * transform-by-transform pipeline; the fused rnnt_b200_features() is the hot path. */
int32_t rnnt_b200_logmel(rnnt_b200_handle h, const float* audio_dev, const int32_t* lens_dev,
                         int32_t B, int64_t n, float* logmel_dev, void* stream);

/* Replaces the `stream` transform pipeline up to Buffer (config/testing.yaml:356-374;
 * transforms.py:301-342, 436-441): window_dev [B, W] is the serving loop's 3-chunk
 * window (api-server.py:26,95-102); feats_dev [B, X] is the single stacked row made of
 * frames F/3+1 .. F/3+n_stack of that window. */
int32_t rnnt_b200_features_stream(rnnt_b200_handle h, const float* window_dev, int32_t B, int64_t W,
                                  float* feats_dev, void* stream);

/* ---- resampling: a2 ------------------------------------------------------------------ */

/* Replaces Resample.encodes (transforms.py:135-144): torchaudio.transforms.Resample(orig_freq = orig_sr,
 * new_freq = cfg.sample_rate) with its defaults (windowed-sinc, lowpass_filter_width 6, rolloff 0.99) for B utterances.
 * audio_dev [B, n] -> out_dev [B, rnnt_b200_resample_len(n, orig_sr)]; orig_sr == cfg.sample_rate copies.
 * ChannelCut (transforms.py:122-132) needs no kernel: the caller passes the first channel. */
int64_t rnnt_b200_resample_len(rnnt_b200_handle h, int64_t n, int32_t orig_sr);
int32_t rnnt_b200_resample(rnnt_b200_handle h, const float* audio_dev, int32_t B, int64_t n, int32_t orig_sr,
                           float* out_dev, void* stream);

/* ---- encoder: a7 + a8 + a9 -------------------------------------------------------- */

/* Replaces Encoder.forward(x, state, lengths, return_state) (models.py:105-113) and
 * CustomRNN.forward (custom_rnn.py:177-232) in eval mode.  feats_dev [B, T, X];
 * lens_T_dev [B] int32 or NULL; state_h_dev / state_c_dev [L, B, H] or NULL.  When
 * use_state_in == 0 the learnable initial state `hs[i]` is used (custom_rnn.py:152-156);
 * when the state pointers are non-NULL the final (h, c) per layer is written back
 * (state at lens_T[b] for ragged batches, as pack_padded_sequence gives,
 * custom_rnn.py:164-170).  enc_out_dev [B, T, H]. */
int32_t rnnt_b200_encode(rnnt_b200_handle h, const float* feats_dev, const int32_t* lens_T_dev,
                         int32_t B, int32_t T, float* state_h_dev, float* state_c_dev,
                         int32_t use_state_in, float* enc_out_dev, void* stream);

/* ---- predictor: a10 + a11 ---------------------------------------------------------- */

/* Replaces Predictor.forward(x, state) for one step (models.py:181-187; cell
 * haste/nbrc.py:46-56).  tokens_dev [B] int32; state_h_dev [Lp, B, H] in/out (required);
 * use_state_in == 0 starts from `hs[i]`; out_dev [B, H]. */
int32_t rnnt_b200_predict(rnnt_b200_handle h, const int32_t* tokens_dev, int32_t B,
                          float* state_h_dev, int32_t use_state_in, float* out_dev, void* stream);

/* ---- joint: a12 --------------------------------------------------------------------- */

/* Replaces Joint.forward(h_pred, h_enc), concat method (models.py:132-140): raw logits,
 * no softmax.  h_pred_dev, h_enc_dev [B, H]; logits_dev [B, V]. */
int32_t rnnt_b200_joint(rnnt_b200_handle h, const float* h_pred_dev, const float* h_enc_dev,
                        int32_t B, float* logits_dev, void* stream);

/* ---- LM shallow fusion: a16 (LMFuser, lm.py:43-83) ---------------------------------------- */

/* With cfg.lm_layers > 0 every greedy decode (decode_greedy / transcribe*) fuses the language
 * model exactly where the reference does: after the blank test, the joint's log_softmax row is
 * standardised ((x - mean) / (std + 1e-5), utils.py:162-164), its blank entry pinned to -10, and
 * the emitted token is arg max(lm_alpha * lm_row + lm_theta * joint_row) (lm.py:56-79); after each
 * emitted token the LM advances one step and its standardised row replaces lm_row (lm.py:50-54);
 * until the first emission there is no lm_row and the joint's arg max stands (lm.py:58,79).
 * By default each call starts from a fresh fuser (models.py:401).  A streaming caller
 * (transcribe_stream keeps one fuser for the whole stream, models.py:478) registers a
 * caller-owned DEVICE blob that carries (h, c) of every LM layer, lm_row and its validity for
 * B streams between calls; a zero-filled blob is a fresh fuser (LMFuser.reset, lm.py:81-83).
 * The blob layout is private; size it with rnnt_b200_lm_state_bytes.  NULL unregisters. */
int32_t rnnt_b200_lm_state_bytes(rnnt_b200_handle h, int32_t B, int64_t* bytes_out);
int32_t rnnt_b200_set_lm_state(rnnt_b200_handle h, void* blob_dev, int32_t B);

/* ---- greedy decode: a13 / a14 inner loop ---------------------------------------------- */

/* Replaces the loops of Transducer.decode_greedy (models.py:403-443) and
 * Transducer.transcribe_stream (models.py:528-571) for B independent utterances/streams,
 * entirely on the device (no host round trip per symbol).
 *   enc_dev [B, T, H], lens_T_dev [B] int32 or NULL
 *   max_iters: 3 offline (models.py:369), 10 streaming (models.py:458)
 *   pred_state_h_dev [Lp, B, H], pred_out_dev [B, H]: predictor state and last predictor
 *     output (`h_t_pred`); in/out.  use_state_in == 0 -> initialised by feeding BOS from the
 *     learnable state (models.py:397-398, 484-489)
 *   tokens_out_dev [B, U_cap] int32, ntok_out_dev [B] int32 (U_cap >= max_iters*T)
 *   neg_logp_out_dev [B] double or NULL: -sum of the chosen log-probs incl. blanks (models.py:422,455)
 *   iters_out_dev [B, T] uint8 or NULL: joint evaluations per frame (models.py:443)
 *   trace_logp_dev [B, trace_cap, V] or NULL: log_softmax rows of the first trace_cap
 *     evaluations of every utterance (parity testing; `extra["outs"]`, models.py:419) */
int32_t rnnt_b200_decode_greedy(rnnt_b200_handle h, const float* enc_dev, const int32_t* lens_T_dev,
                                int32_t B, int32_t T, int32_t max_iters,
                                float* pred_state_h_dev, float* pred_out_dev, int32_t use_state_in,
                                int32_t* tokens_out_dev, int32_t U_cap, int32_t* ntok_out_dev,
                                double* neg_logp_out_dev, uint8_t* iters_out_dev,
                                float* trace_logp_dev, int32_t trace_cap, void* stream);

/* ---- whole path ------------------------------------------------------------------------- */

/* Replaces ASRServicer.Transcribe's compute (api-server.py:68-78: x_tfm -> model.transcribe)
 * for a batch: features -> encoder -> greedy decode with device-resident audio. */
int32_t rnnt_b200_transcribe(rnnt_b200_handle h, const float* audio_dev, const int32_t* lens_dev,
                             int32_t B, int64_t n, int32_t max_iters,
                             int32_t* tokens_out_dev, int32_t U_cap, int32_t* ntok_out_dev,
                             double* neg_logp_out_dev, uint8_t* iters_out_dev, void* stream);

/* Same with HOST buffers (pinned for full speed): copies audio host->device, runs the
 * path, copies tokens / counts / scores back and synchronises `stream` before
 * returning.  lens_host may be NULL.  This is the end-to-end call bench.py times. */
int32_t rnnt_b200_transcribe_host(rnnt_b200_handle h, const float* audio_host, const int32_t* lens_host,
                                  int32_t B, int64_t n, int32_t max_iters,
                                  int32_t* tokens_out_host, int32_t U_cap, int32_t* ntok_out_host,
                                  double* neg_logp_out_host, void* stream);

/* Two-deep pipeline over the same path, for throughput (the analogue of the reference's request thread pool,
 * api-server.py:138-139, which keeps several Transcribe calls in flight).  submit() queues one batch: `audio` / `lens` are
 * host buffers (pinned for full speed) when on_host != 0, device buffers otherwise, complete in `stream` order.  Its
 * host->device copy and front end run on an internal low-priority stream under the recurrent kernels of the batch
 * submitted before it (which occupy 128 of the 148 SMs); encoder and decode follow in submission order on an internal
 * high-priority stream and the results are copied to the HOST buffers given here.  collect() blocks until they are there.
 * slot is 0 or 1; a slot must be collected before it is submitted again, and the plain transcribe / stream_push calls
 * refuse to run while a slot is in flight (shared workspaces).  Results are identical to rnnt_b200_transcribe_host. */
int32_t rnnt_b200_pipeline_submit(rnnt_b200_handle h, const float* audio, int32_t on_host, const int32_t* lens,
                                  int32_t B, int64_t n, int32_t max_iters, int32_t slot,
                                  int32_t* tokens_host, int32_t U_cap, int32_t* ntok_host,
                                  double* neg_logp_host, void* stream);
int32_t rnnt_b200_pipeline_collect(rnnt_b200_handle h, int32_t slot);
/* Non-blocking probe: *front_done / *back_done become 1 when that half of the slot's batch has finished on the device. */
int32_t rnnt_b200_pipeline_query(rnnt_b200_handle h, int32_t slot, int32_t* front_done, int32_t* back_done);

/* ---- streaming sessions: a14 + the serving loop around it ------------------------------------- */

/* B concurrent streams, one chunk per (active) stream per push, each in its own phase: the reference's serving
 * loop (ASRServicer.TranscribeStream, api-server.py:82-135) -- sliding window of `n_window` chunks
 * (api-server.py:26,95-102), stream transforms incl. StreamPostprocess and Buffer(n_buffer)
 * (config/testing.yaml:356-374, transforms.py:326-342,455-471), Transducer.transcribe_stream with
 * carried encoder / predictor / LM-fuser state (models.py:457-577) -- as ONE call per chunk tick, with
 * all per-stream state (audio windows, pending feature rows, LSTM (h, c), GRU h, last predictor output,
 * LM fuser) owned by the session and resident in HBM.
 *   open:  chunk_samples = 1280 (80 ms, api-client.py:14), n_window = 3, n_buffer = 2, max_iters = 10 are the
 *          reference's values.  The session borrows the handle's workspaces: calls on one handle are serialised.
 *   push:  chunks [B, chunk_samples] fp32, DEVICE memory or (chunks_on_host != 0) HOST memory; active_host [B] bytes or
 *          NULL: streams with a zero byte are skipped this tick (no connection / no chunk: window, Buffer and state untouched).
 *          Every stream runs its own phase of the serving loop: its first n_window - 1 chunks only fill the window
 *          (api-server.py:97-98); afterwards every chunk yields one feature row and every n_buffer-th row the encoder +
 *          decode loop run for that stream (streams whose Buffer is not full take part with zero frames).  *advanced_out = 1
 *          when at least one stream ran the model; tokens_host_out [B, U_cap] (U_cap >= max_iters * n_buffer) / ntok_host_out
 *          [B] then hold the tokens each stream emitted in this tick (host memory; ntok = -1 marks a stream that did not run
 *          the model in this tick, 0 one that ran and emitted nothing; the call synchronises the stream).  Otherwise *advanced_out = 0 and the outputs are untouched.
 *   reset: stream `slot` (or all streams when slot = -1) back to the state of a fresh connection: the `reset` closure of
 *          transcribe_stream (models.py:494-500: learnable encoder state, predictor fed BOS, LMFuser.reset) plus an empty
 *          window and Buffer.  Streams may be reset at any tick; the others are not disturbed. */
typedef struct rnnt_b200_stream_s* rnnt_b200_stream;
int32_t rnnt_b200_stream_open(rnnt_b200_handle h, int32_t n_streams, int32_t chunk_samples, int32_t n_window,
                              int32_t n_buffer, int32_t max_iters, rnnt_b200_stream* out);
/* Buffer lifetime: a HOST `chunks` buffer (chunks_on_host = 1) may be reused as soon as the call returns -- every return path
 * has waited for the copy out of it.  rnnt_b200_destroy() refuses (ERR_STATE) while sessions opened on the handle exist. */
int32_t rnnt_b200_stream_push(rnnt_b200_stream s, const float* chunks, int32_t chunks_on_host, const uint8_t* active_host,
                              int32_t* tokens_host_out, int32_t U_cap, int32_t* ntok_host_out,
                              int32_t* advanced_out, void* stream);
int32_t rnnt_b200_stream_reset(rnnt_b200_stream s, int32_t slot);
/* Mid-stream reset = the `reset_fn` that Transducer.transcribe_stream yields (libreasr/lib/models.py:480-500), called by
 * the server's silence logic (api-server.py:133-135): encoder state -> learnable initial state, predictor -> BOS
 * state, LM fuser cleared.  The 3-chunk audio window and the Buffer of the serving loop (api-server.py:83-115,
 * transforms.py:455-471) are NOT touched (rnnt_b200_stream_reset above = new connection clears them too).
 * slot = -1: every stream. */
int32_t rnnt_b200_stream_reset_state(rnnt_b200_stream s, int32_t slot);
int32_t rnnt_b200_stream_close(rnnt_b200_stream s);


/* ---- beam search ------------------------------------------------------------------------------- */

/* Batched RNN-T beam search over encoder output (BASELINE.json configs[3], [4]).  NOT in the reference
 * (libreasr/lib/models.py has greedy decoding only; :8 is an unused PriorityQueue import): the algorithm is the
 * breadth-first, iteration-capped search defined by oracle/beam.py on the reference's Predictor / Joint
 * (models.py:116-187), keeping decode_greedy's max_iters rule (models.py:369).  enc [B,T,H] device, lens_T [B] or NULL;
 * width in [1, 8]; writes the best hypothesis of every utterance: tokens_out [B,U_cap] (U_cap >= max_iters*T),
 * ntok_out [B], score_out [B] (fp64 log-probability, nullable) -- all device pointers.  gemm_mode 1 only. */
int32_t rnnt_b200_decode_beam(rnnt_b200_handle h, const float* enc, const int32_t* lens_T, int32_t B, int32_t T, int32_t width,
                              int32_t max_iters, int32_t* tokens_out, int32_t U_cap, int32_t* ntok_out, double* score_out,
                              void* stream);


/* ---- training-time forward: joint lattice + RNN-T loss (eval mode, no gradients) ---------------------- */

/* Transducer.forward (libreasr/lib/models.py:308-359) in eval mode and the loss get_loss_func("rnnt") computes from its
 * output (libreasr/lib/loss.py:72-110 -> warp_rnnt.rnnt_loss(..., average_frames=False); restated from its definition,
 * oracle/rnnt_loss.py).  feats [N,T,X] (what the transform pipeline yields), lens_T [N] encoder steps or NULL,
 * labels [N,Umax] int32 (padded), label_lens [N].  The predictor is teacher-forced over cat(bos, labels) (U = Umax+1
 * positions).  loss_out [N] fp64 = -log p(labels | audio) per sequence; lattice_out (nullable) [N,T,U,V] = the
 * log_softmax lattice forward() returns (positions beyond (lens_T, label_lens) hold the values of the padded
 * computation).  All pointers device; N <= 256; gemm_mode 1. */
int32_t rnnt_b200_forward_loss(rnnt_b200_handle h, const float* feats, const int32_t* lens_T, const int32_t* labels,
                               const int32_t* label_lens, int32_t N, int32_t T, int32_t Umax, float* lattice_out,
                               double* loss_out, void* stream);
/* The loss alone from a given log-probability lattice [N,T,U,V] (what the reference's _loss_func receives as `inp`). */
int32_t rnnt_b200_rnnt_loss(rnnt_b200_handle h, const float* lattice, const int32_t* lens_T, const int32_t* labels,
                            const int32_t* label_lens, int32_t N, int32_t T, int32_t U, double* loss_out, void* stream);

/* ---- self test -------------------------------------------------------------------------------- */

/* Runs the library's own GEMM (the contraction behind the LSTM gate and joint projections)
 * on caller data: C[M,N] = A[M,K] * W[N,K]^T + bias[N] with the arithmetic of `gemm_mode`.
 * Test hook (no reference counterpart): lets the parity tests measure each arithmetic mode
 * against an fp64 product.  K and N must be multiples of 4.  Synchronises in the TC modes. */
int32_t rnnt_b200_selftest_gemm(rnnt_b200_handle h, const float* A_dev, const float* W_dev, const float* bias_dev,
                                float* C_dev, int64_t M, int32_t N, int32_t K, int32_t gemm_mode, void* stream);

/* ---- introspection ------------------------------------------------------------------------ */

/* Number of kernels this handle has launched so far (bench.py's gpu_launches). */
int64_t rnnt_b200_kernel_launches(rnnt_b200_handle h);
/* Launches of the fp32 cooperative decode kernel so far (gemm_mode 0, or a gemm_mode 1 call that could not take a tcgen05
 * kernel: a registered LM state blob with more than 64 streams).  Tests assert that the tensor-core path was the one that ran. */
int64_t rnnt_b200_fp32_decode_launches(rnnt_b200_handle h);

/* Device time (ms) the most recent transcribe*() spent per stage, measured with CUDA
 * events on `stream`: out[0]=features, [1]=encoder total (LayerNorm + hoisted input GEMMs +
 * recurrent kernels), [2]=the hoisted input GEMMs' share of [1] (incl. operand-image conversion),
 * [3]=joint enc projection GEMM, [4]=decode loop.
 * Synchronises.  Profiling must have been enabled with rnnt_b200_set_profiling(h, 1). */
int32_t rnnt_b200_set_profiling(rnnt_b200_handle h, int32_t enable);
int32_t rnnt_b200_stage_times_ms(rnnt_b200_handle h, float* out5_host);

#ifdef __cplusplus
}
#endif
#endif /* RNNT_B200_H_ */
