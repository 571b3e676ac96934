/*
 * ldiff.h — C ABI of libldiff_sm100.so, the B200 (sm_100a) implementation of the
 * L-Diffusion sampling-and-feature hot path.
 *
 * The reference (Lweihan/LDiffusion) is pure Python and has no FFI of its own:
 * its seams are Python call signatures.  Each entry point below replaces the
 * inline eager-tensor code at the cited reference lines (paths relative to the
 * reference checkout); INTEGRATION.md shows the ctypes stub a maintainer adds.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless named host_*;
 *  - `stream` is a cudaStream_t passed as void*; every call only enqueues work
 *    on it and returns (no allocation, no synchronisation; the only process-wide state is the ldiff_tune knobs
 *    below, per-device launch-attribute bookkeeping - so one process may drive several devices - and five
 *    debugging switches that are read from the environment ONCE, at the first launch that consults them:
 *    LDIFF_HEAD_TMA=0 (register-staged operand load of the tcgen05 head), LDIFF_CELL_TC=0 (CUDA-core cell
 *    classifier: fp32 summation order, so its logits differ from the tensor-core form in the last bits),
 *    LDIFF_PAINT_SMEM_LUT, LDIFF_LIFT_BAND, LDIFF_CONF_BPS (launch shapes; results unchanged));
 *  - `dtype` is LDIFF_F32 / LDIFF_BF16 / LDIFF_U8: the STORAGE type of the
 *    floating tensors; arithmetic is always fp32, in the order documented in
 *    DESIGN.md (bit-exact against oracle/ in fp32);
 *  - return value: 0 on success, a negative LDIFF_E* code otherwise
 *    (ldiff_strerror gives the text).  Argument errors that the reference
 *    raises as ValueError stay in the Python host layer;
 *  - data errors that only the device can see (a predicted label >= K, an
 *    instance id outside the LUT) are reported through an `int* status` word in
 *    device memory that the kernel ORs bits into; the host reads it when it
 *    reads the result.
 */
#ifndef LDIFF_H_
#define LDIFF_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDIFF_ABI_VERSION 2

enum { LDIFF_F32 = 0, LDIFF_BF16 = 1, LDIFF_U8 = 2 };

enum {
  LDIFF_OK = 0,
  LDIFF_EINVAL = -1,       /* bad argument value */
  LDIFF_EALIGN = -2,       /* pointer not aligned as required (16 bytes) */
  LDIFF_ELAUNCH = -3,      /* CUDA reported an error at launch */
  LDIFF_EUNSUPPORTED = -4  /* valid but unsupported combination */
};

/* bits ORed into *status by kernels */
enum { LDIFF_STATUS_PRED_RANGE = 1, LDIFF_STATUS_INST_RANGE = 2, LDIFF_STATUS_SW_INF = 4,
       LDIFF_STATUS_XCHG_TIMEOUT = 8, LDIFF_STATUS_LABEL_RANGE = 16 };

/* Scheduling / variant knobs (no effect on results, except LDIFF_TUNE_PHILOX_ROUNDS).  A value set here applies to
 * the launches issued after the call.  Before the first ldiff_tune of a knob its environment variable, else the default, applies.
 *  LDIFF_TUNE_ARGMAX_VARIANT (env LDIFF_ARGMAX_VARIANT, default 0): ldiff_lift_argmax runs 0 = the envelope
 *    kernel - its row form (a thread sweeps an output row; x32 horizontal lifts with a vertical lift of x29 or more
 *    and a 16-byte aligned mask) where the shape allows it, else its column form -, 1 = the column form always,
 *    4 / 5 = round 1's per-pixel evaluation kernel with 2 / 1 columns per thread (A/B timing);
 *  LDIFF_TUNE_DECODE_TAIL_SMS (env LDIFF_DT_SMS, default 0 = all): the register-staged decode tail sizes its
 *    one-wave grid for that many SMs;
 *  LDIFF_TUNE_DECODE_TAIL_TMA (env LDIFF_DT_TMA, default 6): 0 = register-staged decode tail, 1..6 = the
 *    bulk-TMA staged persistent kernel with (stages x CTAs/SM) = (4x2) / (3x3) / (2x4) / (2x3) / (3x2) / (2x2) for bf16,
 *    7..10 = (2x3) / (3x2) / (2x2) / (4x2) with a dedicated producer warp, 11..13 = (2x1) / (3x1) / (4x1): one CTA per
 *    SM - the slowest alone and the best throughput when three or more passes are in flight (each pass's tails then
 *    leave room for the other passes' kernels; HotPathRing selects 11 for its slots);
 *  LDIFF_TUNE_PHILOX_ROUNDS (env LDIFF_PHILOX_ROUNDS, default 0 = 10): rounds of the Philox4x32 stream behind the
 *    Laplace noise - THIS knob changes the random stream (not the distribution): 7 = the smallest round count that
 *    passes BigCrush in the Philox paper, anything else = the published default of 10. */
enum { LDIFF_TUNE_ARGMAX_VARIANT = 0, LDIFF_TUNE_DECODE_TAIL_SMS = 1, LDIFF_TUNE_DECODE_TAIL_TMA = 2,
       LDIFF_TUNE_PHILOX_ROUNDS = 3, LDIFF_TUNE_COUNT = 4 };
int ldiff_tune(int knob, int value);
/* the value in effect for `knob` (>= 0), or LDIFF_EINVAL */
int ldiff_tune_get(int knob);

int ldiff_abi_version(void);
const char* ldiff_strerror(int code);
/* number of kernels this library has launched in the calling process */
int64_t ldiff_launch_count(void);

/* ---- a-1  Laplace forward noising ---------------------------------------
 * replaces ldiffusion.py:234-237 (and the sampler inside
 * torch.distributions.Laplace.rsample):  out = x + noise,
 * noise = 0 - (b * sign(u)) * log1p(-|u|).
 * Exactly one source of randomness is used:
 *   noise_in != NULL : injected noise (parity mode; out = x + noise_in, one add)
 *   u_in     != NULL : injected uniforms in (-1,1)
 *   otherwise        : Philox4x32-10 (LDIFF_TUNE_PHILOX_ROUNDS = 7: 7 rounds), key = seed.  fp32 storage: element i
 *                      takes word i%4 of counter offset + i/4; bits [22:0] of the word are |u| (23 bits), bit 31 the
 *                      sign.  bf16 storage: element i takes half-word i%8 of counter offset + i/8 (sign + 15-bit
 *                      magnitude; the top magnitude cell is refined by a second 23-bit draw, so the exponential tail
 *                      is not truncated).  Advancing offset by ceil(n/4) between calls never reuses a counter in
 *                      either form.  Both streams are restated in oracle/laplace.py
 * noise_out (optional) receives the fp32/bf16 noise that was added. */
int ldiff_laplace_qsample(const void* x, void* out, const void* noise_in, const void* u_in,
                          void* noise_out, float b, uint64_t seed, uint64_t offset,
                          int64_t n, int dtype, void* stream);

/* a-1, multimodal variant: Laplace(0,1) noise modulated by a per-pixel scale map, and its
 * inverse.  Replaces segmentor.py:339,344-345
 *     latents = vae.encode(rgb_i).latent_dist.sample() * 0.18215
 *     noise = Laplace(0.0, 1.0).sample(latents.shape);  latents_noisy = latents + noise * depth_resized
 * and segmentor.py:375,379
 *     latents_denoised = latents_noisy - noise_pred * depth_resized;  vae.decode(latents_denoised / 0.18215)
 *   ldiff_laplace_qsample_map : out = fl(x_mul * x) + fl(noise * s)      (x_mul = 1: exact, no scaling)
 *   ldiff_scaled_residual     : out = fl(fl(x - fl(eps * s)) / out_div)  (out_div = 1: exact)
 * x, eps, out, noise_*: [B, channels, plane] flattened to n elements.  scale: scale_channels ==
 * channels -> same shape; scale_channels == 1 -> [B, 1, plane], broadcast over the channels (what
 * depth_resized.repeat(1, C, 1, 1) materialises at segmentor.py:341).  Randomness as in
 * ldiff_laplace_qsample (noise_in / u_in / Philox); noise_out receives the UNIT Laplace noise. */
int ldiff_laplace_qsample_map(const void* x, const void* scale, void* out, const void* noise_in,
                              const void* u_in, void* noise_out, float x_mul, uint64_t seed,
                              uint64_t offset, int64_t n, int64_t plane, int channels,
                              int scale_channels, int dtype, void* stream);
int ldiff_scaled_residual(const void* x, const void* eps, const void* scale, void* out, float out_div,
                          int64_t n, int64_t plane, int channels, int scale_channels, int dtype,
                          void* stream);

/* ---- a-2  PLMS reverse step ----------------------------------------------
 * replaces scheduler.step(...).prev_sample at segmentor.py:102-104, :443-445,
 * :525-527, utils.py:200-202, pixel_latent_vector.py:77-79, sample.py:62-64
 * (diffusers PNDMScheduler.step_plms + _get_prev_sample).
 *   mode 0: eh = e0
 *        1: eh = (e0 + e1) / 2
 *        2: eh = (3 e0 - e1) / 2
 *        3: eh = (23 e0 - 16 e1 + 5 e2) / 12
 *        4: eh = (1/24) (55 e0 - 59 e1 + 37 e2 - 9 e3)
 *   prev = sample_coeff * sample - (alpha_diff * eh) / denom
 * e0 is the newest model output, e1..e3 older ones; the three scalars are
 * computed on the host in fp32 exactly as the reference's 0-dim tensors are. */
int ldiff_plms_step(const void* sample, const void* e0, const void* e1, const void* e2,
                    const void* e3, int mode, float sample_coeff, float alpha_diff, float denom,
                    void* prev_sample, int64_t n, int dtype, void* stream);

/* a-2 + a-1 fused ("step_then_noise"): ONE launch that performs the reverse update of
 * ldiff_plms_step on (sample, e0..e3) AND the forward noising of ldiff_laplace_qsample on `clean`
 * (noisy = clean + Laplace(0, b)); the reference does both on latent-sized tensors inside the same
 * loop iteration (segmentor.py:100-104 + ldiffusion.py:233-237).  Arguments as in the two entry
 * points above; every output bit equals what the two separate launches produce. */
int ldiff_plms_step_noise(const void* sample, const void* e0, const void* e1, const void* e2,
                          const void* e3, int mode, float sample_coeff, float alpha_diff, float denom,
                          void* prev_sample, const void* clean, void* noisy, const void* noise_in,
                          const void* u_in, float b, uint64_t seed, uint64_t offset, int64_t n,
                          int dtype, void* stream);

/* ---- a-3  decode tail -> uint8 RGB + PIL gray ----------------------------
 * replaces diffusers decode_latents' tail + numpy_to_pil + PIL convert("L") +
 * the per-pixel stacking loop (pixel_latent_vector.py:80-93,
 * segmentor.py:105-108,446-448,528-530, utils.py:203-205).
 * img: [B,3,H,W] planar fp32/bf16.  rgb_hwc (optional): uint8 [B,H,W,3].
 * gray (optional): plane b is written at gray + b*gray_batch_stride, H*W bytes
 * (slot i of a [B,n+1,H,W] pixel-vector tensor: the concat is free). */
int ldiff_decode_tail_gray(const void* img, uint8_t* rgb_hwc, uint8_t* gray, int B, int H, int W,
                           int64_t gray_batch_stride, int dtype, void* stream);

/* same pass with the segmentor's model input fused in (replaces the PIL ->
 * transforms.ToTensor -> Normalize hand-off at segmentor.py:107-108,533-534):
 * model_input fp32 [B,3,H,W] = ((q / 255) - mean[c]) / std[c] on the quantised image q.
 * host_mean3 / host_std3 are HOST pointers to three floats.  rgb_hwc and gray optional. */
int ldiff_decode_tail_model_input(const void* img, uint8_t* rgb_hwc, uint8_t* gray, float* model_input,
                                  const float* host_mean3, const float* host_std3, int B, int H, int W,
                                  int64_t gray_batch_stride, int dtype, void* stream);

/* the same pass with the per-step consumers of the decoder output fused in — what the tail already
 * streams is not fetched again by separate launches.  Requires H % 16 == 0, W % 16 == 0, gray != NULL; every
 * other pointer is optional (NULL = off), at least one must be given:
 *   feat        channel feat_channel of [B, feat_ctot, H/16, W/16] (feat_dtype = dtype or LDIFF_F32) =
 *               weighted gray of F.interpolate(img, (H/16, W/16), bilinear): the per-step body of
 *               ldiffusion.py:240-247 (lift + gray + torch.cat), bit-identical to ldiff_bilinear_lift(gray=1);
 *   small_rgb   [B,3,H/16,W/16] (image dtype): that lift before the gray (the source of ldiffusion.py:251);
 *   label_plane the uint8 [B,H,W] ground truth `label` copied to label_plane + b*label_plane_stride (the
 *               label slot of the pixel vectors, pixel_latent_vector.py:92 — replaces ldiff_copy_planes_u8);
 *   label_small uint8 [B,1,H/16,W/16] = trunc(bilinear(label)) (ldiffusion.py:224-226). */
int ldiff_decode_tail_fused(const void* img, uint8_t* rgb_hwc, uint8_t* gray, int B, int H, int W,
                            int64_t gray_batch_stride, int dtype, void* feat, int feat_dtype, int feat_ctot,
                            int feat_channel, void* small_rgb, const uint8_t* label, uint8_t* label_plane,
                            int64_t label_plane_stride, uint8_t* label_small, void* stream);

/* ---- a-4  bilinear lift + gray + concat ----------------------------------
 * replaces F.interpolate(mode='bilinear', align_corners=False) (+ weighted gray
 * + torch.cat) at ldiffusion.py:224-226, :240-251 (same primitive at
 * conductor.py:135).  Source [B,C,h,w] planar (strides in elements); result is
 * written into channels [dst_channel, dst_channel + (gray ? 1 : C)) of a
 * [B,Ctot,H,W] planar destination.  gray != 0 requires C == 3 and applies
 * 0.2989 R + 0.5870 G + 0.1140 B after the lift.  A u8 destination truncates. */
int ldiff_bilinear_lift(const void* src, int src_dtype, int C, int h, int w,
                        int64_t src_batch_stride, int64_t src_channel_stride,
                        void* dst, int dst_dtype, int Ctot, int dst_channel, int H, int W,
                        int B, int gray, void* stream);

/* n_src (<= 8) same-shaped sources in ONE launch; source i lands at channel
 * dst_channel + i * (gray ? 1 : C): the per-step lift + gray + torch.cat loop of
 * ldiffusion.py:240-247 as a single gather.  host_srcs: HOST array of device pointers. */
int ldiff_bilinear_lift_multi(const void* const* host_srcs, int n_src, int src_dtype, int C, int h, int w,
                              int64_t src_batch_stride, int64_t src_channel_stride, void* dst,
                              int dst_dtype, int Ctot, int dst_channel, int H, int W, int B, int gray,
                              void* stream);

/* adjoint of ldiff_bilinear_lift for the training caller (ldiffusion.py:240-252; round-2 widening, reached only
 * through autograd): grad_src[b,c,y,x] += w_c * ly * lx * grad_out[b, dst_channel (+c), oy, ox] with the
 * forward's taps (h x w = the SOURCE size, H x W = the lifted size); fp32; grad_src (element strides as the
 * forward's source) must be zeroed by the caller and is accumulated into with atomics. */
int ldiff_bilinear_lift_backward(const float* grad_out, int Ctot, int dst_channel, int H, int W,
                                 float* grad_src, int C, int h, int w, int64_t src_batch_stride,
                                 int64_t src_channel_stride, int B, int gray, void* stream);

/* ---- a-5  classifier head + argmax ---------------------------------------
 * tissue form, replaces conductor.py:127 (1x1 conv 256->K), :135 (bilinear lift
 * to the input size) and segmentor.py:536 (argmax(softmax)) without ever
 * materialising full-resolution logits. */
/* feat [B,Cin,h*w] planar fp32/bf16, weight [K,Cin] same dtype, bias fp32 [K]
 * (or NULL) -> logits fp32 [B,K,h*w].  bf16 runs on tcgen05 tensor cores.
 * clear_i64 / n_clear (optional, NULL / 0): int64 counters this kernel zeroes as a side job — the
 * confusion matrix that ldiff_lift_argmax_hist, enqueued behind it on the same stream, accumulates
 * into — so that a pass needs no memset. */
int ldiff_head_logits(const void* feat, const void* weight, const float* bias, float* logits,
                      int B, int Cin, int K, int hw, int dtype, int64_t* clear_i64, int n_clear,
                      void* stream);
/* logits fp32 [B,K,h,w] -> uint8 mask [B,H,W] */
int ldiff_lift_argmax(const float* logits, uint8_t* mask, int B, int K, int h, int w, int H, int W,
                      void* stream);
/* the same with the a-6 histogram fused in: C[(K+1),K] += counts of (gt, mask) over the whole batch while
 * the mask bytes are still in registers (gt uint8 [B,H,W]; semantics of ldiff_confusion_hist, no gt LUT).
 * xchg != NULL: the kernel's last block also pushes the finished matrix into every rank's peer window, as
 * ldiff_confusion_hist_push does (channel = the window's channel).  K <= 15 and H >= 4h (the envelope
 * kernel's domain); otherwise LDIFF_EUNSUPPORTED: use ldiff_lift_argmax + ldiff_confusion_hist. */
int ldiff_lift_argmax_hist(const float* logits, uint8_t* mask, const uint8_t* gt, int64_t* C, int B, int K,
                           int h, int w, int H, int W, void* xchg, int channel, int* status, void* stream);
/* cell form, replaces conductor.py:218-221: per-instance Linear(Cin,K) ->
 * softmax[:,1:] -> top-1 (+1); writes lut[b*lut_stride + inst_ids[i]] = class.
 * inst_feats [B,n_per_image,Cin] row-major fp32/bf16 (Cin % 8 == 0); inst_ids int32
 * [n_per_image] (shared by the images of the batch); lut uint8 [B,lut_size]. */
int ldiff_cell_classify(const void* inst_feats, const void* weight, const float* bias,
                        const int32_t* inst_ids, uint8_t* lut, int lut_size, int64_t lut_stride,
                        float* logits_out, int n_per_image, int B, int Cin, int K, int dtype,
                        int64_t* clear_i64, int n_clear /* as ldiff_head_logits */, int* status, void* stream);
/* replaces the painting loop conductor.py:224-231 (+ segmentor.py:536):
 * mask[b,p] = lut[b*lut_stride + inst[b,p]]; ids outside [0,lut_size) -> 0 and
 * LDIFF_STATUS_INST_RANGE. */
int ldiff_lut_paint(const int32_t* inst, const uint8_t* lut, uint8_t* mask, int64_t n_per_image,
                    int B, int lut_size, int64_t lut_stride, int* status, void* stream);
/* the same for the uint16 label image Cellpose's `eval` returns below 65 536 labels (conductor.py:180: the
 * reference compares that array as it is, `masks == inst_id`, :193,227): no widening copy on the host, half the
 * bytes over the host link and out of HBM (3 B/pixel instead of 5). */
int ldiff_lut_paint_u16(const uint16_t* inst, const uint8_t* lut, uint8_t* mask, int64_t n_per_image,
                        int B, int lut_size, int64_t lut_stride, int* status, void* stream);
/* painting + the a-6 histogram in one pass (6 B/pixel instead of 5 + 2): C[(K+1),K] += counts of
 * (gt, mask) over the whole batch; xchg / channel as in ldiff_lift_argmax_hist.  K <= 15,
 * n_per_image % 16 == 0, 16-byte aligned planes; otherwise use ldiff_lut_paint + ldiff_confusion_hist. */
int ldiff_lut_paint_hist(const int32_t* inst, const uint8_t* lut, uint8_t* mask, const uint8_t* gt,
                         int64_t* C, int64_t n_per_image, int B, int lut_size, int64_t lut_stride, int K,
                         void* xchg, int channel, int* status, void* stream);
/* ... with uint16 instance ids (see ldiff_lut_paint_u16): 4 B/pixel. */
int ldiff_lut_paint_hist_u16(const uint16_t* inst, const uint8_t* lut, uint8_t* mask, const uint8_t* gt,
                             int64_t* C, int64_t n_per_image, int B, int lut_size, int64_t lut_stride, int K,
                             void* xchg, int channel, int* status, void* stream);
/* first-maximum argmax over the channel axis of [B,K,HW] (the argmax taken
 * inside utils.py:56, :85, evaluate.py:12, :30 on one-hot / logit inputs). */
int ldiff_argmax_channels(const void* x, uint8_t* out, int B, int K, int64_t hw, int dtype,
                          void* stream);

/* ---- a-6  confusion matrix -----------------------------------------------
 * replaces the K^2+8K masked-sum passes of utils.py:55-104 and
 * evaluate.py:11-45.  C is int64 [(K+1),K], row = gt class (row K = gt outside
 * [0,K)), column = predicted class; the kernel ACCUMULATES into C (caller
 * zeroes it; it may be the buffer handed to ncclAllReduce).  gt_lut (optional,
 * 256 bytes) maps raw gray levels to classes first (dataset.py:10-32,48-63).
 * pred >= K sets LDIFF_STATUS_PRED_RANGE (the reference's one_hot raises). */
int ldiff_confusion_hist(const uint8_t* pred, const uint8_t* gt, const uint8_t* gt_lut,
                         int64_t* C, int64_t n, int K, int* status, void* stream);
/* batched form for per-image metrics (evaluate.py:60-102 averages per-image scores):
 * image i = pixels [i*n_per_image, (i+1)*n_per_image), matrix i at C + i*(K+1)*K. */
int ldiff_confusion_hist_batched(const uint8_t* pred, const uint8_t* gt, const uint8_t* gt_lut,
                                 int64_t* C, int64_t n_per_image, int n_images, int K, int* status,
                                 void* stream);

/* ---- a-6 across GPUs: histogram fused with its all-reduce over NVLink peer memory -------------
 * Replaces the only cross-rank exchange the reference has for these counts, nnU-Net's three
 * pickled all_gather_object calls (nnUNetTrainer.py:1004-1012), and a separate NCCL all-reduce
 * behind the histogram.  One process per GPU; each rank owns a small "window" in device memory
 * that every peer maps through CUDA IPC.
 *
 *   ldiff_xchg_create        allocate + zero this rank's window (world <= 16, channels <= 4,
 *                            n_i64 = (K+1)*K counters per channel)
 *   ldiff_xchg_ipc_handle    64-byte CUDA IPC handle of the window (exchange it with any host
 *                            transport, e.g. torch.distributed.all_gather_object)
 *   ldiff_xchg_connect_ipc   handles = world * 64 bytes in rank order (own entry ignored)
 *   ldiff_xchg_connect_local same-process windows (several "ranks" on one GPU: tests)
 *   ldiff_confusion_hist_push  ldiff_confusion_hist whose last block stores the finished matrix C
 *                            into row `rank`, channel `channel` of EVERY rank's window: plain 8-byte
 *                            st.relaxed.sys stores, each word = half a counter + the step number (data
 *                            and "has landed" in one store; no system fence, no separate flag).  The
 *                            fused producers ldiff_lut_paint_hist / ldiff_lift_argmax_hist carry the same tail
 *   ldiff_xchg_reduce        one block: the j-th reduce of a rank waits for the j-th push of
 *                            every rank and channel, then writes the sum of the world rows to
 *                            out [channels][n_i64]
 *   ldiff_xchg_destroy       unmap peers, free the window (synchronise all ranks first)
 *
 * Ordering contract: every rank pushes every channel once per step and reduces once per step;
 * pushes and reduces are matched by count, so a reduce may be enqueued right behind its pushes
 * or a whole step later, concurrently with the next step's work (it then never waits).  Rows
 * live in a ring of 4 slots: push j of a rank must be stream-ordered after that rank's reduce
 * j-2 (at most two steps outstanding).  Integer sums: bit-exact at any world size.  A rank that
 * never arrives trips a device-side timeout (10 s, ldiff_xchg_set_timeout) that sets
 * LDIFF_STATUS_XCHG_TIMEOUT, not a hang; while that bit is set later reduces do not wait again.
 * Both kernels are plain launches and can be captured into CUDA graphs. */
int ldiff_xchg_create(int world, int rank, int channels, int n_i64, void** handle);
int ldiff_xchg_ipc_handle(void* handle, void* out64);
int ldiff_xchg_connect_ipc(void* handle, const void* handles);
int ldiff_xchg_connect_local(void* handle, void* const* peer_handles);
int ldiff_xchg_set_timeout(void* handle, int64_t timeout_ms);
int ldiff_xchg_destroy(void* handle);
int ldiff_confusion_hist_push(const uint8_t* pred, const uint8_t* gt, const uint8_t* gt_lut,
                              int64_t* C, int64_t n, int K, void* xchg, int channel, int* status,
                              void* stream);
/* stand-alone push of all channels at once (C int64 [channels][n_i64], e.g. the totals of a whole evaluation:
 * the reference sums once per evaluation); counts as one push of every channel */
int ldiff_xchg_push(void* xchg, const int64_t* C, void* stream);
int ldiff_xchg_reduce(void* xchg, int64_t* out, int* status, void* stream);
/* B planes of n bytes -> a strided slot (the label plane of the pixel vectors,
 * pixel_latent_vector.py:92); n and dst_stride multiples of 16 */
int ldiff_copy_planes_u8(const uint8_t* src, uint8_t* dst, int64_t n, int B, int64_t dst_stride,
                         void* stream);
/* int64 labels -> uint8 (values outside [0,254] become 255 = "other") */
int ldiff_labels_to_u8(const int64_t* in, uint8_t* out, int64_t n, void* stream);

/* ---- widening N2: nnU-Net sliding-window accumulate / mirror-TTA merge / export ----------
 * (vendored nnU-Net: predict_from_raw_data.py:530-589, label_handling.py:128-173).  All
 * tensors are fp16 as in the reference (its results arrays are torch.half); every reference
 * op's rounding to half is reproduced. */
/* acc[:, y0:y0+th, x0:x0+tw] += pred * gauss ; npred[y0:.., x0:..] += gauss  (gauss NULL: 1) */
int ldiff_sw_accumulate(const void* pred_f16, const void* gauss_f16, void* acc_f16, void* npred_f16,
                        int K, int th, int tw, int H, int W, int y0, int x0, void* stream);
/* out = (preds[0] + sum_j flip(preds[j], flips[j])) / n, adds in order, each rounded to half.
 * host_preds / host_flips are HOST arrays (n <= 8); flip bit 0 = rows, bit 1 = columns. */
int ldiff_sw_tta_merge(const void* const* host_preds_f16, const int* host_flips, int n, void* out_f16,
                       int K, int th, int tw, void* stream);
/* logits = acc / npred (half); seg = argmax(softmax(logits.float(), 0), 0) as uint8.
 * logits_out optional.  An infinite logit sets LDIFF_STATUS_SW_INF (the reference raises). */
int ldiff_sw_finalize_argmax(const void* acc_f16, const void* npred_f16, uint8_t* seg, void* logits_out_f16,
                             int K, int64_t hw, int* status, void* stream);

/* ---- widening N3: pixel-contrastive InfoNCE (model/loss.py:44-109) -------------------------
 * feat fp32 [B,C,hw] (C <= 16: one channel per sampling step).  Triple p = (pair_batch[p],
 * pair_anchor[p], pair_pos[p], pair_neg[p, 0..n_neg)) of pixel indices (int32, device).
 * forward: loss_per_pair[p] = cross_entropy([a.pos, a.neg_1 ..] / T, target 0); lse_per_pair
 * keeps the log-sum-exp for the backward.  The reference's result is mean(loss_per_pair). */
int ldiff_infonce_forward(const float* feat, const int* pair_batch, const int* pair_anchor,
                          const int* pair_pos, const int* pair_neg, float* loss_per_pair,
                          float* lse_per_pair, int C, int64_t hw, int n_neg, int n_pairs,
                          float temperature, void* stream);
/* The sampling of loss.py:64-87 on the device: labels uint8 [B,hw] (values < 32, hw <= 16384).  Per
 * image and per class (ascending) with more than one pixel and more than n_neg pixels outside it:
 * max(1, count/100) anchors = a keyed random permutation's prefix over the class, the positive =
 * a uniform draw among the class's other pixels, n_neg distinct negatives = a keyed random
 * permutation's prefix over the pixels outside the class.  Image b owns pair slots
 * [b*cap, (b+1)*cap) (cap >= hw/100 + 32 holds every case); unused slots are filled with -1, which
 * forward/backward skip (loss 0).  n_valid[b] = triples of image b.  Deterministic in (seed, offset).
 * A label >= 32 sets LDIFF_STATUS_LABEL_RANGE. */
int ldiff_infonce_sample(const uint8_t* labels, int B, int64_t hw, int n_neg, int cap, uint64_t seed,
                         uint64_t offset, int* pair_batch, int* pair_anchor, int* pair_pos, int* pair_neg,
                         int* n_valid, int* status, void* stream);
/* backward: grad_feat (fp32 [B,C,hw], caller-zeroed) += d(sum_p loss_p)/d(feat) * (*grad_scale);
 * grad_scale is a DEVICE scalar (upstream gradient / n_pairs). */
int ldiff_infonce_backward(const float* feat, const int* pair_batch, const int* pair_anchor,
                           const int* pair_pos, const int* pair_neg, const float* lse_per_pair,
                           const float* grad_scale, float* grad_feat, int C, int64_t hw, int n_neg,
                           int n_pairs, float temperature, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LDIFF_H_ */
