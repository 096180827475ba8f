/*
 * eas_b200.h -- C ABI of libeas_b200.so: the B200 (sm_100a) hot path of EAS-SNN.
 *
 * The reference (Windere/EAS-SNN) has no FFI on this path: its seam is Python nn.Module
 * substitution.  These entry points are what a ctypes binding on the
 * reference side calls; each one names the reference code it replaces.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless marked "host";
 *   - the caller owns every buffer including the workspace; the library never allocates,
 *     never synchronises, keeps no mutable global state; all work is enqueued on `stream`
 *     (a cudaStream_t passed as void*, e.g. torch.cuda.current_stream().cuda_stream);
 *   - return value: 0 = ok; negative = argument/shape contract violation detected on the host
 *     before anything is launched (EAS_E_*); positive = cudaError_t of a failed launch;
 *   - reentrant: safe from several host threads / streams of one process (one process per GPU).
 */
#ifndef EAS_B200_H_
#define EAS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EAS_ABI_VERSION 1

enum {
  EAS_OK = 0,
  EAS_E_NULL = -1,        /* a required pointer is NULL                     */
  EAS_E_SHAPE = -2,       /* a dimension is out of the supported range      */
  EAS_E_UNSUPPORTED = -3, /* flag / dtype / kernel-size combination missing */
  EAS_E_WORKSPACE = -4,   /* workspace too small                            */
  EAS_E_ALIGN = -5        /* pointer not aligned as required                */
};

enum { EAS_F32 = 0, EAS_I32 = 1, EAS_BF16 = 2, EAS_U8 = 3, EAS_F16 = 4 };
enum { EAS_READOUT_SUM = 0, EAS_READOUT_LAST = 1, EAS_READOUT_AVG = 2 };
enum { EAS_SG_ATAN = 0, EAS_SG_SIGMOID = 1, EAS_SG_RECT = 2 };

int eas_abi_version(void);
/* Human readable text for a negative EAS_E_* code or a positive cudaError_t. */
const char* eas_error_string(int code);

/* ------------------------------------------------------------------------------------------
 * (a-1) Event binning.  Replaces GEN1Dataset.slice_events + agrregate('micro_sum' -> 'sum')
 *       yolox/data/datasets/gen1.py:313-328, 330-360 (same code: gen4.py, rvt_gen4.py:411-454).
 *
 * B independent, time-sorted windows stored back to back in SoA form (the reference's
 * events_struct, yolox/utils/util.py:119-121); window b is events [offsets[b], offsets[b+1]).
 * Output hist[B][Tm][2][H][W] int32 counts, written completely (no pre-zeroing needed):
 *   tw = (t_last - t_first) / Tm (integer), micro-bin k = [t0 + k*tw, t0 + (k+1)*tw),
 *   events at or after t0 + Tm*tw are dropped, tw == 0 -> all bins empty,
 *   channel 0 <- p == 0, channel 1 <- p != 0, pixel = y*W + x.
 * Events whose (x, y) fall outside [0,W) x [0,H) are ignored (the reference would raise).
 * n_events = offsets[B] (known to the host from the array length) sizes the launch.
 * ---------------------------------------------------------------------------------------- */
size_t eas_bin_events_ws_bytes(int64_t B, int Tm);
int eas_bin_events(const int16_t* x, const int16_t* y, const int64_t* t, const uint8_t* p,
                   const int64_t* offsets, int64_t B, int64_t n_events, int H, int W, int Tm,
                   int32_t* hist, void* ws, size_t ws_bytes, void* stream);
/* Same with options: strategy 0 = auto, 1 = global reductions, 2 = shared-memory tiles (the kernel that holds both
 * polarities of a slab per work item when two slabs fit a CTA and the batch has items for every SM, else one plane per
 * item), 3 / 4 = those two tiles kernels explicitly;
 * out_dtype EAS_I32, or EAS_F32 (the counts as fp32, exact below 2^24: what the sampler's first
 * convolution consumes, and the dtype the reference casts its histogram to on the device), or EAS_U8:
 * the compact form the sampler reads directly (eas_sampler_cfg.in_dtype = EAS_U8) -- a quarter of the
 * bytes to write and to read.  `hist` is then ONE buffer of eas_hist_u8_bytes(B, Tm, H, W) bytes:
 *   [0, nbins)                    uint8 min(count, 255), nbins = B*Tm*2*H*W (< 2^32)
 *   at nbins rounded up to 256    uint32 n_sat, uint32 lost, uint64 (address of the registered sticky flag, or 0),
 *                                 then EAS_HIST_U8_SAT_CAP x {uint32 bin index, uint32 count} for the
 *                                 bins that reached 255: the information of the int histogram, exactly,
 *                                 unless more than EAS_HIST_U8_SAT_CAP bins saturate in one call (lost = 1;
 *                                 callers check it -- the Python side raises).
 * EAS_U8 is written by the shared-memory tiles strategy only (frames that fit it; EAS_E_UNSUPPORTED
 * otherwise).  eas_hist_u8_expand turns it back into the dense int32 / fp32 histogram. */
#define EAS_HIST_U8_SAT_CAP 4096
size_t eas_hist_u8_bytes(int64_t B, int Tm, int H, int W);
int eas_hist_u8_expand(const void* hist_u8, int64_t B, int Tm, int H, int W, void* out, int out_dtype,
                       void* stream);
/* Error reporting without a stream stall: sets *sticky_flag = 1 (never clears it) when the histogram's `lost`
 * word is set.  sticky_flag may be pinned host memory (device-accessible under unified addressing): the host
 * then polls a plain int. */
int eas_hist_u8_report(const void* hist_u8, int64_t B, int Tm, int H, int W, int32_t* sticky_flag, void* stream);
/* The same without a launch: registers (process-wide; NULL = none) an int32 that every later EAS_U8 binning call sets
 * to 1 at the moment it loses a count.  Device-accessible memory, e.g. pinned host memory. */
int eas_hist_u8_set_sticky(int32_t* sticky_flag);
int eas_bin_events_ex(const int16_t* x, const int16_t* y, const int64_t* t, const uint8_t* p,
                      const int64_t* offsets, int64_t B, int64_t n_events, int H, int W, int Tm,
                      void* hist, void* ws, size_t ws_bytes, void* stream, int strategy, int out_dtype);

/* ------------------------------------------------------------------------------------------
 * (f-1) PSEE .dat recordings on the GPU.  rec = n_rec raw 8-byte little-endian Event2D records
 *       {uint32 t_us; uint32 x:14 | y:14 << 14 | p << 28}, time-sorted, resident in device memory:
 *       the payload of a `_td.dat` file after its text header, as the reference's loader reads it
 *       (yolox/utils/psee_loader/io/dat_events_tools.py:24, 40-51).
 *
 * eas_dat_windows replaces GEN1Dataset.search_events (yolox/data/datasets/gen1.py:217-236) over
 * PSEELoader.seek_time / load_delta_t (psee_loader.py:128-238): for every label time t_label[b] the
 * window is [t + win_lo, t + win_hi) us; while it holds no event it is moved back by its own
 * length, at most max_backoff + 1 times (the reference's zero_trigger loop with
 * max_backoff = slice_args['num_slice']).  ranges[b] = {first record, one past the last record}.
 * Reproduces the loader exactly, including seek_time's exact-hit cursor advance while more than
 * 100000 records remain in its bisection.
 *
 * eas_bin_dat = eas_bin_events_ex on record ranges (decode + slice_events + agrregate('micro_sum'),
 * gen1.py:313-360): hist [B][Tm][2][H][W] int32 or f32.
 * ---------------------------------------------------------------------------------------- */
int eas_dat_windows(const void* rec, int64_t n_rec, const int64_t* t_label, int64_t B, int64_t win_lo,
                    int64_t win_hi, int32_t max_backoff, int64_t* ranges, void* stream);
size_t eas_bin_dat_ws_bytes(int64_t B, int Tm);
int eas_bin_dat(const void* rec, int64_t n_rec, const int64_t* ranges, int64_t B, int H, int W, int Tm,
                void* hist, void* ws, size_t ws_bytes, void* stream, int strategy, int out_dtype);

/* ------------------------------------------------------------------------------------------
 * (a-2) Adaptive event sampler.  Replaces AdaptiveRSNNEmbedding.forward / update,
 *       yolox/models/embedding.py:132-226 with the Rectangle surrogate, activation.py:17-30.
 *
 * events [B][Tm][2][H][W] (f32, or the int32 / compact uint8 histogram of eas_bin_events);
 * weights of input_conv / gate_conv in PyTorch layout: w0 [4][2][k][k], b0 [4] and, for
 * depth == 2, w1 [4][4][k][k], b1 [4] (pass NULL for depth 1);
 * out [Ts][B][2][H][W] f32, fully written.
 * For training pass v_seq / gate_seq ([Tm][B][2][H][W] f32 each, in the sampler's own
 * newest-first step order); for inference pass NULL.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int32_t B, H, W, Tm, Ts;
  int32_t ksize;        /* 3, 5 or 7                                            */
  int32_t depth;        /* 1 or 2                                               */
  int32_t readout;      /* EAS_READOUT_*                                        */
  int32_t hard_reset;   /* 1: vm = v*(1-s) + vreset*s ; 0: vm = v - thresh*s     */
  float vreset;
  float thresh;
  int32_t spike_attach; /* SAT: read-out multiplied by the spike                 */
  int32_t write_zero;   /* RPD: residual potential of silent pixels dropped      */
  int32_t use_abs;      /* relu on the aggregated frames                         */
  int32_t in_dtype;     /* EAS_F32, EAS_I32 or EAS_U8 (compact histogram, see above) */
  int32_t algo;         /* EAS_SAMPLER_* (forward only): AUTO picks the row-folded tensor-core kernel for
                           depth 2, k 5, W % 4 == 0, 16 B aligned buffers (and re-runs on the FP32-pipe
                           kernel when an input is not exact in fp16), else the FP32-pipe kernel */
  float surr_alpha;     /* Rectangle.alpha (a class attribute, activation.py:18-30): the backward passes
                           grad * alpha inside |v - thresh| < 0.5 / alpha.  0 is read as 1 (the reference's value) */
} eas_sampler_cfg;

/* TENSOR: row-folded tcgen05 kernel, inputs must be exact in one fp16 plane (event counts <= 2048; the
 * result is undefined otherwise -- AUTO checks).  TENSOR_SPLIT: tcgen05 kernel with hi + lo input planes
 * (real-valued inputs, e.g. letterboxed frames). */
enum { EAS_SAMPLER_AUTO = 0, EAS_SAMPLER_FP32 = 1, EAS_SAMPLER_TENSOR = 2, EAS_SAMPLER_TENSOR_SPLIT = 3 };

typedef struct {
  const float *in_w0, *in_b0, *in_w1, *in_b1;
  const float *gate_w0, *gate_b0, *gate_w1, *gate_b1;
} eas_sampler_weights;

size_t eas_sampler_fwd_ws_bytes(const eas_sampler_cfg* cfg);
int eas_sampler_fwd(const eas_sampler_cfg* cfg, const void* events, const eas_sampler_weights* w,
                    float* out, float* v_seq, float* gate_seq, void* ws, size_t ws_bytes, void* stream);

/* Backward (BPTT, Rectangle surrogate window |v - thresh| < 0.5).  grad_out [Ts][B][2][H][W];
 * v_seq / gate_seq as saved by the forward; the eight gradient buffers have the shapes of the
 * corresponding weights and are OVERWRITTEN; grad_events may be NULL. */
typedef struct {
  float *in_w0, *in_b0, *in_w1, *in_b1;
  float *gate_w0, *gate_b0, *gate_w1, *gate_b1;
} eas_sampler_grads;
size_t eas_sampler_bwd_ws_bytes(const eas_sampler_cfg* cfg);
int eas_sampler_bwd(const eas_sampler_cfg* cfg, const void* events, const eas_sampler_weights* w,
                    const float* v_seq, const float* gate_seq, const float* grad_out,
                    const eas_sampler_grads* gw, float* grad_events, void* ws, size_t ws_bytes,
                    void* stream);

/* ------------------------------------------------------------------------------------------
 * (a-4) Multi-step parametric LIF.  Replaces spikingjelly 0.0.0.0.14
 *       neuron.ParametricLIFNode (step_mode 'm', backend 'torch') as configured at
 *       yolox/utils/utils_snn.py:44-53.
 *
 * x, spikes: [T][N] (N = product of the remaining dims), f32 or bf16.  w: device scalar (f32).
 *   decay_input == 0:  h = v*(1-sigmoid(w)) + x        (v_reset None/0)
 *   decay_input == 1:  h = v + (x - v)*sigmoid(w)
 *   s = (h - v_th >= 0);  soft reset v = h - s*v_th  |  hard reset v = (1-s)*h + s*v_reset
 * v0 (optional, [N] f32) initial potential, NULL = zeros(+v_reset); v_out (optional, [N] f32)
 * final potential.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int64_t T, N;
  float v_threshold;
  int32_t hard_reset;
  float v_reset;
  int32_t decay_input;
  int32_t detach_reset;
  int32_t surrogate;    /* EAS_SG_*  (backward only) */
  float alpha;
  int32_t dtype;        /* EAS_F32 or EAS_BF16 (x, spikes, grads) */
} eas_plif_cfg;

int eas_plif_fwd(const eas_plif_cfg* cfg, const void* x, const float* w, const float* v0,
                 void* spikes, float* v_out, void* stream);
size_t eas_plif_bwd_ws_bytes(const eas_plif_cfg* cfg);
/* grad_w: device scalar, OVERWRITTEN with d loss / d w. */
int eas_plif_bwd(const eas_plif_cfg* cfg, const void* x, const float* w, const float* v0,
                 const void* grad_spikes, void* grad_x, float* grad_w, void* ws, size_t ws_bytes,
                 void* stream);

/* ------------------------------------------------------------------------------------------
 * (a-5) conv -> BatchNorm (folded) -> multi-step PLIF, inference.  Replaces
 *       BaseConv.forward after convert_to_spiking: SeqToANNContainer(Conv2d) ->
 *       layer.BatchNorm2d('m') -> ParametricLIFNode, yolox/models/network_blocks.py:52-53,
 *       yolox/utils/utils_snn.py:25-53; BN folding per yolox/utils/model_utils.py:61-75.
 *
 * Activations are channels-last fp16: x [Tx][B][H][W][Cin] with Tx == T, or Tx == 1 when the
 * input is the same for every time step (sampler output broadcast, spiking_yolox.py:54-55).
 * Weights: n_wsplit fp16 planes [n_wsplit][Cout][kh][kw][Cin] whose sum, times w_unscale[c], is the
 * BN-folded fp32 weight (1 plane = "fast", 2 planes = 22 mantissa bits: fp32-equivalent for the
 * integer-valued spike inputs); n_xsplit input planes likewise for real-valued inputs (|x| < 65504).
 * bias [Cout] f32 (folded BN shift).  Output spikes [T][B][Ho][Wo][Cout] fp16 (0/1).  Tensor-core path (tcgen05 + TMEM + TMA).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int32_t T, Tx, B, H, W, Cin, Cout, ksize, stride;
  int32_t n_wsplit, n_xsplit;
  float v_threshold;
  int32_t hard_reset;
  float v_reset;
  int32_t decay_input;
  int32_t out_mode;     /* EAS_CONV_OUT_* */
  int32_t x_ld;         /* channels per pixel of the input buffer (>= Cin; 0 = Cin): lets the input be
                           a channel slice of a wider concat buffer */
  int32_t out_ld;       /* channels per pixel of the output buffer (>= Cout; 0 = Cout): lets the
                           output land in a channel slice of a concat buffer (CSPLayer / SPP cat) */
  int32_t res_ld;       /* channels per pixel of the residual buffer (0 = Cout) */
  const void* residual; /* optional fp16 [T][B][Ho][Wo][res_ld]: SEW shortcut added to the spikes
                           (Bottleneck, network_blocks.py:99-103); NULL = none */
  const float* w_unscale; /* optional [Cout] f32: the accumulator of channel c is multiplied by
                           w_unscale[c] before the bias (undoes the per-channel power-of-two scale
                           that keeps the fp16 weight planes in range); NULL = 1 */
} eas_conv_cfg;

enum {
  EAS_CONV_OUT_SPIKES = 0, /* fp16 spikes [T][B][Ho][Wo][out_ld]                                  */
  EAS_CONV_OUT_PREACT = 1, /* f32 conv + bias [Tx][B][Ho][Wo][out_ld] (no neuron)                 */
  EAS_CONV_OUT_SILU2 = 2   /* SiLU(conv + bias) as 2 fp16 planes hi/lo [2][Tx][B][Ho][Wo][out_ld]
                              (the ANN stem, network_blocks.py:191-213, feeding a spiking conv)   */
};

size_t eas_conv_bn_plif_ws_bytes(const eas_conv_cfg* cfg);
int eas_conv_bn_plif_fwd(const eas_conv_cfg* cfg, const void* x, const void* w_planes,
                         const float* bias, const float* plif_w, void* out, void* ws,
                         size_t ws_bytes, void* stream);

/* (f-4) RVT-preprocessed stacked histograms -> per-polarity counts: the 'event_sum' branch of
 * RVTGEN4Dataset.generate_slices, yolox/data/datasets/rvt_gen4.py:120-122
 * (ev_repr.reshape(n, 2, -1, H, W).sum(axis=2)).  repr: uint8 [n][2*nb][H][W] (channel = polarity*nb + bin),
 * out: f32 [n][2][H][W] (exact integer counts, <= 255*nb). */
int eas_rvt_event_sum(const uint8_t* repr, int64_t n, int nb, int H, int W, float* out, void* stream);

/* SPP max-pools of the backbone (SPPBottleneck.m, yolox/models/network_blocks.py:128-147; stride 1,
 * padding k/2).  cat: fp16 channels-last [n_images][H][W][ld] whose channels [0, C) hold x; channels
 * [C, 2C), [2C, 3C), [3C, 4C) receive maxpool_k1 / k2 / k3 of x (the torch.cat of :146).  In place. */
int eas_spp_pool_fwd(void* cat, int64_t n_images, int H, int W, int C, int ld, int k1, int k2, int k3,
                     void* stream);

/* ------------------------------------------------------------------------------------------
 * (f-2) Glue between the spiking backbone, the ANN PAFPN / head and the detections.
 * Activations: channels-last fp16, real values as two planes hi + lo (plane stride in elements).
 * ---------------------------------------------------------------------------------------- */
/* `out_features[f].mean(axis=0)`, yolox/models/spiking_yolo_pafpn.py:98.  spikes: fp16 [T][n_pix][x_ld]
 * (channels [0,C) used), out: planes hi/lo [n_pix][out_ld] = sum_t / T (a channel slice of a concat buffer). */
int eas_time_mean_planes(const void* spikes, int T, int64_t n_pix, int C, int x_ld, void* out, int out_ld,
                         int64_t out_plane_stride, void* stream);
/* nn.Upsample(scale_factor=2, mode="nearest"), spiking_yolo_pafpn.py:36,102,107 (yolo_pafpn.py same lines).
 * in: [n_planes][n_images][H][W][in_ld] -> out: [n_planes][n_images][2H][2W][out_ld], channels [0,C). */
int eas_upsample2x_planes(const void* in, int n_planes, int64_t in_plane_stride, int64_t n_images, int H, int W,
                          int C, int in_ld, void* out, int out_ld, int64_t out_plane_stride, void* stream);
/* YOLOXHead inference tail, yolox/models/yolo_head.py:187-199 (sigmoid, flatten, concat over levels) and
 * :232-250 (decode_outputs).  preds: f32 [n_images][H][W][ld], channels = (x, y, w, h, obj, cls...) raw conv
 * outputs of one level; out: f32 [n_images][n_anchors_total][n_ch], this level's anchors start at
 * anchor_offset.  decode = 0 keeps (x, y, w, h) raw (decode_in_inference = False). */
int eas_yolox_decode(const float* preds, int64_t n_images, int H, int W, int n_ch, int ld, float stride,
                     int decode, float* out, int64_t anchor_offset, int64_t n_anchors_total, void* stream);
/* Focus stem front end, yolox/models/network_blocks.py:199-213 (space-to-depth of the 2-channel frames) fused with
 * the im2col of the stem's 3x3 conv: frames f32 [n_images][2][H][W] -> out fp16 planes hi/lo
 * [n_images][H/2][W/2][80], row = [tap (ky, kx)][phase dy + 2 dx][c] (72 values + 8 zeros), zero padding at the
 * borders; the stem conv then runs as a 1x1 conv with K = 80 through eas_conv_bn_plif_fwd. */
int eas_focus_im2col(const float* frames, int64_t n_images, int H, int W, void* out, int64_t out_plane_stride,
                     void* stream);
/* (f-4) SpikeCountEmbedding.forward, yolox/models/embedding.py:9-24: micro-bin histograms summed over the Tm
 * micro-bins (`events.transpose(0, 1).sum(axis=0)`).  hist: [n][Tm][plane_elems] f32 or i32 counts (in_dtype),
 * out: f32 [n][plane_elems]; plane_elems = 2*H*W, a multiple of 4. */
int eas_hist_time_sum(const void* hist, int in_dtype, int64_t n, int Tm, int64_t plane_elems, float* out,
                      void* stream);
/* (f-4) Voxel grid with bilinear interpolation in time: to_voxel_grid_numpy, yolox/utils/event_reps.py:30-89.
 * Events as for eas_bin_events (SoA, windows back to back, offsets[B+1]); per window the timestamps are normalised
 * to [0, n_bins] in float64, every event adds pol * (1 - dt) to bin int(ts) and pol * dt to the next one (bins
 * >= n_bins dropped: the last event of a window contributes nothing).  signed_polarity 1: pol = +1 / -1 (the
 * reference's stated intent); 0: pol = +1 for all events, which is what the reference computes on its own
 * events_struct dtype (p is bool: `pols[pols == 0] = -1` stores True, event_reps.py:62-63).  out: f32
 * [B][n_bins][H][W], fully written.  A window whose first and last timestamps coincide stays zero (the reference
 * produces NaN there). */
int eas_voxel_grid(const int16_t* x, const int16_t* y, const int64_t* t, const uint8_t* p, const int64_t* offsets,
                   int64_t B, int64_t n_events, int H, int W, int n_bins, int signed_polarity, float* out, void* stream);
/* (f-3) Letterbox + bilinear resize of micro-frames: GEN1Dataset.get_random_data(random=False),
 * yolox/data/datasets/gen1.py:433-483 (cv2.resize INTER_LINEAR of every plane to [nh][nw], pasted at (dy, dx) into a
 * zero canvas [oh][ow]).  in: [n_planes][ih][iw] f32 or i32 counts; x0 / fx ([nw]) and y0 / fy ([nh]): source tap and
 * float32 weight of the second tap per output column / row (cv2's half-pixel rule; the second tap is min(tap+1, last));
 * out: f32 [n_planes][oh][ow], ow % 4 == 0, written completely. */
int eas_letterbox_bilinear(const void* in, int in_dtype, int64_t n_planes, int ih, int iw, const int32_t* x0,
                           const float* fx, const int32_t* y0, const float* fy, int nh, int nw, int dy, int dx,
                           float* out, int oh, int ow, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EAS_B200_H_ */
