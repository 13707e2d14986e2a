/*
 * s2v.h - C ABI of libs2v.so, the B200 (sm_100a) kernels behind the per-frame
 * lip-sync path of VideoReTalking (mel front end -> DNet -> LNet).
 *
 * The reference (Ryukhaan/speech-to-video-mpp) has no FFI for this path: its
 * boundary is the Python call surface
 *     futils/audio.py:45        melspectrogram(wav)
 *     inference.py:209-216      mel-window chunking (inline loop)
 *     futils/flow_util.py:3,41  convert_flow_to_deformation / warp_image
 *     models/LNet.py:122        LNet.forward(audio_sequences, face_sequences)
 *     models/DNet.py:20         DNet.forward(input_image, driving_source, stage)
 * whose bodies are stock ATen calls (cuDNN conv, cuBLAS GEMM, cuFFT, ATen
 * elementwise).  Each entry point below replaces the ATen/numpy call sequence
 * named in its comment; the Python mirror in speech-to-video-mpp_b200/ keeps the
 * reference's names, signatures and state_dict keys and calls these through
 * ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless
 *     the name ends in _host; the caller owns every buffer (no allocation here)
 *   - stream-ordered on `stream` (a cudaStream_t passed as void*); no syncs
 *   - returns 0 on success, a negative S2V_E* code otherwise; never throws
 *   - activations between kernels are fp16, channels-last ("NHWC") views:
 *     element (n,y,x,c) of a view lives at ptr + n*sN + y*sH + x*sW + c
 *     (strides in ELEMENTS; channel stride is 1; C and strides multiples of 8)
 *   - thread-safe for distinct streams and devices; the only process-wide state is immutable after its first
 *     use (per-device kernel attributes / SM counts, resolved entry points, S2V_* development knobs)
 */
#ifndef S2V_H_
#define S2V_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S2V_OK            0
#define S2V_EINVAL       -1   /* bad argument / unsupported shape            */
#define S2V_ECUDA        -2   /* a CUDA runtime / driver call failed         */
#define S2V_EUNSUPPORTED -3   /* device is not sm_100 / feature not available */

/* activation codes used by every fused epilogue */
#define S2V_ACT_NONE    0
#define S2V_ACT_RELU    1
#define S2V_ACT_LRELU   2     /* slope = act_param                            */
#define S2V_ACT_SIGMOID 3
#define S2V_ACT_TANH    4
#define S2V_ACT_GELU    5     /* explicit tanh form, models/transformer.py:15 */

/* output element formats */
#define S2V_OUT_F16_NHWC 0
#define S2V_OUT_F32_NCHW 1    /* y: contiguous [N,Cout,OH,OW] float           */

#define S2V_PAD_ZERO    0
#define S2V_PAD_REFLECT 1

/* fp16 channels-last tensor view */
typedef struct {
  void*   ptr;
  int32_t n, h, w, c;
  int64_t sn, sh, sw;          /* strides in elements */
} s2v_view;

const char* s2v_strerror(int code);
int s2v_version(void);
/* 0 if the current device can run the tcgen05/TMA kernels (compute capability 10.x) */
int s2v_device_ok(void);
/* text of the last CUDA error seen by the library's runtime (diagnostics only) */
const char* s2v_last_cuda_error(void);

/* ------------------------------------------------------------------ mel ---
 * replaces futils/audio.py:45-51 (preemphasis :20-23, librosa.stft :57-61,
 * np.abs, _linear_to_mel :92-96, _amp_to_db :104-106, -ref_level_db,
 * _normalize :111-115) for hparams n_fft=800 hop=200 win=800 sr=16000.
 *   wav      float32 [n_samples]
 *   basis    float32 [80,401] mel filterbank (built on the host exactly like
 *            librosa.filters.mel; passed in so hparams stay a host concern)
 *   mel_out  float32 [80, T], T = 1 + n_samples/200
 * pad_reflect: 0 = zero centre padding (librosa>=0.9), 1 = reflect (<=0.8).    */
int s2v_mel_num_frames(int64_t n_samples);
/* band_range (nullable): int32 [80][2] = [first, last+1) non-zero bin of every
 * filterbank row (a pure speed-up: the triangular rows are sparse).            */
int s2v_melspectrogram_f32(const float* wav, int64_t n_samples, const float* basis,
                           const int32_t* band_range, float* mel_out, int pad_reflect, void* stream);
/* uploads the constant DFT-25 twiddle table; call once per device before the first mel call */
int s2v_mel_init(void);

/* ------------------------------------------------------------- load_wav ---
 * replaces futils/audio.py:9-10 (librosa.core.load(path, sr)[0]) behind the file read: PCM decode to float32 mono
 * (soundfile: int16 / 2^15, int32 / 2^31, uint8 (x-128)/2^7; librosa.to_mono = mean over channels) and librosa 0.9.2's
 * default resampler, resampy 'kaiser_best' (band-limited sinc interpolation with a 64-zero-crossing Kaiser window sampled
 * 512 times per crossing; the host builds the half window `win` [nwin] and its forward differences `delta`, float64).
 *   kind: 0 int16, 1 int32, 2 uint8, 3 float32; pcm interleaved [n_frames][channels]
 *   s2v_resample_out_len = int(n_in * sr_new / sr_orig) samples; the caller (librosa.resample) pads / trims the result to
 *   ceil(n_in * sr_new / sr_orig).  float64 accumulation in resampy's order.                                                */
int64_t s2v_resample_out_len(int64_t n_in, int sr_orig, int sr_new);
int s2v_resample_f32(const float* x, int64_t n_in, int sr_orig, int sr_new, const double* win, const double* delta,
                     int nwin, int num_table, float* y, int64_t n_out, void* stream);
int s2v_pcm_to_mono_f32(const void* pcm, int kind, int channels, int64_t n_frames, float* out, void* stream);

/* replaces the loop of inference.py:209-216 (+ layout of :399/:261):
 * number of 80x16 windows for T mel columns at `fps` (bit-exact int(i*80./fps)) */
int64_t s2v_mel_window_count(int64_t n_cols, double fps);
/* host helper: writes the start column of every window (count entries) */
int s2v_mel_window_starts_host(int64_t n_cols, double fps, int32_t* starts_host, int64_t count);
/* gathers windows [first, first+count) of mel [80,T] into out float32 [count,1,80,16] */
int s2v_mel_windows_f32(const float* mel, int64_t n_cols, double fps, int64_t first, int64_t count,
                        float* out, void* stream);

/* ------------------------------------------------- 3DMM coefficient windows ---
 * replaces futils/inference_utils.py:73-76 (obtain_seq_index) + :78-91 (transform_semantic), called once per
 * frame at preprocessing/facing.py:184, for a whole batch of frames in one launch:
 *   semantic   [n_rows, D] float32 (is_f64 = 0) or float64 (is_f64 = 1) device table, D >= 262
 *   frame_idx  nullable int32 [count] device array of frame indices; NULL = frames first .. first+count-1
 *   ratio      crop_norm_ratio (find_crop_norm_ratio, :93-99); applied to crop[:, 0] (column 259) in the table's
 *              dtype when use_ratio != 0 (the reference's `if crop_norm_ratio:` truthiness is the caller's job)
 *   out        float32 [count, 73, 26]: rows exp 80:144 | angle 224:227 | trans 254:257 | crop 259:262, columns =
 *              table rows clamp(i-13 .. i+12, 0, n_rows-1).  Bit-exact.  count <= 65535 per call.            */
int s2v_semantic_windows(const void* semantic, int is_f64, int n_rows, int D, const int32_t* frame_idx,
                         int first, int count, double ratio, int use_ratio, float* out, void* stream);

/* ------------------------------------------------- Laplacian-pyramid blend ---
 * replaces futils/inference_utils.py:181-222 (Laplacian_Pyramid_Blending_with_mask, called per frame at
 * inference.py:312; cv2.pyrDown / cv2.pyrUp / cv2.add on the CPU) for a batch of frames.  Images are channels-last
 * [N,H,W,C] as cv2 holds them, C = 1, 3 or 4.
 *   s2v_pyrdown_u8 / _f32: cv2.pyrDown ([1 4 6 4 1]^2 / 256, BORDER_REFLECT_101, output (H+1)/2 x (W+1)/2); the 8-bit
 *     form is bit-exact ((sum + 128) >> 8).
 *   s2v_lap_blend_level: one level of the collapse, fused (Laplacian levels of A and B, mask blend, reconstruction):
 *       out = pyrUp(coarse_out) + (a_fine - pyrUp(a_coarse)) * m_fine + (b_fine - pyrUp(b_coarse)) * (1 - m_fine)
 *     with a/b the uint8 Gaussian-pyramid levels, m_fine float32 [N,h,w], coarse_out float32 [N,h/2,w/2,C];
 *     coarse_out == NULL: the coarsest level, out = a_fine * m_fine + b_fine * (1 - m_fine).                       */
int s2v_pyrdown_u8(const uint8_t* src, int N, int H, int W, int C, uint8_t* dst, void* stream);
int s2v_pyrdown_f32(const float* src, int N, int H, int W, int C, float* dst, void* stream);
int s2v_lap_blend_level(const float* coarse_out, const uint8_t* a_fine, const uint8_t* b_fine, const float* m_fine,
                        const uint8_t* a_coarse, const uint8_t* b_coarse, int N, int h, int w, int C, float* out,
                        void* stream);

/* ------------------------------------------------ per-frame image glue ---
 * The numpy / OpenCV lines either side of the networks, batched over frames (uint8 images channels-last [N,H,W,3] as cv2
 * holds them, float network tensors NCHW):
 *   s2v_resize_linear_u8   cv2.resize(x, (OW, OH)) (INTER_LINEAR) of 8-bit images, BIT-EXACT (OpenCV's 11-bit fixed-point
 *                          bilinear; an exact 2x down-scale is INTER_AREA): inference.py:292 (generated face -> its box), :308
 *                          (frames -> 512 x 512), :392-393 (crops -> img_size).  dst: image n at dst + n*dst_sn, rows dst_sh bytes
 *                          apart.  boxes_dev (nullable): int32 [N][4] = (y1, y2, x1, x2) per frame - frame n is resized to
 *                          (y2-y1) x (x2-x1) and written INTO that window of dst image n, i.e. the paste of inference.py:295-297
 *                          (dst = a copy of the full frames); max_box_pixels >= the largest box area.  C = 1 or 3.
 *   s2v_resize_linear_f32  the float32 form (the mask at :308; the blended image back to the frame size at :313 with the
 *                          np.clip(x, 0, 255) before (clip_in) and the np.uint8 truncation after (out_u8: dst is uint8) fused in).
 *                          Taps in double precision = what cv2.resize returns through its IPP path (within 1e-4 on 0..255).
 *   s2v_fake_to_bgr_u8     preprocessing/facing.py:190-192: fake [N,3,H,W] -> np.uint8((clamp(x,-1,1) + 1) / 2. * 255), RGB -> BGR
 *   s2v_face_batch         inference.py:394-399 + :260-262: oface / face uint8 [N,S,S,3] (already at img_size) -> img_batch float32
 *                          [N,6,S,S] (rows >= S/2 of the face zeroed | reference, / 255.) and img_original float32 [N,3,S,S]
 *   s2v_compose_pred_u8    inference.py:267, :282-288, :290: clamp(pred,0,1) [mixed with img_original where the masked input is
 *                          non-zero] * 255 -> uint8 [N,S,S,3].  All bit-exact against the reference's lines.                      */
int s2v_resize_linear_u8(const uint8_t* src, int N, int H, int W, int C, uint8_t* dst, int64_t dst_sn, int64_t dst_sh,
                         const int32_t* boxes_dev, int OH, int OW, int max_box_pixels, void* stream);
int s2v_resize_linear_f32(const float* src, int N, int H, int W, int C, void* dst, int OH, int OW, int clip_in, int out_u8,
                          void* stream);
int s2v_fake_to_bgr_u8(const float* fake, int N, int H, int W, uint8_t* out, void* stream);
int s2v_face_batch(const uint8_t* oface, const uint8_t* face, int N, int S, float* img_batch, float* img_original, void* stream);
int s2v_compose_pred_u8(const float* pred, const float* img_batch, const float* img_original, int N, int S, int compose,
                        uint8_t* out, void* stream);

/* ----------------------------------------------------------- flow warp ---
 * replaces futils/flow_util.py:3-15 + :41-56 (convert_flow_to_deformation,
 * bilinear resize of the grid, F.grid_sample bilinear/zeros/align_corners=False)
 * as ONE kernel.  src/out float32 NCHW [B,C,H,W]; flow float32 NCHW [B,2,h,w].
 * out16 (nullable): additionally writes the warped image as fp16 NHWC into the
 * view's channels [c_off, c_off+C) (feeds DNet's editing net).
 * c_off | S2V_WARP_PACK_SRC (needs c_off == C, 2 C <= 8, an 8-channel-aligned
 * texel at channel 0): the same launch also writes src itself (fp16) into channels
 * [0, C) and zeros into [2 C, 8), i.e. the whole 16-byte texel [src | warp | 0]
 * of models/DNet.py:104 (torch.cat([input_image, warp_image], 1)) in one store. */
#define S2V_WARP_PACK_SRC 0x100
int s2v_flow_warp_f32(const float* src, const float* flow, float* out,
                      int B, int C, int H, int W, int h, int w,
                      const s2v_view* out16, int c_off, void* stream);
/* futils/flow_util.py:3-15 alone: flow [B,2,h,w] -> deformation [B,h,w,2] */
int s2v_flow_to_deformation_f32(const float* flow, float* deformation, int B, int h, int w, void* stream);
/* futils/flow_util.py:41-56 for an explicit deformation grid [B,h,w,2] */
int s2v_warp_deformation_f32(const float* src, const float* deformation, float* out,
                             int B, int C, int H, int W, int h, int w, void* stream);

/* bilinear resize of an fp16 channels-last tensor, align_corners = False (F.interpolate(mode='bilinear') as used by the
 * ENet upsampler: models/base_blocks.py:42-46 ResBlock x0.5, :500-503 ModulatedConv2d x2, models/ENet.py:93,104).
 * x, y: views with equal n and c (c a multiple of 8); any h, w.  chan_scale (nullable): float32 [N][C] multiplied into the
 * result per (image, channel) - the StyleGAN2 modulation of the following conv's input; rows of stride scale_stride
 * floats (0 = C; a multiple of 4, 16-byte aligned base).
 * s2v_resize_planes_f32: the same interpolation on float32 planes (models/ENet.py:93,104: the reference frame to 256 x 256,
 * the LNet input to 96 x 96): plane p of image n at src + n*src_sn + p*src_sp (a channel window of an NCHW tensor).        */
int s2v_resize_bilinear(const s2v_view* x, const s2v_view* y, const float* chan_scale, int64_t scale_stride, void* stream);
int s2v_resize_planes_f32(const float* src, int64_t src_sn, int64_t src_sp, int N, int P, int H, int W,
                          float* dst, int64_t dst_sn, int64_t dst_sp, int OH, int OW, void* stream);

/* ------------------------------------------------------------------ ENet ---
 * The memory-bound pieces of the 96 -> 384 upsampler (models/ENet.py:82-139) around the s2v_conv_tc GEMMs.  The per-sample
 * modulated conv (models/base_blocks.py:487-508, a grouped conv over B x Cout x Cin x k x k weights) runs as ONE shared-weight
 * conv:  conv(x, W*s[n,ci]*d[n,co]) == d[n,co] * conv(x*s[n,ci], W).
 *   s2v_style_demod     d[n][co] = gain * rsqrt(sum_ci w2[co][ci] * s[n][ci]^2 + eps)   (:494-496; w2 = sum over taps of W^2)
 *   s2v_style_epilogue  y = lrelu(x*a[n][c] + bias[c] + noise_w[0]*noise[n][h][w], slope) * post[n][c]      (StyleConv :524-536:
 *                       demodulation * sqrt2, noise injection, bias, LeakyReLU(0.2); `post` = the NEXT conv's modulation)
 *                       a, bias, noise, post nullable; a [N][C], post rows of stride post_stride, noise float32 [N][H][W]
 *   s2v_to_rgb          ToRGB :539-553: out[n][k] = sum_c x[..c]*w[k][c]*s[n][c] + bias[k] + bilinear_x2(skip)[n][k], float32 NCHW,
 *                       cropped by `crop` pixels on every side (ENet.py:131); skip float32 NCHW [N][3][H/2][W/2] or NULL
 *   s2v_reflect_pad_nchw_f32   F.pad(x, (p,p,p,p), 'reflect') of `planes` float32 H x W planes (ENet.py:118-119)          */
int s2v_style_demod(const float* w2, const float* s, int64_t s_stride, int N, int cin, int cout, float eps, float gain,
                    float* out, void* stream);
int s2v_style_epilogue(const s2v_view* x, const float* a, const float* bias, const float* noise, const float* noise_w,
                       float slope, const float* post, int64_t post_stride, const s2v_view* y, void* stream);
int s2v_to_rgb(const s2v_view* x, const float* w, const float* s, int64_t s_stride, const float* bias, const float* skip,
               float* out, int crop, void* stream);
int s2v_reflect_pad_nchw_f32(const float* src, int planes, int H, int W, int pad, float* dst, void* stream);

/* ------------------------------------------------------------- layout ---
 * NCHW float32 [N,C,H,W] -> fp16 NHWC view channels [c_off, c_off+C); channels
 * [c_off+C, c_off+c_fill) are zero-filled (channel padding to a multiple of 8).
 * values are written as src*scale + shift.  src_sn = elements between samples
 * (>= C*H*W; lets a channel window of a wider NCHW tensor be packed in place,
 * e.g. the masked / reference halves of LNet's 6-channel face input).           */
int s2v_pack_nchw_f32(const float* src, int N, int C, int H, int W, int64_t src_sn, const s2v_view* dst,
                      int c_off, int c_fill, float scale, float shift, void* stream);
int s2v_unpack_to_nchw_f32(const s2v_view* src, int c_off, int C, float* dst, void* stream);
/* DNet -> LNet glue of the synthetic full-path configuration (harness convention standing in for
 * the CPU image code of inference.py:188-239,341-411; cf. models/ENet.py:103-104):
 *   ref  = bilinear((clamp(fake,-1,1)+1)/2 -> [oh,ow], align_corners=False)
 *   face = cat(ref with rows >= mask_row zeroed, ref)      float32 [B,2C,oh,ow]                 */
int s2v_glue_fake_to_face_f32(const float* fake, float* face, int B, int C, int H, int W, int oh, int ow,
                              int mask_row, void* stream);

/* ----------------------------------------------------------------- conv ---
 * One descriptor for both convolution kernels.  Replaces nn.Conv2d /
 * nn.ConvTranspose2d (as sub-pixel phases) / nn.Conv1d / nn.Linear (1x1) calls
 * of models/base_blocks.py, models/ffc.py, models/transformer.py, models/DNet.py
 * with the epilogue   y = act( conv(x) * scale[c] + bias[c] + res1 ) + res2
 * (res1: added before the activation - the audio encoder's residual
 *  base_blocks.py:22-25; res2: after - residual streams / partial sums).        */
typedef struct {
  s2v_view x;                  /* input  [N,H,W,Cin]                                   */
  s2v_view y;                  /* output [N,OH,OW,Cout] (view; may be strided phases)  */
  const void*  w;              /* packed weights (layout per kernel, see below)        */
  const float* scale;          /* [Cout] or NULL                                       */
  const float* bias;           /* [Cout] or NULL                                       */
  s2v_view res1;               /* ptr NULL => absent; same dims as y                   */
  s2v_view res2;
  int32_t kh, kw;
  int32_t stride_h, stride_w;
  int32_t pad_h, pad_w;        /* top/left padding                                     */
  int32_t dil_h, dil_w;
  int32_t pad_mode;            /* S2V_PAD_*; reflect is applied by index mirroring (simt)
                                  or must be pre-materialised in x (tc)                */
  int32_t up2;                 /* 1: x is nearest-upsampled x2 on the fly (simt only)  */
  int32_t act;  float act_param;
  int32_t out_mode;            /* S2V_OUT_*                                            */
  float*  y_f32;               /* destination for S2V_OUT_F32_NCHW                     */
  /* optional second K segment (tc only): y += conv_{k2}(x2), stride 1, dilation 1, accumulated in the
   * same TMEM tile before the epilogue; w = [Cout][segment 1 K | segment 2 K].  x2.ptr NULL => absent.
   * (FFC: out_g = convl2g(x_l) + conv2(x + fu(x)), models/ffc.py:231,172 - one GEMM, no partial sum in HBM) */
  s2v_view x2;
  int32_t k2h, k2w, pad2_h, pad2_w;
  /* optional fused statistics (tc, fp16 NHWC output only): the epilogue also writes, per image and per
   * spatial tile of the launch, the sum and sum of squares of every output channel it produced:
   *   stats_partial[((n*stats_chunks_total + chunk)*stats_c_total + stats_c_off + c)*2 + {0,1}],
   *   chunk = stats_chunk_off + tile
   * i.e. the [N][chunks][C][2] layout s2v_ln2d_finalize / s2v_adain_finalize consume (replaces a
   * s2v_chan_stats pass over the output).  tile = tile_y*tiles_x + tile_x of the launch's pixel boxes.
   * stats_groups = stats_gmax = 1: one partial per image, tile and channel (the layout above).
   * stats_gmax = 0, stats_groups = 4, stats_c_total = 4, stats_c_off = 0 (s2v_conv_tc, single N tile): LayerNorm2d
   * TOTALS - the consumer normalises over (C,H,W), so per image and tile only the sum / sum of squares over all
   * channels is kept, one entry per group of 32 output rows of the tile:
   *   stats_partial[((n*stats_chunks_total + chunk)*4 + row_group)*2 + {0,1}]      (consumed by s2v_ln2d_finalize_totals)
   * taken from the fp32 accumulators in registers - no shared-memory pass over the staged tile.           */
  float*  stats_partial;
  int32_t stats_c_off, stats_c_total, stats_chunk_off, stats_chunks_total, stats_groups, stats_gmax;
  /* s2v_conv_tc hint: a structural zero block of the weights.  Input channels >= narrow_cin_from (a multiple of 64) of
   * x feed only the first narrow_cout (a multiple of 32) output channels - every other weight of those channels IS
   * zero.  The kernel may then run those K chunks as narrower MMAs and keep only the non-zero rows in shared memory
   * (FFC at 48x48: the global half of the input reaches only the 32 local outputs).  0 = no hint.           */
  int32_t narrow_cin_from, narrow_cout;
} s2v_conv;

/* SIMT direct convolution (small / awkward layers: Cin=3 7x7, Cout=3, audio
 * encoder, AdaIN MLPs, MappingNet).  w: float32 [kh*kw][Cin][Cout_pad],
 * Cout_pad = Cout rounded up to 4.                                             */
int s2v_conv_simt(const s2v_conv* d, void* stream);

/* tcgen05 / TMEM / TMA implicit-GEMM convolution (zero padding via TMA
 * out-of-bounds fill, strides via TMA element strides; reflect padding is
 * pre-materialised by the producer).  Cout must be a multiple of 8 except for
 * S2V_OUT_F32_NCHW heads (weight rows then zero-padded to 8).
 * w: fp16 [Cout][kh*kw][Cin64] K-major, Cin64 = Cin rounded up to 64 (zero
 * filled).  box_w*box_h*box_n must be 128 (the M tile is a box of pixels).
 * pad_h/pad_w are the TOP/LEFT padding only; OH/OW may be smaller than the
 * symmetric-padding formula (asymmetric padding of the sub-pixel phases of
 * nearest-x2 + conv3x3 and of ConvTranspose2d).                                 */
int s2v_conv_tc(const s2v_conv* d, int box_w, int box_h, int box_n, void* stream);
/* N tile (BN) s2v_conv_tc picks for a given Cout (host-side sizing of the statistics groups) */
int s2v_conv_tc_tile_n(int cout);

/* 7x7, stride 1, pad 3 "head" convolutions with Cout <= 8 (FinalBlock2d, models/base_blocks.py:444-457; flow_out,
 * models/DNet.py:72-76) on tcgen05 with the kx taps folded into the GEMM's N dimension (7x fewer MMAs than
 * s2v_conv_tc).  d->x: fp16 NHWC, C a multiple of 64; d->out_mode = S2V_OUT_F32_NCHW, d->y = {n, h, w, c = Cout} shape
 * only, d->y_f32 [N][Cout][H][W]; bias optional; act NONE / RELU / LRELU / SIGMOID / TANH.
 * d->w: fp16 [NB][chunks*7*64] with CP = 2 / 4 / 8 the smallest pad >= Cout and NB = 16 / 32 / 64 rows: row kx*CP + co,
 * column (chunk*7 + ky)*64 + ci; rows >= 7*CP and co >= Cout zero.                                               */
int s2v_conv_head(const s2v_conv* d, void* stream);

/* grouped small linears (all AdaIN gamma/beta heads of a net in one launch,
 * models/base_blocks.py:136-141,149-151):
 *   out[b][g.out_off + j] = bias_g[j] + sum_k hidden[b][g.in_off + k] * wt_g[k][j]
 * hidden fp16 [B][hidden_stride]; wt float32 [K][nout] (transposed); out float32 */
typedef struct {
  const float* wt; const float* bias;
  int32_t in_off, k, out_off, nout;
} s2v_lin_group;
int s2v_grouped_linear(const void* hidden_f16, int64_t hidden_stride, int B,
                       const s2v_lin_group* groups_dev, const int32_t* tile2group_dev, int n_tiles,
                       float* out, int64_t out_stride, void* stream);

/* ---------------------------------------------------------------- norms ---
 * Deterministic (fixed-order) statistics + fused apply.
 * stats: per (n, chunk, c) partial sum / sum-of-squares over the chunk's pixels.
 *   partial float32 [N][chunks][C][2]                                           */
int s2v_chan_stats(const s2v_view* x, int chunks, float* partial, void* stream);
/* LayerNorm2d over (C,H,W) (models/base_blocks.py:52-69, eps 1e-5):
 *   a[n][c] = rstd[n]*gamma[c],  b[n][c] = beta[c] - mean[n]*a[n][c]            */
int s2v_ln2d_finalize(const float* partial, int N, int chunks, int C, int64_t count_per_channel,
                      const float* gamma, const float* beta, float eps, float* a, float* b, void* stream);
/* same, for the LayerNorm2d-totals partial layout [N][chunks][4][2] written by s2v_conv_tc with stats_gmax = 0 */
int s2v_ln2d_finalize_totals(const float* partial, int N, int chunks, int C, int64_t count_per_channel,
                             const float* gamma, const float* beta, float eps, float* a, float* b, void* stream);
/* InstanceNorm2d + AdaIN (models/base_blocks.py:127-157):
 *   a = rstd[n][c]*(1+gamma[n][c]),  b = beta[n][c] - mean[n][c]*a
 * gamma/beta float32 with row stride gb_stride (rows of the grouped_linear out) */
int s2v_adain_finalize(const float* partial, int N, int chunks, int C, int64_t count_per_channel,
                       const float* gamma, const float* beta, int64_t gb_stride, float eps,
                       float* a, float* b, void* stream);
/* y = act(x*a[n][c] + b[n][c]) ; optional 2x2 average pool AFTER the activation
 * (DownBlock2d, base_blocks.py:100); optional +res (after activation/pool);
 * reflect1: additionally mirrors the result into a 1-pixel reflect border around
 * y (y must be the interior view of a buffer padded by 1).                      */
int s2v_affine_act(const s2v_view* x, const float* a, const float* b, int act, float act_param,
                   int pool2, const s2v_view* res, const s2v_view* y, int reflect1, void* stream);
/* y = act(x*a + b) + act(res*ra + rb): the decoder's  up-branch LayerNorm2d + LeakyReLU  and  jump-branch LayerNorm2d +
 * LeakyReLU + add  (models/LNet.py:66-72, base_blocks.py:112-124,429-441) applied in ONE pass over the two raw conv
 * outputs - the normalised up-branch tensor is never materialised (saves one write + one read of it).
 * ra / rb: per-(n,c) scale / shift of `res` (from s2v_ln2d_finalize), same layout as a / b.                          */
int s2v_affine_act2(const s2v_view* x, const float* a, const float* b, int act, float act_param,
                    const s2v_view* res, const float* ra, const float* rb, const s2v_view* y, int reflect1, void* stream);

/* Single-pass InstanceNorm2d + AdaIN + activation [+res] [reflect border] for maps whose (image, channel group)
 * slab fits in shared memory (s2v_adain_fused_fits > 0): one read of x instead of chan_stats + adain_finalize
 * + affine_act.  Same arithmetic and the same fixed reduction order for every batch size.                     */
int s2v_adain_fused_fits(int h, int w, int c);
int s2v_adain_fused(const s2v_view* x, const float* gamma, const float* beta, int64_t gb_stride, float eps,
                    int act, float act_param, const s2v_view* res, const s2v_view* y, int reflect1, void* stream);
/* nn.LayerNorm(C) over the channel dim of every pixel/token (transformer.py:27,35) */
int s2v_token_layernorm(const s2v_view* x, const float* gamma, const float* beta, float eps,
                        const s2v_view* y, void* stream);
/* y = a + b (elementwise over views; used for Jump + Up, base LNet.py:75)       */
int s2v_add(const s2v_view* a, const s2v_view* b, const s2v_view* y, void* stream);
/* fills the 1-pixel reflect border of a padded buffer from its interior view    */
int s2v_reflect_border(const s2v_view* interior, void* stream);

/* ------------------------------------------------------------------ FFT ---
 * FourierUnit halves (models/ffc.py:99-102 and :116-121), H,W in {12,24,48}:
 * rfft2:  x [N,H,W,C] -> spec [N,H,W/2+1,2C] with channel 2c = Re, 2c+1 = Im
 *         (torch.fft.rfftn norm='ortho' + stack/permute/view interleave)
 * irfft2: spec -> y [N,H,W,C] (torch.fft.irfftn norm='ortho', Im of the DC and
 *         Nyquist columns ignored after the H inverse) ; y += add (x + fu(x) of
 *         ffc.py:172) when add.ptr != NULL
 * 48 x 48: dense fp16 DFT matrices on the tensor cores, 8-channel tiles moved by
 * TMA (csrc/fft2d_mma.cu; fp32 accumulate, one more fp16 rounding than the
 * register FFT: <= 1e-3 of the output peak); 24 / 12: register FFT (fft2d.cu).
 * S2V_FFT_MMA selects per direction and size (0 = register FFT everywhere).      */
/* uploads the constant W_48 twiddle table and the DFT matrix fragment tables; call once per device before the first FFT call */
int s2v_fft_init(void);
int s2v_rfft2(const s2v_view* x, const s2v_view* spec, void* stream);
int s2v_irfft2(const s2v_view* spec, const s2v_view* add, const s2v_view* y, void* stream);

/* ------------------------------------------------------------ attention ---
 * softmax(q k^T * scale) v per (n, head) (models/transformer.py:77-86).
 * q,k,v,o: fp16 views [N,1,T,heads*dh] (T tokens <= 256, dh = 64)               */
int s2v_attention(const s2v_view* q, const s2v_view* k, const s2v_view* v, const s2v_view* o,
                  int heads, float scale, void* stream);

/* MappingNet tail: AdaptiveAvgPool1d(1) over L (models/DNet.py:53)
 * x [N,1,L,C] -> y [N,1,1,C]                                                    */
int s2v_mean_over_w(const s2v_view* x, const s2v_view* y, void* stream);

/* ------------------------------------------------------ whole networks ---
 * LNet.forward (models/LNet.py:122-139) and DNet.forward (models/DNet.py:20-28) for hosts without Python.
 * The layer plan of a network at ONE batch size (the ordered list of launcher calls of this header, with the folded and
 * packed weights) is written once by the Python package (s2v_b200.plan_export.export_lnet / export_dnet) into a
 * relocatable plan file; the functions below load it on the host, bind it to two caller-owned device buffers and replay
 * it on a stream.  Nothing is allocated on the device and nothing synchronises (s2v_plan_bind's constant upload is a
 * pageable-memory copy and therefore host-synchronous).  One plan object = one in-flight forward at a time (its workspace
 * holds the activations); use one plan per stream for concurrent forwards.
 *   const_dev      s2v_plan_const_bytes() bytes, 256-byte aligned: weights / epilogue vectors / pointer tables
 *   workspace_dev  s2v_plan_workspace_bytes() bytes, 256-byte aligned: activations, statistics and the I/O slots
 * I/O slots live inside the workspace (s2v_plan_io_info: name, byte offset, size): "mel" [B,1,80,16], "face" [B,6,96,96],
 * "out" [B,3,96,96] for LNet; "img" [B,3,256,256], "coeff" [B,73,1,T], "flow" [B,2,64,64], "warp" / "fake" [B,3,256,256]
 * for DNet - all float32, contiguous, B = the batch the plan was exported for (pad smaller batches with zeros: frames
 * are independent).  s2v_lnet_forward / s2v_dnet_forward copy device buffers into / out of the slots around
 * s2v_plan_run; output pointers of s2v_dnet_forward may be NULL.                                                  */
typedef struct s2v_plan s2v_plan;
int s2v_plan_load(const char* path_host, s2v_plan** out);
int s2v_plan_load_memory(const void* data_host, int64_t bytes, s2v_plan** out);
void s2v_plan_free(s2v_plan* plan);
int64_t s2v_plan_const_bytes(const s2v_plan* plan);
int64_t s2v_plan_workspace_bytes(const s2v_plan* plan);
int s2v_plan_num_ops(const s2v_plan* plan);
int s2v_plan_num_io(const s2v_plan* plan);
int s2v_plan_io_info(const s2v_plan* plan, int i, const char** name, int64_t* offset, int64_t* bytes, int* is_output);
int s2v_plan_bind(s2v_plan* plan, void* const_dev, void* workspace_dev, void* stream);
int s2v_plan_run(const s2v_plan* plan, void* stream);
int s2v_lnet_forward(const s2v_plan* plan, const float* mel, const float* face, float* out, void* stream);
int s2v_dnet_forward(const s2v_plan* plan, const float* input_image, const float* driving_source, float* flow_field,
                     float* warp_image, float* fake_image, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* S2V_H_ */
