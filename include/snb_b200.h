/*
 * snb_b200.h -- C ABI of the B200-native tiled-segmentation hot path.
 *
 * The reference (BloodAxe/segmentation-networks-benchmark) is pure Python and has no FFI of its own; the only
 * native boundary it knows is the call surface of the external `inplace_abn` extension
 * (lib/modules/abn/functions.py:1,46-118).  This header is therefore the boundary a maintainer would bind
 * with `ctypes` to move the hot path of inria_submit.py / lib/tiles.py onto a B200:
 *
 *   group              replaces (reference file:line)
 *   -----------------  ---------------------------------------------------------------------------------
 *   snb_slicer_*       ImageSlicer.__init__ margins + crop list                lib/tiles.py:35-96
 *   snb_split_*        ImageSlicer.split / cut_patch (+ NormalizeImage,        lib/tiles.py:98-135,
 *                      InMemoryDataset HWC->CHW .float(), tta_d4_aug)          lib/augmentations.py:452-491,
 *                                                                              lib/common.py:59-76
 *   snb_merge          ImageSlicer.merge (+ tta_d4_deaug, `mask > 0.5`)        lib/tiles.py:137-161,
 *                                                                              lib/augmentations.py:494-511,
 *                                                                              inria_submit.py:305
 *   snb_conv_*         nn.Conv2d / nn.ConvTranspose2d (+bias, ReLU, 1x1 head,  lib/models/unet16.py:8-49,113-131,
 *                      sigmoid) as executed by UNet16/UNet11/ZF_UNET forward   lib/models/unet11.py:106-122,
 *                                                                              inria_submit.py:250-251
 *   snb_maxpool2x2     nn.MaxPool2d(2, 2)                                      lib/models/unet16.py:64
 *   snb_loss_iou_*     BCEWithLogitsLossAndSmoothJaccard / JaccardScore /      lib/losses.py:31-75,
 *                      PixelAccuracy partial sums and integer counts           lib/metrics.py:9-40
 *   snb_pr_curve_*     PRCurveMeter.update                                     lib/train_utils.py:109-125
 *   snb_abn_*          the `inplace_abn` backend calls of InPlaceABN            lib/modules/abn/functions.py:62-122
 *   snb_conv_scatter_* Conv2d(cin, 16, 3, padding=1) of FCDenseNet's DenseLayer  lib/models/tiramisu.py:9-19
 *   snb_bn_train_nhwc, nn.BatchNorm2d / InPlaceABN in train() mode inside         lib/models/linknet.py:16-31,65-90
 *   snb_bn_backward_*  LinkNet34 (batch statistics; their backward)             (torchvision BasicBlock bn1 / bn2)
 *   snb_conv_generic_* autograd of nn.Conv2d / nn.ConvTranspose2d (input and     torch_train.py:186-189
 *                      weight gradients), snb_maxpool3x3s2_backward, snb_loss_grad
 *
 * Conventions: every pointer named `d_*` is a DEVICE pointer owned by the caller; sizes are int64_t; `stream`
 * is a cudaStream_t passed as void* (0 = legacy default stream).  All entry points enqueue work on `stream`
 * and return without synchronising.  Return value: 0 = ok, negative = error (SNB_E_*); the message of the
 * last error on the calling thread is returned by snb_last_error().  There is no CPU fallback anywhere.
 */
#ifndef SNB_B200_H_
#define SNB_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SNB_API __attribute__((visibility("default")))
#else
#define SNB_API
#endif

#define SNB_OK 0
#define SNB_E_INVALID (-1) /* bad argument: maps to ValueError (lib/tiles.py:56-57,81-85,138-139) */
#define SNB_E_SHAPE (-2)   /* shape mismatch: maps to AssertionError (lib/tiles.py:99-100)          */
#define SNB_E_CUDA (-3)    /* CUDA runtime / driver failure: maps to RuntimeError                   */
#define SNB_E_UNSUPPORTED (-4)

SNB_API int snb_version(void);
SNB_API const char* snb_last_error(void);
/* number of SMs of the current device (grid sizing), or negative error */
SNB_API int snb_device_sm_count(void);

/* ------------------------------------------------------------------------------------------------ slicer */
typedef struct snb_slicer snb_slicer;

/* lib/tiles.py:35-96.  tile_step < 1 or > tile_size, or an image_margin that does not tile -> SNB_E_INVALID */
SNB_API int snb_slicer_create(int64_t image_h, int64_t image_w, int64_t tile_size, int64_t tile_step,
                      int64_t image_margin, snb_slicer** out);
SNB_API void snb_slicer_destroy(snb_slicer* s);
/* info[0..7] = margin_left, margin_right, margin_top, margin_bottom, n_tiles, tiles_x, tiles_y, tile_size */
SNB_API int snb_slicer_info(const snb_slicer* s, int64_t info[8]);
/* xy[2*i] = x, xy[2*i+1] = y of crop i, in crop order (y outer, x inner; lib/tiles.py:94-96) */
SNB_API int snb_slicer_crops(const snb_slicer* s, int64_t* xy);

/* Bit-exact ImageSlicer.split for BORDER_REFLECT101 (border_mode 0) or BORDER_CONSTANT (border_mode 1, the
 * border filled with `elem_bytes` bytes taken from `border_value`):
 *   d_src [H][W][C] elements of elem_bytes  ->  d_dst [tile_count][T][T][C], tiles tile_begin.. in crop order. */
SNB_API int snb_split_hwc(const snb_slicer* s, const void* d_src, int64_t channels, int64_t elem_bytes,
                  int border_mode, const void* border_value, void* d_dst, int64_t tile_begin,
                  int64_t tile_count, void* stream);

/* Fused split for the network input: u8 HWC image -> reflect-101 pad -> crop -> per-channel 256-entry LUT
 * (NormalizeImage evaluated on the host exactly as the reference does) -> D4 view `tta` (0..7, order of
 * tta_d4_aug) -> one of the layouts below.  d_lut is float[channels][256].
 *   SNB_LAYOUT_NCHW_F32   float [n][C][T][T]           (what InMemoryDataset + DataLoader produce)
 *   SNB_LAYOUT_PATCH32    bf16  [n][T][T][32]          (first-layer operand: 3x3 neighbourhood x 3 channels,
 *                                                       k = (ky*3+kx)*3 + c, zero outside the tile, k>=27 zero) */
#define SNB_LAYOUT_NCHW_F32 0
#define SNB_LAYOUT_PATCH32 1
#define SNB_LAYOUT_PATCH32_F32 2 /* same rows as PATCH32 but float rounded to TF32 (128 bytes per pixel): TF32 mode */
#define SNB_LAYOUT_NHWC3_BF16 3  /* bf16 [n][T][T][3], packed (6 bytes per pixel): the input of SNB_CONV_FIRST_3X3, which
                                    builds the first layer's operand rows in shared memory; T % 8 == 0, 3 channels */
SNB_API int snb_split_norm_u8(const snb_slicer* s, const uint8_t* d_src, int64_t channels, const float* d_lut,
                      int tta, int layout, void* d_dst, int64_t tile_begin, int64_t tile_count, void* stream);

/* float [n][3][h][w] (the nn.Module.forward input) -> packed NHWC bf16 [n][h][w][3] (SNB_LAYOUT_NHWC3_BF16); w % 8 == 0 */
SNB_API int snb_nchw_f32_to_nhwc3(const float* d_src, int64_t n, int64_t h, int64_t w, void* d_dst, void* stream);

/* float [n][3][H][W] (the nn.Module.forward input) -> PATCH32 rows, same definition as above; out_f32 != 0 writes
 * the SNB_LAYOUT_PATCH32_F32 form */
SNB_API int snb_nchw_f32_to_patch32(const float* d_src, int64_t n, int64_t channels, int64_t h, int64_t w,
                            void* d_dst, int out_f32, void* stream);

/* ImageSlicer.merge (lib/tiles.py:137-161): weighted overlap-add in float64 in crop order, clip of the norm at
 * DBL_EPSILON, divide, cast, crop.  d_tiles [n_tiles][T][T][C] of tile_dtype (SNB_DT_*), d_weight double[T][T].
 * `tta` = number of D4 views per tile (1 or 8): with 8, d_tiles holds [n_tiles][8][T][T][C] float32 views in
 * tta_d4_aug order and the de-augmented mean (tta_d4_deaug: fp32 sum in listed order, * 0.125f) is merged.
 * Outputs (either may be NULL): d_out [H][W][C] of out_dtype; d_mask [H][W][C] u8 = (value > thr) ? 255 : 0
 * evaluated on the float32 result (inria_submit.py:305). */
#define SNB_DT_U8 0
#define SNB_DT_F32 1
#define SNB_DT_F64 2
#define SNB_DT_I64 3
SNB_API int snb_merge(const snb_slicer* s, const void* d_tiles, int tile_dtype, int64_t channels, int tta,
              const double* d_weight, void* d_out, int out_dtype, uint8_t* d_mask, float thr, void* stream);
/* The same merge restricted to the image rows [row_begin, row_begin + row_count): only those rows of d_out / d_mask (still
 * full-image buffers) are written, and only the tiles of the crop rows covering them are read.  This is the per-rank step
 * of the tile-sharded multi-GPU mode: a rank merges the band of rows it owns from its own tiles plus the seam
 * tiles its neighbours sent, in crop order, so the bytes equal the single-GPU merge. */
SNB_API int snb_merge_rows(const snb_slicer* s, const void* d_tiles, int tile_dtype, int64_t channels, int tta,
                   const double* d_weight, void* d_out, int out_dtype, uint8_t* d_mask, float thr, int64_t row_begin,
                   int64_t row_count, void* stream);

/* ------------------------------------------------------------------------------------------ convolutions */
/* Activations are NHWC bf16 inside channel slabs: pixel stride = *_cstride channels, so a producer can write
 * straight into its slot of a concat buffer (torch.cat of lib/models/unet16.py:122-127 never materialises).
 * Weights are pre-packed bf16 (or fp32 in TF32 mode) [phase][tap][Cout][Cin] (K-major), bias is float[Cout]. */
#define SNB_CONV_3X3 0      /* k3 s1 p1: 1 phase x 9 taps, tap = ky*3+kx                     */
#define SNB_CONV_1X1 1      /* k1: 1 phase x 1 tap                                           */
#define SNB_CONVT_4X4_S2 2  /* ConvTranspose2d k4 s2 p1: 4 sub-pixel phases x 4 taps         */
#define SNB_CONV_2X2 4       /* Conv2d k2 s1 p1: 4 taps (dy,dx in {-1,0}), output (h+1) x (w+1); also the body of a
                               stride-2 conv3x3 after space-to-depth (lib/models/linknet.py:62, resnet34 via :39-48) */
#define SNB_CONVT_3X3_S2_FULL 5 /* ConvTranspose2d k3 s2 p0 uncropped: output (2h+1) x (2w+1) (lib/models/linknet.py:58) */
#define SNB_CONV_2X2_ADJ 6   /* the adjoint tap set of SNB_CONV_2X2: 4 taps (dy,dx in {0,+1}), output h x w (valid: (h-1) x (w-1)):
                               input gradients of stride-2 convolutions / of Conv2d k2 p1 (torch_train.py:186-189)            */
#define SNB_CONV_FIRST_3X3 7 /* the networks' first conv3x3 (Cin = 3, padding 1) from the packed 3-channel bf16 tile
                               [n][h][w][3] (SNB_LAYOUT_NHWC3_BF16): the K = 32 operand rows are built in shared memory;
                               cin = in_cstride = 3, cout 32 or 64, weight bf16 [1][cout][32] with k = (ky*3+kx)*3 + c,
                               bias + optional ReLU, w % 8 == 0 (lib/models/unet16.py:71-73, zf_unet.py:60, tiramisu.py:118) */
#define SNB_CONVT_3X3_S2 3  /* ConvTranspose2d k3 s2 p0 cropped to [0,2h) x [0,2w) (lib/models/tiramisu.py:62-90):
                               4 phases x 4 tap slots, unused slots carry zero weights       */

/* storage / arithmetic type of activations and weights of a convolution */
#define SNB_CONV_BF16 0     /* bf16 storage, tcgen05 kind::f16, fp32 accumulation (default)  */
#define SNB_CONV_TF32 1     /* fp32 storage rounded to TF32, tcgen05 kind::tf32, fp32 accumulation: the "fp32 mode"
                               whose probabilities stay within 1e-4 of the fp32 reference    */

typedef struct snb_conv_desc {
  int32_t kind;          /* SNB_CONV_*                                                      */
  int32_t relu;          /* apply ReLU after bias                                           */
  int64_t n, h, w;       /* input batch and spatial size                                    */
  int64_t cin;           /* input channels read (multiple of 32)                            */
  int64_t in_cstride;    /* pixel stride of the input slab, in channels                     */
  int64_t cout;          /* output channels (multiple of 32)                                */
  int64_t out_cstride;   /* pixel stride of the output slab, in channels                    */
  const void* d_in;      /* bf16, first channel read                                        */
  void* d_out;           /* bf16, first channel written (may be NULL when the head is fused) */
  const void* d_weight;  /* packed bf16                                                     */
  const float* d_bias;   /* float[cout]                                                     */
  /* optional fused 1x1 head (cout must be 32): out = [sigmoid](dot(relu(conv), head_w) + head_b) */
  const float* d_head_w; /* float[32] or NULL                                               */
  float head_b;
  int32_t head_sigmoid;
  float* d_head_out;     /* float [n][h][w]                                                 */
  /* optional fused nn.MaxPool2d(2,2) of the (bias, ReLU) output: a second bf16 slab [n][h/2][w/2] written from the
   * same epilogue (conv3x3 only; h, w even) */
  void* d_pool_out;      /* bf16, first channel written, or NULL                            */
  int64_t pool_cstride;  /* pixel stride of the pooled slab, in channels                    */
  /* nn.Upsample(scale_factor=2) (nearest) fused into the store: d_out is then a slab of [n][2h][2w] pixels and every
   * output pixel is written to its 2x2 block (lib/models/zf_unet.py:42,78-90); conv3x3 / conv1x1 only */
  /* optional pre-activation of the INPUT, y = max(x * scale[c] + shift[c], 0) per input channel, applied inside the
   * kernel to the operand tiles before the multiply (FCDenseNet's per-consumer BatchNorm2d(eval) + ReLU,
   * lib/models/tiramisu.py:12-13); the conv's zero padding applies after it.  bf16 conv3x3 with cout == 32 only */
  const float* d_pre_scale; /* float[cin] or NULL                                           */
  const float* d_pre_shift; /* float[cin]                                                   */
  /* epilogue extras of the ResNet / LinkNet blocks (lib/models/linknet.py:39-48,77-80): an activation slope
   * (relu != 0: y = x > 0 ? x : act_slope * x, so 0 = ReLU, 0.01 = the InPlaceABN / nn.LeakyReLU default) and a residual
   * tensor with the output's pixel grid, added before (res_after_act = 0, BasicBlock) or after the activation (skip) */
  float act_slope;
  int32_t res_after_act;
  const void* d_residual;   /* same element type as d_out, first channel read, or NULL                */
  int64_t res_cstride;
  int32_t valid;            /* conv3x3: 1 = no padding, output (h-2) x (w-2) (lib/models/linknet.py:60), 2 = "full", output
                               (h+2) x (w+2) (the input gradient of the valid conv); conv2x2: 1 = padding on top/left only,
                               output h x w (a stride-2 conv3x3 after space-to-depth); conv2x2-adjoint: 1 = (h-1) x (w-1) */
  int32_t out_upsample2x;
  int32_t dtype;         /* SNB_CONV_BF16 / SNB_CONV_TF32: element type of d_in, d_out, d_pool_out and d_weight
                            (bias and the head stay float); channel counts are multiples of 64 bytes / element size */
} snb_conv_desc;

typedef struct snb_conv snb_conv;
SNB_API int snb_conv_create(const snb_conv_desc* desc, snb_conv** out);
SNB_API int snb_conv_launch(const snb_conv* c, void* stream);
SNB_API void snb_conv_destroy(snb_conv* c);
/* algorithmic FLOPs of one launch (2*MACs, padding excluded), for roofline bookkeeping */
SNB_API double snb_conv_flops(const snb_conv* c);
/* Redirect the fused-head output (float [n][h][w]) of a conv created with d_head_w != NULL: the tiled predictor lets the
 * last layer write straight into its probability-tile buffer (the reference copies every batch, inria_submit.py:251-253).
 * Takes effect at the next snb_conv_launch; a CUDA graph keeps the pointer that was set when the launch was captured. */
SNB_API int snb_conv_set_head_out(snb_conv* c, float* d_head_out);

/* Weight gradient of a convolution of the kinds above on the tensor cores (autograd of nn.Conv2d / nn.ConvTranspose2d,
 * torch_train.py:186-189): with X the forward input and dY the gradient of the forward output (both NHWC bf16 slabs),
 *   d_dweight[(phase * taps + tap)][co][ci] += sum over pixels of dY_phase[pixel][co] * X[pixel + tap offset][ci]
 * in the tap / phase order of the packed forward weights, as float [phases * taps][dw_cout][dw_cin] (dw_cout >= cout,
 * dw_cin >= cin; only the cout x cin corner is touched).  The call ACCUMULATES (split-K float atomics): zero the buffer
 * first.  kind / valid / n / h / w describe the forward convolution (h, w = its input size). */
typedef struct snb_wgrad_desc {
  int32_t kind, valid;
  int64_t n, h, w;
  int64_t cin, in_cstride;      /* forward input channels read, pixel stride of its slab    */
  int64_t cout, dout_cstride;   /* forward output channels, pixel stride of the dY slab     */
  const void* d_in;             /* bf16 X, first channel                                    */
  const void* d_dout;           /* bf16 dY, first channel (the forward's output extent)     */
  float* d_dweight;
  int64_t dw_cout, dw_cin;
} snb_wgrad_desc;
typedef struct snb_wgrad snb_wgrad;
SNB_API int snb_wgrad_create(const snb_wgrad_desc* desc, snb_wgrad** out);
SNB_API int snb_wgrad_launch(const snb_wgrad* c, void* stream);
SNB_API void snb_wgrad_destroy(snb_wgrad* c);
SNB_API double snb_wgrad_flops(const snb_wgrad* c);

/* nn.MaxPool2d(2,2) on NHWC slabs of bf16 (elem_bytes 2) or float (elem_bytes 4); h, w even; channel counts and
 * strides multiples of 16 bytes */
SNB_API int snb_maxpool2x2(const void* d_in, int64_t n, int64_t h, int64_t w, int64_t channels, int64_t in_cstride,
                   void* d_out, int64_t out_cstride, int elem_bytes, void* stream);

/* conv3x3 (stride 1, padding 1) with exactly 16 output channels -- FCDenseNet's growth-rate layers
 * (lib/models/tiramisu.py:9-19: BatchNorm -> ReLU -> Conv2d(cin, 16, 3, padding=1)) -- as one N = 144 GEMM per tile
 * (the nine taps live in the GEMM's N dimension; the epilogue adds the nine shifted partial planes).  NHWC bf16 slabs.
 *   d_weight: bf16 [144][cin], row tap * 16 + co holds W[co][:, ky, kx], tap = ky * 3 + kx; cin % 32 == 0
 *   d_pre_scale / d_pre_shift: float[cin] pre-activation y = relu(x * scale + shift) applied to the operand (or NULL)
 *   d_out: first of the 16 output channels inside the destination slab (32-byte aligned, out_cstride % 16 == 0);
 *          out = conv + bias, no activation */
typedef struct snb_conv_scatter snb_conv_scatter;
SNB_API int snb_conv_scatter_create(const void* d_in, int64_t n, int64_t h, int64_t w, int64_t cin, int64_t in_cstride,
                            const void* d_weight, const float* d_bias, const float* d_pre_scale,
                            const float* d_pre_shift, void* d_out, int64_t out_cstride, snb_conv_scatter** out);
SNB_API int snb_conv_scatter_launch(const snb_conv_scatter* c, void* stream);
SNB_API void snb_conv_scatter_destroy(snb_conv_scatter* c);
SNB_API double snb_conv_scatter_flops(const snb_conv_scatter* c);

/* ResNet-34 encoder helpers of LinkNet34 (lib/models/linknet.py:39-48), NHWC bf16 slabs, channels % 8 == 0:
 *   snb_space_to_depth2: out[n][y][x][(py*2+px)*C + c] = in[n][2y+py][2x+px][c] (zero beyond an odd h / w; the output has
 *                        ceil(h/2) x ceil(w/2) pixels) -- turns the stride-2 conv3x3 / conv1x1 of a down-sampling BasicBlock
 *                        into SNB_CONV_2X2 / SNB_CONV_1X1 launches, and the gradients of stride-2 transposed convolutions
 *                        into stride-1 convolutions over the blocked gradient
 *   snb_depth_to_space2: the inverse, out[n][2y+py][2x+px][c] (+)= in[n][y][x][(py*2+px)*C + c]; (h, w, channels)
 *                        describe the OUTPUT, accumulate != 0 adds to what the output holds (gradient fan-in)
 *   snb_scale_nc_nhwc:   out[n][pixel][c] = in[n][pixel][c] * scale[n][c]: nn.Dropout2d (lib/models/linknet.py:57,83) with
 *                        the keep mask / (1 - p) drawn by the caller, forward and backward
 *   snb_maxpool3x3s2:    nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
 *   snb_stem7x7_rows:    float [n][C][h][w] -> bf16 rows [n][ceil(h/2)][ceil(w/2)][k_pad] of the stride-2 7x7 p3 stem,
 *                        k = (ky*7+kx)*C + c, zero for k >= 49*C: the stem becomes a conv1x1 over these rows */
SNB_API int snb_space_to_depth2(const void* d_in, int64_t n, int64_t h, int64_t w, int64_t channels, int64_t in_cstride,
                        void* d_out, int64_t out_cstride, void* stream);
SNB_API int snb_depth_to_space2(const void* d_in, int64_t n, int64_t h, int64_t w, int64_t channels, int64_t in_cstride,
                        void* d_out, int64_t out_cstride, int accumulate, void* stream);
SNB_API int snb_scale_nc_nhwc(const void* d_in, int64_t n, int64_t hw, int64_t channels, int64_t in_cstride,
                      const float* d_scale, void* d_out, int64_t out_cstride, void* stream);
SNB_API int snb_maxpool3x3s2(const void* d_in, int64_t n, int64_t h, int64_t w, int64_t channels, int64_t in_cstride,
                     void* d_out, int64_t out_cstride, void* stream);

/* Multi-segment index gather, one launch for all layers of a plan: for every segment of the DEVICE table d_segs,
 * dst[i] = idx[i] >= 0 ? src[idx[i]] : 0 for i < count (dst float, or bf16 when dst_bf16 != 0).  first_block = the number
 * of 1024-element blocks of all earlier segments; total_blocks = their sum over the table.  Re-packs fp32 parameters into
 * the [phase][tap][Cout][Cin] operands of the forward / input-gradient convolutions after an optimiser step and scatters
 * packed weight gradients (snb_wgrad_*) back into the parameters' layouts (torch_train.py:186-190). */
typedef struct snb_gather_seg {
  const float* src;
  void* dst;
  const int32_t* idx;
  int64_t count;
  int64_t first_block;
} snb_gather_seg;
SNB_API int snb_gather_segments(const snb_gather_seg* d_segs, int64_t n_segs, int64_t total_blocks, int dst_bf16,
                        void* stream);
SNB_API int snb_stem7x7_rows(const float* d_src, int64_t n, int64_t channels, int64_t h, int64_t w, void* d_dst,
                     int64_t k_pad, void* stream);

/* Generic convolution gradients on NHWC bf16 slabs (the backward half of the LinkNet34 training step; torch_train.py:186-189
 * runs them through autograd + cuDNN).  One geometry: small[n,oy,ox,co] = sum big[n, oy*s+ky-p, ox*s+kx-p, ci] * W[co][ci][ky][kx].
 * nn.Conv2d: big = input, small = output.  nn.ConvTranspose2d(Cin_t, Cout_t): big = its OUTPUT (ci = Cout_t), small = its
 * INPUT (co = Cin_t), weight [Cin_t][Cout_t][kh][kw] = W[co][ci][kh][kw] unchanged.  Weights and weight gradients are
 * float tensors in the PyTorch layout.
 *   snb_conv_generic_fwd   small = conv(big) (+ bias[co])            -> the input gradient of a ConvTranspose2d
 *   snb_conv_generic_dgrad big = conv^T(small) (+ bias[ci])          -> the input gradient of a Conv2d
 *   snb_conv_generic_wgrad dW[co][ci][ky][kx] = sum_p small[p][co] * big[p*s + k - p][ci]   (zeroed, then fp32 atomics)
 * small_h / small_w must equal (big + 2p - k) / s + 1 (SNB_E_SHAPE otherwise). */
typedef struct snb_conv_geom {
  int64_t n, big_h, big_w, big_c, big_cstride;
  int64_t small_h, small_w, small_c, small_cstride;
  int64_t kh, kw, stride, pad;
} snb_conv_geom;
SNB_API int snb_conv_generic_fwd(const snb_conv_geom* g, const void* d_big, const float* d_weight, const float* d_bias,
                         void* d_small, void* stream);
SNB_API int snb_conv_generic_dgrad(const snb_conv_geom* g, const void* d_small, const float* d_weight, const float* d_bias,
                           void* d_big, void* stream);
SNB_API int snb_conv_generic_wgrad(const snb_conv_geom* g, const void* d_big, const void* d_small, float* d_dweight,
                           void* stream);

/* InPlaceABN (lib/modules/abn/bn.py:47-103, functions.py:62-122): in-place activated batch norm on a contiguous NCHW
 * float tensor, hw = H * W.  The reference delegates the arithmetic to the external `inplace_abn` extension
 * (mean_var, forward, leaky_relu_forward/backward, elu_forward/backward, edz_eydz, backward; functions.py:46-118),
 * which is not vendored and has no pinned version: these entry points replace exactly that call surface.
 *   activation: 0 none, 1 leaky_relu(slope), 2 elu.  d_weight / d_bias NULL = not affine.
 *   forward : training != 0 -> batch mean / biased variance into d_mean / d_var (float[c]) and the momentum update of the
 *             running statistics with the unbiased variance (functions.py:84-85); else the running statistics are used.
 *             y = (x - mean) / sqrt(var + eps) * (|weight| + eps) + bias, activation, written over x.
 *   backward: z = forward output, dz = its gradient, d_var = the variance the forward used.  dx may alias dz.
 *             eval mode reproduces the reference (edz = eydz = 0: functions.py:110-112), so dweight = dbias = 0 there.
 *   d_workspace: double[2 * c] scratch. */
SNB_API int snb_abn_forward(float* d_x, int64_t n, int64_t c, int64_t hw, const float* d_weight, const float* d_bias,
                    float* d_running_mean, float* d_running_var, int training, float momentum, float eps,
                    int activation, float slope, float* d_mean, float* d_var, double* d_workspace, void* stream);
SNB_API int snb_abn_backward(const float* d_z, const float* d_dz, int64_t n, int64_t c, int64_t hw, const float* d_var,
                     const float* d_weight, const float* d_bias, int training, float eps, int activation, float slope,
                     float* d_dx, float* d_dweight, float* d_dbias, double* d_workspace, void* stream);

/* Training-mode BatchNorm2d / InPlaceABN on an NHWC bf16 slab (LinkNet34 in train() mode: torchvision BasicBlock bn1/bn2,
 * lib/models/linknet.py:16-31 abn1-3): batch statistics over `pixels` = N*H*W, running statistics updated in place with
 * momentum and the unbiased variance, then out = act(x * scale + shift [+ residual]) [+ residual];
 * abn != 0 uses gamma = |weight| + eps (the InPlaceABN backend); act_slope >= 0: leaky-ReLU with that slope (0 = ReLU),
 * act_slope < 0: no activation.  d_scale / d_shift / d_mean / d_var: float[channels] outputs (the fused normalisation and
 * the statistics, kept for the backward pass).  d_workspace: double[3 * channels + 2], zeroed ONCE by the caller and left
 * zeroed by every call (the last block of the statistics kernel finalises and clears it: two launches per call, no
 * memset); calls that may overlap need their own workspace.  channels % 8 == 0, <= 2048. */
SNB_API int snb_bn_train_nhwc(const void* d_in, int64_t pixels, int64_t channels, int64_t in_cstride, const float* d_gamma,
                      const float* d_beta, int abn, float eps, float momentum, float* d_running_mean,
                      float* d_running_var, float act_slope, const void* d_residual, int64_t res_cstride,
                      int res_after_act, void* d_out, int64_t out_cstride, float* d_scale, float* d_shift, float* d_mean,
                      float* d_var, double* d_workspace, void* stream);

/* Backward of snb_bn_train_nhwc (out = act(x * scale + shift + r_before) + r_after; the gradient of r_after is d_dout
 * itself).  d_x = the raw convolution output the forward normalised, d_scale / d_shift / d_mean / d_var = what the forward
 * returned, d_res_before = the residual added before the activation (or NULL).  Outputs: d_dx (gradient of x, bf16 slab),
 * d_dres (gradient of r_before = dz, or NULL), d_dgamma / d_dbeta (float[channels]; dgamma carries sign(weight) when
 * abn != 0, as the InPlaceABN backend does), d_workspace double[3 * channels + 2] (zeroed once, left zeroed: see above). */
SNB_API int snb_bn_backward_nhwc(const void* d_x, int64_t x_cstride, const void* d_dout, int64_t dout_cstride, int64_t pixels,
                         int64_t channels, const float* d_scale, const float* d_shift, const float* d_mean,
                         const float* d_var, const float* d_gamma, int abn, float eps, float act_slope,
                         const void* d_res_before, int64_t res_cstride, void* d_dx, int64_t dx_cstride, void* d_dres,
                         int64_t dres_cstride, float* d_dgamma, float* d_dbeta, double* d_workspace, void* stream);
/* out[c] = sum over pixels of a bf16 slab (bias gradients); channels % 8 == 0; d_workspace double[3 * channels + 2]
 * (zeroed once, left zeroed) */
SNB_API int snb_channel_sum_nhwc(const void* d_in, int64_t pixels, int64_t channels, int64_t in_cstride, float* d_out,
                         double* d_workspace, void* stream);
/* backward of nn.MaxPool2d(3, 2, 1) (lib/models/linknet.py:44): d_in = the forward input (arg-max recomputed, first
 * maximum wins as in PyTorch), d_dout = gradient of the pooled tensor, d_din = gradient of the input */
SNB_API int snb_maxpool3x3s2_backward(const void* d_in, int64_t n, int64_t h, int64_t w, int64_t channels, int64_t in_cstride,
                              const void* d_dout, int64_t dout_cstride, void* d_din, int64_t din_cstride, void* stream);
/* elementwise on bf16 slabs (channels % 8 == 0): mode 0 out = a + b (gradient fan-in); mode 1 out = a * (b > 0 ? 1 : slope) (gradient a through
 * a leaky-ReLU whose output is b) */
SNB_API int snb_ew_nhwc(const void* d_a, int64_t a_cstride, const void* d_b, int64_t b_cstride, void* d_out, int64_t out_cstride,
                int64_t pixels, int64_t channels, int mode, float slope, void* stream);

/* Pre-activation BatchNorm2d(eval) + ReLU of FCDenseNet's DenseLayer / TransitionDown (lib/models/tiramisu.py:12-13,
 * 50-51): out[.., c] = max(in[.., c] * scale[c] + shift[c], 0) for c < channels, 0 for channels <= c < channels_pad
 * (scale = gamma / sqrt(var + eps), shift = beta - mean * scale, float).  NHWC bf16 slabs, channels % 8 == 0. */
SNB_API int snb_bn_relu_nhwc(const void* d_in, int64_t n, int64_t h, int64_t w, int64_t channels, int64_t in_cstride,
                     const float* d_scale, const float* d_shift, void* d_out, int64_t channels_pad,
                     int64_t out_cstride, void* stream);

/* bf16 NHWC slab -> float NCHW (leaving the network through the nn.Module interface) */
SNB_API int snb_nhwc_bf16_to_nchw_f32(const void* d_in, int64_t n, int64_t h, int64_t w, int64_t channels,
                              int64_t in_cstride, float* d_out, void* stream);

/* ----------------------------------------------------------------------------------------- loss / metrics */
/* Every reduction below is ONE kernel launch: blocks park their partials in `d_workspace`, the last block to finish adds
 * them in a fixed order (deterministic for a given size) and restores the workspace.  The workspace is
 * snb_reduce_workspace_bytes() bytes, 64-byte aligned, zeroed ONCE by the caller and then only touched by these calls
 * ("zero at rest"); calls that may overlap on different streams need different workspaces. */
SNB_API int64_t snb_reduce_workspace_bytes(void);

/* One pass over logits/targets (lib/losses.py:18-101, lib/metrics.py:13-36):
 *   p = sigmoid(x); z = logsigmoid(x); bce_i = max(z,0) - z*t + log1p(exp(-|z|))   (the reference's double squash)
 *   d_sums[0..4]   = sum bce_i, sum p*t, sum p, sum t, sum focal_i                  (double[5])
 *                    focal_i = (1 - exp(-bce_i))^gamma * bce_i (FocalLossBinary, lib/losses.py:78-101) when
 *                    focal_gamma >= 0, else d_sums[4] = 0
 *   d_counts[0..3] = tp, fp, fn, tn at p > 0.5 (float32 compare)                    (int64[4])
 *   d_elem_bce     = optional float[n]: bce_i per element (BCEWithSigmoidLoss(reduce=False), lib/losses.py:46-53)
 * targets of dtype SNB_DT_I64 / SNB_DT_U8 / SNB_DT_F32.  Outputs are overwritten. */
SNB_API int snb_loss_iou_reduce(const float* d_logits, const void* d_targets, int target_dtype, int64_t n,
                        float focal_gamma, float* d_elem_bce, double* d_sums, int64_t* d_counts, void* d_workspace,
                        void* stream);

/* Gradient of a fused loss with respect to the logits (what autograd computes through lib/losses.py:18-101):
 *   grad[i] = g * d/dx_i [ c_bce * sum_i bce_i + c_focal * sum_i focal_i + c_jac * (1 - A / D) ],
 *   A = sum p*t + smooth_num, D = sum p + sum t - sum p*t + smooth_den,
 * with d_sums = the double[5] output of snb_loss_iou_reduce on the same tensors (no host round trip) and g = the
 * upstream gradient on the device: NULL = 1, a scalar, or one float per element when grad_out_per_element != 0.
 * bce_jaccard: c_bce = bce_weight / ((bce_weight + jaccard_weight) * n), c_jac = jaccard_weight / (bce_weight +
 * jaccard_weight), smooth_num = smooth_den = 100; JaccardLoss: smooth_num = 0, smooth_den = 1e-7. */
SNB_API int snb_loss_grad(const float* d_logits, const void* d_targets, int target_dtype, int64_t n, const double* d_sums,
                  const float* d_grad_out, int grad_out_per_element, float c_bce, float c_focal, float focal_gamma,
                  float c_jac, float smooth_num, float smooth_den, float* d_grad_logits, void* stream);

/* Same integer counts from probabilities already on the device (mask parity path): pred = prob > thr. */
SNB_API int snb_confusion_counts(const float* d_probs, const void* d_targets, int target_dtype, int64_t n, float thr,
                         int64_t* d_counts, void* d_workspace, void* stream);

/* PRCurveMeter.update (lib/train_utils.py:109-125): for every threshold k (float32, ascending),
 * pred = sigmoid(x) > thr[k]; adds tp/tn/fp/fn counts into uint64 arrays of length n_thr (accumulating). */
SNB_API int snb_pr_curve_update(const float* d_logits, const void* d_targets, int target_dtype, int64_t n,
                        const float* d_thresholds, int64_t n_thr, uint64_t* d_tp, uint64_t* d_tn,
                        uint64_t* d_fp, uint64_t* d_fn, void* d_workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SNB_B200_H_ */
