/*
 * ssm_b200.h -- C ABI of libssm_b200.so: the B200 (sm_100a) implementation of Super SloMo's
 * per-pixel intermediate-frame synthesis path.
 *
 * The reference (SreenivasVRao/SuperSloMo-VideoInterpolation-PyTorch) has no FFI or operator
 * registry; its boundary for this path is the Python call surface of scripts/models.  Each
 * entry point below names the reference function it replaces.  The Python mirror of that call
 * surface (same names, arguments and error behaviour) lives in
 * superslomo-videointerpolation-pytorch_b200/ and reaches this ABI through ctypes; see
 * INTEGRATION.md for the binding a maintainer of the reference would add.
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer on the current CUDA device unless the entry point's name
 *    ends in _host.  The library never allocates, frees or retains caller memory; outputs are
 *    caller-allocated; inputs are read-only.
 *  - Calls are asynchronous: work is enqueued on `stream` (a cudaStream_t passed as void*, NULL =
 *    legacy default stream) and the call returns without synchronising.
 *  - Re-entrant, no global mutable state apart from the thread-local error string.
 *  - Return value: 0 = ok; < 0 = argument error (SSM_ERR_*); > 0 = cudaError_t of a failed
 *    launch.  ssm_last_error() returns a message for the calling thread's last failure.  No C++
 *    exception crosses the ABI.
 *  - There is no CPU fallback: without a CUDA device every compute entry point fails.
 *
 * Tensors are described by ssm_tensor: planes are dense (element (y, x) of a plane is at
 * y*W + x) while pair, timestep and channel strides are free, so channel-sliced views such as
 * input_tensor[:, 6:10] (flow_interpolation.py:402-403) are passed without a copy.
 * Strides are in ELEMENTS of the storage dtype.
 */
#ifndef SSM_B200_H
#define SSM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSM_ABI_VERSION 8

/* storage dtype of image/flow/output tensors; arithmetic is always fp32 */
#define SSM_DTYPE_F32  0
#define SSM_DTYPE_BF16 1

/* How the division by max(W-1,1) of layers.py:112-113 is rounded (SURVEY.md finding 3b):
 * DIV = IEEE division, bit-matches the reference on CPU; RCP = multiply by fp32 1/(W-1),
 * bit-matches the reference on CUDA (torch's div-by-Python-scalar kernel). */
#define SSM_COORD_DIV 0
#define SSM_COORD_RCP 1

#define SSM_OK                 0
#define SSM_ERR_NULL          -1   /* a required pointer is NULL */
#define SSM_ERR_SHAPE         -2   /* B, N, C, H or W out of range */
#define SSM_ERR_DTYPE         -3   /* unknown dtype / coord_mode */
#define SSM_ERR_ALIGN         -4   /* pointer or stride not aligned for the dtype */
#define SSM_ERR_WORKSPACE     -5   /* workspace missing or too small */
#define SSM_ERR_UNSUPPORTED   -6

typedef struct ssm_tensor {
    void*   data;      /* first element of (pair 0, timestep 0, channel 0) */
    int64_t stride_b;  /* elements between consecutive frame pairs / samples */
    int64_t stride_n;  /* elements between consecutive timesteps (ignored where there is no timestep axis) */
    int64_t stride_c;  /* elements between consecutive channels */
} ssm_tensor;

int         ssm_version(void);
const char* ssm_last_error(void);

/* Self-test (known-answer): compares the kernels' 5-instruction constant-divisor division with the
 * IEEE division for the divisor max(size-1,1) over every finite fp32 dividend s.  mismatches_device
 * points at three zero-initialised device counters: [0] += dividends for which the normalised
 * coordinate rn(s/d - 1) -- the value the path consumes (layers.py:112) -- differs (expected 0);
 * [1] += dividends whose raw quotient differs (only quotients in the denormal range, where the
 * residual underflows, may); [2] = bits of one such dividend. */
int ssm_selftest_division(int size, unsigned long long* mismatches_device, void* stream);

/* ---- a1: layers.warp(x, flo)  [reference scripts/models/layers.py:73-120] --------------------
 * out[b,c,y,x] = bilinear(img[b,c], x + flow[b,0,y,x], y + flow[b,1,y,x]), zeros outside,
 * align_corners=True.  img/out: B x C x H x W, flow: B x 2 x H x W. */
int ssm_warp_fwd(const ssm_tensor* img, const ssm_tensor* flow, const ssm_tensor* out,
                 int B, int C, int H, int W, int dtype, int coord_mode, void* stream);

/* Backward of warp (what autograd derives for layers.py:73-120).  grad_img and/or grad_flow may
 * be NULL (not wanted).  grad_img needs a workspace of ssm_warp_bwd_workspace_bytes(B, C, H, W)
 * (16-byte aligned); it is accumulated deterministically (bit-identical run to run). */
int ssm_warp_bwd(const ssm_tensor* grad_out, const ssm_tensor* img, const ssm_tensor* flow,
                 const ssm_tensor* grad_img, const ssm_tensor* grad_flow,
                 int B, int C, int H, int W, int dtype, int coord_mode,
                 void* workspace, size_t workspace_bytes, void* stream);

/* The same with the image given as an RGBx copy (C = 3 only): ssm_pack_image writes img3 (B x 3 x H x W planar) as
 * B x H x W x 4 elements of the storage dtype (ssm_packed_image_bytes, 16-byte aligned); each bilinear tap of the
 * warp is then one 16-byte request instead of three 4-byte ones.  Worth it when an image is warped more than once
 * (SSMLosses.get_warp_loss warps each frame by two flows, losses.py:152-162).  The image gradient does not read
 * the image, so ssm_warp_bwd_packed takes no planar img. */
size_t ssm_packed_image_bytes(int B, int H, int W, int dtype);
int ssm_pack_image(const ssm_tensor* img3, void* packed, int B, int H, int W, int dtype, void* stream);
int ssm_warp_fwd_packed(const void* packed, const ssm_tensor* flow, const ssm_tensor* out,
                        int B, int H, int W, int dtype, int coord_mode, void* stream);
int ssm_warp_bwd_packed(const ssm_tensor* grad_out, const void* packed, const ssm_tensor* flow,
                        const ssm_tensor* grad_img, const ssm_tensor* grad_flow,
                        int B, int H, int W, int dtype, int coord_mode,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---- frame re-layout (no reference counterpart; an internal staging step of a2/a4) -----------
 * The gathers of a2 and a3+a4 are bound by L1 request throughput when the three colour planes of
 * a frame are fetched separately.  ssm_pack_frames copies img6 (B x 6 x H x W planar) into a
 * pixel-interleaved RGBx buffer `packed` (B x 2 x H x W x 4 elements of the storage dtype,
 * ssm_packed_frames_bytes(B, H, W, dtype) bytes, 16-byte aligned) so each bilinear tap is one
 * request.  It is optional: every entry point below takes `packed` and gathers from the planar
 * frames when it is NULL.  It pays for itself from N >= 2 timesteps per pair. */
size_t ssm_packed_frames_bytes(int B, int H, int W, int dtype);
int ssm_pack_frames(const ssm_tensor* img6, void* packed, int B, int H, int W, int dtype, void* stream);

/* ---- a2: FlowInterpolationModel.compute_inputs(img_tensor, flow_pred_tensor, t)
 *      [reference scripts/models/flow_interpolation.py:338-372], batched over N timesteps so
 *      that the loop of superslomo_r.py:167-179 and its torch.stack become one launch.
 * img6:  B x 6 x H x W        (channels 0-2 = I0, 3-5 = I1)
 * flow4: B x 4 x H x W        (0-1 = F01, 2-3 = F10)
 * t:     B*N fp32 values, t[b*N + n] in (0,1)
 * out16: B x N x 16 x H x W   [I1, g(I1,F_t1), F_t1, F_t0, g(I0,F_t0), I0] */
int ssm_flow_pack_fwd(const ssm_tensor* img6, const void* packed, const ssm_tensor* flow4, const float* t,
                      const ssm_tensor* out16, int B, int N, int H, int W,
                      int dtype, int coord_mode, void* stream);

/* Backward of a2.  grad16: B x N x 16 x H x W.  grad_flow4 (B x 4 x H x W, summed over the N
 * timesteps) and grad_img6 (B x 6 x H x W) may be NULL.  grad_img6 needs a workspace of
 * ssm_flow_pack_bwd_workspace_bytes(B, N, H, W). */
int ssm_flow_pack_bwd(const ssm_tensor* grad16, const ssm_tensor* img6, const void* packed,
                      const ssm_tensor* flow4, const float* t, const ssm_tensor* grad_flow4, const ssm_tensor* grad_img6,
                      int B, int N, int H, int W, int dtype, int coord_mode,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---- a3 + a4: extract_outputs + compute_output_image(img_tensor, input_tensor, output_tensor, t)
 *      [reference scripts/models/flow_interpolation.py:374-429], batched over N timesteps
 *      (replaces the loop of superslomo_r.py:215-238).
 * img6:   B x 6 x H x W
 * flows4: B x N x 4 x H x W   = input_tensor[:, 6:10] (F_t1 then F_t0), usually a strided view
 * out5:   B x N x 5 x H x W   stage-2 U-Net output (visibility logit, dF_t1, dF_t0)
 * out3:   B x N x 3 x H x W   fused frame */
int ssm_fuse_fwd(const ssm_tensor* img6, const void* packed, const ssm_tensor* flows4, const ssm_tensor* out5,
                 const float* t, const ssm_tensor* out3, int B, int N, int H, int W,
                 int dtype, int coord_mode, void* stream);

/* Backward of a3 + a4.  grad3: B x N x 3 x H x W.  grad_out5 (B x N x 5), grad_flows4
 * (B x N x 4: the gradient of input_tensor[:, 6:10]; the other 12 channels of that gradient are
 * zero and are the caller's to fill) and grad_img6 (B x 6) may each be NULL.  grad_img6 needs a
 * workspace of ssm_fuse_bwd_workspace_bytes(B, N, H, W). */
int ssm_fuse_bwd(const ssm_tensor* grad3, const ssm_tensor* img6, const void* packed,
                 const ssm_tensor* flows4, const ssm_tensor* out5, const float* t, const ssm_tensor* grad_out5,
                 const ssm_tensor* grad_flows4, const ssm_tensor* grad_img6,
                 int B, int N, int H, int W, int dtype, int coord_mode,
                 void* workspace, size_t workspace_bytes, void* stream);

/* ---- a2 -> a4 shortcut for callers that hold flow_pred_tensor (the timestep-batched loop) -------
 * Same results as ssm_fuse_fwd / ssm_fuse_bwd given input_tensor = compute_inputs(img, flow4, t):
 * the estimated flows input_tensor[:, 6:10] (flow_interpolation.py:353,356,402-413) are recomputed
 * in-kernel from flow4 (B x 4 x H x W) and t with the arithmetic of ssm_flow_pack_fwd instead of
 * being read back, which removes 4 of the 12 streamed channels per timestep.  In the backward the
 * gradient goes straight to grad_flow4 (B x 4 x H x W, summed over the N timesteps in registers
 * through the coefficients of :353,356): the B x N x 16 gradient of input_tensor that autograd
 * builds for the reference (12 zero channels + 4) never exists. */
int ssm_fuse_flow_fwd(const ssm_tensor* img6, const void* packed, const ssm_tensor* flow4, const ssm_tensor* out5,
                      const float* t, const ssm_tensor* out3, int B, int N, int H, int W,
                      int dtype, int coord_mode, void* stream);
int ssm_fuse_flow_bwd(const ssm_tensor* grad3, const ssm_tensor* img6, const void* packed,
                      const ssm_tensor* flow4, const ssm_tensor* out5, const float* t, const ssm_tensor* grad_out5,
                      const ssm_tensor* grad_flow4, const ssm_tensor* grad_img6,
                      int B, int N, int H, int W, int dtype, int coord_mode,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---- the layouts either side of the stage-2 U-Net ---------------
 * ssm_flow_pack_fwd_nhwc: a2 as ssm_flow_pack_fwd, but the stage-2 input is written channels-last,
 * out16_nhwc = B x N x H x W x 16 elements of out_dtype (32-byte aligned), which is the layout -- and with
 * out_dtype = SSM_DTYPE_BF16 the dtype -- conv1a of the stage-2 U-Net [flow_interpolation.py:36-38]
 * consumes when the U-Net runs channels-last under bf16 autocast: no conversion pass runs between a2
 * and the U-Net.  Values are those of ssm_flow_pack_fwd rounded once (RN) to out_dtype.  Inputs of
 * dtype fp32 give out_dtype fp32 or bf16; bf16 inputs give bf16.
 * ssm_fuse_flow_fwd_mixed: a3+a4 as ssm_fuse_flow_fwd with out5 stored in out5_dtype (bf16 straight
 * from final_conv [flow_interpolation.py:149-157] under autocast) while frames, flows and the result
 * stay in `dtype` (fp32): the result equals ssm_fuse_flow_fwd on out5 converted to fp32. */
int ssm_flow_pack_fwd_nhwc(const ssm_tensor* img6, const void* packed, const ssm_tensor* flow4, const float* t,
                           void* out16_nhwc, int B, int N, int H, int W,
                           int dtype, int out_dtype, int coord_mode, void* stream);
int ssm_fuse_flow_fwd_mixed(const ssm_tensor* img6, const void* packed, const ssm_tensor* flow4,
                            const ssm_tensor* out5, int out5_dtype, const float* t, const ssm_tensor* out3,
                            int B, int N, int H, int W, int dtype, int coord_mode, void* stream);

/* ---- a4 + a9: compute_output_image fused with the loss front-end of SSMLosses
 *      [reference scripts/models/losses.py:104-170, 213-233; SURVEY.md section 8(f) rank 1].
 * Besides out3 (the fused frame, as ssm_fuse_flow_fwd) it returns, per pair b, the sums over all
 * 3*H*W elements of the three L1 maps the reference builds with ~30 more ATen launches and two more
 * warp calls per window:  sums[b*(2N+1) + 2n]   = sum |out3[b,n] - target[b,n]|           (:111)
 *                         sums[b*(2N+1) + 2n+1] = sum |g(I0,F^_t0) - target| + |g(I1,F^_t1) - target|  (:152-154,166-167; 0 if !stage2_loss)
 *                         sums[b*(2N+1) + 2N]   = sum |g(I1,F01) - I0| + |g(I0,F10) - I1|  (:160-163; 0 if !stage1_loss)
 * (the stage-2 warps are the ones compute_output_image performs anyway).  The caller divides by
 * 3*H*W and applies the lambda weights (:213-233).  target: B x N x 3 x H x W.  sums: DEVICE pointer
 * to B*(2N+1) floats.  Deterministic: per-CTA partial sums in `workspace`
 * (ssm_fuse_loss_workspace_bytes) added per pair in a fixed order in fp64.
 * Backward: grad3 (may be NULL) is the dense upstream gradient of out3, grad_sums (device,
 * B*(2N+1) floats) the gradient of `sums`; writes grad_out5 (B x N x 5) and grad_flow4 (B x 4).
 * Frames and targets are data here: no image gradients (use ssm_fuse_flow_bwd + ssm_warp_bwd). */
size_t ssm_fuse_loss_workspace_bytes(int B, int N, int H, int W);
int ssm_fuse_loss_fwd(const ssm_tensor* img6, const void* packed, const ssm_tensor* flow4, const ssm_tensor* out5,
                      const ssm_tensor* target, const float* t, const ssm_tensor* out3, float* sums,
                      int B, int N, int H, int W, int dtype, int coord_mode, int stage1_loss, int stage2_loss,
                      void* workspace, size_t workspace_bytes, void* stream);
int ssm_fuse_loss_bwd(const ssm_tensor* grad3, const float* grad_sums, const ssm_tensor* img6, const void* packed,
                      const ssm_tensor* flow4, const ssm_tensor* out5, const ssm_tensor* target, const ssm_tensor* out3,
                      const float* t, const ssm_tensor* grad_out5, const ssm_tensor* grad_flow4,
                      int B, int N, int H, int W, int dtype, int coord_mode, int stage1_loss, int stage2_loss,
                      void* stream);

/* ---- the steps either side of the path --------------------------
 * ssm_frames_from_u8: F uint8 images (H_in x W_in x 3, cv2 BGR if bgr != 0, else RGB; byte strides
 * src_frame_stride / src_row_stride) -> normalised frames of H x W (H_in x W_in placed at (top, left),
 * the rest filled with pad_value3[c]) written as planar F x 3 x H x W (`planar`, may be NULL) and/or
 * as the RGBx copy of ssm_pack_frames (`rgbx`, F x H x W x 4, may be NULL) in one pass.
 * lut_device: DEVICE float[3][256], lut[c][v] = normalised value of byte v in output channel c
 * (R, G, B); the caller fills it with the reference's expression so the result is bit-identical to
 * [scripts/visualize_interpolation.py:61-88, 257-262] (pad with 0 BEFORE normalising: pad_value3[c] =
 * lut[c][0]) or [scripts/utils/dataloaders/augmentations.py:181-190 + default_reader.py:266-271] (zero
 * pad AFTER normalising: pad_value3 = 0).  pad_value3: HOST float[3].  W must be a multiple of 4.
 * A frame pair (B x 6 x H x W) is two consecutive frames of this layout.
 * ssm_frames_to_u8: the inverse [scripts/evaluate_interpolation_results.py:143-163, 192-202;
 * scripts/visualize_interpolation.py:221-232, 264-268]: crop H_out x W_out at (top, left),
 * (x * std3[c] + mean3[c]) * scale with separately rounded operations, then numpy's astype(uint8)
 * (truncate, out-of-range values wrap; saturate != 0 clamps to [0, 255] first), RGB or BGR bytes.
 * mean3 / std3: HOST float[3]. */
int ssm_frames_from_u8(const unsigned char* src, long long src_frame_stride, int src_row_stride, int bgr,
                       int F, int H_in, int W_in, int H, int W, int top, int left,
                       const float* lut_device, const float* pad_value3, const ssm_tensor* planar, void* rgbx,
                       int dtype, void* stream);
int ssm_frames_to_u8(const ssm_tensor* planar, int F, int H, int W, int top, int left, int H_out, int W_out,
                     const float* mean3, const float* std3, float scale, int bgr, int saturate,
                     unsigned char* dst, long long dst_frame_stride, int dst_row_stride, int dtype, void* stream);

/* ---- frames that arrive as 8-bit images: a2 and a3+a4 gathering from 2x2 byte entries ---------------------
 * The reference reads uint8 images and normalises them on the fly [scripts/visualize_interpolation.py:61-88
 * load_batch, :257-262 normalize_tensor].  Because that normalisation is affine in the byte, v = a*b + c with
 * a = 1/(divisor*std), c = -mean/std, it commutes with bilinear interpolation; the kernels below therefore gather
 * raw bytes from an ENTRY TABLE -- entry(x0, y0) = the 2x2 neighbourhood {(x0,y0),(x0+1,y0),(x0,y0+1),(x0+1,y0+1)}
 * x RGB as 12 bytes in a 16-byte slot, x0 in [-1, W-1], y0 in [-1, H-1], zero bytes outside the frame -- with ONE
 * 16-byte request per bilinear sample instead of four (the fp32 gathers are bound by the L1 data stage), and apply
 * a, c after interpolating: a * sum_k w_k b_k + c * sum_{k inside} w_k.  Results are within 1e-6 of ssm_flow_pack_fwd /
 * ssm_fuse_flow_fwd on the normalised fp32 frames (zeros padding and align_corners=True as in layers.warp).
 * One thread owns two adjacent pixels: W must be even, tensors aligned to two elements with even strides.
 * dtype: storage of img6 / flow4 / out16 / out3 (fp32, or bf16 with fp32 arithmetic and the estimated flows rounded to
 * bf16 before they are used, as in the fp32-frame kernels); out5 has its own out5_dtype.
 * ssm_quads_from_u8: F uint8 images (layout and placement as ssm_frames_from_u8, padding pixels = byte 0, i.e. pad
 *   BEFORE normalising as visualize_interpolation.py:76-87 does) -> F x (H+1) x (W+1) entries
 *   (ssm_quads_bytes(F, H, W) bytes, 16-byte aligned).  A frame pair is two consecutive frames: quads of pair b
 *   start at entry 2*b*(H+1)*(W+1).
 * norm6: HOST float[6] = {a_R, a_G, a_B, c_R, c_G, c_B}.
 * ssm_flow_pack_fwd_q8 [flow_interpolation.py:338-372]: img6 (B x 6 x H x W fp32, the normalised frames from
 *   ssm_frames_from_u8) is only read for the six pass-through channels; out16 B x N x 16 x H x W fp32.
 * ssm_flow_pack_fwd_q8_nhwc: the same written channels-last (B x N x H x W x 16, fp32 or bf16) as ssm_flow_pack_fwd_nhwc.
 * ssm_flow_pack_fwd_q8_lut: ssm_flow_pack_fwd_q8 without img6 -- the pass-through channels are looked up from the tables'
 *   own bytes through lut (DEVICE float[3][256], the normalised value of byte b in channel c: the table ssm_frames_from_u8
 *   applies, so the values are the ones img6 would hold when the frames were padded with byte 0 BEFORE normalising).
 *   The planar frames are not read: 5 % less DRAM traffic for the same result.
 * ssm_fuse_flow_fwd_q8 [flow_interpolation.py:374-429]: as ssm_fuse_flow_fwd(_mixed); out5 in out5_dtype (fp32 or bf16).
 * ssm_fuse_flow_fwd_q8_u8: the same with ssm_frames_to_u8 fused behind it: the fused frames are cropped,
 *   de-normalised and written as B*N uint8 images H_out x W_out x 3 (frame index b*N + n)
 *   [scripts/visualize_interpolation.py:221-232, 264-268]; no fp32 frame is materialised. */
size_t ssm_quads_bytes(int F, int H, int W);
int ssm_quads_from_u8(const unsigned char* src, long long src_frame_stride, int src_row_stride, int bgr,
                      int F, int H_in, int W_in, int H, int W, int top, int left, void* quads, void* stream);
int ssm_flow_pack_fwd_q8(const ssm_tensor* img6, const void* quads, const ssm_tensor* flow4, const float* t,
                         const ssm_tensor* out16, const float* norm6, int B, int N, int H, int W, int dtype, int coord_mode,
                         void* stream);
int ssm_flow_pack_fwd_q8_nhwc(const ssm_tensor* img6, const void* quads, const ssm_tensor* flow4, const float* t,
                              void* out16_nhwc, int out_dtype, const float* norm6, int B, int N, int H, int W,
                              int dtype, int coord_mode, void* stream);
int ssm_flow_pack_fwd_q8_lut(const void* quads, const float* lut, const ssm_tensor* flow4, const float* t,
                             const ssm_tensor* out16, const float* norm6, int B, int N, int H, int W, int dtype, int coord_mode,
                             void* stream);
int ssm_fuse_flow_fwd_q8(const void* quads, const ssm_tensor* flow4, const ssm_tensor* out5, int out5_dtype, const float* t,
                         const ssm_tensor* out3, const float* norm6, int B, int N, int H, int W, int dtype, int coord_mode,
                         void* stream);
int ssm_fuse_flow_fwd_q8_u8(const void* quads, const ssm_tensor* flow4, const ssm_tensor* out5, int out5_dtype, const float* t,
                            unsigned char* dst, long long dst_frame_stride, int dst_row_stride, int top, int left,
                            int H_out, int W_out, const float* mean3, const float* std3, float scale, int bgr, int saturate,
                            const float* norm6, int B, int N, int H, int W, int dtype, int coord_mode, void* stream);

/* ---- element-wise steps between the U-Nets' cuDNN convolutions -------
 * Channels-last activations (M x H x W x C, C a multiple of 8, 16-byte aligned), bf16 or fp32 storage, fp32
 * arithmetic in ATen's operation order.  The convolutions stay on cuDNN.  The *_bwd entry points are the
 * vector-Jacobian products autograd derives for the same torch ops (gathers: deterministic, no atomics);
 * the bias gradient (a per-channel sum of grad_x) is left to the caller.
 * ssm_upsample2x_nhwc: F.interpolate(x, size=(2H, 2W), mode="bilinear", align_corners=False), the upsampleN
 *   lambdas of [scripts/models/flow_computation.py:92-94, 103-105, 113-115, 124-126, 135-137; applied at :236-272] and
 *   [flow_interpolation.py:92-141; applied at :228-267].  `out` may be a channel slice of a wider tensor (out_pixel_stride >= C
 *   elements between pixels), which absorbs the torch.cat in front of the upsampling.
 * ssm_bias_leaky_nhwc: y <- LeakyReLU(y + bias[c], slope) in place, the bias add and activation of
 *   layers.conv [scripts/models/layers.py:21-33]; bias: DEVICE float[C]; pixels = M*H*W.
 * ssm_avgpool2_nhwc: AvgPool2d(2) [scripts/models/layers.py:60-63]: in M x 2H_out x 2W_out x C. */
int ssm_upsample2x_nhwc(const void* in, void* out, int M, int H, int W, int C, long long out_pixel_stride,
                        int dtype, void* stream);
int ssm_bias_leaky_nhwc(void* y, const float* bias, long long pixels, int C, float slope, int dtype, void* stream);
/* the same, written to out1 (pixel stride out1_pixel_stride elements; out1 == y, stride C = in place) and, if out2 is
 * not NULL, also to out2: channel slices of wider tensors, so that the torch.cat that would copy the activation
 * next [flow_computation.py:277] needs no pass of its own */
int ssm_bias_leaky_nhwc_to(const void* y, const float* bias, long long pixels, int C, float slope,
                           void* out1, long long out1_pixel_stride, void* out2, long long out2_pixel_stride,
                           int dtype, void* stream);
int ssm_avgpool2_nhwc(const void* in, void* out, int M, int H_out, int W_out, int C, int dtype, void* stream);
int ssm_upsample2x_bwd_nhwc(const void* grad_out, void* grad_in, int M, int H, int W, int C,
                            long long grad_out_pixel_stride, int dtype, void* stream);   /* grad_in: M x H x W x C */
int ssm_leaky_bwd_nhwc(const void* grad_y, const void* y, void* grad_x, long long pixels, int C, float slope,
                       int dtype, void* stream);       /* y = the activation output; grad_x may alias grad_y */
int ssm_avgpool2_bwd_nhwc(const void* grad_out, void* grad_in, int M, int H_out, int W_out, int C, int dtype, void* stream);

/* Workspace sizes (bytes) needed when the image gradient is wanted (none is needed otherwise):
 * 64-bit fixed-point accumulators of the deterministic, segmented scatter (csrc/ssm_scatter.cuh), plus an fp32 buffer
 * for the direct (non-warped) terms of ssm_flow_pack_bwd. */
size_t ssm_warp_bwd_workspace_bytes(int B, int C, int H, int W);
size_t ssm_flow_pack_bwd_workspace_bytes(int B, int N, int H, int W);
size_t ssm_fuse_bwd_workspace_bytes(int B, int N, int H, int W);

/* ---- host-buffer entry point: the whole path for one batch of frame pairs ---------------------
 * Takes HOST pointers (pinned memory recommended), copies inputs to the device in pair-sized
 * chunks on internal streams, runs a2 then a3+a4 for all N timesteps and copies the fused
 * frames back, overlapping copies with kernels.  Synchronous: returns when out3_host is
 * complete.  Dense NCHW fp32 layouts:
 *   img6_host  B x 6 x H x W, flow4_host B x 4 x H x W, out5_host B x N x 5 x H x W,
 *   t_host B*N floats, out3_host B x N x 3 x H x W, in16_host (optional, may be NULL)
 *   B x N x 16 x H x W receives the packed stage-2 input.
 * `scratch` is caller-owned DEVICE memory of at least ssm_synthesize_host_scratch_bytes(B, N, H, W)
 * bytes (three pair-sized slots), 256-byte aligned. */
size_t ssm_synthesize_host_scratch_bytes(int B, int N, int H, int W);
int ssm_synthesize_host(const float* img6_host, const float* flow4_host, const float* out5_host,
                        const float* t_host, float* out3_host, float* in16_host,
                        int B, int N, int H, int W, int coord_mode, void* scratch, size_t scratch_bytes);

/* ---- host-buffer entry point for 8-bit frames: uint8 images in, uint8 interpolated images out ----------------
 * frames_host B x 2 x H_in x W_in x 3 uint8 (dense), placed at (top, left) of the padded H x W frame (pad = byte 0);
 * flow4_host B x 4 x H x W fp32; out5_host B x N x 5 x H x W in out5_dtype (fp32 or bf16); t_host B*N floats;
 * out_host B x N x H_in x W_in x 3 uint8 (clamped to [0, 255] if saturate, else numpy astype(uint8) wrap-around);
 * lut_host float[3][256] (ssm_frames_from_u8's table: the pass-through channels of the stage-2 input are normalised
 * with it); norm6 / mean3 / std3 host float arrays as above.  Per pair it runs ssm_frames_from_u8, ssm_quads_from_u8,
 * ssm_flow_pack_fwd_q8 (the stage-2 input stays on the device, as with in16_host = NULL above) and
 * ssm_fuse_flow_fwd_q8_u8 on three slots of caller-owned device scratch (256-byte aligned).  Synchronous. */
size_t ssm_synthesize_host_u8_scratch_bytes(int B, int N, int H_in, int W_in, int H, int W, int out5_dtype);
int ssm_synthesize_host_u8(const unsigned char* frames_host, int bgr, const float* flow4_host, const void* out5_host,
                           int out5_dtype, const float* t_host, unsigned char* out_host, const float* lut_host,
                           const float* norm6, const float* mean3, const float* std3, int saturate,
                           int B, int N, int H_in, int W_in, int H, int W, int top, int left, int coord_mode,
                           void* scratch, size_t scratch_bytes);

#ifdef __cplusplus
}
#endif
#endif /* SSM_B200_H */
