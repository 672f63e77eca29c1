/* pf_b200.h -- C ABI of the B200-native PerspectiveFields inference engine (libpf_b200.so).
 *
 * The reference (jinlinyi/PerspectiveFields) is pure Python and has no FFI boundary of its own: the boundary it
 * offers is the class perspective2d.PerspectiveFields (perspective2d/perspectivefields.py:121-272).  This header is
 * what a maintainer binds (ctypes) underneath that class to replace
 *
 *     PerspectiveFields.forward                 perspectivefields.py:223-272
 *       ResizeTransform.apply_image (uint8)     perspectivefields.py:38-46     (pf_forward, images_u8 path)
 *       (x - pixel_mean) / pixel_std, stack     perspectivefields.py:234-236
 *       backbone (MiT-B3), ll_enc               mix_transformers.py:449-485, perspectivefields.py:79-83
 *       persformer_heads.inference/postprocess  persformer_heads.py:73-101, gravity_head.py:139-197,237-261,
 *                                               latitude_head.py:138-219, utils/utils.py:114-162,483-507
 *       param_net                               param_network.py:46-69,193-221, convnext.py:140-152,
 *                                               utils/utils.py:47-91
 *
 * Conventions: plain pointers and sizes only; every DEVICE pointer refers to memory on the engine's device that the
 * caller owns (allocated e.g. through PyTorch); the engine owns nothing but small lookup tables.  All work is
 * enqueued on the caller's stream and the call returns without synchronising.  Functions return 0 on success and a
 * negative pf_status otherwise; pf_last_error() gives the message (thread-local).  A handle is not re-entrant.
 * There is no CPU fallback: without a CUDA device every compute entry point fails with PF_ERR_CUDA.
 */
#ifndef PF_B200_H_
#define PF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PF_ABI_VERSION 2

typedef struct pf_engine* pf_handle;

enum pf_status { PF_OK = 0, PF_ERR_ARG = -1, PF_ERR_CUDA = -2, PF_ERR_WEIGHT = -3, PF_ERR_WORKSPACE = -4 };

enum pf_dtype { PF_F32 = 0, PF_BF16 = 1 };

enum pf_param_net { PF_PARAM_NONE = 0, PF_PARAM_CENTERED = 1 /* ParamNet @320x320 */, PF_PARAM_UNCENTERED = 2 /* ParamNetConvNextRegress @64x64 */ };

/* Model variant (the five yaml files of perspective2d/config/ reduce to these fields). */
typedef struct pf_model_desc {
  int gravity_classes;   /* 2 = regression (L2-normalised up-vector), 73 = classification logits            */
  int latitude_classes;  /* 1 = regression (clamped sin(latitude)),   180 = classification logits           */
  int param_net;         /* enum pf_param_net                                                               */
  int param_input_size;  /* MODEL.PARAM_DECODER.INPUT_SIZE for PF_PARAM_UNCENTERED (64)                     */
  float pixel_mean[3];   /* MODEL.PIXEL_MEAN, channel order of the input image (B, G, R)                    */
  float pixel_std[3];    /* MODEL.PIXEL_STD                                                                 */
} pf_model_desc;

/* One inference_batch call.  Exactly one of images_u8 / images_chw is non-NULL. */
typedef struct pf_batch {
  int n;                       /* number of images                                                          */
  /* uint8 path (PerspectiveFields.inference{,_batch}): DEVICE blob of tightly packed HWC uint8 BGR images   */
  const uint8_t* images_u8;
  const int64_t* image_offset; /* HOST [n] byte offset of each image in the blob                            */
  /* float path (PerspectiveFields.forward called directly): DEVICE fp32 [n,3,320,320], already resized      */
  const float* images_chw;
  const int32_t* height;       /* HOST [n] original heights ("height" key)                                  */
  const int32_t* width;        /* HOST [n] original widths  ("width" key)                                   */
  /* outputs, DEVICE, fp32 */
  float* pred_gravity;         /* [n, gravity_classes, 320, 320]                                            */
  float* pred_latitude;        /* [n, latitude_classes, 320, 320]                                           */
  float* gravity_original;     /* blob; image i occupies [2, H_i, W_i] at gravity_original_offset[i]        */
  const int64_t* gravity_original_offset;   /* HOST [n], in floats                                         */
  float* latitude_original;    /* blob; image i occupies [H_i, W_i] at latitude_original_offset[i]          */
  const int64_t* latitude_original_offset;  /* HOST [n], in floats                                         */
  float* params;               /* [n, 8]: roll, pitch, vfov|general_vfov (deg), rel_cx, rel_cy, rel_focal, raw x2, 0;
                                  may be NULL when param_net == PF_PARAM_NONE                               */
} pf_batch;

int pf_abi_version(void);
const char* pf_last_error(void);

/* Number of CUDA kernels launched by this library in the calling process so far (all handles). */
int64_t pf_kernel_launch_count(void);

/* Engine lifetime.  `device` is the CUDA ordinal; the engine makes it current for its own calls. */
int pf_create(int device, const pf_model_desc* desc, pf_handle* out);
int pf_destroy(pf_handle h);

/* Register one repacked weight tensor (DEVICE pointer, stays owned by the caller and must outlive the handle).
 * Names and layouts are listed in perspectivefields_b200/weights.py; pf_finalize checks that all are present. */
int pf_set_weight(pf_handle h, const char* name, const void* dev_ptr, int64_t numel, int dtype);
int pf_finalize(pf_handle h);

/* Bytes of DEVICE scratch pf_forward needs for a batch of n images whose largest member has max_h rows. */
int64_t pf_workspace_bytes(pf_handle h, int n, int max_h);

/* Whole forward of perspectivefields.py:223-272 for one batch, enqueued on `stream` (a cudaStream_t). */
int pf_forward(pf_handle h, const pf_batch* batch, void* workspace, int64_t workspace_bytes, void* stream);

/* Per-launch timing of the GEMM engine with CUDA events on the launch stream (bench.py roofline leg).  pf_profile_read
 * fills out21[cfg*3 + {0,1,2}] = {milliseconds, algorithmic FLOPs (2*M*N*K), launches} per engine configuration (slots 0-4 are
 * unused since ABI 2 -- the earlier HMMA / register-staged engines were removed; 5: TMA+tcgen05 GEMM mode, 6: TMA+tcgen05 halo
 * 3x3 mode) accumulated since the previous read; synchronise the stream first. */
int pf_profile_enable(pf_handle h, int on /* 0 off, 1 on, n > 1: on + pre-create events for n GEMM launches */);
int pf_profile_read(pf_handle h, double* out21);
/* A CUDA-event pair around EVERY kernel launch of the forward graph: in-pipeline time per kernel (bench.py "per_kernel").
 * enable(max_launches > 0) pre-creates the events and starts recording, enable(0) stops.  read() writes a text table
 * "kernel,launches,ms\n..." (aggregated since the last read) into buf and returns its length; synchronise the stream first. */
int pf_profile_kernels_enable(pf_handle h, int max_launches);
int pf_profile_kernels_read(pf_handle h, char* buf, int cap);

/* Engine options (all default 1 unless noted; the whole graph runs on the persistent TMA -> tcgen05 -> TMEM engine with pre-split
 * bf16 hi/lo activations, gemm_tma.cuh):
 * "attn_tc": attention core on tcgen05 / TMEM (attention_tc.cuh: S = Q K^T into TMEM, softmax one thread per row from TMEM, P V as a
 *   second MMA with V consumed MN-major); 0 = the warp-level mma.sync kernel (attention_mma.cuh).
 * "attn_mma": (with "attn_tc" = 0) mma.sync attention core instead of the exact-softmax CUDA-core kernel.
 * "attn_split": q / kv leave their GEMMs as split planes (0 = fp32, split inside the attention kernel).
 * "stem_tc": the two 7x7 stems as patch gather + TMA GEMM instead of fp32 direct convolution.
 * "phase_conv1": conv_fuse_conv1 composed with the x2 bilinear upsample in front of it (four output phases on the 160x160 grid
 *   + an exact fp32 border-ring kernel); 0 = materialise the upsampled tensor, conv at 320x320.
 * "pair": GEMM-mode launches with at least one 256 x BN tile per TPC run on CTA pairs (gemm2_tma.cuh, tcgen05.mma.cta_group::2);
 *   0 = every launch on the single-CTA kernel.
 * "pdl": programmatic dependent launch of the graph's kernels (a kernel's prologue overlaps its predecessor's tail).
 * "fork" (default 0): the spatial-reduction branch of a MiT block on a second stream beside the q projection (measured 1 % slower).
 * "dw_ln" (default 0): ConvNeXt depthwise 7x7 fused with the LayerNorm behind it (measured slower: profiles/r02_notes.md).
 * "decode_only" (default 0; classification heads): the 73 / 180 logits are never written -- the 1x1 prediction
 *   conv, argmax and bin decode (gravity_head.py:243-244 + utils/utils.py:114-130, latitude_head.py:205-208 + utils.py:148-162)
 *   run in one kernel and pred_gravity / pred_latitude receive the decoded fields [n,2,320,320] / [n,1,320,320] (degrees). */
int pf_set_option(pf_handle h, const char* name, int value);

/* Debug taps (tests only): when enabled, intermediates of the next pf_forward are kept (never recycled) and can be
 * copied out by name (device-to-device, enqueued on `stream`).  Names are listed by pf_debug_name(i). */
int pf_debug_enable(pf_handle h, int on);
int pf_debug_count(pf_handle h);
const char* pf_debug_name(pf_handle h, int i);
int64_t pf_debug_numel(pf_handle h, const char* name);
int pf_debug_copy(pf_handle h, const char* name, float* dst_dev, int64_t numel, void* stream);

/* ---- camera parameters -> dense perspective fields -------------------------------------
 * Replaces PanoCam.get_up_general / PanoCam.get_lat_general (perspective2d/utils/panocam.py:451-513, :515-556), which callers
 * evaluate right after the inference path on ParamNet's output (utils/utils.py:367-385, demo/demo.py:69-78).  One launch per
 * 24 images; float64 arithmetic, float32 results.  No engine handle: the function has no weights. */
typedef struct pf_camera {
  int32_t height, width;        /* im_h, im_w */
  double focal_rel;             /* focal length / image height */
  double elevation, roll;       /* radians */
  double cx_rel, cy_rel;        /* principal point: pixel / size - 0.5 */
  int64_t up_offset;            /* float offset of this image's [H,W,2] (x,y) block in `up` */
  int64_t lat_offset;           /* float offset of this image's [H,W] block (degrees) in `lat` */
} pf_camera;
/* cams: HOST array of n descriptors; up / lat: DEVICE blobs (either may be NULL to skip that field). */
int pf_camera_fields(int device, const pf_camera* cams, int n, float* up, float* lat, void* stream);

/* ---- multi-GPU gather of results (one process per GPU; NCCL point-to-point over NVLink) ---------------------
 * inference_batch shards its list over the ranks; the per-image results live on each rank's device and are gathered to ONE
 * rank with grouped ncclSend / ncclRecv enqueued on the caller's stream (so that the gather of micro-batch k overlaps the
 * forward of micro-batch k+1 when issued on a side stream).  NCCL is resolved at run time from the process
 * (libnccl.so.2 -- the one PyTorch ships); the 128-byte unique id is created on rank 0 and distributed by the caller
 * (torch.distributed / MPI / a file: plumbing). */
typedef struct pf_comm* pf_comm_handle;
int pf_comm_unique_id(void* id128);                                              /* rank 0: ncclGetUniqueId -> 128 bytes */
int pf_comm_create(int device, int rank, int nranks, const void* id128, pf_comm_handle* out);
int pf_comm_destroy(pf_comm_handle c);
/* One grouped exchange.  On every rank != root: send `count` segments (DEVICE pointer, bytes) to root.  On root: receive
 * `count` segments, segment i from rank peer[i].  All segments of one call travel in one ncclGroup on `stream`. */
int pf_gather(pf_comm_handle c, int root, int count, void* const* dev_ptrs, const int64_t* bytes, const int32_t* peer, void* stream);

/* ---- decode front-end: JPEG bytes -> the device blob of BGR uint8 HWC images pf_forward reads ---------
 * Replaces `cv2.imread` in front of the path (demo/demo.py:151) with nvJPEG (bound at run time), the images of a batch fanned out
 * over worker threads / streams and joined into `stream`.  pf_jpeg_info parses the header only (sizes for the blob layout). */
typedef struct pf_jpeg* pf_jpeg_handle;
int pf_jpeg_create(int device, int max_threads /* 0 = half the host's hardware threads, at most 32 */, pf_jpeg_handle* out);
int pf_jpeg_destroy(pf_jpeg_handle j);
int pf_jpeg_info(pf_jpeg_handle j, const uint8_t* data, int64_t length, int32_t* height, int32_t* width);
/* data[i] / length[i]: HOST JPEG streams; blob: DEVICE; image i is written as [height[i], width[i], 3] BGR at byte offset[i]. */
int pf_jpeg_decode_batch(pf_jpeg_handle j, int n, const uint8_t* const* data, const int64_t* length, const int32_t* height,
                         const int32_t* width, uint8_t* blob, const int64_t* offset, void* stream);

/* ---- single-operator entry points (unit tests; the same kernels pf_forward launches) -------------------- */

/* Conv / linear on NHWC fp32 on the TMA -> tcgen05 engine, bf16x3 split precision (the input is split into hi/lo planes first,
 * as a producer kernel of the forward graph would; 3x3/s1/p1 with Cin % 64 == 0 -> halo mode, 1x1 -> GEMM mode, else patch gather).  x: [B,H,W,Cin]; whi/wlo: bf16 [N][KH*KW*Cin]
 * ordered (ky,kx,ci); bias: [N] or NULL; res: [B,OH,OW,N] or NULL; y: [B,OH,OW,N].
 * y = act(conv(relu_in?(x)) + bias) (+ relu_res?(res));  act: 0 none, 1 ReLU, 2 GELU. */
int pf_op_conv_gemm(const float* x, int B, int H, int W, int Cin, const void* whi, const void* wlo, const float* bias,
                    int N, int KH, int KW, int stride, int pad, int in_relu, int act, const float* res, int res_relu,
                    float* y, void* stream);
int pf_op_layernorm(const float* x, float* y, int64_t rows, int C, const float* w, const float* b, float eps, void* stream);
int pf_op_attention(const float* q, const float* kv, float* out, int B, int N, int C, int heads, void* stream);      /* CUDA-core fp32 */
int pf_op_attention_mma(const float* q, const float* kv, float* out, int B, int N, int C, int heads, void* stream);  /* warp-level mma.sync, bf16x3 */
int pf_op_attention_tc(const float* q, const float* kv, float* out, int B, int N, int C, int heads, void* stream);   /* tcgen05 / TMEM, bf16x3 (default) */
int pf_op_dwconv3x3_gelu(const float* x, float* y, int B, int H, int W, int C, const float* w9c, const float* bias, void* stream);
int pf_op_dwconv7x7(const float* x, float* y, int B, int H, int W, int C, const float* w49c, const float* bias, void* stream);
int pf_op_upsample2x(const float* x, float* y, int B, int H, int W, int C, void* stream);
/* Pillow-exact resize + normalise of ONE uint8 HWC image -> [320,320,4] fp32 (b,g,r,0). */
int pf_op_preprocess(const uint8_t* img_dev, int H, int W, const float* mean3, const float* std3, float* y, void* stream);
/* ResizeTransform.apply_image (perspectivefields.py:34-67) on the device, HWC in -> HWC out, C channels:
 *   uint8 (PIL branch, :38-46): Pillow-exact antialiased bilinear (the same integer kernel as pf_forward's pre-process);
 *   float32 (:47-66): F.interpolate(mode="bilinear", align_corners=False), no antialias. */
int pf_op_resize_u8(const uint8_t* img_dev, int H, int W, int new_h, int new_w, uint8_t* out_dev, void* stream);   /* C = 3 */
int pf_op_resize_f32(const float* img_dev, int H, int W, int C, int new_h, int new_w, float* out_dev, void* stream);
/* write-only bandwidth probe: fills dst[numel] (numel % 4 == 0, 16-byte aligned) with 16-byte streaming stores (bench.py measures
 * the store roofline of the write-out kernels with it). */
int pf_op_fill_stream(float* dst, int64_t numel, float value, void* stream);
/* argmax over channels + bin decode of NCHW logits [B,NC,HW] (gravity_head.py:243-244 + utils/utils.py:114-130 when
 * is_gravity, field [B,2,HW]; latitude_head.py:205-208 + utils/utils.py:148-162 otherwise, field [B,1,HW] in degrees). */
int pf_op_argmax_decode(const float* logits, float* field, int B, int HW, int NC, int is_gravity, void* stream);
/* the same decode WITHOUT materialised logits (option "decode_only"): 1x1 conv 32 -> NC on feat [B*HW, ld] (channels coff..coff+31,
 * weights [NC][32], bias [NC]) + argmax + bin decode in one kernel; logits bit-identical to the separate 1x1 conv kernel. */
int pf_op_pred_argmax_decode(const float* feat, int ld, int coff, const float* w, const float* bias, float* field, int B, int HW, int NC,
                             int is_gravity, void* stream);
/* resample of decoded fields to the original sizes (gravity_head.py:246-256, latitude_head.py:209-219, utils/utils.py:483-507):
 * vec [n,2,320,320], lat [n,1,320,320] -> blobs as in pf_batch; lat_is_sin: lat holds sin(latitude) (regression head). */
int pf_op_postprocess(const float* vec, const float* lat, int n, const int32_t* height, const int32_t* width,
                      float* gravity_original, const int64_t* gravity_original_offset, float* latitude_original,
                      const int64_t* latitude_original_offset, int lat_is_sin, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PF_B200_H_ */
