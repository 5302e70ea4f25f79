/* jpb200.h — C ABI of libjpb200.so, the sm_100a CUDA library behind jperceiver_b200.
 *
 * Every entry point takes raw DEVICE pointers, plain sizes and a cudaStream_t passed as void*; it
 * launches asynchronously on that stream, allocates nothing, never synchronises the device and
 * returns 0 on success (1 = bad argument, 2 = unsupported shape, 1000+e = cudaError_t e at launch).
 * Outputs and workspaces are caller-allocated.  Unless stated otherwise tensors are fp32; images are
 * NCHW (the reference's boundary layout), network activations are NHWC ("channels-last").
 *
 * The reference (sunnyHelen/JPerceiver, /root/reference) is pure PyTorch and has no FFI of its own:
 * each function below names the reference Python code (file:line under mono/model/mono_baseline/
 * unless a longer path is given) whose arithmetic it replaces.  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 */
#ifndef JPB200_H
#define JPB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JPB_MAX_SRC 4 /* source frames per snippet (frame_ids[1:]); the reference uses 1 or 2 */

/* ---- library info ------------------------------------------------------------------------- */
int jpb_abi_version(void);       /* bumped on any signature change */
const char* jpb_build_info(void); /* "sm_100a nvcc <ver> <date>" */

/* ---- fused photometric reprojection loss, one launch per scale ------------------------------
 * forward : net.py:690-702 (generate_images_pred), layers.py:41-107 (Backproject, Project, SSIM),
 *           net.py:84-92 (robust_l1, compute_reprojection_loss), net.py:159-175 (automask, min, mean)
 * backward: the autograd graph of the same lines w.r.t. disp_s and cam_T_cam.                   */
typedef struct JpbPhotoArgs {
  const float* target;             /* [B,3,H,W] inputs[("color",0,0)]                              */
  const float* src[JPB_MAX_SRC];   /* [B,3,H,W] inputs[("color",f,0)], f = frame_ids[1:]           */
  const float* T[JPB_MAX_SRC];     /* [B,4,4]   outputs[("cam_T_cam",0,f)], row-major              */
  const float* noise[JPB_MAX_SRC]; /* [B,H,W] explicit automask noise, or NULL                     */
  const float* disp;               /* [B,hs,ws] outputs[("disp",0,s)]                              */
  const float* K;                  /* [B,4,4] inputs[("K",0)]                                      */
  const float* invK;               /* [B,4,4] inputs[("inv_K",0)]                                  */
  int B, H, W, hs, ws, F;
  int automask;                    /* opt.automask: prepend the F identity candidates              */
  float min_disp, max_disp;        /* 1/max_depth, 1/min_depth (layers.py:33-38)                   */
  float noise_scale;               /* 1e-5 in the reference (net.py:163); used when noise[f]==NULL */
  uint64_t seed, stream;           /* Philox key / stream id for the in-kernel N(0,1) draw         */
  const long long* step;           /* optional device step counter: stream += 64 * step[0] (fresh noise per
                                      step even when the launch is replayed from a CUDA graph)               */
  double* loss_sum;                /* [1] += sum over b,y,x of min over candidates                 */
  long long* min_index;            /* [B,H,W] int64 argmin (outputs[("min_index",s)]) or NULL      */
  unsigned char* winner;           /* [B,H,W] same argmin as a byte (kept for backward) or NULL    */
  float* warped[JPB_MAX_SRC];      /* [B,3,H,W] outputs[("color",f,s)] or NULL.  Forward: written.  Backward: when non-NULL, the
                                      frames the forward of this scale wrote are staged instead of being re-projected          */
  float* ident_err;                /* [B,H,W,2] identity-candidate errors (frame 0, frame 1) WITHOUT the tie-breaking noise, or NULL.
                                      They compare the target with the un-warped sources (net.py:159-166) and do not depend on the
                                      scale: one launch of a step computes them, the others read them                          */
  int ident_mode;                  /* 0: compute, do not store (default); 1: compute and store to ident_err; 2: read ident_err
                                      (forward schedule 3 with automask only; the backward ignores both fields)                */
} JpbPhotoArgs;

typedef struct JpbPhotoGrad {
  const float* grad_out;           /* [1] d(total)/d(loss_dict[("min_reconstruct_loss",s)])        */
  float inv_count;                 /* 1 / (num_scales * B*H*W)                                     */
  const unsigned char* winner;     /* [B,H,W] from the forward launch                              */
  float* grad_disp;                /* [B,hs,ws] += ; caller zero-fills                             */
  float* grad_T[JPB_MAX_SRC];      /* [B,4,4] += ; caller zero-fills; NULL to skip                 */
} JpbPhotoGrad;

int jpb_photometric_fwd(const JpbPhotoArgs* args, void* stream);
/* Forward schedule: 2 = one value per instruction (default, the measured one); 3 = the two source frames of a snippet packed
 * into FADD2/FMUL2/FFMA2 pairs (same results within fp32 rounding).  Process-wide; not thread-safe against running launches. */
int jpb_photometric_set_variant(int fwd_variant);
int jpb_photometric_get_variant(void);   /* the forward schedule in force (ident_mode needs schedule 3) */
int jpb_photometric_bwd(const JpbPhotoArgs* args, const JpbPhotoGrad* grad, void* stream);
/* Backward schedule: 4 = register-resident kernel specialised on F <= 2 (default; F > 2 always takes schedule 1); 1 = the
 * generic kernel.  Same gradients within fp32 summation order.  Process-wide; not thread-safe against running launches.       */
int jpb_photometric_set_bwd_variant(int bwd_variant);

/* ---- area-downsampled target pyramid -----------------------------------------------------------
 * level[s] = F.interpolate(img, (H/2^(s+1), W/2^(s+1)), mode="area")  (net.py:762), all levels in one
 * pass over the frame; H and W must be multiples of 2^nlev.                                        */
typedef struct JpbPyramid {
  int nlev;          /* 1..4 */
  float* level[4];   /* [BC, H/2^(s+1), W/2^(s+1)] */
} JpbPyramid;
int jpb_area_pyramid(const float* img, int BC, int H, int W, const JpbPyramid* out, void* stream);

/* ---- edge-aware smoothness (net.py:182-190, 758-786), one scale -------------------------------
 * acc: [B,7] doubles, zero-filled by the caller (six stencil sums + sum of disp per sample), kept for
 * backward.  out[0] = weight * smooth  with weight = smoothness_weight / 2^s / num_scales.        */
int jpb_smooth_fwd(const float* disp, const float* J, int B, int h, int w, int disp_norm, float weight,
                   double* acc, float* out, void* stream);
int jpb_smooth_bwd(const float* disp, const float* J, int B, int h, int w, int disp_norm, float weight,
                   const double* acc, const float* grad_out, float* grad_disp /* += */, void* stream);

/* ---- CGT scale label (net.py:212-310 static, 311-402 dynamic, 403-476 both; layers.py:214-252) ----
 * out[b] = warp(z_map) * warp(label)              (mode 0, "Argo_both")
 *        = warp(z_map) * [warp(label)==1] * quad  (mode 1, "static"/"static_raw"/"Argo_static"; quad = the cv2-filled
 *          rectangle projection of net.py:292-306, rasterised by the host from sample 0's calibration)
 *        = warp(z_map) * quad                     (mode 2, "dynamic"/"Argo_dynamic": net.py:390-401 masks the z-map
 *          by the quad alone; label may be NULL; z_offset is 0 for KITTI, net.py:325-326) */
typedef struct JpbScaleLabelArgs {
  const float* label;       /* [B,occ,occ] inputs[("both_dynamic"|"bothS",0,0)] in {0,1}; unused in mode 2 */
  const float* K3;          /* inputs[("odometry_K",0,0)]: element (i,j) of sample b at K3[b*k_stride+i*k_row+j] */
  int k_stride, k_row;
  const float* Tr;          /* [B,4,4] inputs[("Tr_cam2_velo",0,0)]                                */
  const unsigned char* quad;/* [Hf,Wf] (modes 1, 2) or NULL (mode 0)                                   */
  float* out;               /* [B,Hf,Wf]                                                           */
  int B, occ, Hf, Wf;
  int mode;                 /* 0 both, 1 static, 2 dynamic                                         */
  int align_corners;        /* torchgeometry 0.1.2 leaves it to torch's default: un-pinned, see DESIGN.md */
  float z_offset;           /* 0.27 KITTI (0 in mode 2), 1.9 Argoverse (net.py:229-233, 323-326)   */
  float cam_height;         /* 1.73 KITTI, 0.33 Argoverse (net.py:257-260)                         */
} JpbScaleLabelArgs;
int jpb_scale_label(const JpbScaleLabelArgs* args, void* stream);

/* ---- CGT scale loss (net.py:193-211), one scale ----------------------------------------------
 * acc[0] += sum |g-p|/g, acc[1] += count over label>0 (and the garg/eigen crop for static_raw).    */
typedef struct JpbScaleLossArgs {
  const float* disp;        /* [B,hs,ws]                    */
  const float* label;       /* [B,Hf,Wf] from jpb_scale_label */
  int B, hs, ws, Hf, Wf;
  int crop;                 /* static_raw: rows 153..370, cols 44..1196 */
  float min_disp, max_disp;
  double* acc;              /* [2] */
  float weight;             /* scale_weight / 2^s / num_scales (backward) */
  const float* grad_out;    /* [1] (backward) */
  float* grad_disp;         /* [B,hs,ws] += (backward) */
} JpbScaleLossArgs;
int jpb_scale_loss_fwd(const JpbScaleLossArgs* args, void* stream);
int jpb_scale_loss_bwd(const JpbScaleLossArgs* args, void* stream);

/* ---- signed distance map of binary BEV labels (boundary_loss.py:121-147) ------------------------
 * sdf = EDT(background->foreground) - EDT(foreground->background), 0 on the inner 4-connected boundary,
 * all-zero for an empty mask.  Exact (integer squared distances).  work: 2*B*n*n ints.              */
int jpb_signed_distance(const float* label, int B, int n, int* work, float* sdf, void* stream);

/* ---- BEV head loss (net.py:554-617): loss_weight*region + [loss_sum>=2] loss2_weight*BD + [loss_sum==3] CE(weight=[1,w_fg])
 * region: 0 soft IoU (dice_loss.py:293-331), 1 soft Dice (:255-291), 2 Tversky alpha=.3 beta=.7 (:333-372),
 *         3 focal alpha=.25 gamma=2 smooth=1e-5 (focal_loss.py:7-92)                                             */
typedef struct JpbBevArgs {
  const float* logits;      /* element (b,c,y,x) at logits[b*stride_b + (y*occ+x)*stride_p + c*stride_c] */
  long long stride_b, stride_c, stride_p;
  const float* label;       /* [B,occ,occ] in {0,1} */
  const float* sdf;         /* [B,occ,occ] */
  int B, occ;
  float w_fg, loss_weight, loss2_weight;
  double* acc;              /* [4*B + 4], zero-filled by the caller, kept for backward */
  int region;               /* 0 iou, 1 dice, 2 tversky, 3 focal */
  int loss_sum;             /* 1: region only, 2: + boundary, 3: + boundary + cross entropy (opt.loss_sum) */
} JpbBevArgs;
int jpb_bev_loss_fwd(const JpbBevArgs* args, float* out, void* stream);
int jpb_bev_loss_bwd(const JpbBevArgs* args, const float* grad_out, float* grad_logits, void* stream);

/* ---- mean |x - y| (nn.L1Loss, net.py:619-622) -------------------------------------------------- */
int jpb_l1_mean_fwd(const float* x, const float* y, long long n, double* acc /* [1] += */, void* stream);
int jpb_l1_mean_bwd(const float* x, const float* y, long long n, const float* grad_out, float* gx, float* gy, void* stream);

/* ---- implicit-GEMM convolution on tcgen05 tensor cores (TF32 in, FP32 accumulate in TMEM) ------------
 * Replaces nn.Conv2d plus the passes the reference runs in front of / behind it as separate kernels:
 * ReflectionPad2d/ZeroPad2d (layers.py:156-167), nearest 2x up-sampling (layers.py:110), torch.cat of up to
 * three sources (depth_decoder.py:76,96,115), bias, residual add (layers.py:197) and the activation
 * (F.leaky_relu depth_decoder.py:60, ReLU, sigmoid depth_decoder.py:35-38).
 * Activations are NHWC fp32.  K is walked in 16-byte chunks described by `table` (4 ints per chunk:
 * {source | (tap*nsrc + source) << 8, or -1; dy<<16 | (dx & 0xffff); channel offset; valid bytes}); 8 chunks = one
 * 32-float K block.
 * `weight` is the K-major matrix [N][w_row] whose first w_cols columns follow the same chunk order.      */
#define JPB_CONV_MAX_SRC 3
typedef struct JpbConvArgs {
  const float* src[JPB_CONV_MAX_SRC];   /* [B, H_i, W_i, C_i] */
  int src_C[JPB_CONV_MAX_SRC], src_H[JPB_CONV_MAX_SRC], src_W[JPB_CONV_MAX_SRC];
  int src_up[JPB_CONV_MAX_SRC];         /* 1: source is read through a nearest 2x up-sampling */
  int nsrc;
  int B, Hin, Win;                      /* logical input extent (after up-sampling) */
  int Ho, Wo, N;                        /* output extent and channels */
  int stride, pad, reflect;
  const float* weight;                  /* [N][w_row], 16-byte aligned, w_row % 4 == 0 */
  long long w_row;
  int w_cols;
  const int* table;                     /* [nkb*8][4] */
  int nkb;                              /* number of 32-float K blocks */
  const float* bias;                    /* [N] or NULL */
  const float* residual;                /* [B,Ho,Wo,N] or NULL, added before the activation */
  int act;                              /* 0 none, 1 ReLU, 2 LeakyReLU(0.01), 3 sigmoid */
  float* out;                           /* [B,Ho,Wo,N] (scatter == 0) */
  /* -- data-gradient use of the same kernel: sources = dY, weight = flipped/transposed W, N = Cin of the forward -- */
  int in_div;                           /* 2: gather coordinate must be even and is halved (dgrad of a stride-2 conv) */
  int nt;                               /* N tile override (0 = auto); scatter tiles must not straddle a destination */
  int scatter;                          /* 1: fold the result back into the forward convolution's sources */
  float* dst[JPB_CONV_MAX_SRC];         /* [B, H_j, W_j, C_j] gradient of forward source j (+=, zero-filled by the caller) */
  int dst_C[JPB_CONV_MAX_SRC], dst_H[JPB_CONV_MAX_SRC], dst_W[JPB_CONV_MAX_SRC], dst_up[JPB_CONV_MAX_SRC];
  int ndst;
  int fold_pad, fold_reflect, fold_H, fold_W; /* output pixel (py,px) -> (py-fold_pad, px-fold_pad), reflected into fold_H x fold_W */
  int ntaps, kw;                        /* filter taps (kh*kw) and filter width: tap t reads (dy, dx) = (t / kw, t % kw) */
  int ksplit;                           /* > 1: split the K blocks over this many CTAs per tile (no bias/residual/act; output
                                           zero-filled by the caller, partial tiles are added atomically) */
  const int* kcol;                      /* [nkb] weight column (float index) of K block i, or NULL for i*32: lets the host order
                                           the K blocks so that taps sharing input pixels are consecutive */
  int l1_gather;                        /* 1: gather through L1 (cp.async.ca) — pays off with the K-block order above */
  long long* dbg;                       /* debug only: [512 CTAs][6 warps][8] globaltimer stamps (tools/conv_timeline.py), or NULL */
  int dbg_skip;                         /* timing experiments only (results are wrong): bit 0 = no A gather, bit 1 = no weight TMA */
  double* stats;                        /* optional BatchNorm statistics fused into the epilogue: stats[c] += sum over rows of
                                           out[.][c], stats[N + c] += sum of squares (jpb_bn_stats_accumulator; no split-K, N % 4 == 0) */
  float acc_scale;                      /* 0 = 1.  tcgen05 kind::tf32 TRUNCATES both operands to 10 mantissa bits (cuDNN rounds to nearest):
                                           every product is short by 2 * 2^-11 * ln 2 = 6.8e-4 on average, a coherent bias that compounds
                                           through the BatchNorm-free decoder.  The raw accumulator is multiplied by this factor
                                           (JPB_TF32_TRUNC_COMP for plain fp32 operands, 1 for pre-split 3xTF32 operands) before the epilogue. */
  int patch;                            /* != 0 (1 = one or two tiles per CTA chosen by the library, 2 / 3 = forced to one / two): 3x3 / stride 1 / zero pad 1 over ONE dense source with C % 32 == 0 and W % 8 == 0: the A operand is
                                           read as 18 x 16 pixel TMA patches (one per 32-channel block and 16 x 8 output tile) shared by all nine
                                           taps instead of the per-tap gather; weight columns must be tap-major (tap * C + c); table / kcol unused */
  int patch_ntaps;                      /* patch mode: number of taps (9 for 3x3; 8 for the space-to-depth form of the 7x7 / stride-2 stems) */
  int patch_halo;                       /* patch rows beyond the 16 tile rows (2 for 3x3, 3 for the stems) */
  int patch_org_y, patch_org_x;         /* the patch starts at (tile row - org_y, tile pixel - org_x) (1, 1 for 3x3 / pad 1) */
  int patch_tapoff[16];                 /* per tap: (dy * 2048 + dx * 128) / 16 — start-address offset of the tap inside the patch */
  int patch_desc_mode;                  /* debug: 2 = set the UMMA matrix-descriptor base offset for the shifted patch start addresses (WRONG on
                                           B200: the swizzle follows absolute address bits; kept to reproduce the measurement) */
  int rows;                             /* != 0: the A operand arrives by TMA boxes of 32 pixels x 32 channels instead of the gather: stride 1,
                                           every source dense (no up-sampling) with C % 32 == 0, N tile >= 64, reflect only for pad 1 with
                                           Wo == Win.  2 = one deep CTA per SM for the 256-wide tiles                                         */
  int rows_wv;                          /* rows != 0: row length of the tile raster, a multiple of 32 and >= Wo; pixels Wo .. rows_wv - 1 of a
                                           row are computed and dropped (data gradient of reflection-padded layers: Wo = W + 2)              */
  int dst_mul, dst_oy, dst_ox;          /* scatter: dst_mul > 1 writes output pixel (ty, tx) to (ty * dst_mul + dst_oy, tx * dst_mul + dst_ox) —
                                           one parity class of the data gradient of a stride-2 convolution per launch (4x fewer taps in
                                           all than the zero-stuffed form in_div = 2)                                                        */
} JpbConvArgs;
#define JPB_TF32_TRUNC_COMP 1.00067702f  /* 1 + 2 * 2^-11 * ln 2 */
int jpb_conv2d_fwd(const JpbConvArgs* args, void* stream);

/* weight gradient of the same operator: dw[n][k] (+)= sum_p im2col[p][k] * dy[p][n], k in the chunk order of `table`.
 * dy is the gradient w.r.t. the pre-activation output, [B*Ho*Wo][N] with N % 4 == 0.  The pixel range is split over
 * `splits` CTAs per tile; with splits > 1 results are accumulated atomically into a zero-filled dw.            */
typedef struct JpbConvWgradArgs {
  const float* src[JPB_CONV_MAX_SRC];
  int src_C[JPB_CONV_MAX_SRC], src_H[JPB_CONV_MAX_SRC], src_W[JPB_CONV_MAX_SRC], src_up[JPB_CONV_MAX_SRC];
  int nsrc;
  int B, Hin, Win, Ho, Wo, N;
  int stride, pad, reflect;
  const int* table;
  int nchunks;               /* rows of the table (multiple of 8) */
  const float* dy;
  float* dw;                 /* [N][w_row] */
  long long w_row;
  int w_cols;
  int splits;
  float* dbg;                /* debug only: first pipeline stage (A then B tile) is copied here when non-NULL */
  int accumulate;            /* 1: always add into dw (dw aliases the parameter's slot of the flat gradient buffer) */
  float acc_scale;           /* as JpbConvArgs.acc_scale */
  int dy_pitch;              /* floats between consecutive pixels of dy; 0 = N (dense).  The 3xTF32 mode passes the hi or lo
                                half of a split gradient tensor (jpb_tf32_split): pitch = 2 * N */
  int rows;                  /* != 0: the im2col^T operand arrives by TMA (stride 1, Ho x Wo == Hin x Win, Wo % 32 == 0, no up-sampled
                                source): every group of 8 table rows with gflags != 0 is ONE (source, tap, 32-channel block) and is
                                read as a {32 channels, 32 pixels} box of the source; other groups are gathered.  2 = one deep CTA
                                per SM for the 256-wide tiles (default: two shallow ones)                                        */
  const unsigned char* gflags; /* rows != 0: [nchunks / 8] 1 = regular group                                                      */
  int dz3;                   /* set by the library (ignored on input): dZ is fetched as one 3-D box per step              */
  const int* chunk_col;      /* rows != 0: [nchunks] column of dw that the first K position of each table row accumulates into,
                                -1 for padding rows (the table order is then free)                                              */
} JpbConvWgradArgs;
int jpb_conv2d_wgrad(const JpbConvWgradArgs* args, void* stream);
/* Schedule of the 256-wide forward / data-gradient tiles with more than 148 tiles: 0 = one tile per CTA (default); 1 = CTA pairs
 * (tcgen05 cta_group::2, two pairs per TPC, each CTA streams half of the weight tile); 2 = one pair per TPC, four stages.  Same
 * results; measured slower on B200 (profiles/r2_conv_pair_ab.txt), kept as a tested schedule.  Process-wide (JPB_CONV_PAIR sets the
 * initial value).                                                                                                                 */
int jpb_conv_set_pair(int mode);

/* ---- 3x3 stride-1 pad-1 convolutions with 1..2 output channels on the CUDA cores (disparity heads
 * depth_decoder.py:35-38, BEV topview heads layout_model.py:158), as nine 1x1 projections + a shift-and-add gather.
 * x: [B,Hs,Ws,C] NHWC (read through a nearest 2x up-sampling when up != 0), w / dw: [N][3][3][C], y / dz: [B,Ho,Wo,N].
 * work: [B*Hs*Ws][N*9] floats (forward: scratch; backward: ZERO-FILLED by the caller).  dw is accumulated (+=), dx is
 * overwritten; either may be NULL.                                                                          */
int jpb_conv3x3_smalln_fwd(const float* x, const float* w, const float* bias, float* y, float* work, int B, int Hs, int Ws, int C,
                           int up, int N, int reflect, int act, void* stream);
int jpb_conv3x3_smalln_bwd(const float* x, const float* w, const float* dz, float* work, float* dw, float* dx, int B, int Hs, int Ws,
                           int C, int up, int N, int reflect, void* stream);

/* ---- 3xTF32 operand split (precision mode of the tensor-core convolutions that reproduces the reference's fp32 arithmetic,
 * torch.backends.cudnn.allow_tf32 = False): x [rows][C] -> out [rows][2*Cp], Cp = C rounded up to 4; columns [0,Cp) hold
 * hi = x rounded to TF32 (nearest even), [Cp,2Cp) hold lo = x - hi; padding columns are zero.                    */
int jpb_tf32_split(const float* x, float* out, long long rows, int C, void* stream);

/* ---- space-to-depth form of the stem input for the TMA-patch convolution: x [B,H,W,Cp] (Cp = 4 | 8, H and W even) ->
 * x3 [B,H/2,W/2+1,8*Cp], channel ((dy*4 + dx)*Cp + c) of position (oy, p) = x[b, 2*oy+dy, 2*(p-1)+dx, c] (dy < 2, dx < 4, zero
 * outside the image).                                                                                                        */
int jpb_stem_s2d(const float* x, float* x3, int B, int H, int W, int Cp, void* stream);

/* ---- nearest 2x up-sampling, NHWC: x [B,H,W,C] (C % 4 == 0) -> y [B,2H,2W,C]  (layers.py:16-19) */
int jpb_upsample2x(const float* x, float* y, int B, int H, int W, int C, void* stream);

/* ---- y = ((xs[0] + xs[1]) + xs[2]) + ...  over n <= 8 equally shaped tensors of `count` floats (count % 4 == 0, 16-byte aligned):
 * the residual chain of a CRP block (layers.py:186-199) in one pass.  xs is a HOST array of device pointers.                    */
int jpb_sum_n(const float* const* xs, int n, float* y, long long count, void* stream);

/* ---- zero-padded channel copy: x [rows][C] -> y [rows][Cp], Cp % 4 == 0, Cp >= C */
int jpb_pad_channels(const float* x, float* y, long long rows, int C, int Cp, void* stream);

/* ---- backward of the convolution epilogue: dz = dy * act'(y) (act as in JpbConvArgs, from the OUTPUT y) and
 * dbias[c] += sum over rows of dz.  dz may be NULL (bias gradient only), dbias may be NULL.                  */
int jpb_act_bwd(const float* dy, const float* y, float* dz, long long rows, int C, int act, float* dbias, void* stream);
/* finishing pass of a split-K convolution whose partial tiles were summed without the epilogue: z = act(z + bias + residual)
 * in place; z, residual: [rows][C] NHWC, C % 4 == 0; bias / residual may be NULL.                                  */
int jpb_bias_act(float* z, const float* bias, const float* residual, long long rows, int C, int act, void* stream);

/* ---- BatchNorm2d (training: per-GPU batch statistics) fused with the residual add and ReLU that follow it
 * (resnet.py:28-45, layout_model.py:146-158).  x, res, y, dy, dx, dres: [rows][C] NHWC, C % 4 == 0.
 * stat: [2][C] floats (mean, 1/sqrt(var+eps)) written by the forward and read by the backward.
 * ws: workspace of jpb_bn_workspace_doubles(C) doubles (final sums, per-block partials, a ticket counter); the caller
 * zero-fills it ONCE when allocating it (the kernels leave the counter at zero) and may share one workspace between all
 * BatchNorm calls of a stream.  Each direction is two launches: column sums whose last block folds the partials and
 * finalises (statistics + running_* update + num_batches_tracked += nbt_inc, or the affine gradients), then the apply pass.
 * accumulate != 0: dgamma/dbeta are added to (they alias the flat gradient buffer) instead of overwritten.          */
long long jpb_bn_workspace_doubles(int C);
/* the accumulators inside a BatchNorm workspace that a convolution epilogue (JpbConvArgs.stats) may add into; the next
 * jpb_bn_train_fwd on the same workspace must then be called with stats_ready != 0 (it skips its own statistics pass). */
double* jpb_bn_stats_accumulator(double* ws);
int jpb_bn_train_fwd(const float* x, const float* res, const float* gamma, const float* beta, float* running_mean, float* running_var,
                     long long* num_batches_tracked, int nbt_inc, float momentum, float eps, int relu, float* y, float* stat,
                     double* ws, long long rows, int C, int stats_ready, void* stream);
int jpb_bn_eval_fwd(const float* x, const float* res, const float* gamma, const float* beta, const float* stat, int relu, float* y,
                    long long rows, int C, void* stream);
int jpb_bn_train_bwd(const float* x, const float* dy, const float* y, const float* stat, const float* gamma, int relu, float* dx,
                     float* dres, float* dgamma, float* dbeta, int accumulate, double* ws, long long rows, int C, void* stream,
                     const float* beta /* relu != 0 with y == NULL (BatchNorm + ReLU without residual): the mask y > 0 is re-derived from
                                          x, bit for bit, instead of reading y back; otherwise unused and may be NULL */);

/* ---- NHWC max pooling (nn.MaxPool2d(k, s, p); layers.py:191, resnet.py:91, layout_model.py:84) --------------
 * idx: window-relative arg-max (ky*k + kx) per output element, first maximum wins; C % 4 == 0.
 * Backward: the caller need not zero-fill gx.                                                                   */
int jpb_maxpool_fwd(const float* x, float* y, unsigned char* idx, int B, int H, int W, int C, int k, int s, int p, void* stream);
int jpb_maxpool_bwd(const float* gy, const unsigned char* idx, float* gx, int B, int H, int W, int C, int k, int s, int p, void* stream);
/* Backward schedule: 0 = default (5x5 / stride 1: scatter with red.global.add over a gx the call clears itself; every other
 * geometry: gather, gx written once per element); 1 = scatter for every overlapping window (round 1); 2 = 5x5 as a gather
 * (deterministic; measured slower).  Process-wide.                                                              */
int jpb_maxpool_set_bwd_variant(int variant);

/* ---- flat-buffer optimizer step (mono/core/utils/dist_utils.py:34-60 + torch.optim.Adam) ---------
 * jpb_sumsq: acc[0] += sum g^2 (run on the all-reduced SUM of gradients).
 * jpb_adam_step: g_eff = g * grad_scale (1/world) * min(1, max_norm / (sqrt(normsq)*grad_scale + 1e-6));
 *                Adam(lr, beta1, beta2, eps, L2 weight_decay) with bias correction from the device-side
 *                step counter, which the call increments (CUDA-graph replayable).                    */
typedef struct JpbAdamArgs {
  float lr, beta1, beta2, eps, weight_decay;
  float grad_scale;        /* 1 / world_size */
  float max_norm;          /* <= 0 disables clipping */
  const double* normsq;    /* [1] from jpb_sumsq, or NULL */
  long long* step;         /* [1] device step counter (starts at 0) */
} JpbAdamArgs;
int jpb_sumsq(const float* g, long long n, double* acc, void* stream);
int jpb_adam_step(float* p, const float* g, float* m, float* v, long long n, const JpbAdamArgs* args, void* stream);

/* ---- input prologue of the three ResNet trunks: out[b][y][x][0..Cpad) = ((resize(im) - 0.45) / 0.225, zero padding), NHWC.
 * im0 / im1: [B,3,Hs,Ws] NCHW frames (im1 NULL for a single frame, else the pair is concatenated: pose_encoder.py:84-86,
 * net.py:633-638); bilinear resize (align_corners=False) when (Ho,Wo) != (Hs,Ws); Cpad = 4 (one frame) or 8 (a pair). */
int jpb_image_prep(const float* im0, const float* im1, float* out, int B, int Hs, int Ws, int Ho, int Wo, int Cpad, void* stream);

/* ---- nn.Dropout(p) (depth_decoder.py:47-48): y = x * keep / (1-p); keep from `mask` ([n] of 0/1) when given, else from the
 * counter-based draw (seed, stream_id + 4096*step[0], element index) — the same call with dy as x is the backward.     */
int jpb_dropout(const float* x, const float* mask, float* y, long long n, float p, uint64_t seed, uint64_t stream_id,
                const long long* step, void* stream);

/* ---- pose head: x [B][hw][C>=6] (PoseDecoder output, NHWC) -> spatial mean * 0.01 -> axis-angle / translation ->
 * cam_T_cam [B][4][4] (pose_decoder.py:22-26, net.py:704-756; invert != 0 for the frame before the target).  mean6 [B][6]
 * is written by the forward and read by the backward, which fills gx [B][hw][C] from gT [B][4][4].                    */
int jpb_pose_head_fwd(const float* x, float* T, float* mean6, int B, int hw, int C, int invert, void* stream);
int jpb_pose_head_bwd(const float* gT, const float* mean6, float* gx, int B, int hw, int C, int invert, void* stream);

/* ---- cross-view transformer core (CrossViewTransformer.py:45-92) around its convolutions; all tensors NHWC [B][n][.],
 * n = h*w positions, n <= 256.  select: S[j] = max_i <k_i, q_j> (arg = first maximiser), T[j] = v[arg[j]], attn/argd the same
 * max for the depth pair.  combine: out = front + fused * S + attn @ vd (the reference's broadcast (h x w)(h x w) product over
 * channels; h == w).  The backward entry points return the gradients of every tensor input (g w.r.t. front is g itself). */
int jpb_cct_select_fwd(const float* q, const float* k, const float* v, const float* qd, const float* kd, float* T, float* S, int* arg,
                       float* attn, int* argd, int B, int n, int Cq, int C, void* stream);
int jpb_cct_select_bwd(const float* q, const float* k, const float* qd, const float* kd, const int* arg, const int* argd, const float* gT,
                       const float* gS, const float* gattn, float* gq, float* gk, float* gv, float* gqd, float* gkd, int B, int n, int Cq,
                       int C, void* stream);
int jpb_cct_combine_fwd(const float* front, const float* fused, const float* S, const float* attn, const float* vd, float* out, int B,
                        int h, int w, int C, void* stream);
int jpb_cct_combine_bwd(const float* g, const float* fused, const float* S, const float* attn, const float* vd, float* gfused, float* gS,
                        float* gattn, float* gvd, int B, int h, int w, int C, void* stream);

/* ---- CycledViewProjection transform module (CycledViewProjection.py:27-67): y2 = relu(W2 relu(W1 x + b1) + b2) over the n
 * positions of every (sample, channel); x, y1, y2, g, dz1, dz2, dx: NHWC [B][n][C]; W: [n][n] (out, in).  The backward ADDS
 * the parameter gradients into dW1/db1/dW2/db2 (zero-filled by the caller, or the flat gradient buffer).              */
int jpb_cvp_mlp_fwd(const float* x, const float* W1, const float* b1, const float* W2, const float* b2, float* y1, float* y2, int B, int n,
                    int C, void* stream);
int jpb_cvp_mlp_bwd(const float* x, const float* W1, const float* W2, const float* y1, const float* y2, const float* g, float* dz2, float* dz1,
                    float* dx, float* dW1, float* db1, float* dW2, float* db2, int B, int n, int C, void* stream);

/* ---- batched weight re-layout for the data-gradient GEMMs (autograd's convolution_backward re-lays out each weight
 * separately): dst [Cin][taps][N] = src [N][taps][Cin] with the tap order reversed, for `nent` layers in one launch.
 * entries_dev: device array; block_start = exclusive prefix sum of taps*ceil(N/32)*ceil(Cin/32); nblocks = the total. */
typedef struct JpbWeightT {
  const float* src;
  float* dst;
  int N, Cin, taps, block_start;
} JpbWeightT;
int jpb_weight_flipT(const JpbWeightT* entries_dev, int nent, int nblocks, void* stream);

/* ---- accumulator finalisation: out[i] = (float)(acc[i] / (den ? den[i] : 1) * scale) --------------
 * (the `.mean()` / weight scalings of net.py:175-190, done on device so no loss term syncs the host) */
int jpb_finalize(const double* acc, const double* den, float scale, float* out, int n, void* stream);

/* ---- evaluation metrics on the device (validation hook) ------------------------------------------
 * depth : eval_hooks.py:149-197 (disp_to_depth, cv2.resize to the ground-truth frame, 1/x, validity + crop mask, median
 *         scaling or the fixed stereo scale, clamp) + pixel_error.py:27-40 (compute_errors), one sample per batch entry.
 * BEV   : eval_hooks.py:185-197 + pixel_error.py:62-118 (mean_IU, mean_precision): the three counts both are built from. */
typedef struct JpbDepthEvalArgs {
  const float* disp;          /* [B,h,w]   outputs[("disp",0,0)] (sigmoid output, not yet scaled)                 */
  const float* gt;            /* [B,gh,gw] data['gt_depth']; values outside (min_depth, max_depth) are ignored     */
  int B, h, w, gh, gw;
  float min_disp, max_disp;   /* disp_to_depth: 1/max_depth', 1/min_depth' of the NETWORK range (0.01, 10)         */
  float min_depth, max_depth; /* MIN_DEPTH = 1e-3, MAX_DEPTH = 80 (eval_hooks.py:14-15)                            */
  int crop[4];                /* rows [crop0,crop1) x cols [crop2,crop3) (eval_hooks.py:168-172)                   */
  float fixed_scale;          /* > 0: cfg.data['stereo_scale'] (x36); 0: median scaling                            */
  float* work;                /* [B,2,gh*gw] scratch: compacted (gt, prediction) pairs                             */
  int* count;                 /* [B] valid pixels per sample; caller zero-fills                                    */
  double* out;                /* [B,8] abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3, ratio (NaN when count == 0)    */
} JpbDepthEvalArgs;
int jpb_depth_eval(const JpbDepthEvalArgs* args, void* stream);
/* counts[b] = {#(pred==1 & gt==1), #(pred==1), #(gt==1)} with pred = argmax over the two logits; += (caller zero-fills);
 * logits addressed as base + b*stride_b + c*stride_c + pixel*stride_p (elements); label [B,occ*occ] float {0,1}. */
int jpb_bev_confusion(const float* logits, long long stride_b, long long stride_c, long long stride_p, const float* label,
                      int B, int occ, long long* counts, void* stream);

/* ---- batched image preprocessing (training input pipeline) -------------------------------------------
 * mono_dataset.py:126-171 (preprocess: Resize(ANTIALIAS) twice, ColorJitter, ToTensor), :202-203 (flip / jitter draws),
 * :417-431 (label resize + binarisation).  Bit-exact with Pillow's Resample.c / Blend.c / Convert.c on the same bytes.
 * Frames are uint8 HWC; float outputs are NCHW in [0,1] (ToTensor).  Coefficient / position tables come from the host
 * (built in double precision exactly as Pillow's precompute_coeffs + normalize_coeffs_8bpc do).                        */
typedef struct JpbResizeArgs {
  const unsigned char* src;   /* [B,Hin,Win,3]                                                                     */
  unsigned char* tmp;         /* [B,Hin,Wout,3] scratch: result of the horizontal pass                             */
  unsigned char* dst;         /* [B,Hout,Wout,3] or NULL                                                           */
  float* dst_f;               /* [B,3,Hout,Wout] = ToTensor(dst) or NULL                                           */
  int B, Hin, Win, Hout, Wout;
  const int* kx;              /* [Wout,ksx] 22-bit fixed-point coefficients of the horizontal pass                 */
  const int* bx;              /* [Wout,2]   first source column, tap count                                         */
  int ksx;                    /* 0: Win == Wout (Pillow skips the pass)                                            */
  const int* ky;              /* [Hout,ksy] */
  const int* by;              /* [Hout,2]   */
  int ksy;
  const unsigned char* flip;  /* [B] 1: Image.transpose(FLIP_LEFT_RIGHT) before resizing; NULL: none              */
} JpbResizeArgs;
int jpb_resize_lanczos_u8(const JpbResizeArgs* args, void* stream);

typedef struct JpbJitterArgs {
  const unsigned char* src;   /* [B,H,W,3]                                                                         */
  unsigned char* dst;         /* [B,H,W,3] or NULL                                                                 */
  float* dst_f;               /* [B,3,H,W] = ToTensor(dst) or NULL                                                 */
  int B, H, W;
  const int* order;           /* [B,4] ColorJitter's fn_idx permutation: 0 brightness, 1 contrast, 2 saturation, 3 hue */
  const float* factor;        /* [B,4] brightness, contrast, saturation factors as C floats ([3] unused)           */
  const int* hue_shift;       /* [B]   uint8(int32(hue_factor * 255)) (torchvision _functional_pil.adjust_hue)     */
  const unsigned char* enable;/* [B] do_color_aug per sample (0: pass-through) or NULL (all on)                    */
  unsigned long long* lsum;   /* [B] scratch for the contrast operator's mean; caller zero-fills                   */
} JpbJitterArgs;
int jpb_color_jitter_u8(const JpbJitterArgs* args, void* stream);
/* label[b] = (nearest-resized src == 255) as float; src [B,Hin,Win] uint8, dst [B,size,size]; xtab/ytab [size] source
 * positions (Geometry.c ImagingScaleAffine, accumulated in double on the host).                                       */
int jpb_bev_label_u8(const unsigned char* src, float* dst, int B, int Hin, int Win, int size, const int* xtab, const int* ytab,
                     const unsigned char* flip, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* JPB200_H */
