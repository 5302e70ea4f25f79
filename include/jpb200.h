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
  double* loss_sum;                /* [1] += sum over b,y,x of min over candidates                 */
  long long* min_index;            /* [B,H,W] int64 argmin (outputs[("min_index",s)]) or NULL      */
  unsigned char* winner;           /* [B,H,W] same argmin as a byte (kept for backward) or NULL    */
  float* warped[JPB_MAX_SRC];      /* [B,3,H,W] outputs[("color",f,s)] or NULL                     */
} JpbPhotoArgs;

typedef struct JpbPhotoGrad {
  const float* grad_out;           /* [1] d(total)/d(loss_dict[("min_reconstruct_loss",s)])        */
  float inv_count;                 /* 1 / (num_scales * B*H*W)                                     */
  const unsigned char* winner;     /* [B,H,W] from the forward launch                              */
  float* grad_disp;                /* [B,hs,ws] += ; caller zero-fills                             */
  float* grad_T[JPB_MAX_SRC];      /* [B,4,4] += ; caller zero-fills; NULL to skip                 */
} JpbPhotoGrad;

int jpb_photometric_fwd(const JpbPhotoArgs* args, void* stream);
int jpb_photometric_bwd(const JpbPhotoArgs* args, const JpbPhotoGrad* grad, void* stream);

/* ---- accumulator finalisation: out[i] = (float)(acc[i] / (den ? den[i] : 1) * scale) --------------
 * (the `.mean()` / weight scalings of net.py:175-190, done on device so no loss term syncs the host) */
int jpb_finalize(const double* acc, const double* den, float scale, float* out, int n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* JPB200_H */
