"""CPU fp32 restatement of the JPerceiver training hot path (the parity oracle).

TEST INFRASTRUCTURE ONLY — see ``oracle/__init__.py``.  Not imported by ``jperceiver_b200``.

Plain functional torch (the parity oracle runs it on the CPU in fp32), driven by a flat ``{state_dict key: tensor}`` parameter
dictionary that uses the *reference's* key names, so the same weights can be loaded into the
reference (``oracle/ref_loader.py``), into this port and into the CUDA product.  Every function
cites the reference file:line it restates (paths relative to ``/root/reference``; ``M/`` is
``mono/model/mono_baseline/``).

The same functions run on a CUDA device when parameters and inputs live there and the caller wraps the call in
``with torch.device("cuda"):`` (tensor factories then follow the default device): that is the reference's own arithmetic on
a GPU — eager ATen + cuDNN, TF32 convolutions under torch's default ``cudnn.allow_tf32=True`` — used by
tests/test_model_parity.py to calibrate the product's TF32 deviation and by ``bench.py``'s ``gpu_eager_baseline`` leg.

Parity status
-------------
Pinned against the reference executed in the authoring container (``tests/golden/*.npz`` written by
``oracle/make_golden.py``; ``tests/test_oracle_vs_reference.py`` re-runs the comparison whenever
``/root/reference`` is present) and against SURVEY.md §8c's known-answer vectors.

Un-pinned third-party arithmetic, restated from the published algorithms:
  * torchgeometry 0.1.2 ``warp_perspective`` (``warp_align_corners`` knob, default True = the
    geometrically consistent convention);
  * skimage ``find_boundaries(mode='inner')``.

Pinned semantics where the reference is defective / undefined:
  * per-``type`` loss selection follows ``/net.py:114-159``; ``Argo_both`` follows ``M/net.py:94-138``;
  * ``loss_weightS``/``loss2_weightS`` default to ``loss_weight``/``loss2_weight`` when absent;
  * ``static_eigen`` := depth + pose only (photometric + smoothness);
  * non-square inputs: the layout branch sees the frame bilinearly resized to (4·occ)² and the CCT
    depth feature is ``l4`` bilinearly resized to (occ/8)²  — both are identities at 1024²;
  * ``static`` label mask: a warped-mask pixel counts as "exactly 1" when it is ≥ 1 − 2⁻²⁰
    (the reference's ``uint8`` cast keeps only bit-exact 1.0, which depends on ATen's fp32
    rounding of the bilinear weights and differs between its own CPU and CUDA paths).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

LEAKY = 0.01  # F.leaky_relu default slope (M/depth_decoder.py:60)
ONE_TOL = 1.0 - 2.0 ** -20


# ----------------------------------------------------------------------------------------------
# small helpers
# ----------------------------------------------------------------------------------------------
def _get(opt, key, default=None):
    try:
        return opt[key]
    except (KeyError, AttributeError):
        return default


def conv(P, name, x, stride=1, pad=0, reflect=False):
    """nn.Conv2d, optionally behind ReflectionPad2d (M/layers.py:156-167)."""
    w = P[name + ".weight"]
    b = P.get(name + ".bias")
    if reflect and pad:
        x = F.pad(x, (pad,) * 4, mode="reflect")
        pad = 0
    return F.conv2d(x, w, b, stride=stride, padding=pad)


def bnorm(P, name, x, training, stats=None):
    """nn.BatchNorm2d: batch statistics + running-stat update in training mode."""
    rm, rv = P[name + ".running_mean"], P[name + ".running_var"]
    if training and (name + ".num_batches_tracked") in P:
        P[name + ".num_batches_tracked"] += 1
    return F.batch_norm(x, rm, rv, P[name + ".weight"], P[name + ".bias"], training, 0.1, 1e-5)


# ----------------------------------------------------------------------------------------------
# ResNet-18 trunk  (M/resnet.py:16-136, M/depth_encoder.py:35-44, M/pose_encoder.py:81-92,
#                   M/ResnetEncoder.py:97-110)
# ----------------------------------------------------------------------------------------------
def basic_block(P, pre, x, stride, training):
    y = conv(P, pre + ".conv1", x, stride=stride, pad=1)
    y = F.relu(bnorm(P, pre + ".bn1", y, training))
    y = bnorm(P, pre + ".bn2", conv(P, pre + ".conv2", y, pad=1), training)
    if (pre + ".downsample.0.weight") in P:
        x = bnorm(P, pre + ".downsample.1", conv(P, pre + ".downsample.0", x, stride=stride), training)
    return F.relu(y + x)


def resnet18_features(P, pre, image, training):
    """Five feature levels [64@/2, 64@/4, 128@/8, 256@/16, 512@/32]; input is normalised first."""
    x = (image - 0.45) / 0.225
    x = F.relu(bnorm(P, pre + ".bn1", conv(P, pre + ".conv1", x, stride=2, pad=3), training))
    feats = [x]
    x = F.max_pool2d(x, 3, 2, 1)
    for li, stride in ((1, 1), (2, 2), (3, 2), (4, 2)):
        x = basic_block(P, f"{pre}.layer{li}.0", x, stride, training)
        x = basic_block(P, f"{pre}.layer{li}.1", x, 1, training)
        feats.append(x)
    return feats


# ----------------------------------------------------------------------------------------------
# DepthDecoder  (M/depth_decoder.py:45-137, CRP M/layers.py:184-199)
# ----------------------------------------------------------------------------------------------
def crp(P, pre, x):
    top = x
    for i in range(1, 5):
        top = F.max_pool2d(top, 5, 1, 2)
        top = conv(P, f"{pre}.0.{i}_pointwise.conv", top)
        x = top + x
    return x


def depth_decoder(P, pre, feats, training, drop_masks=None, drop_p=0.5):
    """Returns {scale: disparity}.  ``drop_masks`` = (mask_l4, mask_l3) of 0/1 keeps, or None to draw."""
    l0, l1, l2, l3, l4 = feats
    if training and drop_p > 0:
        if drop_masks is None:
            drop_masks = ((torch.rand_like(l4) >= drop_p).float(), (torch.rand_like(l3) >= drop_p).float())
        l4 = l4 * drop_masks[0] / (1 - drop_p)
        l3 = l3 * drop_masks[1] / (1 - drop_p)
    disp = {}
    x = conv(P, pre + ".reduce4.conv", l4)
    for lvl, skip in ((4, None), (3, l3), (2, l2), (1, l1)):
        if skip is not None:
            x = torch.cat((conv(P, f"{pre}.reduce{lvl}.conv", skip), up, d), 1)
        x = F.leaky_relu(conv(P, f"{pre}.iconv{lvl}.conv", x, pad=1, reflect=True), LEAKY)
        x = crp(P, f"{pre}.crp{lvl}", x)
        x = F.leaky_relu(conv(P, f"{pre}.merge{lvl}.conv", x, pad=1, reflect=True), LEAKY)
        up = F.interpolate(x, scale_factor=2, mode="nearest")
        d = torch.sigmoid(conv(P, f"{pre}.disp{lvl}.0.conv", up, pad=1, reflect=True))
        disp[lvl - 1] = d
    return disp


# ----------------------------------------------------------------------------------------------
# Pose  (M/pose_decoder.py:16-26, M/net.py:630-642,704-756)
# ----------------------------------------------------------------------------------------------
def pose_decoder(P, pre, f4):
    x = F.relu(conv(P, pre + ".reduce", f4))
    x = F.relu(conv(P, pre + ".conv1", x, pad=1))
    x = F.relu(conv(P, pre + ".conv2", x, pad=1))
    x = conv(P, pre + ".conv3", x)
    x = 0.01 * x.mean(3).mean(2)
    return x[:, :3], x[:, 3:]


def rodrigues(aa):
    """Axis-angle (B,3) -> (B,3,3), ``axis = v/(|v|+1e-7)``  (M/net.py:727-756)."""
    ang = aa.norm(dim=1, keepdim=True)
    ax = aa / (ang + 1e-7)
    ca, sa = torch.cos(ang)[:, 0], torch.sin(ang)[:, 0]
    C = 1 - ca
    x, y, z = ax[:, 0], ax[:, 1], ax[:, 2]
    rows = [x * x * C + ca, x * y * C - z * sa, z * x * C + y * sa,
            x * y * C + z * sa, y * y * C + ca, y * z * C - x * sa,
            z * x * C - y * sa, y * z * C + x * sa, z * z * C + ca]
    return torch.stack(rows, 1).view(-1, 3, 3)


def pose_matrix(aa, t, invert):
    """4x4 ``T(t)·R`` or, inverted, ``Rᵀ·T(−t)``  (M/net.py:704-725)."""
    B = aa.shape[0]
    R = torch.zeros(B, 4, 4, dtype=aa.dtype)
    R[:, :3, :3] = rodrigues(aa)
    R[:, 3, 3] = 1
    T = torch.eye(4, dtype=aa.dtype).repeat(B, 1, 1)
    if invert:
        T[:, :3, 3] = -t
        return R.transpose(1, 2) @ T
    T[:, :3, 3] = t
    return T @ R


def predict_poses(P, opt, inputs, training):
    out = {}
    fids = list(opt["frame_ids"])
    small = {f: F.interpolate(inputs[("color_aug", f, 0)], [192, 640], mode="bilinear", align_corners=False)
             for f in fids}
    for f in fids[1:]:
        pair = [small[f], small[0]] if f < 0 else [small[0], small[f]]
        feats = resnet18_features(P, "PoseEncoder.encoder", torch.cat(pair, 1), training)
        aa, t = pose_decoder(P, "PoseDecoder", feats[-1])
        out[("cam_T_cam", 0, f)] = pose_matrix(aa, t, invert=(f < 0))
    return out


# ----------------------------------------------------------------------------------------------
# Layout branch  (M/layout_model.py:76-201, M/CycledViewProjection.py, M/CrossViewTransformer.py)
# ----------------------------------------------------------------------------------------------
def layout_encoder(P, pre, image, training):
    f4 = resnet18_features(P, pre + ".resnet_encoder.encoder", image, training)[-1]
    x = F.max_pool2d(conv(P, pre + ".conv1.conv", f4, pad=1, reflect=True), 2)
    return F.max_pool2d(conv(P, pre + ".conv2.conv", x, pad=1, reflect=True), 2)


def layout_decoder(P, pre, x, training):
    """``decoder.{0..25}`` ModuleList order: per level i=4..0: conv,bn,relu,conv,bn ; then topview."""
    for lvl in range(5):
        k = 5 * lvl
        x = F.relu(bnorm(P, f"{pre}.decoder.{k + 1}", conv(P, f"{pre}.decoder.{k}", x, pad=1), training))
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        x = bnorm(P, f"{pre}.decoder.{k + 4}", conv(P, f"{pre}.decoder.{k + 3}", x, pad=1), training)
    return conv(P, f"{pre}.decoder.25.conv", x, pad=1, reflect=True)


def cvp_mlp(P, pre, x):
    B, C, h, w = x.shape
    y = x.reshape(B, C, h * w)
    y = F.relu(F.linear(y, P[pre + ".fc_transform.0.weight"], P[pre + ".fc_transform.0.bias"]))
    y = F.relu(F.linear(y, P[pre + ".fc_transform.2.weight"], P[pre + ".fc_transform.2.bias"]))
    return y.reshape(B, C, h, w)


def cycled_view_projection(P, pre, x):
    t = cvp_mlp(P, pre + ".transform_module", x)
    return t, cvp_mlp(P, pre + ".retransform_module", t)


def cross_view_transformer(P, pre, front, cross, front_hat, depth_l4):
    """Returns (out, S, attn).  Quirks kept: hard max/argmax over keys; ``attn @ value_d`` is a batched
    8x8 matrix product broadcast over channels (M/CrossViewTransformer.py:45-92)."""
    dfeat = F.max_pool2d(conv(P, pre + ".conv1.conv", depth_l4, pad=1, reflect=True), 2)
    dfeat = F.max_pool2d(conv(P, pre + ".conv2.conv", dfeat, pad=1, reflect=True), 2)
    B, C, a, b = front.shape
    n = a * b
    q = conv(P, pre + ".query_conv", cross).view(B, -1, n)
    k = conv(P, pre + ".key_conv", front).view(B, -1, n).permute(0, 2, 1)
    energy = torch.bmm(k, q)
    star, arg = energy.max(dim=1)
    v = conv(P, pre + ".value_conv", front_hat).view(B, -1, n)
    T = torch.gather(v, 2, arg.view(B, 1, n).expand(-1, v.shape[1], -1)).view(B, -1, a, b)
    S = star.view(B, 1, a, b)
    out = front + conv(P, pre + ".f_conv", torch.cat((front, T), 1), pad=1) * S
    qd = conv(P, pre + ".query_conv_depth", cross).view(B, -1, n)
    kd = conv(P, pre + ".key_conv_depth", front).view(B, -1, n).permute(0, 2, 1)
    vd = conv(P, pre + ".value_conv_depth", dfeat).view(B, -1, a, b)
    attn = (kd @ qd).max(dim=1)[0].view(B, 1, a, b)
    return out + attn @ vd, S, attn


def _layout_head(P, out, feat, l4, training, sfx, car):
    tf, rtf = cycled_view_projection(P, "CycledViewProjection" + sfx, feat)
    fused, S, attn = cross_view_transformer(P, "CrossViewTransformer" + sfx, feat, tf, rtf, l4)
    out["topview" + sfx] = layout_decoder(P, "LayoutDecoder" + sfx, fused, training)
    out["transform_topview" + sfx] = layout_decoder(P, "LayoutTransformDecoder" + sfx, tf, training)
    out["features" + sfx] = out["features_" + car] = fused
    out["transform_feature_" + car] = tf
    out["retransform_features" + sfx] = out["retransform_features_" + car] = rtf
    out["cv_attn_" + car] = S
    out["cm_attn_" + car] = attn


def predict_layout(P, opt, inputs, l4, training, bn_double_update=True):
    """Both BEV heads (M/net.py:644-689).  The reference evaluates the road head twice (M/net.py:73-74);
    forward values are identical, only the road head's BN running stats see two updates — emulated
    by evaluating the road head twice when ``bn_double_update``."""
    occ = opt["occ_map_size"]
    img = inputs[("color_aug", 0, 0)]
    if img.shape[2] != 4 * occ or img.shape[3] != 4 * occ:
        img = F.interpolate(img, (4 * occ, 4 * occ), mode="bilinear", align_corners=False)
    if l4.shape[2] != occ // 8 or l4.shape[3] != occ // 8:
        l4 = F.interpolate(l4, (occ // 8, occ // 8), mode="bilinear", align_corners=False)
    out = {}
    passes = 2 if (training and bn_double_update) else 1
    for _ in range(passes):  # the road head (encoder included) is evaluated twice per step by the reference
        feat = layout_encoder(P, "LayoutEncoder", img, training)
        out["origin_features"] = feat
        _layout_head(P, out, feat, l4, training, "", "road")
    _layout_head(P, out, feat, l4, training, "B", "car")
    return out


# ----------------------------------------------------------------------------------------------
# Photometric chain  (M/layers.py:33-107, M/net.py:84-92,159-175,690-702)
# ----------------------------------------------------------------------------------------------
def disp_to_depth(disp, min_depth=0.1, max_depth=100.0):
    lo, hi = 1.0 / max_depth, 1.0 / min_depth
    return 1.0 / (lo + (hi - lo) * disp)


def pixel_grid(H, W, dtype=torch.float32):
    ys, xs = torch.meshgrid(torch.arange(H, dtype=dtype), torch.arange(W, dtype=dtype), indexing="ij")
    return torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(H * W, dtype=dtype)], 0)


def reproject_grid(depth, K, inv_K, T):
    """Backproject + Project: normalised sampling grid B×H×W×2 (M/layers.py:57-82)."""
    B, _, H, W = depth.shape
    cam = (inv_K[:, :3, :3] @ pixel_grid(H, W, depth.dtype)) * depth.view(B, 1, -1)
    cam = torch.cat([cam, torch.ones(B, 1, H * W, dtype=depth.dtype)], 1)
    p = (K @ T)[:, :3, :] @ cam
    uv = p[:, :2] / (p[:, 2:3] + 1e-7)
    uv = uv.view(B, 2, H, W).permute(0, 2, 3, 1)
    gx = (uv[..., 0] / (W - 1) - 0.5) * 2
    gy = (uv[..., 1] / (H - 1) - 0.5) * 2
    return torch.stack([gx, gy], -1)


def warp_source(disp_s, src, K, inv_K, T, H, W, min_depth=0.1, max_depth=100.0):
    d = F.interpolate(disp_s, [H, W], mode="bilinear", align_corners=False)
    grid = reproject_grid(disp_to_depth(d, min_depth, max_depth), K, inv_K, T)
    return F.grid_sample(src, grid, mode="bilinear", padding_mode="border", align_corners=False)


def ssim_term(x, y):
    x, y = F.pad(x, (1,) * 4, mode="reflect"), F.pad(y, (1,) * 4, mode="reflect")
    mx, my = F.avg_pool2d(x, 3, 1), F.avg_pool2d(y, 3, 1)
    sx = F.avg_pool2d(x * x, 3, 1) - mx * mx
    sy = F.avg_pool2d(y * y, 3, 1) - my * my
    sxy = F.avg_pool2d(x * y, 3, 1) - mx * my
    n = (2 * mx * my + 1e-4) * (2 * sxy + 9e-4)
    d = (mx * mx + my * my + 1e-4) * (sx + sy + 9e-4)
    return torch.clamp((1 - n / d) / 2, 0, 1)


def reprojection_error(pred, target):
    l1 = torch.sqrt((target - pred) ** 2 + 1e-6).mean(1, True)
    return 0.85 * ssim_term(pred, target).mean(1, True) + 0.15 * l1


def photometric_scale(disp_s, target, sources, poses, K, inv_K, automask=True, noise=None,
                      min_depth=0.1, max_depth=100.0):
    """One scale of the min-reprojection loss.  ``sources``/``poses``: lists in frame_ids[1:] order.
    ``noise``: list of B×1×H×W tensors added to the identity terms (None -> no noise).
    Returns (mean(min)  [not yet /num_scales], min_index int64 B×H×W, warped list)."""
    B, _, H, W = target.shape
    cands = []
    if automask:
        for i, s in enumerate(sources):
            e = reprojection_error(s, target)
            if noise is not None:
                e = e + noise[i]
            cands.append(e)
    warped = [warp_source(disp_s, s, K, inv_K, T, H, W, min_depth, max_depth) for s, T in zip(sources, poses)]
    cands += [reprojection_error(w, target) for w in warped]
    m, idx = torch.cat(cands, 1).min(dim=1)
    return m.mean(), idx, warped


# ----------------------------------------------------------------------------------------------
# Smoothness  (M/net.py:182-190,758-786)
# ----------------------------------------------------------------------------------------------
def _dx(a):
    return a[:, :, :, 1:] - a[:, :, :, :-1]


def _dy(a):
    return a[:, :, 1:] - a[:, :, :-1]


def smooth_term(disp, img, disp_norm=True):
    if disp_norm:
        disp = disp / (disp.mean(2, True).mean(3, True) + 1e-7)
    h, w = disp.shape[2:]
    img = F.interpolate(img, (h, w), mode="area")
    total = 0
    for op in ((_dx,), (_dy,), (_dx, _dx), (_dx, _dy), (_dy, _dx), (_dy, _dy)):
        d, g = disp, img
        for f in op:
            d, g = f(d), f(g)
        total = total + (d.abs() * torch.exp(-0.5 * g.abs().mean(1, True))).mean()
    return total


# ----------------------------------------------------------------------------------------------
# CGT scale label + loss  (M/net.py:193-310,403-476,529-543 ; M/layers.py:214-252)
# ----------------------------------------------------------------------------------------------
def warp_perspective(src, M, dsize, align_corners=True):
    """torchgeometry 0.1.2 restated (parity unpinned): dst_pix <- src_pix homography ``M``."""
    b, _, hs, ws = src.shape
    hd, wd = dsize

    def norm_px(h, w):
        return torch.tensor([[2.0 / (w - 1), 0, -1], [0, 2.0 / (h - 1), -1], [0, 0, 1]], dtype=M.dtype)

    dstn_T_srcn = norm_px(hd, wd) @ (M @ torch.inverse(norm_px(hs, ws)))
    srcn_T_dstn = torch.inverse(dstn_T_srcn)
    gy, gx = torch.meshgrid(torch.linspace(-1, 1, hd), torch.linspace(-1, 1, wd), indexing="ij")
    pts = torch.stack([gx, gy, torch.ones_like(gx)], -1).view(1, -1, 3, 1).to(M.dtype)
    q = (srcn_T_dstn.unsqueeze(1) @ pts).squeeze(-1)
    flow = (q[..., :2] / q[..., 2:3]).view(b, hd, wd, 2)
    return F.grid_sample(src, flow, mode="bilinear", padding_mode="zeros", align_corners=align_corners)


def bev_to_image_homography(K3, Tr, split, occ):
    """``inverse(shiftedground_H_img)``: BEV-pixel -> image-pixel homography (M/net.py:250-285)."""
    B = K3.shape[0]
    h = 0.33 if split == "argo" else 1.73
    ego_T_ground = torch.eye(4, dtype=K3.dtype).repeat(B, 1, 1)
    ego_T_ground[:, 2, 3] = -h  # inverse of SE3(I, [0,0,h])
    cam_T_ego = torch.eye(4, dtype=K3.dtype).repeat(B, 1, 1)
    cam_T_ego[:, :3, :3] = Tr[:, :3, :3]
    cam_T_ego[:, :3, 3] = Tr[:, :3, 3]
    cam_T_ground = cam_T_ego @ ego_T_ground
    img_H_ground = K3 @ torch.stack([cam_T_ground[:, :3, 0], cam_T_ground[:, :3, 1], cam_T_ground[:, :3, 3]], 2)
    s = occ / 40.0
    shift = torch.tensor([[s, 0, 0], [0, s, float(int(occ) // 2)], [0, 0, 1]], dtype=K3.dtype)
    return torch.inverse(shift @ torch.inverse(img_H_ground))


def static_quad_mask(Minv0, occ, height, width):
    """cv2-rasterised projection of the BEV rectangle, from sample 0 only (M/net.py:235-248,292-306)."""
    import cv2

    r1 = occ / 40
    pr = [(round(18 * r1), round(31 * r1)), (round(22 * r1), round(31 * r1)),
          (round(18 * r1), round(33 * r1)), (round(22 * r1), round(33 * r1))]
    rot = [[occ - pr[3][1] - 1, pr[0][0] - 1],
           [occ - pr[3][1] + (pr[2][1] - pr[1][1]) - 1, pr[0][0] - 1],
           [occ - pr[3][1] - 1, pr[1][0] - 1],
           [occ - pr[3][1] + (pr[2][1] - pr[1][1]) - 1, pr[1][0] - 1]]
    pts = torch.tensor(rot, dtype=torch.float32)
    ph = torch.cat([pts, torch.ones(4, 1)], 1) @ Minv0.float().T.to(pts.device)
    proj = torch.round(ph[:, :2] / ph[:, 2:3]).int().cpu().numpy()
    poly = np.array([proj[0], proj[2], proj[3], proj[1]], dtype=np.int32).reshape(-1, 1, 2)
    canvas = np.zeros((height, width, 3), dtype=np.uint8)
    canvas = cv2.fillConvexPoly(canvas, poly, (0, 255, 255), 1)
    gray = cv2.cvtColor(canvas, cv2.COLOR_RGB2GRAY)
    return torch.from_numpy((gray > 0).astype(np.float32)).to(Minv0.device)


def scale_label(opt, inputs, warp_align_corners=True):
    typ, split, occ = opt["type"], opt["split"], opt["occ_map_size"]
    height, width = inputs[("color", 0, -1)].shape[2:4]
    dyn = typ in ("dynamic", "Argo_dynamic")
    lab = inputs[("both_dynamic", 0, 0)] if typ == "Argo_both" else inputs[("bothS", 0, 0)]
    B = lab.shape[0]
    # get_scale_label_dynamic (M/net.py:311-402) drops the 0.27 m offset for KITTI (:325-326) and never warps the label:
    # it reads bothS only for its shape (:316-320), the z-map is masked by the cv2 quad alone (:390-401)
    delta = 1.9 if split == "argo" else (0.0 if dyn else 0.27)
    z = (torch.arange(occ, 0, -1, dtype=torch.float32) * (40.0 / occ) - delta).view(1, 1, occ, 1).repeat(B, 1, 1, occ)
    lab = torch.rot90(lab.float(), 3, (2, 3))  # fliplr on dim 1 (size 1) is a no-op; rotate(270) == rot90(k=3)
    z = torch.rot90(z, 3, (2, 3))
    Minv = bev_to_image_homography(inputs[("odometry_K", 0, 0)][:, :3, :3].float(),
                                   inputs[("Tr_cam2_velo", 0, 0)].float(), split, occ)
    wl = warp_perspective(lab, Minv, (height, width), warp_align_corners)
    wz = warp_perspective(z, Minv, (height, width), warp_align_corners)
    if typ == "Argo_both":
        return wz * wl
    quad = static_quad_mask(Minv[0], occ, height, width)
    if dyn:
        return wz * quad.view(1, 1, height, width)
    return wz * ((wl >= ONE_TOL).float() * quad.view(1, 1, height, width))


def scale_term(disp_s, label, typ, min_depth=0.1, max_depth=100.0):
    pred = F.interpolate(disp_to_depth(disp_s, min_depth, max_depth), label.shape[2:4], mode="bilinear",
                         align_corners=False).clamp(1e-3, 80)
    mask = label > 0
    if typ == "static_raw":
        crop = torch.zeros_like(mask)
        crop[:, :, 153:371, 44:1197] = True
        mask = mask & crop
    g, p = label[mask], pred[mask]
    return ((g - p).abs() / g).mean()


# ----------------------------------------------------------------------------------------------
# BEV head losses  (M/net.py:554-622, M/dice_loss.py:31-81,293-331, M/boundary_loss.py:121-192)
# ----------------------------------------------------------------------------------------------
def signed_distance(mask_np):
    """SDF of one binary map: EDT(outside) − EDT(inside), 0 on the inner 4-connected boundary."""
    from scipy import ndimage as ndi

    pos = mask_np.astype(bool)
    if not pos.any():
        return np.zeros(mask_np.shape, dtype=np.float64)
    sdf = ndi.distance_transform_edt(~pos) - ndi.distance_transform_edt(pos)
    cross = ndi.generate_binary_structure(2, 1)
    u8 = pos.astype(np.uint8)
    inner = (ndi.grey_dilation(u8, footprint=cross) != ndi.grey_erosion(u8, footprint=cross)) & pos
    sdf[inner] = 0
    return sdf


def bev_head_loss(logits, label, w_fg, loss_weight=20.0, loss2_weight=20.0, loss_type="iou", loss_sum=3):
    """``compute_topview_loss`` (M/net.py:554-617): region term selected by ``loss_type`` — soft IoU (M/dice_loss.py:293-331),
    soft Dice (:255-291), Tversky alpha=.3 beta=.7 (:333-372), focal alpha=.25 gamma=2 smooth=1e-5 (M/focal_loss.py:7-92) —
    times ``loss_weight``; ``loss_sum`` 2 adds ``loss2_weight`` x boundary loss, 3 adds the weighted cross entropy as well."""
    B = logits.shape[0]
    y = label.reshape(B, logits.shape[2], logits.shape[3]).long()
    p = F.softmax(logits, 1)
    oh = F.one_hot(y, 2).permute(0, 3, 1, 2).to(p.dtype)
    if loss_type == "focal":
        s = 1e-5
        key = torch.clamp(oh, s / (2 - 1), 1.0 - s)
        pt = (key * p).sum(1) + s
        alpha = torch.where(y == 0, torch.tensor(0.25), torch.tensor(0.75)).to(p.dtype)
        region = (-alpha * (1 - pt) ** 2 * pt.log()).mean()
    else:
        tp = (p * oh).sum((2, 3))
        fp = (p * (1 - oh)).sum((2, 3))
        fn = ((1 - p) * oh).sum((2, 3))
        if loss_type == "iou":
            region = -((tp + 1) / (tp + fp + fn + 1)).mean()
        elif loss_type == "dice":
            region = -((2 * tp + 1) / (2 * tp + fp + fn + 1)).mean()
        elif loss_type == "tversky":
            region = -((tp + 1) / (tp + 0.3 * fp + 0.7 * fn + 1)).mean()
        else:
            raise ValueError(loss_type)
    if loss_sum == 1:
        return loss_weight * region
    phi = torch.from_numpy(np.stack([signed_distance(m) for m in y.cpu().numpy()])).to(torch.float32).to(logits.device)
    bd = (p[:, 1] * phi).mean()
    if loss_sum == 2:
        return loss_weight * region + loss2_weight * bd
    ce = F.cross_entropy(logits, y, weight=torch.tensor([1.0, float(w_fg)]))
    return loss_weight * region + ce + loss2_weight * bd      # the reference's order of the three terms (net.py:583-585)


# ----------------------------------------------------------------------------------------------
# compute_losses + full forward  (M/net.py:68-82,94-192 ; /net.py:114-159 ; apis/trainer.py:33-46)
# ----------------------------------------------------------------------------------------------
ROAD_TYPES = ("static", "static_raw", "Argo_static", "Argo_both")
CAR_TYPES = ("dynamic", "Argo_dynamic", "Argo_both")
LABEL_TYPES = ROAD_TYPES + ("dynamic", "Argo_dynamic")   # every type of /net.py:119-124 builds a CGT label


def compute_losses(opt, inputs, outputs, noise=None, warp_align_corners=True):
    """``noise``: {scale: [B×1×H×W per source]} for the automask identity terms, or None (no noise)."""
    typ = opt["type"]
    L = {}
    lw, l2w = opt["loss_weight"], opt["loss2_weight"]
    lwS, l2wS = _get(opt, "loss_weightS", lw), _get(opt, "loss2_weightS", l2w)
    bev_kw = dict(loss_type=_get(opt, "loss_type", "iou"), loss_sum=_get(opt, "loss_sum", 3))
    if typ in ROAD_TYPES:
        y = inputs[("bothS", 0, 0)]
        L["topview_loss"] = bev_head_loss(outputs["topview"], y, opt["static_weight"], lwS, l2wS, **bev_kw)
        L["transform_topview_loss"] = bev_head_loss(outputs["transform_topview"], y, opt["static_weight"], lwS, l2wS, **bev_kw)
        L["transform_loss"] = (outputs["features"] - outputs["retransform_features"]).abs().mean()
        L["layout_loss"] = L["topview_loss"] + 0.001 * L["transform_loss"] + L["transform_topview_loss"]
    if typ in CAR_TYPES:
        y = inputs[("bothD", 0, 0)]
        L["topview_lossB"] = bev_head_loss(outputs["topviewB"], y, opt["dynamic_weight"], lw, l2w, **bev_kw)
        L["transform_topview_lossB"] = bev_head_loss(outputs["transform_topviewB"], y, opt["dynamic_weight"], lw, l2w, **bev_kw)
        L["transform_lossB"] = (outputs["featuresB"] - outputs["retransform_featuresB"]).abs().mean()
        L["layout_lossB"] = L["topview_lossB"] + 0.001 * L["transform_lossB"] + L["transform_topview_lossB"]
    label = None
    if typ in LABEL_TYPES:
        label = scale_label(opt, inputs, warp_align_corners)
        outputs["scale_label"] = label
    fids = list(opt["frame_ids"])
    nsc = len(opt["scales"])
    target = inputs[("color", 0, 0)]
    H, W = opt["height"], opt["width"]
    for s in opt["scales"]:
        disp = outputs[("disp", 0, s)]
        outputs[("depth", 0, s)] = disp_to_depth(disp, opt["min_depth"], opt["max_depth"])
        m, idx, warped = photometric_scale(
            disp, target, [inputs[("color", f, 0)] for f in fids[1:]],
            [outputs[("cam_T_cam", 0, f)] for f in fids[1:]], inputs[("K", 0)], inputs[("inv_K", 0)],
            automask=opt["automask"], noise=None if noise is None else noise[s],
            min_depth=opt["min_depth"], max_depth=opt["max_depth"])
        for f, wimg in zip(fids[1:], warped):
            outputs[("color", f, s)] = wimg
        outputs[("min_index", s)] = idx
        L[("min_reconstruct_loss", s)] = m / nsc
        if label is not None:
            L[("scale_loss", s)] = opt["scale_weight"] * scale_term(disp, label, typ, opt["min_depth"],
                                                                    opt["max_depth"]) / (2 ** s) / nsc
        L[("smooth_loss", s)] = opt["smoothness_weight"] * smooth_term(disp, target, opt["disp_norm"]) / (2 ** s) / nsc
    return L


def total_loss(loss_dict):
    """The trainer sums *every* entry, double-counting the layout terms (apis/trainer.py:33-46)."""
    return sum(v for v in loss_dict.values())


def forward(P, opt, inputs, training=True, drop_masks=None, drop_p=0.5, noise=None,
            warp_align_corners=True, bn_double_update=True):
    """``Baseline.forward``: returns (outputs, loss_dict) in training mode, outputs otherwise."""
    feats = resnet18_features(P, "DepthEncoder.encoder", inputs[("color_aug", 0, 0)], training)
    outputs = {("disp", 0, s): d for s, d in
               depth_decoder(P, "DepthDecoder", feats, training, drop_masks, drop_p).items()}
    if opt["type"] != "static_eigen":
        outputs.update(predict_layout(P, opt, inputs, feats[-1], training, bn_double_update))
    if not training:
        return outputs
    outputs.update(predict_poses(P, opt, inputs, training))
    return outputs, compute_losses(opt, inputs, outputs, noise, warp_align_corners)


# ----------------------------------------------------------------------------------------------
# deterministic weights + synthetic inputs shared by the oracle, the reference runs and the product
# ----------------------------------------------------------------------------------------------
def synth_params(template, seed=0):
    """Fill a {key: tensor} template (shapes from any Baseline state_dict) with reproducible values:
    conv/linear weights ~ N(0, 2/fan_in)·0.7, biases small, BN weight≈1, running_var≈1."""
    out = {}
    for i, (k, v) in enumerate(sorted(template.items())):
        g = torch.Generator().manual_seed(seed * 100003 + i)
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith("running_mean"):
            out[k] = 0.05 * torch.randn(v.shape, generator=g)
        elif k.endswith("running_var"):
            out[k] = 1.0 + 0.1 * torch.rand(v.shape, generator=g)
        elif v.dim() == 1 and k.endswith("weight"):
            out[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        elif v.dim() == 1:
            out[k] = 0.05 * torch.randn(v.shape, generator=g)
        else:
            fan_in = v[0].numel()
            out[k] = torch.randn(v.shape, generator=g) * (0.7 * math.sqrt(2.0 / fan_in))
    return out


def synth_inputs(opt, B, seed=1, hw_full=(375, 1242)):
    """Synthetic batch with the reference's dict keys: smooth field + noise frames."""
    g = torch.Generator().manual_seed(seed)
    H, W, occ = opt["height"], opt["width"], opt["occ_map_size"]
    inp = {}
    base = F.interpolate(torch.rand(B, 3, H // 16 + 2, W // 16 + 2, generator=g), (H, W), mode="bicubic",
                         align_corners=False).clamp(0, 1)
    for f in opt["frame_ids"]:
        shift = torch.roll(base, shifts=(2 * f, 5 * f), dims=(2, 3))
        img = (0.85 * shift + 0.15 * torch.rand(B, 3, H, W, generator=g)).clamp(0, 1)
        inp[("color", f, 0)] = img
        inp[("color_aug", f, 0)] = img.clone()
    inp[("color", 0, -1)] = torch.zeros(B, 3, *hw_full)
    K = torch.tensor([[0.58 * W, 0, 0.5 * W, 0], [0, 1.92 * H, 0.5 * H, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
    inp[("K", 0)] = K.repeat(B, 1, 1)
    inp[("inv_K", 0)] = torch.linalg.pinv(K).repeat(B, 1, 1)
    oK = torch.tensor([[718.856, 0, 607.1928], [0, 718.856, 185.2157], [0, 0, 1.0]])
    inp[("odometry_K", 0, 0)] = oK.repeat(B, 1, 1)
    Tr = torch.tensor([[4.276802385584e-04, -9.999672484946e-01, -8.084491683471e-03, -1.198459927713e-02],
                       [-7.210626507497e-03, 8.081198471645e-03, -9.999413164504e-01, -5.403984729748e-02],
                       [9.999738645903e-01, 4.859485810390e-04, -7.206933692422e-03, -2.921968648686e-01],
                       [0, 0, 0, 1.0]])
    inp[("Tr_cam2_velo", 0, 0)] = Tr.repeat(B, 1, 1)
    yy, xx = torch.meshgrid(torch.arange(occ), torch.arange(occ), indexing="ij")
    for name, frac in (("bothS", 0.55), ("bothD", 0.12), ("both_dynamic", 0.45)):
        m = torch.zeros(B, 1, occ, occ)
        for b in range(B):
            cy, cx = (torch.rand(2, generator=g) * 0.3 + 0.35) * occ
            ry, rx = (torch.rand(2, generator=g) * 0.5 + 0.5) * frac * occ
            m[b, 0] = ((((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2) <= 1).float()
            if name != "bothD":   # the road always covers the strip ahead of the ego car (so the CGT "assumption region" is labelled)
                m[b, 0, int(0.5 * occ):, int(0.3 * occ):int(0.7 * occ)] = 1   # ego car sits at the bottom centre of the BEV map
        inp[(name, 0, 0)] = m
    return inp
