"""Generate ``tests/golden/*`` by executing the REAL reference (authoring container only).

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.   Usage:  ``python -m oracle.make_golden``

Every value written here comes out of the reference's own functions, imported from
``/root/reference`` through ``oracle/ref_loader.py``; the port (``oracle/port.py``) is not involved.
Fixtures:
  state_dict_shapes.json   key -> shape of ``Baseline(opt).state_dict()`` (the checkpoint contract)
  kat.npz                  SURVEY.md §8c closed-form known-answer vectors KAT0..KAT8
  e2e_<type>_1024.npz      one full training-mode forward + backward at 1024², B=1, frames [0,-1,1],
                           weights = ``port.synth_params(seed=3)``, inputs = ``port.synth_inputs(seed=1)``,
                           dropout p forced to 0, automask noise forced to 0, warp align_corners=True:
                           every loss scalar, strided samples of the outputs, gradient checksums.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import port as O  # noqa: E402  (only for synth_params / synth_inputs: data, not arithmetic)
from oracle import ref_loader as R  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")

GRAD_KEYS = [
    "DepthEncoder.encoder.conv1.weight", "DepthEncoder.encoder.layer1.0.bn1.weight",
    "DepthEncoder.encoder.layer2.0.downsample.0.weight", "DepthEncoder.encoder.layer4.1.conv2.weight",
    "DepthDecoder.reduce4.conv.weight", "DepthDecoder.iconv1.conv.weight", "DepthDecoder.iconv1.conv.bias",
    "DepthDecoder.crp2.0.3_pointwise.conv.weight", "DepthDecoder.merge3.conv.weight",
    "DepthDecoder.disp1.0.conv.weight", "DepthDecoder.disp4.0.conv.bias",
    "PoseEncoder.encoder.conv1.weight", "PoseEncoder.encoder.layer3.1.bn2.bias", "PoseDecoder.conv3.weight",
    "PoseDecoder.reduce.bias", "LayoutEncoder.resnet_encoder.encoder.layer1.0.conv1.weight",
    "LayoutEncoder.conv2.conv.weight", "CycledViewProjection.transform_module.fc_transform.0.weight",
    "CycledViewProjection.retransform_module.fc_transform.2.bias", "CrossViewTransformer.query_conv.weight",
    "CrossViewTransformer.f_conv.weight", "CrossViewTransformer.value_conv_depth.bias",
    "CrossViewTransformer.conv1.conv.weight", "LayoutDecoder.decoder.0.weight", "LayoutDecoder.decoder.11.weight",
    "LayoutTransformDecoder.decoder.25.conv.weight", "CrossViewTransformerB.key_conv.weight",
    "LayoutDecoderB.decoder.3.weight", "LayoutTransformDecoderB.decoder.24.bias",
]


def pat(shape, k):
    b, c, h, w = shape
    B, C, I, J = torch.meshgrid(torch.arange(b), torch.arange(c), torch.arange(h), torch.arange(w), indexing="ij")
    return ((7 * C + 13 * I + 29 * J + 3 * k + 11 * B) % 31).float() / 31


def bare_baseline(net, opt):
    """A ``Baseline`` with only the loss-side members (no 53 M-parameter networks) for the KATs."""
    m = net.Baseline.__new__(net.Baseline)
    torch.nn.Module.__init__(m)
    m.opt = opt
    m.ssim = net.SSIM()
    m.backproject = net.Backproject(opt.imgs_per_gpu, opt.height, opt.width)
    m.project_3d = net.Project(opt.imgs_per_gpu, opt.height, opt.width)
    return m


def make_kats():
    net = R.load("registered")
    out = {}
    with R.cpu_cuda_identity():
        opt = R.default_options(height=8, width=12, imgs_per_gpu=1)
        m = bare_baseline(net, opt)
        out["kat0"] = np.float64(net.disp_to_depth(torch.tensor(0.5), 0.1, 100)[1].item())
        r = m.compute_reprojection_loss(pat((1, 3, 8, 12), 1), pat((1, 3, 8, 12), 2))
        out["kat1"] = r.numpy()
        aa, t = torch.tensor([[[.01, -.02, .03]]]), torch.tensor([[[.1, -.05, .2]]])
        Tf = m.transformation_from_parameters(aa, t, invert=False)
        Ti = m.transformation_from_parameters(aa, t, invert=True)
        out["kat2_fwd"], out["kat2_inv"] = Tf.numpy(), Ti.numpy()
        H, W = 8, 12
        K = torch.tensor([[.58 * W, 0, .5 * W, 0], [0, 1.92 * H, .5 * H, 0], [0, 0, 1, 0], [0, 0, 0, 1]]).unsqueeze(0)
        invK = torch.linalg.pinv(K)
        depth = 1 + 4 * pat((1, 1, 8, 12), 3)
        cam = m.backproject(depth, invK)
        pix = m.project_3d(cam, K, Tf)
        out["kat3_cam"], out["kat3_pix"] = cam.numpy(), pix.numpy()
        out["kat3_sample"] = torch.nn.functional.grid_sample(pat((1, 3, 8, 12), 4), pix, padding_mode="border").numpy()
        d = pat((1, 1, 4, 6), 5)
        out["kat4"] = np.float64(m.get_smooth_loss(d / (d.mean(2, True).mean(3, True) + 1e-7), pat((1, 3, 8, 12), 2)).item())
        # KAT5: one-scale photometric, no noise
        opt5 = R.default_options(height=8, width=12, imgs_per_gpu=1, frame_ids=[0, -1, 1])
        m5 = bare_baseline(net, opt5)
        inputs = {("color", 0, 0): pat((1, 3, 8, 12), 2), ("color", -1, 0): pat((1, 3, 8, 12), 6),
                  ("color", 1, 0): pat((1, 3, 8, 12), 7), ("K", 0): K, ("inv_K", 0): invK}
        outputs = {("disp", 0, 0): 0.1 + 0.8 * pat((1, 1, 4, 6), 8), ("cam_T_cam", 0, -1): Ti, ("cam_T_cam", 0, 1): Tf}
        outputs = m5.generate_images_pred(inputs, outputs, 0)
        tgt = inputs[("color", 0, 0)]
        c = [m5.compute_reprojection_loss(inputs[("color", f, 0)], tgt) for f in (-1, 1)]
        c += [m5.compute_reprojection_loss(outputs[("color", f, 0)], tgt) for f in (-1, 1)]
        mn, idx = torch.cat(c, 1).min(1)
        out["kat5_mean"] = np.float64(mn.mean().item())
        out["kat5_hist"] = np.bincount(idx.flatten().numpy(), minlength=4)
        out["kat5_min"] = mn.numpy()
        out["kat5_warp_m1"] = outputs[("color", -1, 0)].numpy()
        # KAT6/7: BEV head losses
        from _jref.model.mono_baseline.dice_loss import IoULoss
        from _jref.model.mono_baseline.boundary_loss import BDLoss, compute_sdf

        logits = torch.cat([4 * pat((1, 1, 16, 16), 9) - 2, 4 * pat((1, 1, 16, 16), 10) - 2], 1)
        lab = torch.zeros(1, 16, 16, dtype=torch.long)
        lab[:, 4:11, 3:9] = 1
        out["kat6_iou"] = np.float64(IoULoss(apply_nonlin=lambda x: torch.softmax(x, 1))(logits, lab).item())
        out["kat6_ce"] = np.float64(torch.nn.CrossEntropyLoss(weight=torch.tensor([1., 5.]))(logits, lab).item())
        out["kat6_bd"] = np.float64(BDLoss()(logits, lab).item())
        oh = torch.nn.functional.one_hot(lab, 2).permute(0, 3, 1, 2).numpy()
        out["kat6_sdf"] = compute_sdf(oh, oh.shape)[0, 1]
        opt7 = R.default_options()
        m7 = bare_baseline(net, opt7)
        big = torch.cat([4 * pat((2, 1, 256, 256), 9) - 2, 4 * pat((2, 1, 256, 256), 10) - 2], 1)
        lab7 = torch.zeros(2, 1, 256, 256)
        lab7[:, :, 64:176, 48:144] = 1
        lab7[1, :, 200:240, 10:250] = 1
        out["kat7_b1_w5"] = np.float64(m7.compute_topview_loss(big[:1], lab7[:1], torch.Tensor([1., 5.]), opt7).item())
        out["kat7_b1_w15"] = np.float64(m7.compute_topview_lossB(big[:1], lab7[:1], torch.Tensor([1., 15.]), opt7).item())
        out["kat7_b2_w5"] = np.float64(m7.compute_topview_loss(big, lab7, torch.Tensor([1., 5.]), opt7).item())
        out["kat8"] = np.float64(m7.compute_transform_losses(pat((1, 128, 8, 8), 11), pat((1, 128, 8, 8), 12)).item())
    np.savez_compressed(os.path.join(GOLD, "kat.npz"), **out)
    print("kat.npz:", {k: (v.shape if hasattr(v, "shape") and v.shape else float(v)) for k, v in out.items()})


def sample(t, n=1024):
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step][:n].double().numpy()


def make_e2e(typ):
    split = "argo" if typ.startswith("Argo") else "odometry"
    opt = R.default_options(frame_ids=[0, -1, 1], height=1024, width=1024, type=typ, split=split)
    if not typ.startswith("Argo"):
        opt.pop("loss_weightS"), opt.pop("loss2_weightS")  # odometry/raw configs do not define them
        opt["loss_weightS"], opt["loss2_weightS"] = opt["loss_weight"], opt["loss2_weight"]
    torch.manual_seed(0)
    model = R.build_baseline(opt)
    shapes = {k: list(v.shape) for k, v in model.state_dict().items()}
    if typ == "Argo_both":
        with open(os.path.join(GOLD, "state_dict_shapes.json"), "w") as f:
            json.dump(shapes, f, indent=0)
    P = O.synth_params(model.state_dict(), seed=3)
    model.load_state_dict(P)
    model.train()
    model.DepthDecoder.do.p = 0.0
    hw = (2056, 2464) if split == "argo" else (375, 1242)
    inp = O.synth_inputs(opt, 1, seed=1, hw_full=hw)
    R.set_warp_align_corners(True)
    real_randn = torch.randn
    torch.randn = lambda *a, **k: torch.zeros(*a, **k)
    try:
        with R.cpu_cuda_identity():
            outs, losses = model({k: v.clone() for k, v in inp.items()})
            total = sum(v.mean() for v in losses.values())  # apis/trainer.py:33-46
            total.backward()
    finally:
        torch.randn = real_randn
    rec = {"total_loss": np.float64(total.item())}
    for k, v in losses.items():
        rec["loss/" + str(k)] = np.float64(float(v))
    for k, v in outs.items():
        if not torch.is_tensor(v):
            continue
        if v.dtype == torch.int64:
            rec["out/" + str(k) + "/hist"] = np.bincount(v.flatten().numpy(), minlength=4)
        else:
            rec["out/" + str(k)] = sample(v)
            rec["out/" + str(k) + "/sum"] = np.float64(v.double().sum().item())
    named = dict(model.named_parameters())
    for k in GRAD_KEYS:
        g = named[k].grad
        if g is None:  # heads without a loss under this ``type``
            g = torch.zeros_like(named[k])
        rec["grad/" + k] = np.array([g.double().sum().item(), g.double().abs().sum().item(), g.double().norm().item()])
        rec["gradv/" + k] = sample(g, 64)
    nograd = sorted(k for k, p in named.items() if p.grad is None)
    rec["nograd"] = np.array(nograd)
    gn = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in named.values() if p.grad is not None))
    rec["grad_total_norm"] = np.float64(gn.item())
    sd = model.state_dict()
    for k in ("DepthEncoder.encoder.bn1.running_mean", "LayoutEncoder.resnet_encoder.encoder.bn1.running_var",
              "LayoutDecoder.decoder.1.running_var", "LayoutDecoderB.decoder.1.running_var",
              "PoseEncoder.encoder.bn1.running_mean", "LayoutDecoder.decoder.1.num_batches_tracked",
              "PoseEncoder.encoder.bn1.num_batches_tracked"):
        rec["buf/" + k] = sd[k].double().numpy()
    np.savez_compressed(os.path.join(GOLD, f"e2e_{typ}_1024.npz"), **rec)
    print(f"e2e_{typ}_1024.npz: total {total.item():.6f}, |g| {gn.item():.6f}, nograd {len(nograd)}")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    os.makedirs(GOLD, exist_ok=True)
    only = sys.argv[1:]
    if not only:
        make_kats()
    for typ in only or ("Argo_both", "static", "static_raw", "dynamic", "Argo_static", "Argo_dynamic"):
        make_e2e(typ)
