"""The sub-networks of ``Baseline`` with the reference's constructor signatures and state_dict keys.

Forward passes are written against ``jperceiver_b200.netops`` (fused channels-last operators), not
against torch.nn layers: reflection padding, nearest up-sampling, channel concatenation, bias,
activation and residual adds are arguments of the convolution operator rather than separate passes.

Reference (paths under /root/reference/mono/model/mono_baseline/):
  DepthEncoder depth_encoder.py:10-44 · PoseEncoder pose_encoder.py:55-92 · ResnetEncoder ResnetEncoder.py:70-110
  DepthDecoder depth_decoder.py:8-137 (+ layers.py:146-199) · PoseDecoder pose_decoder.py:5-26
  Encoder/Decoder layout_model.py:54-201 · CycledViewProjection.py:11-67 · CrossViewTransformer.py:30-92
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from ... import netops as ops
from .params import BNP, BlockP, ConvP, Holder, LinearP, ResNet18P, Wrap


def _load_pretrained(module, path, num_input_images=1):
    if path is None:
        return
    sd = torch.load(path, map_location="cpu")
    if num_input_images > 1:
        sd["conv1.weight"] = torch.cat([sd["conv1.weight"]] * num_input_images, 1) / num_input_images
    module.load_state_dict(sd)


def resnet18_forward(net: ResNet18P, x, training):
    """x: normalised channels-last input.  Returns the five feature levels."""
    # bn_next: the convolution's epilogue accumulates the BatchNorm statistics of its output, so every conv -> BN pair must be
    # adjacent (the shortcut branch of a down-sampling block is therefore evaluated first; the result is the same)
    x = ops.conv2d(x, net.conv1.weight, stride=2, pad=3, bn_next=training)
    x = ops.batchnorm(x, net.bn1, training, relu=True)
    feats = [x]
    x = ops.maxpool(x, 3, 2, 1)
    for li in range(1, 5):
        for blk in getattr(net, "layer%d" % li):
            res = x
            if blk.downsample is not None:
                sc = ops.conv2d(x, blk.downsample[0].weight, stride=blk.stride, bn_next=training)
                res = ops.batchnorm(sc, blk.downsample[1], training)
            y = ops.conv2d(x, blk.conv1.weight, stride=blk.stride, pad=1, bn_next=training)
            y = ops.batchnorm(y, blk.bn1, training, relu=True)
            y = ops.conv2d(y, blk.conv2.weight, pad=1, bn_next=training)
            x = ops.batchnorm(y, blk.bn2, training, relu=True, residual=res)
        feats.append(x)
    return feats


class _ResnetTrunk(nn.Module):
    def __init__(self, num_layers, in_ch):
        super().__init__()
        if num_layers != 18:
            raise ValueError("{} is not a supported number of resnet layers (B200 path implements ResNet-18, "
                             "the only depth the reference configs use)".format(num_layers))
        self.num_ch_enc = np.array([64, 64, 128, 256, 512])
        self.encoder = ResNet18P(in_ch)

    def forward(self, input_image, out_hw=None):
        x = ops.image_prep(input_image, out_hw)
        self.features = resnet18_forward(self.encoder, x, self.training)
        return self.features


class DepthEncoder(_ResnetTrunk):
    def __init__(self, num_layers, pretrained_path=None):
        super().__init__(num_layers, 3)
        _load_pretrained(self.encoder, pretrained_path)


class PoseEncoder(_ResnetTrunk):
    def __init__(self, num_layers, pretrained_path=None, num_input_images=2):
        super().__init__(num_layers, 3 * num_input_images)
        _load_pretrained(self.encoder, pretrained_path, num_input_images)


class ResnetEncoder(_ResnetTrunk):
    def __init__(self, num_layers, pretrained, num_input_images=1):
        # ``pretrained=True`` downloads ImageNet weights in the reference (ResnetEncoder.py:61-66); there is no
        # network here, so weights stay randomly initialised until a checkpoint is loaded.
        super().__init__(num_layers, 3 * num_input_images)


class _CRP(nn.Module):
    def __init__(self, planes, stages=4):
        super().__init__()
        for i in range(stages):
            setattr(self, "%d_pointwise" % (i + 1), Wrap(ConvP(planes, planes, 1, bias=False)))
        self.n_stages = stages


class DepthDecoder(nn.Module):
    def __init__(self, num_ch_enc):
        super().__init__()
        bott = 256
        # attribute order = the reference's registration order (depth_decoder.py:15-41): it fixes the
        # parameter order, hence the optimizer-state layout of interchangeable checkpoints
        self.reduce4 = Wrap(ConvP(int(num_ch_enc[4]), 512, 1, bias=False))
        for lvl in (3, 2, 1):
            setattr(self, "reduce%d" % lvl, Wrap(ConvP(int(num_ch_enc[lvl]), bott, 1, bias=False)))
        self.iconv4 = Wrap(ConvP(512, bott, 3))
        for lvl in (3, 2, 1):
            setattr(self, "iconv%d" % lvl, Wrap(ConvP(2 * bott + 1, bott, 3)))
        for lvl in (4, 3, 2, 1):
            setattr(self, "crp%d" % lvl, nn.ModuleList([_CRP(bott)]))
        for lvl in (4, 3, 2, 1):
            setattr(self, "merge%d" % lvl, Wrap(ConvP(bott, bott, 3)))
        for lvl in (4, 3, 2, 1):
            setattr(self, "disp%d" % lvl, nn.ModuleList([Wrap(ConvP(bott, 1, 3))]))
        self.drop_p = 0.5
        self.drop_masks = None  # tests may inject (mask_l4, mask_l3)

    @staticmethod
    def _crp(crp, x):
        top = x
        terms = [x]
        for i in range(crp.n_stages):
            top = ops.maxpool(top, 5, 1, 2)
            top = ops.conv2d(top, getattr(crp, "%d_pointwise" % (i + 1)).conv.weight)
            terms.append(top)
        # x = top_i + x after every stage (layers.py:197): (((x + t1) + t2) + t3) + t4, one pass instead of four
        return ops.sum_n(terms)

    def forward(self, input_features, frame_id=0):
        l0, l1, l2, l3, l4 = input_features
        m4, m3 = self.drop_masks if self.drop_masks is not None else (None, None)
        step = getattr(self, "step_counter", None)   # device step counter (set by Baseline.forward): fresh masks under graph replay
        l4 = ops.dropout(l4, self.drop_p, self.training, m4, step=step)
        l3 = ops.dropout(l3, self.drop_p, self.training, m3, step=step)
        self.outputs = {}
        skips = {3: l3, 2: l2, 1: l1}
        x = ops.conv2d(l4, self.reduce4.conv.weight)
        prev = disp = None
        for lvl in (4, 3, 2, 1):
            iconv = getattr(self, "iconv%d" % lvl).conv
            if lvl == 4:
                x = ops.conv2d(x, iconv.weight, iconv.bias, pad=1, reflect=True, act="leaky")
            else:
                red = ops.conv2d(skips[lvl], getattr(self, "reduce%d" % lvl).conv.weight)
                x = ops.conv2d([(red, False), (prev, True), (disp, False)], iconv.weight, iconv.bias,
                               pad=1, reflect=True, act="leaky")
            x = self._crp(getattr(self, "crp%d" % lvl)[0], x)
            merge = getattr(self, "merge%d" % lvl).conv
            prev = ops.conv2d(x, merge.weight, merge.bias, pad=1, reflect=True, act="leaky")
            dconv = getattr(self, "disp%d" % lvl)[0].conv
            disp = ops.conv2d([(prev, True)], dconv.weight, dconv.bias, pad=1, reflect=True, act="sigmoid")
            self.outputs[("disp", frame_id, lvl - 1)] = disp
        return self.outputs


class PoseDecoder(nn.Module):
    def __init__(self, num_ch_enc, stride=1):
        super().__init__()
        self.reduce = ConvP(int(num_ch_enc[-1]), 256, 1)
        self.conv1 = ConvP(256, 256, 3)
        self.conv2 = ConvP(256, 256, 3)
        self.conv3 = ConvP(256, 6, 1)
        self.stride = stride

    def features(self, input_features):
        f = input_features[-1]
        x = ops.conv2d(f, self.reduce.weight, self.reduce.bias, act="relu")
        x = ops.conv2d(x, self.conv1.weight, self.conv1.bias, stride=self.stride, pad=1, act="relu")
        x = ops.conv2d(x, self.conv2.weight, self.conv2.bias, stride=self.stride, pad=1, act="relu")
        return ops.conv2d(x, self.conv3.weight, self.conv3.bias)

    def forward(self, input_features):
        out = self.features(input_features)
        out = 0.01 * out.mean(3).mean(2).view(-1, 1, 1, 6)
        return out[..., :3], out[..., 3:]


class Encoder(nn.Module):
    """Layout encoder: ResNet-18 -> reflect 3x3 (512->128) -> pool -> reflect 3x3 -> pool."""

    def __init__(self, num_layers, pretrained=True):
        super().__init__()
        self.resnet_encoder = ResnetEncoder(num_layers, pretrained)
        self.conv1 = Wrap(ConvP(512, 128, 3))
        self.conv2 = Wrap(ConvP(128, 128, 3))

    def forward(self, x, out_hw=None):
        f4 = self.resnet_encoder(x, out_hw)[-1]
        x = ops.maxpool(ops.conv2d(f4, self.conv1.conv.weight, self.conv1.conv.bias, pad=1, reflect=True), 2, 2, 0)
        return ops.maxpool(ops.conv2d(x, self.conv2.conv.weight, self.conv2.conv.bias, pad=1, reflect=True), 2, 2, 0)


class Decoder(nn.Module):
    """Layout decoder; ``decoder`` is the reference's 26-entry ModuleList (layout_model.py:146-158)."""

    def __init__(self, num_ch_enc, num_class=2, type=""):
        super().__init__()
        self.num_output_channels = num_class
        self.num_ch_enc = num_ch_enc
        self.num_ch_dec = np.array([16, 32, 64, 128, 256])
        mods = []
        for i in range(4, -1, -1):
            cin = 128 if i == 4 else int(self.num_ch_dec[i + 1])
            cout = int(self.num_ch_dec[i])
            mods += [ConvP(cin, cout, 3), BNP(cout), nn.Identity(), ConvP(cout, cout, 3), BNP(cout)]
        mods.append(Wrap(ConvP(int(self.num_ch_dec[0]), num_class, 3)))
        self.decoder = nn.ModuleList(mods)

    def forward(self, x, is_training=True):
        d = self.decoder
        for lvl in range(5):
            k = 5 * lvl
            x = ops.conv2d(x, d[k].weight, d[k].bias, pad=1, bn_next=self.training)
            x = ops.batchnorm(x, d[k + 1], self.training, relu=True)
            x = ops.conv2d([(x, True)], d[k + 3].weight, d[k + 3].bias, pad=1, bn_next=self.training)
            x = ops.batchnorm(x, d[k + 4], self.training)
        x = ops.conv2d(x, d[25].conv.weight, d[25].conv.bias, pad=1, reflect=True)
        return x if is_training else torch.softmax(x, 1)


class _TransformModule(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim
        self.fc_transform = nn.ModuleList([LinearP(dim * dim, dim * dim), nn.Identity(), LinearP(dim * dim, dim * dim), nn.Identity()])

    def forward(self, x):
        return ops.cvp_mlp(x, self.fc_transform[0], self.fc_transform[2])


class CycledViewProjection(nn.Module):
    def __init__(self, in_dim):
        super().__init__()
        self.transform_module = _TransformModule(in_dim)
        self.retransform_module = _TransformModule(in_dim)

    def forward(self, x):
        t = self.transform_module(x)
        return t, self.retransform_module(t)


class CrossViewTransformer(nn.Module):
    def __init__(self, in_dim):
        super().__init__()
        self.query_conv = ConvP(in_dim, in_dim // 8, 1)
        self.key_conv = ConvP(in_dim, in_dim // 8, 1)
        self.value_conv = ConvP(in_dim, in_dim, 1)
        self.f_conv = ConvP(in_dim * 2, in_dim, 3)
        self.res_conv = ConvP(in_dim, in_dim // 8, 1)  # never used by the reference forward; kept for the state_dict
        self.query_conv_depth = ConvP(in_dim, in_dim // 8, 1)
        self.key_conv_depth = ConvP(in_dim, in_dim // 8, 1)
        self.value_conv_depth = ConvP(in_dim, in_dim, 1)
        self.conv1 = Wrap(ConvP(512, 128, 3))
        self.conv2 = Wrap(ConvP(128, 128, 3))

    def forward(self, front_x, cross_x, front_x_hat, depth_feature):
        d = ops.maxpool(ops.conv2d(depth_feature, self.conv1.conv.weight, self.conv1.conv.bias, pad=1, reflect=True), 2, 2, 0)
        d = ops.maxpool(ops.conv2d(d, self.conv2.conv.weight, self.conv2.conv.bias, pad=1, reflect=True), 2, 2, 0)
        return ops.cct_attention(front_x, cross_x, front_x_hat, d, self)
