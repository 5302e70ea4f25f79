"""Parameter holders.  They carry no arithmetic — ``forward`` of every network in this package is a
sequence of calls into ``jperceiver_b200.netops`` — but they reproduce the reference's ``state_dict``
key layout and shapes exactly (766 entries), which is the checkpoint contract.

Convolution weights keep the logical ``(Cout, Cin, kh, kw)`` shape but are stored channels-last, i.e.
physically ``[Cout][kh][kw][Cin]`` — the K-major operand layout of the implicit-GEMM kernels."""
import math

import torch
import torch.nn as nn


class ConvP(nn.Module):
    def __init__(self, cin, cout, k, bias=True, init="default"):
        super().__init__()
        w = torch.empty(cout, cin, k, k)
        if init == "kaiming_out":  # reference resnet.py:104-109
            nn.init.kaiming_normal_(w, mode="fan_out", nonlinearity="relu")
        else:                       # nn.Conv2d default
            nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        self.weight = nn.Parameter(w.contiguous(memory_format=torch.channels_last))
        if bias:
            bound = 1 / math.sqrt(cin * k * k)
            self.bias = nn.Parameter(torch.empty(cout).uniform_(-bound, bound))
        else:
            self.register_parameter("bias", None)
        self.k = k


class BNP(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class LinearP(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        w = torch.empty(cout, cin)
        nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        self.weight = nn.Parameter(w)
        bound = 1 / math.sqrt(cin)
        self.bias = nn.Parameter(torch.empty(cout).uniform_(-bound, bound))


class Wrap(nn.Module):
    """``name.conv.*`` nesting of the reference's Conv1x1 / Conv3x3 wrappers (layers.py:146-167)."""

    def __init__(self, conv):
        super().__init__()
        self.conv = conv


class Holder(nn.Module):
    def __init__(self, **children):
        super().__init__()
        for k, v in children.items():
            setattr(self, k, v)


class BlockP(nn.Module):
    """ResNet BasicBlock parameters (resnet.py:16-45)."""

    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = ConvP(cin, cout, 3, bias=False, init="kaiming_out")
        self.bn1 = BNP(cout)
        self.conv2 = ConvP(cout, cout, 3, bias=False, init="kaiming_out")
        self.bn2 = BNP(cout)
        self.stride = stride
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.ModuleList([ConvP(cin, cout, 1, bias=False, init="kaiming_out"), BNP(cout)])


class ResNet18P(nn.Module):
    """ResNet-18 trunk parameters incl. the unused ``fc`` (resnet.py:86-109; it stays in the state_dict)."""

    def __init__(self, in_ch=3):
        super().__init__()
        self.conv1 = ConvP(in_ch, 64, 7, bias=False, init="kaiming_out")
        self.bn1 = BNP(64)
        chans = [64, 64, 128, 256, 512]
        for i in range(1, 5):
            stride = 1 if i == 1 else 2
            setattr(self, "layer%d" % i, nn.ModuleList([BlockP(chans[i - 1], chans[i], stride), BlockP(chans[i], chans[i], 1)]))
        self.fc = LinearP(512, 1000)
