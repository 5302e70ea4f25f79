"""Micro-benchmark of the fused photometric-loss kernels (forward and backward, four scales) on cuda:0.

Prints one JSON line: ms per batch for the four forward launches and the four backward launches and the
achieved algorithmic HBM bandwidth (SURVEY.md §8d bytes) against MEASURED_PEAKS.json.  Inputs are larger
than L2 in aggregate only at B>=8, so an L2 flush (256 MiB write) runs between timed iterations."""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jperceiver_b200 import _lib, functional as JF  # noqa: E402


def make_case(B, H, W, F, dev, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    base = torch.rand(B, 3, H // 16 + 2, W // 16 + 2, generator=g)
    up = torch.nn.functional.interpolate(base, (H, W), mode="bicubic", align_corners=False).clamp(0, 1)
    target = (0.85 * up + 0.15 * torch.rand(B, 3, H, W, generator=g)).clamp(0, 1).to(dev)
    sources = [(0.85 * torch.roll(up, (2 * f, 5 * f), (2, 3)) + 0.15 * torch.rand(B, 3, H, W, generator=g)).clamp(0, 1).to(dev)
               for f in (-1, 1)[:F]]
    disps = [(0.05 + 0.9 * torch.rand(B, 1, H >> (s + 1), W >> (s + 1), generator=g)).to(dev) for s in range(4)]
    K = torch.tensor([[.58 * W, 0, .5 * W, 0], [0, 1.92 * H, .5 * H, 0], [0, 0, 1, 0], [0, 0, 0, 1]]).repeat(B, 1, 1)
    invK = torch.linalg.pinv(K)
    Ts = []
    for f in range(F):
        T = torch.eye(4).repeat(B, 1, 1)
        T[:, :3, 3] = torch.tensor([0.02, -0.01, 0.1 * (1 if f else -1)])
        Ts.append(T.to(dev))
    return target, sources, disps, K.to(dev), invK.to(dev), Ts


def run(B=4, H=320, W=1024, F=2, iters=20, warmup=3, variant=None, bwd_variant=None, debug_outputs=False, ident_once=True):
    dev = torch.device("cuda:0")
    if variant is not None:
        _lib.check(_lib.lib().jpb_photometric_set_variant(variant), "jpb_photometric_set_variant")
    if bwd_variant is not None:
        _lib.check(_lib.lib().jpb_photometric_set_bwd_variant(bwd_variant), "jpb_photometric_set_bwd_variant")
    target, sources, disps, K, invK, Ts = make_case(B, H, W, F, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    disps = [d.requires_grad_(True) for d in disps]
    Ts = [T.requires_grad_(True) for T in Ts]
    ev = lambda: torch.cuda.Event(enable_timing=True)
    fwd_ms, bwd_ms = [], []
    for it in range(warmup + iters):
        flush.zero_()
        e0, e1, e2 = ev(), ev(), ev()
        cache = {} if ident_once else None
        torch.cuda._sleep(1_500_000)     # ~0.75 ms of device spin: the four launches queue up behind it (device time, not launch latency)
        e0.record()
        losses = [JF.photometric_loss(disps[s], target, sources, Ts, K, invK, num_scales=4, seed=it, stream=s, debug_outputs=debug_outputs, ident_cache=cache)[0]
                  for s in range(4)]
        e1.record()
        tot = losses[0] + losses[1] + losses[2] + losses[3]
        flush.zero_()
        e1b = ev()
        torch.cuda._sleep(1_500_000)
        e1b.record()
        tot.backward()
        e2.record()
        torch.cuda.synchronize()
        if it >= warmup:
            fwd_ms.append(e0.elapsed_time(e1))
            bwd_ms.append(e1b.elapsed_time(e2))
        for d in disps:
            d.grad = None
        for T in Ts:
            T.grad = None
    fwd = sorted(fwd_ms)[len(fwd_ms) // 2]
    bwd = sorted(bwd_ms)[len(bwd_ms) // 2]
    dbg = (12 * H * W * F + 8 * H * W) if debug_outputs else 0     # ("color", f, s) frames + int64 min_index written per scale
    bytes_fwd = sum(4 * H * W * (3 + 3 * F) + 4 * (H >> (s + 1)) * (W >> (s + 1)) + dbg for s in range(4)) * B
    bytes_bwd = bytes_fwd + sum(4 * (H >> (s + 1)) * (W >> (s + 1)) for s in range(4)) * B
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm = peaks.get("hbm_gbs", 6650.0)
    return {"metric": "fused photometric-loss ms/batch (4 scales)", "fwd_variant": variant or int(os.environ.get("JPB_PHOTO_FWD", 3)), "bwd_variant": bwd_variant or 4,
            "debug_outputs": debug_outputs, "identity_terms": "once per step" if ident_once else "every scale", "B": B, "H": H, "W": W, "F": F,
            "fwd_ms": fwd, "bwd_ms": bwd, "fwd_alg_bytes": bytes_fwd, "bwd_alg_bytes": bytes_bwd,
            "fwd_gbs": bytes_fwd / fwd / 1e6, "bwd_gbs": bytes_bwd / bwd / 1e6, "hbm_peak_gbs": hbm,
            "fwd_frac": bytes_fwd / fwd / 1e6 / hbm, "bwd_frac": bytes_bwd / bwd / 1e6 / hbm,
            "timing": "cuda events around the 4 launches, queued behind a device spin; L2 flushed between iterations",
            "peak_source": "measured" if peaks else "fallback"}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=4)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--variant", type=int, default=None, help="forward schedule: 2 (default) or 3 (packed fp32 pairs)")
    ap.add_argument("--bwd-variant", type=int, default=None, help="backward schedule: 4 (default) or 1 (generic)")
    ap.add_argument("--debug-outputs", action="store_true", help="materialise (\"color\", f, s) and min_index like the drop-in default")
    ap.add_argument("--ident-every-scale", action="store_true", help="evaluate the identity candidates in all four launches (round-1 behaviour)")
    a = ap.parse_args()
    print(json.dumps(run(B=a.B, iters=a.iters, variant=a.variant, bwd_variant=a.bwd_variant, debug_outputs=a.debug_outputs, ident_once=not a.ident_every_scale)))
