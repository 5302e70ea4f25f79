"""The oracle port against full forward+backward runs of the REAL reference (tests/golden/e2e_*.npz,
written by oracle/make_golden.py): every loss scalar, strided output samples, gradient checksums,
BatchNorm buffer updates (incl. the reference's double road-head evaluation)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import port as O
from oracle.make_golden import GRAD_KEYS, sample
from oracle.ref_loader import default_options


def _run(typ):
    split = "argo" if typ.startswith("Argo") else "odometry"
    opt = default_options(frame_ids=[0, -1, 1], height=1024, width=1024, type=typ, split=split)
    shapes = json.load(open(os.path.join(GOLDEN, "state_dict_shapes.json")))
    tmpl = {k: torch.empty(s) for k, s in shapes.items()}
    P = O.synth_params(tmpl, seed=3)
    for k, v in P.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    hw = (2056, 2464) if split == "argo" else (375, 1242)
    inp = O.synth_inputs(opt, 1, seed=1, hw_full=hw)
    outs, losses = O.forward(P, opt, inp, training=True, drop_p=0.0, noise=None)
    total = O.total_loss(losses)
    total.backward()
    return P, outs, losses, total


@pytest.mark.parametrize("typ", ["Argo_both", "static", "static_raw", "dynamic", "Argo_static", "Argo_dynamic"])
def test_port_matches_reference_run(typ):
    torch.set_num_threads(os.cpu_count() or 1)
    gold = np.load(os.path.join(GOLDEN, f"e2e_{typ}_1024.npz"))
    P, outs, losses, total = _run(typ)
    assert {"loss/" + str(k) for k in losses} == {k for k in gold.files if k.startswith("loss/")}
    for k, v in losses.items():
        ref = float(gold["loss/" + str(k)])
        tol = 5e-5 if (isinstance(k, tuple) and k[0] == "scale_loss" and typ != "Argo_both") else 2e-6
        assert abs(float(v) - ref) <= tol * max(1.0, abs(ref)), (k, float(v), ref)
    assert abs(total.item() - float(gold["total_loss"])) <= 1e-5 * abs(float(gold["total_loss"]))
    for key in gold.files:
        if not key.startswith("out/") or key.endswith("/sum"):
            continue
        name = eval(key[4:].split("/hist")[0]) if key[4] == "(" else key[4:].split("/hist")[0]
        if key.endswith("/hist"):
            assert np.array_equal(np.bincount(outs[name].flatten().numpy(), minlength=4), gold[key]), key
        else:
            ref = gold[key]
            got = sample(outs[name])
            assert np.max(np.abs(got - ref)) <= 1e-5 * max(1.0, np.max(np.abs(ref))), key
    gtol = 2e-4 if typ == "Argo_both" else 2e-3  # static*: the pinned ">= 1-2^-20" mask test moves scale_loss by ~1e-5
    if typ == "Argo_static":   # Argoverse geometry puts more of the mask edge inside the frame: scale_loss (the only term that
        gtol = 5e-3            # differs) moves by 7e-4 of its value at s=2, the depth-decoder gradients by 2.7e-3
    for k in GRAD_KEYS:
        g = P[k].grad
        ref = gold["grad/" + k]
        if g is None:
            assert ref[1] == 0.0, k
            continue
        got = np.array([g.double().sum().item(), g.double().abs().sum().item(), g.double().norm().item()])
        assert abs(got[2] - ref[2]) <= gtol * max(ref[2], 1e-12), (k, got, ref)
        assert np.max(np.abs(sample(g, 64) - gold["gradv/" + k])) <= gtol * max(np.max(np.abs(gold["gradv/" + k])), 1e-12), k
    nograd = sorted(k for k, p in P.items() if p.requires_grad and p.grad is None)
    assert nograd == sorted(gold["nograd"].tolist())
    for key in gold.files:
        if key.startswith("buf/"):
            got = P[key[4:]].detach().double().numpy()
            assert np.max(np.abs(got - gold[key])) <= 1e-5 * max(1.0, np.max(np.abs(gold[key]))), key
