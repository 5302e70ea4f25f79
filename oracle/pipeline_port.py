"""CPU restatement of the reference's per-sample image preprocessing — TEST INFRASTRUCTURE (see
``oracle/__init__.py``).

Follows /root/reference/mono/datasets/mono_dataset.py:126-171 (``preprocess``), :202-203 and :337-343 (random decisions,
``ColorJitter``), :417-431 (``process_topview`` / ``process_topview_both``) with the very libraries the reference calls: Pillow's
``Image.resize`` / ``ImageEnhance`` and torchvision's ``ColorJitter`` / ``ToTensor``.  Pinned against the reference's own
``MonoDataset.preprocess`` executed under an import shim (``oracle/make_golden_pipeline.py`` -> ``tests/golden/kat_pipeline.json``).
``Image.ANTIALIAS`` (removed in Pillow 10) is ``Image.LANCZOS``."""
from __future__ import annotations

import numpy as np
import torch
from PIL import Image
from torchvision import transforms


def preprocess_colour(frames, height, width, full_res=(375, 1242), color_aug=None, flip=False):
    """``frames``: ``{frame_id: PIL RGB image}`` as loaded from disk.  Returns the float tensors ``("color", f, -1)``,
    ``("color", f, 0)``, ``("color_aug", f, 0)`` (mono_dataset.py:133-158; ``get_color`` flips before anything else).
    ``color_aug``: a callable on PIL images (a ``transforms.ColorJitter`` module, called once per frame in frame order — each
    call draws fresh parameters) or None for the identity."""
    to_tensor = transforms.ToTensor()
    lanczos = Image.LANCZOS
    out, scale0 = {}, {}
    for f, im in frames.items():
        if flip:
            im = im.transpose(Image.FLIP_LEFT_RIGHT)
        full = im.resize((full_res[1], full_res[0]), lanczos)          # transforms.Resize((375, 1242), ANTIALIAS)
        scale0[f] = full.resize((width, height), lanczos)              # resize of the ALREADY resized frame (:140-144)
        out[("color", f, -1)] = to_tensor(full)
    for f, im in scale0.items():
        out[("color", f, 0)] = to_tensor(im)
        out[("color_aug", f, 0)] = to_tensor(color_aug(im) if color_aug is not None else im)
    return out


def process_topview_both(label, size, flip=False):
    """mono_dataset.py:425-431 on a PIL ``L`` image -> float64 size×size array of {0, 1}."""
    if flip:
        label = label.transpose(Image.FLIP_LEFT_RIGHT)
    t = np.array(label.resize((size, size), Image.NEAREST))
    out = np.zeros(t.shape)
    out[t == 255] = 1
    return out


def process_topview(label, size, flip=False):
    """mono_dataset.py:417-424: ``convert("1")`` (Floyd-Steinberg dither; the identity on two-level images) first."""
    if flip:
        label = label.transpose(Image.FLIP_LEFT_RIGHT)
    t = np.array(label.convert("1").resize((size, size), Image.NEAREST).convert("L"))
    out = np.zeros(t.shape)
    out[t == 255] = 1
    return out


def synth_frame(seed, h, w):
    """Deterministic textured RGB frame (uint8 HWC)."""
    r = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(np.sin(xx / 9.0 + seed) + np.cos(yy / 5.0)) * 60 + 128, (xx * 3 + yy * 2 + 11 * seed) % 256, (xx * yy // 7) % 256], -1)
    return (0.7 * base + 0.3 * r.randint(0, 256, (h, w, 3))).clip(0, 255).astype(np.uint8)


def synth_label(seed, h, w):
    r = np.random.RandomState(100 + seed)
    a = np.zeros((h, w), np.uint8)
    a[h // 5: h // 2 + seed, w // 7: w // 2] = 255
    a[r.rand(h, w) > 0.98] = 255
    return a


def digest(t):
    import hashlib
    t = t if isinstance(t, np.ndarray) else t.detach().cpu().contiguous().numpy()
    return hashlib.sha256(np.ascontiguousarray(t).tobytes()).hexdigest()[:24]
