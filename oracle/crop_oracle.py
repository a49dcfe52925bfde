"""CPU oracle (test infrastructure only) for the visual half of the reference's data pipeline:

    transforms.RandomResizedCrop(224) / Resize((224, 224)) -> RandomHorizontalFlip -> ToTensor -> Normalize
    (reference dataset/CramedDataset.py:76-89, 96-101; dataset/KSDataset.py:160-173, 183-190)

given the crop box (i, j, h, w) and the flip decision that torchvision drew on the host.  The arithmetic lives
in third-party code that is not under /root/reference: torchvision 0.26 `F.resized_crop` (PIL backend: `img.crop`
then `img.resize((224, 224), BILINEAR)`) and Pillow 12.2 `src/libImaging/Resample.c`
(`precompute_coeffs`, `normalize_coeffs_8bpc`, `ImagingResampleHorizontal_8bpc`, `ImagingResampleVertical_8bpc`),
whose published algorithm is restated here in numpy:

  * per axis: scale = in/out, support = max(scale, 1) (bilinear support 1), window [xmin, xmin+xmax) around
    center = (xx + 0.5) * scale, triangle weights normalised to sum 1 in double precision, then quantised to
    22-bit fixed point with round-half-up;
  * horizontal pass over the rows the vertical pass needs, 8-bit intermediate (rounded, clipped);
  * vertical pass, 8-bit result;  ToTensor = uint8 / 255 in fp32;  Normalize = (t - mean) / std in fp32.

PINNED: tests/test_cpu_crop.py checks this file bit for bit against PIL/torchvision run in the build container
(random images, crop boxes incl. up-scaling and 1-pixel crops) and against tests/golden/crop_golden.npz, which
tests/golden/make_crop_golden.py generated from torchvision itself.  Only tests/, __graft_entry__.smoke() and
bench.py's CPU legs may import this module.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2  # Resample.c: 22-bit fixed-point coefficients for 8-bit channels
MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)  # CramedDataset.py:81
STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)


def precompute_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter and box (0, in_size).
    Returns (bounds [out, 2] int: first input index, count; kk [out, ksize] int32)."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = []
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            if a < 0.0:
                a = -a
            w = 1.0 - a if a < 1.0 else 0.0
            k.append(w)
            ww += w
        for x in range(xmax):
            v = k[x] / ww if ww != 0.0 else k[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _clip8(acc):
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_bilinear_u8(img, out_h, out_w):
    """PIL Image.resize((out_w, out_h), BILINEAR) for an HWC uint8 array: horizontal pass first (when the width
    changes), on the row range the vertical pass reads, then the vertical pass (when the height changes)."""
    in_h, in_w, _ = img.shape
    need_h, need_v = out_w != in_w, out_h != in_h
    bounds_v, kk_v = precompute_coeffs(in_h, out_h)
    src = img
    row0 = 0
    if need_h:
        bounds_h, kk_h = precompute_coeffs(in_w, out_w)
        row0 = int(bounds_v[0, 0]) if need_v else 0
        row1 = int(bounds_v[-1, 0] + bounds_v[-1, 1]) if need_v else in_h
        rows = img[row0:row1].astype(np.int64)
        tmp = np.zeros((row1 - row0, out_w, img.shape[2]), dtype=np.uint8)
        for xx in range(out_w):
            xmin, xmax = bounds_h[xx]
            acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(rows[:, xmin:xmin + xmax, :], kk_h[xx, :xmax], axes=([1], [0]))
            tmp[:, xx, :] = _clip8(acc)
        src = tmp
    if not need_v:
        return src
    out = np.zeros((out_h, src.shape[1], img.shape[2]), dtype=np.uint8)
    s64 = src.astype(np.int64)
    for yy in range(out_h):
        ymin, ymax = bounds_v[yy]
        ymin -= row0
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk_v[yy, :ymax], s64[ymin:ymin + ymax], axes=([0], [0]))
        out[yy] = _clip8(acc)
    return out


def crop_resize_flip_normalize(img, i, j, h, w, flip, size=224):
    """One frame of the reference transform: HWC uint8 -> fp32 [3, size, size] (CramedDataset.py:76-82 with the
    random draws (i, j, h, w, flip) supplied; Resize((224, 224)) of the test split is the full-image box)."""
    crop = np.ascontiguousarray(img[i:i + h, j:j + w])
    out = resize_bilinear_u8(crop, size, size)
    if flip:
        out = out[:, ::-1]
    t = out.astype(np.float32).transpose(2, 0, 1) / np.float32(255.0)  # ToTensor: uint8 -> fp32, true division
    return ((t - MEAN[:, None, None]) / STD[:, None, None]).astype(np.float32)
