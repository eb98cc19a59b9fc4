"""Image pre-processing of demo.py:20-75 and data/synthetic.py:178-210 with the pixel work on the device.

Same function names and argument meaning as the reference; images are uint8 RGBA tensors [H, W, 4] on the GPU instead of PIL
images (a PIL image / numpy array is accepted and uploaded).  The crop + resize reproduces Pillow's default-BICUBIC `Image.resize`
of an RGBA image byte for byte (csrc/preprocess.cu); the coefficient tables are computed here exactly as Pillow's
`precompute_coeffs` / `normalize_coeffs_8bpc` do (src/libImaging/Resample.c) -- double arithmetic on the host, a few KB.
"""
import math

import numpy as np
import torch

from .. import ops

PRECISION_BITS = 32 - 8 - 2      # Pillow Resample.c


def get_1d_bounds(arr):
    """demo.py:20-22."""
    nz = np.flatnonzero(arr)
    return nz[0], nz[-1]


def get_bbox_from_mask(mask, thr, min_pixels=None):
    """demo.py:24-31 (asserts a non-empty mask); with `min_pixels` the variant of data/synthetic.py:183-191 (None if the mask has
    at most that many pixels).  mask: numpy [H, W]."""
    masks_for_box = (np.asarray(mask) > thr).astype(np.float32)
    if min_pixels is not None:
        if masks_for_box.sum() <= min_pixels:
            return None
    else:
        assert masks_for_box.sum() > 0, "Empty mask!"
    x0, x1 = get_1d_bounds(masks_for_box.sum(axis=-2))
    y0, y1 = get_1d_bounds(masks_for_box.sum(axis=-1))
    return x0, y0, x1, y1


def _bicubic(x, a=-0.5):
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resize_coeffs(in_size, out_size, support=2.0):
    """Pillow precompute_coeffs + normalize_coeffs_8bpc for the BICUBIC filter over the whole input range.
    -> (bounds int32 [out, 2] = (first source index, taps), kk int32 [out, ksize])."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    sup = support * filterscale
    ksize = int(math.ceil(sup)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - sup + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + sup + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


_COEFF_CACHE = {}


def _coeffs_on(device, in_size, out_size):
    key = (str(device), in_size, out_size)
    if key not in _COEFF_CACHE:
        b, k = resize_coeffs(in_size, out_size)
        _COEFF_CACHE[key] = (torch.from_numpy(b).to(device), torch.from_numpy(k).to(device))
    return _COEFF_CACHE[key]


def _as_rgba_u8(image, device):
    if isinstance(image, torch.Tensor):
        t = image
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(image)))      # PIL RGBA image or [H, W, 4] array
    assert t.dtype == torch.uint8 and t.dim() == 3 and t.shape[2] == 4, "expected a uint8 RGBA image [H, W, 4]"
    return t.to(device).contiguous()


def crop_box(bbox, crop_ratio=1.):
    """The integer crop window of square_crop (demo.py:33-41, data/synthetic.py:201-210): (top, left, height, width)."""
    x1, y1, x2, y2 = bbox
    h, w = y2 - y1, x2 - x1
    yc, xc = (y1 + y2) / 2, (x1 + x2) / 2
    S = max(h, w) * 1.2
    scale = S * crop_ratio
    return int(yc - scale / 2), int(xc - scale / 2), int(scale), int(scale)


def square_crop_resize(image, bbox, H, W, crop_ratio=1., device="cuda"):
    """square_crop followed by `image.resize((W, H))` when the crop is not already H x W -> uint8 RGBA [H, W, 4] on the device."""
    img = _as_rgba_u8(image, device)
    top, left, ch, cw = crop_box(bbox, crop_ratio)
    if ch == H and cw == W:                                   # the reference skips the resize
        out = torch.zeros(H, W, 4, dtype=torch.uint8, device=img.device)
        y0, y1, x0, x1 = max(top, 0), min(top + ch, img.shape[0]), max(left, 0), min(left + cw, img.shape[1])
        if y1 > y0 and x1 > x0:
            out[y0 - top:y1 - top, x0 - left:x1 - left] = img[y0:y1, x0:x1]
        return out
    xb, xk = _coeffs_on(img.device, cw, W)
    yb, yk = _coeffs_on(img.device, ch, H)
    return ops.rgba_crop_resize(img, left, top, cw, ch, H, W, xb, xk, yb, yk)


def preprocess_image(opt, image, bbox):
    """demo.py:43-53: crop, resize to opt.W x opt.H, to_tensor, composite on opt.data.bgcolor -> (rgb [3,H,W], mask [1,H,W])."""
    img = square_crop_resize(image, bbox, opt.H, opt.W, device=opt.device)
    bg = opt.data.bgcolor
    return ops.rgba_composite(img, bg)


def get_image(opt, image_name, mask_name):
    """demo.py:55-68: read <datadir>/images/<image_name> and <datadir>/masks/<mask_name> (PIL, host), binarise the mask at 127 for
    the bounding box, and run preprocess_image on the device."""
    import os

    from PIL import Image
    image = Image.open(os.path.join(opt.datadir, "images", image_name)).convert("RGB")
    mask = Image.open(os.path.join(opt.datadir, "masks", mask_name)).convert("L")
    mask_np = np.array(mask)
    mask_np[mask_np <= 127] = 0
    mask_np[mask_np >= 127] = 1.0
    rgba = np.dstack([np.asarray(image), np.asarray(mask)])
    bbox = get_bbox_from_mask(mask_np, 0.5)
    return preprocess_image(opt, rgba, bbox)


def erode_mask(mask, iterations=5):
    """demo.py:70-75 (mask -> uint8, cv2.erode with a 3 x 3 kernel `iterations` times, back to float) on the device:
    mask [1, H, W] -> [1, H, W] float."""
    m = mask.to(torch.float32).floor()                 # the reference's astype(np.uint8) on a 0/1 (or 0..255) mask
    H, W = m.shape[-2:]
    return ops.erode_square(m.contiguous().view(1, H, W), iterations).view(1, H, W)
