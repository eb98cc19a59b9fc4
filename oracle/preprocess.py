"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference's image pre-processing.

  crop_resize_rgba  : torchvision_F.crop on a PIL RGBA image (demo.py:33-41) + PIL `Image.resize((W, H))` (demo.py:45), i.e.
                      Pillow's default BICUBIC on RGBA: Image.resize converts RGBA -> RGBa (premultiplied, Convert.c rgbA2rgba),
                      runs ImagingResample (Resample.c: precompute_coeffs, normalize_coeffs_8bpc, horizontal then vertical 8-bit
                      pass) and converts back (rgba2rgbA).  Third-party arithmetic (Pillow; 12.2.0 in the build container, the
                      algorithm is unchanged since 3.x): pinned against Pillow itself in tests/test_oracle_preprocess.py.
  composite         : torchvision to_tensor + demo.py:46-52.
  erode             : demo.py:70-75 (cv2.erode 3x3, n iterations), pinned against OpenCV.
Plain python / numpy loops, written for clarity, not speed."""
import math

import numpy as np

PB = 22


def _bicubic(x):
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def coeffs(in_size, out_size):
    scale = in_size / out_size
    fs = max(scale, 1.0)
    support = 2.0 * fs
    out = []
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) / fs) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        k = [int(-0.5 + v * (1 << PB)) if v < 0 else int(0.5 + v * (1 << PB)) for v in w]
        out.append((xmin, k))
    return out


def _clip8(v):
    return np.clip(v >> PB, 0, 255)


def crop_resize_rgba(img, top, left, ch, cw, H, W):
    """img uint8 [H0, W0, 4] -> uint8 [H, W, 4]."""
    H0, W0 = img.shape[:2]
    crop = np.zeros((ch, cw, 4), np.uint8)
    y0, y1, x0, x1 = max(top, 0), min(top + ch, H0), max(left, 0), min(left + cw, W0)
    if y1 > y0 and x1 > x0:
        crop[y0 - top:y1 - top, x0 - left:x1 - left] = img[y0:y1, x0:x1]
    if (ch, cw) == (H, W):
        return crop
    a = crop[..., 3].astype(np.int64)
    pre = crop.astype(np.int64)
    for b in range(3):                                   # MULDIV255
        t = pre[..., b] * a + 128
        pre[..., b] = ((t >> 8) + t) >> 8
    cur = pre
    if cw != W:
        tmp = np.zeros((ch, W, 4), np.int64)
        for ox, (xmin, k) in enumerate(coeffs(cw, W)):
            ss = np.full((ch, 4), 1 << (PB - 1), np.int64)
            for x, kv in enumerate(k):
                ss += cur[:, xmin + x, :] * kv
            tmp[:, ox, :] = _clip8(ss)
        cur = tmp
    if ch != H:
        tmp = np.zeros((H, cur.shape[1], 4), np.int64)
        for oy, (ymin, k) in enumerate(coeffs(ch, H)):
            ss = np.full((cur.shape[1], 4), 1 << (PB - 1), np.int64)
            for y, kv in enumerate(k):
                ss += cur[ymin + y, :, :] * kv
            tmp[oy] = _clip8(ss)
        cur = tmp
    out = cur.copy()
    al = cur[..., 3]
    part = (al != 255) & (al != 0)
    for b in range(3):
        q = np.where(part, (255 * cur[..., b]) // np.maximum(al, 1), cur[..., b])
        out[..., b] = np.clip(q, 0, 255)
    return out.astype(np.uint8)


def composite(img_u8, bgcolor):
    """uint8 [H, W, 4] -> (rgb [3,H,W], mask [1,H,W]) float32, demo.py:46-52."""
    t = (img_u8.astype(np.float32) / np.float32(255)).transpose(2, 0, 1)
    rgb, mask = t[:3], t[3:]
    if bgcolor is not None:
        rgb = rgb * mask + np.float32(bgcolor) * (np.float32(1) - mask)
        mask = (mask > 0.5).astype(np.float32)
    return rgb.astype(np.float32), mask.astype(np.float32)


def erode(mask, iterations):
    """[H, W] -> minimum over the (2 it + 1)^2 window clipped to the image."""
    m = np.asarray(mask, np.float32)
    H, W = m.shape
    out = np.empty_like(m)
    r = iterations
    for y in range(H):
        for x in range(W):
            out[y, x] = m[max(y - r, 0):y + r + 1, max(x - r, 0):x + r + 1].min()
    return out
