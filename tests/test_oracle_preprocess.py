"""Pins the pre-processing restatement (oracle/preprocess.py) to the libraries the reference calls -- Pillow's RGBA bicubic
resize, torchvision's crop / to_tensor and OpenCV's erode -- and the product's host coefficient tables to the oracle's."""
import numpy as np
import pytest
import torch

from oracle import preprocess as OP


def _rgba(seed, H, W):
    rs = np.random.RandomState(seed)
    img = rs.randint(0, 256, (H, W, 4)).astype(np.uint8)
    yy, xx = np.mgrid[:H, :W]
    soft = np.clip(255 - 3.0 * (np.hypot(yy - H / 2, xx - W / 2) - min(H, W) / 3), 0, 255)
    img[..., 3] = soft.astype(np.uint8)                   # a disc with a soft edge: exercises the premultiplied path
    return img


@pytest.mark.parametrize("H0,W0,box,out", [(300, 400, (20, 50, 260, 260), 224), (90, 70, (-15, -20, 110, 110), 224),
                                           (640, 640, (100, 80, 500, 500), 224), (224, 224, (0, 0, 224, 224), 224),
                                           (50, 60, (10, 10, 31, 31), 64)])
def test_crop_resize_matches_pillow(H0, W0, box, out):
    PIL = pytest.importorskip("PIL.Image")
    tvF = pytest.importorskip("torchvision.transforms.functional")
    img = _rgba(H0 + W0, H0, W0)
    top, left, ch, cw = box
    pil = tvF.crop(PIL.fromarray(img, "RGBA"), top=top, left=left, height=ch, width=cw)
    if pil.size[0] != out or pil.size[1] != out:
        pil = pil.resize((out, out))
    ref = np.asarray(pil)
    got = OP.crop_resize_rgba(img, top, left, ch, cw, out, out)
    assert got.shape == ref.shape
    np.testing.assert_array_equal(got, ref)


def test_composite_matches_torchvision_to_tensor():
    PIL = pytest.importorskip("PIL.Image")
    tvF = pytest.importorskip("torchvision.transforms.functional")
    img = _rgba(3, 64, 48)
    t = tvF.to_tensor(PIL.fromarray(img, "RGBA"))
    rgb, mask = t[:3], t[3:]
    for bg in (None, 1.0, 0.5):
        r, m = OP.composite(img, bg)
        if bg is None:
            r_ref, m_ref = rgb, mask
        else:
            r_ref, m_ref = rgb * mask + bg * (1 - mask), (mask > 0.5).float()
        assert np.array_equal(r, r_ref.numpy()) and np.array_equal(m, m_ref.numpy())


def test_erode_matches_opencv():
    cv2 = pytest.importorskip("cv2")
    rs = np.random.RandomState(0)
    mask = (rs.rand(40, 37) > 0.08).astype(np.uint8)
    for it in (1, 2, 5):
        ref = cv2.erode(mask, np.ones((3, 3), np.uint8), iterations=it)
        np.testing.assert_array_equal(OP.erode(mask, it), ref.astype(np.float32))


def test_host_coefficient_tables_match_the_oracle():
    from zeroshape_b200.data.preprocess import resize_coeffs, crop_box
    for n_in, n_out in ((500, 224), (110, 224), (224, 224), (31, 64), (1000, 7)):
        bounds, kk = resize_coeffs(n_in, n_out)
        for xx, (xmin, k) in enumerate(OP.coeffs(n_in, n_out)):
            assert bounds[xx, 0] == xmin and bounds[xx, 1] == len(k)
            assert kk[xx, :len(k)].tolist() == k and not kk[xx, len(k):].any()
    assert crop_box((10, 20, 110, 90)) == (int(55 - 60.0), int(60 - 60.0), 120, 120)


def test_oracle_reproduces_the_committed_golden():
    """tests/golden/preprocess.npz was produced by Pillow / torchvision / OpenCV (make_golden_preprocess.py); it travels to the GPU box."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden_preprocess import CASES, rgba, digest
    from zeroshape_b200.data.preprocess import crop_box
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "preprocess.npz"))
    for i, (H0, W0, bbox) in enumerate(CASES):
        top, left, ch, cw = crop_box(bbox)
        out = OP.crop_resize_rgba(rgba(10 + i, H0, W0), top, left, ch, cw, 224, 224)
        np.testing.assert_array_equal(out, g[f"rgba224_{i}"])
        r, m = OP.composite(out, 1.0)
        assert np.array_equal(digest(r), g[f"rgb_sha256_{i}"]) and np.array_equal(m.astype(np.uint8), g[f"mask{i}"])
        np.testing.assert_array_equal(OP.erode(m[0], 5).astype(np.uint8), g[f"eroded{i}"])
