"""Input pipeline on the device (csrc/preprocess.cu through zeroshape_b200/data/preprocess.py) against the oracle restatement and
against golden vectors produced with Pillow / torchvision / OpenCV (tests/golden/make_golden_preprocess.py).  Byte / bit exact."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "preprocess.npz")


def _opt(device):
    from zeroshape_b200.utils.util import EasyDict
    return EasyDict(H=224, W=224, device=device, data=dict(bgcolor=1.0))


def test_preprocess_matches_the_reference_libraries_golden(cuda):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden_preprocess import CASES, rgba, digest
    from zeroshape_b200.data import preprocess as PP
    g = np.load(GOLD)
    for i, (H0, W0, bbox) in enumerate(CASES):
        img = rgba(10 + i, H0, W0)
        out = PP.square_crop_resize(torch.from_numpy(img), bbox, 224, 224, device=cuda)
        np.testing.assert_array_equal(out.cpu().numpy(), g[f"rgba224_{i}"])
        rgb, mask = PP.preprocess_image(_opt(cuda), img, bbox)
        assert rgb.shape == (3, 224, 224) and mask.shape == (1, 224, 224) and rgb.dtype == torch.float32
        assert np.array_equal(digest(rgb.cpu().numpy()), g[f"rgb_sha256_{i}"])            # bit-exact with torchvision + demo.py:46-52
        np.testing.assert_array_equal(mask.cpu().numpy().astype(np.uint8), g[f"mask{i}"])
        np.testing.assert_array_equal(PP.erode_mask(mask).cpu().numpy()[0].astype(np.uint8), g[f"eroded{i}"])


@pytest.mark.parametrize("H0,W0,box,out", [(300, 400, (20, 50, 260, 260), 224), (90, 70, (-15, -20, 110, 110), 224),
                                           (1024, 768, (100, 80, 700, 700), 224), (224, 224, (0, 0, 224, 224), 224),
                                           (50, 60, (10, 10, 31, 31), 64), (40, 40, (-100, -100, 30, 30), 16)])
def test_crop_resize_matches_the_oracle(cuda, H0, W0, box, out):
    from oracle import preprocess as OP
    from zeroshape_b200 import ops
    from zeroshape_b200.data.preprocess import resize_coeffs
    rs = np.random.RandomState(H0 * 7 + W0)
    img = rs.randint(0, 256, (H0, W0, 4)).astype(np.uint8)
    top, left, ch, cw = box
    ref = OP.crop_resize_rgba(img, top, left, ch, cw, out, out)
    if (ch, cw) == (out, out):
        pytest.skip("no resize: host slice path, covered by the golden test")
    xb, xk = (torch.from_numpy(a).to(cuda) for a in resize_coeffs(cw, out))
    yb, yk = (torch.from_numpy(a).to(cuda) for a in resize_coeffs(ch, out))
    got = ops.rgba_crop_resize(torch.from_numpy(img).to(cuda), left, top, cw, ch, out, out, xb, xk, yb, yk)
    np.testing.assert_array_equal(got.cpu().numpy(), ref)


def test_composite_and_erode_match_the_oracle(cuda):
    from oracle import preprocess as OP
    from zeroshape_b200 import ops
    rs = np.random.RandomState(4)
    img = rs.randint(0, 256, (37, 53, 4)).astype(np.uint8)
    for bg in (None, 1.0, 0.25):
        rgb, mask = ops.rgba_composite(torch.from_numpy(img).to(cuda), bg)
        r, m = OP.composite(img, bg)
        assert np.array_equal(rgb.cpu().numpy(), r) and np.array_equal(mask.cpu().numpy(), m)
    mk = (rs.rand(2, 33, 41) > 0.1).astype(np.float32)
    for it in (0, 1, 5):
        out = ops.erode_square(torch.from_numpy(mk).to(cuda), it).cpu().numpy()
        for b in range(2):
            np.testing.assert_array_equal(out[b], OP.erode(mk[b], it))


def test_synthetic_dataset_sample_and_demo_get_image(cuda, tmp_path):
    """data/synthetic.py __getitem__ and demo.py get_image from FILES: the device path against PIL / torchvision run on the same
    files (the reference's own calls), incl. the camera pose composition."""
    PIL = pytest.importorskip("PIL.Image")
    tvF = pytest.importorskip("torchvision.transforms.functional")
    from test_data_formats_cpu import make_tree
    from zeroshape_b200.data.synthetic import Dataset
    from zeroshape_b200.data import preprocess as PP
    from zeroshape_b200.utils.util import EasyDict
    base = make_tree(str(tmp_path), H=256, W=256)                    # 256 -> the resize path of preprocess_image runs
    opt = EasyDict(H=224, W=224, device=cuda, data=dict(synthetic=dict(subset="objaverse_LVIS", percentage=1), bgcolor=1.0),
                   training=dict(n_sdf_points=512))
    for o in range(3):                                                # depth maps must be H x H (asserted by the reader)
        name = f"chair_obj{o}_{o:03d}"
        np.save(f"{base}/depth/chair/{name}.npy", np.load(f"{base}/depth/chair/{name}.npy")[:224, :224])
    ds = Dataset(opt, "train", path=str(tmp_path), device=cuda)
    s = ds[1]
    assert set(s) == {"idx", "category_label", "pose_gt", "intr", "rgb_input_map", "mask_input_map", "depth_input_map", "dpc",
                      "gt_sample_points", "gt_sample_sdf"}
    img = PIL.open(f"{base}/images_processed/chair/chair_obj1_001.png").convert("RGB")
    msk = PIL.open(f"{base}/masks/chair/chair_obj1_001.png").convert("L")
    rgba = PIL.merge("RGBA", (*img.split(), msk))
    ref = tvF.to_tensor(rgba.resize((224, 224)))[:3]
    assert torch.equal(s["rgb_input_map"].cpu(), ref)
    Rt = np.load(f"{base}/camera_data/extr/chair/chair_obj1_001.npy")
    assert torch.allclose(s["pose_gt"], torch.from_numpy(Rt[:3]).float(), atol=1e-6)
    assert s["gt_sample_points"].shape == (512, 3) and s["depth_input_map"].shape == (1, 224, 224)
    # demo.py get_image on <datadir>/images, <datadir>/masks
    os.makedirs(tmp_path / "images"); os.makedirs(tmp_path / "masks")
    img.save(str(tmp_path / "images" / "a.png")); msk.save(str(tmp_path / "masks" / "a.png"))
    opt.datadir = str(tmp_path)
    rgb, mask = PP.get_image(opt, "a.png", "a.png")
    mask_np = np.array(msk)
    mask_np[mask_np <= 127] = 0
    mask_np[mask_np >= 127] = 1.0
    x0, y0, x1, y1 = PP.get_bbox_from_mask(mask_np, 0.5)
    h, w = y1 - y0, x1 - x0
    yc, xc = (y0 + y1) / 2, (x0 + x1) / 2
    scale = max(h, w) * 1.2
    pil = tvF.crop(rgba, top=int(yc - scale / 2), left=int(xc - scale / 2), height=int(scale), width=int(scale)).resize((224, 224))
    t = tvF.to_tensor(pil)
    assert torch.equal(rgb.cpu(), t[:3] * t[3:] + 1.0 * (1 - t[3:])) and torch.equal(mask.cpu(), (t[3:] > 0.5).float())


def test_mesh_export_writes_a_binary_ply(cuda, tmp_path):
    from zeroshape_b200 import ops
    from zeroshape_b200.data.formats import read_ply
    from zeroshape_b200.utils.eval_3D import Mesh
    n = 24
    g = torch.linspace(-1, 1, n, device=cuda)
    X, Y, Z = torch.meshgrid(g, g, g, indexing="ij")
    vol = (0.6 - torch.sqrt(X * X + Y * Y + Z * Z)).contiguous()
    v, f = ops.marching_cubes(vol, 0.0)
    mesh = Mesh(v, f, 3.0 / n, -1.5)
    mesh.export(str(tmp_path / "m.ply"))
    v2, f2 = read_ply(str(tmp_path / "m.ply"))
    np.testing.assert_array_equal(f2, mesh.faces.astype(np.int32))
    np.testing.assert_array_equal(v2, mesh.vertices.astype(np.float32))
    assert f2.max() < len(v2) and len(f2) > 100
