"""On-disk formats either side of the hot path (zeroshape_b200/data): PLY writer / reader, the checkpoint dict of
utils/util.py:227-275 and the file layout of data/synthetic.py.  Host logic only -- no CUDA calls."""
import os
import struct

import numpy as np
import pytest
import torch

from zeroshape_b200.data import formats
from zeroshape_b200.utils.util import EasyDict


def test_ply_roundtrip_and_layout(tmp_path):
    rs = np.random.RandomState(0)
    v = rs.rand(57, 3).astype(np.float64) * 3 - 1.5
    f = rs.randint(0, 57, (101, 3))
    for ascii_ in (False, True):
        p = str(tmp_path / ("a.ply" if ascii_ else "b.ply"))
        formats.write_ply(p, v, f, ascii=ascii_)
        v2, f2 = formats.read_ply(p)
        np.testing.assert_array_equal(f2, f.astype(np.int32))
        np.testing.assert_allclose(v2, v.astype(np.float32), rtol=0, atol=0 if not ascii_ else 1e-7)
    raw = open(str(tmp_path / "b.ply"), "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    assert head.startswith(b"ply\nformat binary_little_endian 1.0\n") and b"property list uchar int vertex_indices" in head
    assert len(body) == 57 * 12 + 101 * 13                              # float32 xyz, then (uchar 3, 3 x int32) per face
    assert struct.unpack_from("<fff", body, 0) == tuple(np.float32(v[0]))
    assert struct.unpack_from("<Biii", body, 57 * 12) == (3,) + tuple(int(i) for i in f[0])
    # empty mesh (an empty iso-surface must still produce a valid file)
    formats.write_ply(str(tmp_path / "e.ply"), np.zeros((0, 3)), np.zeros((0, 3), np.int64))
    v0, f0 = formats.read_ply(str(tmp_path / "e.ply"))
    assert v0.shape == (0, 3) and f0.shape == (0, 3)


class _Graph(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.encoder = torch.nn.Linear(3, 4)
        self.impl_network = torch.nn.Sequential(torch.nn.Linear(4, 2))


class _Model:
    def __init__(self):
        self.graph = _Graph()
        self.optim = torch.optim.AdamW(self.graph.parameters(), lr=1e-3)
        self.sched = torch.optim.lr_scheduler.StepLR(self.optim, 3)


def test_checkpoint_dict_save_resume_partial_load(tmp_path):
    opt = EasyDict(output_path=str(tmp_path), device="cpu")
    m = _Model()
    m.graph(torch.zeros(1, 3)) if False else None
    loss = m.graph.impl_network(m.graph.encoder(torch.ones(2, 3))).sum()
    loss.backward()
    m.optim.step(); m.sched.step()
    formats.save_checkpoint(opt, m, ep=7, it=1234, best_val=0.25, best_ep=5, best=True)
    assert sorted(os.listdir(tmp_path)) == ["best.ckpt", "checkpoint", "latest.ckpt"] and os.listdir(tmp_path / "checkpoint") == ["ep_7.ckpt"]
    ck = torch.load(str(tmp_path / "latest.ckpt"))
    assert set(ck) == {"epoch", "iter", "best_val", "best_ep", "graph", "optim", "sched"} and ck["epoch"] == 7 and ck["iter"] == 1234
    assert set(ck["graph"]) == set(m.graph.state_dict())
    m2 = _Model()
    assert formats.restore_checkpoint(opt, m2, resume=True) == (7, 1234, 0.25, 5)
    for a, b in zip(m.graph.state_dict().values(), m2.graph.state_dict().values()):
        assert torch.equal(a, b)
    assert m2.optim.state_dict()["state"][0]["step"] == m.optim.state_dict()["state"][0]["step"]
    # partial file: only the children present are restored (strict per child), DDP "module." prefixes are accepted
    formats.save_checkpoint(opt, m, 8, 1, 0.2, 8, latest=True, children="impl_network")
    ck = torch.load(str(tmp_path / "latest.ckpt"))
    ck["graph"] = {"module." + k: v for k, v in ck["graph"].items()}
    torch.save(ck, str(tmp_path / "partial.ckpt"))
    m3 = _Model()
    enc_before = m3.graph.encoder.weight.clone()
    assert formats.restore_checkpoint(opt, m3, load_name=str(tmp_path / "partial.ckpt")) == (None, None, None, None)
    assert torch.equal(m3.graph.encoder.weight, enc_before) and torch.equal(m3.graph.impl_network[0].weight, m.graph.impl_network[0].weight)
    with pytest.raises(AssertionError):
        formats.restore_checkpoint(opt, m3, load_name="x", resume=True)


def make_tree(root, H=224, W=224, n_obj=3):
    """A tiny dataset in the layout of data/synthetic.py (PNG images / masks, .npy depth, cameras, point clouds, SDF dicts)."""
    from PIL import Image
    rs = np.random.RandomState(1)
    cat = "chair"
    for d in ("lists", f"images_processed/{cat}", f"masks/{cat}", f"depth/{cat}", f"camera_data/intr/{cat}", f"camera_data/extr/{cat}",
              f"pointclouds/{cat}", f"gt_sdf/{cat}"):
        os.makedirs(os.path.join(root, "objaverse_LVIS", d), exist_ok=True)
    base = os.path.join(root, "objaverse_LVIS")
    names = []
    for o in range(n_obj):
        name = f"{cat}_obj{o}_{o:03d}"
        names.append(name + ".png")
        yy, xx = np.mgrid[:H, :W]
        mask = ((yy - 100 - 5 * o) ** 2 + (xx - 120) ** 2 < (50 + 4 * o) ** 2)
        rgb = np.stack([(xx + 3 * o) % 256, (yy * 2) % 256, (xx + yy) % 256], -1).astype(np.uint8)
        Image.fromarray(rgb, "RGB").save(f"{base}/images_processed/{cat}/{name}.png")
        Image.fromarray((mask * 255).astype(np.uint8), "L").save(f"{base}/masks/{cat}/{name}.png")
        np.save(f"{base}/depth/{cat}/{name}.npy", (mask * (1.2 + 0.01 * o)).astype(np.float32))
        np.save(f"{base}/camera_data/intr/{cat}/{name}.npy", np.array([[300., 0, 112], [0, 300, 112], [0, 0, 1]]))
        Rt = np.eye(4)
        Rt[:3, :3] = np.linalg.qr(rs.randn(3, 3))[0]
        Rt[:3, 3] = [0.1 * o, -0.2, 1.5]
        np.save(f"{base}/camera_data/extr/{cat}/{name}.npy", Rt)
        np.save(f"{base}/pointclouds/{cat}/{cat}_obj{o}.npy", rs.rand(500, 3).astype(np.float32))
        np.save(f"{base}/gt_sdf/{cat}/{cat}_obj{o}.npy", {"sample_pt": rs.rand(2000, 3) - 0.5, "sample_sdf": rs.randn(2000) * 0.1}, allow_pickle=True)
    open(f"{base}/lists/{cat}_train.list", "w").write("\n".join(names) + "\n")
    open(f"{base}/lists/{cat}_val.list", "w").write("\n".join(names[:2]) + "\n")
    return base


def test_synthetic_layout_readers(tmp_path):
    pytest.importorskip("PIL.Image")
    from zeroshape_b200.data.synthetic import Dataset
    make_tree(str(tmp_path))
    opt = EasyDict(H=224, W=224, device="cpu", data=dict(synthetic=dict(subset="objaverse_LVIS", percentage=1), bgcolor=1.0),
                   training=dict(n_sdf_points=512))
    ds = Dataset(opt, "train", path=str(tmp_path), device="cpu")
    assert len(ds) == 3 and ds.list[1] == ("objaverse_LVIS", "chair", "obj1", "001") and ds.cat2label == {"chair": 0}
    assert len(Dataset(opt, "test", path=str(tmp_path), device="cpu")) == 2          # "test" reads the val list
    sub, cat, obj, sid = ds.list[2]
    depth, mask = ds.get_depth(sub, cat, obj, sid)
    assert depth.shape == (1, 224, 224) and torch.equal(mask, (depth != 0).float())
    K, Rt = ds.get_camera(sub, cat, obj, sid)
    assert K.shape == (3, 3) and Rt.shape == (4, 4)
    assert ds.get_pointcloud(sub, cat, obj)["points"].shape == (500, 3)
    pts, sdf = ds.get_gt_sdf(sub, cat, obj)
    raw = np.load(f"{tmp_path}/objaverse_LVIS/gt_sdf/chair/chair_obj2.npy", allow_pickle=True).item()
    assert pts.shape == (2000, 3) and torch.allclose(sdf, torch.from_numpy(raw["sample_sdf"]).float() - 0.003)
    rgba, bbox = ds.get_image(sub, cat, obj, sid)
    assert rgba.shape == (224, 224, 4) and rgba.dtype == np.uint8
    m = rgba[..., 3] > 50
    assert bbox == (np.flatnonzero(m.sum(0))[0], np.flatnonzero(m.sum(1))[0], np.flatnonzero(m.sum(0))[-1], np.flatnonzero(m.sum(1))[-1])
