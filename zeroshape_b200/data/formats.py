"""On-disk formats either side of the hot path.

  write_ply / read_ply : the mesh files `mesh.export("x.ply")` leaves behind (utils/util_vis.py:104-108): trimesh's default PLY
                         encoding -- binary_little_endian 1.0, float32 x y z vertices, faces as `uchar 3` + 3 x int32.
  save_checkpoint / load_checkpoint : the checkpoint dict of utils/util.py:227-275: {epoch, iter, best_val, best_ep, graph =
                         state_dict, optim* / sched* / scaler* state dicts}, `latest.ckpt` (+ `best.ckpt`, `checkpoint/ep_N.ckpt`),
                         per-child restore with strict key checking.
"""
import os
import shutil

import numpy as np
import torch

PLY_HEADER = ("ply\nformat {fmt} 1.0\ncomment zeroshape_b200\nelement vertex {nv}\nproperty float x\nproperty float y\n"
              "property float z\nelement face {nf}\nproperty list uchar int vertex_indices\nend_header\n")


def write_ply(path, vertices, faces, ascii=False):
    v = np.ascontiguousarray(np.asarray(vertices), dtype="<f4").reshape(-1, 3)
    f = np.ascontiguousarray(np.asarray(faces), dtype="<i4").reshape(-1, 3)
    with open(path, "wb") as fh:
        fh.write(PLY_HEADER.format(fmt="ascii" if ascii else "binary_little_endian", nv=len(v), nf=len(f)).encode("ascii"))
        if ascii:
            for p in v:
                fh.write(("%.9g %.9g %.9g\n" % tuple(p)).encode("ascii"))
            for t in f:
                fh.write(("3 %d %d %d\n" % tuple(t)).encode("ascii"))
        else:
            fh.write(v.tobytes())
            rec = np.empty(len(f), dtype=[("n", "u1"), ("i", "<i4", (3,))])
            rec["n"] = 3
            rec["i"] = f
            fh.write(rec.tobytes())


def read_ply(path):
    """-> (vertices float32 [V,3], faces int32 [F,3]); the subset of PLY that write_ply and trimesh's exporter produce."""
    with open(path, "rb") as fh:
        assert fh.readline().strip() == b"ply"
        fmt, nv, nf, vprops = None, 0, 0, []
        elem = None
        while True:
            line = fh.readline().decode("ascii").strip()
            if line == "end_header":
                break
            tok = line.split()
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                elem = tok[1]
                if elem == "vertex":
                    nv = int(tok[2])
                elif elem == "face":
                    nf = int(tok[2])
            elif tok[0] == "property" and elem == "vertex":
                vprops.append((tok[2], tok[1]))
        types = {"float": "<f4", "float32": "<f4", "double": "<f8", "uchar": "u1", "uint8": "u1", "int": "<i4", "int32": "<i4"}
        if fmt == "ascii":
            rows = [fh.readline().split() for _ in range(nv)]
            v = np.array([[float(r[i]) for i in range(3)] for r in rows], np.float32).reshape(-1, 3)
            f = np.array([[int(x) for x in fh.readline().split()[1:4]] for _ in range(nf)], np.int32).reshape(-1, 3)
            return v, f
        assert fmt == "binary_little_endian", fmt
        vd = np.dtype([(n, types[t]) for n, t in vprops])
        vraw = np.frombuffer(fh.read(vd.itemsize * nv), dtype=vd)
        v = np.stack([vraw["x"], vraw["y"], vraw["z"]], axis=1).astype(np.float32)
        fd = np.dtype([("n", "u1"), ("i", "<i4", (3,))])
        fraw = np.frombuffer(fh.read(fd.itemsize * nf), dtype=fd)
        assert (fraw["n"] == 3).all()
        return v, fraw["i"].astype(np.int32)


def _graph_of(model):
    g = model.graph
    return g.module if isinstance(g, (torch.nn.DataParallel, torch.nn.parallel.DistributedDataParallel)) else g


def save_checkpoint(opt, model, ep, it, best_val, best_ep, latest=False, best=False, children=None):
    """utils/util.py:251-275."""
    os.makedirs("{0}/checkpoint".format(opt.output_path), exist_ok=True)
    graph = _graph_of(model)
    sd = graph.state_dict()
    if children is not None:
        sd = {k: v for k, v in sd.items() if k.startswith(children)}
    checkpoint = dict(epoch=ep, iter=it, best_val=best_val, best_ep=best_ep, graph=sd)
    for key in model.__dict__:
        if key.split("_")[0] in ["optim", "sched", "scaler"]:
            checkpoint.update({key: getattr(model, key).state_dict()})
    torch.save(checkpoint, "{0}/latest.ckpt".format(opt.output_path))
    if best:
        shutil.copy("{0}/latest.ckpt".format(opt.output_path), "{0}/best.ckpt".format(opt.output_path))
    if not latest:
        shutil.copy("{0}/latest.ckpt".format(opt.output_path), "{0}/checkpoint/ep_{1}.ckpt".format(opt.output_path, ep))


def get_child_state_dict(state_dict, key):
    """utils/util.py:201-210: the entries of a child module, `module.` (DDP) and the child prefix stripped."""
    out = {}
    for k, v in state_dict.items():
        name = k[7:] if k.startswith("module.") else k
        if name.startswith("{}.".format(key)):
            out[".".join(name.split(".")[1:])] = v
    return out


def _load(opt, name):
    dev = opt.device if isinstance(opt.device, torch.device) else torch.device(opt.device)
    return torch.load(name, map_location=dev)


def load_checkpoint(opt, model, load_name):
    """utils/util.py:227-238: restore every child module present in the file (strict per child), skip the others."""
    checkpoint = _load(opt, load_name)
    for name, child in _graph_of(model).named_children():
        child_sd = get_child_state_dict(checkpoint["graph"], name)
        if child_sd:
            child.load_state_dict(child_sd, strict=True)
    return None, None, None, None


def resume_checkpoint(opt, model, best=False):
    """utils/util.py:212-225: latest.ckpt / best.ckpt -> the whole graph (strict) + optimizer / scheduler / scaler state."""
    checkpoint = _load(opt, "{0}/{1}.ckpt".format(opt.output_path, "best" if best else "latest"))
    _graph_of(model).load_state_dict(checkpoint["graph"], strict=True)
    for key in model.__dict__:
        if key.split("_")[0] in ["optim", "sched", "scaler"] and key in checkpoint:
            getattr(model, key).load_state_dict(checkpoint[key])
    return checkpoint["epoch"], checkpoint["iter"], checkpoint["best_val"], checkpoint["best_ep"] if "best_ep" in checkpoint else 0


def restore_checkpoint(opt, model, load_name=None, resume=False, best=False, evaluate=False):
    """utils/util.py:240-249."""
    assert not (load_name is not None and resume)
    if resume:
        return resume_checkpoint(opt, model, best)
    return load_checkpoint(opt, model, load_name)
