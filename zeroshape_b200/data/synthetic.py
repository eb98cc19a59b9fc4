"""Reader of the reference's synthetic training set layout (data/synthetic.py): the same directory tree, file formats and sample
dict, with the image path on the device pre-processing of data/preprocess.py.

    <path>/<subset>/lists/<cat>_{train,val}.list            one image file name per line: <cat>_<object>_<sample>.png
    <path>/<subset>/images_processed/<cat>/<name>.png, masks/<cat>/<name>.png
    <path>/<subset>/depth/<cat>/<name>.npy                   [H, W] float depth, 0 = background
    <path>/<subset>/camera_data/{intr,extr}/<cat>/<name>.npy  3x3 intrinsics, 3x4 (or 4x4) world-to-camera [R|t]
    <path>/<subset>/pointclouds/<cat>/<cat>_<object>.npy     [P, 3]
    <path>/<subset>/gt_sdf/<cat>/<cat>_<object>.npy          pickled dict {sample_pt [S,3], sample_sdf [S]}
"""
import os
from copy import deepcopy

import numpy as np
import torch

from ..utils import camera
from . import preprocess as PP


class Dataset(torch.utils.data.Dataset):
    """data/synthetic.py:10-176 (+ data/base.py).  `device`: where rgb_input_map is produced (the crop / resize kernels)."""

    def __init__(self, opt, split="train", load_3D=True, path="data/train_data", device=None):
        super().__init__()
        if split == "test":
            split = "val"
        self.opt, self.split = deepcopy(opt), split
        self.path, self.load_3D = path, load_3D
        self.device = device if device is not None else getattr(opt, "device", "cuda")
        self.subsets = opt.data.synthetic.subset.split(",")
        self.category_dict, self.category_list = {}, []
        for subset in self.subsets:
            names = sorted(os.listdir("{}/{}/lists".format(self.path, subset)))
            cats = [name[:-11] for name in names if name.endswith("_train.list")]
            self.category_dict[subset] = cats
            self.category_list += cats
        if split == "val":                               # max 10 images per category (data/synthetic.py:28-31)
            self.max_imgs, self.data_percentage = 10, 1
        else:
            self.max_imgs, self.data_percentage = np.inf, opt.data.synthetic.percentage
        self.cat2label = {c: i for i, c in enumerate(self.category_list)}
        self.label2cat = list(self.category_list)
        self.list = self.get_list(opt, split)

    def get_list(self, opt, split):
        data_list = []
        for subset in self.subsets:
            for cat in self.category_dict[subset]:
                list_fname = f"{self.path}/{subset}/lists/{cat}_{split}.list"
                if not os.path.exists(list_fname):
                    continue
                lines = open(list_fname).read().splitlines()
                lines = lines[:round(self.data_percentage * len(lines))]
                for i, img_fname in enumerate(lines):
                    if i >= self.max_imgs:
                        break
                    name = ".".join(img_fname.split(".")[:-1])
                    data_list.append((subset, cat, name.split("_")[-2], name.split("_")[-1]))
        return data_list

    def __len__(self):
        return len(self.list)

    # ---- files -> arrays (host) ---------------------------------------------------------------------------------------
    def get_image(self, subset, category, object_name, sample_id):
        """-> (uint8 RGBA [H,W,4] numpy, bbox or None); the mask is binarised at 50 for the box (data/synthetic.py:77-94)."""
        from PIL import Image
        fname = f"{category}/{category}_{object_name}_{sample_id}"
        image = Image.open(f"{self.path}/{subset}/images_processed/{fname}.png").convert("RGB")
        mask = Image.open(f"{self.path}/{subset}/masks/{fname}.png").convert("L")
        mask_np = np.array(mask)
        mask_np[mask_np <= 50] = 0
        mask_np[mask_np >= 50] = 1.0
        rgba = np.dstack([np.asarray(image), np.asarray(mask)])
        return rgba, PP.get_bbox_from_mask(mask_np, 0.5, min_pixels=10)

    def get_depth(self, subset, category, object_name, sample_id):
        fname = f"{category}/{category}_{object_name}_{sample_id}"
        depth = torch.tensor(np.load(f"{self.path}/{subset}/depth/{fname}.npy")).unsqueeze(0)
        assert depth.shape[1] == self.opt.H
        return depth, 1 - (depth == 0).float()

    def get_camera(self, subset, category, object_name, sample_id):
        fname = f"{category}/{category}_{object_name}_{sample_id}"
        Rt = np.load(f"{self.path}/{subset}/camera_data/extr/{fname}.npy")
        K = torch.from_numpy(np.load(f"{self.path}/{subset}/camera_data/intr/{fname}.npy"))
        return K, Rt

    def get_pointcloud(self, subset, category, object_name):
        pc = np.load(f"{self.path}/{subset}/pointclouds/{category}/{category}_{object_name}.npy")
        return {"points": torch.from_numpy(pc).float()}

    def get_gt_sdf(self, subset, category, object_name):
        gt = np.load(f"{self.path}/{subset}/gt_sdf/{category}/{category}_{object_name}.npy", allow_pickle=True).item()
        return torch.from_numpy(gt["sample_pt"]).float(), torch.from_numpy(gt["sample_sdf"]).float() - 0.003

    def preprocess_image(self, opt, rgba, bbox):
        """data/synthetic.py:193-199: resize to opt.W x opt.H when needed (no crop: the set is pre-cropped), to_tensor, keep rgb."""
        img = PP._as_rgba_u8(rgba, self.device)
        if img.shape[1] != opt.W or img.shape[1] != opt.H:          # the reference tests size[0] twice; kept
            from .. import ops
            xb, xk = PP._coeffs_on(img.device, img.shape[1], opt.W)
            yb, yk = PP._coeffs_on(img.device, img.shape[0], opt.H)
            img = ops.rgba_crop_resize(img, 0, 0, img.shape[1], img.shape[0], opt.H, opt.W, xb, xk, yb, yk)
        from .. import ops
        return ops.rgba_composite(img, None)[0]

    def __getitem__(self, idx):
        opt = self.opt
        subset, category, object_name, sample_id = self.list[idx]
        sample = dict(idx=idx, category_label=self.cat2label[category])
        K, Rt = self.get_camera(subset, category, object_name, sample_id)
        R = np.zeros((3, 4))
        R[:3, :3] = Rt[:3, :3]
        pose = camera.pose.compose([R, camera.pose(t=Rt[:3, 3])])
        sample.update(pose_gt=pose.float(), intr=K.float())
        rgba, bbox = self.get_image(subset, category, object_name, sample_id)
        depth, mask_input_map = self.get_depth(subset, category, object_name, sample_id)
        sample.update(rgb_input_map=self.preprocess_image(opt, rgba, bbox), mask_input_map=mask_input_map, depth_input_map=depth)
        if not self.load_3D:
            return sample
        sample.update(dpc=self.get_pointcloud(subset, category, object_name))
        pts, sdf = self.get_gt_sdf(subset, category, object_name)
        if opt.training.n_sdf_points:
            sel = torch.randperm(pts.shape[0])[:opt.training.n_sdf_points]
            pts, sdf = pts[sel], sdf[sel]
        sample.update(gt_sample_points=pts, gt_sample_sdf=sdf)
        return sample
