"""One launch of each round-2 training kernel at the BASELINE config 3 sizes (batch 32), for `ncu --set full` (tools/gpu/profile_train_kernels.sh)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zeroshape_b200 import ops

dev = torch.device("cuda", 0)
ops.TRAIN_ENGINE, ops.TRAIN_PRECISION = "tc", "bf16"
B, P, L, H, C = 32, 4096, 197, 8, 256
qkv_p = torch.randn(B, P, 3 * C, device=dev) * 0.7
lat = torch.randn(B, L, 3 * C, device=dev) * 0.7
for _ in range(2):
    out = ops.point_attention(qkv_p, lat[..., C:2 * C], lat[..., 2 * C:], H)
    ops.point_attention_bwd(qkv_p, lat[..., C:2 * C], lat[..., 2 * C:], out, torch.randn_like(out), H)
    qkv = torch.randn(B, 197, 3 * 768, device=dev)
    o = ops.mha(qkv, 12, tc=True, precision="fp16")
    ops.mha_bwd(qkv, torch.randn_like(o), 12)
    x = torch.randn(B, 56, 56, 256, device=dev)
    ops.groupnorm_nhwc(x, torch.ones(256, device=dev), torch.zeros(256, device=dev), 32, 1e-5, True)
    ops.conv2d_nhwc_dgrad(torch.randn(B, 112, 112, 64, device=dev), torch.randn(64, 7, 7, 3, device=dev), (B, 224, 224, 3), 2, (3, 3, 3, 3), tc=False)
    y = ops.bilinear_nhwc(torch.randn(B, 112, 112, 128, device=dev), 224, 224, True)
    ops.bilinear_bwd_nhwc(y, 112, 112, True)
torch.cuda.synchronize()
print("done")
