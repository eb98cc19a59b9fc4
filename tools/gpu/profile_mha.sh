set -x
timeout 600 ncu --set full --clock-control none -k regex:'mha_bwd' --launch-count 2 -o gpurun_out/prof_mha_bwd_tc -f python -c "
import torch
from zeroshape_b200 import ops
q = torch.randn(32, 197, 3*768, device='cuda'); d = torch.randn(32, 197, 768, device='cuda')
for _ in range(2): ops.mha_bwd(q, d, 12, tc=True)
torch.cuda.synchronize()
" > gpurun_out/prof_ncu_mha_bwd.log 2>&1
python - <<'PY' > gpurun_out/prof_mha_bwd_times.txt 2>&1
import torch
from zeroshape_b200 import ops
for B in (16, 32):
    q = torch.randn(B, 197, 3*768, device='cuda'); d = torch.randn(B, 197, 768, device='cuda')
    for tc in (True, False):
        for _ in range(3): ops.mha_bwd(q, d, 12, tc=tc)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): ops.mha_bwd(q, d, 12, tc=tc)
        e1.record(); torch.cuda.synchronize()
        print(f"mha_bwd B={B} T=197 heads=12 hd=64 {'tcgen05' if tc else 'FFMA   '}: {e0.elapsed_time(e1)/20*1e3:8.1f} us")
    for tc in (True, False):
        for _ in range(3): ops.mha(q, 12, tc=tc)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): ops.mha(q, 12, tc=tc)
        e1.record(); torch.cuda.synchronize()
        print(f"mha fwd B={B} T=197 heads=12 hd=64 {'tcgen05' if tc else 'FFMA   '}: {e0.elapsed_time(e1)/20*1e3:8.1f} us")
PY
cat gpurun_out/prof_mha_bwd_times.txt
