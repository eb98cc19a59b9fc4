set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
echo "bench rc $?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s_launches.csv python bench.py --steps 1 --warmup 3 --shapes 1 --no-cpu-baseline --no-e2e --no-shard --no-extras --profile-region > gpurun_out/r2s_launch_bench.log 2>&1
echo "launch list rc $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'chain|point_proj' --launch-skip 10 --launch-count 5 -o gpurun_out/r2s_decoder_full -f python tools/diag_decoder.py 2146689 --once > gpurun_out/r2s_ncu_full.log 2>&1
echo "ncu full rc $?"
timeout 600 ncu --set full --clock-control none -k regex:'mha_tc' --launch-count 2 -o gpurun_out/r2s_mha_tc -f python -c "
import torch
from zeroshape_b200 import ops
q = torch.randn(8, 197, 3*768, device='cuda')
for _ in range(3): ops.mha(q, 12, tc=True)
torch.cuda.synchronize()
" > gpurun_out/r2s_ncu_mha.log 2>&1
echo "ncu mha rc $?"
