# The gpurun command set behind profiles/r2_*: bench line, launch list of one step, ncu --set full of the decoder pass and of the ViT attention.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- "bash tools/gpu/profile_round2.sh"   -> gpurun_out/prof_*
set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/prof_bench.json 2> gpurun_out/prof_bench.err
echo "bench rc $?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/prof_launches.csv python bench.py --steps 1 --warmup 3 --shapes 1 --no-cpu-baseline --no-e2e --no-shard --no-extras --profile-region > gpurun_out/prof_launch_bench.log 2>&1
echo "launch list rc $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'chain|point_proj' --launch-skip 10 --launch-count 5 -o gpurun_out/prof_decoder_full -f python tools/diag_decoder.py 2146689 --once > gpurun_out/prof_ncu_full.log 2>&1
echo "ncu full rc $?"
timeout 600 ncu --set full --clock-control none -k regex:'mha_tc' --launch-count 2 -o gpurun_out/prof_mha_tc -f python -c "
import torch
from zeroshape_b200 import ops
q = torch.randn(8, 197, 3*768, device='cuda')
for _ in range(3): ops.mha(q, 12, tc=True)
torch.cuda.synchronize()
" > gpurun_out/prof_ncu_mha.log 2>&1
echo "ncu mha rc $?"
