# Graphed inference encoder: parity tests, then the headline bench with and without it.
set -x
timeout 600 python -m pytest tests/test_gpu_graph.py tests/test_gpu_eval.py tests/test_gpu_bench_config.py tests/test_gpu_ddp.py -x -q -m gpu 2>&1 | tail -8
ZS_ENCODER_GRAPH=0 timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-shard --no-extras > gpurun_out/eg_bench_0.json 2> gpurun_out/eg_bench_0.err; echo "rc $?"
ZS_ENCODER_GRAPH=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/eg_bench_1.json 2> gpurun_out/eg_bench_1.err; echo "rc $?"
python - <<'PY'
import json
for g in (0, 1):
    try:
        d = json.load(open(f"gpurun_out/eg_bench_{g}.json"))
        print(g, {k: d.get(k) for k in ("value", "ms_per_step", "encoder_ms_per_batch", "e2e", "gpu_launches")}, d["roofline"]["avg_launch_ms"], d.get("e2e_reference_api"), {k: (d.get(k) or {}).get("value") for k in ("config2_vox64", "config3_train_bf16_batch32", "config5_eval_256", "config5_eval_bruteforce_64", "shard_config4")})
    except Exception as e:
        print(g, "failed", e); print(open(f"gpurun_out/eg_bench_{g}.err").read()[-2000:])
PY
