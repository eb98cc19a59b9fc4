# Training step as one CUDA graph: the parity tests, then BASELINE config 3 launched op by op and replayed from the graph.
#   /usr/local/graft/bin/gpurun --timeout 900 -- "bash tools/gpu/train_graph.sh"
set -x
timeout 700 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_tc.py tests/test_gpu_ops.py -x -q -m gpu 2>&1 | tail -8
timeout 300 python bench.py --mode train --steps 10 --warmup 3 --train-graph 0 > gpurun_out/train32_eager.json 2> gpurun_out/train32_eager.err; echo "eager rc $?"
timeout 300 python bench.py --mode train --steps 10 --warmup 3 --train-graph 1 > gpurun_out/train32_graph.json 2> gpurun_out/train32_graph.err; echo "graph rc $?"
python - <<'PY'
import json
for n in ("eager", "graph"):
    try:
        d = json.load(open(f"gpurun_out/train32_{n}.json"))
        print(n, {k: d.get(k) for k in ("value", "ms_per_step", "images_per_s", "last_loss", "cuda_graph", "cuda_graph_error", "gpu_launches", "clocks", "peak_memory_gb")})
    except Exception as e:
        print(n, "failed", e)
        print(open(f"gpurun_out/train32_{n}.err").read()[-1500:])
PY
