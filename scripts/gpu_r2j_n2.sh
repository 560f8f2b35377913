#!/bin/bash
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_sampler.py -m gpu -x -q -k "nccl" > gpurun_out/pytest_nccl.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_nccl.log
tail -8 gpurun_out/pytest_nccl.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 2 --no-secondary --steps 10 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench n2 (auto) rc=$?"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29503 bench.py --gpus 2 --no-secondary --steps 10 --exchange nccl > gpurun_out/bench_n2_nccl.log 2> gpurun_out/bench_n2_nccl.err; echo "bench n2 (nccl) rc=$?"
python - <<'PY'
import json
for n in ("n2", "n2_nccl"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{n}.log").read().strip().splitlines()[-1])
        print(n, "%.4g" % d["value"], "%.3f ms" % d["ms_per_step"], d["state_checksum"], d["sampler"]["cuda_graph"], d["sampler"]["exchange"][:5], "e2e %.4g" % d["e2e"]["value"], d["stage_ms_per_launch"]["profiles"])
    except Exception as e:
        print(n, "parse failed", e)
PY
