#!/bin/bash
# 2 GPUs: NCCL rank-invariance tests (eager + graph), bench N=1 vs N=2 checksums, a short cfg-5 chain on 2 ranks
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_sampler.py -m gpu -x -q -k "nccl or graph" > gpurun_out/pytest_nccl.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_nccl.log
tail -15 gpurun_out/pytest_nccl.log
timeout 600 python bench.py --no-secondary --steps 10 > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "bench n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 2 --no-secondary --steps 10 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
tail -c 1500 gpurun_out/bench_n2.err
python - <<'PY'
import json
for n in (1, 2):
    try:
        d = json.loads(open(f"gpurun_out/bench_n{n}.log").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["state_checksum"], d["sampler"], "e2e", d["e2e"]["value"])
    except Exception as e:
        print(n, "parse failed", e)
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 scripts/run_cfg5_chain.py --walkers 16384 --prefit 10 --nburn 10 --nsteps 20 --steps 5 > gpurun_out/cfg5_n2.log 2> gpurun_out/cfg5_n2.err; echo "cfg5 n2 rc=$?"
tail -c 1500 gpurun_out/cfg5_n2.err; tail -c 2500 gpurun_out/cfg5_n2.log
