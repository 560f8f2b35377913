#!/bin/bash
# 2 GPUs: rank-invariance tests (NCCL and peer-memory exchange, eager + graph), bench N=1 vs N=2 checksums
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python -m pytest tests/test_sampler.py -m gpu -x -q -k "nccl" > gpurun_out/pytest_nccl.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_nccl.log
tail -25 gpurun_out/pytest_nccl.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 2 --no-secondary --steps 10 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench n2 (auto) rc=$?"
grep "^\[bench" gpurun_out/bench_n2.err | tail -4; tail -c 600 gpurun_out/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29503 bench.py --gpus 2 --no-secondary --steps 10 --exchange nccl > gpurun_out/bench_n2_nccl.log 2> gpurun_out/bench_n2_nccl.err; echo "bench n2 (nccl) rc=$?"
timeout 300 python bench.py --no-secondary --steps 10 > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "bench n1 rc=$?"
python - <<'PY'
import json
for n in ("n1", "n2", "n2_nccl"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{n}.log").read().strip().splitlines()[-1])
        print(n, "%.4g" % d["value"], "%.3f ms" % d["ms_per_step"], d["state_checksum"], d["sampler"]["cuda_graph"], d["sampler"]["exchange"][:5], d["sampler"].get("p2p_fallback_reason"), "e2e %.4g" % d["e2e"]["value"])
    except Exception as e:
        print(n, "parse failed", e)
PY
