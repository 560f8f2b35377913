# usage: gpu_scale.sh N   -- bench at N ranks (torchrun) into gpurun_out/bench_nN.log
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
fi
echo "rc=$?" >> gpurun_out/bench_n$N.err
