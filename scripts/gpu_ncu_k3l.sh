mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k3l_szmap -s 2 -c 1 -f -o gpurun_out/k3l_full python bench.py --workload synth255 --walkers 8192 --steps 2 --warmup 1 > gpurun_out/ncu_k3l.log 2>&1
