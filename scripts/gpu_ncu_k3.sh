mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k3_szmap -s 2 -c 1 -f -o gpurun_out/k3_full python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_k3.log 2>&1
