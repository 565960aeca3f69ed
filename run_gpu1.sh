mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py 2>gpurun_out/bench_bicgstab.err | tee gpurun_out/bench_bicgstab.json | cut -c1-600
tail -5 gpurun_out/bench_bicgstab.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:apply_kernel_tma -s 6 -c 2 -o gpurun_out/apply_v3_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out
