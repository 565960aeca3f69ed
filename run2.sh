set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
python bench.py --n 60 --steps 30 --warmup 5 --cpu-budget 3 2>gpurun_out/bench_n60.err | tee gpurun_out/bench_n60.json | cut -c1-1500
python bench.py --steps 50 --warmup 5 2>gpurun_out/bench_full.err | tee gpurun_out/bench_full.json | cut -c1-3000
tail -5 gpurun_out/bench_full.err
python bench.py --solver cg --steps 50 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_cg.err | tee gpurun_out/bench_cg.json | cut -c1-3000
