#!/bin/bash
# First GPU commands of the next session (round 1 ended with no GPU minutes left while the code below was written and
# pinned on the CPU only). Run through gpurun, one block per call; outputs go to gpurun_out/<tag>/.
#
#   gpurun --timeout 1500 -- 'bash scripts/gpu_next_session.sh tests'
#   gpurun --timeout 1200 -- 'bash scripts/gpu_next_session.sh sweep'
#   gpurun --gpus 8 --timeout 900 -- 'bash scripts/gpu_next_session.sh config4'
#   gpurun --timeout 900 -- 'bash scripts/gpu_next_session.sh config5'
#   gpurun --gpus 8 --timeout 900 -- 'bash scripts/gpu_next_session.sh scale8'
set -u
what=${1:-tests}
out=gpurun_out/next_$what
mkdir -p "$out"
case "$what" in
  tests)
    # 1. everything already proven, then the tests written without a GPU (tests/conftest.py: RUN_LAST collects them last):
    #    sb_apply_accumulate, the playground's Cahn-Hilliard step + application, sb_eval_group, grouped IDR(s)/BiCGStab(l),
    #    automatic statement grouping. Once green: empty RUN_LAST.
    timeout 1300 python -m pytest tests -m gpu -q 2>&1 | tail -40 > "$out/pytest_gpu.log"
    cat "$out/pytest_gpu.log"
    python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; tail -5 "$out/smoke.log"
    ;;
  sweep)
    # 2. what the statement groups buy at full size: reference templates as written / with automatic grouping / the
    #    grouped classes (expected from the bytes: IDR(4) x1.29, BiCGStab(2) x1.14, GMRES(50) generic x1.25)
    python scripts/solver_sweep.py --solvers idrs,bicgstabl,tfqmr,tfqmr1,cgs,gmres,grouped_idrs,grouped_bicgstabl \
        --out "$out/solver_sweep_as_written.json" > "$out/as_written.log" 2>&1
    python scripts/solver_sweep.py --grouping --solvers idrs,bicgstabl,tfqmr,tfqmr1,cgs,gmres,cg,bicgstab \
        --out "$out/solver_sweep_grouping.json" > "$out/grouping.log" 2>&1
    python scripts/solver_sweep.py --grouping 2 --solvers idrs,bicgstabl,tfqmr,tfqmr1,cgs,gmres,cg,bicgstab \
        --out "$out/solver_sweep_scheduling.json" > "$out/scheduling.log" 2>&1
    tail -3 "$out/as_written.log" "$out/grouping.log" "$out/scheduling.log"
    # the playground's caller at full size, as written and with automatic grouping
    python scripts/playground_cahn_hilliard.py --box 119 --steps 1 --max-iterations 100 --no-vtk --out "$out/ch" > "$out/ch_10M.log" 2>&1
    python scripts/playground_cahn_hilliard.py --box 119 --steps 1 --max-iterations 100 --no-vtk --grouping --out "$out/ch" > "$out/ch_10M_grouping.log" 2>&1
    python scripts/playground_cahn_hilliard.py --box 119 --steps 2 --uniformed --no-vtk --grouping 2 --out "$out/ch" > "$out/ch_10M_uniformed.log" 2>&1
    tail -1 "$out/ch_10M.log" "$out/ch_10M_grouping.log" "$out/ch_10M_uniformed.log"
    ncu --set full --clock-control none --import-source on -k regex:GroupBody -c 3 -o "$out/group_kernel" \
        python scripts/solver_sweep.py --axis 119 --steps 4 --repeats 1 --solvers grouped_idrs > "$out/ncu.log" 2>&1
    ;;
  config4)
    # 3. config 4 at full size: 49.8 M hexahedra on 8 GPUs, every rank builds its own slab (4.6 s)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
        scripts/config4_projection.py --axis 368 --lattice --steps 5 > "$out/config4_368_n8_lattice.json" 2> "$out/config4.err"
    tail -c 1500 "$out/config4_368_n8_lattice.json"
    ;;
  scale8)
    # 5. the 8-GPU efficiency gap (51 % at 10.1 M cells): two cheap A/B runs before any kernel work -- programmatic
    #    dependent launch was measured slower at 10 M cells per GPU, never at 1.26 M per rank where the 8 kernel
    #    boundaries are 20 of 103 us; and the slab partition (2 neighbours per rank instead of up to 7)
    for env in "SB_PDL=0" "SB_PDL=1"; do
      env $env python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 \
          bench.py --gpus 8 --steps 200 --warmup 20 > "$out/bench_n8_$env.json" 2> "$out/bench_n8_$env.err"
      tail -c 600 "$out/bench_n8_$env.json"
    done
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 \
        bench.py --gpus 8 --steps 200 --warmup 20 --partition slab > "$out/bench_n8_slab.json" 2> "$out/bench_n8_slab.err"
    tail -c 600 "$out/bench_n8_slab.json"
    ;;
  config5)
    # 4. the 1e8 / 2e8-cell points of the apply sweep (direct lattice generator)
    python scripts/apply_sweep.py --cells hexlat --sizes 1e8,2e8 --out "$out/apply_sweep_hexlat.json" \
        > "$out/config5.log" 2>&1
    tail -5 "$out/config5.log"
    # and, on an 8-GPU box: python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
    #   --master-port 29518 scripts/apply_sweep.py --cells hexlat --sizes 1e7,1e8,2e8 --out "$out/apply_sweep_hexlat_n8.json"
    ;;
esac
