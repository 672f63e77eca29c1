#!/bin/bash
# Round-end evidence on ONE box (gpurun -- bash tools/final_runs.sh): the GPU test files one process each (a sticky CUDA error in
# one must not hide the others), the bench lines of every configuration, and the ncu launch list of the headline command.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for f in tests/test_gpu_ops.py tests/test_gpu_forward.py tests/test_gpu_viz.py tests/test_gpu_panocam.py tests/test_gpu_dist.py; do
  [ -f $f ] || continue
  timeout 900 python -m pytest $f -x -q -m gpu 2>&1 | tail -2 | sed "s|^|$f: |"
done 2>&1 | tee gpurun_out/final_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/final_bench_c2_n1.json 2> gpurun_out/final_bench_c2_n1.err; tail -c 600 gpurun_out/final_bench_c2_n1.err
for c in C3 C5 P360; do
  timeout 400 python bench.py --config $c --no-cpu-baseline > gpurun_out/final_bench_${c}_n1.json 2> gpurun_out/final_bench_${c}.err
done
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference_arm.json 2> gpurun_out/final_bench_reference.err
for f in gpurun_out/final_bench_*.json; do echo $f; python tools/bench_show.py $f 2>&1 | head -1; done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none -c 3000 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile-passes > gpurun_out/final_ncu_bench.log 2>&1
python tools/ncu_summary.py gpurun_out/final_launches.csv > gpurun_out/final_launches_summary.csv; head -12 gpurun_out/final_launches_summary.csv
