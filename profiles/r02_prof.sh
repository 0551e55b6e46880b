#!/bin/bash
# Round-2 profile collection on ONE B200 (run under gpurun from the repo root):  bash profiles/r02_prof.sh
# 1. bench line (default arguments)                      -> gpurun_out/r02_bench_n1.json
# 2. ncu launch list of the short bench command          -> gpurun_out/r02_launches.csv   (time + tensor pipe + issue slots per launch)
# 3. ncu --set full of the three big tcgen05 launches    -> gpurun_out/r02_big3.ncu-rep   (third update of the same command)
# 4. the other named workloads (ML-1M fit+predict, stress shape)
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
echo "bench rc=$?"
SHORT="python bench.py --steps 2 --warmup 1 --no-cpu"
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,launch__grid_size
$SHORT > gpurun_out/plain.log 2>&1 &&
ncu --metrics $M --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches.csv $SHORT > gpurun_out/ncu_l.log 2>&1
echo "launch list rc=$?"
$SHORT > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:tc_(fwd|bwd1|bwd2)_h2_kernel' -s 6 -c 4 -f -o gpurun_out/r02_big3 $SHORT > gpurun_out/ncu_f.log 2>&1
echo "set full rc=$?"
python bench.py --workload ml1m > gpurun_out/r02_bench_ml1m.json 2> gpurun_out/r02_bench_ml1m.err
echo "ml1m rc=$?"
python bench.py --workload stress --steps 200 > gpurun_out/r02_bench_stress.json 2> gpurun_out/r02_bench_stress.err
echo "stress rc=$?"
python bench.py --workload stress --steps 200 --stress-precision f16x3 > gpurun_out/r02_bench_stress_f16x3.json 2> gpurun_out/r02_bench_stress_f16x3.err
echo "stress f16x3 rc=$?"
tail -c 600 gpurun_out/r02_bench_n1.err gpurun_out/r02_bench_ml1m.err gpurun_out/r02_bench_stress.err gpurun_out/ncu_l.log gpurun_out/ncu_f.log
