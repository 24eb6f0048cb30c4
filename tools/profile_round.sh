#!/bin/bash
# One GPU call that produces the evidence kept under profiles/ (B200_PROFILING.md recipe):
#   1. plain bench run (exit 0 required)             -> gpurun_out/<tag>_bench.json
#   2. ncu launch list of the SAME short command     -> gpurun_out/<tag>_launches.csv
#   3. ncu --set full of one instance of every hot kernel -> gpurun_out/<tag>_full.ncu-rep (+ raw csv)
#   4. device times of the kernels outside the bench step -> gpurun_out/<tag>_aux_kernels.json
# usage: tools/profile_round.sh <tag>
set -u
cd "$(dirname "$0")/.."
tag=${1:-r01}
mkdir -p gpurun_out
SHORT="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-breakdown --no-graph"
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || { echo "bench failed"; tail -5 gpurun_out/${tag}_bench.err; exit 1; }
$SHORT > gpurun_out/${tag}_short.json 2> gpurun_out/${tag}_short.err || { echo "short bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    $SHORT > gpurun_out/${tag}_ncu_launches.log 2>&1
# one instance of each hot kernel, taken from the last (5th) step: 11 matching kernels per step (vq_select runs twice, vq_colsum once), skip the first 4 steps
ncu --set full --import-source on --clock-control none \
    -k regex:'wsum_fwd|wsum_bwd_plain|stream_gemm|vq_select|vq_colsum' -s 44 -c 11 -f -o gpurun_out/${tag}_full \
    $SHORT > gpurun_out/${tag}_ncu_full.log 2>&1
ncu -i gpurun_out/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>/dev/null
# kernels outside the bench step (S1' variants, N1 / N3 / N4): device times and achieved bandwidth
python tools/aux_kernel_times.py > gpurun_out/${tag}_aux_kernels.json 2> gpurun_out/${tag}_aux_kernels.err


tail -2 gpurun_out/${tag}_ncu_full.log
cat gpurun_out/${tag}_bench.json
