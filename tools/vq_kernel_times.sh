#!/bin/bash
# ncu launch times of the VQ kernels (forward + backward at the bench shape) under env settings:
#   tools/vq_kernel_times.sh "A=1 B=2" "A=0" ...      -> gpurun_out/kt_<tag>.csv + a summary on stdout
cd "$(dirname "$0")/.."
for setting in "$@"; do
  tag=$(echo "$setting" | tr ' =' '__')
  env $setting ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/kt_${tag}.csv \
      python tools/vq_bwd_profile.py 256 8 49408 512 3 > /dev/null 2>&1
  python - "$setting" gpurun_out/kt_${tag}.csv <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[2])) if len(r) > 10 and r[0].isdigit()]
print("[%s]" % sys.argv[1])
for r in rows[-11:]:
    if "at::" in r[4]: continue
    print("   %8.1f us  %s" % (float(r[-1]) / 1000, r[4][:90]))
PY
done
