#!/bin/bash
# ncu launch times of the VQ forward kernels under the given env settings: tools/vq_s1_times.sh "A=1 B=2" "A=0" ...
cd "$(dirname "$0")/.."
for setting in "$@"; do
  tag=$(echo "$setting" | tr ' =' '__')
  env $setting ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s1_${tag}.csv \
      python tools/vq_ab.py "" > /dev/null 2>&1
  python - "$setting" gpurun_out/s1_${tag}.csv <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[2])) if len(r) > 10 and r[0].isdigit()]
agg = collections.defaultdict(list)
for r in rows:
    agg[r[4][:70]].append(float(r[-1]))
print("[%s]" % sys.argv[1])
for k, v in agg.items():
    if any(t in k for t in ("Sweep1", "Sweep2", "colsum", "vq_select")):
        v2 = v[len(v) // 2:]
        print("   %8.1f us  %s" % (sum(v2) / len(v2) / 1000, k))
PY
done
