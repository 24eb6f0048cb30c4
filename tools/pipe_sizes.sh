#!/bin/bash
# VQ backward kernel times (ncu launch list) of the two-kernel path (mode 0) and the fused pipeline (mode 1) vs M
for B in 32 64 128 256; do
  for mode in 0 1; do
    SCP_VQ_BWD_PIPE=$mode ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/sz_${B}_$mode.csv python tools/vq_bwd_profile.py $B 8 49408 512 2 > /dev/null 2>&1
    python - "$B" "$mode" <<'PY'
import csv, sys
B, mode = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(f"gpurun_out/sz_{B}_{mode}.csv")) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
names = [(r[ki], float(r[vi]) / 1e3) for r in rows[1:]]
# last iteration: from the last vq_bwd_prep to the end
last = max(i for i, (n, _) in enumerate(names) if "vq_bwd_prep" in n)
bwd = names[last:]
print(f"M={int(B)*8} mode={mode} bwd total {sum(t for _, t in bwd):.1f} us :", ", ".join(f"{n.split('(')[0].split('::')[-1][:22]} {t:.1f}" for n, t in bwd))
PY
  done
done
