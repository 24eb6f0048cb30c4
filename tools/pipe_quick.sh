#!/bin/bash
for cfg in "1 4" "1 16" "1 2"; do
  set -- $cfg
  for B in 32 256; do
    SCP_VQ_BWD_PIPE=$1 SCP_VQ_BWD_RING=$2 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/q.csv python tools/vq_bwd_profile.py $B 8 49408 512 2 > /dev/null 2>&1
    echo "mode=$1 ring=$2 M=$((B*8)): $(grep vq_bwd_pipe_kernel gpurun_out/q.csv | tail -1 | awk -F'","' '{print $NF}' | tr -d '"')"
  done
done
SCP_VQ_BWD_PIPE=1 timeout 200 python tools/vq_bwd_check.py --shapes full > gpurun_out/bwdcheck_1.log 2>&1; echo "check rc=$?"
