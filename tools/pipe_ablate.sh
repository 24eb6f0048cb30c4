#!/bin/bash
# producer / consumer kernel times of the VQ backward pipeline under the timing ablations (run on the GPU box)
# needs a library built with SCP_BUILD_ABLATION=1 python -m speechclip_plus_b200.build (a regular build ignores SCP_PIPE_DEBUG)
# bits: 1 no epilogue maths, 2 no MMAs, 4 no P~/Q~ stores, 8 no TMEM loads
for dbg in ${@:-0 8 12 13 15}; do
  SCP_VQ_BWD_PIPE=2 SCP_PIPE_DEBUG=$dbg ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/abl_$dbg.csv python tools/vq_bwd_profile.py 256 8 49408 512 2 > /dev/null 2>&1
  echo "debug=$dbg (separate): $(grep vq_bwd_pipe_kernel gpurun_out/abl_$dbg.csv | tail -2 | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
done
