#!/bin/bash
# developer sweep: host-pipeline slots x chunk size for the host-buffer (e2e) entry point
for SL in 4 6 8; do
  OMEGA4_NVCC_EXTRA="-DOMEGA4_HOST_SLOTS=$SL" python audio-analyzer-omega_b200/build.py --force > /dev/null
  echo "slots $SL"
  MBS=${MBS:-512,1024,1536} python tools/e2e_probe.py 1024
done
