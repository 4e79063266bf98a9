#!/bin/bash
# sweeps the speculative rANS parameters on the GPU box: a synthetic geometric stream (tools/rans_bench.py)
# and the three attribute streams of config 2 (tools/step_timing.py prints the device step time)
for cfg in "256 2048 4" "256 4096 4" "256 8192 4" "512 4096 4" "512 8192 4" "1024 8192 4" "1024 16384 4" "256 4096 16" "512 8192 16" "2048 16384 4"; do
  set -- $cfg
  export DXO_RANS_DEBUG=1 DXO_RANS_CHUNK=$1 DXO_RANS_WARMUP=$2 DXO_RANS_ROUNDS=$3
  echo "== chunk=$1 warmup=$2 rounds=$3"
  python tools/rans_bench.py 3000000 2 2>&1 | tail -2 | tr '\n' ' ' | sed 's/n=3000000 bytes=[0-9]* //'; echo
  python tools/rans_config2.py 2>&1 | grep -E "att [0-9]|K10|step" | sort | uniq -c | sort -rn | head -8
done
