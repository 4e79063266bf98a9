#!/bin/bash
# sweeps the chunk length / exploration warm-up of K10 on the GPU box: a synthetic geometric stream (tools/rans_bench.py)
# and the three attribute streams of config 2 (tools/rans_config2.py prints K10 per attribute and the step time)
for cfg in "4096 1024" "4096 512" "2048 1024" "2048 512" "1024 512" "8192 1024" "3072 768" "6144 1024"; do
  set -- $cfg
  export DXO_RANS_DEBUG=1 DXO_RANS_CHUNK=$1 DXO_RANS_WARMUP=$2
  echo "== chunk=$1 warmup=$2"
  python tools/rans_bench.py 3000000 2 2>&1 | tail -2 | tr '\n' ' ' | sed 's/n=3000000 bytes=[0-9]* //'; echo
  python tools/rans_config2.py 2>&1 | grep -E "att [0-9]|K10|step" | sort | uniq -c | sort -rn | head -5
done
