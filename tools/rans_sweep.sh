#!/bin/bash
# sweeps the speculative rANS parameters on the GPU box (prints ms per 3M-symbol stream)
for cfg in "4096 8192 3" "4096 8192 6" "2048 8192 6" "2048 4096 8" "2048 16384 4" "1024 8192 8" "1024 4096 12" "8192 16384 3" "4096 16384 4" "2048 12288 6" "1024 16384 6"; do
  set -- $cfg
  echo "chunk=$1 warmup=$2 rounds=$3: $(DXO_RANS_DEBUG=1 DXO_RANS_CHUNK=$1 DXO_RANS_WARMUP=$2 DXO_RANS_ROUNDS=$3 python tools/rans_bench.py 3000000 2 2>&1 | tail -2 | tr '\n' ' ' | sed 's/n=3000000 bytes=[0-9]* //')"
done
