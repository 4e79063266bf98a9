# bench.py's resident-session arm under different host-side schemes: usage: bash tools/session_modes_sweep.sh "<cpu lists>" "<session counts>"
run() { # label, cpus, env..., (ARGS from the environment)
  label=$1; cpus=$2; shift 2
  env "$@" taskset -c $cpus python bench.py --no-config4 --no-cpu-baseline $ARGS 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['config']; print('$label', 'cpus $cpus', round(d['value']), c['units_per_step_per_gpu'][:10], c['host_waits'], c['graph_replay'], c['side_streams'][:12])"
}
for cpus in ${1:-0-3 0-7}; do
ARGS="" run default $cpus X=1
for s in ${2:-8 16}; do
ARGS="--sessions $s" run spin_helpers $cpus X=1
ARGS="--sessions $s --graph-replay off" run spin_inline_nograph $cpus DXO_SIDE_INLINE=1
ARGS="--sessions $s --graph-replay on" run spin_inline_graph $cpus DXO_SIDE_INLINE=1
ARGS="--sessions $s" run block_inline $cpus DXO_BLOCKING_WAIT=1 DXO_SIDE_INLINE=1
done
done
