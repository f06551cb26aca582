for cfg in "1 0" "2 0" "1 1"; do set -- $cfg; python bench.py --qubits 28 --depth 40 --steps 2 --warmup 1 --fuse $1 --tile-debug $2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('fuse $1 debug $2: value %.1f ms/step %.1f  tile_fwd %.1f tile_bwd %.1f launches %d' % (d['value'], d['ms_per_step'], d['profile_ms'].get('tile_fwd',0), d['profile_ms'].get('tile_bwd',0), d['gpu_launches']))"; done
