#!/bin/bash
# dev probe: A/B of library builds inside one GPU session
for rep in 1 2; do for v in "$@"; do for wl in config2 config3; do
FEROX_B200_LIB=$PWD/ferox_b200/_build/ab/libferox_$v.so python bench.py --workload $wl --steps 150 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$v $wl', round(d['ms_per_step'],4), {k: round(x,4) for k,x in d['stage_ms_per_step'].items()})"
done; done; done
