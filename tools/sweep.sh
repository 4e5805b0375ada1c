#!/bin/bash
# usage: tools/sweep.sh "ENV1=a ENV2=b" "ENV1=c" ...   -- runs the 0.1-scale resident bench once per environment set
for envs in "$@"; do
  out=$(env $envs python bench.py --scale ${SCALE:-0.1} --steps 3 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1)
  echo "$envs :: $(echo "$out" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['engine']['stage_ms'].items()})" 2>&1 | tail -1)"
done
