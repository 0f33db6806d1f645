#!/bin/bash
# dev helper: pipelined-sweep tuning at the multi-GPU shard shapes on ONE GPU (rows lag batch [extra env])
out=gpurun_out/r2_tune.jsonl; : > $out
run() { env XCOLUMNS_B200_LAG=$2 $4 python bench.py --steps 20 --warmup 5 --rows $1 --batch $3 --no-e2e --no-cpu --no-secondary 2>/dev/null | grep '^{' | python -c "
import json,sys
b=json.loads(sys.stdin.read()); r=b['roofline']
print(json.dumps({'rows':$1,'lag':$2,'batch':$3,'env':'$4','ms':round(b['ms_per_step'],4),'Minst_s':round(b['value']/1e6,1),'whole':round(r['whole_step']['frac'],3),'commits':b['config']['commits_per_sweep'],'u':b['utility_after_timed_sweeps']}))" >> $out; }
run 38375 1 0
run 38375 2 0
run 38375 3 0
run 76750 1 0
run 76750 2 0
run 76750 2 9594
run 76750 3 0
run 153500 2 0
run 153500 2 19188
run 307000 1 0
run 307000 2 0
cat $out
