#!/bin/bash
# A/B of the round-2 switches on the cfg2 prefill step (ms_per_step of bench.py, device-resident, 30 steps), one process each.
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-roofline --no-library-baseline --no-train-record --no-decode-record --ttft-iters 30"
run() { echo -n "$1: "; env $1 timeout -s KILL 300 $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['launches_per_step'], d['clocks']['sm_mhz'])"; }
for cfg in "$@"; do run "$cfg"; done
