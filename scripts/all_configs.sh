#!/bin/bash
# throughput of the BASELINE.json configs (and their +bvh forms) on one GPU, at a reduced sample count (tuning aid)
run() { python scripts/prof_render.py --scene $1 --nx $2 --ny $3 --ns $4 --reps 2 --flags 8 2>&1 | tail -1; }
run ch01_random 200 100 100
run ch01_random+bvh 200 100 100
run ch01_random+bvh 1000 500 16
run two_perlin 400 200 64
run cornell_box 500 500 32
run cornell_smoke 500 500 32
run final 1000 1000 4
run final+bvh 1000 1000 8
run final_northstar 1000 1000 16
