#!/bin/bash
# usage: scripts/ab_run.sh name [name ...] — time each ab/<name>.so on the bench scene (tuning aid)
for n in "$@"; do echo -n "$n: "; RTNW_LIB=ab/$n.so timeout 120 python scripts/prof_render.py --ns 16 --reps 2 2>&1 | tail -2 | tr '\n' ' '; echo; done
