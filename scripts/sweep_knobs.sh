#!/bin/bash
# usage: scripts/sweep_knobs.sh "<knobs> <knobs> ..."  — k_render scheduling knobs on the bench scene (tuning aid)
for k in $1; do echo -n "KNOBS=$k  "; RTNW_KNOBS=$k python scripts/prof_render.py --ns 8 --reps 1; done
