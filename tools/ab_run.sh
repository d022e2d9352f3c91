#!/bin/bash
# Developer tool: times the libraries built by tools/ab_build.sh against the default build on one box.  usage: tools/ab_run.sh [frames] [all]
echo "== default"; python tools/bm_time.py ${1:-296} $2 2>&1 | tail -9 | awk '{print $1,$2,$3,$4,$7,$8,$NF}'
for f in u96_slam_b200/lib/ab/*.so; do echo "== $f"; U96_LIB=$f python tools/bm_time.py ${1:-296} $2 2>&1 | tail -9 | awk '{print $1,$2,$3,$4,$7,$8,$NF}'; done
