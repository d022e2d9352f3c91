#!/bin/bash
# Collect the round's ncu evidence in one gpurun call (developer tool).  usage: tools/collect_profiles.sh TAG
# Writes gpurun_out/TAG_*.ncu-rep, the bench launch list and bench lines; summarise here with tools/ncu_summary.py.
TAG=${1:-r02}
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2>> $O/${TAG}_bench_n1.err
python tools/bm_time.py 296 all > $O/${TAG}_bm_time.log 2>&1
U96_BM_FUSED=0 python tools/bm_time.py 296 all > $O/${TAG}_bm_time_fast.log 2>&1
python tools/stage_time.py > $O/${TAG}_stage_time.log 2>&1
python tools/post_time.py 296 > $O/${TAG}_post_time.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 4 --warmup 3 --reps 1 --no-cpu-baseline > /dev/null 2>&1
$NCU -k regex:k_bm_fused -s 2 -c 1 -o $O/${TAG}_bm64   python tools/bm_one.py 640 480 64 21 0 296 4 > /dev/null 2>&1
U96_BM_FUSED=0 $NCU -k regex:k_bm_fast -s 2 -c 1 -o $O/${TAG}_bm64fast python tools/bm_one.py 640 480 64 21 0 296 4 > /dev/null 2>&1
$NCU -k regex:k_bm_fused -s 2 -c 1 -o $O/${TAG}_bm64b15 python tools/bm_one.py 640 480 64 15 0 296 4 > /dev/null 2>&1
$NCU -k regex:k_bm_fused -s 2 -c 1 -o $O/${TAG}_bmcv64 python tools/bm_one.py 640 480 64 21 1 296 4 > /dev/null 2>&1
$NCU -k regex:k_bm_fused -s 2 -c 1 -o $O/${TAG}_bm128  python tools/bm_one.py 1242 375 128 15 0 148 4 > /dev/null 2>&1
$NCU -k regex:k_bm_fused -s 2 -c 1 -o $O/${TAG}_bm256  python tools/bm_one.py 1920 1080 256 21 0 37 4 > /dev/null 2>&1
$NCU -k regex:k_rect_remap_tma -s 3 -c 1 -o $O/${TAG}_rect   python tools/stage_time.py 296 > /dev/null 2>&1
$NCU -k regex:k_xsobel -s 3 -c 1 -o $O/${TAG}_xsobel python tools/stage_time.py 296 > /dev/null 2>&1
$NCU -k regex:k_gftt_eig -s 3 -c 1 -o $O/${TAG}_gftt python tools/stage_time.py 296 > /dev/null 2>&1
ls -la $O | grep ${TAG}_ | tail -30
