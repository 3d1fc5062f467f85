# ncu --set full captures (with source counters) of one full-size launch of each fused kernel; reports go to gpurun_out/ (~25 MB each)
set -x
cd $GRAFT_REPO_ROOT
ncu --set full --import-source on --clock-control none -k regex:fused_forward -s 1 -c 1 -f -o gpurun_out/${1:-x}_fwd python bench.py --steps 1 --warmup 0 --no-extras --no-cpu-baseline > gpurun_out/${1:-x}_ncu_fwd.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:fused_backward -s 1 -c 1 -f -o gpurun_out/${1:-x}_bwd python bench.py --steps 1 --warmup 0 --no-extras --no-cpu-baseline > gpurun_out/${1:-x}_ncu_bwd.log 2>&1
ls -la gpurun_out/*.ncu-rep
