# Evidence pass on one GPU: `bash tools/ncu_round.sh <tag>` (run from the repo root on the GPU box).
#   1. bench line without a profiler                          -> gpurun_out/<tag>_bench_1gpu.json
#   2. ncu launch list of a short run                         -> gpurun_out/<tag>_launches.csv
#   3. ncu --set full of the largest launch of the three hot kernels (reports stay in /tmp: too large to bring back;
#      their raw pages are summarised here)                    -> gpurun_out/<tag>_{fused_fwd,fused_bwd,gemm}_ncu.txt
tag=${1:-rXX}
cd ${GRAFT_REPO_ROOT:-.}
python bench.py > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/${tag}_ncu_launches.log 2>&1
for k in fused_forward fused_backward; do
  ncu --set full --import-source on --clock-control none -k regex:$k -s 1 -c 1 -f -o /tmp/${tag}_$k \
      python bench.py --steps 1 --warmup 0 --no-extras --no-cpu-baseline > gpurun_out/${tag}_ncu_$k.log 2>&1
done
# parameter contraction: the batched lower-only NT product of the first full-size layer (grid > 1000)
ncu --set full --import-source on --clock-control none -k regex:gemm_kernel -c 90 -f -o /tmp/${tag}_gemm \
    python bench.py --steps 1 --warmup 0 --no-extras --no-cpu-baseline > gpurun_out/${tag}_ncu_gemm.log 2>&1
python tools/ncu_summary.py raw /tmp/${tag}_fused_forward.ncu-rep gpurun_out/${tag}_fused_fwd_ncu.txt > /dev/null
python tools/ncu_summary.py raw /tmp/${tag}_fused_backward.ncu-rep gpurun_out/${tag}_fused_bwd_ncu.txt > /dev/null
python tools/ncu_summary.py raw /tmp/${tag}_gemm.ncu-rep gpurun_out/${tag}_gemm_ncu.txt > /dev/null
python tools/ncu_lines.py /tmp/${tag}_fused_forward.ncu-rep 30 > gpurun_out/${tag}_fused_fwd_lines.txt 2>&1
python tools/ncu_lines.py /tmp/${tag}_fused_backward.ncu-rep 30 > gpurun_out/${tag}_fused_bwd_lines.txt 2>&1
ls -la gpurun_out/${tag}_*
