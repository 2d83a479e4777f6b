set -x
cd $GRAFT_REPO_ROOT
export D2R_PROFILE_RANGE=1
# (a) launch list of the bench's timed region (2 steps), kernel by kernel
python bench.py --steps 2 --warmup 3 > gpurun_out/plain_bench.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_ncu_launches_bench_step.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
unset D2R_PROFILE_RANGE
# (b) full captures of the dominant kernels
python tools/gemm_bench.py --case fwd_nt,wgrad_sk8,img_nt --iters 2 > gpurun_out/plain_gemm.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 4 -c 1 -o gpurun_out/r02_prof_gemm_fwd_nt python tools/gemm_bench.py --case fwd_nt,wgrad_sk8,img_nt --iters 2 > gpurun_out/ncu_gemm.log 2>&1
python tools/agg_bench.py --iters 2 > gpurun_out/plain_agg.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:agg_ -s 3 -c 2 -o gpurun_out/r02_prof_agg python tools/agg_bench.py --iters 2 > gpurun_out/ncu_agg.log 2>&1
python tools/attn_bench.py --case 0 --iters 2 > gpurun_out/plain_attn.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_ -s 6 -c 2 -o gpurun_out/r02_prof_attn python tools/attn_bench.py --case 0 --iters 2 > gpurun_out/ncu_attn.log 2>&1
python tools/router_bench.py > gpurun_out/plain_router.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"pool_mean|router_head|saf_" --csv --log-file gpurun_out/r02_ncu_router_kernels.csv python tools/router_bench.py > gpurun_out/ncu_router.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_ncu_*.csv
tail -2 gpurun_out/ncu_bench.log
