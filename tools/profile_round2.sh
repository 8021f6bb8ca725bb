# usage: TAG=r2x bash tools/profile_round2.sh
# ncu evidence of round 2 (B200_PROFILING.md recipe): launch list of the default step, then one --set full capture per kernel.
CMD="python bench.py --reads 2000000 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --cpu-sample 20000"
$CMD > gpurun_out/plain_${TAG:-r2p}.log 2>&1 || { echo plain failed; tail -3 gpurun_out/plain_${TAG:-r2p}.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG:-r2p}.csv $CMD > gpurun_out/ncu1_${TAG:-r2p}.log 2>&1; echo launches rc=$?
ncu --set full --clock-control none --import-source on -k regex:prefilter_kernel -s 3 -c 1 -o gpurun_out/prof_${TAG:-r2p}_prefilter $CMD > gpurun_out/ncu2_${TAG:-r2p}.log 2>&1; echo prefilter rc=$?
ncu --set full --clock-control none --import-source on -k regex:score_tasks_kernel -s 3 -c 1 -o gpurun_out/prof_${TAG:-r2p}_tasks $CMD > gpurun_out/ncu3_${TAG:-r2p}.log 2>&1; echo tasks rc=$?
ncu --set full --clock-control none --import-source on -k regex:finalize_best_kernel -s 3 -c 1 -o gpurun_out/prof_${TAG:-r2p}_finalize $CMD > gpurun_out/ncu4_${TAG:-r2p}.log 2>&1; echo finalize rc=$?
ls -la gpurun_out/*${TAG:-r2p}*
