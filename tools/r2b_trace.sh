python tools/copy_sweep.py > gpurun_out/r2b_copy_sweep.log 2>&1
ORPH_AB_GEN_ONLY=1 python tools/e2e_ab.py 10000000 packed > /dev/null 2>&1  # generates the cache
for cfg in "19:" "21:" "19:kernels,h2d" "19:h2d,d2h" "19:kernels,d2h" "21:kernels,h2d" "21:h2d,d2h" "21:kernels,d2h"; do
  lg=${cfg%%:*}; sk=${cfg##*:}
  echo "######## chunk_log2=$lg skip=$sk" >> gpurun_out/r2b_trace.log
  ORPHEUM_B200_TRACE=1 ORPHEUM_B200_CHUNK_LOG2=$lg ORPHEUM_B200_AB_SKIP=$sk python tools/e2e_ab.py run packed 2>&1 | tail -40 >> gpurun_out/r2b_trace.log
done
