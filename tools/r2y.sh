for i in 1 2 3 4; do
  python -m pytest tests/test_gpu_stream.py -x -q > gpurun_out/r2y_$i.log 2>&1 || { echo "run $i failed"; grep -n "Error\|assert\|^E " gpurun_out/r2y_$i.log | head -30; }
  tail -1 gpurun_out/r2y_$i.log
done
