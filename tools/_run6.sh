timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_gpu_tests.log 2>&1; echo "pytest rc $?"
tail -2 gpurun_out/r2w_gpu_tests.log
for w in s64 s150 s150d3; do
  timeout 300 python tools/profile_solve.py $w 4 > gpurun_out/r2w_prof_${w}.log 2>&1
  grep "rep 3" gpurun_out/r2w_prof_${w}.log
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2w_launches_s64.csv python tools/profile_solve.py s64 2 > gpurun_out/r2w_ncu_s64.log 2>&1
grep -E "k_expand" gpurun_out/r2w_launches_s64.csv | tail -4 | awk -F'","' '{print $5, $NF}'
