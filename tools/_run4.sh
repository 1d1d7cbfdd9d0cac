timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2u_gpu_tests.log 2>&1; echo "pytest rc $?"
tail -2 gpurun_out/r2u_gpu_tests.log
for w in s64 s150 s150d3; do
  timeout 300 python tools/profile_solve.py $w 4 > gpurun_out/r2u_prof_${w}.log 2>&1
  grep "rep 3" gpurun_out/r2u_prof_${w}.log
done
ncu --set full --clock-control none --import-source on -k regex:k_score_pair -s 1 -c 1 -o gpurun_out/r2u_pair_s64 -f python tools/profile_solve.py s64 1 > gpurun_out/r2u_ncu_s64.log 2>&1
