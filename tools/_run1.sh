set -x
mkdir -p gpurun_out
timeout 120 tools/ubench/dpx_rates > gpurun_out/r2q_dpx_rates.json 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_gpu_tests.log 2>&1; echo "pytest rc $?"
tail -5 gpurun_out/r2q_gpu_tests.log
for w in s64 s150 s150d3; do
  timeout 300 python tools/profile_solve.py $w 4 > gpurun_out/r2q_prof_${w}_pair.log 2>&1
  NW_PROFILE_DP32=1 timeout 300 python tools/profile_solve.py $w 4 > gpurun_out/r2q_prof_${w}_dp32.log 2>&1
  grep "rep 3" gpurun_out/r2q_prof_${w}_pair.log gpurun_out/r2q_prof_${w}_dp32.log
done
timeout 600 python bench.py > gpurun_out/r2q_bench_default.json 2> gpurun_out/r2q_bench_default.err; echo "bench rc $?"
cat gpurun_out/r2q_dpx_rates.json
