timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2y_gpu_tests.log 2>&1; echo "pytest rc $?"
tail -2 gpurun_out/r2y_gpu_tests.log
for w in s64 s150; do
  timeout 300 python tools/profile_solve.py $w 4 > gpurun_out/r2y_prof_${w}.log 2>&1
  grep "rep 3" gpurun_out/r2y_prof_${w}.log
done
python tools/_prof_many.py 1024 512
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2y_launches_many.csv python tools/_prof_many.py 1024 512 > gpurun_out/r2y_many.log 2>&1
timeout 600 python bench.py --workload regions --no-cpu > gpurun_out/r2y_bench_regions.json 2> gpurun_out/r2y_bench_regions.err; echo "bench regions rc $?"
python -c "
import json; d=json.load(open('gpurun_out/r2y_bench_regions.json')); print('regions/s', d['value'])"
