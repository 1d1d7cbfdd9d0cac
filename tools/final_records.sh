# every single-GPU record of a round in one call (tests, smoke, bench lines, launch lists, verifier / 32-bit A/B): bash tools/final_records.sh
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2H_gpu_tests.log 2>&1; echo "pytest rc $?"
tail -2 gpurun_out/r2H_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2H_smoke.log 2>&1; tail -1 gpurun_out/r2H_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2H_bench_default.json 2> gpurun_out/r2H_bench_default.err; echo "bench rc $?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2H_bench_reference.json 2> gpurun_out/r2H_bench_reference.err; echo "ref rc $?"
for w in s150 s150d3; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r2H_bench_$w.json 2> gpurun_out/r2H_bench_$w.err; echo "bench $w rc $?"
done
for w in regions stage genome; do
  timeout 600 python bench.py --workload $w --no-cpu > gpurun_out/r2H_bench_$w.json 2> gpurun_out/r2H_bench_$w.err; echo "bench $w rc $?"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2H_launches_s64.csv python tools/profile_solve.py s64 2 > gpurun_out/r2H_ncu_s64.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2H_launches_s150.csv python tools/profile_solve.py s150 2 > gpurun_out/r2H_ncu_s150.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2H_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2H_bench_under_ncu.log 2>&1
for w in s64 s150 s150d3; do
  timeout 300 python tools/profile_solve.py $w 4 > gpurun_out/r2H_prof_${w}.log 2>&1
  NW_PROFILE_NO_VERIFY=1 timeout 300 python tools/profile_solve.py $w 4 > gpurun_out/r2H_prof_${w}_noverify.log 2>&1
  NW_PROFILE_DP32=1 timeout 300 python tools/profile_solve.py $w 4 > gpurun_out/r2H_prof_${w}_dp32.log 2>&1
  grep -h "rep 3" gpurun_out/r2H_prof_${w}.log gpurun_out/r2H_prof_${w}_noverify.log gpurun_out/r2H_prof_${w}_dp32.log | cut -c1-90
done
python - <<'P'
import json
for w in ('default','reference','s150','s150d3','regions','stage','genome'):
    try:
        d=json.load(open('gpurun_out/r2H_bench_%s.json'%w))
        print(w, d['metric'], '%.4g'%d['value'], d['unit'], 'ms/step %.3f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], 'frac', d.get('roofline',{}).get('frac'))
    except Exception as e: print(w,'ERR',e)
P
