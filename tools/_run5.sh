timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2v_gpu_tests.log 2>&1; echo "pytest rc $?"
tail -2 gpurun_out/r2v_gpu_tests.log
timeout 600 python bench.py > gpurun_out/r2v_bench_default.json 2> gpurun_out/r2v_bench_default.err; echo "bench rc $?"
for w in s150 s150d3 regions stage genome; do
  timeout 600 python bench.py --workload $w --no-cpu > gpurun_out/r2v_bench_$w.json 2> gpurun_out/r2v_bench_$w.err; echo "bench $w rc $?"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2v_launches_s64.csv python tools/profile_solve.py s64 2 > gpurun_out/r2v_ncu_s64.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2v_launches_s150.csv python tools/profile_solve.py s150 2 > gpurun_out/r2v_ncu_s150.log 2>&1
python - <<'P'
import json
for w in ('default','s150','s150d3','regions','stage','genome'):
    try:
        d=json.load(open('gpurun_out/r2v_bench_%s.json'%w))
        print(w, d['metric'], '%.4g'%d['value'], d['unit'], 'ms/step %.3f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], 'frac', d.get('roofline',{}).get('frac'))
    except Exception as e: print(w,'ERR',e)
P
