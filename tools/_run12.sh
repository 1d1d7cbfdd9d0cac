for m in 592 1500 3000 6000 12000 1000000; do
  NW_CLASS_MERGE_CTAS=$m python tools/_prof_many.py 1024 512 | head -1 | sed "s/^/merge $m: /"
  NW_CLASS_MERGE_CTAS=$m timeout 600 python bench.py --workload regions --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('   regions/s', d['value'])"
done
