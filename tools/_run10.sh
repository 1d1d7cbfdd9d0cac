python tools/_prof_many.py 1024 512
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2x_launches_many.csv python tools/_prof_many.py 1024 512 > gpurun_out/r2x_many.log 2>&1
tail -3 gpurun_out/r2x_many.log
