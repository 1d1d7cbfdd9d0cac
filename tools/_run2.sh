set -x
ncu --set full --clock-control none --import-source on -k regex:k_score_pair -s 1 -c 1 -o gpurun_out/r2r_pair_s64 -f python tools/profile_solve.py s64 1 > gpurun_out/r2r_ncu_s64.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_score_pair -s 1 -c 1 -o gpurun_out/r2r_pair_s150d3 -f python tools/profile_solve.py s150d3 1 > gpurun_out/r2r_ncu_s150d3.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
