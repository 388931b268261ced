# usage: bash tools/gpu_balance.sh -- SM activity balance of every k_pass launch of one 512-pair step
mkdir -p gpurun_out
timeout 600 ncu --metrics sm__cycles_active.avg,sm__cycles_active.max,sm__cycles_active.min,sm__cycles_elapsed.max,gpu__time_duration.sum --clock-control none -k regex:k_pass -c 108 --csv --log-file gpurun_out/balance_all.csv python bench.py --one-step > gpurun_out/b_ncu_bal.log 2>&1
tail -2 gpurun_out/b_ncu_bal.log
