CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,smsp__inst_executed.sum --clock-control none -k regex:"k_mdch_train|k_weak_update_batch|k_haar_matrix_batch|k_build_colormap" -s 20 -c 24 --csv --log-file gpurun_out/mtr_list.csv $CMD > gpurun_out/ncu_mtr2.log 2>&1
tail -n 1 gpurun_out/ncu_mtr2.log
