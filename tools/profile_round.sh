#!/bin/bash
# Run on the GPU box (gpurun): launch list + full-set captures of the score-path kernels of bench.py.
#   tools/profile_round.sh <tag>       e.g. r01
# Outputs land in gpurun_out/ ; tools/extract_profiles.py turns them into the tracked summaries under profiles/.
TAG=${1:-r01}
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_list_${TAG}.log 2>&1
$CMD > gpurun_out/plain_${TAG}b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_mdch_sel|k_classify_staged|k_pf_update|k_frame_prep" -s 60 -c 5 -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -n 2 gpurun_out/ncu_full_${TAG}.log
# UpdateClassifier kernels (full frames leg)
ncu --set full --clock-control none --import-source on -k regex:"k_milboost_select_batch|k_mdch_train_main|k_weak_update_batch|k_haar_matrix_batch|k_train_gen_default" -s 16 -c 5 -o gpurun_out/prof_${TAG}_upd $CMD > gpurun_out/ncu_full_${TAG}_upd.log 2>&1
tail -n 2 gpurun_out/ncu_full_${TAG}_upd.log
