# Round measurement on one B200: launch list of the default bench command, `ncu --set full` captures of the dominant kernels.
# Outputs under gpurun_out/ (copy the summaries into profiles/).
R=${ROUND:-r2}
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-per-config > gpurun_out/plain_ll.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 600 --csv --log-file gpurun_out/launches_${R}_tc.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-per-config > gpurun_out/ncu_ll.log 2>&1
cap() {  # config kernel-regex out-name [env]
  env $4 ncu --set full --clock-control none --graph-profiling node --import-source on -k regex:$2 -s 4 -c 1 -f -o gpurun_out/$3 python bench.py --config $1 --steps 2 --warmup 3 --no-cpu-baseline --no-per-config > gpurun_out/$3_ncu.log 2>&1
}
cap c2 rollout_gnn_tc prof_${R}_tc_c2 A=1
cap c3 rollout_gnn_tc prof_${R}_tc_c3 A=1
cap c4 rollout_gnn_tc prof_${R}_tc_c4 A=1
cap c5 rollout_mlp_tc prof_${R}_tc_c5 A=1
cap c2 cost_compact prof_${R}_cost_c2 A=1
cap c5 cost_compact prof_${R}_cost_c5 A=1
cap c2 rollout_gnn_tc prof_${R}_tc_c2_tail CEEUS_COST_TAIL=1
cap c2 rollout_gnn_tc prof_${R}_tc_c2_tail_nodiscard "CEEUS_COST_TAIL=1 CEEUS_NO_DISCARD=1"
ls -la gpurun_out/*.ncu-rep
# the reports are too large to travel together: summarise on the box, keep only the summaries (+ the c2 rollout report)
for f in gpurun_out/prof_${R}_*.ncu-rep; do python profiles/summarize_ncu.py $f > ${f%.ncu-rep}.summary.txt 2>/dev/null; done
for f in gpurun_out/prof_${R}_*.ncu-rep; do case $f in *prof_${R}_tc_c2.ncu-rep) ;; *) rm -f $f;; esac; done
rm -f gpurun_out/*_ncu.log
