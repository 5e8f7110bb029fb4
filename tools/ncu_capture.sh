# ncu evidence for profiles/ (run under gpurun on ONE B200): launch list of the bench + full captures of the
# step kernel (exact, fast) and of the boundary kernels.  Output < 64 MiB so that gpurun brings it back.
TAG=${1:-r2a}
O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-sweep --no-fast --no-djds"
$B > $O/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${TAG}_launches.csv $B > $O/${TAG}_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:l96_rk4_stream -s 160 -c 1 -o $O/${TAG}_step $B > $O/${TAG}_ncu_step.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:gram_mma|apply_mma|dilation|small_single|tree_level|extract_col0|sum_partials' -s 48 -c 7 -o $O/${TAG}_boundary $B > $O/${TAG}_ncu_bnd.log 2>&1
F="$B --math fast"
$F > $O/${TAG}_plain_fast.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:l96_rk4_stream -s 160 -c 1 -o $O/${TAG}_step_fast $F > $O/${TAG}_ncu_stepf.log 2>&1
# the primal marches (run-up, recorded solver call): outside the bench's timed region, captured from the probe
ncu --set full --clock-control none --import-source on -k regex:dilation_fused -s 8 -c 1 -o $O/${TAG}_runup python tools/runup_probe.py > $O/${TAG}_ncu_runup.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dilation_fused -s 140 -c 1 -o $O/${TAG}_recorded python tools/runup_probe.py > $O/${TAG}_ncu_rec.log 2>&1
# gpurun brings back at most 64 MiB: keep the raw-page exports (what tools/summarize_profiles.py reads), drop the reports
for r in step boundary step_fast runup recorded; do ncu -i $O/${TAG}_$r.ncu-rep --page raw --csv > $O/${TAG}_${r}_raw.csv 2>/dev/null && rm -f $O/${TAG}_$r.ncu-rep; done
du -sh $O; ls -la $O
