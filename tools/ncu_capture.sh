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
for r in step boundary step_fast; do ncu -i $O/${TAG}_$r.ncu-rep --page raw --csv > $O/${TAG}_${r}_raw.csv 2>/dev/null; done
du -sh $O; ls -la $O
