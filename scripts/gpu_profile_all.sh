#!/bin/bash
# ncu --set full of every kernel that carries a BASELINE config (one capture each), summarised ON THE BOX (the reports are
# ~7 MB each; gpurun brings back at most 64 MiB): gpurun_out/<P>_<kernel>.txt = scripts/ncu_summary.py, plus the hottest
# source lines.  Then the launch list of the default bench command.
mkdir -p gpurun_out /tmp/reps
P=${1:-r02}
run() {  # target kernel-regex skip tag
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$2" -s "$3" -c 1 -o /tmp/reps/$4 python scripts/prof_target.py $1 > /tmp/reps/ncu_$4.log 2>&1
  python scripts/ncu_summary.py /tmp/reps/$4.ncu-rep > gpurun_out/${P}_$4.txt 2>&1
  python scripts/ncu_source.py /tmp/reps/$4.ncu-rep 25 >> gpurun_out/${P}_$4.txt 2>&1
}
run chamfer "chamfer_fwd_fused" 2 chamfer_fwd
run chamfer "chamfer_finalize3" 2 chamfer_finalize
run chamfer "chamfer_bwd_kernel" 2 chamfer_bwd
run group "group_fused" 2 group_fused
run group "group_bwd" 2 group_bwd
run fps "fps_blk" 2 fps
run fps8k "fps_blk" 2 fps8k
run fps_cluster8 "fps_cluster" 2 fps_cluster8
run crop "crop_split" 2 crop
run knn "knn_warp" 2 knn
run scatter "rows_scatter_add" 2 rows_scatter
run scatter "gather_points_grad" 2 gather_grad
run interp "interp_blend" 2 interp_blend
run interp "interp_bwd_stream" 2 interp_bstream
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/launches_${P}.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-configs > /dev/null 2>&1
cp /tmp/reps/chamfer_fwd.ncu-rep gpurun_out/prof_${P}_chamfer_fwd.ncu-rep 2>/dev/null
ls gpurun_out | wc -l
