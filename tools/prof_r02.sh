set -x
cd $GRAFT_REPO_ROOT
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 3 --warmup 3 --no-other-configs > gpurun_out/bench_under_ncu_r02.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vit_cosched -s 6 -c 2 -o gpurun_out/prof_r02_cosched python bench.py --steps 3 --warmup 3 --no-other-configs > gpurun_out/prof_r02.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"xca_tc|attn_wide|attn_core|lepe_tiled|layernorm|sr_conv|dwconv3|class_attn_core" -c 14 -o gpurun_out/prof_r02_variants python tools/launch_breakdown.py > gpurun_out/prof_r02v.log 2>&1
ls -la gpurun_out/*.ncu-rep
