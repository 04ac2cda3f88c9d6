# end-of-round evidence run (one GPU): bench line, ncu launch list of the bench command, --set full captures of the headline
# kernel and of every variant kernel, summarised ON THE BOX (the .ncu-rep files exceed what gpurun copies back)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 300 --warmup 20 > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 3 --warmup 3 --no-other-configs > gpurun_out/bench_under_ncu_r02.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vit_cosched -s 6 -c 2 -o gpurun_out/prof_r02_cosched -f python bench.py --steps 3 --warmup 3 --no-other-configs > gpurun_out/prof_r02.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_r02_cosched.ncu-rep gpurun_out/launches_r02.csv r02
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pa::" --csv --log-file gpurun_out/launch_breakdown_r02.csv python tools/launch_breakdown.py > gpurun_out/lb_r02.log 2>&1
ncu --set full --clock-control none -k regex:"xca_tc|attn_wide|attn_win|attn_single_slot|attn_proj|attn_core|layernorm|sr_conv|sr_patchify|dwconv3|class_attn_core" -c 30 -o gpurun_out/prof_r02_variants -f python tools/launch_breakdown.py > gpurun_out/prof_r02v.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_r02_variants.ncu-rep - r02_variants --no-traffic
cp profiles/ncu_summary_r02.md profiles/ncu_summary_r02_variants.md profiles/traffic.json gpurun_out/
rm -f gpurun_out/prof_r02_variants.ncu-rep
ls -la gpurun_out/
