set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 3 --warmup 3 --no-other-configs > gpurun_out/bench_under_ncu_r02.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vit_cosched -s 6 -c 2 -o gpurun_out/prof_r02_cosched -f python bench.py --steps 3 --warmup 3 --no-other-configs > gpurun_out/prof_r02.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_r02_cosched.ncu-rep gpurun_out/launches_r02.csv r02
cp profiles/ncu_summary_r02.md profiles/traffic.json gpurun_out/
rm -f gpurun_out/prof_r02_cosched.ncu-rep
