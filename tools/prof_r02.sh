mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "fps_mbarrier or fps_bit" 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python bench.py --short --steps 1 --warmup 1 --pairs-per-step 2 --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1
wc -l gpurun_out/r02_launches.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_sd -s 8 -c 8 -o gpurun_out/r02_sd_v2 python tools/conv_bench.py 9000 1 2>&1 | tail -3
timeout 200 python tools/conv_bench.py 9000 5 > gpurun_out/r02_conv_layers_K9000.txt 2>&1; cat gpurun_out/r02_conv_layers_K9000.txt | tail -10
BX_FPS_SYNC=1 timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_racecheck_fpssync.txt 2>&1; tail -5 gpurun_out/r02_sanitizer_racecheck_fpssync.txt
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; tail -4 gpurun_out/r02_sanitizer_memcheck.txt
