# Round-end evidence run (one GPU): bench lines, launch list + full ncu captures of the bench command's kernels, sanitizer.
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python tools/profile_step.py --steps 6 > gpurun_out/stage_times.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'quantise_count|place_kernel|pillar_prep|scan_cells|tile_sums' -c 5 -o gpurun_out/vox_r02 -f python tools/profile_step.py --steps 1 > gpurun_out/ncu_vox.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'pfn_slot|canvas_v8|pfn_finish' -c 3 -o gpurun_out/pfn_canvas_r02 -f python tools/profile_step.py --steps 1 > gpurun_out/ncu_pfn.log 2>&1
python tools/config_sweep.py --iters 20 > gpurun_out/config_sweep.json 2> gpurun_out/config_sweep.err
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_next.py tests/test_gpu_modar.py -x -q -k "not reference_kernel and not reference_composition" > gpurun_out/memcheck_next.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -k "agree and (777 or 1000 or 32768)" > gpurun_out/memcheck_methods.log 2>&1
timeout 600 compute-sanitizer --tool initcheck python -m pytest tests/test_gpu_parity.py -x -q -k "agree and (777 or 1000)" > gpurun_out/initcheck_methods.log 2>&1
tail -n 3 gpurun_out/memcheck_next.log; tail -n 3 gpurun_out/memcheck_methods.log; tail -n 3 gpurun_out/initcheck_methods.log
cut -c1-300 gpurun_out/bench.json
