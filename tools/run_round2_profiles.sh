# Round-2 GPU pass (one B200): tests, bench, ncu launch list of the bench command, ncu --set full of the dominant kernels.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_gputests.log
python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
B="python bench.py --steps 2 --warmup 3 --no-extras --no-configs --no-cpu-baseline"
$B > gpurun_out/r02_plain_launches.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_1M.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
B2="python bench.py --bodies 65536 --steps 2 --warmup 3 --no-extras --no-configs --no-cpu-baseline"
$B2 > gpurun_out/r02_plain_launches2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_65536.csv $B2 > gpurun_out/r02_ncu_launches2.log 2>&1
G="python tools/prof_gravity.py 1048576 2"
$G > gpurun_out/r02_plain_g1m.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nbody_gravity_kernel -s 1 -c 1 -o gpurun_out/r02_prof_gravity_1M $G > gpurun_out/r02_ncu_g1m.log 2>&1
G2="python tools/prof_gravity.py 65536 4"
$G2 > gpurun_out/r02_plain_g64k.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nbody_gravity_kernel -s 2 -c 1 -o gpurun_out/r02_prof_gravity_65536 $G2 > gpurun_out/r02_ncu_g64k.log 2>&1
S="python tools/prof_streaming_once.py"
$S > gpurun_out/r02_plain_stream.log 2>&1 && ncu --set full --clock-control none -k regex:"axpy_rn_vec_kernel|rain_tma_kernel|recs_generic_system" -c 12 -o gpurun_out/r02_prof_streaming $S > gpurun_out/r02_ncu_stream.log 2>&1
tail -3 gpurun_out/r02_gputests.log
