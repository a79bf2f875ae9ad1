set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; tail -c 3000 gpurun_out/bench_r1c.json
python bench.py --steps 20 --warmup 3 --bodies 65536 --no-extras --no-cpu-baseline > gpurun_out/bench_65536_r1c.json 2>> gpurun_out/bench_r1c.err; tail -c 1500 gpurun_out/bench_65536_r1c.json
python tools/prof_gravity.py 65536 4 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nbody_gravity_kernel -s 2 -c 1 -o gpurun_out/prof_r1c python tools/prof_gravity.py 65536 4 > gpurun_out/ncu_r1c.log 2>&1
python bench.py --steps 2 --warmup 3 --bodies 65536 --no-extras --no-cpu-baseline > gpurun_out/plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 2 --warmup 3 --bodies 65536 --no-extras --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -3 gpurun_out/ncu_launches.log
