# usage: bash tools/run_scale.sh "8 4"   (on a box with enough GPUs)
for N in $1; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2> gpurun_out/bench_${N}gpu.err | tail -1 > gpurun_out/bench_${N}gpu.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_${N}gpu.json')); print($N, {k:d[k] for k in ('value','ms_per_step','ms_per_step_by_rank')}, d['parity']['acc_max_rel_err_vs_oracle'], d['roofline']['kernel_ms_per_step'], d['e2e']['value'], d['clocks'])"
done
nvidia-smi --query-gpu=index,clocks.sm,power.draw,temperature.gpu --format=csv,noheader | head -8
