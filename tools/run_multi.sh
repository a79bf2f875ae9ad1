# usage: bash tools/run_multi.sh N  -> tests + bench on N GPUs
N=${1:-2}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/tools/multi_gpu_check.py 100000 2 2>&1 | grep -E "MULTI_GPU_CHECK|Error|error" | head
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 5 --warmup 3 --no-extras 2> gpurun_out/bench_${N}gpu.err | tail -1 > gpurun_out/bench_${N}gpu.json
tail -c 2500 gpurun_out/bench_${N}gpu.json
