"""Streaming systems only (simple_iter 10 M entities, fused rain tick 4 Mi drops) for ncu / quick timing."""
import json
import sys

sys.path.insert(0, ".")
import bench  # noqa: E402
from recs_b200 import GpuContext  # noqa: E402

with GpuContext(device=0, use_cuda_graph=False) as ctx:
    print(json.dumps(bench.streaming_extras(ctx, bench.measured_traffic())))
