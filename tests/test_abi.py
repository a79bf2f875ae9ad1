"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol the headers
declare, and refuses to compute without a GPU (no fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(recs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from recs_b200 import _lib, ecs

    names = _declared("recs_gpu.h")
    assert len(names) >= 30
    for name in names:
        assert hasattr(_lib.lib, name), f"librecs_gpu.so does not export {name}"
    # and the ctypes table binds exactly the declared set
    assert sorted(_lib.GPU_SYMBOLS) == names

    ecs_names = [n for n in _declared("recs_ecs.h") if n not in names and n != "recs_cpu_system_fn"]
    assert len(ecs_names) >= 40
    for name in ecs_names:
        assert hasattr(_lib.lib, name), f"librecs_gpu.so does not export {name}"
    assert sorted(ecs.ECS_SYMBOLS) == sorted(ecs_names)


def test_abi_version_and_default_config():
    from recs_b200 import _lib

    assert _lib.lib.recs_gpu_abi_version() == 1
    cfg = _lib.Config()
    _lib.lib.recs_gpu_default_config(C.byref(cfg))
    # crates/n-body/src/lib.rs:14,118,123; crates/rain-simulation/src/lib.rs:15,59,60,72,76
    assert cfg.nbody_dt == C.c_float(1.0 / 20000.0).value
    assert cfg.nbody_g == C.c_float(6.67e-11).value
    assert cfg.nbody_epsilon == C.c_float(2.0**-23).value
    assert cfg.rain_dt == C.c_float(0.05).value
    assert (cfg.rain_gravity, cfg.rain_terminal, cfg.rain_wind_accel, cfg.rain_wind_deg_per_s) == (
        C.c_float(9.82).value,
        9.0,
        10.0,
        10.0,
    )
    assert cfg.world_size == 1 and cfg.rank == 0


def test_no_cpu_fallback_without_gpu():
    from tests.conftest import HAS_GPU

    if HAS_GPU:
        pytest.skip("a GPU is present")
    from recs_b200 import GpuContext, RecsGpuError

    with pytest.raises(RecsGpuError) as e:
        GpuContext(device=0)
    assert e.value.code == 2  # RECS_ERR_NO_DEVICE


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "recs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "liboracle" not in text and "nbody_oracle" not in text, f


def test_only_tests_bench_and_smoke_touch_the_oracle():
    """The oracle is test infrastructure: besides tests/ only bench.py (cpu_baseline / --impl reference) and
    __graft_entry__.smoke() may import it; the package and the developer tools must not."""
    allowed = {os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")}
    for top in ("recs_b200", "tools", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".sh", ".cu", ".cuh", ".cpp", ".h")):
                    text = open(os.path.join(dirpath, f)).read()
                    assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, os.path.join(dirpath, f)
    for f in os.listdir(ROOT):
        path = os.path.join(ROOT, f)
        if f.endswith(".py") and path not in allowed:
            assert "import oracle" not in open(path).read(), f


def test_reference_arm_never_maps_the_product_library():
    """`bench.py --impl reference` times the CPU port alone: the process must not even load librecs_gpu.so (round-1 did,
    through `from recs_b200 import scenes`)."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, runpy\n"
        "sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '1', '--bodies', '3000']\n"
        "runpy.run_path('bench.py', run_name='__main__')\n"
        "libs = sorted({l.split()[-1] for l in open('/proc/self/maps') if '.so' in l and sys.argv[0] and '/recs_b200/' in l})\n"
        "print('MAPPED', libs)\n"
    )
    proc = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stderr[-2000:]
    assert '"impl": "reference"' in proc.stdout
    assert "MAPPED []" in proc.stdout, proc.stdout[-500:]
