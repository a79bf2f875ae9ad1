"""the Rust host shim as files (integration/rust), checkable without rustc.

  * integration/rust/recs-gpu-sys/src/lib.rs is generated from include/recs_gpu.h (tools/gen_rust_sys.py): it must be in
    step with the header, and -- parsed independently here -- agree with the ctypes table the tests drive the library
    through (recs_b200/_lib.py): every function, argument count, argument / return width, `#[repr(C)]` struct layout
    and constant.
  * integration/rust/ecs-gpu/ecs_gpu_system.rs (`impl System for GpuSystem`, `with_gpu()` decorator; reference
    crates/ecs/src/systems/mod.rs:48-62, crates/gfx-plugin/src/lib.rs:214-377) may only use symbols, constants and
    struct fields the -sys crate declares, with the declared arity.
"""
import ctypes as C
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SYS_RS = os.path.join(ROOT, "integration", "rust", "recs-gpu-sys", "src", "lib.rs")
SHIM_RS = os.path.join(ROOT, "integration", "rust", "ecs-gpu", "ecs_gpu_system.rs")

RUST_SCALARS = {"i32": (4, "int"), "u32": (4, "int"), "u64": (8, "int"), "u8": (1, "int"), "usize": (8, "int"), "f32": (4, "float")}


def _rust_functions(text):
    block = text[text.index('extern "C" {'):]
    block = block[: block.index("\n}\n") + 2]
    funcs = {}
    for name, args, ret in re.findall(r"pub fn (\w+)\(\s*(.*?)\s*\)\s*(?:->\s*([^;]+?))?\s*;", block, flags=re.S):
        params = [a.strip() for a in re.split(r",\s*(?![^\[]*\])", args) if a.strip()]
        funcs[name] = ([p.split(":", 1)[1].strip() for p in params], (ret or "()").strip())
    return funcs


def _rust_structs(text):
    structs = {}
    for name, body in re.findall(r"#\[repr\(C\)\]\s*(?:#\[derive\([^)]*\)\]\s*)?pub struct (\w+) \{(.*?)\n\}", text, flags=re.S):
        fields = re.findall(r"pub (\w+): ([^,\n]+),", body)
        structs[name] = fields
    return structs


def _rust_consts(text):
    out = {}
    for name, ty, value in re.findall(r"pub const (\w+): (\w+) = ([^;]+);", text):
        out[name] = (ty, int(value, 0))
    return out


def _width(rust_type, structs, consts):
    """(size, alignment, class) of a Rust FFI type under the C ABI of x86-64 / aarch64 Linux."""
    t = rust_type.strip()
    if t.startswith("*"):
        return 8, 8, "pointer"
    if t in RUST_SCALARS:
        size, cls = RUST_SCALARS[t]
        return size, size, cls
    m = re.match(r"\[(.+); (\w+)\]", t)
    if m:
        size, align, _ = _width(m.group(1), structs, consts)
        n = consts[m.group(2)][1] if m.group(2) in consts else int(m.group(2))
        return size * n, align, "array"
    size, align, _ = _layout(structs[t], structs, consts)
    return size, align, "struct"


def _layout(fields, structs, consts):
    offset, max_align, offsets = 0, 1, []
    for _, ty in fields:
        size, align, _ = _width(ty, structs, consts)
        offset = (offset + align - 1) // align * align
        offsets.append(offset)
        offset += size
        max_align = max(max_align, align)
    return (offset + max_align - 1) // max_align * max_align, max_align, offsets


def _ctypes_class(ct):
    if ct is None:
        return 0, "void"
    if isinstance(ct, type) and issubclass(ct, (C._Pointer, C.c_char_p, C.c_void_p)) or ct in (C.c_char_p, C.c_void_p):
        return 8, "pointer"
    if ct in (C.c_float,):
        return 4, "float"
    return C.sizeof(ct), "int"


def test_sys_crate_is_in_step_with_the_header():
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_sys.py"), "--check"])
    assert proc.returncode == 0, "include/recs_gpu.h changed: run `python tools/gen_rust_sys.py`"


def test_every_function_argument_and_width_agrees_with_the_ctypes_table():
    from recs_b200 import _lib

    text = open(SYS_RS).read()
    funcs, structs, consts = _rust_functions(text), _rust_structs(text), _rust_consts(text)
    assert set(funcs) == set(_lib.GPU_SYMBOLS), set(funcs) ^ set(_lib.GPU_SYMBOLS)
    header = open(os.path.join(ROOT, "include", "recs_gpu.h")).read()
    declared = set(re.findall(r"\b(recs_gpu_\w+)\s*\(", re.sub(r"/\*.*?\*/", " ", header, flags=re.S)))
    assert set(funcs) == declared
    for name, (restype, argtypes) in _lib.GPU_SYMBOLS.items():
        rust_args, rust_ret = funcs[name]
        assert len(rust_args) == len(argtypes), name
        for i, (ra, ca) in enumerate(zip(rust_args, argtypes)):
            size, _, cls = _width(ra, structs, consts)
            csize, ccls = _ctypes_class(ca)
            assert (size, cls) == (csize, ccls), f"{name} argument {i}: Rust {ra} vs ctypes {ca}"
        if rust_ret == "()":
            assert restype is None, name
        else:
            size, _, cls = _width(rust_ret, structs, consts)
            assert (size, cls) == _ctypes_class(restype), name


def test_repr_c_struct_layouts_match_the_c_layout():
    from recs_b200 import _lib

    text = open(SYS_RS).read()
    structs, consts = _rust_structs(text), _rust_consts(text)
    pairs = {"recs_gpu_config": _lib.Config, "recs_gpu_access": _lib.Access, "recs_gpu_system_desc": _lib.SystemDesc,
             "recs_gpu_kernel_param": _lib.KernelParam, "recs_gpu_timing": _lib.Timing}
    assert set(pairs) == set(structs) - {"recs_gpu_ctx"}
    for name, cstruct in pairs.items():
        size, _, offsets = _layout(structs[name], structs, consts)
        assert size == C.sizeof(cstruct), name
        assert [f for f, _ in structs[name]] == [f for f, _ in cstruct._fields_], name
        assert offsets == [getattr(cstruct, f).offset for f, _ in cstruct._fields_], name


def test_constants_match():
    from recs_b200 import _lib

    consts = _rust_consts(open(SYS_RS).read())
    for code, name in _lib.ERR_NAMES.items():
        assert consts[name][1] == code
    for attr in dir(_lib):
        if attr.startswith("SYS_"):
            assert consts["RECS_" + attr][1] == getattr(_lib, attr), attr
        if attr.startswith("KPARAM_"):
            assert consts["RECS_" + attr][1] == getattr(_lib, attr), attr
    assert consts["RECS_GPU_ENTITY_COMPONENT"] == ("u64", _lib.ENTITY_COMPONENT)
    assert consts["RECS_GPU_NCCL_ID_BYTES"][1] == _lib.NCCL_ID_BYTES and consts["RECS_GPU_IPC_HANDLE_BYTES"][1] == _lib.IPC_HANDLE_BYTES
    assert consts["RECS_GPU_MAX_ACCESSES"][1] == _lib.MAX_ACCESSES
    assert consts["RECS_GPU_ABI_VERSION"][1] == _lib.lib.recs_gpu_abi_version()


def _call_arity(text, start):
    """Number of top-level arguments of the call whose '(' is at text[start]."""
    depth, args, i, seen = 0, 0, start, False
    while True:
        ch = text[i]
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
            if depth == 0:
                return args + (1 if seen else 0)
        elif ch == "," and depth == 1:
            args += 1
            seen = False
        elif not ch.isspace() and depth >= 1:
            seen = True
        i += 1


def test_shim_uses_only_what_the_sys_crate_declares():
    sys_text, shim = open(SYS_RS).read(), open(SHIM_RS).read()
    funcs, structs, consts = _rust_functions(sys_text), _rust_structs(sys_text), _rust_consts(sys_text)
    code = re.sub(r"//.*", "", shim)  # comments may mention anything
    for m in re.finditer(r"sys::(recs_gpu_\w+)\s*\(", code):
        name = m.group(1)
        assert name in funcs, name
        assert _call_arity(code, m.end() - 1) == len(funcs[name][0]), f"{name}: wrong number of arguments"
    for name in set(re.findall(r"sys::(RECS_\w+)", code)):
        assert name in consts, name
    for name in set(re.findall(r"sys::(recs_gpu_\w+)\b(?!\s*\()", code)):
        assert name in structs or name == "recs_gpu_ctx", name
    # struct literals name every field of the C struct, and only those
    for m in re.finditer(r"sys::(recs_gpu_\w+)\s*\{", code):
        end = code.index("}", m.end())
        fields = re.findall(r"(\w+):", code[m.end():end])
        assert fields == [f for f, _ in structs[m.group(1)]], m.group(1)
    # the trait the reference schedules and runs: all six methods of `System`, plus `run`
    for method in ("fn name(&self)", "fn component_accesses(&self)", "fn try_as_sequentially_iterable(&self)",
                   "fn try_as_segment_iterable(&self)", "fn command_buffer(&self)", "fn command_receiver(&self)",
                   "fn run(&self, world: &World) -> SystemResult<()>"):
        assert method in shim, method
    assert "impl System for GpuSystem" in shim and "impl SequentiallyIterable for GpuSystem" in shim
    for opener, closer in ("{}", "()", "[]"):
        assert code.count(opener) == code.count(closer), f"unbalanced {opener}{closer}"
