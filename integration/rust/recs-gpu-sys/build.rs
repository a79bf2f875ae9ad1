//! Tells cargo where librecs_gpu.so lives.  The library is built from this repository with `make` (nvcc, sm_100a) and
//! is NOT built by cargo: set RECS_GPU_LIB_DIR to the directory holding it (default: ../../../recs_b200 relative to
//! this crate, i.e. the in-tree build).
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("RECS_GPU_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").expect("cargo sets CARGO_MANIFEST_DIR")).join("../../../recs_b200")
    });
    println!("cargo:rerun-if-env-changed=RECS_GPU_LIB_DIR");
    println!("cargo:rerun-if-changed=../../../include/recs_gpu.h");
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=recs_gpu");
    // librecs_gpu.so carries its own CUDA runtime (static cudart); NCCL is dlopen'ed at run time when sharding is used.
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
}
