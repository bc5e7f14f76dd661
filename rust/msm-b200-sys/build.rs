//! Replaces ag-cuda-ec/build.rs + ag_build::generate (ag-build/src/compile.rs:44-130): no source
//! generation, no fatbin; the hand-written sm_100a kernels are compiled by the repository's Makefile
//! (nvcc -gencode arch=compute_100a,code=sm_100a) into one shared library that is linked here.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("0g-ec-gpu_b200/csrc");
    let status = Command::new("make").arg("-C").arg(&csrc).arg("-j4").status().expect("make not found");
    assert!(status.success(), "building libmsm_b200.so failed (nvcc with sm_100a support required)");
    println!("cargo:rustc-link-search=native={}", root.join("0g-ec-gpu_b200").display());
    println!("cargo:rustc-link-lib=dylib=msm_b200");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", root.join("include/msm_b200.h").display());
}
