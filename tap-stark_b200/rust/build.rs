// build.rs -- compiles the CUDA sources for sm_100a with nvcc and links the resulting shared library.
// No Triton, no multi-backend dispatch, no CPU fallback: if nvcc is missing the build fails.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libtapstark_b200.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let status = Command::new(nvcc)
        .args(["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-shared",
               "-Xcompiler", "-fPIC", "-o"])
        .arg(&lib)
        .arg(root.join("tap-stark_b200/csrc/tapstark.cu"))
        .status()
        .expect("nvcc not found: tapstark-gpu has no CPU fallback");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=tapstark_b200");
    println!("cargo:rerun-if-changed={}", root.join("tap-stark_b200/csrc").display());
    println!("cargo:rerun-if-changed={}", root.join("include/tapstark.h").display());
}
