// Builds libpf_cuda.so from the CUDA sources of this repository for sm_100a and links it.
// PF_CUDA_SRC points at pathfinder_b200/csrc (default: a checkout next to the workspace).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let src = PathBuf::from(env::var("PF_CUDA_SRC").unwrap_or_else(|_| "../../pathfinder_b200/csrc".into()));
    let status = Command::new("nvcc")
        .args(&["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
                "-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "-shared", "-o"])
        .arg(out.join("libpf_cuda.so"))
        .arg(src.join("kernels.cu"))
        .arg(src.join("renderer.cu"))
        .args(&["-x", "cu"])
        .arg(src.join("scene.cpp"))
        .status()
        .expect("nvcc not found");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=pf_cuda");
    println!("cargo:rerun-if-changed={}", src.display());
}
