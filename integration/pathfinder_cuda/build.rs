// Builds libpf_cuda.so from the CUDA sources of this repository for sm_100a and links it.
// PF_CUDA_SRC points at pathfinder_b200/csrc (default: a checkout next to the workspace). The source list and
// the flags are those of pathfinder_b200/csrc/Makefile (tests/test_capi_symbols.py keeps the two in step):
// every file with -fmad=false (dice / bin reproduce the CPU tiler's roundings), except composite.cu.
use std::{env, path::PathBuf, process::Command};

const EXACT_SOURCES: &[&str] = &["kernels.cu", "renderer.cu", "scene.cpp", "stroke.cpp", "svg.cpp", "dilate.cpp", "font.cpp"];
const CONTRACTED_SOURCES: &[&str] = &["composite.cu"];

fn nvcc(src: &PathBuf, out: &PathBuf, file: &str, exact: bool) -> PathBuf {
    let object = out.join(format!("{}.o", file));
    let mut command = Command::new("nvcc");
    command.args(&["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
                   "-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "-x", "cu", "-c", "-o"]);
    command.arg(&object).arg(src.join(file));
    if exact {
        command.arg("-fmad=false");
    }
    assert!(command.status().expect("nvcc not found").success(), "nvcc failed on {}", file);
    object
}

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let src = PathBuf::from(env::var("PF_CUDA_SRC").unwrap_or_else(|_| "../../pathfinder_b200/csrc".into()));
    let mut objects = vec![];
    for file in EXACT_SOURCES {
        objects.push(nvcc(&src, &out, file, true));
    }
    for file in CONTRACTED_SOURCES {
        objects.push(nvcc(&src, &out, file, false));
    }
    let status = Command::new("nvcc")
        .args(&["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o"])
        .arg(out.join("libpf_cuda.so"))
        .args(&objects)
        .status()
        .expect("nvcc not found");
    assert!(status.success(), "nvcc link failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=pf_cuda");
    println!("cargo:rerun-if-changed={}", src.display());
}
