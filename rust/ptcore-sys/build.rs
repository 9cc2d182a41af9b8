// UNVERIFIED (no Rust toolchain in the build image).
// Builds libptcore.so from the CUDA sources of this repository with nvcc for sm_100a and links it.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = root.join("raytracer-rust_b200/csrc");
    let lib = out.join("libptcore.so");
    let status = Command::new(env::var("NVCC").unwrap_or_else(|_| "nvcc".into()))
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--fmad=false"])
        .args(["-Xcompiler", "-fPIC,-ffp-contract=off,-pthread", "-shared", "-o"])
        .arg(&lib)
        .arg(csrc.join("ptcore.cu"))
        .arg(csrc.join("pt_build.cpp"))
        .status()
        .expect("nvcc not found");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=ptcore");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", root.join("include/ptcore.h").display());
}
