// build.rs -- compiles the CUDA layer for sm_100a with nvcc and links it.  No build-dependencies (no `cc` crate):
// plain std::process::Command, because crates cannot be fetched in the target environment.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = root.join("..").join("csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libaus_b200.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let status = Command::new(nvcc)
        .args(["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC", "-shared", "-o"])
        .arg(&lib)
        .arg(csrc.join("aus_b200.cu"))
        .status()
        .expect("nvcc not found: this crate has no CPU fallback");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=aus_b200");
    for f in ["aus_b200.cu", "aus_device.cuh", "aus_stft_kernel.cuh", "aus_istft_kernel.cuh"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", root.join("..").join("..").join("include").join("aus_b200.h").display());
}
