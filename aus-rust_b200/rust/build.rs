// build.rs -- compiles the CUDA layer for sm_100a with nvcc and links it.  No build-dependencies (no `cc` crate):
// plain std::process::Command, because crates cannot be fetched in the target environment.
// The library is one translation unit per fft_size (aus_kernels.cu -DAUS_LG=6..14) plus the host side (aus_b200.cu);
// the units are compiled concurrently and linked into one shared object, like csrc/Makefile does.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = root.join("..").join("csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libaus_b200.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let flags = ["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"];
    let mut jobs = Vec::new();
    let mut objs = Vec::new();
    let host = out.join("aus_b200.o");
    jobs.push(Command::new(&nvcc).args(flags).arg("-c").arg("-o").arg(&host).arg(csrc.join("aus_b200.cu"))
        .spawn().expect("nvcc not found: this crate has no CPU fallback"));
    objs.push(host);
    for lg in 6..=14 {
        let o = out.join(format!("aus_kernels_{lg}.o"));
        jobs.push(Command::new(&nvcc).args(flags).arg(format!("-DAUS_LG={lg}")).arg("-c").arg("-o").arg(&o)
            .arg(csrc.join("aus_kernels.cu")).spawn().expect("nvcc not found"));
        objs.push(o);
    }
    for mut j in jobs { assert!(j.wait().expect("nvcc").success(), "nvcc failed"); }
    let status = Command::new(&nvcc).args(["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o"]).arg(&lib).args(&objs)
        .status().expect("nvcc");
    assert!(status.success(), "link failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=aus_b200");
    for f in ["aus_b200.cu", "aus_kernels.cu", "aus_host.cuh", "aus_device.cuh", "aus_stft_kernel.cuh", "aus_stft_r.cuh",
              "aus_istft_kernel.cuh"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", root.join("..").join("..").join("include").join("aus_b200.h").display());
}
