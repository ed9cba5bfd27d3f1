// Builds liblbm_b200.so with nvcc for sm_100a and links it.
// Mirrors simuverse_b200/csrc/build.sh; NVCC / LBM_CSRC may be overridden from the environment.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = env::var("LBM_CSRC")
        .map(PathBuf::from)
        .unwrap_or_else(|_| manifest.join("../../simuverse_b200/csrc"));
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let lib = out.join("liblbm_b200.so");
    let status = Command::new(&nvcc)
        .args([
            "-std=c++17", "-O3", "-lineinfo",
            "-gencode", "arch=compute_100a,code=sm_100a",
            "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
            "-Xcompiler", "-fPIC,-O2,-ffp-contract=off",
            "-shared", "-cudart", "static", "-o",
        ])
        .arg(&lib)
        .arg(csrc.join("lbm_b200.cu"))
        .arg(csrc.join("host_logic.cpp"))
        .status()
        .expect("failed to run nvcc");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=lbm_b200");
    println!("cargo:rerun-if-changed={}", csrc.display());
}
