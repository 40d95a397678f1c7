// build.rs of crate `ptb200-sys` (un-compiled here: no Rust toolchain in this image).
// Compiles the CUDA sources of raytracing-rust_b200/csrc for sm_100a with nvcc and links the result.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("PTB200_ROOT").unwrap_or_else(|_| "../..".into()));
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let pkg = root.join("raytracing-rust_b200");
    let mut objs = Vec::new();
    for src in ["csrc/context.cu", "csrc/lbvh_build.cu", "csrc/wavefront.cu"] {
        let obj = out.join(src.replace('/', "_")).with_extension("o");
        // -fmad=false: rustc never contracts a*b+c; the device must round like the CPU reference does
        let ok = Command::new("nvcc")
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
                   "-Xcompiler", "-fPIC", "-c"])
            .arg(pkg.join(src)).arg("-o").arg(&obj)
            .status().expect("nvcc not found").success();
        assert!(ok, "nvcc failed on {src}");
        objs.push(obj);
    }
    for src in ["host/ssml_loader.cpp", "host/image_out.cpp"] {
        let obj = out.join(src.replace('/', "_")).with_extension("o");
        let ok = Command::new("g++")
            .args(["-O2", "-std=c++17", "-fPIC", "-fno-fast-math", "-ffp-contract=off", "-c"])
            .arg(pkg.join(src)).arg("-o").arg(&obj)
            .status().expect("g++ not found").success();
        assert!(ok, "g++ failed on {src}");
        objs.push(obj);
    }
    let lib = out.join("libptb200.so");
    let ok = Command::new("nvcc").args(["-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-o"])
        .arg(&lib).args(&objs).status().unwrap().success();
    assert!(ok, "link failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=ptb200");
    println!("cargo:rerun-if-changed={}", pkg.join("csrc").display());
    println!("cargo:rerun-if-changed={}", root.join("include/ptb200.h").display());
}
