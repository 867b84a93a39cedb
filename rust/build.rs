// build.rs -- compiles the CUDA sources for sm_100a with nvcc and links the resulting shared library.
// (Equivalent to mercer_research_b200/csrc/Makefile; no cudarc / cc crate needed.)
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..");
    let csrc = root.join("mercer_research_b200").join("csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let sources = ["model.cu", "features.cu", "dense.cu", "smallnet.cu", "conv.cu", "extops.cu", "dp.cu", "ozaki.cu"];
    let mut objs = Vec::new();
    for s in sources {
        let obj = out.join(s.replace(".cu", ".o"));
        let ok = Command::new(&nvcc)
            .args(["-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr"])
            .args(["-gencode", "arch=compute_100a,code=sm_100a"])
            .args(["-Xcompiler", "-fPIC", "-c"])
            .arg(csrc.join(s))
            .arg("-o")
            .arg(&obj)
            .status()
            .expect("nvcc not found")
            .success();
        assert!(ok, "nvcc failed on {s}");
        objs.push(obj);
        println!("cargo:rerun-if-changed={}", csrc.join(s).display());
    }
    let lib = out.join("librcn_cuda.so");
    let ok = Command::new(&nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o"])
        .arg(&lib)
        .args(&objs)
        .arg("-lcudart")
        .status()
        .unwrap()
        .success();
    assert!(ok, "link failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=rcn_cuda");
    // src/cuda.rs binds a handful of libcudart entry points directly (device / pinned buffers)
    let cuda_lib = env::var("CUDA_LIB_DIR").unwrap_or_else(|_| "/usr/local/cuda/lib64".into());
    println!("cargo:rustc-link-search=native={cuda_lib}");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rerun-if-changed={}", root.join("include").join("rcn_cuda.h").display());
}
