// build.rs — compiles the CUDA engine for sm_100a and links it (reference side; not built in this image).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from("../kzg_poly_commit_exploration_b200/csrc");
    let mut objs = vec![];
    // the same list as kzg_poly_commit_exploration_b200/build.py SOURCES
    for f in ["abi.cu", "engine_one.cu", "host_helpers.cu", "poly.cu", "setup.cu", "setup_g2.cu", "g2_import.cu", "pairing.cu", "msm.cu", "msm_tail.cu",
              "msm_tree.cu", "msm_tree_r1.cu", "batch.cu", "batch_tree.cu", "probe.cu"] {
        let obj = out.join(f.replace(".cu", ".o"));
        let st = Command::new("nvcc")
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-c"])
            .arg(csrc.join(f)).arg("-o").arg(&obj)
            .status().expect("nvcc not found");
        assert!(st.success(), "nvcc failed on {f}");
        objs.push(obj);
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    let lib = out.join("libkzg_b200.a");
    assert!(Command::new("ar").arg("crs").arg(&lib).args(&objs).status().unwrap().success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=kzg_b200");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
}
