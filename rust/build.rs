// build.rs — builds the CUDA library with nvcc and links it (source only: no cargo/rustc in the build image).
// The reference's own build.rs (build.rs:1-12) only embeds the git hash; this one would replace it when the
// `cuda` feature is on.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = root.join("dbgphmm_b200/csrc");
    // the same list and flags as dbgphmm_b200/build.py (tests/test_abi.py keeps the two in step)
    let srcs = ["model.cu", "dense.cu", "sparse.cu", "mapx.cu", "engine.cu", "products.cu", "api.cu", "formats.cu", "score.cu"];
    let mut objs = Vec::new();
    for s in srcs {
        let o = out.join(s.replace(".cu", ".o"));
        let st = Command::new("nvcc")
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-c"])
            .arg(csrc.join(s)).arg("-o").arg(&o).status().expect("nvcc not found");
        assert!(st.success(), "nvcc failed on {s}");
        objs.push(o);
        println!("cargo:rerun-if-changed={}", csrc.join(s).display());
    }
    let lib = out.join("libdbgphmm_b200.so");
    let st = Command::new("nvcc").arg("-shared").arg("-o").arg(&lib).args(&objs).args(["-lcudart", "-lz"]).status().unwrap();
    assert!(st.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=dbgphmm_b200");
}
