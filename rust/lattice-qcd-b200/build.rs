// AUTHORED, NOT COMPILED here (no rustc/cargo in the image).
//
// Builds lattice_qcd_rs_b200/csrc/lq_capi.cu for sm_100a with nvcc (through the `cc` crate's CUDA mode) and links
// it statically, so that `cargo build` of a user crate needs nothing but the CUDA toolkit.
// Set LQCD_B200_LIB_DIR to link a prebuilt liblqcd_b200.so instead (python -m lattice_qcd_rs_b200.build).
use std::{env, path::PathBuf};

fn main() {
    println!("cargo:rerun-if-env-changed=LQCD_B200_LIB_DIR");
    if let Ok(dir) = env::var("LQCD_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=lqcd_b200");
        return;
    }
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("lattice_qcd_rs_b200/csrc");
    for f in ["lq_capi.cu", "lq_kernels.cuh", "lq_common.cuh", "lq_local.cuh", "lq_tuned.cuh", "lq_geom_host.h"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", root.join("include/lqcd_b200.h").display());
    cc::Build::new()
        .cuda(true)
        .cudart("shared")
        .flag("-gencode")
        .flag("arch=compute_100a,code=sm_100a")
        .flag("-lineinfo")
        .flag("-O3")
        .flag("-std=c++17")
        .define("LQ_BUILD_CUDA", "1")
        .define("LQ_HAVE_TUNED", "1")
        .include(root.join("include"))
        .file(csrc.join("lq_capi.cu"))
        .compile("lqcd_b200");
}
