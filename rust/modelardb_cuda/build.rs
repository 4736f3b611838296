// Links libmodelardb_cuda.so.  MODELARDB_CUDA_LIB_DIR names the directory that holds it (the repository's
// modelardb_rs_b200/ directory after `python -m modelardb_rs_b200.build`); the library itself has no link-time
// dependency besides the C++ runtime (the CUDA runtime is linked statically, NCCL is loaded at run time).
fn main() {
    println!("cargo:rerun-if-env-changed=MODELARDB_CUDA_LIB_DIR");
    if let Ok(dir) = std::env::var("MODELARDB_CUDA_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=modelardb_cuda");
}
