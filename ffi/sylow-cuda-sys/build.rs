// Tells cargo where libsylow_b200.so lives.  The library itself is built by `python -m sylow_b200.build`
// (one nvcc invocation, sm_100a); this crate never compiles CUDA.
fn main() {
    let dir = std::env::var("SYLOW_B200_LIB_DIR").unwrap_or_else(|_| "../../sylow_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=sylow_b200");
    println!("cargo:rerun-if-env-changed=SYLOW_B200_LIB_DIR");
}
