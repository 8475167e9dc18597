// Links the C ABI library.  FLOCKGPU_LIB_DIR points at the directory holding libflockgpu.so
// (flock_b200/ in this repository).
fn main() {
    let dir = std::env::var("FLOCKGPU_LIB_DIR").unwrap_or_else(|_| "../../flock_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=flockgpu");
    println!("cargo:rerun-if-env-changed=FLOCKGPU_LIB_DIR");
}
