// Links the prebuilt libhnsw_b200.so; set HNSW_B200_LIB_DIR to the directory that holds it
// (hnsw_rs_b200/ in the engine's repository after `make -C hnsw_rs_b200/csrc`).
fn main() {
    let dir = std::env::var("HNSW_B200_LIB_DIR").expect("set HNSW_B200_LIB_DIR to the directory of libhnsw_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=hnsw_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=HNSW_B200_LIB_DIR");
}
