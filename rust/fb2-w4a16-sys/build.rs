// Links libfb2_w4a16.so.  FB2_W4A16_DIR names the directory that holds it (the output of
// `make -C ferrum-infer-rs_b200/csrc`); the library itself is built by nvcc for sm_100a, not by cargo.
fn main() {
    println!("cargo:rerun-if-env-changed=FB2_W4A16_DIR");
    if let Ok(dir) = std::env::var("FB2_W4A16_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=fb2_w4a16");
}
