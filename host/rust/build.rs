// build.rs for the reference crate when it links the B200 leaf layer.
// LEAFLINE_B200_LIB_DIR = the directory holding libleafline_b200.so.
fn main() {
    let dir = std::env::var("LEAFLINE_B200_LIB_DIR").expect("set LEAFLINE_B200_LIB_DIR to the directory of libleafline_b200.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=leafline_b200");
    println!("cargo:rerun-if-env-changed=LEAFLINE_B200_LIB_DIR");
}
