// Links the prebuilt librnmf_b200.so (make -C rnmf_b200/csrc; nvcc, sm_100a).
fn main() {
    let dir = std::env::var("RNMF_B200_LIB_DIR").expect("set RNMF_B200_LIB_DIR to the directory that holds librnmf_b200.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=rnmf_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=RNMF_B200_LIB_DIR");
}
