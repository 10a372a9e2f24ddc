// HBMPC_B200_LIB_DIR = directory holding libhbmpc_b200.so (built by `python mpc-protocols_b200/build.py`, nvcc, sm_100a).
fn main() {
    let dir = std::env::var("HBMPC_B200_LIB_DIR").expect("set HBMPC_B200_LIB_DIR to the directory of libhbmpc_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=hbmpc_b200");
    println!("cargo:rerun-if-env-changed=HBMPC_B200_LIB_DIR");
}
