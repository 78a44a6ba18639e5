// Links libvarpro_b200.so. VARPRO_B200_LIB_DIR = the directory holding the shared library
// (in this repository: <repo>/varpro_b200 after `python -m varpro_b200.build`).
fn main() {
    println!("cargo:rerun-if-env-changed=VARPRO_B200_LIB_DIR");
    if let Ok(dir) = std::env::var("VARPRO_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=varpro_b200");
}
