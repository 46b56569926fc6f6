// Links the prebuilt CUDA library.  OHWHISPER_LIB_DIR = directory holding libohwhisper.so
// (openhush_b200/csrc after `python -c "import __graft_entry__ as g; g.build()"`).
fn main() {
    let dir = std::env::var("OHWHISPER_LIB_DIR").expect("set OHWHISPER_LIB_DIR to the directory that holds libohwhisper.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=ohwhisper");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=OHWHISPER_LIB_DIR");
}
