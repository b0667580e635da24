// Link against libspenso_b200.so.  SPENSO_B200_LIB_DIR points at the directory holding it
// (e.g. <repo>/spenso_b200 after `python -c "import __graft_entry__ as g; g.build()"`).
fn main() {
    if let Ok(dir) = std::env::var("SPENSO_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=spenso_b200");
    println!("cargo:rerun-if-env-changed=SPENSO_B200_LIB_DIR");
}
