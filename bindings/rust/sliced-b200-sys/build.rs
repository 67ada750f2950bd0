// Points rustc at libsliced_b200.so.  SLICED_B200_LIB_DIR = directory that holds the library built by
// `python -m sliced_b200.build` (default: the in-tree location relative to this crate).
fn main() {
    let dir = std::env::var("SLICED_B200_LIB_DIR").unwrap_or_else(|_| {
        let here = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{here}/../../../sliced_b200")
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=sliced_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=SLICED_B200_LIB_DIR");
}
