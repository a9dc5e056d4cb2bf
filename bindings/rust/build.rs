// Link against the in-tree shared library (s-rack_b200/libsrack_b200.so, built by
// `make -C s-rack_b200/csrc`).  SRACK_B200_LIB_DIR overrides the search path.
fn main() {
    let dir = std::env::var("SRACK_B200_LIB_DIR").unwrap_or_else(|_| {
        format!("{}/../../s-rack_b200", std::env::var("CARGO_MANIFEST_DIR").unwrap())
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=srack_b200");
    println!("cargo:rerun-if-env-changed=SRACK_B200_LIB_DIR");
}
