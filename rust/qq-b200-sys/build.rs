// NOT BUILT IN THIS ENVIRONMENT. Links the C-ABI library produced by __graft_entry__.build().
fn main() {
    let dir = std::env::var("QQ_B200_LIB_DIR").unwrap_or_else(|_| "../../quisquis-rust_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=qq_b200");
}
