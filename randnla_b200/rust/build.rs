fn main() {
    // librnla.so is built by `make -C randnla_b200/csrc`; point RNLA_LIB_DIR at the directory holding it
    let dir = std::env::var("RNLA_LIB_DIR").unwrap_or_else(|_| "../".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=rnla");
}
