fn main() {
    // libfheram_cuda.so is built by `python -c "import __graft_entry__ as g; g.build()"`
    let dir = std::env::var("FHERAM_LIB_DIR").unwrap_or_else(|_| "../fhe_ram_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=fheram_cuda");
}
