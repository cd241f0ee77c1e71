fn main() {
    // libfloria_b200.so is built by `python -m floria_b200.build` (nvcc, sm_100a)
    let dir = std::env::var("FLORIA_B200_LIB_DIR").unwrap_or_else(|_| "../floria_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=floria_b200");
}
