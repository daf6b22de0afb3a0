// build.rs — link flags for liblbft_b200.so (INTEGRATION.md "Build / link").
//
// LBFT_B200_LIB_DIR names the directory holding liblbft_b200.so (in this repository:
// librabft_simulator_b200/csrc, produced by `python -c "import __graft_entry__ as g; g.build()"`).  Default:
// ../librabft_simulator_b200/csrc relative to this crate, i.e. the layout where the crate lives in the repository root.
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("LBFT_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../librabft_simulator_b200/csrc")
    });
    println!("cargo:rerun-if-env-changed=LBFT_B200_LIB_DIR");
    println!("cargo:rerun-if-changed=../include/lbft.h");
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=lbft_b200");
    // so that `cargo test` / `cargo run` find the library without LD_LIBRARY_PATH
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
}
