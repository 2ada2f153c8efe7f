// Points the linker at libws_b200.so.  WS_B200_LIB_DIR overrides the in-tree location
// (../rustronomy-watershed_b200, where `python rustronomy-watershed_b200/build.py` leaves the library).
use std::path::PathBuf;

fn main() {
  let dir = std::env::var("WS_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
    PathBuf::from(std::env::var("CARGO_MANIFEST_DIR").unwrap()).join("..").join("rustronomy-watershed_b200")
  });
  println!("cargo:rustc-link-search=native={}", dir.display());
  println!("cargo:rustc-link-lib=dylib=ws_b200");
  println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
  println!("cargo:rerun-if-env-changed=WS_B200_LIB_DIR");
}
