// Links the reference crate against the sm_100a backend when the `gpu` feature is on.
fn main() {
    if std::env::var_os("CARGO_FEATURE_GPU").is_some() {
        let dir = std::env::var("GKR_MSM_B200_LIB_DIR").expect("set GKR_MSM_B200_LIB_DIR to <gkr-msm-b200>/gkr-msm_b200/lib");
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=gkr_msm_b200");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
        println!("cargo:rerun-if-env-changed=GKR_MSM_B200_LIB_DIR");
    }
}
