// build.rs -- feature "b200": link the B200 engine the way feature "simd" links hsdlib (reference build.rs:6-40).
//
// The engine is CUDA, so instead of the `cc` crate the build runs the repository's own nvcc recipe
// (`python -m vq_b200.build` == nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false
//  -Xcompiler -fPIC vq_b200/csrc/*.cu, then nvcc -shared -cudart static -o vq_b200/libvqb200.so) and links the result.
fn main() {
    #[cfg(feature = "simd")]
    {
        // unchanged: the reference's cc::Build of external/hsdlib (build.rs:6-40)
    }

    #[cfg(feature = "b200")]
    {
        let vqb = std::env::var("VQB200_DIR").expect("VQB200_DIR must point at a checkout of the vq_b200 repository");
        let status = std::process::Command::new(std::env::var("PYTHON").unwrap_or_else(|_| "python".into()))
            .args(["-m", "vq_b200.build"])
            .current_dir(&vqb)
            .status()
            .expect("failed to run the nvcc build of libvqb200.so");
        assert!(status.success(), "nvcc build of libvqb200.so failed");
        println!("cargo:rustc-link-search=native={vqb}/vq_b200");
        println!("cargo:rustc-link-lib=dylib=vqb200");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{vqb}/vq_b200");
        println!("cargo:rerun-if-changed={vqb}/vq_b200/csrc");
        println!("cargo:rerun-if-changed={vqb}/include/vqb200.h");
        println!("cargo:rerun-if-env-changed=VQB200_DIR");
    }
}
