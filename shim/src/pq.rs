//! Product quantizer over the B200 engine (feature `b200`).
//!
//! Same public surface as the CPU implementation (reference src/pq.rs:39-209): `ProductQuantizer::new(training_data, m,
//! k, max_iters, distance, seed)`, the getters, and the `Quantizer` impl returning `Vec<f16>`.  What changes underneath:
//! * validation stays here, in the reference's order (pq.rs:91-117, vector.rs:396-410), so callers see the same errors
//!   before the GPU is touched;
//! * the RNG stays here: `choose_multiple` / `choose` depend only on the slice LENGTH, so drawing row indices from
//!   `StdRng::seed_from_u64(seed + i)` consumes exactly the stream `lbg_quantize` consumes (vector.rs:412-413, 450);
//! * `lbg_quantize` for all m subspaces is one `vqb_pq_train` call; `quantize` is a batch of one through `vqb_pq_encode`;
//! * new batch methods (`quantize_batch`, `encode`, `decode`) avoid the per-vector call overhead (ROADMAP.md:30-31).

use std::os::raw::{c_int, c_void};
use std::ptr::null_mut;

use half::f16;
use rand::prelude::*;
use rand::rngs::StdRng;

use crate::core::distance::Distance;
use crate::core::error::{VqError, VqResult};
use crate::core::quantizer::Quantizer;
use crate::core::vqb200_ffi::*;

pub struct ProductQuantizer {
    handle: *mut VqbPq,
    codebooks: Vec<f32>, // [m][k][sub_dim]
    sub_dim: usize,
    m: usize,
    k: usize,
    dim: usize,
    distance: Distance,
}

// The handle is immutable after creation and the engine serialises calls on its context.
unsafe impl Send for ProductQuantizer {}
unsafe impl Sync for ProductQuantizer {}

/// Distance -> VQB_* id (enum order of src/core/distance.rs:8-17).
fn metric_id(d: &Distance) -> c_int {
    match d {
        Distance::SquaredEuclidean => 0,
        Distance::Euclidean => 1,
        Distance::Manhattan => 2,
        Distance::CosineDistance => 3,
    }
}

struct ReseedState<'a> {
    rngs: &'a mut [StdRng],
    ids: &'a [usize],
}

unsafe extern "C" fn reseed_cb(user: *mut c_void, subspace: u32) -> u64 {
    // SAFETY: `user` is the ReseedState passed to vqb_pq_train below and outlives the call.
    let st = unsafe { &mut *(user as *mut ReseedState) };
    *st.ids.choose(&mut st.rngs[subspace as usize]).unwrap() as u64 // vector.rs:450
}

impl ProductQuantizer {
    pub fn new(
        training_data: &[&[f32]],
        m: usize,
        k: usize,
        max_iters: usize,
        distance: Distance,
        seed: u64,
    ) -> VqResult<Self> {
        // ---- pq.rs:91-117, unchanged
        if training_data.is_empty() {
            return Err(VqError::EmptyInput);
        }
        let dim = training_data[0].len();
        for vec in training_data.iter() {
            if vec.len() != dim {
                return Err(VqError::DimensionMismatch { expected: dim, found: vec.len() });
            }
        }
        if dim < m {
            return Err(VqError::InvalidParameter {
                parameter: "m",
                reason: format!("must be at most the data dimension ({})", dim),
            });
        }
        if dim % m != 0 {
            return Err(VqError::InvalidParameter {
                parameter: "m",
                reason: format!("dimension ({}) must be divisible by m", dim),
            });
        }
        // ---- vector.rs:396-410 (lbg_quantize's own checks)
        let n = training_data.len();
        if k == 0 {
            return Err(VqError::InvalidParameter { parameter: "k", reason: "must be greater than 0".to_string() });
        }
        if n < k {
            return Err(VqError::InvalidParameter {
                parameter: "k",
                reason: format!("not enough data points ({}) for {} clusters", n, k),
            });
        }
        let sub_dim = dim / m;

        // one contiguous row-major [n, dim] block instead of n*m small Vecs (pq.rs:122-129)
        let mut flat = Vec::with_capacity(n * dim);
        for v in training_data {
            flat.extend_from_slice(v);
        }

        // the reference's index stream: one StdRng per subspace, seed + i (pq.rs:130)
        let ids: Vec<usize> = (0..n).collect();
        let mut rngs: Vec<StdRng> = (0..m).map(|i| StdRng::seed_from_u64(seed.wrapping_add(i as u64))).collect();
        let mut init = Vec::with_capacity(m * k);
        for r in rngs.iter_mut() {
            init.extend(ids.choose_multiple(r, k).map(|&i| i as u64)); // vector.rs:413
        }
        let mut st = ReseedState { rngs: &mut rngs, ids: &ids };
        let opts = VqbTrainOpts {
            struct_size: std::mem::size_of::<VqbTrainOpts>() as u32,
            update_mode: VQB_UPDATE_ORDERED, // the reference's summation order: codebooks bit-identical with the CPU build
            assign_mode: VQB_ASSIGN_AUTO,
            flags: 0,
            reseed: Some(reseed_cb),
            reseed_user: &mut st as *mut ReseedState as *mut c_void,
            allreduce: None,
            allreduce_user: null_mut(),
            row_offset: 0,
            n_global: 0,
            iter_ms: null_mut(),
        };
        let eng = engine()?;
        let mut codebooks = vec![0f32; m * k * sub_dim];
        // SAFETY: all pointers are valid for the sizes passed; the call blocks until training is complete.
        check(eng.0, unsafe {
            vqb_pq_train(eng.0, flat.as_ptr(), n, dim, m, k, max_iters, init.as_ptr(), &opts, codebooks.as_mut_ptr(), null_mut())
        })?;
        let mut handle: *mut VqbPq = null_mut();
        // SAFETY: codebooks holds m*k*sub_dim floats; handle is a valid out-pointer.
        check(eng.0, unsafe { vqb_pq_create(eng.0, codebooks.as_ptr(), m, k, sub_dim, metric_id(&distance), &mut handle) })?;
        Ok(Self { handle, codebooks, sub_dim, m, k, dim, distance })
    }

    pub fn num_subspaces(&self) -> usize { self.m }
    pub fn sub_dim(&self) -> usize { self.sub_dim }
    pub fn dim(&self) -> usize { self.dim }
    pub fn num_centroids(&self) -> usize { self.k }
    pub fn distance_metric(&self) -> &'static str { self.distance.name() }
    /// Trained codebooks, [m][k][sub_dim] row-major.
    pub fn codebooks(&self) -> &[f32] { &self.codebooks }

    /// n vectors (row-major `x`, n * dim floats) -> the reference's output format for each: dim f16 values.
    pub fn quantize_batch(&self, x: &[f32]) -> VqResult<Vec<f16>> {
        if x.len() % self.dim != 0 {
            return Err(VqError::DimensionMismatch { expected: self.dim, found: x.len() % self.dim });
        }
        let n = x.len() / self.dim;
        let mut out = vec![f16::ZERO; x.len()];
        let eng = engine()?;
        // SAFETY: x holds n*dim floats, out n*dim u16-sized values.
        check(eng.0, unsafe {
            vqb_pq_encode(self.handle, x.as_ptr(), n, VQB_ASSIGN_AUTO, null_mut(), 1, out.as_mut_ptr() as *mut u16)
        })?;
        Ok(out)
    }

    /// n vectors -> n * m compact codes (u8 when k <= 256, which `new` does not require: wider codes via `encode_u32`).
    pub fn encode(&self, x: &[f32]) -> VqResult<Vec<u8>> {
        if self.k > 256 {
            return Err(VqError::InvalidParameter { parameter: "k", reason: "codes do not fit u8; use encode_u32".to_string() });
        }
        if x.len() % self.dim != 0 {
            return Err(VqError::DimensionMismatch { expected: self.dim, found: x.len() % self.dim });
        }
        let n = x.len() / self.dim;
        let mut codes = vec![0u8; n * self.m];
        let eng = engine()?;
        // SAFETY: codes holds n*m bytes.
        check(eng.0, unsafe {
            vqb_pq_encode(self.handle, x.as_ptr(), n, VQB_ASSIGN_AUTO, codes.as_mut_ptr() as *mut c_void, 1, null_mut())
        })?;
        Ok(codes)
    }

    pub fn encode_u32(&self, x: &[f32]) -> VqResult<Vec<u32>> {
        if x.len() % self.dim != 0 {
            return Err(VqError::DimensionMismatch { expected: self.dim, found: x.len() % self.dim });
        }
        let n = x.len() / self.dim;
        let mut codes = vec![0u32; n * self.m];
        let eng = engine()?;
        // SAFETY: codes holds n*m u32 values.
        check(eng.0, unsafe {
            vqb_pq_encode(self.handle, x.as_ptr(), n, VQB_ASSIGN_AUTO, codes.as_mut_ptr() as *mut c_void, 4, null_mut())
        })?;
        Ok(codes)
    }

    /// codes (n * m, u8) -> n * dim reconstructed f32 values (through f16, like quantize + dequantize).
    pub fn decode(&self, codes: &[u8]) -> VqResult<Vec<f32>> {
        if codes.len() % self.m != 0 {
            return Err(VqError::DimensionMismatch { expected: self.m, found: codes.len() % self.m });
        }
        let n = codes.len() / self.m;
        let mut out = vec![0f32; n * self.dim];
        let eng = engine()?;
        // SAFETY: out holds n*dim floats.
        check(eng.0, unsafe { vqb_pq_decode(self.handle, codes.as_ptr() as *const c_void, 1, n, out.as_mut_ptr()) })?;
        Ok(out)
    }
}

impl Drop for ProductQuantizer {
    fn drop(&mut self) {
        if !self.handle.is_null() {
            // SAFETY: the handle came from vqb_pq_create and is destroyed once.
            unsafe { vqb_pq_destroy(self.handle) };
        }
    }
}

impl Quantizer for ProductQuantizer {
    type QuantizedOutput = Vec<f16>;

    /// pq.rs:167-199 as a batch of one: per subspace the nearest centroid (strict '<', lowest index wins, the metric's
    /// hsdlib arithmetic), output = that centroid's values as f16.
    fn quantize(&self, vector: &[f32]) -> VqResult<Self::QuantizedOutput> {
        if vector.len() != self.dim {
            return Err(VqError::DimensionMismatch { expected: self.dim, found: vector.len() });
        }
        self.quantize_batch(vector)
    }

    /// pq.rs:201-209, unchanged: f16 -> f32 element-wise.
    fn dequantize(&self, quantized: &Self::QuantizedOutput) -> VqResult<Vec<f32>> {
        if quantized.len() != self.dim {
            return Err(VqError::DimensionMismatch { expected: self.dim, found: quantized.len() });
        }
        Ok(quantized.iter().map(|&x| f16::to_f32(x)).collect())
    }
}

