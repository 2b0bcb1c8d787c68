//! Tree-structured vector quantizer over the B200 engine (feature `b200`); public surface of reference src/tsvq.rs:159-265.
//!
//! `TSVQNode::build` (tsvq.rs:31-115: mean, max-variance dimension, median split, recursion) runs level-synchronously on
//! the GPU inside `vqb_tsvq_train`; `find_leaf` + the f16 output (tsvq.rs:117-132, 239-255) is `vqb_tsvq_encode`.

use std::os::raw::c_int;
use std::ptr::null_mut;

use half::f16;

use crate::core::distance::Distance;
use crate::core::error::{VqError, VqResult};
use crate::core::quantizer::Quantizer;
use crate::core::vqb200_ffi::*;

pub struct TSVQ {
    handle: *mut VqbTsvq,
    dim: usize,
    distance: Distance,
}

unsafe impl Send for TSVQ {}
unsafe impl Sync for TSVQ {}

fn metric_id(d: &Distance) -> c_int {
    match d {
        Distance::SquaredEuclidean => 0,
        Distance::Euclidean => 1,
        Distance::Manhattan => 2,
        Distance::CosineDistance => 3,
    }
}

impl TSVQ {
    pub fn new(training_data: &[&[f32]], max_depth: usize, distance: Distance) -> VqResult<Self> {
        // tsvq.rs:196-210, unchanged
        if training_data.is_empty() {
            return Err(VqError::EmptyInput);
        }
        let dim = training_data[0].len();
        for v in training_data.iter() {
            if v.len() != dim {
                return Err(VqError::DimensionMismatch { expected: dim, found: v.len() });
            }
        }
        let n = training_data.len();
        let mut flat = Vec::with_capacity(n * dim);
        for v in training_data {
            flat.extend_from_slice(v);
        }
        let eng = engine()?;
        let mut handle: *mut VqbTsvq = null_mut();
        // SAFETY: flat holds n*dim floats; handle is a valid out-pointer.
        check(eng.0, unsafe { vqb_tsvq_train(eng.0, flat.as_ptr(), n, dim, max_depth, metric_id(&distance), &mut handle) })?;
        Ok(TSVQ { handle, dim, distance })
    }

    pub fn dim(&self) -> usize { self.dim }
    pub fn distance_metric(&self) -> &'static str { self.distance.name() }

    /// n vectors (row-major) -> n * dim f16 leaf-centroid values (the reference's output format per vector).
    pub fn quantize_batch(&self, x: &[f32]) -> VqResult<Vec<f16>> {
        if x.len() % self.dim != 0 {
            return Err(VqError::DimensionMismatch { expected: self.dim, found: x.len() % self.dim });
        }
        let n = x.len() / self.dim;
        let mut out = vec![f16::ZERO; x.len()];
        // SAFETY: out holds n*dim u16-sized values.
        check(engine()?.0, unsafe { vqb_tsvq_encode(self.handle, x.as_ptr(), n, null_mut(), out.as_mut_ptr() as *mut u16) })?;
        Ok(out)
    }

    /// n vectors -> the breadth-first node id of the leaf each one reaches.
    pub fn leaf_ids(&self, x: &[f32]) -> VqResult<Vec<u32>> {
        if x.len() % self.dim != 0 {
            return Err(VqError::DimensionMismatch { expected: self.dim, found: x.len() % self.dim });
        }
        let n = x.len() / self.dim;
        let mut out = vec![0u32; n];
        // SAFETY: out holds n u32 values.
        check(engine()?.0, unsafe { vqb_tsvq_encode(self.handle, x.as_ptr(), n, out.as_mut_ptr(), null_mut()) })?;
        Ok(out)
    }
}

impl Drop for TSVQ {
    fn drop(&mut self) {
        if !self.handle.is_null() {
            // SAFETY: the handle came from vqb_tsvq_train and is destroyed once.
            unsafe { vqb_tsvq_destroy(self.handle) };
        }
    }
}

impl Quantizer for TSVQ {
    type QuantizedOutput = Vec<f16>;

    fn quantize(&self, vector: &[f32]) -> VqResult<Self::QuantizedOutput> {
        if vector.len() != self.dim {
            return Err(VqError::DimensionMismatch { expected: self.dim, found: vector.len() });
        }
        self.quantize_batch(vector)
    }

    fn dequantize(&self, quantized: &Self::QuantizedOutput) -> VqResult<Vec<f32>> {
        if quantized.len() != self.dim {
            return Err(VqError::DimensionMismatch { expected: self.dim, found: quantized.len() });
        }
        Ok(quantized.iter().map(|&x| f16::to_f32(x)).collect())
    }
}
