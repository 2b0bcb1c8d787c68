//! Batch methods for BinaryQuantizer over the B200 engine (feature `b200`).
//!
//! The struct, `new`, the getters and the per-vector `Quantizer` impl stay as they are (reference src/bq.rs:55-118): a
//! single short vector is faster on the CPU than a kernel launch.  These `impl` blocks add the bulk path -- the case the
//! engine is for (BASELINE config 2: 100M x 1536 values) -- with bit-identical results: `x >= threshold ? high : low`
//! (NaN -> low, -0.0 >= 0.0 is true) and `c >= high ? high : low` as f32.

use crate::core::error::VqResult;
use crate::core::vqb200_ffi::*;

impl crate::bq::BinaryQuantizer {
    /// bq.rs:94-105 over any number of values.
    pub fn quantize_bulk(&self, values: &[f32]) -> VqResult<Vec<u8>> {
        let mut out = vec![0u8; values.len()];
        let eng = engine()?;
        // SAFETY: in / out buffers hold values.len() elements.
        check(eng.0, unsafe {
            vqb_bq_quantize(eng.0, values.as_ptr(), values.len(), self.threshold(), self.low(), self.high(), out.as_mut_ptr())
        })?;
        Ok(out)
    }

    /// bq.rs:107-118 over any number of codes.
    pub fn dequantize_bulk(&self, codes: &[u8]) -> VqResult<Vec<f32>> {
        let mut out = vec![0f32; codes.len()];
        let eng = engine()?;
        // SAFETY: in / out buffers hold codes.len() elements.
        check(eng.0, unsafe { vqb_bq_dequantize(eng.0, codes.as_ptr(), codes.len(), self.low(), self.high(), out.as_mut_ptr()) })?;
        Ok(out)
    }
}
