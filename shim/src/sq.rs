//! Batch methods for ScalarQuantizer over the B200 engine (feature `b200`).
//!
//! The struct, `new` (which computes `step = (max - min) / (levels - 1) as f32`, reference src/sq.rs:94), the getters and
//! the per-vector `Quantizer` impl stay as they are.  The bulk path is bit-identical with src/sq.rs:123-151: clamp, true
//! IEEE division by `step`, `round()` half away from zero, saturating cast (NaN -> 0), `min(levels - 1)`; dequantize is
//! `min + idx as f32 * step` with two roundings (no FMA).

use crate::core::error::VqResult;
use crate::core::vqb200_ffi::*;

impl crate::sq::ScalarQuantizer {
    pub fn quantize_bulk(&self, values: &[f32]) -> VqResult<Vec<u8>> {
        let mut out = vec![0u8; values.len()];
        let eng = engine()?;
        // SAFETY: in / out buffers hold values.len() elements.
        check(eng.0, unsafe {
            vqb_sq_quantize(eng.0, values.as_ptr(), values.len(), self.min(), self.max(), self.step(), self.levels() as u32,
                            out.as_mut_ptr())
        })?;
        Ok(out)
    }

    pub fn dequantize_bulk(&self, codes: &[u8]) -> VqResult<Vec<f32>> {
        let mut out = vec![0f32; codes.len()];
        let eng = engine()?;
        // SAFETY: in / out buffers hold codes.len() elements.
        check(eng.0, unsafe { vqb_sq_dequantize(eng.0, codes.as_ptr(), codes.len(), self.min(), self.step(), out.as_mut_ptr()) })?;
        Ok(out)
    }
}
