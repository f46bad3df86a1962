//! Drop-in for the `vectors` crate of Gumo-A/hnsw_rs (vectors/src/lib.rs:10-37, quant.rs, full.rs, serializer.rs).
//! The quantiser (`QuantVec::new`, quant.rs:41-66) and every distance run on the B200 engine with the reference's exact
//! arithmetic (bit-identical results, see the engine's parity tests).  A call on two stand-alone vectors uploads them;
//! bulk work belongs to `points::SimplePoints`, whose vectors stay on the device.
//! Not compiled in the engine's CI (no Rust toolchain there); see INTEGRATION.md.
pub mod serializer;

use hnsw_b200_sys as sys;
use sys::engine::{check, with_ctx};

pub use serializer::Serializer;

/// vectors/src/lib.rs:10-27
pub trait VecBase {
    fn new(vector: &Vec<f32>) -> Self;
    fn dim(&self) -> usize;
    fn iter_vals(&self) -> impl Iterator<Item = f32>;
    fn distance(&self, other: &impl VecBase) -> f32;
    fn dist2other(&self, other: &Self) -> f32;
    fn dist2many<'a, I>(&'a self, others: I) -> impl Iterator<Item = f32> + 'a
    where
        I: Iterator<Item = &'a Self> + 'a,
        Self: 'a,
    {
        others.map(move |other| self.dist2other(other))
    }
    fn get_vals(&self) -> Vec<f32> { self.iter_vals().collect() }
}

/// generic distance (full.rs:23-29, quant.rs:67-73): one strictly sequential f32 sum over the zipped values
fn full_distance(x: &[f32], y: &[f32]) -> f32 {
    let n = x.len().min(y.len());
    if n == 0 { return 0.0; }
    let mut out = 0f32;
    with_ctx(|c| check(unsafe { sys::hnswb200_dist_full_pairs(c, x.as_ptr(), y.as_ptr(), 1, n as u32, &mut out) }))
        .expect("distance");
    out
}

/// vectors/src/full.rs:3-6
#[derive(Debug, Clone, PartialEq)]
pub struct FullVec {
    pub vector: Vec<f32>,
}
impl VecBase for FullVec {
    fn new(vector: &Vec<f32>) -> Self { FullVec { vector: vector.clone() } }
    fn dim(&self) -> usize { self.vector.len() }
    fn iter_vals(&self) -> impl Iterator<Item = f32> { self.vector.iter().copied() }
    fn distance(&self, other: &impl VecBase) -> f32 { full_distance(&self.vector, &other.get_vals()) }
    fn dist2other(&self, other: &Self) -> f32 { full_distance(&self.vector, &other.vector) }
}

/// vectors/src/quant.rs:6-11 (fields private there too)
#[derive(Debug, Clone, PartialEq)]
pub struct QuantVec {
    delta: f32,
    min: f32,
    codes: Vec<u8>,
}
impl QuantVec {
    pub fn from_parts(delta: f32, min: f32, codes: Vec<u8>) -> QuantVec { QuantVec { delta, min, codes } }
    pub fn parts(&self) -> (f32, f32, &[u8]) { (self.delta, self.min, &self.codes) }
    /// distance_unrolled (quant.rs:14-37) against many others in one device call
    pub fn dist2slice(&self, others: &[&QuantVec]) -> Vec<f32> {
        let dim = self.codes.len();
        let n = others.len() + 1;
        let mut codes = Vec::with_capacity(n * dim);
        let (mut mins, mut deltas) = (Vec::with_capacity(n), Vec::with_capacity(n));
        for v in std::iter::once(self).chain(others.iter().copied()) {
            assert_eq!(v.codes.len(), dim, "vectors of different dimensions");
            codes.extend_from_slice(&v.codes);
            mins.push(v.min);
            deltas.push(v.delta);
        }
        let a = vec![0u32; others.len()];
        let b: Vec<u32> = (1..n as u32).collect();
        let mut out = vec![0f32; others.len()];
        with_ctx(|c| -> Result<(), String> {
            let mut p = std::ptr::null_mut();
            check(unsafe {
                sys::hnswb200_points_upload(c, codes.as_ptr(), mins.as_ptr(), deltas.as_ptr(), std::ptr::null(), n as u64,
                                            dim as u32, &mut p)
            })?;
            let rc = unsafe { sys::hnswb200_dist_pairs(c, p, a.as_ptr(), b.as_ptr(), others.len() as u64, out.as_mut_ptr()) };
            unsafe { sys::hnswb200_points_destroy(p) };
            check(rc)
        })
        .expect("dist2other");
        out
    }
}
impl VecBase for QuantVec {
    /// quant.rs:41-66 on the device (a NaN panics like `partial_cmp().unwrap()` there)
    fn new(vector: &Vec<f32>) -> Self {
        let dim = vector.len();
        let mut codes = vec![0u8; dim];
        let (mut min, mut delta) = (0f32, 0f32);
        with_ctx(|c| check(unsafe { sys::hnswb200_quantise(c, vector.as_ptr(), 1, dim as u32, codes.as_mut_ptr(), &mut min, &mut delta) }))
            .expect("QuantVec::new");
        QuantVec { delta, min, codes }
    }
    fn dim(&self) -> usize { self.codes.len() }
    /// quant.rs:79-84: dequantised values, `code * delta + min` with two roundings
    fn iter_vals(&self) -> impl Iterator<Item = f32> {
        let (d, m) = (self.delta, self.min);
        self.codes.iter().map(move |&c| (c as f32) * d + m)
    }
    fn distance(&self, other: &impl VecBase) -> f32 { full_distance(&self.get_vals(), &other.get_vals()) }
    fn dist2other(&self, other: &Self) -> f32 { self.dist2slice(&[other])[0] }
}

/// vectors/src/quant.rs:90-125: `min f32 | delta f32 | codes`, big-endian
impl Serializer for QuantVec {
    fn size(&self) -> usize { 8 + self.codes.len() }
    fn serialize(&self) -> Vec<u8> {
        let mut b = Vec::with_capacity(self.size());
        b.extend_from_slice(&self.min.to_be_bytes());
        b.extend_from_slice(&self.delta.to_be_bytes());
        b.extend_from_slice(&self.codes);
        b
    }
    fn deserialize(data: Vec<u8>) -> Self {
        let min = f32::from_be_bytes(data[0..4].try_into().unwrap());
        let delta = f32::from_be_bytes(data[4..8].try_into().unwrap());
        QuantVec { delta, min, codes: data[8..].to_vec() }
    }
}
/// vectors/src/full.rs:44-70: `dim` big-endian f32
impl Serializer for FullVec {
    fn size(&self) -> usize { 4 * self.vector.len() }
    fn serialize(&self) -> Vec<u8> { self.vector.iter().flat_map(|v| v.to_be_bytes()).collect() }
    fn deserialize(data: Vec<u8>) -> Self {
        FullVec { vector: data.chunks_exact(4).map(|c| f32::from_be_bytes(c.try_into().unwrap())).collect() }
    }
}

/// vectors/src/lib.rs:29-37 without the `rand` dependency: uniform values in [0, 1) from a xorshift stream (the
/// reference draws from an unseeded thread_rng, so there is nothing to match)
pub fn gen_rand_vecs(dim: usize, n: usize) -> Vec<Vec<f32>> {
    assert!(n > 0);
    let mut s = std::time::SystemTime::now().duration_since(std::time::UNIX_EPOCH).map(|d| d.as_nanos() as u64).unwrap_or(1) | 1;
    let mut next = move || {
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        ((s >> 40) as f32) / (1u64 << 24) as f32
    };
    (0..n).map(|_| (0..dim).map(|_| next()).collect()).collect()
}
