//! `HNSW` with the public entry points of `hnsw/src/template.rs` of Gumo-A/hnsw_rs, backed by
//! libhnsw_b200.so.  Same names, argument meaning and error behaviour (`Result<_, String>`;
//! a dimension mismatch panics like `check_points_dim`, template.rs:253-262).
//! Not compiled in the engine's CI (no Rust toolchain there); see INTEGRATION.md.
use hnsw_b200_sys as sys;
use std::ffi::{CStr, CString};
use std::path::Path;
use std::ptr;

pub type NodeID = u32; // graph/src/lib.rs:1

fn last_error() -> String {
    unsafe { CStr::from_ptr(sys::hnswb200_last_error()).to_string_lossy().into_owned() }
}
fn check(rc: i32) -> Result<(), String> {
    if rc == sys::HNSWB200_OK { Ok(()) } else { Err(last_error()) }
}

/// hnsw/src/params.rs:4-12
pub type Params = sys::hnswb200_params;

/// `type VecType = QuantVec;` (points/src/point.rs:4) is a compile-time choice in the reference; here it is the cargo
/// feature `full-vec` (FullVec: f32 vectors, strictly sequential distance) handed to the engine's context.
#[cfg(not(feature = "full-vec"))]
const VEC_TYPE: i32 = sys::HNSWB200_VEC_QUANT;
#[cfg(feature = "full-vec")]
const VEC_TYPE: i32 = sys::HNSWB200_VEC_FULL;

pub struct HNSW {
    ctx: *mut sys::hnswb200_ctx,
    ix: *mut sys::hnswb200_index,
    pub params: Params,
}
unsafe impl Send for HNSW {}
unsafe impl Sync for HNSW {} // read-only calls are thread-safe per handle

impl HNSW {
    /// template.rs:133-144
    pub fn new(m: usize, ef_cons: Option<usize>, dim: usize) -> HNSW {
        let mut params = Params::default();
        unsafe { sys::hnswb200_params_default(m as u64, ef_cons.map(|e| e as i64).unwrap_or(-1), dim as u64, &mut params) };
        let mut ctx = ptr::null_mut();
        check(unsafe { sys::hnswb200_ctx_create(0, &mut ctx) }).expect("no CUDA device: this engine has no CPU fallback");
        check(unsafe { sys::hnswb200_ctx_set_vec_type(ctx, VEC_TYPE) }).unwrap();
        let mut ix = ptr::null_mut();
        check(unsafe { sys::hnswb200_build(ctx, ptr::null(), 0, dim as u32, &params, ptr::null(), 0, &mut ix) }).unwrap();
        HNSW { ctx, ix, params }
    }

    /// template.rs:388-444.  `nb_threads` selects the insertion batch: 1 = the reference's
    /// deterministic single-thread order, > 1 = concurrent inserts against a frozen snapshot.
    pub fn insert_bulk(mut self, vectors: Vec<Vec<f32>>, nb_threads: usize, _verbose: bool) -> Result<HNSW, String> {
        let dim = self.params.dim as usize;
        for v in &vectors {
            if v.len() != dim {
                panic!("The current index dimension is {}, but tried inserting points of dimension {}", dim, v.len());
            }
        }
        let flat: Vec<f32> = vectors.into_iter().flatten().collect();
        let n = (flat.len() / dim.max(1)) as u64;
        let batch = if nb_threads <= 1 { 1 } else { 0 };
        check(unsafe { sys::hnswb200_index_insert_bulk(self.ctx, self.ix, flat.as_ptr(), n, dim as u32, ptr::null(), batch) })?;
        check(unsafe { sys::hnswb200_index_params(self.ix, &mut self.params) })?;
        Ok(self)
    }

    /// template.rs:165-173
    pub fn insert_vec(&mut self, vector: &Vec<f32>) -> Result<NodeID, String> {
        let mut id = 0u32;
        check(unsafe { sys::hnswb200_index_insert_vec(self.ctx, self.ix, vector.as_ptr(), vector.len() as u32, &mut id) })?;
        check(unsafe { sys::hnswb200_index_params(self.ix, &mut self.params) })?;
        Ok(id)
    }

    /// template.rs:306-335: ids only, at most `n` of them, ascending (dist, id).
    pub fn ann_by_vector(&self, vector: &Vec<f32>, n: usize, ef: usize) -> Result<Vec<NodeID>, String> {
        Ok(self.ann_batch(std::slice::from_ref(vector), n, ef)?.pop().unwrap())
    }

    /// Many queries per call -- the device boundary sits here (one C-ABI call = one kernel launch).
    pub fn ann_batch(&self, queries: &[Vec<f32>], n: usize, ef: usize) -> Result<Vec<Vec<NodeID>>, String> {
        let dim = self.params.dim as usize;
        let flat: Vec<f32> = queries.iter().flat_map(|q| q.iter().copied()).collect();
        let nq = queries.len();
        let mut ids = vec![sys::HNSWB200_NO_ID; nq * n];
        let mut counts = vec![0u32; nq];
        check(unsafe {
            sys::hnswb200_search(self.ctx, self.ix, flat.as_ptr(), nq as u64, dim as u32, n as u32, ef as u32,
                                 ids.as_mut_ptr(), ptr::null_mut(), counts.as_mut_ptr(), ptr::null())
        })?;
        Ok((0..nq).map(|q| ids[q * n..q * n + counts[q] as usize].to_vec()).collect())
    }

    pub fn len(&self) -> usize { unsafe { sys::hnswb200_index_len(self.ix) as usize } } // template.rs:146-148

    /// template.rs:150-152
    pub fn distance(&self, a: NodeID, b: NodeID) -> Option<f32> {
        if (a as usize) >= self.len() || (b as usize) >= self.len() { return None; }
        let mut out = 0f32;
        let p = unsafe { sys::hnswb200_index_points(self.ix) };
        check(unsafe { sys::hnswb200_dist_pairs(self.ctx, p, &a, &b, 1, &mut out) }).ok()?;
        Some(out)
    }

    /// template.rs:43-73 (same directory layout and byte formats)
    pub fn save(&self, index_dir: &Path) -> std::io::Result<()> {
        let c = CString::new(index_dir.to_str().unwrap()).unwrap();
        check(unsafe { sys::hnswb200_index_save_dir(self.ctx, self.ix, c.as_ptr()) })
            .map_err(|e| std::io::Error::new(std::io::ErrorKind::Other, e))
    }

    /// template.rs:75-131
    pub fn load(index_dir: &Path) -> Result<HNSW, String> {
        let mut ctx = ptr::null_mut();
        check(unsafe { sys::hnswb200_ctx_create(0, &mut ctx) })?;
        let c = CString::new(index_dir.to_str().unwrap()).unwrap();
        let mut ix = ptr::null_mut();
        check(unsafe { sys::hnswb200_index_load_dir(ctx, c.as_ptr(), &mut ix) })?;
        let mut params = Params::default();
        check(unsafe { sys::hnswb200_index_params(ix, &mut params) })?;
        Ok(HNSW { ctx, ix, params })
    }
}

impl Drop for HNSW {
    fn drop(&mut self) {
        unsafe {
            sys::hnswb200_index_destroy(self.ix);
            sys::hnswb200_ctx_destroy(self.ctx);
        }
    }
}

/// hnsw/src/helpers/glove.rs:73-92: exact top-k under the quantised metric with (dist, id) order.
pub fn brute_force_nns(index: &HNSW, queries: &[Vec<f32>], nb_nns: usize) -> Result<Vec<Vec<NodeID>>, String> {
    let flat: Vec<f32> = queries.iter().flat_map(|q| q.iter().copied()).collect();
    let mut ids = vec![sys::HNSWB200_NO_ID; queries.len() * nb_nns];
    let p = unsafe { sys::hnswb200_index_points(index.ix) };
    check(unsafe {
        sys::hnswb200_bruteforce_topk(index.ctx, p, flat.as_ptr(), queries.len() as u64, nb_nns as u32, 0,
                                      ids.as_mut_ptr(), ptr::null_mut())
    })?;
    Ok(ids.chunks(nb_nns).map(|c| c.iter().copied().filter(|&i| i != sys::HNSWB200_NO_ID).collect()).collect())
}
