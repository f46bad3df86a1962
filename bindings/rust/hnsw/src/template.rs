//! `HNSW` with the public entry points of `hnsw/src/template.rs` of Gumo-A/hnsw_rs, backed by libhnsw_b200.so.
//! Same names, argument meaning and error behaviour (`Result<_, String>`; a dimension mismatch panics like
//! `check_points_dim`, template.rs:253-262).
//!
//! Threading: the reference's `HNSW` is `Send + Sync` (adjacency behind `Arc<Mutex<_>>`, graph/src/graph.rs:9).  An
//! engine context serves one host thread at a time, so the context and the index handle live behind a `Mutex`: `&self`
//! methods may be called from many threads, and run one at a time (batch your queries with `ann_batch` instead).
use std::ffi::CString;
use std::path::Path;
use std::ptr;
use std::sync::Mutex;

use graph::{Graph, NodeID};
use hnsw_b200_sys as sys;
use points::Point;
use sys::engine::check;
use vectors::VecBase;

use crate::params::Params;

/// `type VecType = QuantVec;` (points/src/point.rs:4) is a compile-time choice in the reference; here it is the cargo
/// feature `full-vec` (FullVec: f32 vectors, strictly sequential distance) handed to the engine's context.
#[cfg(not(feature = "full-vec"))]
const VEC_TYPE: i32 = sys::HNSWB200_VEC_QUANT;
#[cfg(feature = "full-vec")]
const VEC_TYPE: i32 = sys::HNSWB200_VEC_FULL;

struct Handles {
    ctx: *mut sys::hnswb200_ctx,
    ix: *mut sys::hnswb200_index,
}
unsafe impl Send for Handles {} // the handles are only ever used with the mutex held

pub struct HNSW {
    h: Mutex<Handles>,
    pub params: Params,
}

impl HNSW {
    /// runs `f` with the context and the index locked
    pub fn with_handles<R>(&self, f: impl FnOnce(*mut sys::hnswb200_ctx, *mut sys::hnswb200_index) -> R) -> R {
        let g = self.h.lock().unwrap_or_else(|e| e.into_inner());
        f(g.ctx, g.ix)
    }
    fn refresh_params(&mut self) -> Result<(), String> {
        let mut p = sys::hnswb200_params::default();
        self.with_handles(|_, ix| check(unsafe { sys::hnswb200_index_params(ix, &mut p) }))?;
        self.params = Params::from_c(&p);
        Ok(())
    }

    /// template.rs:133-144
    pub fn new(m: usize, ef_cons: Option<usize>, dim: usize) -> HNSW {
        let params = match ef_cons { Some(e) => Params::from_m_efcons(m, e, dim), None => Params::from_m(m, dim) };
        let mut ctx = ptr::null_mut();
        check(unsafe { sys::hnswb200_ctx_create(0, &mut ctx) }).expect("no CUDA device: this engine has no CPU fallback");
        check(unsafe { sys::hnswb200_ctx_set_vec_type(ctx, VEC_TYPE) }).unwrap();
        let mut ix = ptr::null_mut();
        check(unsafe { sys::hnswb200_build(ctx, ptr::null(), 0, dim as u32, &params.to_c(), ptr::null(), 0, &mut ix) }).unwrap();
        HNSW { h: Mutex::new(Handles { ctx, ix }), params }
    }

    /// template.rs:388-444.  `nb_threads` selects the insertion batch: 1 = the reference's deterministic single-thread
    /// order, > 1 = concurrent inserts against a frozen snapshot (the deterministic analogue of its threaded mode).
    pub fn insert_bulk(mut self, vectors: Vec<Vec<f32>>, nb_threads: usize, _verbose: bool) -> Result<HNSW, String> {
        let dim = self.params.dim;
        for v in &vectors {
            if v.len() != dim {
                panic!("The current index dimension is {}, but tried inserting points of dimension {}", dim, v.len());
            }
        }
        let flat: Vec<f32> = vectors.into_iter().flatten().collect();
        let n = (flat.len() / dim.max(1)) as u64;
        let batch = if nb_threads <= 1 { 1 } else { 0 };
        self.with_handles(|ctx, ix| check(unsafe { sys::hnswb200_index_insert_bulk(ctx, ix, flat.as_ptr(), n, dim as u32, ptr::null(), batch) }))?;
        self.refresh_params()?;
        Ok(self)
    }

    /// template.rs:165-173
    pub fn insert_vec(&mut self, vector: &Vec<f32>) -> Result<NodeID, String> {
        let mut id = 0u32;
        self.with_handles(|ctx, ix| check(unsafe { sys::hnswb200_index_insert_vec(ctx, ix, vector.as_ptr(), vector.len() as u32, &mut id) }))?;
        self.refresh_params()?;
        Ok(id)
    }

    /// template.rs:306-335: ids only, at most `n` of them, ascending (dist, id).
    pub fn ann_by_vector(&self, vector: &Vec<f32>, n: usize, ef: usize) -> Result<Vec<NodeID>, String> {
        Ok(self.ann_batch(std::slice::from_ref(vector), n, ef)?.pop().unwrap())
    }

    /// Many queries per call -- the device boundary sits here (one C-ABI call = one kernel launch).
    pub fn ann_batch(&self, queries: &[Vec<f32>], n: usize, ef: usize) -> Result<Vec<Vec<NodeID>>, String> {
        let dim = self.params.dim;
        // every query is checked BEFORE the flat buffer goes to C (a short query would make the engine read past it)
        for (i, q) in queries.iter().enumerate() {
            if q.len() != dim {
                return Err(format!("query {} has dimension {}, the index has dimension {}", i, q.len(), dim));
            }
        }
        let flat: Vec<f32> = queries.iter().flat_map(|q| q.iter().copied()).collect();
        let nq = queries.len();
        let mut ids = vec![sys::HNSWB200_NO_ID; nq * n];
        let mut counts = vec![0u32; nq];
        self.with_handles(|ctx, ix| {
            check(unsafe {
                sys::hnswb200_search(ctx, ix, flat.as_ptr(), nq as u64, dim as u32, n as u32, ef as u32, ids.as_mut_ptr(),
                                     ptr::null_mut(), counts.as_mut_ptr(), ptr::null())
            })
        })?;
        Ok((0..nq).map(|q| ids[q * n..q * n + counts[q] as usize].to_vec()).collect())
    }

    /// template.rs:146-148
    pub fn len(&self) -> usize { self.with_handles(|_, ix| unsafe { sys::hnswb200_index_len(ix) as usize }) }

    /// template.rs:150-152
    pub fn distance(&self, a: NodeID, b: NodeID) -> Option<f32> {
        if (a as usize) >= self.len() || (b as usize) >= self.len() { return None; }
        let mut out = 0f32;
        self.with_handles(|ctx, ix| {
            let p = unsafe { sys::hnswb200_index_points(ix) };
            check(unsafe { sys::hnswb200_dist_pairs(ctx, p, &a, &b, 1, &mut out) })
        }).ok()?;
        Some(out)
    }

    /// template.rs:154-156 (`Option<&Point>` there; the points live on the device, so an owned copy is returned).
    /// `get_vals()` of the result is what the index stores: the dequantised values of a QuantVec point.
    pub fn get_point(&self, id: NodeID) -> Option<Point> {
        let n = self.len();
        if (id as usize) >= n { return None; }
        let dim = self.params.dim;
        let mut rows = vec![0f32; n * dim];
        let mut levels = vec![0u8; n];
        self.with_handles(|ctx, ix| {
            let p = unsafe { sys::hnswb200_index_points(ix) };
            check(unsafe { sys::hnswb200_points_values(ctx, p, rows.as_mut_ptr(), levels.as_mut_ptr()) })
        }).ok()?;
        let v = rows[id as usize * dim..(id as usize + 1) * dim].to_vec();
        let mut point = Point::new(&v);
        point.id = id;
        point.level = levels[id as usize];
        Some(point)
    }

    /// template.rs:192-194 (`&Graph` there): a read-only snapshot of one layer
    pub fn get_layer(&self, layer_nb: usize) -> Graph {
        self.with_handles(|_, ix| {
            let g = unsafe { sys::hnswb200_index_graph(ix) };
            let nl = unsafe { sys::hnswb200_graph_nb_layers(g) } as usize;
            if layer_nb >= nl { panic!("Could not get layer {layer_nb} of the index."); } // layers.rs:28
            let l = layer_nb as u32;
            let nn = unsafe { sys::hnswb200_graph_layer_nb_nodes(g, l) } as usize;
            let ne = unsafe { sys::hnswb200_graph_layer_nb_edges(g, l) } as usize;
            let cap = unsafe { sys::hnswb200_graph_layer_cap(g, l) } as usize;
            let (mut ids, mut off, mut nb) = (vec![0u32; nn], vec![0u64; nn + 1], vec![0u32; ne.max(1)]);
            check(unsafe { sys::hnswb200_graph_export_layer(g, l, ids.as_mut_ptr(), off.as_mut_ptr(), nb.as_mut_ptr()) }).unwrap();
            Graph::from_csr(layer_nb, cap, &ids, &off, &nb)
        })
    }

    /// template.rs:158-163: prints and returns the degrees of one layer
    pub fn layer_degrees(&self, layer_nb: usize) -> Vec<usize> {
        let g = self.get_layer(layer_nb);
        let d: Vec<usize> = g.nodes.values().map(|v| v.len()).collect();
        println!("layer {layer_nb}: {} nodes, degrees min {:?} max {:?}", d.len(), d.iter().min(), d.iter().max());
        d
    }

    /// template.rs:341-370: degree <= ceil(1.1 * cap) on every layer (caps: mmax0 on layer 0, mmax above)
    pub fn assert_param_compliance(&self) {
        let nl = self.with_handles(|_, ix| unsafe { sys::hnswb200_graph_nb_layers(sys::hnswb200_index_graph(ix)) }) as usize;
        for l in 0..nl {
            let cap = if l == 0 { self.params.mmax0 } else { self.params.mmax };
            let lim = ((cap as f32) * 1.1).ceil() as usize;
            for (id, nb) in self.get_layer(l).nodes.iter() {
                assert!(nb.len() <= lim, "node {id} of layer {l} has degree {} > {lim}", nb.len());
            }
        }
    }

    /// template.rs:43-73 (same directory layout and byte formats)
    pub fn save(&self, index_dir: &Path) -> std::io::Result<()> {
        let c = CString::new(index_dir.to_str().unwrap()).unwrap();
        self.with_handles(|ctx, ix| check(unsafe { sys::hnswb200_index_save_dir(ctx, ix, c.as_ptr()) }))
            .map_err(|e| std::io::Error::new(std::io::ErrorKind::Other, e))
    }

    /// template.rs:75-131
    pub fn load(index_dir: &Path) -> Result<HNSW, String> {
        let mut ctx = ptr::null_mut();
        check(unsafe { sys::hnswb200_ctx_create(0, &mut ctx) })?;
        let c = CString::new(index_dir.to_str().unwrap()).unwrap();
        let mut ix = ptr::null_mut();
        if let Err(e) = check(unsafe { sys::hnswb200_index_load_dir(ctx, c.as_ptr(), &mut ix) }) {
            unsafe { sys::hnswb200_ctx_destroy(ctx) };
            return Err(e);
        }
        let mut p = sys::hnswb200_params::default();
        check(unsafe { sys::hnswb200_index_params(ix, &mut p) })?;
        Ok(HNSW { h: Mutex::new(Handles { ctx, ix }), params: Params::from_c(&p) })
    }
}

impl Drop for HNSW {
    fn drop(&mut self) {
        let g = self.h.get_mut().unwrap_or_else(|e| e.into_inner());
        unsafe {
            sys::hnswb200_index_destroy(g.ix);
            sys::hnswb200_ctx_destroy(g.ctx);
        }
    }
}
