//! points/src/points.rs:16-160: the `Points` trait and `SimplePoints`.  The host-side collection keeps the reference's
//! shape (`pub collection: Vec<Point>`); an index keeps its own device-resident copy (hnswb200_points).
use graph::NodeID;
use hnsw_b200_sys as sys;
use sys::engine::{check, with_ctx};
use vectors::VecBase;

use crate::point::Point;

/// points.rs:16-31
pub trait Points {
    fn new(vecs: Vec<Vec<f32>>, ml: f32) -> Self;
    fn len(&self) -> usize;
    fn ids(&self) -> impl Iterator<Item = NodeID>;
    fn dim(&self) -> Option<usize>;
    fn push(&mut self, point: Point) -> NodeID;
    fn extend(&mut self, other: Self) -> Vec<NodeID>;
    fn get_point(&self, idx: NodeID) -> Option<&Point>;
    fn get_points_iter<I>(&self, indices: I) -> impl Iterator<Item = &Point>
    where
        I: Iterator<Item = NodeID>;
    fn distance(&self, a_idx: NodeID, b_idx: NodeID) -> Option<f32>;
    fn distance2point(&self, point: &Point, idx: NodeID) -> Option<f32>;
}

/// points.rs:33-36
#[derive(Debug, Clone)]
pub struct SimplePoints {
    pub collection: Vec<Point>,
}

/// points.rs:148-160: `floor(-ln(u) * ml)` for a uniform u in (0, 1)
pub fn new_layer(ml: f32, next_uniform: &mut impl FnMut() -> f32) -> usize {
    let mut rand_nb = 0.0f32;
    while rand_nb == 0.0 || rand_nb == 1.0 {
        rand_nb = next_uniform();
    }
    (-rand_nb.ln() * ml).floor() as usize
}

impl Points for SimplePoints {
    /// points.rs:39-48.  Quantises every row on the device in one call; levels follow the engine's seeded generator
    /// (the reference's StdRng stream is not pinned by any of its tests).
    fn new(vecs: Vec<Vec<f32>>, ml: f32) -> Self {
        let mut s = 0x9E37_79B9_7F4A_7C15u64; // seed 0 of a fixed xorshift stream
        let mut uni = move || {
            s ^= s << 13;
            s ^= s >> 7;
            s ^= s << 17;
            ((s >> 40) as f32) / (1u64 << 24) as f32
        };
        let collection = vecs
            .iter()
            .enumerate()
            .map(|(idx, v)| Point::with_level_and_id(v, new_layer(ml, &mut uni), idx))
            .collect();
        SimplePoints { collection }
    }
    fn len(&self) -> usize { self.collection.len() }
    fn ids(&self) -> impl Iterator<Item = NodeID> { self.collection.iter().map(|p| p.id) }
    fn dim(&self) -> Option<usize> { self.collection.first().map(|p| p.dim()) }
    fn push(&mut self, mut point: Point) -> NodeID {
        point.id = self.len() as NodeID;
        let id = point.id;
        self.collection.push(point);
        id
    }
    fn extend(&mut self, other: Self) -> Vec<NodeID> { other.collection.into_iter().map(|p| self.push(p)).collect() }
    fn get_point(&self, idx: NodeID) -> Option<&Point> { self.collection.get(idx as usize) }
    fn get_points_iter<I>(&self, indices: I) -> impl Iterator<Item = &Point>
    where
        I: Iterator<Item = NodeID>,
    {
        indices.filter_map(|i| self.collection.get(i as usize))
    }
    /// points.rs:86-93
    fn distance(&self, a_idx: NodeID, b_idx: NodeID) -> Option<f32> {
        Some(self.get_point(a_idx)?.dist2other(self.get_point(b_idx)?))
    }
    /// points.rs:95-101
    fn distance2point(&self, point: &Point, idx: NodeID) -> Option<f32> { Some(point.dist2other(self.get_point(idx)?)) }
}

impl SimplePoints {
    /// distance2point for many ids in one device call (VecBase::dist2many, vectors/src/lib.rs:17-22): the f32 query
    /// becomes a point like a stored one and is compared with `ids` of a device-resident point set
    pub fn dist_query_many(points: *const sys::hnswb200_points, query: &[f32], ids: &[NodeID]) -> Result<Vec<f32>, String> {
        let mut out = vec![0f32; ids.len()];
        with_ctx(|c| check(unsafe { sys::hnswb200_dist_query_many(c, points, query.as_ptr(), ids.as_ptr(), ids.len() as u64, out.as_mut_ptr()) }))?;
        Ok(out)
    }
}
