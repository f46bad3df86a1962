//! points/src/point.rs:4-76
use graph::NodeID;
use vectors::{serializer::Serializer, VecBase};

/// `type VecType = QuantVec;` (point.rs:4) is a compile-time choice in the reference; here it is the cargo feature
/// `full-vec`.
#[cfg(not(feature = "full-vec"))]
pub type VecType = vectors::QuantVec;
#[cfg(feature = "full-vec")]
pub type VecType = vectors::FullVec;

#[derive(Debug, Clone)]
pub struct Point {
    pub id: NodeID,
    pub level: u8,
    vector: VecType,
}

impl Point {
    /// point.rs:13-18
    pub fn with_level_and_id(vector: &Vec<f32>, level: usize, id: usize) -> Point {
        let mut point = Self::new(vector);
        point.id = id as NodeID;
        point.level = level as u8;
        point
    }
    pub fn from_vector(id: NodeID, level: u8, vector: VecType) -> Point { Point { id, level, vector } }
    pub fn vector(&self) -> &VecType { &self.vector }
}

impl VecBase for Point {
    /// point.rs:24-30: id 0 and level 0 by default
    fn new(vector: &Vec<f32>) -> Point { Point { id: 0, level: 0, vector: VecType::new(vector) } }
    fn distance(&self, other: &impl VecBase) -> f32 { self.vector.distance(other) }
    fn dist2other(&self, other: &Self) -> f32 { self.vector.dist2other(&other.vector) } // point.rs:35-37
    fn iter_vals(&self) -> impl Iterator<Item = f32> { self.vector.iter_vals() }
    fn dim(&self) -> usize { self.vector.dim() }
}

/// point.rs:46-76: `level u8 | vector`
impl Serializer for Point {
    fn size(&self) -> usize { 1 + self.vector.size() }
    fn serialize(&self) -> Vec<u8> {
        let mut b = Vec::with_capacity(self.size());
        b.push(self.level);
        b.extend(self.vector.serialize());
        b
    }
    fn deserialize(data: Vec<u8>) -> Self {
        Point { id: 0, level: data[0], vector: VecType::deserialize(data[1..].to_vec()) }
    }
}
