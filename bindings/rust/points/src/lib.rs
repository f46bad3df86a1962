//! Drop-in for the `points` crate of Gumo-A/hnsw_rs (points/src/point.rs, points/src/points.rs).
//! Not compiled in the engine's CI (no Rust toolchain there); see INTEGRATION.md.
pub mod point;
pub mod points;

pub use point::Point;
pub use points::{new_layer, Points, SimplePoints};
