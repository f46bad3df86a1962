//! hnsw/src/helpers (mod.rs of the reference: args, data, glove)
pub mod args;
pub mod glove;
