//! Drop-in for the `hnsw` crate of Gumo-A/hnsw_rs: the module paths its callers import
//! (`hnsw::template::HNSW`, `hnsw::helpers::glove::load_glove_array`, `hnsw::helpers::args::parse_args_eval`,
//! `hnsw::params::Params`; eval_glove/src/main.rs:8-15 of the reference) resolve to the B200 engine.
//! Not compiled in the engine's CI (no Rust toolchain there); see INTEGRATION.md.
pub mod helpers;
pub mod params;
pub mod template;
