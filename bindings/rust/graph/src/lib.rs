//! Drop-in for the `graph` crate of Gumo-A/hnsw_rs (graph/src/lib.rs, dist.rs, graph.rs, layers.rs, errors.rs).
//! The authoritative adjacency lives in the engine (device rows + host mirror); `Graph` / `Layers` here are read-only
//! snapshots exported from an index (`hnsw::template::HNSW::get_layer`), with the reference's read accessors.
//! Not compiled in the engine's CI (no Rust toolchain there); see INTEGRATION.md.
use std::cmp::Ordering;
use std::collections::BTreeMap;

pub type NodeID = u32; // graph/src/lib.rs:1

/// graph/src/dist.rs:4-37: total order by distance, ties by id; equality needs both fields; NaN panics.
#[derive(Debug, Clone, Copy)]
pub struct Dist {
    pub id: NodeID,
    pub dist: f32,
}
impl Dist {
    pub fn new(id: NodeID, dist: f32) -> Dist { Dist { id, dist } }
}
impl PartialEq for Dist {
    fn eq(&self, other: &Self) -> bool { self.dist == other.dist && self.id == other.id }
}
impl Eq for Dist {}
impl PartialOrd for Dist {
    fn partial_cmp(&self, other: &Self) -> Option<Ordering> { Some(self.cmp(other)) }
}
impl Ord for Dist {
    fn cmp(&self, other: &Self) -> Ordering {
        match self.dist.partial_cmp(&other.dist).unwrap() { // dist.rs:32: NaN panics
            Ordering::Equal => self.id.cmp(&other.id),
            o => o,
        }
    }
}

/// graph/src/errors.rs
#[derive(Debug, Clone, PartialEq, Eq)]
pub enum GraphError {
    NodeNotInGraph(NodeID),
    IsolatedNode(NodeID),
    SelfConnection(NodeID),
    MExceeded(NodeID),
}
impl std::fmt::Display for GraphError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result { write!(f, "{:?}", self) }
}

/// One layer (graph/src/graph.rs:11-16), read-only: node -> ascending neighbour ids.
#[derive(Debug, Clone, Default)]
pub struct Graph {
    pub nodes: BTreeMap<NodeID, Vec<NodeID>>,
    pub level: usize,
    pub m: usize,
}
impl Graph {
    /// from the CSR export of `hnswb200_graph_export_layer`
    pub fn from_csr(level: usize, m: usize, node_ids: &[NodeID], offsets: &[u64], nbrs: &[NodeID]) -> Graph {
        let mut nodes = BTreeMap::new();
        for (i, id) in node_ids.iter().enumerate() {
            nodes.insert(*id, nbrs[offsets[i] as usize..offsets[i + 1] as usize].to_vec());
        }
        Graph { nodes, level, m }
    }
    pub fn iter_nodes(&self) -> impl Iterator<Item = NodeID> + '_ { self.nodes.keys().copied() } // graph.rs:27-29
    pub fn nb_nodes(&self) -> usize { self.nodes.len() }                                        // graph.rs:157-159
    pub fn contains(&self, node_id: NodeID) -> bool { self.nodes.contains_key(&node_id) }        // graph.rs:161-163
    /// graph.rs:103-113
    pub fn neighbors_vec(&self, node_id: NodeID) -> Result<Vec<NodeID>, GraphError> {
        self.nodes.get(&node_id).cloned().ok_or(GraphError::NodeNotInGraph(node_id))
    }
    /// graph.rs:96-101
    pub fn neighbors(&self, node_id: NodeID) -> Result<&Vec<NodeID>, GraphError> {
        self.nodes.get(&node_id).ok_or(GraphError::NodeNotInGraph(node_id))
    }
    /// graph.rs:146-155
    pub fn degree(&self, node_id: NodeID) -> Result<usize, GraphError> { Ok(self.neighbors(node_id)?.len()) }
}

/// graph/src/layers.rs:13-71, read-only
#[derive(Debug, Clone, Default)]
pub struct Layers {
    layers: Vec<Graph>,
}
impl Layers {
    pub fn from_graphs(layers: Vec<Graph>) -> Layers { Layers { layers } }
    pub fn len(&self) -> usize { self.layers.len() }
    pub fn is_empty(&self) -> bool { self.layers.is_empty() }
    /// layers.rs:25-30: a missing layer panics
    pub fn get_layer(&self, layer_nb: usize) -> &Graph {
        self.layers.get(layer_nb).unwrap_or_else(|| panic!("Could not get layer {layer_nb} of the index."))
    }
    pub fn iter_layers(&self) -> impl Iterator<Item = &Graph> { self.layers.iter() }
}
