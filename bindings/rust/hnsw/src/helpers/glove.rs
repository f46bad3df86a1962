//! hnsw/src/helpers/glove.rs:14-109: GloVe text loader and brute-force ground truth.
use std::collections::HashMap;
use std::fs::File;
use std::io::{BufRead, BufReader};

use graph::NodeID;
use hnsw_b200_sys as sys;
use sys::engine::check;

use crate::template::HNSW;

/// glove.rs:14-71: `word v1 .. vd` rows; at most `lim` of them (0 = all).  Values are parsed straight to f32.
pub fn load_glove_array(lim: usize, file: File, _verbose: bool) -> Result<(Vec<String>, Vec<Vec<f32>>), String> {
    let reader = BufReader::new(file);
    let (mut words, mut embeddings): (Vec<String>, Vec<Vec<f32>>) = (Vec::new(), Vec::new());
    for (idx, line) in reader.lines().enumerate() {
        if lim > 0 && idx >= lim { break; }
        let line = line.map_err(|e| e.to_string())?;
        let mut parts = line.split_whitespace();
        let word = match parts.next() { Some(w) => w.to_string(), None => continue };
        let v: Vec<f32> = parts.filter_map(|x| x.parse::<f32>().ok()).collect();
        if let Some(first) = embeddings.first() {
            if first.len() != v.len() {
                return Err(format!("Line {}: vector is not the same size as others.", idx + 1));
            }
        }
        words.push(word);
        embeddings.push(v);
    }
    Ok((words, embeddings))
}

/// glove.rs:73-92: {query id -> ids of its `nb_nns` nearest stored points}, exact under the metric of the index's
/// points with (dist, id) order (the tensor-core filter + exact re-rank of the engine; ids equal the reference's
/// full sort, glove.rs:99-109)
pub fn brute_force_nns(nb_nns: usize, index: &HNSW, test_vectors: &[Vec<f32>], ids: &[usize]) -> Result<HashMap<NodeID, Vec<NodeID>>, String> {
    let dim = index.params.dim;
    let mut flat = Vec::with_capacity(ids.len() * dim);
    for &i in ids {
        if test_vectors[i].len() != dim { return Err(format!("query {i} has dimension {}, the index {dim}", test_vectors[i].len())); }
        flat.extend_from_slice(&test_vectors[i]);
    }
    let mut out = vec![sys::HNSWB200_NO_ID; ids.len() * nb_nns];
    index.with_handles(|ctx, ix| {
        let p = unsafe { sys::hnswb200_index_points(ix) };
        check(unsafe { sys::hnswb200_bruteforce_topk(ctx, p, flat.as_ptr(), ids.len() as u64, nb_nns as u32, 0, out.as_mut_ptr(), std::ptr::null_mut()) })
    })?;
    Ok(ids.iter().enumerate()
        .map(|(j, &i)| (i as NodeID, out[j * nb_nns..(j + 1) * nb_nns].iter().copied().filter(|&x| x != sys::HNSWB200_NO_ID).collect()))
        .collect())
}
