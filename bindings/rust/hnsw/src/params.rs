//! hnsw/src/params.rs:4-62
use hnsw_b200_sys as sys;

#[derive(Debug, Clone, Copy, PartialEq)]
pub struct Params {
    pub ep: u32,
    pub m: usize,
    pub mmax: usize,
    pub mmax0: usize,
    pub ml: f32,
    pub ef_cons: usize,
    pub dim: usize,
}

/// params.rs:58-62
pub fn get_default_ml(m: usize) -> f32 { 1.0 / (m as f32).ln() }

impl Params {
    /// params.rs:19-30: mmax = M, mmax0 = 2M, ml = 1/ln M, ef_cons = 2M
    pub fn from_m(m: usize, dim: usize) -> Params { Params::from_c(&Self::c_default(m, -1, dim)) }
    /// params.rs:32-44
    pub fn from_m_efcons(m: usize, ef_cons: usize, dim: usize) -> Params { Params::from_c(&Self::c_default(m, ef_cons as i64, dim)) }
    /// params.rs:46-56
    pub fn from(m: usize, ef_cons: Option<usize>, mmax: Option<usize>, mmax0: Option<usize>, ml: Option<f32>, dim: usize) -> Params {
        let mut p = match ef_cons { Some(e) => Self::from_m_efcons(m, e, dim), None => Self::from_m(m, dim) };
        if let Some(v) = mmax { p.mmax = v; }
        if let Some(v) = mmax0 { p.mmax0 = v; }
        if let Some(v) = ml { p.ml = v; }
        p
    }
    fn c_default(m: usize, ef_cons: i64, dim: usize) -> sys::hnswb200_params {
        let mut p = sys::hnswb200_params::default();
        unsafe { sys::hnswb200_params_default(m as u64, ef_cons, dim as u64, &mut p) };
        p
    }
    pub fn from_c(p: &sys::hnswb200_params) -> Params {
        Params { ep: p.ep, m: p.m as usize, mmax: p.mmax as usize, mmax0: p.mmax0 as usize, ml: p.ml, ef_cons: p.ef_cons as usize, dim: p.dim as usize }
    }
    pub fn to_c(&self) -> sys::hnswb200_params {
        sys::hnswb200_params { ep: self.ep, m: self.m as u64, mmax: self.mmax as u64, mmax0: self.mmax0 as u64, ml: self.ml,
                               ef_cons: self.ef_cons as u64, dim: self.dim as u64 }
    }
}
