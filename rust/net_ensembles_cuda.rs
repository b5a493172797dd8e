//! `net_ensembles_cuda` — the module a net_ensembles maintainer adds to call the B200-native
//! `vertex_load` (SOURCE ONLY: no Rust toolchain exists in the build image, so this file has not
//! been compiled; it mirrors include/net_ensembles_cuda.h one to one and is deliberately tiny).
//!
//! It sits beside `MeasurableGraphQuantities` (src/traits/graph_traits.rs:161-331): the same
//! bounds and the same blanket impl, so `Graph`, `SwGraph`, `ErEnsembleC`, `ErEnsembleM`,
//! `SwEnsemble`, `BAensemble` (and every other `AsRef<GenericGraph<T, A>>`) get
//! `vertex_load_cuda(include_endpoints)` with the signature of `vertex_load`
//! (src/generic_graph/generic_graph.rs:1310).
//!
//! Cargo: `[features] cuda = []`, `build.rs`: `println!("cargo:rustc-link-lib=dylib=net_ensembles_cuda");`
#![cfg(feature = "cuda")]
use crate::{traits::*, GenericGraph};
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct NecCtx { _private: [u8; 0] }
#[repr(C)] pub struct NecGraph { _private: [u8; 0] }

extern "C" {
    fn nec_init(out: *mut *mut NecCtx, devices: *const c_int, n_devices: c_int) -> c_int;
    fn nec_destroy(ctx: *mut NecCtx);
    fn nec_last_error(ctx: *const NecCtx) -> *const c_char;
    fn nec_graph_upload(ctx: *mut NecCtx, n: u32, nnz: u64, row_ptr: *const u64, col_idx: *const u32,
                        out: *mut *mut NecGraph) -> c_int;
    fn nec_graph_free(ctx: *mut NecCtx, g: *mut NecGraph);
    fn nec_vertex_load(ctx: *mut NecCtx, g: *const NecGraph, include_endpoints: c_int, out: *mut f64) -> c_int;
    fn nec_vertex_load_sources(ctx: *mut NecCtx, g: *const NecGraph, sources: *const u32, n_sources: u32,
                               include_endpoints: c_int, out: *mut f64) -> c_int;
    fn nec_distance_measures(ctx: *mut NecCtx, g: *const NecGraph, sources: *const u32, n_sources: u32,
                             ecc: *mut u32, sum_dist: *mut u64, reached: *mut u32) -> c_int;
    #[allow(dead_code)]
    fn nec_set_stream(ctx: *mut NecCtx, cuda_stream: *mut c_void) -> c_int;
    fn nec_wave_size(ctx: *mut NecCtx, g: *const NecGraph, n_sources_total: u64, sources_per_wave: *mut u32) -> c_int;
}

/// One GPU context (device, stream, workspace).  Not `Sync`: one per thread, like the C ABI says.
pub struct CudaContext { raw: *mut NecCtx }

impl CudaContext {
    /// Panics if no usable B200 is present — there is no CPU fallback by design.
    pub fn new(device: i32) -> Self {
        let mut raw = std::ptr::null_mut();
        let dev: c_int = device;
        let rc = unsafe { nec_init(&mut raw, &dev, 1) };
        if rc != 0 { panic!("nec_init failed ({}): {}", rc, last_error(std::ptr::null())); }
        CudaContext { raw }
    }
}
impl Drop for CudaContext { fn drop(&mut self) { unsafe { nec_destroy(self.raw) } } }

fn last_error(ctx: *const NecCtx) -> String {
    unsafe { CStr::from_ptr(nec_last_error(ctx)).to_string_lossy().into_owned() }
}

/// Order-preserving CSR of the adjacency lists: the flattening of `container(i).neighbors()`
/// (graph_traits.rs:72; graph.rs:62-64; sw_graph.rs:195-199).  Once per sampled graph.
fn export_csr<T, A>(g: &GenericGraph<T, A>) -> (Vec<u64>, Vec<u32>)
where T: Node, A: AdjContainer<T>
{
    let n = g.vertex_count();
    assert!(n < u32::MAX as usize);
    let mut row_ptr = Vec::with_capacity(n + 1);
    let mut col_idx = Vec::with_capacity(2 * g.edge_count());
    row_ptr.push(0u64);
    for c in g.container_iter() {
        col_idx.extend(c.neighbors().map(|&j| j as u32));
        row_ptr.push(col_idx.len() as u64);
    }
    (row_ptr, col_idx)
}

/// `diameter()` (generic_graph.rs:1116-1136) and `closeness_centrality()` (:1387-1403) on the GPU.
pub fn diameter_and_closeness_cuda<T, A>(ctx: &CudaContext, g: &GenericGraph<T, A>) -> (Option<usize>, Vec<f64>)
where T: Node, A: AdjContainer<T>
{
    let n = g.vertex_count();
    if n == 0 { return (None, Vec::new()); }
    let (row_ptr, col_idx) = export_csr(g);
    let sources: Vec<u32> = (0..n as u32).collect();
    let (mut ecc, mut sum, mut reached) = (vec![0u32; n], vec![0u64; n], vec![0u32; n]);
    unsafe {
        let mut dg = std::ptr::null_mut();
        let rc = nec_graph_upload(ctx.raw, n as u32, col_idx.len() as u64, row_ptr.as_ptr(), col_idx.as_ptr(), &mut dg);
        if rc != 0 { panic!("nec_graph_upload failed ({}): {}", rc, last_error(ctx.raw)); }
        let rc = nec_distance_measures(ctx.raw, dg, sources.as_ptr(), n as u32, ecc.as_mut_ptr(), sum.as_mut_ptr(), reached.as_mut_ptr());
        nec_graph_free(ctx.raw, dg);
        if rc != 0 { panic!("nec_distance_measures failed ({}): {}", rc, last_error(ctx.raw)); }
    }
    let diameter = if reached[0] as usize == n { Some(ecc[1..].iter().copied().max().unwrap_or(0) as usize) } else { None };
    let val = (n - 1) as f64;
    (diameter, sum.iter().map(|&c| val / c as f64).collect())
}

pub trait VertexLoadCuda {
    /// Drop-in for `vertex_load(include_endpoints)`; returns a fresh `Vec<f64>` of length N.
    fn vertex_load_cuda(&self, ctx: &CudaContext, include_endpoints: bool) -> Vec<f64>;
    /// Only the contributions of the listed sources (sharding / sampling primitive).
    fn vertex_load_cuda_sources(&self, ctx: &CudaContext, sources: &[u32], include_endpoints: bool) -> Vec<f64>;
}

impl<T, A, E> VertexLoadCuda for E
where T: Node, A: AdjContainer<T>, GenericGraph<T, A>: Clone, E: AsRef<GenericGraph<T, A>>
{
    fn vertex_load_cuda(&self, ctx: &CudaContext, include_endpoints: bool) -> Vec<f64> {
        let g = self.as_ref();
        let (row_ptr, col_idx) = export_csr(g);
        let mut out = vec![0.0f64; g.vertex_count()];
        unsafe {
            let mut dg = std::ptr::null_mut();
            let rc = nec_graph_upload(ctx.raw, out.len() as u32, col_idx.len() as u64, row_ptr.as_ptr(), col_idx.as_ptr(), &mut dg);
            if rc != 0 { panic!("nec_graph_upload failed ({}): {}", rc, last_error(ctx.raw)); }
            let rc = nec_vertex_load(ctx.raw, dg, include_endpoints as c_int, out.as_mut_ptr());
            nec_graph_free(ctx.raw, dg);
            if rc != 0 { panic!("nec_vertex_load failed ({}): {}", rc, last_error(ctx.raw)); }
        }
        out
    }

    fn vertex_load_cuda_sources(&self, ctx: &CudaContext, sources: &[u32], include_endpoints: bool) -> Vec<f64> {
        let g = self.as_ref();
        let (row_ptr, col_idx) = export_csr(g);
        let mut out = vec![0.0f64; g.vertex_count()];
        unsafe {
            let mut dg = std::ptr::null_mut();
            let rc = nec_graph_upload(ctx.raw, out.len() as u32, col_idx.len() as u64, row_ptr.as_ptr(), col_idx.as_ptr(), &mut dg);
            if rc != 0 { panic!("nec_graph_upload failed ({}): {}", rc, last_error(ctx.raw)); }
            let rc = nec_vertex_load_sources(ctx.raw, dg, sources.as_ptr(), sources.len() as u32,
                                             include_endpoints as c_int, out.as_mut_ptr());
            nec_graph_free(ctx.raw, dg);
            if rc != 0 { panic!("nec_vertex_load_sources failed ({}): {}", rc, last_error(ctx.raw)); }
        }
        out
    }
}
