//! `net_ensembles_cuda` — the module a net_ensembles maintainer adds to call the B200-native
//! `vertex_load`.  UNTESTED SOURCE: the build image has no `rustc`/`cargo`, so this file has never
//! been compiled; what IS compiled and tested is the C header it binds (include/net_ensembles_cuda.h,
//! exercised from C++ and through ctypes).  Every function of the header has its `extern "C"`
//! declaration below (tests/test_cabi.py checks the list against the header).
//!
//! It sits beside `MeasurableGraphQuantities<G>` (src/traits/graph_traits.rs:161-331) and copies its
//! shape: a trait generic over the graph type with the same blanket impl for every
//! `E: AsRef<GenericGraph<T, A>>` (graph_traits.rs:272-278), so `Graph`, `SwGraph`, `ErEnsembleC`,
//! `ErEnsembleM`, `SwEnsemble`, `BAensemble` (and every other ensemble) get
//! `vertex_load_cuda(ctx, include_endpoints)` with the meaning of `vertex_load`
//! (src/generic_graph/generic_graph.rs:1310).
//!
//! Cargo: `[features] cuda = []`, `build.rs`: `println!("cargo:rustc-link-lib=dylib=net_ensembles_cuda");`
#![cfg(feature = "cuda")]
use crate::{traits::*, GenericGraph};
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct NecCtx { _private: [u8; 0] }
#[repr(C)] pub struct NecGraph { _private: [u8; 0] }
#[repr(C)] pub struct NecChains { _private: [u8; 0] }

/// `nec_stats` of the header, field for field.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct NecStats {
    pub n_sources: u64, pub n_batches: u64, pub reached_pairs: u64, pub edge_pairs: u64, pub vertex_visits: u64,
    pub max_depth: u32, pub kernel_launches: u32, pub wide_batches: u32, pub n_slots: u32,
    pub kernel_ms: f32, pub engine_ms: f32, pub frac_reset: f32, pub frac_forward: f32, pub frac_backward: f32,
    pub workspace_bytes: u64, pub engine: u32, pub batch_width: u32, pub team_size: u32, pub cta_threads: u32,
    pub pull_levels: u32, pub push_levels: u32, pub requeued_batches: u32, pub reserved0: u32,
}

extern "C" {
    fn nec_version() -> *const c_char;
    fn nec_init(out: *mut *mut NecCtx, devices: *const c_int, n_devices: c_int) -> c_int;
    fn nec_destroy(ctx: *mut NecCtx);
    fn nec_last_error(ctx: *const NecCtx) -> *const c_char;
    fn nec_set_stream(ctx: *mut NecCtx, cuda_stream: *mut c_void) -> c_int;
    fn nec_set_workspace_limit(ctx: *mut NecCtx, bytes: u64) -> c_int;
    fn nec_set_engine(ctx: *mut NecCtx, engine: c_int) -> c_int;
    fn nec_graph_upload(ctx: *mut NecCtx, n: u32, nnz: u64, row_ptr: *const u64, col_idx: *const u32,
                        out: *mut *mut NecGraph) -> c_int;
    fn nec_graph_upload_adjacency(ctx: *mut NecCtx, n: u32, rows: *const *const c_void, degrees: *const u64,
                                  elem_stride: u32, to_offset: u32, to_width: u32, out: *mut *mut NecGraph) -> c_int;
    fn nec_graph_free(ctx: *mut NecCtx, g: *mut NecGraph);
    fn nec_graph_dims(g: *const NecGraph, n: *mut u32, nnz: *mut u64) -> c_int;
    fn nec_graph_download(ctx: *mut NecCtx, g: *const NecGraph, row_ptr: *mut u64, col_idx: *mut u32) -> c_int;
    fn nec_erc_randomize(ctx: *mut NecCtx, n: u32, prob: f64, rng_in: *const u64, rng_out: *mut u64,
                         out: *mut *mut NecGraph) -> c_int;
    fn nec_vertex_load(ctx: *mut NecCtx, g: *const NecGraph, include_endpoints: c_int, out: *mut f64) -> c_int;
    fn nec_vertex_load_sources(ctx: *mut NecCtx, g: *const NecGraph, sources: *const u32, n_sources: u32,
                               include_endpoints: c_int, out: *mut f64) -> c_int;
    fn nec_vertex_load_sources_device(ctx: *mut NecCtx, g: *const NecGraph, sources: *const u32, n_sources: u32,
                                      include_endpoints: c_int, d_out: *mut f64) -> c_int;
    fn nec_vertex_load_batch(ctx: *mut NecCtx, n_graphs: u32, n_per_graph: *const u32, row_ptr_offsets: *const u64,
                             row_ptr_concat: *const u32, col_offsets: *const u64, col_idx_concat: *const u32,
                             include_endpoints: c_int, out_concat: *mut f64) -> c_int;
    fn nec_chains_create(ctx: *mut NecCtx, n_graphs: u32, n_per_graph: *const u32, row_ptr_offsets: *const u64,
                         row_ptr_concat: *const u32, col_offsets: *const u64, col_idx_concat: *const u32,
                         slack: u32, out: *mut *mut NecChains) -> c_int;
    fn nec_chains_step_and_load(ctx: *mut NecCtx, chains: *mut NecChains, op_offsets: *const u32, ops: *const u64,
                                include_endpoints: c_int, out_concat: *mut f64) -> c_int;
    fn nec_chains_result(chains: *const NecChains) -> *const f64;
    fn nec_chains_download(ctx: *mut NecCtx, chains: *mut NecChains, row_ptr_concat: *mut u32, col_idx_concat: *mut u32,
                           col_capacity: u64, nnz: *mut u64) -> c_int;
    fn nec_chains_free(ctx: *mut NecCtx, chains: *mut NecChains);
    fn nec_distance_measures(ctx: *mut NecCtx, g: *const NecGraph, sources: *const u32, n_sources: u32,
                             ecc: *mut u32, sum_dist: *mut u64, reached: *mut u32) -> c_int;
    fn nec_wave_size(ctx: *mut NecCtx, g: *const NecGraph, n_sources_total: u64, sources_per_wave: *mut u32) -> c_int;
    fn nec_bfs_debug(ctx: *mut NecCtx, g: *const NecGraph, source: u32, dist: *mut u32, npred: *mut u32,
                     sigma: *mut f64, bk: *mut f64) -> c_int;
    fn nec_get_stats(ctx: *const NecCtx, out: *mut NecStats) -> c_int;
}

/// One context: a GPU (device, stream, workspace) or, with `with_devices`, several GPUs of the box
/// driven by this one process (sources dealt to the devices, one NCCL all-reduce behind the C ABI).
/// Not `Sync`: one per thread, like the C ABI says.
pub struct CudaContext { raw: *mut NecCtx }

impl CudaContext {
    /// Panics if no usable B200 is present — there is no CPU fallback by design.
    pub fn new(device: i32) -> Self { Self::with_devices(&[device]) }
    pub fn with_devices(devices: &[i32]) -> Self {
        let mut raw = std::ptr::null_mut();
        let devs: Vec<c_int> = devices.iter().map(|&d| d as c_int).collect();
        let rc = unsafe { nec_init(&mut raw, devs.as_ptr(), devs.len() as c_int) };
        if rc != 0 { panic!("nec_init failed ({}): {}", rc, last_error(std::ptr::null())); }
        CudaContext { raw }
    }
    pub fn version() -> String { unsafe { CStr::from_ptr(nec_version()).to_string_lossy().into_owned() } }
    pub fn stats(&self) -> NecStats {
        let mut s = NecStats::default();
        unsafe { nec_get_stats(self.raw, &mut s) };
        s
    }
    /// 0 auto, 1 source-minor batched engine, 2 one-source-per-CTA engine, 3 compact batched engine (tests / profiling).
    pub fn set_engine(&self, engine: i32) { self.check(unsafe { nec_set_engine(self.raw, engine) }, "nec_set_engine") }
    pub fn set_workspace_limit(&self, bytes: u64) { self.check(unsafe { nec_set_workspace_limit(self.raw, bytes) }, "nec_set_workspace_limit") }
    /// Run on a caller's cudaStream_t (single-device contexts).
    pub unsafe fn set_stream(&self, cuda_stream: *mut c_void) { self.check(nec_set_stream(self.raw, cuda_stream), "nec_set_stream") }
    fn check(&self, rc: c_int, what: &str) {
        if rc != 0 { panic!("{} failed ({}): {}", what, rc, last_error(self.raw)); }
    }
}
impl Drop for CudaContext { fn drop(&mut self) { unsafe { nec_destroy(self.raw) } } }

fn last_error(ctx: *const NecCtx) -> String {
    unsafe { CStr::from_ptr(nec_last_error(ctx)).to_string_lossy().into_owned() }
}

/// A sampled graph on the device ("export once per sampled graph, measure many times").
pub struct DeviceGraph<'c> { ctx: &'c CudaContext, raw: *mut NecGraph, n: usize }

impl<'c> DeviceGraph<'c> {
    /// Export fast path: every container keeps its neighbours in one contiguous slice
    /// (`AdjList::edges()`: `&[usize]` src/graph.rs:145-150, `&[SwEdge]` src/sw_graph.rs:164-170); the slices
    /// are handed over as they lie and the library flattens / narrows / validates them into pinned staging.
    /// `to_of` maps an edge to the address of its neighbour id (`|e: &usize| e`, `|e: &SwEdge| e.to_ref()`).
    pub fn from_graph<T, A, Edge, F>(ctx: &'c CudaContext, g: &GenericGraph<T, A>, to_of: F) -> Self
    where T: Node, A: AdjContainer<T> + AdjList<Edge>, F: Fn(&Edge) -> &usize
    {
        let n = g.vertex_count();
        assert!(n < u32::MAX as usize);
        let mut rows: Vec<*const c_void> = Vec::with_capacity(n);
        let mut degrees: Vec<u64> = Vec::with_capacity(n);
        let mut to_offset = 0usize;
        for c in g.container_iter() {
            let edges: &[Edge] = c.edges();
            if let Some(first) = edges.first() {
                to_offset = (to_of(first) as *const usize as usize) - (first as *const Edge as usize);
            }
            rows.push(edges.as_ptr() as *const c_void);
            degrees.push(edges.len() as u64);
        }
        let mut raw = std::ptr::null_mut();
        let rc = unsafe {
            nec_graph_upload_adjacency(ctx.raw, n as u32, rows.as_ptr(), degrees.as_ptr(), std::mem::size_of::<Edge>() as u32,
                                       to_offset as u32, std::mem::size_of::<usize>() as u32, &mut raw)
        };
        ctx.check(rc, "nec_graph_upload_adjacency");
        DeviceGraph { ctx, raw, n }
    }

    /// Generic path through `neighbors()` (graph_traits.rs:72; graph.rs:62-64; sw_graph.rs:195-199): an
    /// order-preserving CSR built on the host.
    pub fn from_neighbors<T, A>(ctx: &'c CudaContext, g: &GenericGraph<T, A>) -> Self
    where T: Node, A: AdjContainer<T>
    {
        let (row_ptr, col_idx) = export_csr(g);
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { nec_graph_upload(ctx.raw, g.vertex_count() as u32, col_idx.len() as u64, row_ptr.as_ptr(), col_idx.as_ptr(), &mut raw) };
        ctx.check(rc, "nec_graph_upload");
        DeviceGraph { ctx, raw, n: g.vertex_count() }
    }

    /// `ErEnsembleC::randomize` (src/er_c.rs:128-143) drawn on the device from the ensemble's Pcg64 state
    /// (state low/high, increment low/high); returns the device graph and the generator state afterwards.
    pub fn erc_randomize(ctx: &'c CudaContext, n: usize, prob: f64, rng: [u64; 4]) -> (Self, [u64; 4]) {
        let mut raw = std::ptr::null_mut();
        let mut after = [0u64; 4];
        let rc = unsafe { nec_erc_randomize(ctx.raw, n as u32, prob, rng.as_ptr(), after.as_mut_ptr(), &mut raw) };
        ctx.check(rc, "nec_erc_randomize");
        (DeviceGraph { ctx, raw, n }, after)
    }

    pub fn vertex_count(&self) -> usize { self.n }
    pub fn adjacency_entries(&self) -> u64 {
        let (mut n, mut nnz) = (0u32, 0u64);
        unsafe { nec_graph_dims(self.raw, &mut n, &mut nnz) };
        nnz
    }
    /// (row_ptr, col_idx) of the device graph.
    pub fn download(&self) -> (Vec<u64>, Vec<u32>) {
        let mut rp = vec![0u64; self.n + 1];
        let mut col = vec![0u32; self.adjacency_entries() as usize];
        self.ctx.check(unsafe { nec_graph_download(self.ctx.raw, self.raw, rp.as_mut_ptr(), col.as_mut_ptr()) }, "nec_graph_download");
        (rp, col)
    }
    pub fn vertex_load(&self, include_endpoints: bool) -> Vec<f64> {
        let mut out = vec![0.0f64; self.n];
        self.ctx.check(unsafe { nec_vertex_load(self.ctx.raw, self.raw, include_endpoints as c_int, out.as_mut_ptr()) }, "nec_vertex_load");
        out
    }
    pub fn vertex_load_sources(&self, sources: &[u32], include_endpoints: bool) -> Vec<f64> {
        let mut out = vec![0.0f64; self.n];
        let rc = unsafe { nec_vertex_load_sources(self.ctx.raw, self.raw, sources.as_ptr(), sources.len() as u32, include_endpoints as c_int, out.as_mut_ptr()) };
        self.ctx.check(rc, "nec_vertex_load_sources");
        out
    }
    /// Partial load into a caller-owned DEVICE buffer of n doubles (one-process-per-GPU sharding: hand it to ncclAllReduce).
    pub unsafe fn vertex_load_sources_device(&self, sources: &[u32], include_endpoints: bool, d_out: *mut f64) {
        let rc = nec_vertex_load_sources_device(self.ctx.raw, self.raw, sources.as_ptr(), sources.len() as u32, include_endpoints as c_int, d_out);
        self.ctx.check(rc, "nec_vertex_load_sources_device");
    }
    /// (eccentricity, sum of distances, reached count) per source — diameter / closeness (generic_graph.rs:1116-1144, :1387-1403).
    pub fn distance_measures(&self, sources: &[u32]) -> (Vec<u32>, Vec<u64>, Vec<u32>) {
        let k = sources.len();
        let (mut ecc, mut sum, mut reached) = (vec![0u32; k], vec![0u64; k], vec![0u32; k]);
        let rc = unsafe { nec_distance_measures(self.ctx.raw, self.raw, sources.as_ptr(), k as u32, ecc.as_mut_ptr(), sum.as_mut_ptr(), reached.as_mut_ptr()) };
        self.ctx.check(rc, "nec_distance_measures");
        (ecc, sum, reached)
    }
    pub fn wave_size(&self, n_sources_total: u64) -> u32 {
        let mut w = 0u32;
        self.ctx.check(unsafe { nec_wave_size(self.ctx.raw, self.raw, n_sources_total, &mut w) }, "nec_wave_size");
        w
    }
    /// (dist, npred, sigma, bk) of one source, for exactness checks.
    pub fn bfs_debug(&self, source: u32) -> (Vec<u32>, Vec<u32>, Vec<f64>, Vec<f64>) {
        let n = self.n;
        let (mut d, mut np, mut s, mut b) = (vec![0u32; n], vec![0u32; n], vec![0f64; n], vec![0f64; n]);
        let rc = unsafe { nec_bfs_debug(self.ctx.raw, self.raw, source, d.as_mut_ptr(), np.as_mut_ptr(), s.as_mut_ptr(), b.as_mut_ptr()) };
        self.ctx.check(rc, "nec_bfs_debug");
        (d, np, s, b)
    }
}
impl<'c> Drop for DeviceGraph<'c> { fn drop(&mut self) { unsafe { nec_graph_free(self.ctx.raw, self.raw) } } }

/// Order-preserving CSR of the adjacency lists: the flattening of `container(i).neighbors()`.
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

/// Local CSR batch of many small graphs: the input of `nec_vertex_load_batch` / `nec_chains_create`.
struct CsrBatch { n_per: Vec<u32>, rp_off: Vec<u64>, rp: Vec<u32>, col_off: Vec<u64>, col: Vec<u32> }

fn export_batch<T, A>(graphs: &[&GenericGraph<T, A>]) -> CsrBatch
where T: Node, A: AdjContainer<T>
{
    let mut b = CsrBatch { n_per: vec![], rp_off: vec![], rp: vec![], col_off: vec![], col: vec![] };
    for g in graphs {
        b.n_per.push(g.vertex_count() as u32);
        b.rp_off.push(b.rp.len() as u64);
        b.col_off.push(b.col.len() as u64);
        let base = b.col.len();
        b.rp.push(0);
        for c in g.container_iter() {
            b.col.extend(c.neighbors().map(|&j| j as u32));
            b.rp.push((b.col.len() - base) as u32);
        }
    }
    b
}

/// Many small graphs, one CTA each (a Markov chain measured after every step, benches/common/mod.rs:79-90).
pub fn vertex_load_batch_cuda<T, A>(ctx: &CudaContext, graphs: &[&GenericGraph<T, A>], include_endpoints: bool) -> Vec<Vec<f64>>
where T: Node, A: AdjContainer<T>
{
    let b = export_batch(graphs);
    let total: usize = b.n_per.iter().map(|&n| n as usize).sum();
    let mut flat = vec![0.0f64; total];
    let rc = unsafe {
        nec_vertex_load_batch(ctx.raw, graphs.len() as u32, b.n_per.as_ptr(), b.rp_off.as_ptr(), b.rp.as_ptr(), b.col_off.as_ptr(),
                              b.col.as_ptr(), include_endpoints as c_int, flat.as_mut_ptr())
    };
    ctx.check(rc, "nec_vertex_load_batch");
    let mut out = Vec::with_capacity(graphs.len());
    let mut at = 0usize;
    for &n in &b.n_per { out.push(flat[at..at + n as usize].to_vec()); at += n as usize; }
    out
}

/// One list operation of a Markov step, as the device replays it (see the header): what `ErStepC`, `ErStepM`
/// and `SwChangeState` boil down to on the adjacency lists.
#[derive(Clone, Copy)]
pub enum ListOp { Push { v: u32, x: u32 }, SwapRemove { v: u32, x: u32 }, Retarget { v: u32, x: u32, y: u32 } }
impl ListOp {
    fn encode(self) -> u64 {
        match self {
            ListOp::Push { v, x } => (1u64 << 60) | ((v as u64) << 40) | ((x as u64) << 20),
            ListOp::SwapRemove { v, x } => (2u64 << 60) | ((v as u64) << 40) | ((x as u64) << 20),
            ListOp::Retarget { v, x, y } => (3u64 << 60) | ((v as u64) << 40) | ((x as u64) << 20) | y as u64,
        }
    }
    /// `add_edge(a, b)` / `remove_edge(a, b)` of GenericGraph (generic_graph.rs:487, :503).
    pub fn add_edge(a: u32, b: u32) -> [ListOp; 2] { [ListOp::Push { v: a, x: b }, ListOp::Push { v: b, x: a }] }
    pub fn remove_edge(a: u32, b: u32) -> [ListOp; 2] { [ListOp::SwapRemove { v: a, x: b }, ListOp::SwapRemove { v: b, x: a }] }
}

/// Device-resident chains: upload once, then ship each step's list operations (`nec_chains_*`).
pub struct DeviceChains<'c> { ctx: &'c CudaContext, raw: *mut NecChains, n_per: Vec<u32> }
impl<'c> DeviceChains<'c> {
    pub fn new<T, A>(ctx: &'c CudaContext, graphs: &[&GenericGraph<T, A>], slack: u32) -> Self
    where T: Node, A: AdjContainer<T>
    {
        let b = export_batch(graphs);
        let mut raw = std::ptr::null_mut();
        let rc = unsafe {
            nec_chains_create(ctx.raw, graphs.len() as u32, b.n_per.as_ptr(), b.rp_off.as_ptr(), b.rp.as_ptr(), b.col_off.as_ptr(),
                              b.col.as_ptr(), slack, &mut raw)
        };
        ctx.check(rc, "nec_chains_create");
        DeviceChains { ctx, raw, n_per: b.n_per }
    }
    /// `ops[g]` = the list operations chain g performed since the last call, in order; returns the load vectors.
    pub fn step_and_load(&mut self, ops: &[Vec<ListOp>], include_endpoints: bool) -> Vec<Vec<f64>> {
        let mut off = Vec::with_capacity(ops.len() + 1);
        let mut flat_ops = Vec::new();
        off.push(0u32);
        for o in ops { flat_ops.extend(o.iter().map(|x| x.encode())); off.push(flat_ops.len() as u32); }
        let total: usize = self.n_per.iter().map(|&n| n as usize).sum();
        let rc = unsafe {
            nec_chains_step_and_load(self.ctx.raw, self.raw, if flat_ops.is_empty() { std::ptr::null() } else { off.as_ptr() },
                                     flat_ops.as_ptr(), include_endpoints as c_int, std::ptr::null_mut())
        };
        self.ctx.check(rc, "nec_chains_step_and_load");
        let flat = unsafe { std::slice::from_raw_parts(nec_chains_result(self.raw), total) };
        let mut out = Vec::with_capacity(self.n_per.len());
        let mut at = 0usize;
        for &n in &self.n_per { out.push(flat[at..at + n as usize].to_vec()); at += n as usize; }
        out
    }
    /// The device copy's adjacency as (local row pointers back to back, neighbour ids), for checks.
    pub fn download(&mut self, col_capacity: usize) -> (Vec<u32>, Vec<u32>) {
        let rp_len: usize = self.n_per.iter().map(|&n| n as usize + 1).sum();
        let (mut rp, mut col, mut nnz) = (vec![0u32; rp_len], vec![0u32; col_capacity], 0u64);
        let rc = unsafe { nec_chains_download(self.ctx.raw, self.raw, rp.as_mut_ptr(), col.as_mut_ptr(), col_capacity as u64, &mut nnz) };
        self.ctx.check(rc, "nec_chains_download");
        col.truncate(nnz as usize);
        (rp, col)
    }
}
impl<'c> Drop for DeviceChains<'c> { fn drop(&mut self) { unsafe { nec_chains_free(self.ctx.raw, self.raw) } } }

/// `diameter()` (generic_graph.rs:1116-1136) and `closeness_centrality()` (:1387-1403) on the GPU.
pub fn diameter_and_closeness_cuda<T, A>(ctx: &CudaContext, g: &GenericGraph<T, A>) -> (Option<usize>, Vec<f64>)
where T: Node, A: AdjContainer<T>
{
    let n = g.vertex_count();
    if n == 0 { return (None, Vec::new()); }
    let dg = DeviceGraph::from_neighbors(ctx, g);
    let sources: Vec<u32> = (0..n as u32).collect();
    let (ecc, sum, reached) = dg.distance_measures(&sources);
    let diameter = if reached[0] as usize == n { Some(ecc[1..].iter().copied().max().unwrap_or(0) as usize) } else { None };
    let val = (n - 1) as f64;
    (diameter, sum.iter().map(|&c| val / c as f64).collect())
}

/// Mirrors `MeasurableGraphQuantities<G>` (graph_traits.rs:161): generic over the graph type, so that the
/// blanket impl below constrains `T` and `A` the way the reference's does (graph_traits.rs:272-278).
pub trait VertexLoadCuda<G> {
    /// Drop-in for `vertex_load(include_endpoints)`; returns a fresh `Vec<f64>` of length N.
    fn vertex_load_cuda(&self, ctx: &CudaContext, include_endpoints: bool) -> Vec<f64>;
    /// Only the contributions of the listed sources (sharding / sampling primitive).
    fn vertex_load_cuda_sources(&self, ctx: &CudaContext, sources: &[u32], include_endpoints: bool) -> Vec<f64>;
}

impl<T, A, E> VertexLoadCuda<GenericGraph<T, A>> for E
where T: Node, A: AdjContainer<T>, GenericGraph<T, A>: Clone, E: AsRef<GenericGraph<T, A>>
{
    fn vertex_load_cuda(&self, ctx: &CudaContext, include_endpoints: bool) -> Vec<f64> {
        DeviceGraph::from_neighbors(ctx, self.as_ref()).vertex_load(include_endpoints)
    }
    fn vertex_load_cuda_sources(&self, ctx: &CudaContext, sources: &[u32], include_endpoints: bool) -> Vec<f64> {
        DeviceGraph::from_neighbors(ctx, self.as_ref()).vertex_load_sources(sources, include_endpoints)
    }
}
