//! Device-backed batched tensors for spenso's concrete-data path.
//!
//! `DeviceBatchTensor` holds one `DenseTensor<Complex<f64>, OrderedStructure>` per phase-space point on a B200
//! and implements `spenso::contraction::Contract` (spenso/src/contraction.rs:32-35) by calling
//! `spn_tensor_contract`; `BatchedNetwork` records a product/sum expression and evaluates it for every point
//! with the fused executor (`spn_net_*`), the counterpart of `Network::execute` (spenso/src/network.rs:1217-1241).
//!
//! This crate is NOT compiled in the repository that ships libspenso_b200.so (its build image has no Rust
//! toolchain); it is the binding a spenso maintainer adds.  See INTEGRATION.md.
pub mod executor;
pub mod sys;

use std::ffi::CStr;
use std::sync::Arc;

use spenso::algebra::complex::Complex;
use spenso::contraction::{Contract, ContractionError};
use spenso::structure::abstract_index::AbstractIndex;
use spenso::structure::dimension::Dimension;
use spenso::structure::representation::LibraryRep;
use spenso::structure::slot::{IsAbstractSlot, Slot};
use spenso::structure::{HasStructure, OrderedStructure, TensorStructure};
use spenso::tensors::data::DenseTensor;

use sys::*;

fn last_error() -> String {
    unsafe { CStr::from_ptr(spn_last_error()) }.to_string_lossy().into_owned()
}

/// Status code -> `ContractionError` (spenso/src/contraction.rs:20-30), variant for variant as INTEGRATION.md section 3
/// lists them; codes without a counterpart in the reference (CUDA, no device, invalid argument, JIT) are `Other`.
fn check(rc: i32) -> Result<(), ContractionError> {
    match rc {
        SPN_OK => Ok(()),
        SPN_ERR_EMPTY_SPARSE => Err(ContractionError::EmptySparse),
        SPN_ERR_STRUCTURE => Err(ContractionError::StructureError(spenso::structure::StructureError::Other(anyhow::anyhow!(
            "spenso_b200: {}",
            last_error()
        )))),
        SPN_ERR_DIMENSION => Err(ContractionError::DimensionError(anyhow::anyhow!("spenso_b200: {}", last_error()).into())),
        _ => Err(ContractionError::Other(anyhow::anyhow!("spenso_b200 ({rc}): {}", last_error()))),
    }
}

/// `Slot<LibraryRep, AbstractIndex>` -> `spn_slot`.
pub fn to_spn_slot(s: &Slot<LibraryRep, AbstractIndex>) -> spn_slot {
    let (rep_kind, rep_id) = match s.rep_name() {
        LibraryRep::Dummy => (SPN_REP_DUMMY, 0),
        LibraryRep::SelfDual(n) => (SPN_REP_SELF_DUAL, n as i16),
        LibraryRep::InlineMetric(n) => (SPN_REP_INLINE_METRIC, n as i16),
        LibraryRep::Dualizable(k) => (SPN_REP_DUALIZABLE, k),
    };
    let (aind_kind, aind) = match s.aind() {
        AbstractIndex::Normal(i) => (0u8, i as u64),
        AbstractIndex::Dualize(i) => (1, i as u64),
        AbstractIndex::Double(a, b) => (2, ((a as u64) << 16) | b as u64),
        AbstractIndex::Dummy(i) => (3, i as u64),
        AbstractIndex::Added(i) => (4, i as u64),
        #[allow(unreachable_patterns)]
        _ => panic!("symbolic abstract indices must be interned to integers before crossing the FFI"),
    };
    let dim = match s.dim() {
        Dimension::Concrete(d) => d as u32,
        #[allow(unreachable_patterns)]
        _ => panic!("symbolic dimensions have no numeric tensor"),
    };
    spn_slot { rep_kind, aind_kind, rep_id, dim, aind }
}

fn slots_of(structure: &OrderedStructure) -> Vec<spn_slot> {
    structure.external_structure_iter().map(|s| to_spn_slot(&s)).collect()
}

/// One CUDA device + stream.
pub struct Device(*mut spn_ctx);
unsafe impl Send for Device {}
unsafe impl Sync for Device {}
impl Device {
    pub fn new(index: i32) -> Result<Arc<Self>, ContractionError> {
        let mut h = std::ptr::null_mut();
        check(unsafe { spn_ctx_create(index, &mut h) })?;
        Ok(Arc::new(Device(h)))
    }
}
impl Drop for Device {
    fn drop(&mut self) {
        unsafe { spn_ctx_destroy(self.0) }
    }
}

/// Owner of one device tensor handle.
struct Handle(*mut spn_tensor);
unsafe impl Send for Handle {}
unsafe impl Sync for Handle {}
impl Drop for Handle {
    fn drop(&mut self) {
        unsafe { spn_tensor_free(self.0) }
    }
}

/// `n_points` dense complex tensors of one structure, resident on the device (or one shared, point-independent
/// tensor).  Device tensors are immutable once built, so `Clone` shares the handle (the reference's executor clones
/// tensors only to consume them in `neg()` / `+=`, network.rs:1320-1325, 1381).
#[derive(Clone)]
pub struct DeviceBatchTensor {
    h: Arc<Handle>,
    dev: Arc<Device>,
    pub structure: OrderedStructure,
    pub n_points: u64,
}

impl DeviceBatchTensor {
    /// Upload one `DenseTensor` per point (all of the same, canonically ordered, structure).
    pub fn from_points(dev: &Arc<Device>, points: &[DenseTensor<Complex<f64>, OrderedStructure>]) -> Result<Self, ContractionError> {
        let structure = points[0].structure.clone();
        let slots = slots_of(&structure);
        let size = structure.size().map_err(|e| ContractionError::Other(e.into()))?;
        let mut h = std::ptr::null_mut();
        check(unsafe {
            spn_tensor_batched(dev.0, slots.as_ptr(), slots.len() as u32, points.len() as u64, SPN_C64, SPN_POINT_MAJOR, &mut h)
        })?;
        // Complex<f64> is two f64 (spenso/src/algebra/complex.rs:55-58): the points are packed end to end — the wire
        // format — and cross the bus in ONE copy (a copy per point would pay the launch latency n_points times)
        let mut packed: Vec<Complex<f64>> = Vec::with_capacity(size * points.len());
        for t in points {
            debug_assert_eq!(t.data.len(), size);
            packed.extend_from_slice(&t.data);
        }
        let this = Self { h: Arc::new(Handle(h)), dev: dev.clone(), structure, n_points: points.len() as u64 };
        check(unsafe { spn_tensor_upload(h, packed.as_ptr().cast(), 0, points.len() as u64) })?;
        Ok(this)
    }

    /// A point-independent rank-0 tensor (`Sc::ref_one()` and scalar leaves).
    pub fn shared_scalar(dev: &Arc<Device>, re: f64, im: f64) -> Result<Self, ContractionError> {
        let v = [re, im];
        let mut h = std::ptr::null_mut();
        check(unsafe { spn_tensor_shared_dense(dev.0, std::ptr::null(), 0, v.as_ptr(), &mut h) })?;
        Ok(Self { h: Arc::new(Handle(h)), dev: dev.clone(), structure: OrderedStructure::default(), n_points: 1 })
    }

    pub fn is_scalar(&self) -> bool {
        unsafe { spn_tensor_order(self.h.0) == 0 }
    }

    fn unary(&self, f: unsafe extern "C" fn(*mut spn_ctx, *const spn_tensor, *mut *mut spn_tensor) -> i32) -> Result<Self, ContractionError> {
        let mut out = std::ptr::null_mut();
        check(unsafe { f(self.dev.0, self.h.0, &mut out) })?;
        Ok(self.wrap(out, self.structure.clone()))
    }
    /// `Neg`
    pub fn neg(&self) -> Result<Self, ContractionError> { self.unary(spn_tensor_neg) }
    /// FUN_LIB conj (spenso-hep-lib/src/lib.rs:511-520)
    pub fn conj(&self) -> Result<Self, ContractionError> { self.unary(spn_tensor_conj) }
    /// `s.ref_one() / s`
    pub fn recip(&self) -> Result<Self, ContractionError> { self.unary(spn_tensor_recip) }
    /// `self += rhs` as a new tensor (`AddAssign<T::Ref>`)
    pub fn add(&self, rhs: &Self) -> Result<Self, ContractionError> {
        let mut out = std::ptr::null_mut();
        check(unsafe { spn_tensor_add(self.dev.0, self.h.0, rhs.h.0, 0, &mut out) })?;
        Ok(self.wrap(out, self.structure.clone()))
    }

    /// Download point `p` as a reference `DenseTensor`.
    pub fn point(&self, p: u64) -> Result<DenseTensor<Complex<f64>, OrderedStructure>, ContractionError> {
        let size = unsafe { spn_tensor_size(self.h.0) } as usize;
        let mut data = vec![Complex { re: 0.0, im: 0.0 }; size];
        check(unsafe { spn_tensor_download(self.h.0, data.as_mut_ptr().cast(), p, 1) })?;
        Ok(DenseTensor { data, structure: self.structure.clone() })
    }

    fn wrap(&self, h: *mut spn_tensor, structure: OrderedStructure) -> Self {
        Self { h: Arc::new(Handle(h)), dev: self.dev.clone(), structure, n_points: unsafe { spn_tensor_n_points(h) } }
    }
}

/// The unit operation of the path: `a.contract(&b)` for every point at once.
impl Contract<DeviceBatchTensor> for DeviceBatchTensor {
    type LCM = DeviceBatchTensor;
    fn contract(&self, other: &DeviceBatchTensor) -> Result<Self::LCM, ContractionError> {
        // the result structure is the host-side merge, exactly as in the blanket impl (contraction.rs:148-153);
        // the engine computes the same merge for its stride plan and re-checks it
        let (structure, _, _, _) = self.structure.merge(&other.structure).map_err(ContractionError::StructureError)?;
        let mut out = std::ptr::null_mut();
        check(unsafe { spn_tensor_contract(self.dev.0, self.h.0, other.h.0, &mut out) })?;
        Ok(self.wrap(out, structure))
    }
}

impl std::ops::Neg for &DeviceBatchTensor {
    type Output = DeviceBatchTensor;
    fn neg(self) -> DeviceBatchTensor {
        DeviceBatchTensor::neg(self).expect("neg")
    }
}

impl std::ops::Add for &DeviceBatchTensor {
    type Output = DeviceBatchTensor;
    fn add(self, rhs: &DeviceBatchTensor) -> DeviceBatchTensor {
        DeviceBatchTensor::add(self, rhs).expect("add")
    }
}

/// Host data tensors -> shared device tensors (library tensors with concrete indices, local point-independent tensors).
impl executor::ToDevice for spenso::tensors::data::DataTensor<Complex<f64>, OrderedStructure> {
    fn to_device(&self, dev: &Arc<Device>) -> Result<DeviceBatchTensor, ContractionError> {
        use spenso::tensors::data::DataTensor;
        let structure = self.structure().clone();
        let slots = slots_of(&structure);
        let mut h = std::ptr::null_mut();
        match self {
            DataTensor::Dense(d) => check(unsafe {
                spn_tensor_shared_dense(dev.0, slots.as_ptr(), slots.len() as u32, d.data.as_ptr().cast(), &mut h)
            })?,
            DataTensor::Sparse(s) => {
                // HashMap order is irrelevant: entries are keyed by flat index (spenso/src/tensors/data/sparse.rs:48-53)
                let flat: Vec<u64> = s.elements.keys().map(|k| usize::from(*k) as u64).collect();
                let vals: Vec<Complex<f64>> = s.elements.values().cloned().collect();
                check(unsafe {
                    spn_tensor_shared_sparse(dev.0, slots.as_ptr(), slots.len() as u32, flat.as_ptr(), vals.as_ptr().cast(), flat.len() as u64, &mut h)
                })?
            }
        }
        Ok(DeviceBatchTensor { h: Arc::new(Handle(h)), dev: dev.clone(), structure, n_points: 1 })
    }
}

impl DeviceBatchTensor {
    /// `ScalarMul<Complex<f64>>` (spenso/src/algebra/scalar_mul.rs:11-57)
    pub fn scalar_mul(&self, s: Complex<f64>) -> DeviceBatchTensor {
        let mut out = std::ptr::null_mut();
        check(unsafe { spn_tensor_scalar_mul(self.dev.0, self.h.0, s.re, s.im, &mut out) }).expect("scalar_mul");
        self.wrap(out, self.structure.clone())
    }
}

/// A product / sum expression over per-point leaves, shared tensors and library tensors, compiled once and
/// evaluated for all points by one fused kernel — what `Network::execute::<Sequential, SmallestDegree, ..>` does per
/// point on the CPU.  Host data in, host data out (`spn_net_execute_host`).
pub struct BatchedNetwork {
    h: *mut spn_net,
    _dev: Arc<Device>,
    n_inputs: u32,
}
impl Drop for BatchedNetwork {
    fn drop(&mut self) {
        unsafe { spn_net_destroy(self.h) }
    }
}
#[derive(Clone, Copy)]
pub struct Node(i32);

impl BatchedNetwork {
    pub fn new(dev: &Arc<Device>) -> Result<Self, ContractionError> {
        let mut h = std::ptr::null_mut();
        check(unsafe { spn_net_create(dev.0, &mut h) })?;
        Ok(Self { h, _dev: dev.clone(), n_inputs: 0 })
    }
    fn node(rc: i32) -> Result<Node, ContractionError> {
        if rc < 0 { check(-rc).map(|_| Node(0)) } else { Ok(Node(rc)) }
    }
    /// `NetworkLeaf::LocalTensor` that differs per point: the k-th call declares the k-th input buffer.
    pub fn per_point(&mut self, slots: &[Slot<LibraryRep, AbstractIndex>]) -> Result<Node, ContractionError> {
        let s: Vec<_> = slots.iter().map(to_spn_slot).collect();
        self.n_inputs += 1;
        Self::node(unsafe { spn_net_leaf_batched(self.h, s.as_ptr(), s.len() as u32, SPN_C64) })
    }
    /// `NetworkLeaf::LibraryKey` with concrete indices (`which` = SPN_LIB_*; HEP_LIB of spenso-hep-lib).
    pub fn library(&mut self, which: i32, slots: &[Slot<LibraryRep, AbstractIndex>]) -> Result<Node, ContractionError> {
        let s: Vec<_> = slots.iter().map(to_spn_slot).collect();
        Self::node(unsafe { spn_net_leaf_library(self.h, which, s.as_ptr(), s.len() as u32) })
    }
    pub fn scalar(&mut self, c: Complex<f64>) -> Result<Node, ContractionError> {
        Self::node(unsafe { spn_net_leaf_scalar(self.h, c.re, c.im) })
    }
    fn op(&mut self, op: i32, children: &[Node], power: i32) -> Result<Node, ContractionError> {
        let ids: Vec<i32> = children.iter().map(|n| n.0).collect();
        Self::node(unsafe { spn_net_op(self.h, op, ids.as_ptr(), ids.len() as u32, power) })
    }
    pub fn product(&mut self, c: &[Node]) -> Result<Node, ContractionError> { self.op(SPN_OP_PRODUCT, c, 1) }
    pub fn sum(&mut self, c: &[Node]) -> Result<Node, ContractionError> { self.op(SPN_OP_SUM, c, 1) }
    pub fn neg(&mut self, c: Node) -> Result<Node, ContractionError> { self.op(SPN_OP_NEG, &[c], 1) }
    pub fn pow(&mut self, c: Node, n: i32) -> Result<Node, ContractionError> { self.op(SPN_OP_POWER, &[c], n) }

    pub fn compile(&mut self, root: Node) -> Result<(), ContractionError> {
        check(unsafe { spn_net_set_root(self.h, root.0) })?;
        check(unsafe { spn_net_compile(self.h, SPN_EXEC_AUTO) })
    }

    /// Evaluate at `n_points` points.  `inputs[k]` holds the k-th per-point leaf for all points, point after point
    /// (canonical order).  Returns the per-point results (point after point) and their sum.
    pub fn execute(&self, inputs: &[&[Complex<f64>]], n_points: u64) -> Result<(Vec<Complex<f64>>, Vec<Complex<f64>>), ContractionError> {
        assert_eq!(inputs.len() as u32, self.n_inputs);
        let size = unsafe { spn_net_result_size(self.h) } as usize;
        let mut out = vec![Complex { re: 0.0, im: 0.0 }; size * n_points as usize];
        let mut sum = vec![Complex { re: 0.0, im: 0.0 }; size];
        let ptrs: Vec<*const std::ffi::c_void> = inputs.iter().map(|s| s.as_ptr().cast()).collect();
        check(unsafe {
            spn_net_execute_host(self.h, ptrs.as_ptr(), self.n_inputs, n_points, out.as_mut_ptr().cast(), sum.as_mut_ptr().cast(), 0)
        })?;
        Ok((out, sum))
    }
}
