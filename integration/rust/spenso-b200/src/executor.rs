//! `ExecuteOp` and a `ContractionStrategy` for a device-resident store: the plug points of
//! `Network::execute::<Strat, C, ..>` (spenso/src/network.rs:1197-1241).
//!
//! The reference's blanket `impl ExecuteOp for NetworkStore<T, Sc>` (network.rs:1244-1706) and its
//! `SmallestDegree` / `ContractScalars` / `SingleSmallestDegree` strategies (network/contract.rs:43-498) are written for
//! `NetworkStore<T, Sc>`; the orphan rule does not let this crate implement them for `NetworkStore<DeviceBatchTensor, _>`,
//! so the store is a local type, [`DeviceStore`], and the two impls below reproduce the reference's decisions — which
//! leaf is folded, contracted or summed when, in which operand order, and how the graph is rewired (`iter_nodes`,
//! `involved_node_id`, `get_lib_data`, `identify_nodes_without_self_edges*`) — while every tensor operation is one call
//! into libspenso_b200.so, for ALL phase-space points at once:
//!
//! | reference (per point, host)                   | here (all points, device)                       |
//! |-----------------------------------------------|-------------------------------------------------|
//! | `tensors[l1].contract(&tensors[l2])`          | `spn_tensor_contract`                           |
//! | `accumulator += tensors[t].refer()`           | `spn_tensor_add`                                |
//! | `tensors[t].clone().neg()`                    | `spn_tensor_neg`                                |
//! | `tensors[l].scalar_mul(&acc)`                 | `spn_tensor_contract` with the rank-0 scalar    |
//! | `s *= scalar[si].refer()`, `s.ref_one() / s`  | `spn_tensor_contract`, `spn_tensor_recip`       |
//! | `T::from(graph.get_lib_data(lib, id))`        | [`ToDevice::to_device`] (`spn_tensor_shared_*`) |
//! | `fn_lib.apply(&f, t)`                         | `FunctionLibrary<DeviceBatchTensor, ..>`        |
//!
//! The same call sequence, with the same match arms, is exercised in the repository that ships the library by
//! `tests/cpp/blanket_executor_test.cpp` (C++ over the same C ABI; no Rust toolchain exists in that image), which
//! also checks that handing this executor's contraction order to `spn_net_set_schedule` makes the engine's own
//! executor reproduce it bit for bit.
use std::fmt::{Debug, Display};
use std::sync::Arc;

use anyhow::anyhow;
use spenso::contraction::Contract;
use spenso::network::graph::{NetworkGraph, NetworkLeaf, NetworkNode, NetworkOp};
use spenso::network::library::{FunctionLibrary, Library, LibraryTensor};
use spenso::network::{ContractionStrategy, ExecuteOp, TensorNetworkError};
use spenso::structure::abstract_index::AbsInd;
use spenso::structure::permuted::PermuteTensor;
use spenso::structure::slot::IsAbstractSlot;
use spenso::structure::{HasStructure, PermutedStructure, TensorStructure};

use crate::{Device, DeviceBatchTensor};

/// Host tensor -> shared (point-independent) device tensor: what `T::from(lib_tensor_with_indices)` is in the
/// blanket impl.  Implemented in lib.rs for `DataTensor<Complex<f64>, OrderedStructure>` and its dense/sparse arms.
pub trait ToDevice {
    fn to_device(&self, dev: &Arc<Device>) -> Result<DeviceBatchTensor, spenso::contraction::ContractionError>;
}

/// `NetworkStore { tensors, scalar }` (network/store.rs:125-129) with device tensors.  Scalars are rank-0 device
/// tensors (one value per point, or shared), so `Sc`'s algebra is the tensor algebra.  Append only, like the
/// reference's store.
pub struct DeviceStore {
    pub dev: Arc<Device>,
    pub tensors: Vec<DeviceBatchTensor>,
    pub scalar: Vec<DeviceBatchTensor>,
    /// (n1, n2) store positions' graph nodes of every pairwise contraction, in execution order: what
    /// `spn_net_set_schedule` takes when the same network is later run through the fused executor
    pub contraction_log: Vec<(usize, usize)>,
}

impl DeviceStore {
    pub fn new(dev: &Arc<Device>) -> Self {
        Self { dev: dev.clone(), tensors: Vec::new(), scalar: Vec::new(), contraction_log: Vec::new() }
    }
    pub fn add_tensor(&mut self, t: DeviceBatchTensor) -> usize {
        self.tensors.push(t);
        self.tensors.len() - 1
    }
    pub fn add_scalar(&mut self, s: DeviceBatchTensor) -> usize {
        self.scalar.push(s);
        self.scalar.len() - 1
    }
    fn one(&self) -> Result<DeviceBatchTensor, spenso::contraction::ContractionError> {
        DeviceBatchTensor::shared_scalar(&self.dev, 1.0, 0.0)
    }
}

/// `SmallestDegree` (network/contract.rs:239-263) for [`DeviceStore`]: `ContractScalars`, then
/// `SingleSmallestDegree` until nothing is left to contract, then `ContractScalars` again.
pub struct DeviceSmallestDegree;

impl<LT, L, K, FK, Aind> ContractionStrategy<DeviceStore, L, K, FK, Aind> for DeviceSmallestDegree
where
    LT: LibraryTensor + Clone,
    LT::WithIndices: PermuteTensor<Permuted = LT::WithIndices> + ToDevice,
    <<LT::WithIndices as HasStructure>::Structure as TensorStructure>::Slot: IsAbstractSlot<Aind = Aind>,
    L: Library<<LT as HasStructure>::Structure, Key = K, Value = PermutedStructure<LT>>,
    K: Display + Debug + Clone,
    FK: Display + Debug + Clone,
    Aind: AbsInd,
{
    fn contract(
        executor: &mut DeviceStore,
        graph: NetworkGraph<K, FK, Aind>,
        lib: &L,
    ) -> Result<(NetworkGraph<K, FK, Aind>, bool), TensorNetworkError<K, FK>> {
        let (mut graph, mut didsmth) = contract_scalars::<LT, L, K, FK, Aind>(executor, graph, lib)?;
        loop {
            let (g, smth) = single_smallest_degree::<LT, L, K, FK, Aind>(executor, graph, lib)?;
            graph = g;
            if !smth {
                break;
            }
            didsmth = true;
        }
        let (graph, _) = contract_scalars::<LT, L, K, FK, Aind>(executor, graph, lib)?;
        Ok((graph, didsmth))
    }
}

/// A leaf of the graph as a device tensor: a store entry, or a library tensor instantiated with the node's indices
/// (`graph.get_lib_data`, network/graph.rs:291-322) and uploaded as a shared tensor.  The flag says "library key".
fn leaf_on_device<LT, L, K, FK, Aind>(
    store: &DeviceStore,
    graph: &NetworkGraph<K, FK, Aind>,
    lib: &L,
    nid: spenso::network::graph::NodeIndex,
) -> Result<Option<(DeviceBatchTensor, bool)>, TensorNetworkError<K, FK>>
where
    LT: LibraryTensor + Clone,
    LT::WithIndices: PermuteTensor<Permuted = LT::WithIndices> + ToDevice,
    <<LT::WithIndices as HasStructure>::Structure as TensorStructure>::Slot: IsAbstractSlot<Aind = Aind>,
    L: Library<<LT as HasStructure>::Structure, Key = K, Value = PermutedStructure<LT>>,
    K: Display + Debug + Clone,
    FK: Display + Debug + Clone,
    Aind: AbsInd,
{
    Ok(match &graph.graph[nid] {
        NetworkNode::Leaf(NetworkLeaf::LocalTensor(i)) => Some((store.tensors[*i].clone(), false)),
        NetworkNode::Leaf(NetworkLeaf::LibraryKey(_)) => Some((graph.get_lib_data(lib, nid).unwrap().to_device(&store.dev)?, true)),
        _ => None,
    })
}

/// What ContractScalars does (network/contract.rs:73-208), for a device store: fold every Scalar leaf of the product into
/// one; when at most one tensor factor is left, multiply it in and dissolve the op node.
fn contract_scalars<LT, L, K, FK, Aind>(
    executor: &mut DeviceStore,
    mut graph: NetworkGraph<K, FK, Aind>,
    lib: &L,
) -> Result<(NetworkGraph<K, FK, Aind>, bool), TensorNetworkError<K, FK>>
where
    LT: LibraryTensor + Clone,
    LT::WithIndices: PermuteTensor<Permuted = LT::WithIndices> + ToDevice,
    <<LT::WithIndices as HasStructure>::Structure as TensorStructure>::Slot: IsAbstractSlot<Aind = Aind>,
    L: Library<<LT as HasStructure>::Structure, Key = K, Value = PermutedStructure<LT>>,
    K: Display + Debug + Clone,
    FK: Display + Debug + Clone,
    Aind: AbsInd,
{
    graph.sync_order();
    // census of the product: scalar leaves (store position, node), tensor factors, the op node itself
    let mut scalar_leaves = Vec::new();
    let mut tensor_factors = Vec::new();
    let mut op_node = None;
    for (nid, _, node) in graph.graph.iter_nodes() {
        match node {
            NetworkNode::Leaf(NetworkLeaf::Scalar(pos)) => scalar_leaves.push((*pos, nid)),
            NetworkNode::Leaf(_) => tensor_factors.push(nid),
            NetworkNode::Op(NetworkOp::Product) => op_node = Some(nid),
            NetworkNode::Op(_) => {}
        }
    }
    let dissolves = tensor_factors.len() <= 1; // nothing left to contract afterwards: the op node goes away

    if scalar_leaves.is_empty() {
        // no scalar at all: a product of one tensor is that tensor
        if let (true, Some(&only), Some(op)) = (dissolves, tensor_factors.first(), op_node) {
            let leaf = graph.graph[only].clone();
            graph.identify_nodes_without_self_edges(&[op, only], leaf);
            return Ok((graph, true));
        }
        return Ok((graph, false));
    }
    if !dissolves && scalar_leaves.len() == 1 {
        return Ok((graph, false)); // a single scalar beside several tensors stays where it is
    }

    // the reference starts from the LAST scalar and multiplies the others in, in node order
    let (last_pos, _) = *scalar_leaves.last().unwrap();
    let mut product = executor.scalar[last_pos].clone();
    for (pos, _) in &scalar_leaves[..scalar_leaves.len() - 1] {
        product = product.contract(&executor.scalar[*pos])?;
    }
    let mut merged: Vec<_> = scalar_leaves.iter().map(|(_, nid)| *nid).collect();
    let replacement = if dissolves {
        merged.extend(op_node);
        match tensor_factors.first() {
            Some(&only) => {
                merged.push(only);
                let (t, _) = leaf_on_device::<LT, L, K, FK, Aind>(executor, &graph, lib, only)?.expect("a tensor factor");
                NetworkLeaf::LocalTensor(executor.add_tensor(t.contract(&product)?)) // scalar_mul
            }
            None => NetworkLeaf::Scalar(executor.add_scalar(product)),
        }
    } else {
        NetworkLeaf::Scalar(executor.add_scalar(product))
    };
    if dissolves {
        graph.identify_nodes_without_self_edges(&merged, NetworkNode::Leaf(replacement));
    } else {
        graph.identify_nodes_without_self_edges_merge_heads(&merged, NetworkNode::Leaf(replacement));
    }
    Ok((graph, true))
}

/// One step of SingleSmallestDegree (network/contract.rs:346-497): among the tensor leaves take the one with the fewest
/// internal slot edges (first minimum; a leaf without any pairs up with the tensor visited before it, at maximal
/// weight), contract it with the neighbour its last internal edge leads to, and put the result in its place.
fn single_smallest_degree<LT, L, K, FK, Aind>(
    executor: &mut DeviceStore,
    mut graph: NetworkGraph<K, FK, Aind>,
    lib: &L,
) -> Result<(NetworkGraph<K, FK, Aind>, bool), TensorNetworkError<K, FK>>
where
    LT: LibraryTensor + Clone,
    LT::WithIndices: PermuteTensor<Permuted = LT::WithIndices> + ToDevice,
    <<LT::WithIndices as HasStructure>::Structure as TensorStructure>::Slot: IsAbstractSlot<Aind = Aind>,
    L: Library<<LT as HasStructure>::Structure, Key = K, Value = PermutedStructure<LT>>,
    K: Display + Debug + Clone,
    FK: Display + Debug + Clone,
    Aind: AbsInd,
{
    graph.sync_order();
    let mut choice: Option<(i32, _, _)> = None; // (weight, n1, n2)
    let mut visited_before = None;
    for (nid, hedges, node) in graph.graph.iter_nodes() {
        if !node.is_tensor() {
            continue;
        }
        let mut internal_edges = 0;
        let mut last_internal = None;
        for h in hedges {
            if graph.graph[[&h]].is_slot() && graph.graph.inv(h) != h {
                internal_edges += 1;
                last_internal = Some(h);
            }
        }
        let (weight, partner) = match last_internal {
            Some(h) => (internal_edges, graph.graph.involved_node_id(h)),
            None => (i32::MAX, visited_before), // unconnected: outer product with the previous tensor, as a last resort
        };
        // (the first unconnected tensor has no partner yet: it only registers itself as "visited")
        let pair = partner.map(|p| (weight, nid, p));
        visited_before = Some(nid);
        if let Some(candidate) = pair {
            if choice.map_or(true, |(w, _, _)| candidate.0 < w) {
                choice = Some(candidate);
            }
        }
    }
    let Some((_, n1, n2)) = choice else { return Ok((graph, false)) };

    if matches!((&graph.graph[n1], &graph.graph[n2]), (NetworkNode::Op(NetworkOp::Product), _) | (_, NetworkNode::Op(NetworkOp::Product))) {
        return Err(TensorNetworkError::SlotEdgeToProdNode);
    }
    if matches!(&graph.graph[n1], NetworkNode::Leaf(NetworkLeaf::Scalar(_))) || matches!(&graph.graph[n2], NetworkNode::Leaf(NetworkLeaf::Scalar(_))) {
        return Err(TensorNetworkError::SlotEdgeToScalarNode);
    }
    let a = leaf_on_device::<LT, L, K, FK, Aind>(executor, &graph, lib, n1)?;
    let b = leaf_on_device::<LT, L, K, FK, Aind>(executor, &graph, lib, n2)?;
    let (Some((a, a_is_lib)), Some((b, b_is_lib))) = (a, b) else {
        return Err(TensorNetworkError::CannotContractEdgeBetween(graph.graph[n1].clone(), graph.graph[n2].clone()));
    };
    // the receiver is the local tensor whenever exactly one side is a library key (contract.rs:428-461): the operand
    // order decides the rounding, so it is kept
    let contracted = if a_is_lib && !b_is_lib { b.contract(&a)? } else { a.contract(&b)? };
    executor.contraction_log.push((usize::from(n1), usize::from(n2)));
    let leaf = NetworkLeaf::LocalTensor(executor.add_tensor(contracted));
    graph.identify_nodes_without_self_edges_merge_heads(&[n1, n2], NetworkNode::Leaf(leaf));
    Ok((graph, true))
}

/// `impl ExecuteOp for NetworkStore<T, Sc>` (network.rs:1244-1706) with T = Sc = [`DeviceBatchTensor`].
impl<LT, L, K, FK, FL, Aind> ExecuteOp<FL, L, K, FK, Aind> for DeviceStore
where
    LT: LibraryTensor + Clone,
    LT::WithIndices: PermuteTensor<Permuted = LT::WithIndices> + ToDevice,
    <<LT::WithIndices as HasStructure>::Structure as TensorStructure>::Slot: IsAbstractSlot<Aind = Aind>,
    L: Library<<LT as HasStructure>::Structure, Key = K, Value = PermutedStructure<LT>>,
    K: Display + Debug + Clone,
    FK: Display + Debug,
    FL: FunctionLibrary<DeviceBatchTensor, DeviceBatchTensor, Key = FK>,
    Aind: AbsInd,
{
    fn execute<C: ContractionStrategy<Self, L, K, FK, Aind>>(
        &mut self,
        mut graph: NetworkGraph<K, FK, Aind>,
        lib: &L,
        fn_lib: &FL,
        op: NetworkOp<FK>,
    ) -> Result<NetworkGraph<K, FK, Aind>, TensorNetworkError<K, FK>> {
        graph.sync_order();
        // the single leaf child of a unary op (Neg / Function / Power): network.rs:1284-1302, 1475-1493, 1540-1559
        let unary_child = |graph: &NetworkGraph<K, FK, Aind>, is_op: &dyn Fn(&NetworkNode<K, FK>) -> bool| {
            let (opid, children, _) = graph.graph.iter_nodes().find(|(_, _, d)| is_op(d)).unwrap();
            let mut child = None;
            for c in children {
                if let Some(id) = graph.graph.involved_node_id(c)
                    && let NetworkNode::Leaf(_) = &graph.graph[id]
                    && child.is_none()
                {
                    child = Some(id);
                }
            }
            (opid, child)
        };
        match op {
            NetworkOp::Neg => {
                let (opid, child) = unary_child(&graph, &|d| matches!(d, NetworkNode::Op(NetworkOp::Neg)));
                let child_id = child.ok_or(TensorNetworkError::ChildlessNeg)?;
                let NetworkNode::Leaf(leaf) = &graph.graph[child_id] else { unreachable!() };
                let new_node = match leaf {
                    NetworkLeaf::Scalar(s) => NetworkLeaf::Scalar(self.add_scalar(self.scalar[*s].neg()?)),
                    NetworkLeaf::LibraryKey(_) => {
                        let t = graph.get_lib_data(lib, child_id).unwrap().to_device(&self.dev)?.neg()?;
                        NetworkLeaf::LocalTensor(self.add_tensor(t))
                    }
                    NetworkLeaf::LocalTensor(t) => NetworkLeaf::LocalTensor(self.add_tensor(self.tensors[*t].neg()?)),
                };
                graph.identify_nodes_without_self_edges(&[child_id, opid], NetworkNode::Leaf(new_node));
                Ok(graph)
            }
            NetworkOp::Product => {
                let (graph, _) = C::contract(self, graph, lib)?;
                Ok(graph)
            }
            NetworkOp::Sum => {
                // network.rs:1342-1467: the accumulator starts as the first leaf, the others are added in node order
                let mut targets = Vec::new();
                let mut all_nodes = Vec::new();
                for (n, _, v) in graph.graph.iter_nodes() {
                    all_nodes.push(n);
                    if let NetworkNode::Leaf(l) = &v {
                        targets.push((n, l.clone()));
                    }
                }
                let value_of = |store: &DeviceStore, graph: &NetworkGraph<K, FK, Aind>, nid, l: &NetworkLeaf<K>| match l {
                    NetworkLeaf::Scalar(s) => Ok::<_, TensorNetworkError<K, FK>>(store.scalar[*s].clone()),
                    NetworkLeaf::LocalTensor(t) => Ok(store.tensors[*t].clone()),
                    NetworkLeaf::LibraryKey(_) => Ok(graph.get_lib_data(lib, nid).unwrap().to_device(&store.dev)?),
                };
                let (nf, first) = &targets[0];
                let mut accumulator = value_of(self, &graph, *nf, first)?;
                let scalar_mode = matches!(first, NetworkLeaf::Scalar(_)) || accumulator.is_scalar();
                for (nid, t) in &targets[1..] {
                    let rhs = value_of(self, &graph, *nid, t)?;
                    match (scalar_mode, t, rhs.is_scalar()) {
                        (true, NetworkLeaf::LibraryKey(_), _) => return Err(TensorNetworkError::ScalarLibSum("".to_string())),
                        (true, _, false) => return Err(TensorNetworkError::NotAllScalars("".to_string())),
                        (false, NetworkLeaf::Scalar(_), _) => return Err(TensorNetworkError::SumScalarTensor("".to_string())),
                        _ => {}
                    }
                    accumulator = accumulator.add(&rhs)?; // accumulator += rhs
                }
                let new_node = if scalar_mode {
                    NetworkLeaf::Scalar(self.add_scalar(accumulator))
                } else {
                    NetworkLeaf::LocalTensor(self.add_tensor(accumulator))
                };
                graph.identify_nodes_without_self_edges(&all_nodes, NetworkNode::Leaf(new_node));
                Ok(graph)
            }
            NetworkOp::Function(f) => {
                let (opid, child) = unary_child(&graph, &|d| matches!(d, NetworkNode::Op(NetworkOp::Function(_))));
                let child_id = child.ok_or(TensorNetworkError::ChildlessNeg)?;
                let NetworkNode::Leaf(leaf) = &graph.graph[child_id] else { unreachable!() };
                let new_node = match leaf {
                    NetworkLeaf::Scalar(s) => {
                        let s = fn_lib.apply_scalar(&f, self.scalar[*s].clone())?;
                        NetworkLeaf::Scalar(self.add_scalar(s))
                    }
                    NetworkLeaf::LibraryKey(_) => {
                        let t = fn_lib.apply(&f, graph.get_lib_data(lib, child_id).unwrap().to_device(&self.dev)?)?;
                        NetworkLeaf::LocalTensor(self.add_tensor(t))
                    }
                    NetworkLeaf::LocalTensor(t) => {
                        let t = fn_lib.apply(&f, self.tensors[*t].clone())?;
                        NetworkLeaf::LocalTensor(self.add_tensor(t))
                    }
                };
                graph.identify_nodes_without_self_edges(&[child_id, opid], NetworkNode::Leaf(new_node));
                Ok(graph)
            }
            NetworkOp::Power(pow) => {
                // network.rs:1526-1703, including its literal behaviour for odd powers
                let (opid, child) = unary_child(&graph, &|d| matches!(d, NetworkNode::Op(NetworkOp::Power(_))));
                let child_id = child.ok_or(TensorNetworkError::ChildlessNeg)?;
                let NetworkNode::Leaf(leaf) = &graph.graph[child_id] else { unreachable!() };
                let n = pow.abs();
                let new_node = match leaf {
                    NetworkLeaf::Scalar(si) => {
                        if n == 0 {
                            NetworkLeaf::Scalar(*si)
                        } else {
                            let mut s = self.scalar[*si].clone();
                            for _ in 1..n {
                                s = s.contract(&self.scalar[*si])?;
                            }
                            if pow < 0 {
                                s = s.recip()?; // s.ref_one() / s
                            }
                            NetworkLeaf::Scalar(self.add_scalar(s))
                        }
                    }
                    other => {
                        let mut t = match other {
                            NetworkLeaf::LibraryKey(_) => graph.get_lib_data(lib, child_id).unwrap().to_device(&self.dev)?,
                            NetworkLeaf::LocalTensor(ti) => self.tensors[*ti].clone(),
                            NetworkLeaf::Scalar(_) => unreachable!(),
                        };
                        match pow {
                            0 => NetworkLeaf::Scalar(self.add_scalar(self.one()?)),
                            1 => other.clone(),
                            _ => {
                                let squares = n / 2;
                                let mut square = t.contract(&t)?;
                                if n % 2 == 1 {
                                    for _ in 0..squares {
                                        square = square.contract(&square)?;
                                    }
                                    t = square.contract(&t)?;
                                    if pow < 0 {
                                        if !t.is_scalar() {
                                            return Err(TensorNetworkError::NegativeExponentNonScalar("".to_string()));
                                        }
                                        NetworkLeaf::Scalar(self.add_scalar(t.recip()?))
                                    } else {
                                        NetworkLeaf::LocalTensor(self.add_tensor(t))
                                    }
                                } else {
                                    if !square.is_scalar() {
                                        return Err(TensorNetworkError::Other(anyhow!("even power of a tensor that does not contract to a scalar")));
                                    }
                                    let mut s = square.clone();
                                    for _ in 1..squares {
                                        s = s.contract(&square)?;
                                    }
                                    if pow < 0 {
                                        s = s.recip()?;
                                    }
                                    NetworkLeaf::Scalar(self.add_scalar(s))
                                }
                            }
                        }
                    }
                };
                graph.identify_nodes_without_self_edges(&[child_id, opid], NetworkNode::Leaf(new_node));
                Ok(graph)
            }
        }
    }
}

/// Run a whole graph to one leaf on the device: `Sequential::execute_all::<DeviceSmallestDegree>` (network.rs:1126-1156).
pub fn execute_on_device<LT, L, K, FK, FL, Aind>(
    store: &mut DeviceStore,
    graph: &mut NetworkGraph<K, FK, Aind>,
    lib: &L,
    fn_lib: &FL,
) -> Result<(), TensorNetworkError<K, FK>>
where
    LT: LibraryTensor + Clone,
    LT::WithIndices: PermuteTensor<Permuted = LT::WithIndices> + ToDevice,
    <<LT::WithIndices as HasStructure>::Structure as TensorStructure>::Slot: IsAbstractSlot<Aind = Aind>,
    L: Library<<LT as HasStructure>::Structure, Key = K, Value = PermutedStructure<LT>>,
    K: Display + Debug + Clone,
    FK: Display + Debug + Clone,
    FL: FunctionLibrary<DeviceBatchTensor, DeviceBatchTensor, Key = FK>,
    Aind: AbsInd,
{
    use spenso::network::{ExecutionStrategy, Sequential};
    <Sequential as ExecutionStrategy<DeviceStore, FL, L, K, FK, Aind>>::execute_all::<DeviceSmallestDegree>(store, graph, lib, fn_lib)
}
