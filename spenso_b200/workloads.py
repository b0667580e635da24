"""The BASELINE.json workloads as neutral network descriptions (NetSpec).

A NetSpec is pure data (leaves + ops): `build_engine` turns it into a spenso_b200.Network; the
tests and bench.py's cpu_baseline leg turn the SAME description into an oracle network, so both
sides evaluate the same expression on the same inputs.  Definitions follow SURVEY.md §8(d):

  cfg1  4-gamma trace  p(1) (p+q)(3) gamma(1,2,1) gamma(2,3,2) gamma(3,4,3) gamma(4,1,4)
        (spenso-hep-lib/tests/gamma_algebra_validate.rs:29-70)
  cfg2  e+e- -> mu+mu- tree:  [vbar_i gamma^mu_ij u_j] [ubar_k gamma_mu,kl v_l]
  cfg3  SU(3) colour:  X[a,i,jbar] T[a,j,ibar] + D[a,b,c] f[a,b,c]   (Gell-Mann lambda/2, TR = 1/2)
  cfg5  one-loop numerator gl_03 (spenso-hep-lib/tests/gamma_algebra_validate.rs:160-248)
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

from . import (C64, FUSED, LIB_GAMMA_WEYL, LIB_METRIC, LIB_PROJM_WEYL, OP_NEG, OP_POWER, OP_PRODUCT, OP_SUM, Bispinor,
               ColorAdjoint, ColorFundamental, Minkowski, Network, Slot)


@dataclass
class NetSpec:
    name: str
    # leaves: ("batched", [Slot], input_name) | ("reindex", leaf_idx, [Slot]) | ("lib", which, [Slot])
    #         | ("custom", [Slot], {idx tuple: value}) | ("scalar", complex)
    # ops:    ("op", kind, [node idx], power)
    nodes: List[tuple] = field(default_factory=list)
    root: int = -1
    inputs: List[Tuple[str, int]] = field(default_factory=list)  # (name, components per point)
    input_dtypes: List[int] = field(default_factory=list)        # C64 / F64 per input
    # algorithmic bytes per evaluation: full dense inputs read once + result written once
    algorithmic_bytes: int = 0

    def add(self, *node) -> int:
        self.nodes.append(tuple(node))
        return len(self.nodes) - 1

    def batched(self, sl, name, dtype: int = C64) -> int:
        size = int(np.prod([s.dim for s in sl])) if sl else 1
        self.inputs.append((name, size))
        self.input_dtypes.append(dtype)
        return self.add("batched", list(sl), name, dtype)

    def reindex(self, leaf, sl) -> int:
        return self.add("reindex", leaf, list(sl))

    def lib(self, which, sl) -> int:
        return self.add("lib", which, list(sl))

    def custom(self, sl, entries) -> int:
        return self.add("custom", list(sl), dict(entries))

    def scalar(self, c) -> int:
        return self.add("scalar", complex(c))

    def function(self, sl, build) -> int:
        """Device-filled leaf: build(params_of) -> list of expr components, where params_of(node, n) gives the
        expr.Param components of an earlier batched node of THIS spec (ids are mapped when the network is built)."""
        return self.add("function", list(sl), build)

    def product(self, *ch) -> int:
        return self.add("op", OP_PRODUCT, list(ch), 1)

    def sum(self, *ch) -> int:
        return self.add("op", OP_SUM, list(ch), 1)

    def neg(self, c) -> int:
        return self.add("op", OP_NEG, [c], 1)

    def power(self, c, n) -> int:
        return self.add("op", OP_POWER, [c], n)


def build_engine(spec: NetSpec, ctx, mode: int = FUSED, schedule=None) -> Network:
    """NetSpec -> compiled spenso_b200.Network (ctx may be None for a plan-only network).  `schedule`: the host's
    contraction order as (n1, n2) pairs of SPEC node indices (Network.set_schedule naming)."""
    net = Network(ctx)
    ids: Dict[int, int] = {}
    for k, node in enumerate(spec.nodes):
        kind = node[0]
        if kind == "batched":
            ids[k] = net.batched(node[1], node[3])
        elif kind == "reindex":
            ids[k] = net.reindex(ids[node[1]], node[2])
        elif kind == "lib":
            ids[k] = net.library(node[1], node[2])
        elif kind == "custom":
            ids[k] = net.library_custom(node[1], node[2])
        elif kind == "scalar":
            ids[k] = net.scalar(node[1])
        elif kind == "function":
            from .expr import Param

            ids[k] = net.function(node[1], node[2](lambda n, size: Param.vector(ids[n], size)))
        else:
            ids[k] = net.op(node[1], [ids[c] for c in node[2]], node[3])
    net.set_root(ids[spec.root])
    net.spec_ids = ids
    if schedule is not None:
        net.set_schedule([(ids[a], ids[b]) for a, b in schedule])
    net.compile(mode)
    return net


def _b(a):
    return Bispinor.new_slot(4, a)


def _m(a):
    return Minkowski.new_slot(4, a)


# ---- cfg1 ---------------------------------------------------------------------------------------------
def gamma_trace4() -> NetSpec:
    s = NetSpec("cfg1_gamma_trace4")
    p1 = s.batched([_m(1)], "p")
    q3 = s.batched([_m(3)], "q")
    p3 = s.reindex(p1, [_m(3)])
    g = [s.lib(LIB_GAMMA_WEYL, [_b(1), _b(2), _m(1)]), s.lib(LIB_GAMMA_WEYL, [_b(2), _b(3), _m(2)]),
         s.lib(LIB_GAMMA_WEYL, [_b(3), _b(4), _m(3)]), s.lib(LIB_GAMMA_WEYL, [_b(4), _b(1), _m(4)])]
    s.root = s.product(p1, s.sum(p3, q3), *g)
    s.algorithmic_bytes = 2 * 64 + 256
    return s


def gamma_trace4_closed_form(p: np.ndarray, q: np.ndarray) -> np.ndarray:
    """R[mu2, mu4] = 4 (p^mu2 r^mu4 + r^mu2 p^mu4 - (p.r) eta^{mu2 mu4}), r = p + q."""
    eta = np.diag([1.0, -1.0, -1.0, -1.0])
    r = p + q
    pr = np.einsum("...i,ij,...j->...", p, eta, r)
    return 4 * (np.einsum("...i,...j->...ij", p, r) + np.einsum("...i,...j->...ij", r, p)
                - pr[..., None, None] * eta)


# ---- cfg2 ---------------------------------------------------------------------------------------------
def eemumu() -> NetSpec:
    s = NetSpec("cfg2_eemumu_tree")
    vb = s.batched([_b(1)], "vbar_p2")
    u = s.batched([_b(2)], "u_p1")
    ub = s.batched([_b(3)], "ubar_p3")
    v = s.batched([_b(4)], "v_p4")
    g1 = s.lib(LIB_GAMMA_WEYL, [_b(1), _b(2), _m(1)])
    g2 = s.lib(LIB_GAMMA_WEYL, [_b(3), _b(4), _m(1)])
    s.root = s.product(vb, g1, u, ub, g2, v)
    s.algorithmic_bytes = 4 * 64 + 16
    return s


def massless_spinors(p):
    """Helicity spinors of a massless momentum p = (E, px, py, pz) in the chiral (Weyl) basis of
    spenso-hep-lib's gamma_data_weyl (gamma^0 = [[0,1],[1,0]], gamma^k = [[0,s_k],[-s_k,0]]), as expression
    components.  u_plus solves pslash u = 0 with gamma^5 u = +u; ubar = u^dagger gamma^0."""
    from .expr import re, sqrt

    E, px, py, pz = (re(c) for c in p)
    a = sqrt(E + pz)                       # sqrt(p+)
    b = (px + 1j * py) / a                 # p_perp / sqrt(p+)
    u_plus = [0.0, 0.0, a, b]              # right-handed: lower two components
    u_minus = [-b.conj(), a, 0.0, 0.0]     # left-handed: upper two components
    ubar_plus = [a, b.conj(), 0.0, 0.0]    # u_plus^dagger gamma^0
    ubar_minus = [0.0, 0.0, -b, a]         # u_minus^dagger gamma^0
    return {"u+": u_plus, "u-": u_minus, "ubar+": ubar_plus, "ubar-": ubar_minus}


def eemumu_from_momenta(dtype: int = C64) -> NetSpec:
    """cfg2 with device-side leaf filling: the per-point inputs are the four massless momenta
    (4 x 64 B as c64 Minkowski vectors, or 4 x 32 B as f64) and the spinors vbar(p2), u(p1), ubar(p3), v(p4)
    (fixed helicities; massless v(p) of helicity h == u(p) of helicity -h) are built inside the kernel."""
    s = NetSpec("cfg2m_eemumu_from_momenta" + ("_f64" if dtype != C64 else ""))
    p = [s.batched([_m(100 + k)], f"p{k + 1}", dtype) for k in range(4)]
    sp1 = lambda k, key, a: s.function([_b(a)], lambda P, k=k, key=key: massless_spinors(P(p[k], 4))[key])
    vb = sp1(1, "ubar-", 1)   # vbar_+(p2) = ubar_-(p2)
    u = sp1(0, "u-", 2)
    ub = sp1(2, "ubar-", 3)
    v = sp1(3, "u-", 4)       # v_+(p4) = u_-(p4)
    g1 = s.lib(LIB_GAMMA_WEYL, [_b(1), _b(2), _m(1)])
    g2 = s.lib(LIB_GAMMA_WEYL, [_b(3), _b(4), _m(1)])
    s.root = s.product(vb, g1, u, ub, g2, v)
    s.algorithmic_bytes = 4 * (64 if dtype == C64 else 32) + 16
    return s


def eemumu_from_real_momenta() -> NetSpec:
    from . import F64

    return eemumu_from_momenta(F64)


def massless_momenta(n_points: int, seed: int):
    """Four massless momenta per point (random directions and energies; no momentum conservation needed for a
    throughput workload), [n, 4] complex128 each with zero imaginary parts."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(4):
        E = rng.uniform(10.0, 100.0, n_points)
        ct = rng.uniform(-0.95, 0.95, n_points)
        ph = rng.uniform(0, 2 * np.pi, n_points)
        st = np.sqrt(1 - ct * ct)
        out.append(np.stack([E, E * st * np.cos(ph), E * st * np.sin(ph), E * ct], axis=1).astype(np.complex128))
    return out


# ---- cfg3 ---------------------------------------------------------------------------------------------
def gell_mann_T() -> dict:
    """T^a_{i j} = lambda^a_{ij} / 2 as {(a, i, j): value}."""
    lam = np.zeros((8, 3, 3), dtype=np.complex128)
    lam[0][0, 1] = lam[0][1, 0] = 1
    lam[1][0, 1], lam[1][1, 0] = -1j, 1j
    lam[2][0, 0], lam[2][1, 1] = 1, -1
    lam[3][0, 2] = lam[3][2, 0] = 1
    lam[4][0, 2], lam[4][2, 0] = -1j, 1j
    lam[5][1, 2] = lam[5][2, 1] = 1
    lam[6][1, 2], lam[6][2, 1] = -1j, 1j
    lam[7][0, 0] = lam[7][1, 1] = 1 / np.sqrt(3)
    lam[7][2, 2] = -2 / np.sqrt(3)
    T = lam / 2
    return {(a, i, j): T[a, i, j] for a in range(8) for i in range(3) for j in range(3) if T[a, i, j] != 0}


def su3_f() -> dict:
    """f^{abc} = -2i Tr(T^a [T^b, T^c]) as {(a, b, c): value}; exact values set by hand."""
    base = {(0, 1, 2): 1.0, (0, 3, 6): 0.5, (0, 4, 5): -0.5, (1, 3, 5): 0.5, (1, 4, 6): 0.5, (2, 3, 4): 0.5,
            (2, 5, 6): -0.5, (3, 4, 7): np.sqrt(3) / 2, (5, 6, 7): np.sqrt(3) / 2}
    out = {}
    for (a, b, c), v in base.items():
        for perm, sgn in (((a, b, c), 1), ((b, c, a), 1), ((c, a, b), 1), ((b, a, c), -1), ((a, c, b), -1),
                          ((c, b, a), -1)):
            out[perm] = sgn * v
    return out


def su3_colour() -> NetSpec:
    s = NetSpec("cfg3_su3_colour")
    co = lambda a: ColorAdjoint.new_slot(8, a)
    cf = lambda a: ColorFundamental.new_slot(3, a)
    cfb = lambda a: ColorFundamental.dual().new_slot(3, a)
    X = s.batched([co(1), cf(2), cfb(3)], "X")
    D = s.batched([co(4), co(5), co(6)], "D")
    # T[a, j, ibar]: contracts X's cof|2 with coaf|2 and X's coaf|3 with cof|3
    T = s.custom([co(1), cfb(2), cf(3)], gell_mann_T())
    f = s.custom([co(4), co(5), co(6)], su3_f())
    s.root = s.sum(s.product(X, T), s.product(D, f))
    s.algorithmic_bytes = 72 * 16 + 512 * 16 + 16
    return s


# ---- cfg5 ---------------------------------------------------------------------------------------------
GL03_MC, GL03_MW = 11232.0, 1231.0


def one_loop_gl03(P0=(0.0, 1.0, 2.0, 3.0), massive: bool = False, dtype: int = C64) -> NetSpec:
    """gl_03 of the reference's gamma_algebra_validate.rs: inputs K0, K1 (per point); P0, MC, MW fixed.

    The literal gl_03 numerator vanishes identically: P- (MC - K0slash) gamma P- kills the mass term
    and what is left is a trace of P- with five gamma matrices.  (The reference test only checks
    that it equals its simplified form.)  `massive=True` is the benchmark variant of the same
    topology: the second fermion propagator numerator V5slash becomes (MC + V5slash), which leaves
    -MC Tr[K0slash gamma^h7 P- gamma^h11 V3slash] W_{h7 h8} g^{h0 h8} != 0.
    """
    s = NetSpec(("cfg5_one_loop_massive" if massive else "gl03_literal") + ("_f64" if dtype != C64 else ""))
    H = lambda n: n              # hedge(n)
    E = lambda a: 100 + a        # edge(a, 1)
    V11 = 201                    # vertex(1, 1)
    K0 = {E(1): s.batched([_m(E(1))], "K0", dtype)}
    for lab in (E(3), E(5)):
        K0[lab] = s.reindex(K0[E(1)], [_m(lab)])
    K1 = {E(3): s.batched([_m(E(3))], "K1", dtype)}
    for lab in (H(7), H(8), E(5)):
        K1[lab] = s.reindex(K1[E(3)], [_m(lab)])
    P = {lab: s.custom([_m(lab)], {(i,): P0[i] for i in range(4) if P0[i] != 0}) for lab in (H(7), H(8), E(5))}
    pref = s.scalar((1.0 / 6.0 ** (4.0 ** -2.0)) * np.sqrt(0.5))
    # (MC g(bis h1, bis h2) - K0(e11) gamma(h1, h2, e11))
    f1 = s.sum(s.product(s.scalar(GL03_MC), s.lib(LIB_METRIC, [_b(H(1)), _b(H(2))])),
               s.neg(s.product(K0[E(1)], s.lib(LIB_GAMMA_WEYL, [_b(H(1)), _b(H(2)), _m(E(1))]))))
    v3 = s.sum(s.neg(K0[E(3)]), s.neg(K1[E(3)]))
    w = s.sum(s.neg(s.lib(LIB_METRIC, [_m(H(7)), _m(H(8))])),
              s.product(s.scalar(GL03_MW ** -2), s.sum(s.neg(P[H(7)]), s.neg(K1[H(7)])),
                        s.sum(s.neg(P[H(8)]), s.neg(K1[H(8)]))))
    v5 = s.sum(P[E(5)], K0[E(5)], K1[E(5)])
    v5slash = [v5, s.lib(LIB_GAMMA_WEYL, [_b(H(9)), _b(H(10)), _m(E(5))])]
    if massive:
        v5slash = [s.sum(s.product(s.scalar(GL03_MC), s.lib(LIB_METRIC, [_b(H(9)), _b(H(10))])), s.product(*v5slash))]
    s.root = s.product(
        pref, f1, v3, w, *v5slash,
        s.lib(LIB_METRIC, [_m(H(0)), _m(H(8))]),
        s.lib(LIB_GAMMA_WEYL, [_b(H(10)), _b(H(6)), _m(H(11))]),
        s.lib(LIB_GAMMA_WEYL, [_b(H(2)), _b(V11), _m(H(7))]),
        s.lib(LIB_GAMMA_WEYL, [_b(H(6)), _b(H(5)), _m(E(3))]),
        s.lib(LIB_PROJM_WEYL, [_b(H(5)), _b(H(1))]),
        s.lib(LIB_PROJM_WEYL, [_b(V11), _b(H(9))]))
    s.algorithmic_bytes = 2 * (64 if dtype == C64 else 32)
    return s


def one_loop_massive() -> NetSpec:
    return one_loop_gl03(massive=True)


def one_loop_massive_real() -> NetSpec:
    """cfg5 with the (real) loop momenta declared as f64 tensors: half the bytes, and every term that would
    multiply an imaginary part is gone from the kernel."""
    from . import F64

    return one_loop_gl03(massive=True, dtype=F64)


WORKLOADS = {"cfg1": gamma_trace4, "cfg2": eemumu, "cfg3": su3_colour, "cfg5": one_loop_massive,
             "gl03": one_loop_gl03, "cfg2m": eemumu_from_momenta, "cfg2mr": eemumu_from_real_momenta,
             "cfg5r": one_loop_massive_real}


def synthetic_inputs(spec: NetSpec, n_points: int, seed: int, scale: float = 1.0, real: bool = False) -> List[np.ndarray]:
    """i.i.d. U(-1,1) re/im per component, [n_points, size] complex128 per input, canonical order."""
    rng = np.random.default_rng(seed)
    out = []
    for _, size in spec.inputs:
        a = rng.uniform(-1.0, 1.0, size=(n_points, size, 2)) * scale
        if real:
            a[..., 1] = 0.0
        out.append(np.ascontiguousarray(a).view(np.complex128).reshape(n_points, size))
    return out
