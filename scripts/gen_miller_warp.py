#!/usr/bin/env python3
"""Assembler + checker for the warp-per-pairing Miller program (lighthouse_b200/csrc/bls/miller_warp.cuh).

Small batches are latency bound: one lane (or six) per SignatureSet leaves a 63-iteration chain of ~60 dependent
Fp products per iteration.  miller_warp.cuh gives ONE WARP to one pairing and runs the arithmetic at Fp granularity:
every lane computes one Fp value per phase,

    MUL phase (K terms):   slot[d] = ( sum_{q<K} X_q * slot[y_q] ) / R mod p,   X_q = +-slot[x_q] or +-2 slot[x_q]
    LIN phase:             slot[d] = ( sum_q c_q * slot[s_q] ) / 2^h mod p,     c_q small signed integers

with a warp barrier between phases.  This script writes the phase tables (which lane computes what) from the same
formulas bls/miller_coop.cuh uses, runs them on Python integers and checks the result against the oracle:
    * one whole Miller loop + the oracle's final exponentiation == the oracle's pairing,
    * the dense product section == f12_mul.
It then emits bls/miller_warp_tables.inc.  Run:  python scripts/gen_miller_warp.py
"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bls_ref as B  # noqa: E402

P = B.P
KMAX = 4
NLANES = 32
# x-operand modes of a MUL term
X_ZERO, X_POS, X_NEG, X_DBL, X_NEGDBL = 0, 1, 2, 3, 4


class Prog:
    def __init__(self):
        self.slots = {}
        self.sections = {}      # name -> list of phases
        self.cur = None

    def slot(self, name):
        if name not in self.slots:
            self.slots[name] = len(self.slots)
        return self.slots[name]

    def section(self, name):
        self.cur = []
        self.sections[name] = self.cur

    # ---- phases
    def mul_phase(self, K, ops):
        """ops: list of (dst, [(mode, xslot, yslot), ...])  one per lane"""
        assert 1 <= K <= KMAX and len(ops) <= NLANES, (K, len(ops))
        for d, terms in ops:
            assert len(terms) <= K
        self.cur.append(("mul", K, ops))

    def lin_phase(self, ops):
        """ops: list of (dst, [(coef, slot), ...], halvings)"""
        assert len(ops) <= NLANES, len(ops)
        for d, terms, h in ops:
            assert 1 <= len(terms) <= 4 and all(abs(c) in (1, 2, 3, 12) for c, _ in terms) and 0 <= h <= 2
        self.cur.append(("lin", 0, ops))


def run_section(prog, name, mem):
    for kind, K, ops in prog.sections[name]:
        out = []
        if kind == "mul":
            for d, terms in ops:
                acc = 0
                for mode, xs, ys in terms:
                    x = mem[xs]
                    xv = {X_ZERO: 0, X_POS: x, X_NEG: P - x, X_DBL: 2 * x, X_NEGDBL: 2 * P - 2 * x}[mode]
                    assert 0 <= mem[ys] < P and 0 <= xv <= 2 * P
                    acc += xv * mem[ys]
                out.append((d, acc % P))       # (Montgomery factors cancel: values here are plain residues)
        else:
            for d, terms, h in ops:
                acc = sum(c * mem[s] for c, s in terms) % P
                for _ in range(h):
                    acc = acc * pow(2, -1, P) % P
                out.append((d, acc))
        for d, v in out:                       # all lanes read before any lane writes (barrier semantics)
            mem[d] = v


# ------------------------------------------------------------------------------------------------ Fp2-level builders
class B2:
    """collects Fp2-level operations of one phase and lowers them to lane ops"""

    def __init__(self, prog):
        self.p = prog
        self.mops = []
        self.lops = []

    def s2(self, name):
        return (self.p.slot(name + ".0"), self.p.slot(name + ".1"))

    # d = a * b   (K = 2):  re = a0 b0 - a1 b1,  im = a0 b1 + a1 b0
    def mul(self, d, a, b):
        d, a, b = self.s2(d), self.s2(a), self.s2(b)
        self.mops.append((d[0], [(X_POS, a[0], b[0]), (X_NEG, a[1], b[1])]))
        self.mops.append((d[1], [(X_POS, a[0], b[1]), (X_POS, a[1], b[0])]))

    # d = a^2:  re = a0 a0 - a1 a1,  im = 2 a0 a1
    def sqr(self, d, a):
        d, a = self.s2(d), self.s2(a)
        self.mops.append((d[0], [(X_POS, a[0], a[0]), (X_NEG, a[1], a[1])]))
        self.mops.append((d[1], [(X_DBL, a[0], a[1])]))

    # d = a * s, s an Fp slot
    def mul_fp(self, d, a, s):
        d, a = self.s2(d), self.s2(a)
        s = self.p.slot(s)
        self.mops.append((d[0], [(X_POS, s, a[0])]))
        self.mops.append((d[1], [(X_POS, s, a[1])]))

    def flush_mul(self, K=2):
        if self.mops:
            self.p.mul_phase(K, self.mops)
            self.mops = []

    # d = sum c_k * a_k  (component-wise), / 2^h
    def lin(self, d, terms, h=0):
        d = self.s2(d)
        for comp in (0, 1):
            self.lops.append((d[comp], [(c, self.s2(a)[comp]) for c, a in terms], h))

    # d = k * xi * a:   re = k (a0 - a1),  im = k (a0 + a1)
    def lin_xi(self, d, a, k):
        d, a = self.s2(d), self.s2(a)
        self.lops.append((d[0], [(k, a[0]), (-k, a[1])], 0))
        self.lops.append((d[1], [(k, a[0]), (k, a[1])], 0))

    def flush_lin(self):
        if self.lops:
            self.p.lin_phase(self.lops)
            self.lops = []


def coef_slots(prog, base, j):
    """the four stored forms of coefficient j of an Fp12 value: a0, a1, s = a0 + a1, d = a0 - a1"""
    return [prog.slot(f"{base}{j}.{c}") for c in ("0", "1", "s", "d")]


def build():
    pr = Prog()
    # fixed slots first (the kernel addresses them by name): f coefficients, T, Q, P
    for j in range(6):
        coef_slots(pr, "f", j)
    for n in ("X", "Y", "Z", "QX", "QY", "QZ", "HX", "HY", "HZ"):
        pr.slot(n + ".0"); pr.slot(n + ".1")
    for n in ("px", "py", "pz"):
        pr.slot(n)
    for j in range(6):
        coef_slots(pr, "g", j)
    pr.slot("dummy")

    # ---------------------------------------------------------------- INIT: H (Jacobian) -> T = Q = (X Z, Y, Z^3)
    pr.section("init")
    b = B2(pr)
    b.mul("X", "HX", "HZ"); b.sqr("t0", "HZ"); b.flush_mul()
    b.mul("Z", "t0", "HZ"); b.flush_mul()
    b.lin("Y", [(1, "HY")]); b.lin("QX", [(1, "X")]); b.lin("QY", [(1, "HY")]); b.lin("QZ", [(1, "Z")]); b.flush_lin()

    # ---------------------------------------------------------------- DBL (Costello-Lange-Naehrig, as miller_coop.cuh)
    # A = X Y / 2, B = Y^2, C = Z^2, E = 3 b' C = 12 xi C, F = 3 E, X3 = A (B - F), G = (B + F) / 2, Y3 = G^2 - 3 E^2,
    # H = (Y + Z)^2 - B - C, Z3 = B H;  line c0 = (B - E) pz, c1 = -3 X^2 px, c4 = H py
    pr.section("dbl")
    b = B2(pr)
    b.lin("s1", [(1, "X"), (1, "Y")]); b.lin("s2", [(1, "Y"), (1, "Z")]); b.flush_lin()
    b.sqr("XX", "X"); b.sqr("B", "Y"); b.sqr("C", "Z"); b.sqr("S2", "s2"); b.sqr("S1", "s1"); b.flush_mul()
    b.lin_xi("E", "C", 12)
    b.lin("H", [(1, "S2"), (-1, "B"), (-1, "C")])
    b.lin("A", [(1, "S1"), (-1, "XX"), (-1, "B")], h=2)          # (2 X Y) / 4
    b.flush_lin()
    b.lin("BmF", [(1, "B"), (-3, "E")]); b.lin("BmE", [(1, "B"), (-1, "E")]); b.lin("G", [(1, "B"), (3, "E")], h=1)
    b.flush_lin()
    b.mul("X", "A", "BmF"); b.mul("Z", "B", "H"); b.sqr("GG", "G"); b.sqr("EE", "E")
    b.mul_fp("l1p", "XX", "px"); b.mul_fp("l0", "BmE", "pz"); b.mul_fp("l4", "H", "py"); b.flush_mul()
    b.lin("Y", [(1, "GG"), (-3, "EE")]); b.lin("l1", [(-3, "l1p")]); b.flush_lin()

    # ---------------------------------------------------------------- ADD (T <- T + Q, add-1998-cmo-2, as miller_coop.cuh)
    pr.section("add")
    b = B2(pr)
    b.mul("d", "Y", "QZ"); b.mul("e", "X", "QZ"); b.mul("ff", "Z", "QZ"); b.mul("y2z1", "QY", "Z"); b.mul("x2z1", "QX", "Z")
    b.flush_mul()
    b.lin("u", [(1, "y2z1"), (-1, "d")]); b.lin("v", [(1, "x2z1"), (-1, "e")]); b.flush_lin()
    b.sqr("vv", "v"); b.sqr("uu", "u"); b.mul("ux2", "u", "QX"); b.mul("vy2", "v", "QY"); b.mul("uz2", "u", "QZ")
    b.mul("vz2", "v", "QZ"); b.flush_mul()
    b.lin("l0p", [(1, "ux2"), (-1, "vy2")]); b.flush_lin()
    b.mul("vvv", "v", "vv"); b.mul("R", "vv", "e"); b.mul("uuz", "uu", "ff"); b.mul_fp("l0", "l0p", "pz")
    b.mul_fp("l1p", "uz2", "px"); b.mul_fp("l4", "vz2", "py"); b.flush_mul()
    b.lin("Aa", [(1, "uuz"), (-1, "vvv"), (-2, "R")]); b.lin("l1", [(-1, "l1p")]); b.flush_lin()
    b.lin("RmA", [(1, "R"), (-1, "Aa")]); b.flush_lin()
    b.mul("Z", "vvv", "ff"); b.mul("dY", "vvv", "d"); b.mul("X", "v", "Aa"); b.mul("uRA", "u", "RmA"); b.flush_mul()
    b.lin("Y", [(1, "uRA"), (-1, "dY")]); b.flush_lin()

    # ---------------------------------------------------------------- products into f (w-basis, w^6 = xi)
    def product_section(name, terms_of):
        """terms_of(t) -> list of (xname (Fp2 slot name), j, c, xi): out_t = sum c * [xi] * x * a_j over the f
        coefficients a_j.  Schoolbook on Fp level with the stored forms (a0, a1, s, d) of a_j so that a multiplication by
        xi costs nothing:  re = c (x0 a0 - x1 a1) | c (x0 d - x1 s),   im = c (x0 a1 + x1 a0) | c (x0 s + x1 d)."""
        pr.section(name)
        parts = {}
        ops = []
        for t in range(6):
            terms = terms_of(t)
            re, im = [], []
            for xn, j, c, xi in terms:
                x0, x1 = pr.slot(xn + ".0"), pr.slot(xn + ".1")
                a0, a1, s, d = coef_slots(pr, "f", j)
                pos, neg = (X_POS, X_NEG) if c == 1 else (X_DBL, X_NEGDBL)
                if not xi:
                    re += [(pos, x0, a0), (neg, x1, a1)]
                    im += [(pos, x0, a1), (pos, x1, a0)]
                else:
                    re += [(pos, x0, d), (neg, x1, s)]
                    im += [(pos, x0, s), (pos, x1, d)]
            for comp, lst in (("0", re), ("1", im)):
                chunks = [lst[i:i + KMAX] for i in range(0, len(lst), KMAX)]
                parts[(t, comp)] = []
                for ci, ch in enumerate(chunks):
                    dst = pr.slot(f"n{t}.{comp}.{ci}")
                    parts[(t, comp)].append(dst)
                    ops.append((dst, ch))
        # mul phases of at most 32 lanes
        for i in range(0, len(ops), NLANES):
            pr.mul_phase(KMAX, ops[i:i + NLANES])
        nparts = max(len(v) for v in parts.values())
        if 2 * nparts <= 4:
            # one LIN phase: a0, a1 and the stored forms s = a0 + a1, d = a0 - a1 straight from the partial sums
            lops = []
            for t in range(6):
                a0, a1, s, d = coef_slots(pr, "f", t)
                p0, p1 = parts[(t, "0")], parts[(t, "1")]
                lops.append((a0, [(1, x) for x in p0], 0))
                lops.append((a1, [(1, x) for x in p1], 0))
                lops.append((s, [(1, x) for x in p0] + [(1, x) for x in p1], 0))
                lops.append((d, [(1, x) for x in p0] + [(-1, x) for x in p1], 0))
            pr.lin_phase(lops)
            return
        lops = []
        for t in range(6):
            for comp in ("0", "1"):
                lops.append((pr.slot(f"f{t}.{comp}"), [(1, s) for s in parts[(t, comp)]], 0))
        pr.lin_phase(lops)
        lops = []
        for t in range(6):
            a0, a1, s, d = coef_slots(pr, "f", t)
            lops.append((s, [(1, a0), (1, a1)], 0))
            lops.append((d, [(1, a0), (-1, a1)], 0))
        pr.lin_phase(lops)

    # f^2: out_t = sum_{i<=j, i+j = t} c a_i a_j + xi sum_{i<=j, i+j = t+6} c a_i a_j   (c = 2 unless i == j)
    def sqr_terms(t):
        out = []
        for i in range(6):
            for j in range(i, 6):
                if i + j == t or i + j == t + 6:
                    out.append((f"f{i}", j, 1 if i == j else 2, i + j == t + 6))
        return out
    product_section("sqr", sqr_terms)

    # f * (l0 + l1 w^2 + l4 w^3): out_t = l0 a_t + l1 a_{t-2} + l4 a_{t-3}  (indices below 0 wrap with xi)
    def sparse_terms(t):
        out = []
        for ln, k in (("l0", 0), ("l1", 2), ("l4", 3)):
            j = t - k
            out.append((ln, j % 6, 1, j < 0))
        return out
    product_section("sparse", sparse_terms)

    # f * g (g: another warp's value, copied into the g slots): out_t = sum_{i+j = t} g_i a_j + xi sum_{i+j = t+6} g_i a_j
    def dense_terms(t):
        out = []
        for i in range(6):
            for j in range(6):
                if i + j == t or i + j == t + 6:
                    out.append((f"g{i}", j, 1, i + j == t + 6))
        return out
    product_section("dense", dense_terms)

    # conj: negate the odd coefficients (and refresh s, d)
    pr.section("conj")
    lops = []
    for t in (1, 3, 5):
        a0, a1, s, d = coef_slots(pr, "f", t)
        lops += [(a0, [(-1, a0)], 0), (a1, [(-1, a1)], 0), (s, [(-1, s)], 0), (d, [(-1, d)], 0)]
    pr.lin_phase(lops)
    return pr


# ------------------------------------------------------------------------------------------------ checks
def f12_from_mem(pr, mem, base="f"):
    k = [(mem[pr.slots[f"{base}{j}.0"]], mem[pr.slots[f"{base}{j}.1"]]) for j in range(6)]
    return ((k[0], k[2], k[4]), (k[1], k[3], k[5]))


def f12_to_mem(pr, mem, v, base):
    k = [v[0][0], v[1][0], v[0][1], v[1][1], v[0][2], v[1][2]]
    for j in range(6):
        a0, a1 = k[j]
        for c, val in (("0", a0), ("1", a1), ("s", (a0 + a1) % P), ("d", (a0 - a1) % P)):
            mem[pr.slots[f"{base}{j}.{c}"]] = val


def check(pr):
    rnd = random.Random(2024)
    mem = [0] * len(pr.slots)
    Pt = B.g1_mul(B.G1_GEN, rnd.randrange(1, B.R))
    Q = B.g2_mul(B.G2_GEN, rnd.randrange(1, B.R))
    # H as a Jacobian point with a random Z, P projective with a random pz
    z = (rnd.randrange(1, P), rnd.randrange(1, P))
    z2 = B.f2_sqr(z)
    H = (B.f2_mul(Q[0], z2), B.f2_mul(Q[1], B.f2_mul(z2, z)), z)
    for n, v in zip(("HX", "HY", "HZ"), H):
        mem[pr.slots[n + ".0"]], mem[pr.slots[n + ".1"]] = v
    pz = rnd.randrange(1, P)
    mem[pr.slots["px"]], mem[pr.slots["py"]], mem[pr.slots["pz"]] = Pt[0] * pz % P, Pt[1] * pz % P, pz
    f12_to_mem(pr, mem, B.F12_ONE, "f")
    run_section(pr, "init", mem)
    for i in range(B.X_ABS.bit_length() - 2, -1, -1):
        run_section(pr, "sqr", mem)
        run_section(pr, "dbl", mem)
        run_section(pr, "sparse", mem)
        if (B.X_ABS >> i) & 1:
            run_section(pr, "add", mem)
            run_section(pr, "sparse", mem)
    run_section(pr, "conj", mem)
    f = f12_from_mem(pr, mem)
    assert B.final_exp(f) == B.pairing(Pt, Q), "Miller program disagrees with the oracle pairing"
    for j in range(6):   # stored forms consistent
        a0, a1, s, d = (mem[x] for x in coef_slots(pr, "f", j))
        assert s == (a0 + a1) % P and d == (a0 - a1) % P
    g = B.f12_pow(f, 5)
    f12_to_mem(pr, mem, g, "g")
    run_section(pr, "dense", mem)
    assert f12_from_mem(pr, mem) == B.f12_mul(f, g), "dense product section disagrees with f12_mul"



# ================================================================================================ G2 programs
# Homogeneous projective points (X : Y : Z) on E2: y^2 = x^3 + 4 xi.  Accumulators A (and A2), base Bp; the same
# doubling (Costello-Lange-Naehrig) and addition (add-1998-cmo-2) formulas as the Miller sections, without the lines.
def g2_dbl_stages(b, pts):
    """T <- 2 T for every point name in pts, all in the same phases (temporaries are suffixed with the point name)"""
    for T in pts:
        b.lin(f"s1{T}", [(1, f"{T}X"), (1, f"{T}Y")]); b.lin(f"s2{T}", [(1, f"{T}Y"), (1, f"{T}Z")])
    b.flush_lin()
    for T in pts:
        b.sqr(f"XX{T}", f"{T}X"); b.sqr(f"B{T}", f"{T}Y"); b.sqr(f"C{T}", f"{T}Z"); b.sqr(f"S2{T}", f"s2{T}")
        b.sqr(f"S1{T}", f"s1{T}")
    b.flush_mul()
    for T in pts:
        b.lin_xi(f"E{T}", f"C{T}", 12)
        b.lin(f"H{T}", [(1, f"S2{T}"), (-1, f"B{T}"), (-1, f"C{T}")])
        b.lin(f"Ah{T}", [(1, f"S1{T}"), (-1, f"XX{T}"), (-1, f"B{T}")], h=2)
    b.flush_lin()
    for T in pts:
        b.lin(f"BmF{T}", [(1, f"B{T}"), (-3, f"E{T}")]); b.lin(f"G{T}", [(1, f"B{T}"), (3, f"E{T}")], h=1)
    b.flush_lin()
    for T in pts:
        b.mul(f"{T}X", f"Ah{T}", f"BmF{T}"); b.mul(f"{T}Z", f"B{T}", f"H{T}"); b.sqr(f"GG{T}", f"G{T}")
        b.sqr(f"EE{T}", f"E{T}")
    b.flush_mul()
    for T in pts:
        b.lin(f"{T}Y", [(1, f"GG{T}"), (-3, f"EE{T}")])
    b.flush_lin()


def g2_add_stages(b, T, Q):
    """T <- T + Q (both projective; T != +-Q, neither at infinity: see the callers)"""
    b.mul("d", f"{T}Y", f"{Q}Z"); b.mul("e", f"{T}X", f"{Q}Z"); b.mul("ff", f"{T}Z", f"{Q}Z")
    b.mul("y2z1", f"{Q}Y", f"{T}Z"); b.mul("x2z1", f"{Q}X", f"{T}Z"); b.flush_mul()
    b.lin("u", [(1, "y2z1"), (-1, "d")]); b.lin("v", [(1, "x2z1"), (-1, "e")]); b.flush_lin()
    b.sqr("vv", "v"); b.sqr("uu", "u"); b.flush_mul()
    b.mul("vvv", "v", "vv"); b.mul("R", "vv", "e"); b.mul("uuz", "uu", "ff"); b.flush_mul()
    b.lin("Aa", [(1, "uuz"), (-1, "vvv"), (-2, "R")]); b.flush_lin()
    b.lin("RmA", [(1, "R"), (-1, "Aa")]); b.flush_lin()
    b.mul(f"{T}Z", "vvv", "ff"); b.mul("dY", "vvv", "d"); b.mul(f"{T}X", "v", "Aa"); b.mul("uRA", "u", "RmA"); b.flush_mul()
    b.lin(f"{T}Y", [(1, "uRA"), (-1, "dY")]); b.flush_lin()


def g2_copy(b, D, S, neg=False):
    b.lin(f"{D}X", [(1, f"{S}X")]); b.lin(f"{D}Y", [(-1 if neg else 1, f"{S}Y")]); b.lin(f"{D}Z", [(1, f"{S}Z")])


def g2_psi(b, D, S):
    """D <- psi(S) = (conj(X) CX : conj(Y) CY : conj(Z)); one MUL phase + one LIN phase (caller flushes)"""
    for c, k in (("X", "CX"), ("Y", "CY")):
        d, a, kk = b.s2(f"{D}{c}"), b.s2(f"{S}{c}"), b.s2(k)
        b.mops.append((d[0], [(X_POS, a[0], kk[0]), (X_POS, a[1], kk[1])]))      # (a0 - a1 i)(k0 + k1 i)
        b.mops.append((d[1], [(X_POS, a[0], kk[1]), (X_NEG, a[1], kk[0])]))
    dz, az = b.s2(f"{D}Z"), b.s2(f"{S}Z")
    b.lops.append((dz[0], [(1, az[0])], 0)); b.lops.append((dz[1], [(-1, az[1])], 0))


def build_g2():
    pr = Prog()
    for n in ("AX", "AY", "AZ", "A2X", "A2Y", "A2Z", "BpX", "BpY", "BpZ", "HX", "HY", "HZ", "CX", "CY", "JX", "JY", "JZ"):
        pr.slot(n + ".0"); pr.slot(n + ".1")
    for n in ("c2x", "e1.0", "e1.1", "e2.0", "e2.1", "dummy"):
        pr.slot(n)
    b = B2(pr)
    # ---- shared by both kernels
    pr.section("setA"); g2_copy(b, "A", "Bp"); b.flush_lin()
    pr.section("setA2"); g2_copy(b, "A2", "Bp"); b.flush_lin()
    pr.section("dblA"); g2_dbl_stages(b, ["A"])
    pr.section("dblAA2"); g2_dbl_stages(b, ["A", "A2"])
    pr.section("addA"); g2_add_stages(b, "A", "Bp")
    pr.section("addA2"); g2_add_stages(b, "A2", "Bp")
    # A (projective) -> Jacobian (X Z, Y Z^2, Z) in J
    pr.section("tojac")
    b.mul("JX", "AX", "AZ"); b.sqr("zz", "AZ"); b.flush_mul()
    b.mul("JY", "AY", "zz"); b.flush_mul()
    b.lin("JZ", [(1, "AZ")]); b.flush_lin()
    # ---- signatures: Bp = (x, y, 1) affine sigma; A = [r] sigma, A2 = [|x|] sigma.
    # in G2  <=>  psi(sigma) == [x] sigma = -A2:  e1 = psi_x Z2 - X2 = 0 and e2 = psi_y Z2 + Y2 = 0 (and Z2 != 0)
    pr.section("sigcheck")
    g2_psi(b, "Ps", "Bp"); b.flush_mul(); b.flush_lin()
    b.mul("t1", "PsX", "A2Z"); b.mul("t2", "PsY", "A2Z"); b.flush_mul()
    b.lin("e1", [(1, "t1"), (-1, "A2X")]); b.lin("e2", [(1, "t2"), (1, "A2Y")]); b.flush_lin()
    # ---- hash_to_curve: clear_cofactor (Budroni-Pintore, RFC 9380 G.3), as g2_clear_cofactor in ec.cuh
    # input H Jacobian -> P0 projective (X Z, Y, Z^3)
    pr.section("hinit")
    b.mul("P0X", "HX", "HZ"); b.sqr("zz", "HZ"); b.flush_mul()
    b.mul("P0Z", "zz", "HZ"); b.flush_mul()
    b.lin("P0Y", [(1, "HY")]); b.flush_lin()
    g2_copy(b, "Bp", "P0"); g2_copy(b, "A", "P0"); b.flush_lin()
    # sums of Jacobian points (the sum r_i sig_i tree): Bp <- projective form of the Jacobian point in H, A untouched
    pr.section("hbp")
    b.mul("BpX", "HX", "HZ"); b.sqr("zz", "HZ"); b.flush_mul()
    b.mul("BpZ", "zz", "HZ"); b.flush_mul()
    b.lin("BpY", [(1, "HY")]); b.flush_lin()
    # after ladder 1 (A = [|x|] P0):  T1 = -A = [x] P0;  T2 = psi(P0);  A = P0 (to be doubled)
    pr.section("hmid1")
    g2_copy(b, "T1", "A", neg=True); g2_psi(b, "T2", "P0"); b.flush_mul(); b.flush_lin()
    g2_copy(b, "A", "P0"); b.flush_lin()
    # (dblA) then A = psi^2(2 P0) = (X c2x : -Y : Z);  Bp = -T2
    pr.section("hmid2")
    b.mul_fp("AX", "AX", "c2x"); b.flush_mul()
    b.lin("AY", [(-1, "AY")]); g2_copy(b, "Bp", "T2", neg=True); b.flush_lin()
    # (addA: A = t3 = psi^2(2P) - psi(P))  T3 = A;  A = T2;  Bp = T1
    pr.section("hmid3")
    g2_copy(b, "T3", "A"); g2_copy(b, "Bp", "T1"); b.flush_lin()
    g2_copy(b, "A", "T2"); b.flush_lin()
    # (addA: A = t1 + t2)  Bp = A  (base of ladder 2; A stays the accumulator start)
    pr.section("hmid4")
    g2_copy(b, "Bp", "A"); b.flush_lin()
    # (ladder 2: A = [|x|](t1 + t2))  A = -A;  Bp = T3
    pr.section("hmid5")
    b.lin("AY", [(-1, "AY")]); g2_copy(b, "Bp", "T3"); b.flush_lin()
    # (addA)  Bp = -T1
    pr.section("hmid6")
    g2_copy(b, "Bp", "T1", neg=True); b.flush_lin()
    # (addA)  Bp = -P0
    pr.section("hmid7")
    g2_copy(b, "Bp", "P0", neg=True); b.flush_lin()
    # (addA) (tojac)
    return pr


def g2_set_point(pr, mem, name, pt, z=None):
    """store an affine oracle point projectively (x z : y z : z)"""
    z = z or (1, 0)
    for c, v in (("X", B.f2_mul(pt[0], z)), ("Y", B.f2_mul(pt[1], z)), ("Z", z)):
        mem[pr.slots[f"{name}{c}.0"]], mem[pr.slots[f"{name}{c}.1"]] = v


def g2_get_affine(pr, mem, name):
    X, Y, Z = [(mem[pr.slots[f"{name}{c}.0"]], mem[pr.slots[f"{name}{c}.1"]]) for c in "XYZ"]
    if Z == (0, 0):
        return None
    zi = B.f2_inv(Z)
    return (B.f2_mul(X, zi), B.f2_mul(Y, zi))


def g2_get_jac_affine(pr, mem):
    X, Y, Z = [(mem[pr.slots[f"J{c}.0"]], mem[pr.slots[f"J{c}.1"]]) for c in "XYZ"]
    zi = B.f2_inv(Z); zi2 = B.f2_sqr(zi)
    return (B.f2_mul(X, zi2), B.f2_mul(Y, B.f2_mul(zi2, zi)))


def g2_consts(pr, mem):
    mem[pr.slots["CX.0"]], mem[pr.slots["CX.1"]] = B.PSI_CX
    mem[pr.slots["CY.0"]], mem[pr.slots["CY.1"]] = B.PSI_CY
    cx2 = B.f2_mul(B.f2_conj(B.PSI_CX), B.PSI_CX)
    assert cx2[1] == 0
    mem[pr.slots["c2x"]] = cx2[0]


def ladder(pr, mem, k, two=False, k2=0):
    """A = [k] Bp (and A2 = [k2] Bp), left to right, exactly the control flow of the kernels"""
    started = started2 = False
    for i in range(63, -1, -1):
        if started or started2:
            run_section(pr, "dblAA2" if two else "dblA", mem)
        if (k >> i) & 1:
            run_section(pr, "addA" if started else "setA", mem); started = True
        if two and (k2 >> i) & 1:
            run_section(pr, "addA2" if started2 else "setA2", mem); started2 = True


def check_g2(pr):
    rnd = random.Random(77)
    mem = [0] * len(pr.slots)
    g2_consts(pr, mem)
    # ---- signature program
    sig = B.g2_mul(B.G2_GEN, rnd.randrange(1, B.R))
    r = rnd.randrange(1, 1 << 64)
    g2_set_point(pr, mem, "Bp", sig)
    ladder(pr, mem, r, True, B.X_ABS)
    assert g2_get_affine(pr, mem, "A") == B.g2_mul(sig, r)
    assert g2_get_affine(pr, mem, "A2") == B.g2_mul(sig, B.X_ABS)
    run_section(pr, "sigcheck", mem)
    assert all(mem[pr.slots[n]] == 0 for n in ("e1.0", "e1.1", "e2.0", "e2.1")), "a G2 point failed the subgroup check"
    run_section(pr, "tojac", mem)
    assert g2_get_jac_affine(pr, mem) == B.g2_mul(sig, r)
    # a curve point outside G2 must fail it
    x = (5, 1)
    while True:
        y = B.f2_sqrt(B.f2_add(B.f2_mul(B.f2_sqr(x), x), B.B2))
        if y and not B.g2_in_subgroup((x, y)):
            break
        x = (x[0] + 1, 1)
    g2_set_point(pr, mem, "Bp", (x, y))
    ladder(pr, mem, 3, True, B.X_ABS)
    run_section(pr, "sigcheck", mem)
    assert any(mem[pr.slots[n]] != 0 for n in ("e1.0", "e1.1", "e2.0", "e2.1"))
    # ---- clear_cofactor program on a curve point outside G2 given in Jacobian form
    z = (rnd.randrange(1, P), rnd.randrange(1, P)); z2 = B.f2_sqr(z)
    Hj = (B.f2_mul(x, z2), B.f2_mul(y, B.f2_mul(z2, z)), z)
    for n, v in zip(("HX", "HY", "HZ"), Hj):
        mem[pr.slots[n + ".0"]], mem[pr.slots[n + ".1"]] = v
    run_section(pr, "hinit", mem)
    ladder(pr, mem, B.X_ABS)
    run_section(pr, "hmid1", mem); run_section(pr, "dblA", mem); run_section(pr, "hmid2", mem)
    run_section(pr, "addA", mem); run_section(pr, "hmid3", mem); run_section(pr, "addA", mem)
    run_section(pr, "hmid4", mem)
    ladder(pr, mem, B.X_ABS)
    run_section(pr, "hmid5", mem); run_section(pr, "addA", mem); run_section(pr, "hmid6", mem)
    run_section(pr, "addA", mem); run_section(pr, "hmid7", mem); run_section(pr, "addA", mem)
    run_section(pr, "tojac", mem)
    assert g2_get_jac_affine(pr, mem) == B.g2_mul((x, y), B.H_EFF), "clear_cofactor program disagrees with [h_eff]P"
    # ---- sum of Jacobian points: hinit (first), then hbp + addA per further point
    pts = [B.g2_mul(B.G2_GEN, rnd.randrange(1, B.R)) for _ in range(3)]
    acc = None
    for k, pt in enumerate(pts):
        z = (rnd.randrange(1, P), rnd.randrange(1, P)); z2 = B.f2_sqr(z)
        for n, v in zip(("HX", "HY", "HZ"), (B.f2_mul(pt[0], z2), B.f2_mul(pt[1], B.f2_mul(z2, z)), z)):
            mem[pr.slots[n + ".0"]], mem[pr.slots[n + ".1"]] = v
        if k == 0:
            run_section(pr, "hinit", mem)
        else:
            run_section(pr, "hbp", mem); run_section(pr, "addA", mem)
        acc = B.g2_add(acc, pt)
    run_section(pr, "tojac", mem)
    assert g2_get_jac_affine(pr, mem) == acc, "point sum program disagrees with g2_add"



# ================================================================================================ final exponentiation
FE_REGS = ("f", "t0", "t1", "t2")
# the sequence of coop_final_exp (bls/coop.cuh) = final_exp (bls/pairing.cuh): f^(3 (p^12 - 1) / r)
FE_MULS = [("f", "t0", "t1"), ("f", "t0", "f"), ("t0", "t0", "t1"), ("t0", "t1", "t2"), ("t2", "t2", "t1"), ("t1", "t1", "f"),
           ("f", "t2", "t1"), ("t0", "t0", "f"), ("t1", "t1", "t0"), ("t2", "t2", "t1"), ("f", "f", "t0")]
FE_COPIES = [("t0", "f"), ("t1", "t0"), ("t2", "t1"), ("t1", "f")]
FE_CONJS = [("t0", "t0"), ("t1", "t1"), ("t2", "t2"), ("t0", "f"), ("t1", "f"), ("t2", "t0"), ("t1", "t0")]
FE_FROB2 = [("t0", "f"), ("t1", "t0")]
FE_FROB = [("t2", "t0")]


def build_fe():
    pr = Prog()
    for V in FE_REGS:
        for k in range(6):
            pr.slot(f"{V}{k}.0"); pr.slot(f"{V}{k}.1")
    for k in range(6):
        pr.slot(f"g1_{k}.0"); pr.slot(f"g1_{k}.1")
    for k in range(6):
        pr.slot(f"g2_{k}")
    pr.slot("dummy")

    def c(V, k, comp):
        return pr.slot(f"{V}{k}.{comp}")

    done = set()
    for dst, a, bb in FE_MULS:
        name = f"mul_{dst}_{a}_{bb}"
        if name in done:
            continue
        done.add(name)
        pr.section(name)
        lops = []
        for j in range(6):
            lops.append((pr.slot(f"ys{j}"), [(1, c(bb, j, 0)), (1, c(bb, j, 1))], 0))
            lops.append((pr.slot(f"yd{j}"), [(1, c(bb, j, 0)), (-1, c(bb, j, 1))], 0))
        pr.lin_phase(lops)
        ops, parts = [], {}
        for t in range(6):
            re, im = [], []
            for i in range(6):
                for j in range(6):
                    if i + j == t:
                        re += [(X_POS, c(a, i, 0), c(bb, j, 0)), (X_NEG, c(a, i, 1), c(bb, j, 1))]
                        im += [(X_POS, c(a, i, 0), c(bb, j, 1)), (X_POS, c(a, i, 1), c(bb, j, 0))]
                    elif i + j == t + 6:
                        re += [(X_POS, c(a, i, 0), pr.slot(f"yd{j}")), (X_NEG, c(a, i, 1), pr.slot(f"ys{j}"))]
                        im += [(X_POS, c(a, i, 0), pr.slot(f"ys{j}")), (X_POS, c(a, i, 1), pr.slot(f"yd{j}"))]
            for comp, lst in ((0, re), (1, im)):
                assert len(lst) == 12
                parts[(t, comp)] = []
                for ci in range(3):
                    d = pr.slot(f"mp{t}.{comp}.{ci}")
                    parts[(t, comp)].append(d)
                    ops.append((d, lst[4 * ci:4 * ci + 4]))
        pr.mul_phase(KMAX, ops[:18]); pr.mul_phase(KMAX, ops[18:])
        pr.lin_phase([(c(dst, t, comp), [(1, x) for x in parts[(t, comp)]], 0) for t in range(6) for comp in (0, 1)])

    for V in ("t0", "t1", "t2"):    # cyclotomic squaring in place (Granger-Scott, pairs (a0,a3) (a1,a4) (a2,a5))
        pr.section(f"cyc_{V}")
        ops = []
        for pp in range(3):
            x0, x1, y0, y1 = c(V, pp, 0), c(V, pp, 1), c(V, pp + 3, 0), c(V, pp + 3, 1)
            ops.append((pr.slot(f"sx{pp}.0"), [(X_POS, x0, x0), (X_NEG, x1, x1)]))
            ops.append((pr.slot(f"sx{pp}.1"), [(X_DBL, x0, x1)]))
            ops.append((pr.slot(f"sy{pp}.0"), [(X_POS, y0, y0), (X_NEG, y1, y1)]))
            ops.append((pr.slot(f"sy{pp}.1"), [(X_DBL, y0, y1)]))
            ops.append((pr.slot(f"bb{pp}.0"), [(X_DBL, x0, y0), (X_NEGDBL, x1, y1)]))      # B = 2 x y
            ops.append((pr.slot(f"bb{pp}.1"), [(X_DBL, x0, y1), (X_DBL, x1, y0)]))
        pr.mul_phase(2, ops)
        S = pr.slot
        lops = []

        def A_minus(dst_k, pp):      # 3 (sx + xi sy) - 2 a
            lops.append((c(V, dst_k, 0), [(3, S(f"sx{pp}.0")), (3, S(f"sy{pp}.0")), (-3, S(f"sy{pp}.1")), (-2, c(V, dst_k, 0))], 0))
            lops.append((c(V, dst_k, 1), [(3, S(f"sx{pp}.1")), (3, S(f"sy{pp}.0")), (3, S(f"sy{pp}.1")), (-2, c(V, dst_k, 1))], 0))

        def B_plus(dst_k, pp, xi):   # 3 [xi] B + 2 a
            if not xi:
                for comp in (0, 1):
                    lops.append((c(V, dst_k, comp), [(3, S(f"bb{pp}.{comp}")), (2, c(V, dst_k, comp))], 0))
            else:
                lops.append((c(V, dst_k, 0), [(3, S(f"bb{pp}.0")), (-3, S(f"bb{pp}.1")), (2, c(V, dst_k, 0))], 0))
                lops.append((c(V, dst_k, 1), [(3, S(f"bb{pp}.0")), (3, S(f"bb{pp}.1")), (2, c(V, dst_k, 1))], 0))
        A_minus(0, 0); B_plus(3, 0, False); B_plus(1, 2, True); A_minus(4, 2); A_minus(2, 1); B_plus(5, 1, False)
        pr.lin_phase(lops)

    for dst, src in FE_COPIES:
        pr.section(f"copy_{dst}_{src}")
        pr.lin_phase([(c(dst, k, comp), [(1, c(src, k, comp))], 0) for k in range(6) for comp in (0, 1)])
    for dst, src in FE_CONJS:     # a^(p^6): negate the odd powers of w
        pr.section(f"conj_{dst}_{src}")
        pr.lin_phase([(c(dst, k, comp), [(-1 if k & 1 else 1, c(src, k, comp))], 0) for k in range(6) for comp in (0, 1)])
    for dst, src in FE_FROB2:     # a^(p^2): coefficient k times the Fp constant gamma2_k
        pr.section(f"frob2_{dst}_{src}")
        pr.mul_phase(2, [(c(dst, k, comp), [(X_POS, pr.slot(f"g2_{k}"), c(src, k, comp))]) for k in range(6) for comp in (0, 1)])
    for dst, src in FE_FROB:      # a^p: conj(a_k) * gamma_k
        pr.section(f"frob_{dst}_{src}")
        ops = []
        for k in range(6):
            s0, s1, g0, g1 = c(src, k, 0), c(src, k, 1), pr.slot(f"g1_{k}.0"), pr.slot(f"g1_{k}.1")
            ops.append((c(dst, k, 0), [(X_POS, s0, g0), (X_POS, s1, g1)]))       # (s0 - s1 i)(g0 + g1 i)
            ops.append((c(dst, k, 1), [(X_POS, s0, g1), (X_NEG, s1, g0)]))
        pr.mul_phase(2, ops)
    return pr


def fe_get(pr, mem, V):
    k = [(mem[pr.slots[f"{V}{j}.0"]], mem[pr.slots[f"{V}{j}.1"]]) for j in range(6)]
    return ((k[0], k[2], k[4]), (k[1], k[3], k[5]))


def fe_put(pr, mem, V, v):
    k = [v[0][0], v[1][0], v[0][1], v[1][1], v[0][2], v[1][2]]
    for j in range(6):
        mem[pr.slots[f"{V}{j}.0"]], mem[pr.slots[f"{V}{j}.1"]] = k[j]


def fe_program(run, inv):
    """the control flow of k_final_warp: run(section name), inv() = t1 <- f^-1 (single-lane code on the device)"""
    def pow_x(r, a):
        run(f"copy_{r}_{a}")
        for i in range(62, -1, -1):
            run(f"cyc_{r}")
            if (B.X_ABS >> i) & 1:
                run(f"mul_{r}_{r}_{a}")
        run(f"conj_{r}_{r}")
    inv()
    run("conj_t0_f"); run("mul_f_t0_t1"); run("frob2_t0_f"); run("mul_f_t0_f")
    pow_x("t0", "f"); run("conj_t1_f"); run("mul_t0_t0_t1")
    pow_x("t1", "t0"); run("conj_t2_t0"); run("mul_t0_t1_t2")
    pow_x("t1", "t0"); run("frob_t2_t0"); run("mul_t0_t1_t2")
    pow_x("t1", "t0"); pow_x("t2", "t1"); run("frob2_t1_t0"); run("mul_t2_t2_t1"); run("conj_t1_t0"); run("mul_t2_t2_t1")
    run("copy_t1_f"); run("cyc_t1"); run("mul_t1_t1_f"); run("mul_f_t2_t1")


def check_fe(pr):
    rnd = random.Random(5)
    mem = [0] * len(pr.slots)
    G = [B.f2_pow(B.XI, k * (P - 1) // 6) for k in range(6)]
    for k in range(6):
        mem[pr.slots[f"g1_{k}.0"]], mem[pr.slots[f"g1_{k}.1"]] = G[k]
        n = B.f2_mul(B.f2_conj(G[k]), G[k])
        assert n[1] == 0
        mem[pr.slots[f"g2_{k}"]] = n[0]
    f_in = B.miller_loop(B.g1_mul(B.G1_GEN, rnd.randrange(1, B.R)), B.g2_mul(B.G2_GEN, rnd.randrange(1, B.R)))
    fe_put(pr, mem, "f", f_in)
    # product section used by the kernel to fold its inputs: f <- f * t0
    other = B.f12_pow(f_in, 7)
    fe_put(pr, mem, "t0", other)
    run_section(pr, "mul_f_f_t0", mem)
    assert fe_get(pr, mem, "f") == B.f12_mul(f_in, other)
    fe_put(pr, mem, "f", f_in)
    fe_program(lambda name: run_section(pr, name, mem), lambda: fe_put(pr, mem, "t1", B.f12_inv(fe_get(pr, mem, "f"))))
    g = B.final_exp(f_in)
    assert fe_get(pr, mem, "f") == B.f12_mul(B.f12_sqr(g), g), "final exponentiation program != oracle GT value cubed"


# ------------------------------------------------------------------------------------------------ emit
def emit(pr, path, PX="MW", exported=("f0.0", "X.0", "Y.0", "Z.0", "HX.0", "HY.0", "HZ.0", "px", "py", "pz", "g0.0", "dummy"),
         structs=True):
    out = ["// generated by scripts/gen_miller_warp.py — do not edit (phase tables of bls/miller_warp.cuh)"]
    out.append(f"constexpr int {PX}_NSLOTS = {len(pr.slots)};")
    for n in exported:
        out.append(f"constexpr int {PX}_S_{n.replace('.', '_').upper()} = {pr.slots[n]};")
    if PX == "MW":   # every coefficient block is (a0, a1, s, d) contiguous: f_j at MW_S_F0_0 + 4 j, g likewise
        for j in range(6):
            assert [pr.slots[f"f{j}.{c}"] for c in ("0", "1", "s", "d")] == [pr.slots["f0.0"] + 4 * j + k for k in range(4)]
            assert [pr.slots[f"g{j}.{c}"] for c in ("0", "1", "s", "d")] == [pr.slots["g0.0"] + 4 * j + k for k in range(4)]
    mul_rows, lin_rows, phases, sect = [], [], [], []
    dummy = pr.slots["dummy"]
    for name, phs in pr.sections.items():
        first = len(phases)
        for kind, K, ops in phs:
            if kind == "mul":
                base = len(mul_rows)
                for lane in range(NLANES):
                    if lane < len(ops):
                        d, terms = ops[lane]
                        terms = list(terms) + [(X_ZERO, dummy, dummy)] * (KMAX - len(terms))
                    else:
                        d, terms = dummy, [(X_ZERO, dummy, dummy)] * KMAX
                    mul_rows.append("{%d, {%s}, {%s}, {%s}, {0,0,0}}" % (d, ",".join(str(t[0]) for t in terms),
                                                               ",".join(str(t[1]) for t in terms),
                                                               ",".join(str(t[2]) for t in terms)))
                phases.append("{1, %d, %d}" % (K, base // NLANES))
            else:
                base = len(lin_rows)
                for lane in range(NLANES):
                    if lane < len(ops):
                        d, terms, h = ops[lane]
                    else:
                        d, terms, h = dummy, [(1, dummy)], 0
                    n = len(terms)
                    terms = list(terms) + [(0, dummy)] * (4 - n)
                    lin_rows.append("{%d, %d, %d, {%s}, {%s}, 0}" % (d, n, h, ",".join(str(t[0]) for t in terms),
                                                                 ",".join(str(t[1]) for t in terms)))
                phases.append("{0, 0, %d}" % (base // NLANES))
        sect.append((name, first, len(phases) - first))
    if structs:
        out.append("struct MwMulOp { uint8_t d; uint8_t xm[4]; uint8_t xs[4]; uint8_t ys[4]; uint8_t pad[3]; };   // 16 B")
        out.append("struct MwLinOp { uint8_t d; uint8_t n; uint8_t h; int8_t c[4]; uint8_t s[4]; uint8_t pad; };   // 12 B")
        out.append("struct MwPhase { uint8_t is_mul; uint8_t k; uint16_t table; };")
    out.append(f"MW_TABLE MwMulOp {PX}_MUL[{len(mul_rows)}] = {{\n" + ",\n".join(mul_rows) + "};")
    out.append(f"MW_TABLE MwLinOp {PX}_LIN[{len(lin_rows)}] = {{\n" + ",\n".join(lin_rows) + "};")
    out.append(f"MW_TABLE MwPhase {PX}_PHASES[{len(phases)}] = {{" + ", ".join(phases) + "};")
    out.append(f"constexpr int {PX}_N_MUL = {len(mul_rows)}, {PX}_N_LIN = {len(lin_rows)}, {PX}_N_PHASES = {len(phases)};")
    for name, first, cnt in sect:
        out.append(f"constexpr int {PX}_SEC_{name.upper()}_FIRST = {first}, {PX}_SEC_{name.upper()}_COUNT = {cnt};")
    assert len(pr.slots) < 256
    open(path, "w").write("\n".join(out) + "\n")
    return len(mul_rows) // NLANES, len(lin_rows) // NLANES


if __name__ == "__main__":
    pr = build()
    check(pr)
    nm, nl = emit(pr, os.path.join(ROOT, "lighthouse_b200", "csrc", "bls", "miller_warp_tables.inc"))
    print(f"ok: {len(pr.slots)} slots, {nm} mul phases, {nl} lin phases;",
          {n: len(p) for n, p in pr.sections.items()})
    g2 = build_g2()
    check_g2(g2)
    nm, nl = emit(g2, os.path.join(ROOT, "lighthouse_b200", "csrc", "bls", "g2_warp_tables.inc"), PX="GW",
                  exported=("AX.0", "A2X.0", "BpX.0", "HX.0", "CX.0", "CY.0", "JX.0", "c2x", "e1.0", "e2.0", "dummy"),
                  structs=False)
    print(f"ok: g2 programs, {len(g2.slots)} slots, {nm} mul phases, {nl} lin phases;",
          {n: len(p) for n, p in g2.sections.items()})
    fe = build_fe()
    check_fe(fe)
    nm, nl = emit(fe, os.path.join(ROOT, "lighthouse_b200", "csrc", "bls", "fe_warp_tables.inc"), PX="FE",
                  exported=("f0.0", "t0" + "0.0", "t10.0", "t20.0", "g1_0.0", "g2_0", "dummy"), structs=False)
    print(f"ok: final exponentiation, {len(fe.slots)} slots, {nm} mul phases, {nl} lin phases, {len(fe.sections)} sections")
