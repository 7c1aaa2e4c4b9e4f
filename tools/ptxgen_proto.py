#!/usr/bin/env python
"""Prototype of the K-spec v2 code generator (BITS descriptors, PTX output).

Used to compare code shapes on the B200 before the winner is ported into
bayescard_b200/csrc/spec_codegen.cc:

  S  one query per thread; u = selp(lambda, 0, bit); m_p = fma(u, T, m_p)
  P  one query per thread; @bit fma(m_p, lambda, T, m_p)   (predicated, no select)
  F  two queries per thread packed in f32x2 pairs; fma.rn.f32x2 with the CPT entry broadcast

  python tools/ptxgen_proto.py census F out.ptx [--threads 128 --minblocks 2]
"""
from __future__ import annotations

import argparse
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def fhex(x: float) -> str:
    return "0f%08X" % struct.unpack("<I", struct.pack("<f", float(x)))[0]


def f2hex(x: float) -> str:
    b = struct.unpack("<I", struct.pack("<f", float(x)))[0]
    return "0x%08X%08X" % (b, b)


class Gen:
    def __init__(self, tm, variant: str, threads: int, minblocks: int):
        self.tm = tm
        self.v = variant
        self.threads = threads
        self.minblocks = minblocks
        self.n = tm.n_nodes
        self.card = [int(c) for c in tm.card]
        self.parent = [int(p) for p in tm.parent]
        self.kids = [[] for _ in range(self.n)]
        for v in range(1, self.n):
            self.kids[self.parent[v]].append(v)
        sub = [0] * self.n
        for v in range(self.n - 1, -1, -1):
            sub[v] += self.card[v]
            if v:
                sub[self.parent[v]] += sub[v]
        for v in range(self.n):
            self.kids[v].sort(key=lambda k: -sub[k])
        self.T = [np.asarray(t, dtype=np.float64).astype(np.float32) for t in tm.cpts]
        self.bit_off = np.concatenate([[0], np.cumsum(self.card)[:-1]]).astype(int)
        self.words = -(-int(sum(self.card)) // 128) * 4
        self.fan = [tm.fan_vector(v) for v in range(self.n)]
        self.mask_words = (self.n + 31) // 32
        self.L = []
        self.nf = self.np_ = self.nr = self.nd = 0
        self.n_fma = 0
        self.Q = 2 if variant == "F" else 1

    # ---- registers
    def f(self):
        self.nf += 1
        return f"%f{self.nf}"

    def p(self):
        self.np_ += 1
        return f"%p{self.np_}"

    def r(self):
        self.nr += 1
        return f"%r{self.nr}"

    def d(self):
        self.nd += 1
        return f"%rd{self.nd}"

    def emit(self, s):
        self.L.append("    " + s)

    # ---- bit predicate of element (v, c) for query slot s
    def bitpred(self, s, v, c):
        b = int(self.bit_off[v]) + c
        w, k = b >> 5, b & 31
        t, p = self.r(), self.p()
        self.emit(f"and.b32 {t}, {self.words_reg[s][w]}, {1 << k};")
        self.emit(f"setp.ne.u32 {p}, {t}, 0;")
        return p

    # ---- message of node v: list of registers (one per parent state); F: .b64 pairs
    def message(self, v):
        card, root = self.card[v], v == 0
        self.emit(f"// ---- node {v}: card {card}" + (" (root)" if root else f", parent {self.parent[v]}"))
        lam = None
        for k in self.kids[v]:
            mk = self.message(k)
            if lam is None:
                lam = mk
            else:
                for c in range(card):
                    if self.v == "F":
                        self.emit(f"mul.f32x2 {lam[c]}, {lam[c]}, {mk[c]};")
                    else:
                        self.emit(f"mul.f32 {lam[c]}, {lam[c]}, {mk[c]};")
        fan = self.fan[v]
        cols = 1 if root else self.card[self.parent[v]]
        T = self.T[v].reshape(card, cols) if not root else self.T[v].reshape(card, 1)
        if self.v == "F":
            return self._message_F(v, lam, fan, T, cols)
        # per-slot fan predicate
        pfb = None
        if fan is not None:
            t = self.r()
            pfb = self.p()
            self.emit(f"and.b32 {t}, {self.fm_reg[0][v >> 5]}, {1 << (v & 31)};")
            self.emit(f"setp.ne.u32 {pfb}, {t}, 0;")
        m = [None] * cols
        for c in range(card):
            nz = [p for p in range(cols) if T[c, p] != 0.0]
            if not nz and not (self.v == "H" and c == 0):
                continue
            pc = self.bitpred(0, v, c)
            lc = lam[c] if lam is not None else None
            if fan is not None:
                if lc is None:
                    lc = self.f()
                    self.emit(f"selp.f32 {lc}, {fhex(fan[c])}, 0f3F800000, {pfb};")
                else:
                    self.emit(f"@{pfb} mul.f32 {lc}, {lc}, {fhex(fan[c])};")
            if self.v == "S":
                u = self.f()
                self.emit(f"selp.f32 {u}, {lc if lc else '0f3F800000'}, 0f00000000, {pc};")
                for p in nz:
                    self.n_fma += 1
                    if m[p] is None:
                        m[p] = self.f()
                        self.emit(f"mul.f32 {m[p]}, {u}, {fhex(T[c, p])};")
                    else:
                        self.emit(f"fma.rn.f32 {m[p]}, {u}, {fhex(T[c, p])}, {m[p]};")
            elif self.v == "H" and c == 0:
                # first row seeds every accumulator through one select (no zero-init, no predicates)
                u = self.f()
                self.emit(f"selp.f32 {u}, {lc if lc else '0f3F800000'}, 0f00000000, {pc};")
                for p in range(cols):
                    m[p] = self.f()
                    if T[c, p] != 0.0:
                        self.n_fma += 1
                        self.emit(f"mul.f32 {m[p]}, {u}, {fhex(T[c, p])};")
                    else:
                        self.emit(f"mov.f32 {m[p]}, 0f00000000;")
            else:  # P, and rows >= 1 of H
                for p in nz:
                    self.n_fma += 1
                    if m[p] is None:
                        m[p] = self.f()
                        self.emit(f"mov.f32 {m[p]}, 0f00000000;")
                    if lc is None:
                        self.emit(f"@{pc} add.f32 {m[p]}, {m[p]}, {fhex(T[c, p])};")
                    else:
                        self.emit(f"@{pc} fma.rn.f32 {m[p]}, {lc}, {fhex(T[c, p])}, {m[p]};")
        for p in range(cols):
            if m[p] is None:
                m[p] = self.f()
                self.emit(f"mov.f32 {m[p]}, 0f00000000;")
        return m

    def _message_F(self, v, lam, fan, T, cols):
        card = self.card[v]
        pfb = [None, None]
        if fan is not None:
            for s in range(2):
                t = self.r()
                pfb[s] = self.p()
                self.emit(f"and.b32 {t}, {self.fm_reg[s][v >> 5]}, {1 << (v & 31)};")
                self.emit(f"setp.ne.u32 {pfb[s]}, {t}, 0;")
        m = [None] * cols
        for c in range(card):
            nz = [p for p in range(cols) if T[c, p] != 0.0]
            if not nz:
                continue
            halves = []
            if lam is not None:
                la, lb = self.f(), self.f()
                self.emit(f"mov.b64 {{{la}, {lb}}}, {lam[c]};")
                src = [la, lb]
            else:
                src = [None, None]
            for s in range(2):
                pc = self.bitpred(s, v, c)
                ls = src[s]
                if fan is not None:
                    if ls is None:
                        ls = self.f()
                        self.emit(f"selp.f32 {ls}, {fhex(fan[c])}, 0f3F800000, {pfb[s]};")
                    else:
                        self.emit(f"@{pfb[s]} mul.f32 {ls}, {ls}, {fhex(fan[c])};")
                u = self.f()
                self.emit(f"selp.f32 {u}, {ls if ls else '0f3F800000'}, 0f00000000, {pc};")
                halves.append(u)
            U = self.d()
            self.emit(f"mov.b64 {U}, {{{halves[0]}, {halves[1]}}};")
            for p in nz:
                self.n_fma += 1
                tt = self.d()
                self.emit(f"mov.b64 {tt}, {f2hex(T[c, p])};")
                if m[p] is None:
                    m[p] = self.d()
                    self.emit(f"mul.f32x2 {m[p]}, {U}, {tt};")
                else:
                    self.emit(f"fma.rn.f32x2 {m[p]}, {U}, {tt}, {m[p]};")
        for p in range(cols):
            if m[p] is None:
                m[p] = self.d()
                self.emit(f"mov.b64 {m[p]}, 0;")
        return m

    def generate(self):
        Q = self.Q
        self.words_reg = [[self.r() for _ in range(self.words)] for _ in range(Q)]
        any_fan = any(f is not None for f in self.fan)
        self.fm_reg = [[self.r() for _ in range(self.mask_words)] for _ in range(Q)]
        body_start = len(self.L)
        res = self.message(0)[0]
        body = self.L[body_start:]
        self.L = self.L[:body_start]
        H = []
        H.append("//\n// Generated by bayescard_b200 K-spec v2 prototype (variant %s) -- do not edit.\n//" % self.v)
        H.append(f"// nodes: {self.n}  FMA emitted per query: {self.n_fma}  BC_SPEC_FFMA={self.n_fma}")
        H.append(".version 8.7\n.target sm_100a\n.address_size 64\n")
        H.append(f".visible .global .align 4 .u32 bc_spec_meta[4] = {{{self.threads}, {Q}, 8, 0}};\n")
        H.append(".visible .entry bc_spec_bits(\n    .param .u64 p_desc,\n    .param .u64 p_stride,\n    .param .u64 p_fmask,\n"
                 "    .param .u64 p_out,\n    .param .u64 p_nq\n)\n"
                 f".maxntid {self.threads}, 1, 1\n.minnctapersm {self.minblocks}\n{{")
        decl_at = len(H)
        P = []

        def e(s):
            P.append("    " + s)

        e("ld.param.u64 %rdesc, [p_desc];")
        e("ld.param.u64 %rstride, [p_stride];")
        e("ld.param.u64 %rfmask, [p_fmask];")
        e("ld.param.u64 %rout, [p_out];")
        e("ld.param.u64 %rnq, [p_nq];")
        e("cvta.to.global.u64 %rdesc, %rdesc;")
        e("cvta.to.global.u64 %rout, %rout;")
        e("setp.ne.u64 %pfm, %rfmask, 0;")
        e("@%pfm cvta.to.global.u64 %rfmask, %rfmask;")
        e("mov.u32 %t0, %tid.x;")
        e("mov.u32 %t1, %ctaid.x;")
        e("mov.u32 %t2, %ntid.x;")
        e("mov.u32 %t3, %nctaid.x;")
        e("mad.wide.u32 %rq, %t1, %t2, 0;")
        e("cvt.u64.u32 %rtmp, %t0;")
        e("add.u64 %rq, %rq, %rtmp;")
        e("mul.wide.u32 %rstep, %t2, %t3;")
        if Q == 2:
            e("shl.b64 %rq, %rq, 1;")
            e("shl.b64 %rstep, %rstep, 1;")
        e("LOOP:")
        e("setp.ge.u64 %pdone, %rq, %rnq;")
        e("@%pdone bra DONE;")
        for s in range(Q):
            if s == 0:
                e("mov.u64 %rqs0, %rq;")
            else:
                e("add.u64 %rqs1, %rq, 1;")
                e("setp.lt.u64 %phasb, %rqs1, %rnq;")
                e("@!%phasb mov.u64 %rqs1, %rq;")
            e(f"mad.lo.u64 %rrow{s}, %rqs{s}, %rstride, %rdesc;")
            for w in range(0, self.words, 4):
                regs = ", ".join(self.words_reg[s][w:w + 4])
                e(f"ld.global.nc.v4.u32 {{{regs}}}, [%rrow{s}+{4 * w}];")
            if any_fan:
                for w in range(self.mask_words):
                    e(f"mov.u32 {self.fm_reg[s][w]}, 0;")
                    e(f"mad.lo.u64 %rtmp, %rqs{s}, {4 * self.mask_words}, %rfmask;")
                    e(f"@%pfm ld.global.nc.u32 {self.fm_reg[s][w]}, [%rtmp+{4 * w}];")
        P.extend(body)
        if Q == 1:
            e("shl.b64 %rtmp, %rq, 2;")
            e("add.u64 %rtmp, %rtmp, %rout;")
            e(f"st.global.f32 [%rtmp], {res};")
        else:
            e(f"mov.b64 {{%fra, %frb}}, {res};")
            e("shl.b64 %rtmp, %rq, 2;")
            e("add.u64 %rtmp, %rtmp, %rout;")
            e("st.global.f32 [%rtmp], %fra;")
            e("@%phasb st.global.f32 [%rtmp+4], %frb;")
        e("add.u64 %rq, %rq, %rstep;")
        e("bra LOOP;")
        e("DONE:")
        e("ret;")
        D = []
        D.append(f"    .reg .pred %p<{self.np_ + 1}>;")
        D.append("    .reg .pred %pfm, %pdone, %phasb;")
        D.append(f"    .reg .f32 %f<{self.nf + 1}>;")
        D.append("    .reg .f32 %fra, %frb;")
        D.append(f"    .reg .b32 %r<{self.nr + 1}>;")
        D.append("    .reg .b32 %t0, %t1, %t2, %t3;")
        D.append(f"    .reg .b64 %rd<{self.nd + 1}>;")
        D.append("    .reg .b64 %rdesc, %rstride, %rfmask, %rout, %rnq, %rq, %rstep, %rtmp, %rqs0, %rqs1, %rrow0, %rrow1;")
        return "\n".join(H[:decl_at] + D + P) + "\n}\n"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("model")
    ap.add_argument("variant", choices=["S", "P", "F", "H"])
    ap.add_argument("out")
    ap.add_argument("--threads", type=int, default=128)
    ap.add_argument("--minblocks", type=int, default=4)
    a = ap.parse_args()
    from bayescard_b200.loader import TreeModel

    tm = TreeModel.load(os.path.join(ROOT, "tests", "golden", "models", a.model + ".npz"))
    g = Gen(tm, a.variant, a.threads, a.minblocks)
    src = g.generate()
    with open(a.out, "w") as f:
        f.write(src)
    print(f"{a.out}: variant {a.variant}, {g.n_fma} FMA/query, {len(src.splitlines())} PTX lines")


if __name__ == "__main__":
    main()
