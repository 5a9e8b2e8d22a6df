"""Prints the per-vertex stage programs (tools/adgen/stages.py) as C++ over a scalar type T:
langevin-mcmc_b200/csrc/core/pathgrad_stages.inc (committed; `python tools/adgen/gen.py` regenerates it).

For every stage two functions are written:
  st_<name>_fwd<T>(scene, vert, lvert, lightType, in, out)                      forward value of the stage
  st_<name>_rev<T, COMPAT>(scene, vert, lvert, lightType, in, oadj, out, iadj)  forward + reverse sweep:
        iadj = J^T oadj with the statements in the order the reference's generator emits them;
        COMPAT = true merges `if` outputs by assignment like the reference (src/chad.cpp:284-287),
        COMPAT = false accumulates (the true adjoint).
T = float gives the gradient; T = a dual number gives forward-over-reverse second derivatives.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import chadlike as cl
import stages as st

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.join(ROOT, "langevin-mcmc_b200", "csrc", "core", "pathgrad_stages.inc")


def lit(v):
    s = "%.9g" % v
    if "e" not in s and "." not in s and "inf" not in s and "nan" not in s:
        s += ".0"
    return s + "f"


class Emitter:
    def __init__(self, prog):
        self.prog = prog
        self.lines = []
        self.ind = 1
        self.decl_t, self.decl_f = [], []
        self.declared = set()
        self.partner = {}      # sin node <-> cos node of the same argument recorded at the same nesting level

    def w(self, s):
        self.lines.append("    " * self.ind + s)

    def is_t(self, n):
        return n.diff

    def name(self, n):
        if n.op == "const":
            return lit(n.val)
        if n.op == "in":
            if n.diff:
                return n.name
            return "lightType" if n.name.startswith("lightType") else n.name
        nm = ("t%d" if n.diff else "f%d") % n.uid
        if n not in self.declared:
            self.declared.add(n)
            (self.decl_t if n.diff else self.decl_f).append(nm)
        return nm

    def as_t(self, n):
        s = self.name(n)
        return s if n.diff else "ad_const<T>(%s)" % s

    def expr(self, n):
        a = [self.name(x) for x in n.args]
        op = n.op
        if op == "add": return "%s + %s" % (a[0], a[1])
        if op == "sub": return "%s - %s" % (a[0], a[1])
        if op == "mul": return "%s * %s" % (a[0], a[1])
        if op == "div": return "%s / %s" % (a[0], a[1])
        if op == "neg": return "-%s" % a[0]
        if op == "sq": return "ad_square(%s)" % a[0]
        if op == "inv": return "ad_inverse(%s)" % a[0]
        if op in ("sin", "cos", "sqrt", "exp", "log", "acos"): return "ad_%s(%s)" % (op, a[0])
        if op == "pow":
            assert not n.args[1].diff
            return "ad_pow(%s, %s)" % (a[0], a[1])
        if op == "atan2":
            if n.args[0].diff != n.args[1].diff:
                return "ad_atan2(%s, %s)" % (self.as_t(n.args[0]), self.as_t(n.args[1]))
            return "ad_atan2(%s, %s)" % (a[0], a[1])
        if op == "dot3": return "%s * %s + %s * %s + %s * %s" % (a[0], a[3], a[1], a[4], a[2], a[5])
        if op == "len3": return "ad_sqrt(%s * %s + %s * %s + %s * %s)" % (a[0], a[0], a[1], a[1], a[2], a[2])
        raise KeyError(op)

    def cond(self, c):
        if c.kind == "and":
            return "(%s) && (%s)" % (self.cond(c.a), self.cond(c.b))
        sym = {"gt": ">", "ge": ">=", "lt": "<", "le": "<=", "eq": "=="}[c.kind]
        va = self.name(c.a) if c.a.op == "const" else "ad_val(%s)" % self.name(c.a)
        vb = self.name(c.b) if c.b.op == "const" else "ad_val(%s)" % self.name(c.b)
        return "%s %s %s" % (va, sym, vb)

    def forward(self, items):
        # sin(x) and cos(x) of the same x at the same nesting level: ONE evaluation of the sincos kernel, and the
        # reverse sweep reads the partner's value as the derivative instead of evaluating it again
        by_arg = {}
        for it in items:
            if it[0] == "def" and it[1].op in ("sin", "cos"):
                by_arg.setdefault(it[1].args[0], {})[it[1].op] = it[1]
        skip = set()
        for arg, d in by_arg.items():
            if len(d) == 2:
                self.partner[d["sin"]] = d["cos"]
                self.partner[d["cos"]] = d["sin"]
        for it in items:
            if it[0] == "def":
                n = it[1]
                if n in skip:
                    continue
                if n in self.partner:
                    m = self.partner[n]
                    s_, c_ = (n, m) if n.op == "sin" else (m, n)
                    self.w("ad_sincos(%s, %s, %s);" % (self.name(n.args[0]), self.name(s_), self.name(c_)))
                    skip.add(m)
                    continue
                self.w("%s = %s;" % (self.name(n), self.expr(n)))
            else:
                _, sp, branches = it
                for i, (cond, body, outs) in enumerate(branches):
                    if i == 0:
                        self.w("if (%s) {" % self.cond(cond))
                    elif cond is not None:
                        self.w("} else if (%s) {" % self.cond(cond))
                    else:
                        self.w("} else {")
                    self.ind += 1
                    self.forward(body)
                    for o, e in outs:
                        self.w("%s = %s;" % (self.name(o), self.as_t(e) if o.diff else self.name(e)))
                    self.ind -= 1
                self.w("}")

    def acc(self, n):
        nm = "a%d" % n.uid
        self.accs.add(nm)
        return nm

    def partial_stmt(self, c, e, k):
        a = [self.name(x) for x in e.args]
        r = self.name(e)
        ac, ae = self.acc(c), self.acc(e)
        op = e.op
        plus = lambda x: "%s = %s + %s;" % (ac, ac, x)
        minus = lambda x: "%s = %s - %s;" % (ac, ac, x)
        if op == "add": return plus(ae)
        if op == "sub": return plus(ae) if k == 0 else minus(ae)
        if op == "mul": return plus("%s * %s" % (ae, a[1 - k]))
        if op == "div":
            return plus("%s / %s" % (ae, a[1])) if k == 0 else minus("%s * (%s / %s)" % (ae, r, a[1]))
        if op == "neg": return minus(ae)
        if op == "sq": return plus("%s * (2.0f * %s)" % (ae, a[0]))
        if op == "inv": return minus("%s * (%s * %s)" % (ae, r, r))
        if op == "sin": return plus("%s * %s" % (ae, self.name(self.partner[e]) if e in self.partner else "ad_cos(%s)" % a[0]))
        if op == "cos": return minus("%s * %s" % (ae, self.name(self.partner[e]) if e in self.partner else "ad_sin(%s)" % a[0]))
        if op == "sqrt": return plus("%s * (0.5f / %s)" % (ae, r))
        if op == "exp": return plus("%s * %s" % (ae, r))
        if op == "log": return plus("%s / %s" % (ae, a[0]))
        if op == "acos": return minus("%s / ad_sqrt(1.0f - %s * %s)" % (ae, a[0], a[0]))
        if op == "pow": return plus("%s * (%s * %s / %s)" % (ae, a[1], r, a[0]))      # y x^(y-1) = y x^y / x
        if op == "atan2":
            den = "(%s * %s + %s * %s)" % (a[0], a[0], a[1], a[1])
            return plus("%s * (%s / %s)" % (ae, a[1], den)) if k == 0 else minus("%s * (%s / %s)" % (ae, a[0], den))
        if op == "dot3": return plus("%s * %s" % (ae, a[(k + 3) % 6]))
        if op == "len3": return plus("%s * (%s / %s)" % (ae, a[k], r))
        raise KeyError(op)

    def reverse(self, items):
        for it in items:
            if it[0] == "acc":
                _, c, e, k = it
                self.w(self.partial_stmt(c, e, k))
            else:
                _, sp, branches = it
                live = [(i, b) for i, b in enumerate(branches) if b[1] or b[2]]
                if not live:
                    continue
                first = True
                for i, (cond, assigns, body) in enumerate(branches):
                    if first:
                        self.w("if (%s) {" % self.cond(cond))
                        first = False
                    elif cond is not None:
                        self.w("} else if (%s) {" % self.cond(cond))
                    else:
                        self.w("} else {")
                    self.ind += 1
                    for e, o in assigns:
                        ae, ao = self.acc(e), self.acc(o)
                        self.w("if (COMPAT) %s = %s; else %s = %s + %s;" % (ae, ao, ae, ae, ao))
                    self.reverse(body)
                    self.ind -= 1
                self.w("}")


def emit_stage(name):
    prog = st.program(name)
    nin, nout = st.STAGES[name]
    terminal = nout == 1
    sig = "const float *scene, const float *vert, const float *lvert, float lightType, const T *in, const T *lp"
    text = []
    if not terminal:
        em = Emitter(prog)
        em.forward(prog.fwd)
        for i, o in enumerate(prog.outputs):
            em.w("out[%d] = %s;" % (i, em.as_t(o)))
        text.append("template <class T>\nLMC_HD_NOINLINE void st_%s_fwd(%s, T *out) {" % (name, sig))
        text += _decls(em)
        text += em.lines
        text.append("}")
    # forward + reverse
    em = Emitter(prog)
    em.accs = set()
    em.forward(prog.fwd)
    if terminal:
        em.w("out[0] = %s;" % em.as_t(prog.outputs[0]))
    fwd_lines = em.lines
    em.lines = []
    for i, o in enumerate(prog.outputs):
        if cl._has_acc(o):
            em.w("%s = %s + oadj[%d];" % (em.acc(o), em.acc(o), i))
    em.reverse(prog.rev)
    for key, n in prog.inputs.items():
        if not n.diff:
            continue
        arr, idx = key[:2], int(key[3:-1])
        dst = "iadj" if arr == "in" else "ladj"
        em.w("%s[%d] = %s;" % (dst, idx, em.acc(n) if n in prog.nz else "ad_const<T>(0.0f)"))
    text.append("template <class T, bool COMPAT>\nLMC_HD_NOINLINE void st_%s_rev(%s, const T *oadj, T *out, T *iadj, T *ladj) {" % (name, sig))
    text += _decls(em)
    text += fwd_lines
    accs = sorted(em.accs, key=lambda s: int(s[1:]))
    for i in range(0, len(accs), 12):
        text.append("    T " + ", ".join("%s = ad_const<T>(0.0f)" % a for a in accs[i:i + 12]) + ";")
    text += em.lines
    text.append("}")
    return "\n".join(text) + "\n"


def _decls(em):
    out = []
    for names, ty in ((em.decl_f, "float"), (em.decl_t, "T")):
        for i in range(0, len(names), 16):
            out.append("    %s %s;" % (ty, ", ".join(names[i:i + 16])))
    return out


def main():
    parts = ["// GENERATED by tools/adgen/gen.py from tools/adgen/{pathfn,stages}.py -- do not edit.\n"
             "// Per-vertex stages of the path function: forward value and reverse sweep (see pathgrad_rev.h).\n"]
    for name in st.STAGES:
        parts.append("// ---- stage %s: %d differentiable inputs, %d outputs\n" % ((name,) + st.STAGES[name]))
        parts.append(emit_stage(name))
    with open(OUT, "w") as f:
        f.write("\n".join(parts))
    print("wrote", OUT, sum(p.count("\n") for p in parts), "lines")


if __name__ == "__main__":
    main()
