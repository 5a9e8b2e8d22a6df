"""Expression tracer + reverse-mode program builder with the reference AD library's semantics.

The reference differentiates the path contribution with its own compile-time AD ("chad",
/root/reference/src/chad.{h,cpp}).  Its reverse sweep has a documented quirk that changes the numbers
(SURVEY.md App. B#13): when an `if` is merged, the adjoint of the value a branch forwards is ASSIGNED
(`_acc<value> = _acc<output>`, src/chad.cpp:284-287) after everything that follows the `if` has been
swept, so what the forwarded value had collected from later uses is dropped.  "Gradient parity with the
reference" means reproducing that (`compat`), the true adjoint (`exact`) is the same program with `+=`.

This module is OUR OWN tracer (nothing of chad is used or copied): it records a program with structured
control flow the way the reference's recorder does -- node identity, constant folding that aliases
operands (x*1 -> x, x+0 -> x, src/chad.h:1254-1340), blocks / splits in creation order
(src/chad.h:1444-1533) -- builds the forward and the reverse statement lists in the order
src/chad.cpp:109-170 / 230-331 emit them, and can (a) interpret them numerically (tests: whole-path
programs against the reference's compiled code) and (b) print them as C++ over a scalar type T
(product: per-vertex stage functions, tools/adgen/gen.py).
"""
import math

# ------------------------------------------------------------------------------------------------
# nodes
# ------------------------------------------------------------------------------------------------
_counter = [0]


class Node:
    __slots__ = ("op", "args", "val", "diff", "uid", "name")

    def __init__(self, op, args=(), val=None, diff=False, name=None):
        self.op = op
        self.args = list(args)
        self.val = val
        self.diff = diff
        self.name = name
        _counter[0] += 1
        self.uid = _counter[0]
        if op not in ("const", "in") and G.f is not None and G.f.cur is not None:
            G.f.cur.exprs.append(self)

    # arithmetic with the reference's folding rules -------------------------------------------
    def __add__(self, o): return add(self, lift(o))
    def __radd__(self, o): return add(lift(o), self)
    def __sub__(self, o): return sub(self, lift(o))
    def __rsub__(self, o): return sub(lift(o), self)
    def __mul__(self, o): return mul(self, lift(o))
    def __rmul__(self, o): return mul(lift(o), self)
    def __neg__(self): return neg(self)

    def __truediv__(self, o):
        if isinstance(o, Node):
            return _mk("div", self, o)          # Divide::Create, never folded (src/chad.h operator/)
        return mul(self, const(1.0 / o))        # expr / float -> expr * (1/float)

    def __rtruediv__(self, o):                  # float / expr -> float * inverse(expr)
        return mul(const(o), inverse(self))


class G:
    f = None


def const(v):
    return Node("const", val=float(v))


def lift(x):
    return x if isinstance(x, Node) else const(x)


def is_const(n, v=None):
    return n.op == "const" and (v is None or n.val == v)


def _mk(op, *args):
    return Node(op, args, diff=any(a.diff for a in args))


def inp(name, diff):
    return Node("in", name=name, diff=diff)


def neg(a):
    if is_const(a): return const(-a.val)
    return _mk("neg", a)


def add(a, b):
    if is_const(a, 0.0): return b
    if is_const(b, 0.0): return a
    if is_const(a) and is_const(b): return const(a.val + b.val)
    return _mk("add", a, b)


def sub(a, b):
    if is_const(a, 0.0): return neg(b)
    if is_const(b, 0.0): return a
    if is_const(a) and is_const(b): return const(a.val - b.val)
    return _mk("sub", a, b)


def mul(a, b):
    if is_const(a):
        if a.val == 0.0: return const(0.0)
        if a.val == 1.0: return b
        if a.val == -1.0: return neg(b)
    if is_const(b):
        if b.val == 0.0: return const(0.0)
        if b.val == 1.0: return a
        if b.val == -1.0: return neg(a)
    if is_const(a) and is_const(b): return const(a.val * b.val)
    return _mk("mul", a, b)


def _unary(op, fn):
    def f(a):
        a = lift(a)
        if is_const(a): return const(fn(a.val))
        return _mk(op, a)
    return f


square = _unary("sq", lambda x: x * x)
inverse = _unary("inv", lambda x: 1.0 / x)
sin = _unary("sin", math.sin)
cos = _unary("cos", math.cos)
sqrt = _unary("sqrt", math.sqrt)
exp = _unary("exp", math.exp)
log = _unary("log", math.log)
acos = _unary("acos", math.acos)


def pow_(a, b):
    a, b = lift(a), lift(b)
    if is_const(a) and is_const(b): return const(math.pow(a.val, b.val))
    return _mk("pow", a, b)


def atan2(y, x):
    y, x = lift(y), lift(x)
    if is_const(y) and is_const(x): return const(math.atan2(y.val, x.val))
    return _mk("atan2", y, x)


def dot3(a, b):
    return _mk("dot3", a[0], a[1], a[2], b[0], b[1], b[2])     # Dot3D::Create, never folded


def length3(v):
    if all(is_const(c) for c in v): return const(math.sqrt(sum(c.val * c.val for c in v)))
    return _mk("len3", v[0], v[1], v[2])


# ------------------------------------------------------------------------------------------------
# control flow (src/chad.h:1444-1533)
# ------------------------------------------------------------------------------------------------
class Block:
    def __init__(self):
        self.exprs = []
        self.next = None


class Split:
    _n = [0]

    def __init__(self):
        self.conds = []
        self.children = []
        self.outputs = []
        self.next = None
        Split._n[0] += 1
        self.sid = Split._n[0]


class Function:
    def __init__(self):
        self.first = self.cur = Block()
        self.stack = []


def begin_function():
    G.f = Function()
    return G.f


def end_function():
    f = G.f
    assert not f.stack
    f.cur = None
    return f


class Cond:
    def __init__(self, kind, a=None, b=None):
        self.kind, self.a, self.b = kind, a, b


def Gt(a, b): return Cond("gt", lift(a), lift(b))
def Gte(a, b): return Cond("ge", lift(a), lift(b))
def Lt(a, b): return Cond("lt", lift(a), lift(b))
def Lte(a, b): return Cond("le", lift(a), lift(b))
def Eq(a, b): return Cond("eq", lift(a), lift(b))
def And(a, b): return Cond("and", a, b)


def begin_if(cond, nout):
    sp = Split()
    outs = [Node.__new__(Node) for _ in range(nout)]
    for o in outs:                      # CondExpr::Create does not enter any block
        o.op, o.args, o.val, o.diff, o.name = "cond", [], None, False, None
        _counter[0] += 1
        o.uid = _counter[0]
    sp.outputs = outs
    G.f.cur.next = sp
    G.f.stack.append(sp)
    sp.conds.append(cond)
    b = Block()
    sp.children.append(b)
    G.f.cur = b
    return outs


def begin_else_if(cond):
    sp = G.f.stack[-1]
    sp.conds.append(cond)
    b = Block()
    sp.children.append(b)
    G.f.cur = b


def begin_else():
    begin_else_if(None)


def set_cond_output(exprs):
    sp = G.f.stack[-1]
    assert len(exprs) == len(sp.outputs)
    for o, e in zip(sp.outputs, exprs):
        e = lift(e)
        o.args.append(e)
        o.diff = o.diff or e.diff


def end_if():
    sp = G.f.stack.pop()
    b = Block()
    sp.next = b
    G.f.cur = b


def if_else(cond, x, y):                 # src/utils.h:92-102
    r = begin_if(cond, 1)
    set_cond_output([x])
    begin_else()
    set_cond_output([y])
    end_if()
    return r[0]


def fabs(x):                             # src/chad.h:1226-1234
    x = lift(x)
    r = begin_if(Gte(x, 0.0), 1)
    set_cond_output([x])
    begin_else()
    set_cond_output([-x])
    end_if()
    return r[0]


def fmax(a, b):                          # src/chad.h:1236-1244
    a, b = lift(a), lift(b)
    r = begin_if(Gte(a, b), 1)
    set_cond_output([a])
    begin_else()
    set_cond_output([b])
    end_if()
    return r[0]


# ------------------------------------------------------------------------------------------------
# forward / reverse statement lists
# ------------------------------------------------------------------------------------------------
def build_forward(block):
    """[('def', node) | ('if', split, [(cond, body, [(output, expr)])])] in emission order
    (src/chad.cpp:109-170)."""
    items = [("def", e) for e in block.exprs]
    if block.next is not None:
        sp = block.next
        branches = []
        for i, child in enumerate(sp.children):
            body = build_forward(child)
            outs = [(o, o.args[i]) for o in sp.outputs]
            branches.append((sp.conds[i], body, outs))
        items.append(("if", sp, branches))
        if sp.next is not None:
            items += build_forward(sp.next)
    return items


def _has_acc(n):
    return n.diff and n.op != "const"


def build_reverse(block, nz):
    """Reverse statement list in the order src/chad.cpp:230-331 emits it.
    [('acc', child, expr, k) | ('if', split, [(cond, assigns [(expr, output)], body)])].
    `nz` is the growing set of nodes some emitted statement accumulates into (nonZeroAccId)."""
    items = []
    if block.next is not None:
        sp = block.next
        if sp.next is not None:
            items += build_reverse(sp.next, nz)
        if any(o in nz for o in sp.outputs):
            branches = []
            for i, child in enumerate(sp.children):
                assigns = []
                for o in sp.outputs:
                    if o not in nz:
                        continue
                    e = o.args[i]
                    if not _has_acc(e):
                        continue
                    assigns.append((e, o))
                    nz.add(e)
                body = build_reverse(child, nz)
                branches.append((sp.conds[i], assigns, body))
            items.append(("if", sp, branches))
    for e in reversed(block.exprs):
        if e not in nz:
            continue
        for k, c in enumerate(e.args):
            if not _has_acc(c):
                continue
            if e.op == "pow" and k == 1:
                continue
            items.append(("acc", c, e, k))
            nz.add(c)
    return items


# ------------------------------------------------------------------------------------------------
# numeric interpretation (float64)
# ------------------------------------------------------------------------------------------------
def _eval_cond(c, env):
    if c.kind == "and":
        return _eval_cond(c.a, env) and _eval_cond(c.b, env)
    a, b = _value(c.a, env), _value(c.b, env)
    return {"gt": a > b, "ge": a >= b, "lt": a < b, "le": a <= b, "eq": a == b}[c.kind]


def _value(n, env):
    if n.op == "const":
        return n.val
    return env[n]


def _safe(fn, *a):
    try:
        return fn(*a)
    except (ValueError, ZeroDivisionError, OverflowError):
        return float("nan")


def _eval_node(n, env):
    a = [_value(x, env) for x in n.args]
    op = n.op
    if op == "add": return a[0] + a[1]
    if op == "sub": return a[0] - a[1]
    if op == "mul": return a[0] * a[1]
    if op == "div": return _safe(lambda: a[0] / a[1])
    if op == "neg": return -a[0]
    if op == "sq": return a[0] * a[0]
    if op == "inv": return _safe(lambda: 1.0 / a[0])
    if op == "sin": return _safe(math.sin, a[0])
    if op == "cos": return _safe(math.cos, a[0])
    if op == "sqrt": return _safe(math.sqrt, a[0])
    if op == "exp": return _safe(math.exp, a[0])
    if op == "log": return _safe(math.log, a[0])
    if op == "acos": return _safe(math.acos, a[0])
    if op == "pow": return _safe(math.pow, a[0], a[1])
    if op == "atan2": return math.atan2(a[0], a[1])
    if op == "dot3": return a[0] * a[3] + a[1] * a[4] + a[2] * a[5]
    if op == "len3": return math.sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2])
    raise KeyError(op)


def _partial(n, k, env):
    a = [_value(x, env) for x in n.args]
    r = env[n]
    op = n.op
    if op == "add": return 1.0
    if op == "sub": return 1.0 if k == 0 else -1.0
    if op == "mul": return a[1 - k]
    if op == "div": return _safe(lambda: 1.0 / a[1]) if k == 0 else _safe(lambda: -r / a[1])
    if op == "neg": return -1.0
    if op == "sq": return 2.0 * a[0]
    if op == "inv": return -r * r
    if op == "sin": return math.cos(a[0])
    if op == "cos": return -math.sin(a[0])
    if op == "sqrt": return _safe(lambda: 0.5 / r)
    if op == "exp": return r
    if op == "log": return _safe(lambda: 1.0 / a[0])
    if op == "acos": return _safe(lambda: -1.0 / math.sqrt(1.0 - a[0] * a[0]))
    if op == "pow": return _safe(lambda: a[1] * math.pow(a[0], a[1] - 1.0))
    if op == "atan2":
        d = a[0] * a[0] + a[1] * a[1]
        return _safe(lambda: a[1] / d) if k == 0 else _safe(lambda: -a[0] / d)
    if op == "dot3": return a[(k + 3) % 6]
    if op == "len3": return _safe(lambda: a[k] / r)
    raise KeyError(op)


def run_forward(items, env, taken):
    for it in items:
        if it[0] == "def":
            env[it[1]] = _eval_node(it[1], env)
        else:
            _, sp, branches = it
            for i, (cond, body, outs) in enumerate(branches):
                if cond is None or _eval_cond(cond, env):
                    taken[sp] = i
                    run_forward(body, env, taken)
                    for o, e in outs:
                        env[o] = _value(e, env)
                    break


def run_reverse(items, env, taken, acc, compat=True):
    for it in items:
        if it[0] == "acc":
            _, c, e, k = it
            acc[c] = acc.get(c, 0.0) + acc.get(e, 0.0) * _partial(e, k, env)
        else:
            _, sp, branches = it
            i = taken.get(sp)
            if i is None:
                continue
            cond, assigns, body = branches[i]
            for e, o in assigns:
                if compat:
                    acc[e] = acc.get(o, 0.0)
                else:
                    acc[e] = acc.get(e, 0.0) + acc.get(o, 0.0)
            run_reverse(body, env, taken, acc, compat)


class Program:
    """A traced function: inputs (dict name -> node), outputs (list of nodes)."""

    def __init__(self, func, inputs, outputs):
        self.func, self.inputs, self.outputs = func, inputs, outputs
        self.fwd = build_forward(func.first)
        nz = set(o for o in outputs if _has_acc(o))
        self.rev = build_reverse(func.first, nz)
        self.nz = nz

    def evaluate(self, values, out_adj=None, compat=True):
        """values: dict name -> float.  Returns (output values, dict name -> adjoint of diff inputs)."""
        env, taken = {}, {}
        for name, n in self.inputs.items():
            env[n] = float(values[name])
        run_forward(self.fwd, env, taken)
        outs = [_value(o, env) for o in self.outputs]
        if out_adj is None:
            return outs, None
        acc = {}
        for o, a in zip(self.outputs, out_adj):
            if _has_acc(o):
                acc[o] = acc.get(o, 0.0) + a
        run_reverse(self.rev, env, taken, acc, compat)
        return outs, {name: acc.get(n, 0.0) for name, n in self.inputs.items() if n.diff}
