"""Flatten a module tree into the three-address layer program of include/cnngp.h.

The reference evaluates the tree recursively, materialising every intermediate patch in HBM
(Sequential.propagate cnn_gp/kernels.py:184-187, Sum.propagate :252-254, Mixture.propagate
:221-225, element-wise combination cnn_gp/kernel_patch.py:43-63).  Here the tree is compiled
once into ops over a handful of numbered map slots that the kernels keep on chip.

Ownership protocol used by ``module._emit(builder, src, owned)``:
  owned=True   the callee may overwrite ``src``; it must return it or release it
  owned=False  ``src`` is still needed by the caller (a Sum / Mixture input); do not write it
The return value is ``(slot, owned)``; ``owned=False`` only when the module is the identity
and hands back the caller's slot.
"""
import numpy as np

from . import _native as nat


class ProgramBuilder:
    def __init__(self):
        self.ops = []
        self.n_slots = 1  # slot 0 = initial covariance map (kernels.py:43-49)
        self._free = []

    # -- slots --------------------------------------------------------------------------
    def alloc(self):
        if self._free:
            return self._free.pop()
        self.n_slots += 1
        return self.n_slots - 1

    def release(self, slot):
        assert slot not in self._free
        self._free.append(slot)

    # -- ops ----------------------------------------------------------------------------
    def _op(self, opcode, src, dst, **kw):
        self.ops.append(nat.Op(opcode=opcode, src=src, dst=dst, ke=kw.get("ke", 0),
                               zero_first=kw.get("zero_first", 0), stride=kw.get("stride", 1),
                               pad=kw.get("pad", 0), dil=kw.get("dil", 1),
                               scale=kw.get("scale", 1.0), bias=kw.get("bias", 0.0)))

    def _dst_for(self, src, owned):
        return src if owned else self.alloc()

    def conv(self, src, owned, kernel_size, zero_first, stride, pad, dil, var_weight, var_bias):
        dst = self._dst_for(src, owned)
        # the reference stores the tap as a float32 buffer (kernels.py:87-88); a .double()
        # model widens that rounded value, so the float32 rounding is part of the semantics
        tap = float(np.float32(float(var_weight) / int(kernel_size) ** 2))
        self._op(nat.OP_CONV, src, dst, ke=int(kernel_size) + (1 if zero_first else 0),
                 zero_first=int(bool(zero_first)), stride=int(stride), pad=int(pad), dil=int(dil),
                 scale=tap, bias=float(var_bias))
        return dst, True

    def relu(self, src, owned):
        dst = self._dst_for(src, owned)
        self._op(nat.OP_RELU, src, dst)
        return dst, True

    def scale(self, src, owned, factor):
        dst = self._dst_for(src, owned)
        self._op(nat.OP_SCALE, src, dst, scale=float(factor))
        return dst, True

    def copy(self, src):
        dst = self.alloc()
        self._op(nat.OP_COPY, src, dst)
        return dst

    def add_into(self, dst, src):
        self._op(nat.OP_ADD, src, dst)

    # -- combinators --------------------------------------------------------------------
    def sequential(self, mods, src, owned):
        cur, own = src, owned
        for m in mods:
            cur, own = m._emit(self, cur, own)
        return cur, own

    def branches(self, mods, src, owned, weights=None):
        """Sum (weights=None) or Mixture: every branch reads the same input, results are added
        left to right.  a + b == b + a exactly in IEEE arithmetic, so the accumulator may be
        whichever operand we own."""
        acc = None  # (slot, owned)
        src_consumed = False
        n = len(mods)
        for idx, m in enumerate(mods):
            last = idx == n - 1
            give = bool(owned and last and not (acc is not None and acc[0] == src))
            r, r_owned = m._emit(self, src, give)
            if give:
                src_consumed = True
            if weights is not None:
                r, r_owned = self.scale(r, r_owned, weights[idx])
            if acc is None:
                acc = (r, r_owned)
                continue
            a, a_owned = acc
            if a_owned:
                self.add_into(a, r)
                if r_owned:
                    self.release(r)
            elif r_owned:
                self.add_into(r, a)
                acc = (r, True)
            else:  # two borrowed operands (e.g. Sum of two identities)
                t = self.copy(a)
                self.add_into(t, r)
                acc = (t, True)
        if acc is None:
            raise ValueError("Sum / Mixture needs at least one module")
        if owned and not src_consumed:
            if acc[0] == src:
                acc = (src, True)
            else:
                self.release(src)
        return acc

    def finish(self, result_slot):
        if not self.ops or self.ops[-1].dst != result_slot:
            if result_slot != 0 or self.ops:
                # the plan takes the last op's dst as the result slot
                t = self.alloc()
                self._op(nat.OP_COPY, result_slot, t)
        return self.ops, self.n_slots


def compile_model(model):
    """-> (ops, n_slots).  The root module owns slot 0."""
    b = ProgramBuilder()
    slot, _ = model._emit(b, 0, True)
    return b.finish(slot)


def describe(ops):
    """Human-readable listing (debugging / DESIGN.md)."""
    names = {nat.OP_CONV: "CONV", nat.OP_RELU: "RELU", nat.OP_COPY: "COPY", nat.OP_ADD: "ADD",
             nat.OP_SCALE: "SCALE"}
    out = []
    for o in ops:
        s = f"{names[o.opcode]:5s} s{o.src}->s{o.dst}"
        if o.opcode == nat.OP_CONV:
            s += f" ke={o.ke} z={o.zero_first} st={o.stride} p={o.pad} d={o.dil} tap={o.scale:.6g} b={o.bias:.6g}"
        elif o.opcode == nat.OP_SCALE:
            s += f" x{o.scale:.6g}"
        out.append(s)
    return "\n".join(out)
