"""Generate quisquis-rust_b200/csrc/fe25519_mp.inc: straight-line multi-precision product / square bodies for the
saturated 8 x 32-bit field representation.

Every 32x32->64 partial product is a (mad.lo.cc, madc.hi.cc) PTX pair that ptxas fuses into ONE
IMAD.WIDE.U32[.X] whose carry travels in a predicate, so accumulation costs no extra instructions.  To keep every
product on a 64-bit aligned accumulator slot, products whose limb position i+j is even go to accumulator E and
those with odd position go to accumulator O (O[k] has weight 2^(32(k+1))); E and O are merged with one add chain.

The generator tracks which limbs have been written ("touched") and exact magnitude bounds, so that first touches
use non-accumulating forms and carry chains stop exactly where the bound proves they can.
Run: python tools/gen_field_ops.py
"""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W = 32
MAXL = (1 << W) - 1


class Acc:
    """One accumulator (E or O): tracks touched limbs and an upper bound on its integer value."""

    def __init__(self, name, nl):
        self.name, self.nl = name, nl
        self.touched = [False] * nl
        self.bound = 0
        self.out = []

    def limb(self, k):
        return "%s[%d]" % (self.name, k)

    def row(self, prods, a_name, b_name):
        """prods: list of (slot_index k, i, j) sorted by k, slots adjacent (k, k+2, ...): acc[k..k+1] += a[i]*b[j]."""
        if not prods:
            return
        ks = [p[0] for p in prods]
        assert all(ks[n + 1] == ks[n] + 2 for n in range(len(ks) - 1)), ks
        any_acc = any(self.touched[k] or self.touched[k + 1] for k in ks)
        for k, i, j in prods:
            self.bound += (MAXL * MAXL) << (W * k)
        if not any_acc:
            for k, i, j in prods:  # independent first touches: plain mul.lo / mul.hi (one IMAD.WIDE, no carry)
                self.out.append("%s = mul_lo(%s[%d], %s[%d]); %s = mul_hi(%s[%d], %s[%d]);" % (
                    self.limb(k), a_name, i, b_name, j, self.limb(k + 1), a_name, i, b_name, j))
                self.touched[k] = self.touched[k + 1] = True
            return
        first = True
        for k, i, j in prods:
            lo_c = self.limb(k) if self.touched[k] else "0u"
            hi_c = self.limb(k + 1) if self.touched[k + 1] else "0u"
            self.out.append("%s = %s(%s[%d], %s[%d], %s);" % (self.limb(k), "mad_lo_cc" if first else "madc_lo_cc", a_name, i, b_name, j, lo_c))
            self.out.append("%s = madc_hi_cc(%s[%d], %s[%d], %s);" % (self.limb(k + 1), a_name, i, b_name, j, hi_c))
            self.touched[k] = self.touched[k + 1] = True
            first = False
        # propagate the carry as far as the value bound requires
        need = (self.bound.bit_length() + W - 1) // W
        t = ks[-1] + 2
        assert all(not self.touched[q] for q in range(max(need, t), self.nl)), (self.name, need, t)
        while t < need:
            last = t == need - 1
            src = self.limb(t) if self.touched[t] else "0u"
            self.out.append("%s = %s(%s, 0u);" % (self.limb(t), "addc" if last else "addc_cc", src))
            self.touched[t] = True
            t += 1


def gen_mul(n, fname):
    """r[2n] = a[n] * b[n] (schoolbook, n even)."""
    E, O = Acc("E", 2 * n), Acc("O", 2 * n)
    lines = []
    for i in range(n):
        pe, po = [], []
        for j in range(n):
            pos = i + j
            if pos % 2 == 0:
                pe.append((pos, j, i))
            else:
                po.append((pos - 1, j, i))
        for acc, pr in ((E, pe), (O, po)):
            acc.out = []
            acc.row(pr, "a", "b")
            lines += acc.out
    return finish(fname, n, E, O, lines, "const u32* a, const u32* b")


def finish(fname, n, E, O, lines, args):
    # r = E + (O << 32)
    top_e = max(k for k in range(2 * n) if E.touched[k])
    top_o = max(k for k in range(2 * n) if O.touched[k])
    assert top_e == 2 * n - 1 and top_o + 1 <= 2 * n - 1, (top_e, top_o)
    lines.append("r[0] = E[0];")
    for k in range(1, 2 * n):
        o = "O[%d]" % (k - 1) if (k - 1) <= top_o and O.touched[k - 1] else "0u"
        e = "E[%d]" % k if E.touched[k] else "0u"
        op = "add_cc" if k == 1 else ("addc" if k == 2 * n - 1 else "addc_cc")
        lines.append("r[%d] = %s(%s, %s);" % (k, op, e, o))
    body = "\n".join("    " + l for l in lines)
    return "QQ_HD void %s(u32* r, %s) {\n    u32 E[%d], O[%d];\n%s\n}\n" % (fname, args, 2 * n, 2 * n, body)


def gen_sq_offdiag(n, fname):
    """r[2n] = sum_{i<j} a[i] a[j] 2^(32(i+j))   (NOT doubled, no diagonal)."""
    E, O = Acc("E", 2 * n), Acc("O", 2 * n)
    lines = []
    for i in range(n - 1):
        pe, po = [], []
        for j in range(i + 1, n):
            pos = i + j
            if pos % 2 == 0:
                pe.append((pos, i, j))
            else:
                po.append((pos - 1, i, j))
        for acc, pr in ((E, pe), (O, po)):
            acc.out = []
            acc.row(pr, "a", "a")
            lines += acc.out
    # off-diagonal sum: E limbs 0,1 and the top limb are never touched
    lines.append("r[0] = 0u;")
    lines.append("r[1] = O[0];")
    top = 2 * n - 1
    first = True
    for k in range(2, 2 * n):
        e = "E[%d]" % k if E.touched[k] else "0u"
        o = "O[%d]" % (k - 1) if O.touched[k - 1] else "0u"
        if e == "0u" and o == "0u" and k == top:
            lines.append("r[%d] = addc(0u, 0u);" % k)
            continue
        op = "add_cc" if first else ("addc" if k == top else "addc_cc")
        first = False
        lines.append("r[%d] = %s(%s, %s);" % (k, op, e, o))
    body = "\n".join("    " + l for l in lines)
    return "QQ_HD void %s(u32* r, const u32* a) {\n    u32 E[%d], O[%d];\n%s\n}\n" % (fname, 2 * n, 2 * n, body)


out = ["// generated by tools/gen_field_ops.py -- do not edit\n"]
out.append(gen_mul(4, "mp_mul4"))
out.append(gen_mul(8, "mp_mul8"))
out.append(gen_sq_offdiag(4, "mp_sqoff4"))
out.append(gen_sq_offdiag(8, "mp_sqoff8"))
path = os.path.join(ROOT, "quisquis-rust_b200", "csrc", "fe25519_mp.inc")
open(path, "w").write("\n".join(out))
print("wrote", path)
