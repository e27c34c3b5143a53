"""Static shape of one HPMN memory side and the flat parameter layout (host mirror of
include/hpmn_b200.h).  Variable names are the ones the reference's TF scopes produce
(/root/reference/code/hpmn.py:117,173-174,137-139,190-195,433-465), so a TF checkpoint maps 1:1."""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

from . import _lib

ATT_FC1, ATT_FC2, HEAD_FC1, HEAD_FC2 = 80, 40, 200, 80


@dataclass
class HpmnShape:
    B: int
    T: int
    F: int
    E: int
    H: int
    periods: Sequence[int]      # li_layer (only the first L-1 entries are used, hpmn.py:122-128)
    L: int
    hops: int
    V: int
    front_pad: int = 0          # Hpmn_Industry: 23 (hpmn.py:288-289)
    mask_id0: bool = True       # Hpmn: True (hpmn.py:417-423); Hpmn_Industry: False
    last_offset: int = 1        # Hpmn: 1 (hpmn.py:439); Hpmn_Industry: 2 (hpmn.py:292)
    scope: str = "User"

    @property
    def D(self) -> int:
        return self.F * self.E

    @property
    def Tpad(self) -> int:
        return self.T + self.front_pad

    def steps(self) -> List[int]:
        s, out = self.Tpad, []
        for k in range(self.L):
            out.append(s)
            if k < self.L - 1:
                p = self.periods[k]
                if s % p:
                    raise ValueError("layer %d: %d steps not divisible by period %d" % (k, s, p))
                s //= p
        return out

    def with_batch(self, B: int) -> "HpmnShape":
        return HpmnShape(B, self.T, self.F, self.E, self.H, list(self.periods), self.L, self.hops, self.V,
                         self.front_pad, self.mask_id0, self.last_offset, self.scope)

    def to_c(self) -> "_lib.hpmn_shape":
        c = _lib.hpmn_shape()
        c.B, c.T, c.F, c.E, c.H, c.L, c.hops = self.B, self.T, self.F, self.E, self.H, self.L, self.hops
        c.front_pad, c.mask_id0, c.last_offset, c.V = self.front_pad, int(self.mask_id0), self.last_offset, self.V
        for k in range(min(self.L - 1, len(self.periods))):
            c.periods[k] = int(self.periods[k])
        return c

    # algorithmic work (DESIGN.md / BASELINE.md section 3)
    def gru_flops_fwd_per_sample(self) -> int:
        tot = 0
        for k, s in enumerate(self.steps()):
            din = self.D if k == 0 else self.H
            tot += s * 2 * (din + self.H) * 3 * self.H
        return tot

    def gather_bytes(self) -> int:
        return self.B * self.T * self.F * (4 + 4 * self.E)


def param_names(sh: HpmnShape) -> "OrderedDict[str, Tuple[int, ...]]":
    o: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    H, D, sc = sh.H, sh.D, sh.scope
    for k in range(sh.L):
        din = D if k == 0 else H
        base = "%s/GRU%d/rnn/gru_cell/" % (sc, k)
        o[base + "gates/kernel"] = (din + H, 2 * H)
        o[base + "gates/bias"] = (2 * H,)
        o[base + "candidate/kernel"] = (din + H, H)
        o[base + "candidate/bias"] = (H,)
    o[sc + "/dense/kernel"] = (D, H)
    o[sc + "/dense/bias"] = (H,)
    o[sc + "/map"] = (H, H)
    n = 1
    for _ in range(sh.hops):
        for (a, b) in ((4 * H, ATT_FC1), (ATT_FC1, ATT_FC2), (ATT_FC2, 1)):
            o["%s/dense_%d/kernel" % (sc, n)] = (a, b)
            o["%s/dense_%d/bias" % (sc, n)] = (b,)
            n += 1
    R = H + D
    o["output/bn1/gamma"] = (R,)
    o["output/bn1/beta"] = (R,)
    o["output/fc1/kernel"] = (R, HEAD_FC1)
    o["output/fc1/bias"] = (HEAD_FC1,)
    o["output/fc2/kernel"] = (HEAD_FC1, HEAD_FC2)
    o["output/fc2/bias"] = (HEAD_FC2,)
    o["output/fc3/kernel"] = (HEAD_FC2, 1)
    o["output/fc3/bias"] = (1,)
    return o


def param_layout(sh: HpmnShape) -> Tuple["OrderedDict[str, Tuple[int, Tuple[int, ...]]]", int]:
    """name -> (float offset, shape); total floats.  Every tensor starts on a 4-float boundary."""
    out: "OrderedDict[str, Tuple[int, Tuple[int, ...]]]" = OrderedDict()
    off = 0
    for name, shape in param_names(sh).items():
        n = 1
        for d in shape:
            n *= d
        out[name] = (off, shape)
        off = (off + n + 3) & ~3
    return out, off
