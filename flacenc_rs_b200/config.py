"""Host-side mirror of the reference's ``config`` module (/root/reference/src/config.rs).

Same names, nesting, defaults and verification rules as ``config::Encoder`` and its children; the
flattened POD handed to the C ABI is ``_ffi.Config`` (``fb200_config``).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

from . import _ffi
from .error import VerifyError


@dataclass
class StereoCoding:
    """src/config.rs:137-151"""
    use_leftside: bool = True
    use_rightside: bool = True
    use_midside: bool = True


@dataclass
class Prc:
    """src/config.rs:211-229"""
    max_parameter: int = 30


@dataclass
class OrderSel:
    """src/config.rs:400-432: ``BitCount`` or ``ApproxEnt { partitions }``"""
    type: str = "ApproxEnt"
    partitions: int = 16

    @staticmethod
    def BitCount() -> "OrderSel":
        return OrderSel("BitCount", 16)

    @staticmethod
    def ApproxEnt(partitions: int = 16) -> "OrderSel":
        return OrderSel("ApproxEnt", partitions)


@dataclass
class Fixed:
    """src/config.rs:236-264"""
    max_order: int = 4
    order_sel: OrderSel = field(default_factory=OrderSel)


@dataclass
class Window:
    """src/config.rs:344-387: ``Rectangle`` or ``Tukey { alpha }``"""
    type: str = "Tukey"
    alpha: float = 0.4

    @staticmethod
    def Rectangle() -> "Window":
        return Window("Rectangle", 0.0)

    @staticmethod
    def Tukey(alpha: float = 0.4) -> "Window":
        return Window("Tukey", alpha)


@dataclass
class Qlpc:
    """src/config.rs:271-326"""
    lpc_order: int = 10
    quant_precision: int = 15
    use_direct_mse: bool = False
    mae_optimization_steps: int = 0
    window: Window = field(default_factory=Window)
    # extension beyond the reference (opt-in, see include/flacenc_b200.h): also try this many lower LPC orders
    ext_order_search: int = 0
    ext_precision_search: int = 0


@dataclass
class SubFrameCoding:
    """src/config.rs:167-204"""
    use_constant: bool = True
    use_fixed: bool = True
    use_lpc: bool = True
    fixed: Fixed = field(default_factory=Fixed)
    qlpc: Qlpc = field(default_factory=Qlpc)
    prc: Prc = field(default_factory=Prc)


@dataclass
class Encoder:
    """src/config.rs:85-130"""
    block_size: int = 4096
    multithread: bool = True
    workers: Optional[int] = None
    stereo_coding: StereoCoding = field(default_factory=StereoCoding)
    subframe_coding: SubFrameCoding = field(default_factory=SubFrameCoding)

    def to_pod(self) -> _ffi.Config:
        s, q, f = self.subframe_coding, self.subframe_coding.qlpc, self.subframe_coding.fixed
        if f.order_sel.type not in ("BitCount", "ApproxEnt"):
            raise VerifyError("subframe_coding.fixed.order_sel", "unknown selector")
        if q.window.type not in ("Rectangle", "Tukey"):
            raise VerifyError("subframe_coding.qlpc.window", "unknown window")
        return _ffi.Config(
            block_size=self.block_size, multithread=int(self.multithread), workers=int(self.workers or 0),
            use_leftside=int(self.stereo_coding.use_leftside), use_rightside=int(self.stereo_coding.use_rightside),
            use_midside=int(self.stereo_coding.use_midside),
            use_constant=int(s.use_constant), use_fixed=int(s.use_fixed), use_lpc=int(s.use_lpc),
            fixed_max_order=f.max_order, fixed_order_sel=0 if f.order_sel.type == "BitCount" else 1,
            approx_ent_partitions=f.order_sel.partitions,
            lpc_order=q.lpc_order, quant_precision=q.quant_precision, use_direct_mse=int(q.use_direct_mse),
            mae_optimization_steps=q.mae_optimization_steps,
            window_type=0 if q.window.type == "Rectangle" else 1, tukey_alpha=float(q.window.alpha),
            prc_max_parameter=s.prc.max_parameter, ext_lpc_order_search=q.ext_order_search,
            ext_lpc_precision_search=q.ext_precision_search)

    # ---- the reference's TOML form (serde layout of config::Encoder, src/config.rs:20-60; every table `serde(default)`:
    # missing keys keep their defaults; flacenc-bin's `--config` / `--dump-config`, flacenc-bin/src/main.rs:96-101,204-214)
    @staticmethod
    def from_toml(text: str) -> "Encoder":
        import tomllib
        try:
            doc = tomllib.loads(text)
        except tomllib.TOMLDecodeError as exc:
            raise VerifyError("config", f"not TOML: {exc}") from None
        e = Encoder()

        def take(table, obj, keys, path):
            if not isinstance(table, dict):
                raise VerifyError(path, "expected a table")
            for k, v in table.items():
                if k in keys:
                    want = type(getattr(obj, k)) if getattr(obj, k) is not None else int
                    if want is float and isinstance(v, int) and not isinstance(v, bool):
                        v = float(v)
                    if not isinstance(v, want) or (want is int and isinstance(v, bool)):
                        raise VerifyError(f"{path}.{k}".lstrip("."), f"expected {want.__name__}")
                    setattr(obj, k, v)
                elif not isinstance(v, dict):
                    raise VerifyError(f"{path}.{k}".lstrip("."), "unknown key")

        take(doc, e, ("block_size", "multithread", "workers"), "")
        take(doc.get("stereo_coding", {}), e.stereo_coding, ("use_leftside", "use_rightside", "use_midside"), "stereo_coding")
        sc = doc.get("subframe_coding", {})
        take(sc, e.subframe_coding, ("use_constant", "use_fixed", "use_lpc"), "subframe_coding")
        fx = sc.get("fixed", {}) if isinstance(sc, dict) else {}
        take(fx, e.subframe_coding.fixed, ("max_order",), "subframe_coding.fixed")
        osel = fx.get("order_sel") if isinstance(fx, dict) else None
        if osel is not None:
            if not isinstance(osel, dict) or osel.get("type") not in ("BitCount", "ApproxEnt"):
                raise VerifyError("subframe_coding.fixed.order_sel", "type must be \"BitCount\" or \"ApproxEnt\"")
            e.subframe_coding.fixed.order_sel = (OrderSel.BitCount() if osel["type"] == "BitCount"
                                                 else OrderSel.ApproxEnt(int(osel.get("partitions", 16))))
        ql = sc.get("qlpc", {}) if isinstance(sc, dict) else {}
        take(ql, e.subframe_coding.qlpc, ("lpc_order", "quant_precision", "use_direct_mse", "mae_optimization_steps",
                                          "ext_order_search", "ext_precision_search"), "subframe_coding.qlpc")
        win = ql.get("window") if isinstance(ql, dict) else None
        if win is not None:
            if not isinstance(win, dict) or win.get("type") not in ("Rectangle", "Tukey"):
                raise VerifyError("subframe_coding.qlpc.window", "type must be \"Rectangle\" or \"Tukey\"")
            e.subframe_coding.qlpc.window = (Window.Rectangle() if win["type"] == "Rectangle"
                                             else Window.Tukey(float(win.get("alpha", 0.4))))
        take(sc.get("prc", {}) if isinstance(sc, dict) else {}, e.subframe_coding.prc, ("max_parameter",), "subframe_coding.prc")
        return e

    def to_toml(self) -> str:
        s, q, f = self.subframe_coding, self.subframe_coding.qlpc, self.subframe_coding.fixed
        b = lambda v: "true" if v else "false"  # noqa: E731
        lines = [f"block_size = {self.block_size}", f"multithread = {b(self.multithread)}"]
        if self.workers:
            lines.append(f"workers = {int(self.workers)}")
        lines += ["", "[stereo_coding]", f"use_leftside = {b(self.stereo_coding.use_leftside)}",
                  f"use_rightside = {b(self.stereo_coding.use_rightside)}", f"use_midside = {b(self.stereo_coding.use_midside)}",
                  "", "[subframe_coding]", f"use_constant = {b(s.use_constant)}", f"use_fixed = {b(s.use_fixed)}",
                  f"use_lpc = {b(s.use_lpc)}", "", "[subframe_coding.fixed]", f"max_order = {f.max_order}",
                  "", "[subframe_coding.fixed.order_sel]", f"type = \"{f.order_sel.type}\""]
        if f.order_sel.type == "ApproxEnt":
            lines.append(f"partitions = {f.order_sel.partitions}")
        lines += ["", "[subframe_coding.qlpc]", f"lpc_order = {q.lpc_order}", f"quant_precision = {q.quant_precision}",
                  f"use_direct_mse = {b(q.use_direct_mse)}", f"mae_optimization_steps = {q.mae_optimization_steps}"]
        if q.ext_order_search or q.ext_precision_search:  # (extensions of this library; absent = the reference's frames)
            lines += [f"ext_order_search = {q.ext_order_search}", f"ext_precision_search = {q.ext_precision_search}"]
        lines += ["", "[subframe_coding.qlpc.window]", f"type = \"{q.window.type}\""]
        if q.window.type == "Tukey":
            lines.append(f"alpha = {float(q.window.alpha)!r}")
        lines += ["", "[subframe_coding.prc]", f"max_parameter = {s.prc.max_parameter}", ""]
        return "\n".join(lines)

    def verify(self) -> None:
        """``Verify::verify`` (src/config.rs:109-130 and children); raises VerifyError."""
        pod = self.to_pod()
        if _ffi.lib().fb200_config_verify(C.byref(pod)) != _ffi.OK:
            raise VerifyError("config", "configuration out of range")

    def into_verified(self) -> "Verified":
        """``Verify::into_verified`` (src/error.rs:278-304)."""
        self.verify()
        return Verified(self)


class Verified:
    """``Verified<config::Encoder>`` (src/error.rs:268-276): only obtainable through verification."""

    def __init__(self, inner: Encoder):
        self._inner = inner
        self._pod = inner.to_pod()

    @property
    def inner(self) -> Encoder:
        return self._inner

    @property
    def pod(self) -> _ffi.Config:
        return self._pod
