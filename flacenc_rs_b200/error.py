"""Error types mirroring /root/reference/src/error.rs (VerifyError :195-266, SourceError :500-560,
EncodeError :458-463) and the mapping from the C ABI's return codes."""
from __future__ import annotations


class VerifyError(ValueError):
    """error::VerifyError: an input (config, samples, frame number) is out of the accepted range."""

    def __init__(self, component: str, reason: str):
        super().__init__(f"verification error: `{component}` is not valid. reason: {reason}")
        self.component = component
        self.reason = reason


class SourceError(IOError):
    """error::SourceError: reading from a Source failed / bad buffer."""


class EncodeError(RuntimeError):
    """error::EncodeError wrapper for failures that are neither Config nor Source (CUDA, capacity)."""


def raise_for_code(code: int, detail: str = "") -> None:
    """FB200_* return code -> exception (0 = ok)."""
    if code == 0:
        return
    if code == 1:
        raise VerifyError("input", detail or "config/sample/frame-number verification failed")
    if code == 2:
        raise SourceError(detail or "source/argument error")
    if code == 3:
        raise EncodeError("CUDA error (flacenc_rs_b200 has no CPU fallback): " + detail)
    if code == 4:
        raise EncodeError("output capacity too small: " + detail)
    raise EncodeError(f"unknown error code {code}: {detail}")
