"""flacenc_rs_b200 -- B200-native backend for flacenc-rs's per-frame encode path.

Host-side mirror of the reference API (config::Encoder, source::MemSource / FrameBuf,
encode_with_fixed_block_size, encode_fixed_size_frame) bound over ctypes to the C ABI of
csrc/libflacenc_b200.so (include/flacenc_b200.h).  No CPU fallback: the CUDA library must be built
(`__graft_entry__.build()`) and a device must be visible.
"""
__all__ = ["config", "source", "encoder", "error", "sigen"]
