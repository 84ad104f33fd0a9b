"""Checks the CUDA kernel bodies' logic against the oracle on the CPU, by running
flacenc_rs_b200/csrc/fb_kernels.cuh under the phase-by-phase CTA emulation in tests/emu.
The same comparisons run against the real kernels in tests/test_gpu_parity.py (-m gpu)."""
import os

import numpy as np
import pytest

from conftest import crafted_huge_residual_stereo, fuzz_frame_case, load_fixture, pack_pcm, random_case, shift_cases
from flacenc_rs_b200 import sigen
from oracle import oracle as O
from emu import emu as E


def _cfg_pair(**kw):
    return O.default_config(**kw), E.default_config(**kw)


def _compare(signal, channels, bps, rate, block_size, container=None, first_frame=0, **cfgkw):
    ocfg, ecfg = _cfg_pair(**cfgkw)
    signal = np.ascontiguousarray(signal, np.int32).reshape(-1, channels)
    n = len(signal)
    container = container or (bps + 7) // 8
    ref, ref_sizes = O.encode_frames(ocfg, signal, channels, bps, rate, block_size, first_frame_number=first_frame)
    # the device paths: the fused per-frame kernels (when the batch is eligible; 16-bit stereo in a 2-byte container
    # also with the pack kernel staging planes instead of the PCM pairs) and the generic K2/K3 kernels
    modes = [(False, True), (True, True)]
    if channels == 2 and bps == 16 and container == 2:
        modes.insert(1, (False, False))
    for force_generic, kp_pairs in modes:
        E.set_force_generic(force_generic)
        E.set_kp_pairs(kp_pairs)
        try:
            rc, got, sizes, _ = E.encode_interleaved(ecfg, pack_pcm(signal, container), container, n, channels, bps,
                                                     rate, block_size, first_frame)
        finally:
            E.set_force_generic(False)
            E.set_kp_pairs(True)
        assert rc == 0
        assert list(sizes) == list(ref_sizes), f"force_generic={force_generic}"
        if got != ref:
            # locate the first differing frame for the report
            off = 0
            for i, s in enumerate(ref_sizes):
                if got[off:off + s] != ref[off:off + s]:
                    raise AssertionError(f"frame {i} differs (size {s}), force_generic={force_generic}")
                off += s
        assert got == ref
    out, nf = O.decode_frames(got, channels, bps)
    assert np.array_equal(out, signal)


def test_log2f_matches_libm_exhaustive():
    """fb_log2f -- the device code of estimate_entropy's log2 (/root/reference/src/coding.rs:200-227), compiled for
    the CPU -- against the host glibc log2f for EVERY non-negative float bit pattern (zero, subnormals, normals,
    +inf, and the first NaNs), plus a slice of the negative ones.  Bit-equal or both NaN.  (The -m gpu suite runs
    the same sweep on the device build of the function.)"""
    import ctypes as C
    L = E.lib()
    first_bad = C.c_uint32(0)
    threads = min(16, os.cpu_count() or 1)
    bad = L.fbemu_log2f_sweep(0, 0x7F800000 + 4096, threads, C.byref(first_bad))
    assert bad == 0, f"{bad} mismatches, first at bits {first_bad.value:#010x}"
    bad = L.fbemu_log2f_sweep(0x80000000, 1 << 22, threads, C.byref(first_bad))   # -0, negative: NaN / -inf
    assert bad == 0, f"{bad} mismatches, first at bits {first_bad.value:#010x}"


def test_find_shift_matches_libm_around_powers_of_two():
    """fb_find_shift (device code) == the oracle's ceil(log2()) through the host libm, on crafted coefficients"""
    import ctypes as C
    L = E.lib()
    vals = shift_cases()
    n_inexact = 0
    for i, v in enumerate(vals):
        # every value at the default precision; every 40th also negated and at the other precisions
        for prec in ((1, 4, 8, 15) if i % 40 == 0 else (15,)):
            for sign in ((1.0, -1.0) if i % 40 == 0 else (1.0,)):
                c = (C.c_double * 3)(sign * v, 0.0, v * 0.5)
                got, want = L.fbemu_find_shift(c, 3, prec), O.find_shift([sign * v, 0.0, v * 0.5], prec)
                assert got == want, (v.hex() if isinstance(v, float) else v, prec, got, want)
    # the sweep really contains values whose rounded log2 falls back onto the integer (the inexact case)
    import math
    for v in vals:
        if 0 < v < float("inf"):
            m, e = math.frexp(v)
            if m != 0.5 and math.log2(v) == float(e - 1):
                n_inexact += 1
    assert n_inexact > 50


def test_default_config_matches_oracle():
    o, e = _cfg_pair()
    assert bytes(o) == bytes(e)


def test_frame_header_matches_oracle():
    L = E.lib()
    for n in (192, 256, 576, 1000, 4096, 4608, 100, 2728, 32767, 16384):
        for ch_tag in (0, 1, 7, 8, 9, 10):
            for bps, rate in ((16, 44100), (24, 96000), (8, 12345), (20, 48000), (16, 11025), (24, 95800)):
                for num in (0, 1, 127, 128, 2047, 2048, 65535, 65536, (1 << 21), (1 << 26) + 5, (1 << 31) - 1):
                    a = np.zeros(16, np.uint8)
                    b = np.zeros(16, np.uint8)
                    ka = L.fbemu_frame_header(n, ch_tag, bps, rate, num, a.ctypes.data_as(E.C.POINTER(E.C.c_uint8)))
                    kb = O.lib().fo_frame_header_bytes(n, ch_tag, bps, rate, 0, num, O._p(b, O.C.c_uint8))
                    assert ka == kb and bytes(a[:ka]) == bytes(b[:kb])


def test_cd_stereo_noisy_sine_c1_slice():
    """config 1 shape: 44.1 kHz / 16-bit / stereo, block 4096, short tail"""
    x = sigen.noisy_sine_pcm(4096 * 3 + 2728, 2, 16, 44100)
    _compare(x, 2, 16, 44100, 4096)


@pytest.mark.parametrize("name", ["sus109", "sus6", "ras22", "ras103"])
def test_fixture_clips(name):
    x = np.stack([load_fixture(name, 0), load_fixture(name, 1)], axis=1)
    _compare(x, 2, 16, 44100, 4096)
    _compare(x[:, 0], 1, 16, 44100, 4096)


def test_96k_24bit_order24_block4608():
    """config 3 shape"""
    x = sigen.noisy_sine_pcm(4608 * 2, 2, 24, 96000, config_id=3)
    _compare(x, 2, 24, 96000, 4608, lpc_order=24, quant_precision=15)


def test_8ch_24bit():
    """config 5 shape"""
    x = sigen.noisy_sine_pcm(4096 + 1000, 8, 24, 48000, config_id=5)
    _compare(x, 8, 24, 48000, 4096)


def test_rectangle_window_c4_shape():
    x = sigen.noisy_sine_pcm(4096 * 2, 2, 16, 44100, config_id=4)
    _compare(x, 2, 16, 44100, 4096, window_type=0)


@pytest.mark.parametrize("cfgkw", [
    dict(use_lpc=0), dict(use_fixed=0), dict(use_constant=0), dict(fixed_order_sel=0),
    dict(use_leftside=0, use_rightside=0), dict(use_midside=0), dict(use_leftside=0, use_rightside=0, use_midside=0),
    dict(lpc_order=1, quant_precision=3), dict(lpc_order=7, quant_precision=12, tukey_alpha=0.1),
    dict(fixed_max_order=0), dict(fixed_max_order=2, approx_ent_partitions=32), dict(approx_ent_partitions=1),
    dict(approx_ent_partitions=64), dict(prc_max_parameter=0), dict(prc_max_parameter=3), dict(prc_max_parameter=14),
    dict(fixed_max_order=9), dict(tukey_alpha=1.0), dict(tukey_alpha=0.0),
])
def test_config_variations(cfgkw):
    x = sigen.Sine(60, 0.5).noise(0.05, seed=2).to_vec_quantized(16, 2 * 3000).reshape(3000, 2)
    _compare(x, 2, 16, 44100, 1024, **cfgkw)


@pytest.mark.parametrize("block_size", [32, 63, 64, 65, 100, 192, 576, 1000, 1152, 4000, 4095, 4097, 16384])
def test_block_sizes(block_size):
    n = block_size * 2 + block_size // 3 + 1
    x = sigen.Sine(77, 0.6).noise(0.02, seed=block_size).to_vec_quantized(16, n)
    _compare(x, 1, 16, 32000, block_size)


def test_max_block_size_8ch_global_pack_path():
    """32767-sample blocks with 8 channels do not fit shared memory: frames are assembled in their global slot"""
    x = sigen.noisy_sine_pcm(32767 + 500, 8, 24, 96000, config_id=7)
    _compare(x, 8, 24, 96000, 32767)


@pytest.mark.parametrize("channels", [1, 2, 3, 5, 8])
@pytest.mark.parametrize("bps,container", [(8, 1), (12, 2), (16, 2), (20, 3), (24, 3), (24, 4), (16, 4)])
def test_channels_and_sample_formats(channels, bps, container):
    n = 2500
    chans = [sigen.Sine(30 + 7 * c, 0.7).noise(0.01, seed=c).to_vec_quantized(bps, n) for c in range(channels)]
    _compare(np.stack(chans, axis=1), channels, bps, 48000, 1024, container=container)


def test_8bit_in_16bit_container():
    x = sigen.Sine(40, 0.9).noise(0.1, seed=3).to_vec_quantized(8, 5000).reshape(2500, 2)
    _compare(x, 2, 8, 22050, 512, container=2)


def test_special_signals():
    z = np.zeros((3000, 2), np.int32)
    _compare(z, 2, 16, 44100, 1024)                                  # constant zero
    c = np.full((3000, 2), -1234, np.int32)
    c[:, 1] = 77
    _compare(c, 2, 16, 44100, 1024)                                  # constant non-zero, S constant too
    fs = np.tile(np.array([[32767, -32768], [-32768, 32767]], np.int32), (1500, 1))
    _compare(fs, 2, 16, 44100, 1024)                                 # full-scale alternation: verbatim / huge residuals
    rng = np.random.default_rng(5)
    w = rng.integers(-(1 << 23), 1 << 23, (3000, 2)).astype(np.int32)
    _compare(w, 2, 24, 96000, 1024)                                  # white noise 24-bit: verbatim
    imp = np.zeros((4096, 1), np.int32)
    imp[1000] = 30000
    _compare(imp, 1, 16, 44100, 4096)                                # impulse: mixed leaf parameters
    step = np.concatenate([np.zeros(2000, np.int32), np.full(2096, 20000, np.int32)]).reshape(-1, 1)
    _compare(step, 1, 16, 44100, 4096)
    loudquiet = np.concatenate([rng.integers(-30000, 30000, 2048), rng.integers(-3, 3, 2048)]).astype(np.int32)
    _compare(loudquiet.reshape(-1, 1), 1, 16, 44100, 4096)           # wide spread of leaf parameters


def test_pathological_lpc_gain_chunked_saturation_path():
    """smooth signal then full-scale alternation: the LPC residual exceeds 2^27 (mode 2 of the Rice search,
    src/rice.rs:75-98 chunked saturating accumulation) and wraps in 32 bits (src/lpc.rs:375-389)"""
    t = np.arange(4096)
    smooth = (np.sin(t / 300.0) * 8000000).astype(np.int32)
    smooth[3000:] = np.where(np.arange(1096) % 2 == 0, 8388607, -8388608)
    _compare(smooth.reshape(-1, 1), 1, 24, 96000, 4096, lpc_order=24)
    _compare(np.stack([smooth, smooth[::-1]], axis=1), 2, 24, 96000, 4096, lpc_order=12, prc_max_parameter=5)


def test_fused_kernel_is_the_path_taken_and_falls_back_only_when_it_must():
    """Default config: every frame goes through the fused kernel, also tails whose finest Rice partitions are not
    a multiple of 4 samples (the 2728-sample tail: 8 x 341; the kernels' ODD instances).  Residuals >= 2^27 (the reference's chunked saturating
    sums become order dependent) and saturated table minima are handed to the generic kernels; BitCount order
    selection (one Rice search per fixed order, /root/reference/src/coding.rs:241-262) also runs fused; frames that do
    not fit shared memory never enter the fused kernel."""
    E.fused_counts()
    x = sigen.noisy_sine_pcm(4096 * 3 + 2728, 2, 16, 44100)
    _compare(x, 2, 16, 44100, 4096)
    assert E.fused_counts() == [8, 0]  # (16-bit stereo runs the fused path twice: PCM pairs and planes in the packer)
    _compare(x[: 4096 * 3 + 2048], 2, 16, 44100, 4096)
    assert E.fused_counts() == [8, 0]
    E.mode_counts()
    _compare(crafted_huge_residual_stereo(), 2, 24, 96000, 4096, lpc_order=24)
    assert E.fused_counts() == [0, 1]
    assert E.mode_counts()[2] > 0  # the generic kernels replayed the chunked saturating sums (mode 2)
    t = np.arange(4096)
    smooth = (np.sin(t / 300.0) * 8000000).astype(np.int32)
    smooth[3000:] = np.where(np.arange(1096) % 2 == 0, 8388607, -8388608)
    _compare(np.stack([smooth, smooth[::-1]], axis=1), 2, 24, 96000, 4096, lpc_order=12, prc_max_parameter=5)
    assert E.fused_counts() == [0, 1]  # saturated table minimum (max_parameter too small for the residuals)
    _compare(x, 2, 16, 44100, 4096, fixed_order_sel=0)
    assert E.fused_counts() == [8, 0]
    y = sigen.noisy_sine_pcm(32767 + 500, 8, 24, 96000, config_id=7)
    _compare(y, 8, 24, 96000, 32767)
    assert E.fused_counts() == [0, 0]


def test_bitcount_order_selection_on_the_fused_path():
    """OrderSel::BitCount (/root/reference/src/coding.rs:241-262: bps * order + code_bits of an exact Rice search per
    order, first minimum wins, accepted below the verbatim size) in the plan kernel: signals whose best order is each of
    0..4, max orders below 4, no LPC, noise (verbatim), a constant channel, RICE2 parameters, and the reference's KAT
    shapes ([255..], [256..], [128..] picks order 0, src/coding.rs:945-979)"""
    E.fused_counts()
    t = np.arange(2048, dtype=np.float64)
    rng = np.random.default_rng(21)
    sigs = [
        rng.integers(-20000, 20000, 2048),                                             # order 0 / verbatim territory
        np.cumsum(rng.integers(-300, 300, 2048)),                                      # order 1
        np.cumsum(np.cumsum(rng.integers(-3, 4, 2048))) % 30000,                       # order 2
        (20000 * np.sin(t / 40.0)).astype(np.int64) + rng.integers(-2, 3, 2048),       # higher orders
        (3000 * np.sin(t / 9.0) + 9000 * np.sin(t / 100.0)).astype(np.int64),
    ]
    for s in sigs:
        x = np.stack([s, s[::-1]], axis=1).astype(np.int32)
        _compare(x, 2, 16, 44100, 1024, fixed_order_sel=0)
        _compare(x[:, :1], 1, 16, 44100, 2048, fixed_order_sel=0, use_lpc=0)
    for mo in (0, 1, 2, 3):
        _compare(np.stack([sigs[3], sigs[4]], axis=1).astype(np.int32), 2, 16, 44100, 1024, fixed_order_sel=0, fixed_max_order=mo)
    _compare(np.stack([sigs[2], np.full(2048, 77)], axis=1).astype(np.int32), 2, 16, 44100, 512, fixed_order_sel=0)
    loud = (rng.integers(-(1 << 22), 1 << 22, (1024, 2))).astype(np.int32)           # Rice parameters above 14
    _compare(loud, 2, 24, 96000, 512, fixed_order_sel=0)
    for v in (255, 256, 128):
        k = np.zeros((64, 1), np.int32)
        k[::2] = v
        _compare(np.tile(k, (4, 1)), 1, 16, 44100, 256, fixed_order_sel=0, use_lpc=0)
    fused, fb = E.fused_counts()
    # (the loud 24-bit noise is handed over: its order-4 differences reach 2^26, the zigzag values 2^27)
    assert fused >= 50 and fb <= 2


def test_ext_lpc_order_search_extension():
    """EXTENSION beyond the reference (config.ext_lpc_order_search, opt-in): besides lpc_order the lower orders
    P - i * ceil(P / (k + 1)) are tried, each the Levinson solution of that order on the same autocorrelation; the LPC
    candidate with the fewest bits wins.  Emulated kernels (fused and generic) byte-equal to the oracle's statement of it;
    the streams decode to the input (inside _compare) and are never larger than the reference-compatible ones."""
    assert O.ext_lpc_orders(10, 1) == [10, 5] and O.ext_lpc_orders(10, 4) == [10, 8, 6, 4, 2]
    assert O.ext_lpc_orders(24, 8) == [24, 21, 18, 15, 12, 9, 6, 3] and O.ext_lpc_orders(3, 8) == [3, 2, 1]
    assert O.ext_lpc_orders(1, 3) == [1]
    rng = np.random.default_rng(33)
    t = np.arange(3000, dtype=np.float64)
    ar2 = np.zeros(3000)
    e = rng.normal(0, 200, 3000)
    for i in range(2, 3000):
        ar2[i] = 1.027 * ar2[i - 1] - 0.9025 * ar2[i - 2] + 1.5 * e[i]  # a resonant order-2 process: high orders only cost bits
    sigs = [
        np.stack([ar2, np.roll(ar2, 7)], axis=1),
        np.stack([8000 * np.sin(t / 15.0) + rng.normal(0, 30, 3000), 3000 * np.sin(t / 4.0) + rng.normal(0, 300, 3000)], axis=1),
        sigen.noisy_sine_pcm(3000, 2, 16, 44100, config_id=2),
    ]
    for x in sigs:
        x = np.clip(np.round(x), -32768, 32767).astype(np.int32)
        base, _ = O.encode_frames(O.default_config(), x, 2, 16, 44100, 1024)
        for k in (1, 2, 4, 8):
            _compare(x, 2, 16, 44100, 1024, ext_lpc_order_search=k)
            ext, _ = O.encode_frames(O.default_config(ext_lpc_order_search=k), x, 2, 16, 44100, 1024)
            assert len(ext) <= len(base)
    x = np.clip(np.round(sigs[0]), -32768, 32767).astype(np.int32)
    small, _ = O.encode_frames(O.default_config(ext_lpc_order_search=4), x, 2, 16, 44100, 1024)
    base, _ = O.encode_frames(O.default_config(), x, 2, 16, 44100, 1024)
    assert len(small) < len(base)   # the order-2 process really is cheaper at a lower order
    # with other orders / precisions / selectors / formats
    _compare(x, 2, 16, 44100, 1024, ext_lpc_order_search=8, lpc_order=24)
    _compare(x, 2, 16, 44100, 1000, ext_lpc_order_search=3, lpc_order=3, quant_precision=7)
    _compare(x[:, :1], 1, 16, 44100, 512, ext_lpc_order_search=2, lpc_order=1)
    _compare(x, 2, 16, 44100, 1024, ext_lpc_order_search=2, fixed_order_sel=0)
    _compare(x, 2, 16, 44100, 1024, ext_lpc_order_search=2, use_fixed=0, window_type=0)
    y = sigen.noisy_sine_pcm(1152 * 2 + 100, 3, 24, 96000, config_id=3)
    _compare(y, 3, 24, 96000, 1152, ext_lpc_order_search=4, lpc_order=16)
    _compare(np.zeros((600, 2), np.int32), 2, 16, 44100, 256, ext_lpc_order_search=2, use_constant=0)
    # second extension: the order-P coefficients quantised with fewer bits (config.ext_lpc_precision_search), alone and
    # together with the order search
    base, _ = O.encode_frames(O.default_config(), x, 2, 16, 44100, 1024)
    for ko, kp in ((0, 1), (0, 4), (4, 4), (2, 3), (5, 3)):
        _compare(x, 2, 16, 44100, 1024, ext_lpc_order_search=ko, ext_lpc_precision_search=kp)
        ext, _ = O.encode_frames(O.default_config(ext_lpc_order_search=ko, ext_lpc_precision_search=kp), x, 2, 16, 44100, 1024)
        assert len(ext) <= len(base)
    _compare(x, 2, 16, 44100, 1000, ext_lpc_precision_search=4, quant_precision=3, lpc_order=5)   # precisions 2 and 1 only
    _compare(x, 2, 16, 44100, 1024, ext_lpc_precision_search=2, fixed_order_sel=0, lpc_order=24)
    _compare(y, 3, 24, 96000, 1152, ext_lpc_order_search=3, ext_lpc_precision_search=2, lpc_order=16)
    assert E.lib().fbemu_config_verify(E.default_config(ext_lpc_precision_search=5)) != 0
    assert E.lib().fbemu_config_verify(E.default_config(ext_lpc_order_search=6, ext_lpc_precision_search=3)) != 0
    assert E.lib().fbemu_config_verify(E.default_config(ext_lpc_precision_search=1, use_direct_mse=1)) != 0
    # the estimators of the `experimental` feature have no lower orders to offer: rejected
    bad = E.default_config(ext_lpc_order_search=1, use_direct_mse=1)
    assert E.lib().fbemu_config_verify(bad) != 0
    assert E.lib().fbemu_config_verify(E.default_config(ext_lpc_order_search=9)) != 0


def test_large_blocks_with_256_finest_partitions():
    """block sizes 30720 .. 32512 in steps of 256 have 2^8 finest Rice partitions AND nearly the maximum length: the generic
    rice kernel's shared-memory layout only fits them without its bank padding (found by the fuzz campaign: 31744)"""
    rng = np.random.default_rng(77)
    for block in (31744, 32512, 30720):
        x = (rng.normal(0, 2000, (block + 300, 2)).cumsum(axis=0) % 60000 - 30000).astype(np.int32)
        _compare(x, 2, 16, 44100, block)
    x4 = (rng.normal(0, 200000, (31744, 4))).astype(np.int32)
    _compare(x4, 4, 24, 44100, 31744, lpc_order=6, quant_precision=10, fixed_order_sel=0, prc_max_parameter=6, use_midside=0)


def test_crc8_closed_form_matches_bit_serial_definition():
    """CRC-8/SMBUS (poly 0x07, init 0): the kernels' closed-form byte step against the bit-serial definition, for every
    byte value and for random strings"""
    import ctypes as C

    def ref(data):
        c = 0
        for b in data:
            c ^= b
            for _ in range(8):
                c = ((c << 1) ^ 0x07) & 0xFF if c & 0x80 else (c << 1) & 0xFF
        return c
    L = E.lib()
    L.fbemu_crc8.argtypes = [C.c_char_p, C.c_int]
    L.fbemu_crc8.restype = C.c_int
    for v in range(256):
        assert L.fbemu_crc8(bytes([v]), 1) == ref([v])
    rng = np.random.default_rng(5)
    for n in (2, 5, 8, 16):
        for _ in range(50):
            d = bytes(rng.integers(0, 256, n, dtype=np.uint8))
            assert L.fbemu_crc8(d, n) == ref(d)
    assert L.fbemu_crc8(bytes([0xFF, 0xF9, 0x10, 0x10, 0x00]), 5) == 0b01101001  # src/component/bitrepr.rs:645-666


def test_launch_geometry_helpers():
    """host/device shared launch geometry: a shorter last frame starts a warp of its own in the analysis launch, the
    fused kernels' ODD instances are chosen from the frame geometry, the staged rows cover what a warp can span"""
    # (channels, block, samples) -> expectations
    frames, slots, full, odd, rows = E.launch_geometry(2, 16, 44100, 4096, 4096 * 107 + 2728)
    assert (frames, full, slots, odd, rows) == (108, 107 * 4, ((107 * 4 + 31) // 32) * 32 + 4, 1, 16)  # 2728 = 8 x 341
    frames, slots, full, odd, rows = E.launch_geometry(2, 16, 44100, 4096, 4096 * 108)
    assert (frames, full, slots, odd, rows) == (108, 432, 432, 0, 16)
    frames, slots, full, odd, rows = E.launch_geometry(2, 16, 44100, 4096, 4096 * 9 + 3136)
    assert (frames, full, odd) == (10, 36, 1) and slots == 64 + 4                                  # 3136 = 32 x 98
    frames, slots, full, odd, rows = E.launch_geometry(2, 16, 44100, 1000, 7333)
    assert (frames, odd) == (8, 2)                                                                   # 1000 = 8 x 125
    frames, slots, full, odd, rows = E.launch_geometry(2, 16, 44100, 4096, 3136)
    assert (frames, slots, full, odd) == (1, 4, 4, 2)                                                # a single odd frame
    frames, slots, full, odd, rows = E.launch_geometry(1, 16, 44100, 4096, 4096 * 40 + 2048)
    assert (frames, full, slots, odd, rows) == (41, 40, 64 + 1, 0, 32)
    for ch, want in ((3, 48), (4, 32), (5, 48), (6, 48), (7, 48), (8, 32)):
        assert E.launch_geometry(ch, 16, 44100, 4096, 4096 * 3)[4] == want


def test_odd_block_size_every_frame_on_the_odd_instances():
    E.fused_counts()
    x = sigen.noisy_sine_pcm(1000 * 7 + 333, 2, 16, 44100, config_id=9)
    _compare(x, 2, 16, 44100, 1000)
    assert E.fused_counts() == [16, 0]


def test_fused_geometry_covers_every_block_size():
    """units tile every leaf exactly, are <= 112 samples and there are >= 32 of them (host-side geometry check
    through the encode of odd sizes, incl. leaves that are not a multiple of 4 samples)"""
    rng = np.random.default_rng(11)
    for n in (64, 66, 127, 128, 341 * 8, 3136, 98 * 32, 5000, 4100, 8192, 9216, 12345, 16383):
        x = (rng.normal(0, 300, n + 17).cumsum() % 20000 - 10000).astype(np.int32)
        _compare(x, 1, 16, 44100, n)


def test_first_frame_number_and_utf8_lengths():
    x = sigen.Sine(50, 0.3).noise(0.01, seed=1).to_vec_quantized(16, 600).reshape(300, 2)
    for first in (0, 127, 128, 2047, 65535, (1 << 21) - 1, (1 << 26), (1 << 31) - 3):
        _compare(x, 2, 16, 44100, 128, first_frame=first)


def test_out_of_range_sample_is_a_config_error():
    """src/coding.rs:587-593 + src/source.rs:262-275: VerifyError -> FB200_ERR_CONFIG"""
    x = np.zeros((200, 2), np.int32)
    x[77, 1] = 2048  # 12-bit range is [-2048, 2047]
    rc, _, _, _ = E.encode_interleaved(E.default_config(), pack_pcm(x, 2), 2, 200, 2, 12, 44100, 64)
    assert rc == 1
    x[77, 1] = -2048
    rc, _, _, _ = E.encode_interleaved(E.default_config(), pack_pcm(x, 2), 2, 200, 2, 12, 44100, 64)
    assert rc == 0
    # frame number must stay below 2^31
    rc, _, _, _ = E.encode_interleaved(E.default_config(), pack_pcm(x, 2), 2, 200, 2, 12, 44100, 64, (1 << 31) - 2)
    assert rc == 1


def test_config_verify_matches_oracle():
    cases = [dict(), dict(lpc_order=0), dict(lpc_order=25), dict(quant_precision=0), dict(quant_precision=16),
             dict(tukey_alpha=1.5), dict(tukey_alpha=-0.1), dict(prc_max_parameter=31), dict(use_direct_mse=1),
             dict(mae_optimization_steps=2), dict(block_size=31), dict(block_size=32768), dict(fixed_max_order=9),
             dict(window_type=0, tukey_alpha=9.0)]
    for kw in cases:
        o, e = _cfg_pair(**kw)
        assert O.lib().fo_config_verify(O.C.byref(o)) == E.lib().fbemu_config_verify(E.C.byref(e)), kw


def test_float_tier_taps_match_oracle():
    """autocorrelation / LPC floats are bit-identical to the scalar reference order (tolerance tier: 1e-5 rel)"""
    x = sigen.noisy_sine_pcm(4096 * 2, 2, 16, 44100)
    ecfg = E.default_config()
    rc, taps, nv = E.analyze(ecfg, pack_pcm(x, 2), 2, len(x), 2, 16, 44100, 4096)
    assert rc == 0 and nv == 8
    for f in range(2):
        blk = x[f * 4096:(f + 1) * 4096]
        L, R = blk[:, 0], blk[:, 1]
        variants = [L, R, (L + R) >> 1, L - R]
        for v, sig in enumerate(variants):
            coefs, corr = O.lpc_from_autocorr(sig, 1, 0.4, 10)
            tp = taps[f * 4 + v]
            got_corr = np.array(tp.autocorr[:11])
            got_lpc = np.array(tp.lpc[:10])
            np.testing.assert_allclose(got_corr, corr, rtol=1e-5)
            np.testing.assert_allclose(got_lpc, coefs, rtol=1e-5, atol=1e-12)
            assert np.array_equal(got_corr, corr) and np.array_equal(got_lpc, coefs)  # in fact bit-identical
            q, order, shift = O.quantize_parameters(coefs, 15)
            assert tp.qlp_order == order and tp.qlp_shift == shift
            assert list(tp.qlp[:order]) == q.tolist()
            e5 = O.fixed_lpc_errors(sig)
            bps_v = 17 if v == 3 else 16
            est = [O.estimate_entropy(e5[k], k, 16) + bps_v * k for k in range(5)]
            assert list(tp.fixed_est_bits) == est


@pytest.mark.parametrize("seed", range(60))
def test_randomised_formats_signals_and_configs(seed):
    """seeded fuzz over channel counts, sample sizes, block sizes (incl. odd tails), signal classes and encoder options:
    the kernel bodies (fused and generic paths) must reproduce the oracle's frame bytes, which must decode losslessly"""
    rng = np.random.default_rng(1000 + seed)
    x, channels, bps, rate, block, first, cfg = random_case(rng)
    _compare(x, channels, bps, rate, block, first_frame=first, **cfg)


def test_direct_mse_estimator_matches_oracle():
    """K1C (the `experimental` covariance-method LPC, /root/reference/src/lpc.rs:852-913) under emulation, frames spanning
    several staging tiles; incl. signals whose covariance matrix is not positive definite"""
    x = sigen.noisy_sine_pcm(1024 * 3 + 300, 2, 16, 44100, config_id=4)
    _compare(x, 2, 16, 44100, 1024, use_direct_mse=1, window_type=0)
    _compare(x[:2048], 2, 16, 44100, 1024, use_direct_mse=1)
    _compare(x[:1500, 0], 1, 16, 44100, 500, use_direct_mse=1, lpc_order=24, tukey_alpha=0.1)
    _compare(x[:1100, 1], 1, 16, 44100, 1000, use_direct_mse=1, lpc_order=1, quant_precision=5)
    _compare(np.zeros((600, 2), np.int32), 2, 16, 44100, 256, use_direct_mse=1, window_type=0)
    c = np.full((600, 1), 1000, np.int32)
    c[300:] = -77
    _compare(c, 1, 16, 44100, 256, use_direct_mse=1, window_type=0, use_constant=0)
    y = sigen.noisy_sine_pcm(1152 + 100, 2, 24, 96000, config_id=3)
    _compare(y, 2, 24, 96000, 1152, use_direct_mse=1, window_type=0, lpc_order=16)
    # float tier: coefficients are bit-identical to the oracle's
    sig = x[:1024, 0]
    coefs, corr, _ = O.lpc_with_direct_mse(sig, 0, 0.0, 10)
    cfg = E.default_config(use_direct_mse=1, window_type=0)
    rc, taps, nv = E.analyze(cfg, pack_pcm(sig.reshape(-1, 1), 2), 2, 1024, 1, 16, 44100, 1024)
    assert rc == 0 and nv == 1
    assert np.array_equal(np.array(taps[0].lpc[:10]), coefs) and np.array_equal(np.array(taps[0].autocorr[:11]), corr)


def test_irls_weight_matches_libm_exhaustive():
    """fb_irls_weight -- the device code of the IRLS weight (/root/reference/src/lpc.rs:828: max, divide, max, then
    powf(-1.2) through fb_powf_pos, a restatement of glibc's powf) compiled for the CPU -- against the oracle's, which
    calls the host libm, for EVERY non-negative raw-error bit pattern (zero .. +inf and the first NaNs) under one
    normalizer, and slices of it under others (incl. the normalizer 0 of a silent frame and negative errors)"""
    L = E.lib()
    threads = min(16, os.cpu_count() or 1)

    def sweep(first, end, norm):
        step = 1 << 26
        while first < end:
            cnt = min(step, end - first)
            got = np.empty(cnt, np.uint32)
            L.fbemu_irls_weight_bits(first, cnt, norm, threads, got.ctypes.data)
            want = O.irls_weight_bits(first, cnt, norm, threads)
            bad = np.flatnonzero(got != want)
            assert len(bad) == 0, (norm, hex(first + int(bad[0])), hex(int(got[bad[0]])), hex(int(want[bad[0]])))
            first += cnt

    sweep(0, 0x7F800000 + 4096, 100.0)          # weights from 0.01^-1.2 down to 0
    for norm in (1.0, 3.0, 32767.0, 8388607.0, 0.0):
        sweep(0x3F000000, 0x4C000000, norm)     # |err| in [0.5, 2^25): what 24-bit audio can produce
    sweep(0xBF000000, 0xC1000000, 1000.0)       # negative errors


def test_irls_mae_estimator_matches_oracle():
    """K1I (the `experimental` IRLS-MAE refinement, /root/reference/src/lpc.rs:814-850) under emulation, frames spanning
    several staging tiles, orders and step counts; whole frames byte-equal and coefficients bit-equal to the oracle"""
    x = sigen.noisy_sine_pcm(1024 * 3 + 300, 2, 16, 44100, config_id=4)
    _compare(x, 2, 16, 44100, 1024, use_direct_mse=1, mae_optimization_steps=2, window_type=0)
    _compare(x[:2048], 2, 16, 44100, 1024, use_direct_mse=1, mae_optimization_steps=1)
    _compare(x[:1500, 0], 1, 16, 44100, 500, use_direct_mse=1, mae_optimization_steps=3, lpc_order=24, tukey_alpha=0.1)
    _compare(x[:1100, 1], 1, 16, 44100, 1000, use_direct_mse=1, mae_optimization_steps=5, lpc_order=1, quant_precision=5)
    _compare(np.zeros((600, 2), np.int32), 2, 16, 44100, 256, use_direct_mse=1, mae_optimization_steps=2, window_type=0)
    c = np.full((600, 1), 1000, np.int32)
    c[300:] = -77
    _compare(c, 1, 16, 44100, 256, use_direct_mse=1, mae_optimization_steps=2, window_type=0, use_constant=0)
    y = sigen.noisy_sine_pcm(1152 + 100, 2, 24, 96000, config_id=3)
    _compare(y, 2, 24, 96000, 1152, use_direct_mse=1, mae_optimization_steps=2, window_type=0, lpc_order=16)
    # the steps are ignored unless the direct-MSE estimator is selected (src/coding.rs:337-351)
    _compare(x[:2048], 2, 16, 44100, 1024, use_direct_mse=0, mae_optimization_steps=3)
    # float tier: the winning coefficients are bit-identical to the oracle's, and the refinement does change them
    sig = x[:1024, 0]
    cfg = E.default_config(use_direct_mse=1, mae_optimization_steps=4, window_type=0)
    rc, taps, nv = E.analyze(cfg, pack_pcm(sig.reshape(-1, 1), 2), 2, 1024, 1, 16, 44100, 1024)
    assert rc == 0 and nv == 1
    coefs, sums = O.lpc_with_irls_mae(sig, 0, 0.0, 10, 4)
    plain, corr, _ = O.lpc_with_direct_mse(sig, 0, 0.0, 10)
    assert np.array_equal(np.array(taps[0].lpc[:10]), coefs) and np.array_equal(np.array(taps[0].autocorr[:11]), corr)
    assert not np.array_equal(coefs, plain) and sums.min() < sums[0]


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_frames_at_reference_scale_emu(seed):
    """the reference's fuzz shape (one frame, block up to 32767, 1..8 channels, 8..24 bits, all toggles) under emulation;
    the -m gpu suite runs 60 seeds of it on the device"""
    rng = np.random.default_rng(9000 + seed)
    x, channels, bps, rate, block, cfg = fuzz_frame_case(rng)
    _compare(x, channels, bps, rate, block, **cfg)
