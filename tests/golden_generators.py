"""Deterministic generators for the reference's golden assets.

Each function restates the generator in pco/src/tests/compatibility.rs (cited
per function) in numpy; the .pco bytes live in tests/golden/pco_assets.json.
"""
import numpy as np


def v0_0_0_classic():  # compatibility.rs:70-82
    return np.concatenate([np.arange(0, 1000), np.arange(2000, 3000)]).astype(np.int32)


def v0_0_0_delta_float_mult():  # compatibility.rs:85-97
    nums = np.arange(2000, dtype=np.float32)
    nums[1337] = np.float32(nums[1337] + np.float32(1.001))
    return nums


def v0_1_0_delta_int_mult():  # compatibility.rs:100-114
    nums = (np.arange(2000, dtype=np.int32) * 1000).astype(np.int32)
    nums[1337] -= 1
    return nums


def v0_1_1_standalone_versioned():  # compatibility.rs:117-126
    return np.zeros(0, dtype=np.float32)


def _pseudorandom_f16s():  # compatibility.rs:128-142
    num = np.float32(0.1)
    out = []
    for _ in range(2000):
        num = np.float32(np.fmod(np.float32(np.float32(num * np.float32(77.7)) + np.float32(0.1)), np.float32(2.0)))
        if num < np.float32(1.0):
            out.append(np.float16(np.float32(np.float32(-1.0) - num)))
        else:
            out.append(np.float16(num))
    return np.array(out, dtype=np.float16)


def v0_3_0_f16():  # compatibility.rs:145-155
    return _pseudorandom_f16s()


def v0_3_0_float_quant():  # compatibility.rs:157-178
    x = _pseudorandom_f16s().astype(np.float32)
    bits = x.view(np.uint32).copy()
    small = np.abs(x) < np.float32(1.1)
    bits[small] += 1
    return bits.view(np.float32)


def v0_4_0_lookback_delta():  # compatibility.rs:181-197
    base = [1121827092, 729032807, 3968137854, 2875434067, 3775328080, 431649926, 1048116090, 1906978350, 14752788, 1180462487]
    return np.array(base * 100, dtype=np.uint32)


def v0_4_5_uniform_type():  # compatibility.rs:200-222
    return np.array([1, 2, 3, 4, 5], dtype=np.uint32)


def v0_4_8_minor_version():  # compatibility.rs:225-245
    return np.array([1, 2, 3, 4, 5], dtype=np.uint32)


def v1_0_0_dict():  # compatibility.rs:248-259
    return np.array([8924659283, 234897984367, 9827358920] * 1000, dtype=np.uint64)


def v1_0_0_conv1():  # compatibility.rs:262-279
    xm1 = np.float32(0.0)
    xm2 = np.float32(0.0)
    out = []
    for i in range(2000):
        c = np.float32((i * 47) % 77 - 38)
        x = np.float32(np.float32(np.float32(xm1 * np.float32(1.99)) - xm2) + c)
        out.append(int(np.float32(x + np.float32(10000.0))))  # `as i32` truncates toward zero
        xm2 = xm1
        xm1 = x
    return np.array(out, dtype=np.int32)


def v1_0_0_u8():  # compatibility.rs:282-291
    return np.concatenate([np.arange(0, 65), np.arange(192, 256)]).astype(np.uint8)


def v1_0_0_i8():  # compatibility.rs:294-303
    return np.concatenate([np.arange(-128, -63), np.arange(64, 128)]).astype(np.int8)


GENERATORS = {
    "v0_0_0_classic": v0_0_0_classic,
    "v0_0_0_delta_float_mult": v0_0_0_delta_float_mult,
    "v0_1_0_delta_int_mult": v0_1_0_delta_int_mult,
    "v0_1_1_standalone_versioned": v0_1_1_standalone_versioned,
    "v0_3_0_f16": v0_3_0_f16,
    "v0_3_0_float_quant": v0_3_0_float_quant,
    "v0_4_0_lookback_delta": v0_4_0_lookback_delta,
    "v0_4_5_uniform_type": v0_4_5_uniform_type,
    "v0_4_8_minor_version": v0_4_8_minor_version,
    "v1_0_0_dict": v1_0_0_dict,
    "v1_0_0_conv1": v1_0_0_conv1,
    "v1_0_0_u8": v1_0_0_u8,
    "v1_0_0_i8": v1_0_0_i8,
}


def load_assets():
    import hashlib
    import json
    import os

    path = os.path.join(os.path.dirname(__file__), "golden", "pco_assets.json")
    with open(path) as f:
        raw = json.load(f)
    out = {}
    for name, ent in raw.items():
        data = bytes.fromhex(ent["hex"])
        assert hashlib.sha256(data).hexdigest() == ent["sha256"]
        out[name] = data
    return out


def bits_view(a):
    """Raw bit patterns (equality on to_latent_ordered == equality on bits)."""
    return a.view({1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[a.dtype.itemsize])
