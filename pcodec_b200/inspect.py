"""`pco inspect` for standalone files (pco_cli/src/inspect/mod.rs:26-234, summary.rs): format version, number type, per-chunk
mode / delta encoding and per latent variable n_bins, ans_size_log, the bins and the approximate bits per latent - the quantities
SURVEY.md 8(d) asks to record beside every throughput number.

Host-side metadata reading only (docs/format.md; pco/src/metadata/{chunk,chunk_latent_var,mode,delta_encoding,bin}.rs,
pco/src/standalone/{compressor.rs:85-105,decompressor.rs:85-148}): nothing here decodes numbers.  A chunk's page length is not
stored in the file; pass `chunk_offsets` (from the side index, pcodec_b200.standalone.build_index) for large files, otherwise
the page's tANS stream is walked here in Python (symbols only, ~0.5 s per 2^18-number chunk) to find where the next chunk starts.
"""
import math

import numpy as np

MODE_NAMES = ("Classic", "IntMult", "FloatMult", "FloatQuant", "Dict")
DELTA_NAMES = ("NoOp", "Consecutive", "Lookback", "Conv1")
DTYPE_NAMES = {1: "u32", 2: "u64", 3: "i32", 4: "i64", 5: "f32", 6: "f64", 7: "u16", 8: "i16", 9: "f16", 10: "u8", 11: "i8"}  # pco_c/include/cpcodec.h:10-20
DTYPE_BITS = {1: 32, 2: 64, 3: 32, 4: 64, 5: 32, 6: 64, 7: 16, 8: 16, 9: 16, 10: 8, 11: 8}


class InspectError(ValueError):
    pass


class _Bits:
    """Little-endian bit cursor over bytes (pco/src/bit_reader.rs); reads past the end fail."""

    def __init__(self, buf, byte_pos=0):
        self.buf, self.bit = buf, 8 * byte_pos

    def uint(self, n):
        if n == 0:
            return 0
        lo, hi = self.bit // 8, (self.bit + n + 7) // 8
        if hi > len(self.buf):
            raise InspectError("InsufficientData: metadata runs past the end of the file")
        v = (int.from_bytes(self.buf[lo:hi], "little") >> (self.bit % 8)) & ((1 << n) - 1)
        self.bit += n
        return v

    def align(self):
        self.bit = (self.bit + 7) // 8 * 8

    @property
    def byte(self):
        return (self.bit + 7) // 8


def _offset_bits_bits(latent_bits):  # pco/src/metadata/bin.rs:20-22: log2(L::BITS) + 1
    return int(math.log2(latent_bits)) + 1


def _read_var(r, latent_bits):  # pco/src/metadata/chunk_latent_var.rs:22-53, :102-143
    ans_size_log = r.uint(4)
    n_bins = r.uint(15)
    obb = _offset_bits_bits(latent_bits)
    bins = [(r.uint(ans_size_log) + 1, r.uint(latent_bits), r.uint(obb)) for _ in range(n_bins)]
    return {"latent_bits": latent_bits, "ans_size_log": ans_size_log, "bins": bins}


def read_chunk_meta(buf, byte_pos, number_bits, format_major=4):
    """ChunkMeta::read_from (pco/src/metadata/chunk.rs:127-174) at `byte_pos`; returns (meta dict, byte position after it)."""
    r = _Bits(buf, byte_pos)
    mode = {"kind": r.uint(4)}  # metadata/mode.rs:102-167
    if mode["kind"] >= len(MODE_NAMES):
        raise InspectError(f"Corruption: unknown mode variant {mode['kind']}")
    if mode["kind"] in (1, 2):
        mode["base_latent"] = r.uint(number_bits)
    elif mode["kind"] == 3:
        mode["k"] = r.uint(8)
    elif mode["kind"] == 4:
        n_unique = r.uint(25)
        r.align()
        mode["dict"] = [r.uint(number_bits) for _ in range(n_unique)]
    delta = {"kind": 0, "order": 0, "secondary_uses_delta": False}  # metadata/delta_encoding.rs:118-202
    if format_major < 3:  # before delta variants: a bare 3-bit consecutive order
        delta["order"] = r.uint(3)
        delta["kind"] = 1 if delta["order"] else 0
    else:
        delta["kind"] = r.uint(4)
        if delta["kind"] == 1:
            delta["order"] = r.uint(3)
            delta["secondary_uses_delta"] = bool(r.uint(1))
        elif delta["kind"] == 2:
            delta["window_n_log"] = 1 + r.uint(5)
            delta["state_n_log"] = r.uint(4)
            delta["secondary_uses_delta"] = bool(r.uint(1))
        elif delta["kind"] == 3:
            delta["quantization"] = r.uint(5)
            bias = r.uint(64) ^ (1 << 63)  # i64::from_latent_ordered
            delta["bias"] = bias - (1 << 64) if bias >> 63 else bias
            delta["weights"] = []
            for _ in range(1 + r.uint(5)):
                w = r.uint(32) ^ 0x80000000
                delta["weights"].append(w - (1 << 32) if w >> 31 else w)
        elif delta["kind"] != 0:
            raise InspectError(f"Corruption: unknown delta encoding value {delta['kind']}")
    meta = {"mode": mode, "delta": delta, "vars": {}}
    if delta["kind"] == 2:
        meta["vars"]["delta"] = _read_var(r, 32)
    meta["vars"]["primary"] = _read_var(r, 32 if mode["kind"] == 4 else number_bits)
    if mode["kind"] in (1, 2, 3):
        meta["vars"]["secondary"] = _read_var(r, number_bits)
    r.align()
    return meta, r.byte


def _ordered_to_float(v, bits):
    mid = 1 << (bits - 1)
    raw = v ^ mid if v & mid else (~v) & ((1 << bits) - 1)
    dt = {16: np.float16, 32: np.float32, 64: np.float64}[bits]
    return float(np.array([raw], dtype=f"u{bits // 8}").view(dt)[0])


def describe_mode(mode, dtype_byte):
    """The mode the way the reference's Debug print shows it, with the base in the number's own units."""
    bits, name = DTYPE_BITS[dtype_byte], MODE_NAMES[mode["kind"]]
    if mode["kind"] == 1:
        return f"IntMult({mode['base_latent']})"
    if mode["kind"] == 2:
        return f"FloatMult({_ordered_to_float(mode['base_latent'], bits)!r})"
    if mode["kind"] == 3:
        return f"FloatQuant({mode['k']})"
    if mode["kind"] == 4:
        return f"Dict({len(mode['dict'])} values)"
    return name


def describe_delta(delta):
    if delta["kind"] == 1:
        return f"Consecutive(order={delta['order']}, secondary_uses_delta={str(delta['secondary_uses_delta']).lower()})"
    if delta["kind"] == 2:
        return f"Lookback(window_n_log={delta['window_n_log']}, state_n_log={delta['state_n_log']}, secondary_uses_delta={str(delta['secondary_uses_delta']).lower()})"
    if delta["kind"] == 3:
        return f"Conv1(quantization={delta['quantization']}, bias={delta['bias']}, weights={delta['weights']})"
    return DELTA_NAMES[delta["kind"]]


def var_summary(var):
    """LatentVarSummary (pco_cli/src/inspect/mod.rs:75-123): approx_avg_bits = sum_b w_b (offset_bits_b + ans_size_log - log2 w_b) / 2^ans_size_log,
    reported here also split into its tANS and offset parts."""
    total = float(1 << var["ans_size_log"])
    ans = sum(w * (var["ans_size_log"] - math.log2(w)) for w, _, _ in var["bins"]) / total
    off = sum(w * ob for w, _, ob in var["bins"]) / total
    return {"latent_type": f"U{var['latent_bits']}", "n_bins": len(var["bins"]), "ans_size_log": var["ans_size_log"], "approx_avg_bits": ans + off,
            "approx_avg_ans_bits": ans, "approx_avg_offset_bits": off, "bins": list(var["bins"])}


# ---- finding the end of a page without an index: walk its tANS symbols ------------------------------------------------
def _decoder_nodes(var):  # pco/src/ans/spec.rs:24-59 + ans/decoding.rs:15-48
    size_log = var["ans_size_log"]
    size = 1 << size_log
    weights = [w for w, _, _ in var["bins"]] or [1]
    if sum(weights) != size:
        raise InspectError("Corruption: bin weights do not add up to the tANS table size")
    stride = (3 * size) // 5
    stride += 1 - stride % 2
    symbols = [0] * size
    step = 0
    for s, w in enumerate(weights):
        for _ in range(w):
            symbols[(stride * step) & (size - 1)] = s
            step += 1
    x_s = list(weights)
    base, nbits = [0] * size, [0] * size
    for i, s in enumerate(symbols):
        b = size_log - (x_s[s].bit_length() - 1)
        base[i], nbits[i] = (x_s[s] << b) - size, b
        x_s[s] += 1
    return symbols, base, nbits


def page_size(buf, byte_pos, meta, n):
    """Bytes of the page that starts at `byte_pos` (page meta + batches, pco/src/metadata/page.rs:36-57,
    pco/src/page_latent_decompressor.rs:89-177, pco/src/wrapped/page_decompressor.rs:115-191)."""
    delta = meta["delta"]
    # latents held in the page meta instead of the stream (metadata/delta_encoding.rs n_latents_per_state): the consecutive order,
    # the lookback state, the conv1 weights' span
    n_state = {0: 0, 1: delta["order"], 2: 1 << delta.get("state_n_log", 0), 3: len(delta.get("weights", []))}[delta["kind"]]
    r = _Bits(buf, byte_pos)
    walkers = []
    for key in ("delta", "primary", "secondary"):
        if key not in meta["vars"]:
            continue
        var = meta["vars"][key]
        uses_delta = n_state > 0 and (key == "primary" or (key == "secondary" and delta["secondary_uses_delta"]))
        if uses_delta:
            for _ in range(n_state):
                r.uint(var["latent_bits"])
        states = [r.uint(var["ans_size_log"]) for _ in range(4)]
        stored = max(0, n - n_state) if (uses_delta or key == "delta") else n  # the lookback var has one entry per stored primary latent
        walkers.append({"n": stored, "states": states, "nodes": _decoder_nodes(var),
                        "offset_bits": [ob for _, _, ob in var["bins"]] or [0], "coded": len(var["bins"]) > 1})
    r.align()
    for start in range(0, n, 256):
        for w in walkers:
            cnt = max(0, min(256, w["n"] - start))
            if cnt == 0:
                continue
            symbols, base, nbits = w["nodes"]
            ob, states = w["offset_bits"], w["states"]
            if not w["coded"]:
                r.bit += cnt * ob[0]
                continue
            total_offset_bits = 0
            for i in range(cnt):
                st = states[i & 3]
                total_offset_bits += ob[symbols[st]]
                states[i & 3] = base[st] + r.uint(nbits[st])
            r.bit += total_offset_bits
    if r.byte > len(buf):
        raise InspectError("InsufficientData: page runs past the end of the file")
    return r.byte - byte_pos


def inspect(src, chunk_offsets=None):
    """Summary of a standalone file, shaped like the reference's `pco inspect` output (pco_cli/src/inspect/summary.rs)."""
    buf = bytes(memoryview(src))
    if len(buf) < 5 or buf[:4] != b"pco!":
        raise InspectError("Corruption: magic header does not match")
    r = _Bits(buf, 4)
    standalone_version = r.uint(8)
    uniform_type = 0
    if standalone_version < 2:
        r.bit -= 8
    else:
        if standalone_version >= 3:
            uniform_type = r.uint(8)
        power = 1 + r.uint(6)
        r.uint(power)
        r.align()
    major = r.uint(8)
    minor = r.uint(8) if major >= 4 else 0
    header_size = pos = r.byte
    chunks, meta_size, page_bytes, dtype_byte = [], 0, 0, uniform_type
    while True:
        if pos >= len(buf):
            raise InspectError("InsufficientData: file ends without a terminator")
        if chunk_offsets is not None and len(chunks) < len(chunk_offsets) and int(chunk_offsets[len(chunks)]) != pos:
            raise InspectError(f"chunk_offsets[{len(chunks)}] = {int(chunk_offsets[len(chunks)])} but the chunk starts at byte {pos}")
        t = buf[pos]
        if t == 0:  # standalone/decompressor.rs: terminator
            break
        if t not in DTYPE_BITS:
            raise InspectError(f"Corruption: unknown number type byte {t}")
        dtype_byte = t
        n = int.from_bytes(buf[pos + 1:pos + 4], "little") + 1  # standalone/compressor.rs:85-105: 24 bits of n - 1
        meta, after_meta = read_chunk_meta(buf, pos + 4, DTYPE_BITS[t], major)
        meta_size += after_meta - pos
        if chunk_offsets is not None and len(chunks) + 1 < len(chunk_offsets):
            after_page = int(chunk_offsets[len(chunks) + 1])
        elif chunk_offsets is not None and len(chunks) + 1 == len(chunk_offsets):
            after_page = len(buf) - 1 if buf[-1] == 0 else None
            if after_page is None:
                after_page = after_meta + page_size(buf, after_meta, meta, n)
        else:
            after_page = after_meta + page_size(buf, after_meta, meta, n)
        page_bytes += after_page - after_meta
        chunks.append({"idx": len(chunks), "n": n, "byte_offset": pos, "meta_size": after_meta - pos, "page_size": after_page - after_meta,
                       "mode": describe_mode(meta["mode"], t), "delta_encoding": describe_delta(meta["delta"]),
                       "latent_var": {k: var_summary(v) for k, v in meta["vars"].items()}})
        pos = after_page
    n_total = sum(c["n"] for c in chunks)
    unc = n_total * DTYPE_BITS.get(dtype_byte, 0) // 8
    total = header_size + meta_size + page_bytes + 1
    return {"format_version": f"{major}.{minor}", "standalone_version": standalone_version, "number_type": DTYPE_NAMES.get(dtype_byte, "<none>"), "n": n_total,
            "n_chunks": len(chunks), "uncompressed_size": unc,
            "compressed": {"ratio": unc / total if n_total else 0.0, "total_size": total, "header_size": header_size, "meta_size": meta_size, "page_size": page_bytes,
                           "footer_size": 1, "unknown_trailing_bytes": len(buf) - pos - 1},
            "chunk": chunks}


def index_chunk_offsets(index):
    """Chunk byte offsets out of a side index (include/pco_b200.h IndexHeader / IndexChunk), for inspect(..., chunk_offsets=)."""
    ib = bytes(memoryview(index))
    n_chunks, chunks_off = int.from_bytes(ib[8:16], "little"), int.from_bytes(ib[32:40], "little")
    return [int.from_bytes(ib[chunks_off + 32 * i:chunks_off + 32 * i + 8], "little") for i in range(n_chunks)]


__all__ = ["inspect", "read_chunk_meta", "page_size", "var_summary", "describe_mode", "describe_delta", "index_chunk_offsets", "InspectError"]


if __name__ == "__main__":  # python -m pcodec_b200.inspect file.pco  (the summary as JSON; bins left out unless --bins)
    import json
    import sys

    with open(sys.argv[1], "rb") as fh:
        summary = inspect(fh.read())
    if "--bins" not in sys.argv:
        for ch in summary["chunk"]:
            for v in ch["latent_var"].values():
                v.pop("bins")
    print(json.dumps(summary, indent=1))
