"""Names and result files of the reference's bench tool, so that numbers measured here can be merged with
`docs/benchmark_results/*.csv`:

* codec strings `pco:level=8:delta=Consecutive@1:mode=FloatMult@0.01:chunk-n=262144` - parsed like
  pco_cli/src/bench/codecs/mod.rs:241-292 (name, then key=value pairs) with the spec grammars of pco_cli/src/parse.rs:8-48
  and the defaults of pco_cli/src/chunk_config_opt.rs:8-25; printed like `Display for CodecConfig`
  (codecs/mod.rs:137-148, :294-303): only the settings that differ from the defaults, in the order level, delta, mode, chunk-n;
* the results CSV `input,codec,compress_dt,decompress_dt,compressed_size,uncompressed_size` (seconds as f32), one line per
  (input, codec), sorted, later runs replacing earlier ones (pco_cli/src/bench/mod.rs:325-372).

Host-side text handling only - nothing here touches the codec.
"""
import os

import numpy as np

from ._lib import ChunkConfig, DeltaSpec, ModeSpec, PagingSpec

DEFAULT_MAX_PAGE_N = 1 << 18  # pco::DEFAULT_MAX_PAGE_N (pco/src/constants.rs)
CSV_HEADER = "input,codec,compress_dt,decompress_dt,compressed_size,uncompressed_size"


def parse_delta_spec(s):  # pco_cli/src/parse.rs:8-26
    low = s.lower()
    if low == "auto":
        return DeltaSpec.auto()
    if low == "noop":
        return DeltaSpec.no_op()
    if low == "lookback":
        return DeltaSpec.try_lookback()
    name, sep, value = low.partition("@")
    if not sep or name not in ("consecutive", "conv1"):
        raise ValueError(f"invalid delta spec: {s}")
    return DeltaSpec.try_consecutive(int(value)) if name == "consecutive" else DeltaSpec.try_conv1(int(value))


def parse_mode_spec(s):  # pco_cli/src/parse.rs:28-48
    low = s.lower()
    if low == "auto":
        return ModeSpec.auto()
    if low == "classic":
        return ModeSpec.classic()
    if low == "dict":
        return ModeSpec.try_dict()
    name, sep, value = low.partition("@")
    if not sep or name not in ("floatmult", "floatquant", "intmult"):
        raise ValueError(f"invalid mode spec: {s}")
    if name == "floatmult":
        return ModeSpec.try_float_mult(float(value))
    return ModeSpec.try_float_quant(int(value)) if name == "floatquant" else ModeSpec.try_int_mult(int(value))


def _rust_f64(x):
    """How Rust's `{}` prints an f64: shortest round-trip digits, never an exponent, integers without a fraction."""
    r = repr(float(x))
    if "e" in r or "E" in r:
        r = np.format_float_positional(float(x), trim="-")
    return r[:-2] if r.endswith(".0") else r


def unparse_delta_spec(spec):  # pco_cli/src/bench/codecs/pco.rs:7-16
    return {0: "Auto", 1: "NoOp", 2: f"Consecutive@{spec.order}", 3: "Lookback", 4: f"Conv1@{spec.order}"}.get(spec.kind, "Unknown")


def unparse_mode_spec(spec):  # pco_cli/src/bench/codecs/pco.rs:18-28
    return {0: "Auto", 1: "Classic", 2: f"FloatMult@{_rust_f64(spec.base)}", 3: f"FloatQuant@{spec.k}", 4: f"IntMult@{spec.int_base}",
            5: "Dict"}.get(spec.kind, "Unknown")


class PcoCodec:
    """The `pco` codec entry of the reference's bench (ChunkConfigOpt, pco_cli/src/chunk_config_opt.rs:8-37)."""

    def __init__(self, level=8, delta=None, mode=None, chunk_n=DEFAULT_MAX_PAGE_N):
        self.level, self.delta, self.mode, self.chunk_n = int(level), delta or DeltaSpec.auto(), mode or ModeSpec.auto(), int(chunk_n)

    def confs(self):  # codecs/pco.rs:35-42
        return [("level", str(self.level)), ("delta", unparse_delta_spec(self.delta)), ("mode", unparse_mode_spec(self.mode)), ("chunk-n", str(self.chunk_n))]

    def name(self, explicit=False):
        default = dict(PcoCodec().confs())
        return "pco" + "".join(f":{k}={v}" for k, v in self.confs() if explicit or v != default[k])

    __str__ = name

    def chunk_config(self, enable_8_bit=True):  # chunk_config_opt.rs:28-36 (the bench passes enable_8_bit = true, codecs/pco.rs:45-47)
        return ChunkConfig(compression_level=self.level, mode_spec=self.mode, delta_spec=self.delta,
                           paging_spec=PagingSpec.equal_pages_up_to(self.chunk_n), enable_8_bit=enable_8_bit)


def parse_codec(s):
    parts = s.split(":")
    if parts[0] not in ("pco", "pcodec"):
        raise ValueError(f"Unknown codec: {parts[0]}")
    kw = {}
    for part in parts[1:]:
        kv = part.split("=")
        if len(kv) != 2:
            raise ValueError(f"codec config {part} is not a key=value pair")
        k, v = kv
        if k == "level":
            kw["level"] = int(v)
        elif k == "delta":
            kw["delta"] = parse_delta_spec(v)
        elif k == "mode":
            kw["mode"] = parse_mode_spec(v)
        elif k in ("chunk-n", "chunk_n"):
            kw["chunk_n"] = int(v)
        else:
            raise ValueError(f"unexpected argument --{k} for codec pco")
    return PcoCodec(**kw)


def _f32_seconds(x):
    """Duration::as_secs_f32 printed with `{}`: the shortest digits that round-trip the f32, no exponent."""
    return np.format_float_positional(np.float32(x), unique=True, trim="-")


def merge_results_csv(path, rows):
    """rows: dicts with input, codec, compress_dt, decompress_dt (seconds), compressed_size, uncompressed_size.  Lines already in
    `path` are kept unless a row has the same (input, codec); output is sorted by that key (pco_cli/src/bench/mod.rs:325-372)."""
    lines = {}
    if os.path.exists(path):
        with open(path) as f:
            for i, line in enumerate(f.read().split("\n")):
                if i == 0 or not line.strip():
                    continue
                fields = line.split(",")
                lines[(fields[0], fields[1])] = fields[2:6]
    for r in rows:
        lines[(str(r["input"]), str(r["codec"]))] = [_f32_seconds(r["compress_dt"]), _f32_seconds(r["decompress_dt"]), str(int(r["compressed_size"])),
                                                     str(int(r["uncompressed_size"]))]
    out = [CSV_HEADER] + [",".join([k[0], k[1]] + v) for k, v in sorted(lines.items())]
    with open(path, "w") as f:
        f.write("\n".join(out))
    return len(out) - 1
