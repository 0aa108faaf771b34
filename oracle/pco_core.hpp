// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of pcodec/pcodec (pco crate v1.0.3) for the per-chunk
// encode/decode hot path.  Written from scratch in C++17 from a reading of the
// reference; every function cites the reference file:line it follows (paths
// relative to /root/reference/).  Nothing here is shipped: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may build or call it.  The product (pcodec_b200/) never links this code.
//
// Parity pinning: this oracle decodes all 13 golden assets in pco/assets to
// the generators in pco/src/tests/compatibility.rs, reproduces the reference's
// unit known-answer tests, and re-encodes the format-4.1 assets byte for byte
// (see tests/test_oracle_golden.py).  The reference itself (Rust) cannot be
// compiled in this image (no cargo/rustc), so there is no oracle/_ref.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace pco_oracle {

// pco/src/constants.rs:1-62
using Bitlen = uint32_t;
using Weight = uint32_t;
using AnsState = uint32_t;
using Symbol = uint32_t;
using DeltaLookback = uint32_t;

constexpr Bitlen BITS_TO_ENCODE_ANS_SIZE_LOG = 4;
constexpr Bitlen BITS_TO_ENCODE_MODE_VARIANT = 4;
constexpr Bitlen BITS_TO_ENCODE_DELTA_ENCODING_VARIANT = 4;
constexpr Bitlen BITS_TO_ENCODE_DELTA_ENCODING_ORDER = 3;
constexpr Bitlen BITS_TO_ENCODE_DELTA_CONV_QUANTIZATION = 5;
constexpr Bitlen BITS_TO_ENCODE_DELTA_CONV_N_WEIGHTS = 5;
constexpr Bitlen BITS_TO_ENCODE_DELTA_LOOKBACK_WINDOW_N_LOG = 5;
constexpr Bitlen BITS_TO_ENCODE_DELTA_LOOKBACK_STATE_N_LOG = 4;
constexpr Bitlen BITS_TO_ENCODE_N_BINS = 15;
constexpr Bitlen BITS_TO_ENCODE_QUANTIZE_K = 8;
constexpr Bitlen BITS_TO_ENCODE_DICT_LEN = 25;
constexpr size_t OVERSHOOT_PADDING = 9;
constexpr Bitlen MAX_ANS_BITS = 14;
constexpr Bitlen LIMITED_UNOPTIMIZED_BINS_LOG = 6;
constexpr size_t MAX_COMPRESSION_LEVEL = 12;
constexpr size_t MAX_CONSECUTIVE_DELTA_ORDER = 7;
constexpr size_t MAX_CONV1_DELTA_ORDER = 32;
constexpr Bitlen MAX_CONV1_DELTA_QUANTIZATION = (1u << BITS_TO_ENCODE_DELTA_CONV_QUANTIZATION) - 1;
constexpr size_t MAX_ENTRIES = size_t(1) << 24;
constexpr Bitlen MAX_DELTA_LOOKBACK_WINDOW_N_LOG = 24;
constexpr size_t DEFAULT_COMPRESSION_LEVEL = 8;
constexpr size_t DEFAULT_MAX_PAGE_N = size_t(1) << 18;
constexpr size_t ANS_INTERLEAVING = 4;
constexpr size_t FULL_BATCH_N = 256;

// pco/src/standalone/constants.rs:4-9
constexpr uint8_t MAGIC_HEADER[4] = {112, 99, 111, 33};
constexpr uint8_t MAGIC_TERMINATION_BYTE = 0;
constexpr Bitlen BITS_TO_ENCODE_N_ENTRIES = 24;
constexpr Bitlen BITS_TO_ENCODE_STANDALONE_VERSION = 8;
constexpr Bitlen BITS_TO_ENCODE_VARINT_POWER = 6;
constexpr size_t CURRENT_STANDALONE_VERSION = 3;

// Every decode entry point copies its input into a buffer with this much zero
// padding, which plays the role of the reference's BitReaderBuilder eof buffer
// (pco/src/bit_reader.rs:271-300): any single "with_reader" section reads less
// than this past a position that was verified in bounds.
constexpr size_t READ_PADDING = 8192 + 64;

// pco/src/errors.rs:8-24
enum class ErrorKind : int { None = 0, Corruption = 1, InsufficientData = 2, InvalidArgument = 3, Io = 4 };

struct PcoError : std::exception {
  ErrorKind kind;
  std::string msg;
  PcoError(ErrorKind k, std::string m) : kind(k), msg(std::move(m)) {}
  const char* what() const noexcept override { return msg.c_str(); }
};
[[noreturn]] inline void corruption(const std::string& m) { throw PcoError(ErrorKind::Corruption, m); }
[[noreturn]] inline void insufficient_data(const std::string& m) { throw PcoError(ErrorKind::InsufficientData, m); }
[[noreturn]] inline void invalid_argument(const std::string& m) { throw PcoError(ErrorKind::InvalidArgument, m); }
[[noreturn]] inline void io_error(const std::string& m) { throw PcoError(ErrorKind::Io, m); }

// ---------------------------------------------------------------------------
// Latent traits (pco/src/data_types/latent_priv.rs:11-64, unsigned.rs:89-155)
// ---------------------------------------------------------------------------
template <typename L>
struct LatentTraits {
  static constexpr Bitlen BITS = sizeof(L) * 8;
  static constexpr L ZERO = 0;
  static constexpr L ONE = 1;
  static constexpr L MID = L(L(1) << (BITS - 1));
  static constexpr L MAX = std::numeric_limits<L>::max();
};

template <typename L>
inline Bitlen leading_zeros(L x) {
  constexpr Bitlen BITS = sizeof(L) * 8;
  if (x == 0) return BITS;
  return Bitlen(__builtin_clzll((unsigned long long)x)) - (64 - BITS);
}
inline Bitlen ilog2_u64(uint64_t x) { return 63 - Bitlen(__builtin_clzll(x)); }

// pco/src/bits.rs:19-26
template <typename L>
inline Bitlen bits_to_encode_offset(L max_offset) {
  return Bitlen(sizeof(L) * 8) - leading_zeros<L>(max_offset);
}
inline Bitlen bits_to_encode_offset_bits(Bitlen latent_bits) {
  // (Bitlen::BITS - L::BITS.leading_zeros()): 8->4, 16->5, 32->6, 64->7
  return 32 - Bitlen(__builtin_clz(latent_bits));
}
inline uint64_t lowest_bits_u64(uint64_t x, Bitlen n) { return n >= 64 ? x : (x & ((uint64_t(1) << n) - 1)); }

// ---------------------------------------------------------------------------
// Bit writer (pco/src/bit_writer.rs:22-147).  The reference stages into a
// scratch Vec and flushes to a generic Write; here the destination is always a
// growing byte vector, which is observably the same byte stream.
// ---------------------------------------------------------------------------
struct BitWriter {
  std::vector<uint8_t>& dst;
  size_t base;      // byte index in dst where this writer started
  size_t bit_idx;   // bits written since base
  explicit BitWriter(std::vector<uint8_t>& d) : dst(d), base(d.size()), bit_idx(0) {}

  inline void ensure(size_t extra_bytes) {
    size_t need = base + bit_idx / 8 + extra_bytes;
    if (dst.size() < need) dst.resize(need, 0);
  }
  // little-endian OR of the low n bits of x at the cursor; n <= 64
  // (bit_writer.rs:22-42 write_uint_to + :94-125 write_uint)
  inline void write_uint(uint64_t x, Bitlen n) {
    if (n == 0) return;
    x = lowest_bits_u64(x, n);
    ensure(17);
    uint8_t* p = dst.data() + base + bit_idx / 8;
    Bitlen bpb = Bitlen(bit_idx % 8);
    uint64_t w;
    std::memcpy(&w, p, 8);
    w |= x << bpb;
    std::memcpy(p, &w, 8);
    if (bpb + n > 64) {
      uint64_t w2;
      std::memcpy(&w2, p + 8, 8);
      w2 |= x >> (64 - bpb);
      std::memcpy(p + 8, &w2, 8);
    }
    bit_idx += n;
  }
  inline void write_bool(bool b) { write_uint(b ? 1 : 0, 1); }
  // bit_writer.rs:80-92
  inline void write_aligned_bytes(const uint8_t* bytes, size_t n) {
    if (bit_idx % 8 != 0) invalid_argument("cannot write aligned bytes to unaligned writer");
    ensure(n + 17);
    std::memcpy(dst.data() + base + bit_idx / 8, bytes, n);
    bit_idx += 8 * n;
  }
  // bit_writer.rs:136-139
  inline void finish_byte() { bit_idx = (bit_idx + 7) / 8 * 8; }
  // Trim the destination to exactly the bytes written (the reference's flush
  // only emits whole bytes; every section ends with finish_byte()).
  inline void finish() {
    finish_byte();
    dst.resize(base + bit_idx / 8);
  }
};

// ---------------------------------------------------------------------------
// Bit reader (pco/src/bit_reader.rs:14-247).  `src` must be padded with
// READ_PADDING zero bytes past `unpadded_len`.
// ---------------------------------------------------------------------------
struct BitReader {
  const uint8_t* src;
  size_t unpadded_len;  // bytes
  size_t bit_idx;

  BitReader(const uint8_t* s, size_t len, size_t start_bit = 0) : src(s), unpadded_len(len), bit_idx(start_bit) {}

  inline uint64_t u64_at(size_t byte_idx) const {
    uint64_t w;
    std::memcpy(&w, src + byte_idx, 8);
    return w;
  }
  // bit_reader.rs:30-106; n <= 64.  Reads are clamped to the padded buffer.
  inline uint64_t read_uint(Bitlen n) {
    if (n == 0) return 0;
    size_t byte_idx = bit_idx / 8;
    Bitlen bpb = Bitlen(bit_idx % 8);
    bit_idx += n;
    if (byte_idx + 16 > unpadded_len + READ_PADDING) return 0;  // far past EOF: zeros
    uint64_t w = u64_at(byte_idx) >> bpb;
    if (bpb + n > 64) w |= u64_at(byte_idx + 8) << (64 - bpb);
    return lowest_bits_u64(w, n);
  }
  inline bool read_bool() { return read_uint(1) != 0; }
  inline size_t byte_idx() const { return bit_idx / 8; }
  // bit_reader.rs:218-232
  inline void check_in_bounds() const {
    if (bit_idx > unpadded_len * 8)
      insufficient_data("[BitReader] out of bounds at bit " + std::to_string(bit_idx) + " / " +
                        std::to_string(unpadded_len * 8));
  }
  // bit_reader.rs:164-170.  Returns pointer to n bytes at the (aligned) cursor.
  inline const uint8_t* read_aligned_bytes(size_t n) {
    if (bit_idx % 8 != 0) invalid_argument("cannot get aligned byte index on misaligned bit reader");
    size_t b = bit_idx / 8;
    bit_idx += 8 * n;
    // Callers wrap this in a section that checks bounds afterwards; keep the
    // pointer valid by clamping into the padding.
    if (b + n > unpadded_len + READ_PADDING) insufficient_data("[BitReader] aligned read far out of bounds");
    return src + b;
  }
  // bit_reader.rs:237-247
  inline void drain_empty_byte(const char* message) {
    check_in_bounds();
    Bitlen bpb = Bitlen(bit_idx % 8);
    if (bpb != 0) {
      if ((src[bit_idx / 8] >> bpb) > 0) corruption(message);
      bit_idx += 8 - bpb;
    }
  }
};

// ---------------------------------------------------------------------------
// tANS (pco/src/ans/spec.rs, encoding.rs, decoding.rs)
// ---------------------------------------------------------------------------
struct AnsSpec {
  Bitlen size_log = 0;
  std::vector<Symbol> state_symbols;
  std::vector<Weight> symbol_weights;
  size_t table_size() const { return size_t(1) << size_log; }
};

// pco/src/ans/spec.rs:24-30
inline Weight choose_stride(Weight table_size) {
  Weight res = (3 * table_size) / 5;
  if (res % 2 == 0) res += 1;
  return res;
}

// pco/src/ans/spec.rs:37-59, :61-75
inline AnsSpec ans_spec_from_weights(Bitlen size_log, std::vector<Weight> symbol_weights) {
  if (symbol_weights.empty()) symbol_weights = {1};
  uint64_t table_size = 0;
  for (Weight w : symbol_weights) table_size += w;
  if (table_size != (uint64_t(1) << size_log))
    corruption("table size log of " + std::to_string(size_log) + " does not agree with total weight of " +
               std::to_string(table_size));
  AnsSpec spec;
  spec.size_log = size_log;
  spec.state_symbols.assign(table_size, 0);
  Weight step = 0;
  Weight stride = choose_stride(Weight(table_size));
  Weight mod_table_size = Weight(table_size - 1);
  for (size_t symbol = 0; symbol < symbol_weights.size(); symbol++) {
    for (Weight k = 0; k < symbol_weights[symbol]; k++) {
      Weight state_idx = (stride * step) & mod_table_size;
      spec.state_symbols[state_idx] = Symbol(symbol);
      step += 1;
    }
  }
  spec.symbol_weights = std::move(symbol_weights);
  return spec;
}

// pco/src/ans/encoding.rs:8-92
struct AnsEncoder {
  struct SymbolInfo {
    AnsState renorm_bit_cutoff;
    Bitlen min_renorm_bits;
    std::vector<AnsState> next_states;
  };
  std::vector<SymbolInfo> symbol_infos;
  Bitlen size_log = 0;

  AnsEncoder() = default;
  explicit AnsEncoder(const AnsSpec& spec) : size_log(spec.size_log) {
    size_t table_size = spec.table_size();
    for (Weight weight : spec.symbol_weights) {
      Weight max_x_s = 2 * weight - 1;
      Bitlen min_renorm_bits = spec.size_log - ilog2_u64(max_x_s);
      AnsState cutoff = AnsState(2 * weight * (1u << min_renorm_bits));
      SymbolInfo si{cutoff, min_renorm_bits, {}};
      si.next_states.reserve(weight);
      symbol_infos.push_back(std::move(si));
    }
    for (size_t state_idx = 0; state_idx < spec.state_symbols.size(); state_idx++)
      symbol_infos[spec.state_symbols[state_idx]].next_states.push_back(AnsState(table_size + state_idx));
  }
  // returns new state; *bits = number of low bits of the old state to emit
  inline AnsState encode(AnsState state, Symbol symbol, Bitlen* bits) const {
    const SymbolInfo& si = symbol_infos[symbol];
    Bitlen renorm_bits = state >= si.renorm_bit_cutoff ? si.min_renorm_bits + 1 : si.min_renorm_bits;
    *bits = renorm_bits;
    return si.next_states[(state >> renorm_bits) - si.next_states.size()];
  }
  AnsState default_state() const { return AnsState(1) << size_log; }
};

// pco/src/ans/decoding.rs:15-48
struct AnsNode {
  uint16_t next_state_idx_base;
  uint8_t offset_bits;
  uint8_t bits_to_read;
};

inline std::vector<AnsNode> ans_decoder_nodes(const AnsSpec& spec, const std::vector<Bitlen>& bin_offset_bits) {
  size_t table_size = spec.table_size();
  std::vector<AnsNode> nodes;
  nodes.reserve(table_size);
  std::vector<Weight> symbol_x_s = spec.symbol_weights;
  for (Symbol symbol : spec.state_symbols) {
    AnsState next_state_base = symbol_x_s[symbol];
    Bitlen bits_to_read = Bitlen(__builtin_clz(next_state_base)) - Bitlen(__builtin_clz(AnsState(table_size)));
    next_state_base <<= bits_to_read;
    Bitlen offset_bits = symbol < bin_offset_bits.size() ? bin_offset_bits[symbol] : 0;
    nodes.push_back(AnsNode{uint16_t(next_state_base - AnsState(table_size)), uint8_t(offset_bits), uint8_t(bits_to_read)});
    symbol_x_s[symbol] += 1;
  }
  return nodes;
}

// pco/src/ans/encoding.rs:95-151
inline std::vector<Weight> quantize_weights_to(const std::vector<Weight>& counts, size_t total_count, Bitlen size_log) {
  if (size_log == 0) return {1};
  Weight required_weight_sum = Weight(1) << size_log;
  float multiplier = float(required_weight_sum) / float(total_count);
  std::vector<float> desired_surplus_per_bin(counts.size());
  for (size_t i = 0; i < counts.size(); i++) {
    float v = float(counts[i]) * multiplier - 1.0f;
    // f32::max(0.0): NaN-free here
    desired_surplus_per_bin[i] = v > 0.0f ? v : 0.0f;
  }
  float desired_surplus = 0.0f;
  for (float s : desired_surplus_per_bin) desired_surplus = desired_surplus + s;
  Weight required_surplus = required_weight_sum - Weight(counts.size());
  float surplus_mult = desired_surplus == 0.0f ? 0.0f : float(required_surplus) / desired_surplus;
  std::vector<float> float_weights(counts.size());
  for (size_t i = 0; i < counts.size(); i++) float_weights[i] = 1.0f + desired_surplus_per_bin[i] * surplus_mult;
  std::vector<Weight> weights(counts.size());
  Weight weight_sum = 0;
  for (size_t i = 0; i < counts.size(); i++) {
    // f32::round = half away from zero; `as u32` saturates (values are small positives here)
    weights[i] = Weight(std::round(float_weights[i]));
    weight_sum += weights[i];
  }
  size_t i = 0;
  while (weight_sum > required_weight_sum) {
    if (weights[i] > 1 && float(weights[i]) > float_weights[i]) {
      weights[i] -= 1;
      weight_sum -= 1;
    }
    i += 1;
  }
  i = 0;
  while (weight_sum < required_weight_sum) {
    if (float(weights[i]) < float_weights[i]) {
      weights[i] += 1;
      weight_sum += 1;
    }
    i += 1;
  }
  return weights;
}

// pco/src/ans/encoding.rs:156-175
inline std::pair<Bitlen, std::vector<Weight>> quantize_weights(const std::vector<Weight>& counts, size_t total_count,
                                                                Bitlen max_size_log) {
  if (counts.size() == 1) return {0, {1}};
  size_t m = counts.size() - 1;
  Bitlen min_size_log = m == 0 ? 0 : Bitlen(64 - __builtin_clzll((unsigned long long)m));
  Bitlen size_log = std::max(min_size_log, max_size_log);
  std::vector<Weight> weights = quantize_weights_to(counts, total_count, size_log);
  Bitlen power_of_2 = 32;
  for (Weight w : weights) power_of_2 = std::min(power_of_2, w == 0 ? Bitlen(32) : Bitlen(__builtin_ctz(w)));
  size_log -= power_of_2;
  for (Weight& w : weights) w >>= power_of_2;
  return {size_log, weights};
}

}  // namespace pco_oracle
