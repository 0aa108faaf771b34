// ORACLE — TEST INFRASTRUCTURE ONLY (see pco_core.hpp header).
// Number types, chunk/page metadata and their bit-exact (de)serialisation.
#pragma once
#include "pco_core.hpp"

namespace pco_oracle {

// ---------------------------------------------------------------------------
// Number types (pco/src/data_types/{unsigned,signed,float}.rs, dynamic.rs)
// type bytes: pco_c/include/cpcodec.h:10-20
// ---------------------------------------------------------------------------
enum NumberType : uint8_t {
  NT_U32 = 1, NT_U64 = 2, NT_I32 = 3, NT_I64 = 4, NT_F32 = 5, NT_F64 = 6,
  NT_U16 = 7, NT_I16 = 8, NT_F16 = 9, NT_U8 = 10, NT_I8 = 11,
};
inline bool number_type_valid(uint8_t b) { return b >= 1 && b <= 11; }
inline Bitlen number_type_bits(uint8_t b) {
  switch (b) {
    case NT_U8: case NT_I8: return 8;
    case NT_U16: case NT_I16: case NT_F16: return 16;
    case NT_U32: case NT_I32: case NT_F32: return 32;
    default: return 64;
  }
}
inline bool number_type_is_float(uint8_t b) { return b == NT_F16 || b == NT_F32 || b == NT_F64; }
inline bool number_type_is_signed(uint8_t b) { return b == NT_I8 || b == NT_I16 || b == NT_I32 || b == NT_I64; }

// Order-preserving maps number bits <-> latent, on raw bit patterns.
// unsigned.rs:155-161, signed.rs:46-52, float.rs:392-411
template <typename L>
inline L to_latent_ordered_bits(L bits, bool is_float, bool is_signed) {
  constexpr L MID = LatentTraits<L>::MID;
  if (is_float) return (bits & MID) ? L(~bits) : L(bits ^ MID);
  if (is_signed) return L(bits - MID);  // x.wrapping_sub(MIN) as L
  return bits;
}
template <typename L>
inline L from_latent_ordered_bits(L l, bool is_float, bool is_signed) {
  constexpr L MID = LatentTraits<L>::MID;
  if (is_float) return (l & MID) ? L(l ^ MID) : L(~l);
  if (is_signed) return L(l + MID);  // (l as Self).wrapping_add(MIN)
  return l;
}

// ---------------------------------------------------------------------------
// Metadata (pco/src/metadata/*)
// ---------------------------------------------------------------------------
enum class ModeKind : int { Classic = 0, IntMult = 1, FloatMult = 2, FloatQuant = 3, Dict = 4 };
struct Mode {
  ModeKind kind = ModeKind::Classic;
  uint64_t base_latent = 0;      // IntMult: base; FloatMult: to_latent_ordered(base)
  Bitlen k = 0;                  // FloatQuant
  std::vector<uint64_t> dict;    // Dict (latent-ordered values)
};

enum class DeltaKind : int { NoOp = 0, Consecutive = 1, Lookback = 2, Conv1 = 3 };
struct DeltaEncoding {
  DeltaKind kind = DeltaKind::NoOp;
  size_t order = 0;                   // Consecutive
  bool secondary_uses_delta = false;  // Consecutive / Lookback
  Bitlen window_n_log = 0, state_n_log = 0;  // Lookback
  Bitlen quantization = 0;            // Conv1
  int64_t bias = 0;
  std::vector<int64_t> weights;
  // pco/src/metadata/delta_encoding.rs:111-116
  static constexpr size_t MAX_BIT_SIZE = (BITS_TO_ENCODE_DELTA_ENCODING_VARIANT + BITS_TO_ENCODE_DELTA_CONV_QUANTIZATION +
                                          BITS_TO_ENCODE_DELTA_CONV_N_WEIGHTS) + 64 + MAX_CONV1_DELTA_ORDER * 32;
};

enum class VarKey : int { Delta = 0, Primary = 1, Secondary = 2 };

// pco/src/metadata/delta_encoding.rs:55-73, :308-351 (for_latent_var)
struct LatentVarDelta {
  DeltaKind kind = DeltaKind::NoOp;
  const DeltaEncoding* enc = nullptr;
  size_t n_latents_per_state() const {
    switch (kind) {
      case DeltaKind::NoOp: return 0;
      case DeltaKind::Consecutive: return enc->order;
      case DeltaKind::Lookback: return size_t(1) << enc->state_n_log;
      case DeltaKind::Conv1: return enc->weights.size();
    }
    return 0;
  }
};
inline LatentVarDelta delta_for_latent_var(const DeltaEncoding& d, VarKey key) {
  LatentVarDelta r;
  r.enc = &d;
  if (d.kind == DeltaKind::NoOp || key == VarKey::Delta) return r;
  if (key == VarKey::Primary) { r.kind = d.kind; return r; }
  // secondary
  if ((d.kind == DeltaKind::Consecutive || d.kind == DeltaKind::Lookback) && d.secondary_uses_delta) r.kind = d.kind;
  return r;
}

// pco/src/metadata/bin.rs:8-38
struct Bin {
  Weight weight;
  uint64_t lower;
  Bitlen offset_bits;
};

// pco/src/metadata/chunk_latent_var.rs:81-92
struct LatentVarMeta {
  Bitlen latent_bits = 64;
  Bitlen ans_size_log = 0;
  std::vector<Bin> bins;
  // chunk_latent_var.rs:166-187, bin.rs:20-22
  size_t exact_bit_size() const {
    size_t per_bin = ans_size_log + latent_bits + bits_to_encode_offset_bits(latent_bits);
    return BITS_TO_ENCODE_ANS_SIZE_LOG + BITS_TO_ENCODE_N_BINS + bins.size() * per_bin;
  }
  size_t exact_page_meta_bit_size(const LatentVarDelta& d) const {
    return size_t(ans_size_log) * ANS_INTERLEAVING + size_t(latent_bits) * d.n_latents_per_state();
  }
  Bitlen max_offset_bits() const {
    Bitlen m = 0;
    for (const Bin& b : bins) m = std::max(m, b.offset_bits);
    return m;
  }
  // metadata/bins.rs:7-9
  bool are_trivial() const { return bins.empty() || (bins.size() == 1 && bins[0].offset_bits == 0); }
};

struct ChunkMeta {
  Bitlen number_bits = 64;  // L::BITS of the number type
  Mode mode;
  DeltaEncoding delta;
  bool has_delta_var = false, has_secondary = false;
  LatentVarMeta delta_var, primary, secondary;

  // vars in file order (per_latent_var.rs:140-150)
  std::vector<std::pair<VarKey, const LatentVarMeta*>> vars() const {
    std::vector<std::pair<VarKey, const LatentVarMeta*>> r;
    if (has_delta_var) r.push_back({VarKey::Delta, &delta_var});
    r.push_back({VarKey::Primary, &primary});
    if (has_secondary) r.push_back({VarKey::Secondary, &secondary});
    return r;
  }
  // metadata/mode.rs:222-233
  size_t mode_max_bit_size() const {
    size_t payload = 0;
    switch (mode.kind) {
      case ModeKind::Classic: payload = 0; break;
      case ModeKind::Dict: payload = BITS_TO_ENCODE_DICT_LEN + 7 + mode.dict.size() * number_bits; break;
      case ModeKind::FloatMult: case ModeKind::IntMult: payload = number_bits; break;
      case ModeKind::FloatQuant: payload = BITS_TO_ENCODE_QUANTIZE_K; break;
    }
    return BITS_TO_ENCODE_MODE_VARIANT + payload;
  }
  // metadata/chunk.rs:105-125
  size_t max_size() const {
    size_t bits = 0;
    for (auto& kv : vars()) bits += kv.second->exact_bit_size();
    size_t n_bits = mode_max_bit_size() + DeltaEncoding::MAX_BIT_SIZE + bits;
    return (n_bits + 7) / 8;
  }
  size_t exact_page_meta_size() const {
    size_t bits = 0;
    for (auto& kv : vars()) bits += kv.second->exact_page_meta_bit_size(delta_for_latent_var(delta, kv.first));
    return (bits + 7) / 8;
  }
};

struct FormatVersion {
  uint8_t major = 4, minor = 1;
  bool used_old_gcds() const { return major == 0; }
  bool supports_delta_variants() const { return major >= 3; }
};

// metadata/mode.rs:190-220 latent types of the vars
inline Bitlen primary_latent_bits(const Mode& m, Bitlen number_bits) { return m.kind == ModeKind::Dict ? 32 : number_bits; }
inline bool mode_has_secondary(const Mode& m) {
  return m.kind == ModeKind::IntMult || m.kind == ModeKind::FloatMult || m.kind == ModeKind::FloatQuant;
}

// ----- mode validity (unsigned.rs:80-86, float.rs:372-384) -----------------
template <typename F> inline F float_from_bits_generic(uint64_t bits);
template <> inline float float_from_bits_generic<float>(uint64_t bits) { uint32_t b = uint32_t(bits); float f; std::memcpy(&f, &b, 4); return f; }
template <> inline double float_from_bits_generic<double>(uint64_t bits) { double f; std::memcpy(&f, &bits, 8); return f; }

inline float f16_bits_to_f32(uint16_t h) {
  uint32_t sign = uint32_t(h & 0x8000) << 16;
  uint32_t exp = (h >> 10) & 0x1f;
  uint32_t man = h & 0x3ff;
  uint32_t out;
  if (exp == 0) {
    if (man == 0) out = sign;
    else {
      int e = -1;
      do { e++; man <<= 1; } while ((man & 0x400) == 0);
      out = sign | uint32_t(127 - 15 - e) << 23 | (man & 0x3ff) << 13;
    }
  } else if (exp == 31) out = sign | 0x7f800000u | man << 13;
  else out = sign | (exp + 127 - 15) << 23 | man << 13;
  float f; std::memcpy(&f, &out, 4); return f;
}
// round-to-nearest-even f32 -> f16 (half crate semantics)
inline uint16_t f32_to_f16_bits(float f) {
  uint32_t x; std::memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000;
  uint32_t exp = (x >> 23) & 0xff;
  uint32_t man = x & 0x7fffff;
  if (exp == 255) return uint16_t(sign | 0x7c00 | (man ? (0x200 | (man >> 13)) : 0));
  int32_t e = int32_t(exp) - 127 + 15;
  if (e >= 31) return uint16_t(sign | 0x7c00);
  if (e <= 0) {
    if (e < -10) return uint16_t(sign);
    man |= 0x800000;
    uint32_t shift = uint32_t(14 - e);
    uint32_t half_man = man >> shift;
    uint32_t rem = man & ((1u << shift) - 1);
    uint32_t halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (half_man & 1))) half_man++;
    return uint16_t(sign | half_man);
  }
  uint32_t half = sign | uint32_t(e) << 10 | man >> 13;
  uint32_t rem = man & 0x1fff;
  if (rem > 0x1000 || (rem == 0x1000 && (half & 1))) half++;
  return uint16_t(half);
}

inline bool mode_is_valid(const Mode& m, uint8_t number_type) {
  Bitlen bits = number_type_bits(number_type);
  if (number_type_is_float(number_type)) {
    switch (m.kind) {
      case ModeKind::Classic: case ModeKind::Dict: return true;
      case ModeKind::IntMult: return false;
      case ModeKind::FloatQuant: {
        Bitlen precision = bits == 64 ? 52 : bits == 32 ? 23 : 10;
        return m.k > 0 && m.k <= precision;
      }
      case ModeKind::FloatMult: {
        if (bits == 64) {
          double b = float_from_bits_generic<double>(from_latent_ordered_bits<uint64_t>(m.base_latent, true, false));
          return std::isfinite(b) && std::fabs(b) > 0.0;
        } else if (bits == 32) {
          float b = float_from_bits_generic<float>(from_latent_ordered_bits<uint32_t>(uint32_t(m.base_latent), true, false));
          return std::isfinite(b) && std::fabs(b) > 0.0f;
        } else {
          float b = f16_bits_to_f32(from_latent_ordered_bits<uint16_t>(uint16_t(m.base_latent), true, false));
          return std::isfinite(b) && std::fabs(b) > 0.0f;
        }
      }
    }
    return false;
  }
  switch (m.kind) {
    case ModeKind::Classic: case ModeKind::Dict: return true;
    case ModeKind::FloatMult: case ModeKind::FloatQuant: return false;
    case ModeKind::IntMult: return m.base_latent > 0;
  }
  return false;
}

// ---------------------------------------------------------------------------
// Serialisation
// ---------------------------------------------------------------------------
// metadata/mode.rs:169-188
inline void write_mode(const Mode& m, Bitlen number_bits, BitWriter& w) {
  w.write_uint(uint64_t(m.kind), BITS_TO_ENCODE_MODE_VARIANT);
  switch (m.kind) {
    case ModeKind::Classic: break;
    case ModeKind::IntMult: case ModeKind::FloatMult: w.write_uint(m.base_latent, number_bits); break;
    case ModeKind::FloatQuant: w.write_uint(m.k, BITS_TO_ENCODE_QUANTIZE_K); break;
    case ModeKind::Dict:
      w.write_uint(m.dict.size(), BITS_TO_ENCODE_DICT_LEN);
      w.finish_byte();
      for (uint64_t v : m.dict) w.write_uint(v, number_bits);
      break;
  }
}
// metadata/delta_encoding.rs:204-261
inline void write_delta_encoding(const DeltaEncoding& d, BitWriter& w) {
  w.write_uint(uint64_t(d.kind), BITS_TO_ENCODE_DELTA_ENCODING_VARIANT);
  switch (d.kind) {
    case DeltaKind::NoOp: break;
    case DeltaKind::Consecutive:
      w.write_uint(d.order, BITS_TO_ENCODE_DELTA_ENCODING_ORDER);
      w.write_bool(d.secondary_uses_delta);
      break;
    case DeltaKind::Lookback:
      w.write_uint(d.window_n_log - 1, BITS_TO_ENCODE_DELTA_LOOKBACK_WINDOW_N_LOG);
      w.write_uint(d.state_n_log, BITS_TO_ENCODE_DELTA_LOOKBACK_STATE_N_LOG);
      w.write_bool(d.secondary_uses_delta);
      break;
    case DeltaKind::Conv1:
      w.write_uint(d.quantization, BITS_TO_ENCODE_DELTA_CONV_QUANTIZATION);
      w.write_uint(uint64_t(d.bias) ^ (uint64_t(1) << 63), 64);  // i64::to_latent_ordered
      w.write_uint(d.weights.size() - 1, BITS_TO_ENCODE_DELTA_CONV_N_WEIGHTS);
      for (int64_t wt : d.weights) w.write_uint(uint32_t(int32_t(wt)) ^ 0x80000000u, 32);
      break;
  }
}
// metadata/chunk_latent_var.rs:55-71, :145-158
inline void write_latent_var_meta(const LatentVarMeta& v, BitWriter& w) {
  w.write_uint(v.ans_size_log, BITS_TO_ENCODE_ANS_SIZE_LOG);
  w.write_uint(v.bins.size(), BITS_TO_ENCODE_N_BINS);
  Bitlen obb = bits_to_encode_offset_bits(v.latent_bits);
  for (const Bin& b : v.bins) {
    w.write_uint(b.weight - 1, v.ans_size_log);
    w.write_uint(b.lower, v.latent_bits);
    w.write_uint(b.offset_bits, obb);
  }
}
// metadata/chunk.rs:176-189
inline void write_chunk_meta(const ChunkMeta& m, std::vector<uint8_t>& dst) {
  BitWriter w(dst);
  write_mode(m.mode, m.number_bits, w);
  write_delta_encoding(m.delta, w);
  for (auto& kv : m.vars()) write_latent_var_meta(*kv.second, w);
  w.finish();
}

// metadata/mode.rs:102-167
inline Mode read_mode(BitReader& r, const FormatVersion& version, Bitlen number_bits) {
  Mode m;
  Bitlen variant = Bitlen(r.read_uint(BITS_TO_ENCODE_MODE_VARIANT));
  size_t n_unique = 0;
  switch (variant) {
    case 0: m.kind = ModeKind::Classic; break;
    case 1:
      if (version.used_old_gcds())
        corruption("unable to decompress data from yanked v0.0.0 of pco with different GCD encoding");
      m.kind = ModeKind::IntMult;
      m.base_latent = r.read_uint(number_bits);
      break;
    case 2: m.kind = ModeKind::FloatMult; m.base_latent = r.read_uint(number_bits); break;
    case 3: m.kind = ModeKind::FloatQuant; m.k = Bitlen(r.read_uint(BITS_TO_ENCODE_QUANTIZE_K)); break;
    case 4:
      m.kind = ModeKind::Dict;
      n_unique = size_t(r.read_uint(BITS_TO_ENCODE_DICT_LEN));
      r.drain_empty_byte("expected zeros between dict mode length and values");
      break;
    default: r.check_in_bounds(); corruption("unknown mode variant " + std::to_string(variant));
  }
  r.check_in_bounds();  // end of the FIXED_READ_SIZE section
  if (m.kind == ModeKind::Dict) {
    // dyn_latents.rs:33-55: batches of 512 values, bounds-checked per batch
    m.dict.reserve(std::min<size_t>(n_unique, 1 << 16));
    for (size_t start = 0; start < n_unique; start += 512) {
      size_t end = std::min(start + 512, n_unique);
      for (size_t i = start; i < end; i++) m.dict.push_back(r.read_uint(number_bits));
      r.check_in_bounds();
    }
  }
  return m;
}

// metadata/delta_encoding.rs:118-202
inline DeltaEncoding read_delta_encoding(BitReader& r, const FormatVersion& version) {
  DeltaEncoding d;
  if (!version.supports_delta_variants()) {
    size_t order = size_t(r.read_uint(BITS_TO_ENCODE_DELTA_ENCODING_ORDER));
    if (order != 0) { d.kind = DeltaKind::Consecutive; d.order = order; }
    return d;
  }
  Bitlen variant = Bitlen(r.read_uint(BITS_TO_ENCODE_DELTA_ENCODING_VARIANT));
  switch (variant) {
    case 0: break;
    case 1: {
      size_t order = size_t(r.read_uint(BITS_TO_ENCODE_DELTA_ENCODING_ORDER));
      if (order == 0) corruption("Consecutive delta encoding order must not be 0");
      d.kind = DeltaKind::Consecutive;
      d.order = order;
      d.secondary_uses_delta = r.read_bool();
      break;
    }
    case 2: {
      Bitlen window_n_log = 1 + Bitlen(r.read_uint(BITS_TO_ENCODE_DELTA_LOOKBACK_WINDOW_N_LOG));
      Bitlen state_n_log = Bitlen(r.read_uint(BITS_TO_ENCODE_DELTA_LOOKBACK_STATE_N_LOG));
      if (window_n_log > MAX_DELTA_LOOKBACK_WINDOW_N_LOG) corruption("LZ delta encoding window size log exceeds max");
      if (state_n_log > window_n_log) corruption("LZ delta encoding state size log exceeded window size log");
      d.kind = DeltaKind::Lookback;
      d.window_n_log = window_n_log;
      d.state_n_log = state_n_log;
      d.secondary_uses_delta = r.read_bool();
      break;
    }
    case 3: {
      d.kind = DeltaKind::Conv1;
      d.quantization = Bitlen(r.read_uint(BITS_TO_ENCODE_DELTA_CONV_QUANTIZATION));
      d.bias = int64_t(r.read_uint(64) ^ (uint64_t(1) << 63));
      size_t order = 1 + size_t(r.read_uint(BITS_TO_ENCODE_DELTA_CONV_N_WEIGHTS));
      for (size_t i = 0; i < order; i++) d.weights.push_back(int64_t(int32_t(uint32_t(r.read_uint(32)) ^ 0x80000000u)));
      break;
    }
    default: corruption("unknown delta encoding value: " + std::to_string(variant));
  }
  return d;
}

// metadata/chunk_latent_var.rs:22-53, :102-143
inline LatentVarMeta read_latent_var_meta(BitReader& r, Bitlen latent_bits) {
  LatentVarMeta v;
  v.latent_bits = latent_bits;
  v.ans_size_log = Bitlen(r.read_uint(BITS_TO_ENCODE_ANS_SIZE_LOG));
  size_t n_bins = size_t(r.read_uint(BITS_TO_ENCODE_N_BINS));
  r.check_in_bounds();
  if ((size_t(1) << v.ans_size_log) < n_bins) corruption("ANS size log is too small for number of bins");
  if (n_bins == 1 && v.ans_size_log > 0) corruption("Only 1 bin but ANS size log is > 0");
  if (v.ans_size_log > MAX_ANS_BITS) corruption("ANS size log should not be greater than 14");
  Bitlen obb = bits_to_encode_offset_bits(latent_bits);
  v.bins.reserve(n_bins);
  for (size_t start = 0; start < n_bins; start += 128) {
    size_t end = std::min(start + 128, n_bins);
    for (size_t i = start; i < end; i++) {
      Bin b;
      b.weight = Weight(r.read_uint(v.ans_size_log)) + 1;
      b.lower = r.read_uint(latent_bits);
      b.offset_bits = Bitlen(r.read_uint(obb));
      if (b.offset_bits > latent_bits) {
        r.check_in_bounds();
        corruption("offset bits of " + std::to_string(b.offset_bits) + " exceeds type of " + std::to_string(latent_bits) + " bits");
      }
      v.bins.push_back(b);
    }
    r.check_in_bounds();
  }
  return v;
}

// metadata/chunk.rs:32-103 ChunkMeta::new validation
inline void validate_chunk_meta(const ChunkMeta& m) {
  const DeltaEncoding& d = m.delta;
  if (d.kind == DeltaKind::Lookback) {
    uint64_t window_n = uint64_t(1) << d.window_n_log;
    for (const Bin& b : m.delta_var.bins)
      if (b.lower < 1 || b.lower > window_n) corruption("delta lookback bin had invalid lower bound outside window");
  } else if (d.kind == DeltaKind::Conv1) {
    Bitlen l_bits = m.primary.latent_bits;
    if (l_bits == 64) corruption("Conv1 delta encodings are not supported on types larger than 32 bits");
    Bitlen conv_bits = l_bits == 8 ? 16 : l_bits == 16 ? 32 : 64;
    Bitlen max_q = std::min(MAX_CONV1_DELTA_QUANTIZATION, conv_bits - 1);
    if (d.quantization > max_q) corruption("Conv1 delta encoding quantization exceeds max");
    double wsum = 0.0;
    for (int64_t w : d.weights) wsum += double(w < 0 ? -w : w);  // (w.abs() as f64) summed in order
    double max_pred = std::fabs(double(d.bias)) + std::pow(2.0, int(l_bits)) * wsum;
    if (max_pred >= std::pow(2.0, int(conv_bits) - 1)) corruption("Conv1 delta encoding weights and bias risk overflowing");
  }
}

// metadata/chunk.rs:127-174
inline ChunkMeta read_chunk_meta(BitReader& r, const FormatVersion& version, Bitlen number_bits) {
  ChunkMeta m;
  m.number_bits = number_bits;
  m.mode = read_mode(r, version, number_bits);
  m.delta = read_delta_encoding(r, version);
  r.check_in_bounds();
  if (m.delta.kind == DeltaKind::Lookback) {
    m.has_delta_var = true;
    m.delta_var = read_latent_var_meta(r, 32);
  }
  m.primary = read_latent_var_meta(r, primary_latent_bits(m.mode, number_bits));
  if (mode_has_secondary(m.mode)) {
    m.has_secondary = true;
    m.secondary = read_latent_var_meta(r, number_bits);
  }
  r.drain_empty_byte("nonzero bits in end of final byte of chunk metadata");
  validate_chunk_meta(m);
  return m;
}

// ----- page meta (metadata/page.rs, page_latent_var.rs) ---------------------
struct PageVarMeta {
  std::vector<uint64_t> delta_state;
  AnsState ans_final_state_idxs[ANS_INTERLEAVING] = {0, 0, 0, 0};
};
struct PageMeta {
  PageVarMeta delta_var, primary, secondary;
};

inline void write_page_meta(const ChunkMeta& cm, const PageMeta& pm, BitWriter& w) {
  auto one = [&](const LatentVarMeta& vm, const PageVarMeta& pv) {
    for (uint64_t s : pv.delta_state) w.write_uint(s, vm.latent_bits);
    for (size_t j = 0; j < ANS_INTERLEAVING; j++) w.write_uint(pv.ans_final_state_idxs[j], vm.ans_size_log);
  };
  if (cm.has_delta_var) one(cm.delta_var, pm.delta_var);
  one(cm.primary, pm.primary);
  if (cm.has_secondary) one(cm.secondary, pm.secondary);
  w.finish_byte();
}

inline PageMeta read_page_meta(BitReader& r, const ChunkMeta& cm) {
  PageMeta pm;
  auto one = [&](const LatentVarMeta& vm, VarKey key, PageVarMeta& pv) {
    size_t n_state = delta_for_latent_var(cm.delta, key).n_latents_per_state();
    // A corrupt header can ask for a 2^15-entry lookback state; reads far past
    // EOF return zeros and the bounds check below reports InsufficientData.
    pv.delta_state.resize(n_state);
    for (size_t i = 0; i < n_state; i++) pv.delta_state[i] = r.read_uint(vm.latent_bits);
    for (size_t j = 0; j < ANS_INTERLEAVING; j++) pv.ans_final_state_idxs[j] = AnsState(r.read_uint(vm.ans_size_log));
  };
  if (cm.has_delta_var) one(cm.delta_var, VarKey::Delta, pm.delta_var);
  one(cm.primary, VarKey::Primary, pm.primary);
  if (cm.has_secondary) one(cm.secondary, VarKey::Secondary, pm.secondary);
  r.drain_empty_byte("non-zero bits at end of data page metadata");
  return pm;
}

// ----- standalone framing (standalone/compressor.rs:12-16,85-105; decompressor.rs) ---
inline void write_varint(uint64_t n, BitWriter& w) {
  Bitlen power = n == 0 ? 1 : ilog2_u64(n) + 1;
  w.write_uint(power - 1, BITS_TO_ENCODE_VARINT_POWER);
  w.write_uint(lowest_bits_u64(n, power), power);
}

inline void write_standalone_header(std::vector<uint8_t>& dst, size_t n_hint, uint8_t uniform_type) {
  BitWriter w(dst);
  w.write_aligned_bytes(MAGIC_HEADER, 4);
  w.write_uint(CURRENT_STANDALONE_VERSION, BITS_TO_ENCODE_STANDALONE_VERSION);
  w.write_aligned_bytes(&uniform_type, 1);
  write_varint(n_hint, w);
  w.finish_byte();
  uint8_t ver[2] = {4, 1};  // metadata/format_version.rs:30-34, :87-91
  w.write_aligned_bytes(ver, 2);
  w.finish();
}

struct StandaloneHeader {
  size_t standalone_version = 0;
  uint8_t uniform_type = 0;  // 0 = none
  size_t n_hint = 0;
  FormatVersion format;
};

// standalone/decompressor.rs:85-148 + metadata/format_version.rs:65-85
inline StandaloneHeader read_standalone_header(BitReader& r) {
  StandaloneHeader h;
  const uint8_t* magic = r.read_aligned_bytes(4);
  r.check_in_bounds();
  if (std::memcmp(magic, MAGIC_HEADER, 4) != 0) corruption("magic header does not match");
  h.standalone_version = size_t(r.read_uint(BITS_TO_ENCODE_STANDALONE_VERSION));
  if (h.standalone_version < 2) {
    r.bit_idx -= BITS_TO_ENCODE_STANDALONE_VERSION;  // rewind: byte is the wrapped major version
  } else {
    if (h.standalone_version >= 3) {
      uint8_t byte = r.read_aligned_bytes(1)[0];
      if (byte != MAGIC_TERMINATION_BYTE) {
        if (!number_type_valid(byte)) corruption("unknown number type byte: " + std::to_string(byte));
        h.uniform_type = byte;
      }
    }
    Bitlen power = 1 + Bitlen(r.read_uint(BITS_TO_ENCODE_VARINT_POWER));
    h.n_hint = size_t(r.read_uint(power));
    r.drain_empty_byte("standalone size hint");
  }
  r.check_in_bounds();
  if (h.standalone_version > CURRENT_STANDALONE_VERSION) corruption("file's standalone version exceeds max supported");
  h.format.major = r.read_aligned_bytes(1)[0];
  h.format.minor = h.format.major >= 4 ? r.read_aligned_bytes(1)[0] : 0;
  if (h.format.major > 4) corruption("File's format version definitely cannot be decompressed by this library version");
  r.check_in_bounds();
  return h;
}

}  // namespace pco_oracle
