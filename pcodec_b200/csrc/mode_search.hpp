// ModeSpec::Auto: which mode the reference would pick for one chunk of numbers.  HOST code (no device work): the search reads a
// sample of ~n/40 numbers at data-independent positions and does a few thousand scalar operations on it - planner logic like the
// reference's, not part of the per-number hot path.
//   ints   (pco/src/data_types/unsigned.rs:28-35): IntMult(base) when int_mult::choose_base finds one, else Classic
//   floats (pco/src/data_types/float.rs:70-98):    the best of Classic, FloatMult (trailing-zeros and Euclidean candidates, centred
//                                                  and snapped) and FloatQuant by estimated bits saved; f16 in the half crate's arithmetic
// Sample: pco/src/sampling.rs:62-103 (Floyd's algorithm over Xoroshiro128++ seeded with 0; rand_xoshiro 0.6.0 is a crates.io
// dependency of the reference, restated from its published algorithm).  Where the reference iterates a std HashMap
// (mode/int_mult.rs:187-204, sampling.rs:110-141) its order is random per process; keys are visited in ascending order here, so
// every answer is one the reference can give.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <utility>
#include <vector>

namespace pcob200 {
namespace mode_search {

struct Choice {
  int kind = 0;            // 0 Classic, 1 IntMult, 2 FloatMult, 3 FloatQuant (the format's mode numbers)
  uint64_t int_base = 0;   // IntMult: the base, in latent units
  double base = 0.0;       // FloatMult: base and the inverse the splitter multiplies by (snapping can make inv_base != 1 / base)
  double inv_base = 0.0;
  uint32_t k = 0;          // FloatQuant
  double bits_saved_per_num = 0.0;
};

constexpr size_t MIN_SAMPLE = 10, SAMPLE_RATIO = 40;  // sampling.rs:9-12
constexpr double MULT_REQUIRED_BITS_SAVED = 0.5, QUANT_REQUIRED_BITS_SAVED = 1.5, MEMORIZABLE_BINS = 256.0;  // constants.rs:48-50

// ---- the sample ---------------------------------------------------------------------------------------------------------
struct Xoroshiro128pp {
  uint64_t a, b;
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  static uint64_t splitmix(uint64_t& s) {
    uint64_t z = (s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
  }
  explicit Xoroshiro128pp(uint64_t seed) { a = splitmix(seed); b = splitmix(seed); }
  uint64_t next() {
    const uint64_t out = rotl(a + b, 17) + a;
    b ^= a;
    a = rotl(a, 49) ^ b ^ (b << 21);
    b = rotl(b, 28);
    return out;
  }
};

// positions choose_mode_sample visits, in visiting order (sampling.rs:73-95); empty when n < MIN_SAMPLE
inline std::vector<size_t> sample_positions(size_t n) {
  std::vector<size_t> out;
  if (n < MIN_SAMPLE) return out;
  const size_t target = MIN_SAMPLE + (n - MIN_SAMPLE) / SAMPLE_RATIO;
  Xoroshiro128pp rng(0);
  std::vector<bool> seen(n, false);
  out.reserve(target);
  for (size_t j = n - target; j < n; j++) {
    const size_t t = size_t(rng.next() % (uint64_t(j) + 1));
    const size_t pick = seen[t] ? j : t;
    seen[pick] = true;
    out.push_back(pick);
  }
  return out;
}

// sampling.rs:105-141: (primary, bits saved) per sample element; groups no larger than the cutoff count as savings
template <typename L>
inline double bits_saved_per_num(std::vector<std::pair<L, double>> items) {
  const size_t n = items.size();
  std::stable_sort(items.begin(), items.end(), [](const std::pair<L, double>& x, const std::pair<L, double>& y) { return x.first < y.first; });
  const size_t cutoff = std::max<size_t>(1, size_t(double(n) / MEMORIZABLE_BINS));
  double total = 0.0;
  for (size_t i = 0; i < n;) {
    size_t j = i;
    double group = 0.0;
    for (; j < n && items[j].first == items[i].first; j++) group += items[j].second;
    if (j - i <= cutoff) total += group;
    i = j;
  }
  return total / double(n);
}

// ---- int mult (mode/int_mult.rs) ----------------------------------------------------------------------------------------
template <typename L>
inline L gcd(L x, L y) {  // :57-70
  if (x == 0) return y;
  while (y != 0) {
    x = L(x % y);
    std::swap(x, y);
  }
  return x;
}
template <typename L>
inline L triple_gcd(L a, L b, L c) {  // :99-115
  if (a > b) std::swap(a, b);
  if (b > c) std::swap(b, c);
  if (a > b) std::swap(a, b);
  return gcd<L>(L(b - a), L(c - a));
}
template <typename Fn>
inline bool false_position_root(Fn f, double lb, double ub, double* root) {  // :72-97
  double flb = f(lb), fub = f(ub);
  if (flb > 0.0 || fub < 0.0) return false;
  while (ub - lb > 1E-4 && fub - flb > 0.0) {
    const double lb_prop = 0.001 + 0.998 * fub / (fub - flb);
    const double mid = lb_prop * lb + (1.0 - lb_prop) * ub;
    const double fmid = f(mid);
    if (fmid < 0.0) { lb = mid; flb = fmid; } else { ub = mid; fub = fmid; }
  }
  *root = (lb + ub) / 2.0;
  return true;
}
inline double category_entropy(double p) { return (p == 0.0 || p == 1.0) ? 0.0 : -p * std::log2(p); }  // mode/mod.rs:7-18
inline double worst_case_entropy(double concentrated_p, double others) {
  return category_entropy(concentrated_p) + others * category_entropy((1.0 - concentrated_p) / others);
}
// :117-185: is this gcd statistically there, and how many bits per number would its residues save in the worst case
inline bool score_triple_gcd(double g, size_t with_gcd, size_t triples, double* score) {
  const double zeta2 = 3.14159265358979323846264338327950288 * 3.14159265358979323846264338327950288 / 6.0;
  const double w = double(with_gcd), t = double(triples);
  const double natural = 1.0 / (zeta2 * g * g);
  const double stdev = std::sqrt(natural * (1.0 - natural) / t);
  if ((w / t - natural) / stdev < 3.0) return false;
  const double w_lcb = w - 1.0 * std::sqrt(w);
  if (w_lcb <= 0.0) return false;
  const double congruent = std::fmin(zeta2 * w_lcb / t, 1.0);
  const double gm1 = g - 1.0, gm1_inv_sq = 1.0 / (gm1 * gm1);
  auto cube = [](double x) { return x * x * x; };
  auto f = [&](double p) { return cube(p) + cube(1.0 - p) * gm1_inv_sq - congruent; };
  double p;
  if (!false_position_root(f, 1.0 / g, std::cbrt(congruent) + std::numeric_limits<double>::epsilon(), &p)) return false;
  const double saved = std::log2(g) - worst_case_entropy(p, gm1);
  if (saved < MULT_REQUIRED_BITS_SAVED) return false;
  *score = saved;
  return true;
}
template <typename L>
inline bool candidate_base(const std::vector<L>& sample, L* base, double* saved) {  // :187-214
  std::vector<L> gcds;
  for (size_t i = 0; i + 3 <= sample.size(); i += 3) {
    const L g = triple_gcd<L>(sample[i], sample[i + 1], sample[i + 2]);
    if (g > 1) gcds.push_back(g);
  }
  std::sort(gcds.begin(), gcds.end());
  bool found = false;
  for (size_t i = 0; i < gcds.size();) {
    size_t j = i;
    while (j < gcds.size() && gcds[j] == gcds[i]) j++;
    double s;
    if (score_triple_gcd(double(uint64_t(gcds[i])), j - i, sample.size() / 3, &s) && (!found || s >= *saved)) {
      found = true;
      *base = gcds[i];
      *saved = s;
    }
    i = j;
  }
  return found;
}
template <typename L>
inline bool int_mult_base(const std::vector<L>& sample, L* base) {  // :216-230
  L cand;
  double per_adj;
  if (!candidate_base<L>(sample, &cand, &per_adj)) return false;
  std::vector<std::pair<L, double>> items;
  items.reserve(sample.size());
  for (L x : sample) items.emplace_back(L(x / cand), per_adj);
  if (bits_saved_per_num<L>(std::move(items)) > MULT_REQUIRED_BITS_SAVED) {
    *base = cand;
    return true;
  }
  return false;
}

// ---- floats ---------------------------------------------------------------------------------------------------------------
// f16 as the reference computes with it (half 2.7.1; data_types/float.rs:254-366): every operator widens to f32 and rounds the result
// back to nearest-even, from_f64 rounds once from the double
struct Half { uint16_t bits; };
inline float half_to_float(Half h) {
  const int e = (h.bits >> 10) & 0x1f, m = h.bits & 0x3ff;
  float v = e == 0 ? std::ldexp(float(m), -24) : e == 31 ? (m ? std::numeric_limits<float>::quiet_NaN() : std::numeric_limits<float>::infinity()) : std::ldexp(float(m | 0x400), e - 25);
  return (h.bits & 0x8000) ? -v : v;
}
inline Half half_from_double(double d) {  // one rounding, ties to even (the current rounding mode is round-to-nearest)
  const uint16_t sign = std::signbit(d) ? 0x8000 : 0;
  const double a = std::fabs(d);
  if (std::isnan(a)) return Half{uint16_t(sign | 0x7e00)};
  if (std::isinf(a)) return Half{uint16_t(sign | 0x7c00)};
  if (a == 0.0) return Half{sign};
  int ex;
  std::frexp(a, &ex);
  int e = ex - 1;
  if (e < -14) return Half{uint16_t(sign | uint16_t(std::nearbyint(std::ldexp(a, 24))))};
  double r = std::nearbyint(std::ldexp(a, 10 - e));
  if (r == 2048.0) { r = 1024.0; e += 1; }
  if (e > 15) return Half{uint16_t(sign | 0x7c00)};
  return Half{uint16_t(sign | uint16_t((e + 15) << 10) | uint16_t(uint32_t(r) - 1024))};
}
inline Half operator+(Half a, Half b) { return half_from_double(double(half_to_float(a) + half_to_float(b))); }
inline Half operator-(Half a, Half b) { return half_from_double(double(half_to_float(a) - half_to_float(b))); }
inline Half operator*(Half a, Half b) { return half_from_double(double(half_to_float(a) * half_to_float(b))); }
inline Half operator/(Half a, Half b) { return half_from_double(double(half_to_float(a) / half_to_float(b))); }
inline bool operator<(Half a, Half b) { return half_to_float(a) < half_to_float(b); }
inline bool operator<=(Half a, Half b) { return half_to_float(a) <= half_to_float(b); }
inline bool operator==(Half a, Half b) { return half_to_float(a) == half_to_float(b); }
inline bool operator!=(Half a, Half b) { return half_to_float(a) != half_to_float(b); }

template <typename F> struct Fl;
template <> struct Fl<float> {
  using L = uint32_t;
  static constexpr uint32_t P = 23, BITS = 32;
  static constexpr int32_t BIAS = 127;
  static float of(double x) { return float(x); }
  static double wide(float x) { return double(x); }
  static float of_int(uint32_t l) { return float(l); }
  static float round_(float x) { return std::round(x); }
  static float abs_(float x) { return std::fabs(x); }
  static float max_(float a, float b) { return std::fmax(a, b); }
  static float min_(float a, float b) { return std::fmin(a, b); }
  static bool normal(float x) { return std::isnormal(x); }
  static float sample_cap() { return std::numeric_limits<float>::max() * 0.5f; }  // MAX_FOR_SAMPLING
};
template <> struct Fl<double> {
  using L = uint64_t;
  static constexpr uint32_t P = 52, BITS = 64;
  static constexpr int32_t BIAS = 1023;
  static double of(double x) { return x; }
  static double wide(double x) { return x; }
  static double of_int(uint64_t l) { return double(l); }
  static double round_(double x) { return std::round(x); }
  static double abs_(double x) { return std::fabs(x); }
  static double max_(double a, double b) { return std::fmax(a, b); }
  static double min_(double a, double b) { return std::fmin(a, b); }
  static bool normal(double x) { return std::isnormal(x); }
  static double sample_cap() { return std::numeric_limits<double>::max() * 0.5; }
};
template <> struct Fl<Half> {
  using L = uint16_t;
  static constexpr uint32_t P = 10, BITS = 16;
  static constexpr int32_t BIAS = 15;
  static Half of(double x) { return half_from_double(x); }
  static double wide(Half x) { return double(half_to_float(x)); }
  static Half of_int(uint16_t l) { return half_from_double(double(float(l))); }
  static Half round_(Half x) { return half_from_double(double(std::round(half_to_float(x)))); }
  static Half abs_(Half x) { return Half{uint16_t(x.bits & 0x7fff)}; }
  static Half max_(Half a, Half b) { return half_from_double(double(std::fmax(half_to_float(a), half_to_float(b)))); }
  static Half min_(Half a, Half b) { return half_from_double(double(std::fmin(half_to_float(a), half_to_float(b)))); }
  static bool normal(Half x) { const uint16_t e = x.bits & 0x7c00; return e != 0 && e != 0x7c00; }
  static Half sample_cap() { return Half{30719}; }
};
template <typename F> inline typename Fl<F>::L bits_of(F x) { typename Fl<F>::L b; std::memcpy(&b, &x, sizeof b); return b; }
template <typename F> inline F from_bits(typename Fl<F>::L b) { F x; std::memcpy(&x, &b, sizeof x); return x; }
template <typename F> inline F pow2(int32_t p) {  // data_types/float.rs:158-160
  using L = typename Fl<F>::L;
  return from_bits<F>(L(L(int64_t(Fl<F>::BIAS + p)) << Fl<F>::P));
}
template <typename F> inline int32_t exponent_of(F x) { return int32_t(bits_of<F>(Fl<F>::abs_(x)) >> Fl<F>::P) - Fl<F>::BIAS; }  // :183-185
template <typename L> inline uint32_t ctz(L x) { return x == 0 ? 8 * sizeof(L) : (sizeof(L) == 8 ? uint32_t(__builtin_ctzll(uint64_t(x))) : uint32_t(__builtin_ctz(uint32_t(x)))); }
template <typename L> inline uint32_t clz(L x) { return x == 0 ? 8 * sizeof(L) : (sizeof(L) == 8 ? uint32_t(__builtin_clzll(uint64_t(x))) : uint32_t(__builtin_clz(uint32_t(x))) - uint32_t(32 - 8 * sizeof(L))); }
template <typename F> inline typename Fl<F>::L ordered(F x) {  // data_types/float.rs:402-411
  using L = typename Fl<F>::L;
  const L b = bits_of<F>(x), mid = L(1) << (Fl<F>::BITS - 1);
  return (b & mid) ? L(~b) : L(b ^ mid);
}
template <typename F> inline typename Fl<F>::L int_float_to_latent(F x) {  // data_types/float.rs:229-244
  using L = typename Fl<F>::L;
  const L mid = L(1) << (Fl<F>::BITS - 1), b = bits_of<F>(x), abs_bits = L(b & ~mid);
  const L gpi = L(1) << (Fl<F>::P + 1);
  const F gpi_f = Fl<F>::of_int(gpi), a = from_bits<F>(abs_bits);
  const L abs_int = a < gpi_f ? L(Fl<F>::wide(a)) : L(gpi + (abs_bits - bits_of<F>(gpi_f)));
  return (b & mid) ? L(mid - 1 - abs_int) : L(mid + abs_int);
}

template <typename F> struct MultConfig { F base, inv_base; };
template <typename F> inline MultConfig<F> from_base(F b) { return {b, Fl<F>::of(1.0) / b}; }
template <typename F> inline MultConfig<F> from_inv_base(F i) { return {Fl<F>::of(1.0) / i, i}; }

template <typename F> inline bool approx_zero(F small, F big) { return small <= big * pow2<F>(-int32_t(Fl<F>::P - 6)); }  // mode/float_mult.rs:85-92
template <typename F>
inline bool pair_gcd(F greater, F lesser, F* out) {  // :102-142: Euclid with an error bound carried along
  if (approx_zero<F>(lesser, greater) || lesser == greater) return false;
  const F eps = pow2<F>(-int32_t(Fl<F>::P));
  F gv = greater, ge = Fl<F>::of(0.0), lv = lesser, le = Fl<F>::of(0.0);
  for (;;) {
    const F prev = gv, ratio = Fl<F>::round_(gv / lv);
    ge = ge + (ratio * le + gv * eps);
    gv = Fl<F>::abs_(gv - ratio * lv);
    if (gv <= prev * pow2<F>(-16) || gv <= ge) {
      *out = lv;
      return true;
    }
    if (approx_zero<F>(gv, greater) || gv <= ge * pow2<F>(6)) return false;
    std::swap(gv, lv);
    std::swap(ge, le);
  }
}
template <typename F>
inline bool config_by_trailing_zeros(const std::vector<F>& sample, MultConfig<F>* out) {  // :145-194
  using L = typename Fl<F>::L;
  const uint32_t P = Fl<F>::P, BITS = Fl<F>::BITS;
  auto pow2_divisor = [&](int32_t e, uint32_t tz) { return e - int32_t(P > tz ? P - tz : 0); };
  int32_t k = std::numeric_limits<int32_t>::max();
  size_t count = 0;
  for (F x : sample) {
    const uint32_t tz = ctz(bits_of<F>(x));
    if (x != Fl<F>::of(0.0) && tz >= 5) {
      count++;
      k = std::min(k, pow2_divisor(exponent_of<F>(x), tz));
    }
  }
  const size_t required = std::max<size_t>(size_t(std::ceil(double(sample.size()) * 0.5)), MIN_SAMPLE);
  if (count < required) return false;
  std::vector<L> ints;
  for (F x : sample) {
    const int32_t e = exponent_of<F>(x);
    if (pow2_divisor(e, ctz(bits_of<F>(x))) >= k && e < k + int32_t(BITS)) {
      const L with_mantissa_bit = L(L(bits_of<F>(x) << (BITS - P - 1)) | (L(1) << (BITS - 1)));
      ints.push_back(L(with_mantissa_bit >> (BITS - 1 - uint32_t(e - k))));
    }
  }
  if (ints.size() < required) return false;
  L int_base;
  double unused;
  if (!candidate_base<L>(ints, &int_base, &unused)) int_base = 1;
  *out = from_base<F>(Fl<F>::of_int(int_base) * pow2<F>(k));
  return true;
}
template <typename F>
inline bool sample_gcd_euclidean(const std::vector<F>& sample, F* out) {  // :197-229
  std::vector<F> gcds;
  for (size_t i = 0; i + 1 < sample.size(); i += 2) {
    F g;
    if (pair_gcd<F>(Fl<F>::max_(sample[i], sample[i + 1]), Fl<F>::min_(sample[i], sample[i + 1]), &g)) gcds.push_back(g);
  }
  const size_t required = 1 + size_t(std::ceil(double(sample.size()) * 0.001));
  if (gcds.size() < required) return false;
  std::sort(gcds.begin(), gcds.end());
  for (double percentile : {0.1, 0.3, 0.5}) {
    const F cand = gcds[size_t(percentile * double(gcds.size()))];
    size_t similar = 0;
    for (F g : gcds) similar += Fl<F>::abs_(g - cand) < Fl<F>::of(0.01) * cand;
    if (similar >= required) {
      *out = cand;
      return true;
    }
  }
  return false;
}
template <typename F>
inline F center_base(F base, const std::vector<F>& sample) {  // :239-259
  const F inv = Fl<F>::of(1.0) / base;
  F tweak = Fl<F>::of(0.0), weight_sum = Fl<F>::of(0.0);
  for (F x : sample) {
    const F mult = Fl<F>::round_(x * inv);
    const uint32_t me = uint32_t(exponent_of<F>(mult));
    if (me < Fl<F>::P && mult != Fl<F>::of(0.0)) {
      const F weight = Fl<F>::of(double(Fl<F>::P - me));
      tweak = tweak + weight * (((mult * base) - x) / mult);
      weight_sum = weight_sum + weight;
    }
  }
  return base - tweak / weight_sum;
}
template <typename F>
inline MultConfig<F> snap_to_int_reciprocal(F base) {  // :261-275
  const F inv = Fl<F>::of(1.0) / base, rounded = Fl<F>::round_(inv);
  const F decimal = Fl<F>::of(std::pow(10.0, std::round(std::log10(Fl<F>::wide(inv)))));
  if (Fl<F>::abs_(inv - rounded) < Fl<F>::of(0.02)) return from_inv_base<F>(rounded);
  if (Fl<F>::abs_(inv - decimal) / inv < Fl<F>::of(0.01)) return from_inv_base<F>(decimal);
  return from_base<F>(base);
}
template <typename F>
inline bool mult_savings(const MultConfig<F>& c, const std::vector<F>& sample, double* out) {  // :277-315
  using L = typename Fl<F>::L;
  std::vector<std::pair<L, double>> items;
  items.reserve(sample.size());
  for (F x : sample) {
    const F mult = Fl<F>::round_(x * c.inv_base);
    const uint32_t me = uint32_t(exponent_of<F>(mult));
    const uint32_t inter_base_bits = Fl<F>::P > me ? Fl<F>::P - me : 0;
    const L approx = ordered<F>(mult * c.base), exact = ordered<F>(x);
    const uint32_t adj_bits = 1 + 2 * (Fl<F>::BITS - clz<L>(L(std::max(exact, approx) - std::min(exact, approx))));
    items.emplace_back(int_float_to_latent<F>(mult), double(inter_base_bits) - double(adj_bits));
  }
  const double s = bits_saved_per_num<L>(std::move(items));
  if (s < MULT_REQUIRED_BITS_SAVED) return false;
  *out = s;
  return true;
}
inline uint64_t total_order(double x) { return ordered<double>(x); }  // f64::total_cmp as an integer key
template <typename F>
inline bool float_mult_bid(const std::vector<F>& sample, MultConfig<F>* config, double* saved) {  // :338-358
  bool found = false;
  for (int which = 0; which < 2; which++) {
    MultConfig<F> c{Fl<F>::of(0.0), Fl<F>::of(0.0)};
    double s = 0.0;
    bool ok = which == 0 ? config_by_trailing_zeros<F>(sample, &c) : false;
    if (which == 1) {
      F g;
      ok = sample_gcd_euclidean<F>(sample, &g);
      if (ok) c = snap_to_int_reciprocal<F>(center_base<F>(g, sample));
    }
    if (!ok || !mult_savings<F>(c, sample, &s)) continue;
    if (!found || total_order(s) >= total_order(*saved)) {
      found = true;
      *config = c;
      *saved = s;
    }
  }
  return found;
}
template <typename F>
inline bool float_quant_bid(const std::vector<F>& sample, uint32_t* k_out, double* saved) {  // mode/float_quant.rs:73-151
  using L = typename Fl<F>::L;
  const uint32_t P = Fl<F>::P;
  std::vector<uint32_t> at_least(P + 1, 0);
  for (F x : sample) at_least[std::min<uint32_t>(P, ctz(bits_of<F>(x)))]++;
  for (size_t i = P; i-- > 0;) at_least[i] += at_least[i + 1];
  uint32_t best_k = 0;
  double best = 0.0;
  for (uint32_t k = 1; k <= P; k++) {
    if (at_least[k] == 0) continue;
    const double s = double(k) - worst_case_entropy(double(at_least[k]) / double(sample.size()), double((uint64_t(1) << k) - 1));
    if (!(s > best)) break;
    best_k = k;
    best = s;
  }
  std::vector<std::pair<L, double>> items;
  items.reserve(sample.size());
  for (F x : sample) items.emplace_back(L(bits_of<F>(x) >> best_k), best);
  const double s = bits_saved_per_num<L>(std::move(items));
  if (!(s > QUANT_REQUIRED_BITS_SAVED)) return false;
  *k_out = best_k;
  *saved = s;
  return true;
}

// `sample_bits`: the chunk's numbers at sample_positions(n), in visiting order, as raw bits (e.g. gathered on the device)
template <typename F>
inline Choice choose_float_from_sample(const typename Fl<F>::L* sample_bits, size_t m) {
  Choice best;
  std::vector<F> sample;
  for (size_t j = 0; j < m; j++) {  // data_types/float.rs:70-80: normal, not huge, by magnitude
    const F x = from_bits<F>(sample_bits[j]);
    if (Fl<F>::normal(x) && Fl<F>::abs_(x) <= Fl<F>::sample_cap()) sample.push_back(Fl<F>::abs_(x));
  }
  if (sample.size() < MIN_SAMPLE) return best;
  MultConfig<F> c{Fl<F>::of(0.0), Fl<F>::of(0.0)};
  double s = 0.0;
  uint32_t k = 0;
  // bids in the order classic (0 bits saved), float mult, float quant; the last maximum wins (compression_intermediates.rs:79-84)
  if (float_mult_bid<F>(sample, &c, &s) && total_order(s) >= total_order(best.bits_saved_per_num)) {
    best.kind = 2;
    best.base = Fl<F>::wide(c.base);
    best.inv_base = Fl<F>::wide(c.inv_base);
    best.bits_saved_per_num = s;
  }
  if (float_quant_bid<F>(sample, &k, &s) && total_order(s) >= total_order(best.bits_saved_per_num)) {
    best = Choice();
    best.kind = 3;
    best.k = k;
    best.bits_saved_per_num = s;
  }
  return best;
}
template <typename F>
inline Choice choose_float(const typename Fl<F>::L* num_bits, size_t n) {
  std::vector<typename Fl<F>::L> picked;
  for (size_t i : sample_positions(n)) picked.push_back(num_bits[i]);
  return choose_float_from_sample<F>(picked.data(), picked.size());
}

// `sample_bits`: the chunk's numbers at sample_positions(n) as raw bits of width L; is_signed says how they order (to_latent_ordered)
template <typename L>
inline Choice choose_int_from_sample(const L* sample_bits, size_t m, bool is_signed) {
  Choice best;
  std::vector<L> sample;
  const L mid = L(L(1) << (8 * sizeof(L) - 1));
  for (size_t j = 0; j < m; j++) sample.push_back(is_signed ? L(sample_bits[j] ^ mid) : sample_bits[j]);
  L base;
  if (sample.size() >= MIN_SAMPLE && int_mult_base<L>(sample, &base)) {
    best.kind = 1;
    best.int_base = uint64_t(base);
  }
  return best;
}
// nums: the chunk's numbers as raw bits of width L
template <typename L>
inline Choice choose_int(const L* nums, size_t n, bool is_signed) {
  std::vector<L> picked;
  for (size_t i : sample_positions(n)) picked.push_back(nums[i]);
  return choose_int_from_sample<L>(picked.data(), picked.size(), is_signed);
}

}  // namespace mode_search
}  // namespace pcob200
