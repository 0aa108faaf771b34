// Host orchestration of the compress path: validates the ChunkConfig like the reference, lays out the
// chunks, runs the encode kernels on one stream and returns the .pco bytes (+ optional side index).
#pragma once
#include <cstdlib>
#include <cub/device/device_segmented_radix_sort.cuh>

#include <cmath>

#include "encode_kernels.cuh"
#include "host_common.hpp"
#include "mode_search.hpp"

#ifndef PCOB_ONE_PASS_DEFAULT
#define PCOB_ONE_PASS_DEFAULT true  // PCOB200_ONE_PASS_FRONT_END=0 selects the two-kernel front end (A/B runs)
#endif

namespace pcob200 {

// internal flag of compress_typed (not in the public enum): the call's chunks are the PAGES of one wrapped chunk and share
// bins trained on all of them (pco/src/wrapped/chunk_compressor.rs:129-140)
constexpr uint32_t PCO_B200_INTERNAL_SHARED_BINS = 1u << 16;


struct CompressScratch {
  DevBuf lat0, lat1, keys_a, keys_b, sym0, sym1, ans0, ans1, ob_sum, ans_sum, entries, plans, chunks, starts, seg, cub_tmp, out, small, index, probes, sample, sample_starts, key16_0, key16_1;
  bool plan_attr_set = false, union_attr_set = false;
  void release() {
    for (DevBuf* b : {&lat0, &lat1, &keys_a, &keys_b, &sym0, &sym1, &ans0, &ans1, &ob_sum, &ans_sum, &entries, &plans, &chunks, &starts, &seg, &cub_tmp, &out, &small,
                      &index, &probes, &sample, &sample_starts, &key16_0, &key16_1})
      b->release();
  }
};

// pco/src/wrapped/chunk_compressor.rs:362-371
inline uint32_t choose_unoptimized_bins_log(uint32_t level, size_t n) {
  uint32_t log_n = uint32_t(std::floor(std::log2(double(n))));
  uint32_t fast = log_n >= 4 ? log_n - 4 : 0;
  if (level <= fast) return level;
  return fast + (level - fast) / 2;
}

// pco/src/chunk_config.rs:134-183
inline PcoB200Error n_per_page(const PcoB200ChunkConfig& cfg, size_t n, std::vector<uint64_t>* out) {
  out->clear();
  if (cfg.paging_spec == PCO_B200_PAGING_EXACT) {
    out->assign(cfg.exact_page_ns, cfg.exact_page_ns + cfg.n_exact_pages);
  } else {
    if (n == 0) return PCO_B200_OK;
    size_t max_page_n = cfg.max_page_n == 0 ? (size_t(1) << 18) : size_t(cfg.max_page_n);
    size_t n_pages = (n + max_page_n - 1) / max_page_n;
    size_t low = n / n_pages, r = n % n_pages;
    out->assign(n_pages, low);
    for (size_t i = 0; i < r; i++) (*out)[i] = low + 1;
  }
  uint64_t summed = 0;
  for (uint64_t p : *out) summed += p;
  if (summed != n)
    return fail(PCO_B200_INVALID_ARGUMENT, "paging spec suggests " + std::to_string(summed) + " numbers but " + std::to_string(n) + " were given");
  for (uint64_t p : *out)
    if (p == 0) return fail(PCO_B200_INVALID_ARGUMENT, "cannot write data page of 0 numbers");
  return PCO_B200_OK;
}

// standalone header bytes (pco/src/standalone/compressor.rs:12-16,85-105)
inline std::vector<uint8_t> make_standalone_header(uint64_t n_hint, uint8_t uniform_type) {
  std::vector<uint8_t> h = {112, 99, 111, 33, 3, uniform_type};
  uint32_t power = n_hint == 0 ? 1 : 64 - uint32_t(__builtin_clzll(n_hint));
  unsigned __int128 v = (unsigned __int128)(power - 1) | ((unsigned __int128)(power >= 64 ? n_hint : (n_hint & ((uint64_t(1) << power) - 1))) << 6);
  uint32_t nbytes = (6 + power + 7) / 8;
  for (uint32_t i = 0; i < nbytes; i++) h.push_back(uint8_t(v >> (8 * i)));
  h.push_back(4);
  h.push_back(1);
  return h;
}

__global__ void range_bits_kernel(ChunkEnc* chunks, uint32_t n_chunks, int v, uint32_t* out_bits) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chunks) return;
  uint64_t a = chunks[c].vmin[v], b = chunks[c].vmax[v];
  uint32_t bits = b > a ? 64 - __clzll((long long)(b - a)) : 0;
  atomicMax(out_bits, bits);
  chunks[c].key_base[v] = a;  // plan_probe_kernel's keys are latent - vmin
}

// internal entries [(c, v)][batches_per_chunk] -> compact side index
__global__ void emit_index_kernel(EncParams ep, uint32_t batches_per_chunk, const ChunkEnc* chunks, const BatchEntry* entries, uint8_t* index,
                                  uint64_t chunks_offset, const uint64_t* entry_offsets, const uint64_t* total_bytes, uint32_t has_terminator) {
  const uint32_t c = blockIdx.x;
  if (c == 0 && threadIdx.x == 0) {  // the file size is known on the device first: the host need not wait for it to write the header
    IndexHeader* ih = reinterpret_cast<IndexHeader*>(index);
    ih->file_len = *total_bytes;
    ih->end_byte = has_terminator ? *total_bytes : 0;
  }
  const uint64_t cs = ep.chunk_starts[c], ce = ep.chunk_starts[c + 1];
  const uint32_t n = uint32_t(ce - cs);
  const uint32_t nb = n_batches_of(n);
  const uint32_t n_vars = chunks[c].fallback ? 1 : ep.n_vars;
  if (threadIdx.x == 0) {
    IndexChunk ic;
    ic.chunk_offset = chunks[c].out_offset;
    ic.n = n;
    ic.n_vars = n_vars;
    ic.entries_offset = entry_offsets[c];
    ic.out_offset = cs;
    reinterpret_cast<IndexChunk*>(index + chunks_offset)[c] = ic;
  }
  BatchEntry* dst = reinterpret_cast<BatchEntry*>(index + entry_offsets[c]);
  for (uint32_t i = threadIdx.x; i < n_vars * nb; i += blockDim.x) {
    uint32_t v = i / nb, b = i % nb;
    dst[size_t(v) * nb + b] = entries[(size_t(c) * MAX_VARS + v) * batches_per_chunk + b];
  }
}

struct CompressResult {
  uint64_t total_bytes = 0;
  uint64_t index_bytes = 0;
};

template <typename L>
static PcoB200Error compress_typed(CompressScratch& S, const void* nums, size_t n, uint32_t dtype, const PcoB200ChunkConfig& cfg, bool uniform_type,
                                   void* dst, size_t dst_cap, void* index_dst, size_t index_cap, uint32_t flags, cudaStream_t stream,
                                   CompressResult* res) {
  const bool src_dev = flags & PCO_B200_SRC_ON_DEVICE, dst_dev = flags & PCO_B200_DST_ON_DEVICE;
  const uint32_t lbits = sizeof(L) * 8;
  const bool is_float = nt_is_float(dtype);
  // ---- config validation (pco/src/chunk_config.rs:269-314)
  if (cfg.compression_level > 12) return fail(PCO_B200_INVALID_ARGUMENT, "compression level may not exceed 12");
  if (cfg.delta_spec == PCO_B200_DELTA_TRY_CONSECUTIVE && cfg.delta_order > 7)
    return fail(PCO_B200_INVALID_ARGUMENT, "consecutive delta order may not exceed 7");
  if (lbits == 8 && !cfg.enable_8_bit)
    return fail(PCO_B200_INVALID_ARGUMENT, "compressing 8-bit types with Pco is often a mistake; enable them on the ChunkConfig if you know what you're doing");
  std::vector<uint64_t> pages;
  if (PcoB200Error e = n_per_page(cfg, n, &pages)) return e;
  for (uint64_t p : pages)
    if (p > (uint64_t(1) << 24)) return fail(PCO_B200_INVALID_ARGUMENT, "count may not exceed 16777216 per chunk");
  // ---- specs the GPU hot path implements (Auto, Dict, Lookback, Conv1 are "next")
  EncParams ep;
  std::memset(&ep, 0, sizeof(ep));
  ep.dtype = dtype;
  ep.uniform_type = uniform_type ? dtype : 0;
  switch (cfg.mode_spec) {
    case PCO_B200_MODE_CLASSIC: ep.mode = MODE_CLASSIC; break;
    case PCO_B200_MODE_TRY_INT_MULT:
      if (is_float) return fail(PCO_B200_INVALID_ARGUMENT, "unable to use int mult mode on floats");
      ep.mode = MODE_INT_MULT;
      ep.mode_base = lbits == 64 ? cfg.int_mult_base : (cfg.int_mult_base & ((uint64_t(1) << lbits) - 1));
      if (ep.mode_base == 0) return fail(PCO_B200_INVALID_ARGUMENT, "The chosen mode of IntMult(0) was invalid");
      break;
    case PCO_B200_MODE_TRY_FLOAT_QUANT: {
      if (!is_float) return fail(PCO_B200_INVALID_ARGUMENT, "unable to use float mode for ints");
      uint32_t precision = lbits == 64 ? 52 : lbits == 32 ? 23 : 10;
      if (cfg.float_quant_k == 0 || cfg.float_quant_k > precision) return fail(PCO_B200_INVALID_ARGUMENT, "The chosen mode of FloatQuant was invalid");
      ep.mode = MODE_FLOAT_QUANT;
      ep.mode_k = cfg.float_quant_k;
      break;
    }
    case PCO_B200_MODE_TRY_FLOAT_MULT: {
      if (!is_float) return fail(PCO_B200_INVALID_ARGUMENT, "unable to use float mode for ints");
      if (lbits == 16) return fail(PCO_B200_UNSUPPORTED, "f16 FloatMult is outside the GPU hot path");
      ep.mode = MODE_FLOAT_MULT;
      if (lbits == 64) {
        double base = cfg.float_mult_base, inv = 1.0 / base;
        std::memcpy(&ep.base_bits, &base, 8);
        std::memcpy(&ep.inv_base_bits, &inv, 8);
        if (!std::isfinite(base) || base == 0.0) return fail(PCO_B200_INVALID_ARGUMENT, "The chosen mode of FloatMult was invalid");
        ep.mode_base = (ep.base_bits >> 63) ? ~ep.base_bits : (ep.base_bits ^ (uint64_t(1) << 63));
      } else {
        float base = float(cfg.float_mult_base), inv = 1.0f / base;
        uint32_t bb, ib;
        std::memcpy(&bb, &base, 4);
        std::memcpy(&ib, &inv, 4);
        if (!std::isfinite(base) || base == 0.0f) return fail(PCO_B200_INVALID_ARGUMENT, "The chosen mode of FloatMult was invalid");
        ep.base_bits = bb;
        ep.inv_base_bits = ib;
        ep.mode_base = (bb >> 31) ? uint32_t(~bb) : (bb ^ 0x80000000u);
      }
      break;
    }
    case PCO_B200_MODE_AUTO: {
      // Default: Auto on the GPU path means Classic, which is always valid.  With PCOB200_AUTO_MODE_SEARCH=1 the reference's mode
      // search (mode_search.hpp: int_mult::choose_base, FloatMult / FloatQuant bids on the reference's sample) runs on the host over
      // the call's FIRST chunk and its answer is used for every chunk of the call (the reference searches per chunk; one array is
      // usually one kind of data).  Opt-in until it has been measured on a GPU box.
      ep.mode = MODE_CLASSIC;
      static const bool search = [] { const char* e = std::getenv("PCOB200_AUTO_MODE_SEARCH"); return e && e[0] == '1'; }();
      const bool multi_page_chunk = (flags & PCO_B200_INTERNAL_SHARED_BINS) && pages.size() > 1;  // shared bins exist for one latent var only
      if (search && !pages.empty() && !multi_page_chunk) {
        const size_t n0 = size_t(pages[0]);
        std::vector<L> staged;
        const L* first = static_cast<const L*>(nums);
        if (src_dev) {
          staged.resize(n0);
          PCOB_CUDA_TRY(cudaMemcpyAsync(staged.data(), nums, n0 * sizeof(L), cudaMemcpyDeviceToHost, stream));
          PCOB_CUDA_TRY(cudaStreamSynchronize(stream));
          first = staged.data();
        }
        mode_search::Choice c;
        if constexpr (sizeof(L) == 8) c = is_float ? mode_search::choose_float<double>(first, n0) : mode_search::choose_int<L>(first, n0, nt_is_signed(dtype));
        else if constexpr (sizeof(L) == 4) c = is_float ? mode_search::choose_float<float>(first, n0) : mode_search::choose_int<L>(first, n0, nt_is_signed(dtype));
        else if constexpr (sizeof(L) == 2) {
          c = is_float ? mode_search::choose_float<mode_search::Half>(first, n0) : mode_search::choose_int<L>(first, n0, nt_is_signed(dtype));
          if (is_float && c.kind == 2) c = mode_search::Choice();  // there is no f16 FloatMult kernel: Classic instead
        } else c = mode_search::choose_int<L>(first, n0, nt_is_signed(dtype));
        if (c.kind == 1) {
          ep.mode = MODE_INT_MULT;
          ep.mode_base = c.int_base;
        } else if (c.kind == 3) {
          ep.mode = MODE_FLOAT_QUANT;
          ep.mode_k = c.k;
        } else if (c.kind == 2 && lbits == 64) {
          ep.mode = MODE_FLOAT_MULT;
          std::memcpy(&ep.base_bits, &c.base, 8);
          std::memcpy(&ep.inv_base_bits, &c.inv_base, 8);
          ep.mode_base = (ep.base_bits >> 63) ? ~ep.base_bits : (ep.base_bits ^ (uint64_t(1) << 63));
        } else if (c.kind == 2) {
          const float base = float(c.base), inv = float(c.inv_base);
          uint32_t bb, ib;
          std::memcpy(&bb, &base, 4);
          std::memcpy(&ib, &inv, 4);
          ep.mode = MODE_FLOAT_MULT;
          ep.base_bits = bb;
          ep.inv_base_bits = ib;
          ep.mode_base = (bb >> 31) ? uint32_t(~bb) : (bb ^ 0x80000000u);
        }
      }
      break;
    }
    default: return fail(PCO_B200_UNSUPPORTED, "ModeSpec::TryDict is outside the GPU hot path");
  }
  bool auto_delta = false;
  switch (cfg.delta_spec) {
    case PCO_B200_DELTA_NOOP: ep.order = 0; break;
    case PCO_B200_DELTA_TRY_CONSECUTIVE: ep.order = cfg.delta_order; break;
    case PCO_B200_DELTA_AUTO: auto_delta = true; ep.order = 0; break;  // resolved below by the sampled order search
    case PCO_B200_DELTA_TRY_CONV1:
      if (cfg.delta_order == 0) { ep.order = 0; break; }
      return fail(PCO_B200_UNSUPPORTED, "DeltaSpec::TryConv1 is outside the GPU hot path");
    default: return fail(PCO_B200_UNSUPPORTED, "DeltaSpec::TryLookback is outside the GPU hot path");
  }
  ep.n_vars = ep.mode == MODE_CLASSIC ? 1 : 2;
  ep.n_total = n;
  ep.n_chunks = uint32_t(pages.size());
  const bool chunks_only = flags & PCO_B200_CHUNKS_ONLY;
  std::vector<uint8_t> header = make_standalone_header(n, uint8_t(ep.uniform_type));
  if (chunks_only) header.clear();
  // empty input: header + terminator only (standalone/simple.rs:62-91)
  if (n == 0) {
    if (!chunks_only) header.push_back(0);
    if (header.size() > dst_cap) return fail(PCO_B200_IO, "failed to write whole buffer");
    if (dst_dev && !header.empty()) PCOB_CUDA_TRY(cudaMemcpyAsync(dst, header.data(), header.size(), cudaMemcpyHostToDevice, stream));
    else if (!header.empty()) std::memcpy(dst, header.data(), header.size());
    PCOB_CUDA_TRY(cudaStreamSynchronize(stream));
    res->total_bytes = header.size();
    if (index_dst && index_cap >= sizeof(IndexHeader)) {
      IndexHeader ih;
      std::memset(&ih, 0, sizeof(ih));
      ih.magic = INDEX_MAGIC; ih.version = 1; ih.file_len = header.size(); ih.chunks_offset = sizeof(IndexHeader); ih.end_byte = header.size();
      std::memcpy(index_dst, &ih, sizeof(ih));
      res->index_bytes = sizeof(ih);
    }
    return PCO_B200_OK;
  }
  std::vector<uint64_t> starts(pages.size() + 1, 0);
  uint64_t max_chunk_n = 0;
  for (size_t i = 0; i < pages.size(); i++) { starts[i + 1] = starts[i] + pages[i]; max_chunk_n = std::max<uint64_t>(max_chunk_n, pages[i]); }
  ep.max_chunk_n = uint32_t(max_chunk_n);
  // unoptimized_bins_log is a function of each chunk's n; the kernels take one value per call, so every chunk must agree
  const bool shared_bins = (flags & PCO_B200_INTERNAL_SHARED_BINS) && pages.size() > 1;
  uint32_t bins_log = choose_unoptimized_bins_log(cfg.compression_level, shared_bins ? n : size_t(pages[0]));
  if (!shared_bins)
    for (uint64_t p : pages)
      if (choose_unoptimized_bins_log(cfg.compression_level, size_t(p)) != bins_log)
        return fail(PCO_B200_UNSUPPORTED, "chunks whose sizes imply different unoptimized_bins_log in one call");
  if (shared_bins && ep.mode != MODE_CLASSIC)
    return fail(PCO_B200_UNSUPPORTED, "wrapped chunks with several pages: classic mode only, for now");
  if (bins_log > 8) return fail(PCO_B200_UNSUPPORTED, "compression levels that train more than 256 bins are outside the GPU hot path");
  ep.bins_log[0] = bins_log;
  ep.bins_log[1] = std::min<uint32_t>(bins_log, 6);  // LIMITED_UNOPTIMIZED_BINS_LOG (chunk_compressor.rs:238-248)

  // ---- device buffers
  const uint32_t n_chunks = ep.n_chunks;
  const uint32_t bpc = n_batches_of(uint32_t(max_chunk_n));
  const uint32_t tiles_per_chunk = uint32_t((max_chunk_n + SPLIT_TILE - 1) / SPLIT_TILE);
  const void* d_nums = nums;
  // every chunk's rows start on a 256-slot boundary in the latent / symbol / ans arrays (vector accesses per batch row)
  std::vector<uint64_t> rows(pages.size() + 1, 0);
  for (size_t i = 0; i < pages.size(); i++) rows[i + 1] = rows[i] + ((pages[i] + BATCH_N - 1) / BATCH_N) * BATCH_N;
  const size_t n_slots = size_t(rows.back());
  PCOB_CUDA_TRY(S.lat0.reserve(n_slots * sizeof(L) + 64));
  if (ep.n_vars > 1) PCOB_CUDA_TRY(S.lat1.reserve(n_slots * sizeof(L) + 64));
  PCOB_CUDA_TRY(S.sym0.reserve(n_slots + 64));
  PCOB_CUDA_TRY(S.ans0.reserve(n_slots * 2 + 64));
  if (ep.n_vars > 1) { PCOB_CUDA_TRY(S.sym1.reserve(n_slots + 64)); PCOB_CUDA_TRY(S.ans1.reserve(n_slots * 2 + 64)); }
  const size_t n_cvb = size_t(n_chunks) * MAX_VARS * bpc;
  PCOB_CUDA_TRY(S.ob_sum.reserve(n_cvb * 4));
  PCOB_CUDA_TRY(S.ans_sum.reserve(n_cvb * 4));
  PCOB_CUDA_TRY(S.entries.reserve(n_cvb * sizeof(BatchEntry)));
  PCOB_CUDA_TRY(S.plans.reserve(size_t(n_chunks) * MAX_VARS * sizeof(VarPlan)));
  PCOB_CUDA_TRY(S.chunks.reserve(size_t(n_chunks) * sizeof(ChunkEnc)));
  PCOB_CUDA_TRY(S.starts.reserve(starts.size() * 16));
  PCOB_CUDA_TRY(S.seg.reserve(size_t(n_chunks) * 16));
  PCOB_CUDA_TRY(S.small.reserve(256 + header.size()));
  // a dedicated input staging buffer when nums live on the host (kept apart from the sort buffers)
  DevBuf& in_stage = S.index;
  if (!src_dev) {
    PCOB_CUDA_TRY(in_stage.reserve(n * sizeof(L) + 64));
    PCOB_CUDA_TRY(cudaMemcpyAsync(in_stage.p, nums, n * sizeof(L), cudaMemcpyHostToDevice, stream));
    d_nums = in_stage.p;
  }
  ep.nums = d_nums;
  PCOB_CUDA_TRY(cudaMemcpyAsync(S.starts.p, starts.data(), starts.size() * 8, cudaMemcpyHostToDevice, stream));
  ep.chunk_starts = S.starts.as<uint64_t>();
  PCOB_CUDA_TRY(cudaMemcpyAsync(S.starts.as<uint64_t>() + starts.size(), rows.data(), rows.size() * 8, cudaMemcpyHostToDevice, stream));
  ep.row_base = S.starts.as<uint64_t>() + starts.size();
  ChunkEnc* d_chunks = S.chunks.as<ChunkEnc>();
  VarPlan* d_plans = S.plans.as<VarPlan>();
  PCOB_CUDA_TRY(S.probes.reserve(size_t(n_chunks) * sizeof(PlanProbes)));
  PlanProbes* d_probes = S.probes.as<PlanProbes>();
  L* d_lat[2] = {S.lat0.as<L>(), S.lat1.as<L>()};
  uint8_t* d_sym[2] = {S.sym0.as<uint8_t>(), S.sym1.as<uint8_t>()};
  uint16_t* d_ans[2] = {S.ans0.as<uint16_t>(), S.ans1.as<uint16_t>()};
  uint32_t* d_small = S.small.as<uint32_t>();  // [0]: range bits, [2..3]: total bytes (u64), header at byte 64
  uint64_t* d_total = reinterpret_cast<uint64_t*>(d_small + 2);
  uint8_t* d_header = reinterpret_cast<uint8_t*>(d_small) + 64;
  PCOB_CUDA_TRY(cudaMemcpyAsync(d_header, header.data(), header.size(), cudaMemcpyHostToDevice, stream));
  PCOB_CUDA_TRY(cudaMemsetAsync(S.ob_sum.p, 0, n_cvb * 4, stream));

  // ---- K1+K2 and the planner for one set of chunks (the call's chunks, or their samples during the Auto delta search)
  uint32_t var_range_bits[MAX_VARS] = {64, 64};
  auto front = [&](const EncParams& e, uint32_t tiles, size_t slots, uint32_t (&vrb)[MAX_VARS], const std::vector<uint64_t>& sizes) -> PcoB200Error {
    init_chunks_kernel<<<(n_chunks + 255) / 256, 256, 0, stream>>>(d_chunks, n_chunks);
    if (!S.plan_attr_set) {
      S.plan_attr_set = true;
      PCOB_CUDA_TRY(cudaFuncSetAttribute(plan_probe_kernel<L, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(((size_t(1) << PLAN_MAX_COUNT_BITS) + 1) * 4 + 16)));
      PCOB_CUDA_TRY(cudaFuncSetAttribute(split_count_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(size_t(SC_N) * 4)));
      PCOB_CUDA_TRY(cudaFuncSetAttribute(pack_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PackSmem)));
      // several counting CTAs per SM: ask for the largest shared-memory carveout
      PCOB_CUDA_TRY(cudaFuncSetAttribute(plan_probe_kernel<L, true>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
      PCOB_CUDA_TRY(cudaFuncSetAttribute(pack_kernel<L>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
      PCOB_CUDA_TRY(cudaFuncSetAttribute(bin_lut_kernel<L>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    }
#ifndef PCOB_NO_SPLIT_COUNT
    static const bool one_pass = [] { const char* v = std::getenv("PCOB200_ONE_PASS_FRONT_END"); return v ? v[0] != '0' : PCOB_ONE_PASS_DEFAULT; }();
    if (one_pass && e.mode == MODE_CLASSIC && e.n_vars == 1) {
      // one-pass front end: split + delta + counting histogram + 16-bit keys, no 64-bit latents (split_count_kernel).
      // It assumes every chunk's stored latents span < 2^15; a chunk that does not raises flags[1] and the call is redone
      // on the two-kernel path below (the speculation costs one read of the input).
      PCOB_CUDA_TRY(S.key16_0.reserve(slots * 2 + 64));
      PCOB_CUDA_TRY(cudaMemsetAsync(d_small, 0, 8, stream));
      profiler().begin("split_count_kernel", stream);
      split_count_kernel<L><<<n_chunks, SC_THREADS, size_t(SC_N) * 4, stream>>>(e, d_chunks, d_probes, S.key16_0.as<uint16_t>(), d_small);
      profiler().end(stream);
      uint32_t fl[2] = {0, 0};
      PCOB_CUDA_TRY(cudaMemcpyAsync(fl, d_small, 8, cudaMemcpyDeviceToHost, stream));
      PCOB_CUDA_TRY(cudaStreamSynchronize(stream));
      if (fl[1] == 0 && shared_bins) {
        // pages of one chunk: one histogram over all of them, one plan, copied to every page's slot
        if (!S.union_attr_set) {
          S.union_attr_set = true;
          PCOB_CUDA_TRY(cudaFuncSetAttribute(union_probe_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(size_t(SC_N) * 4)));
        }
        PCOB_CUDA_TRY(cudaMemsetAsync(d_small, 0, 8, stream));
        union_probe_kernel<L><<<1, SC_THREADS, size_t(SC_N) * 4, stream>>>(e, n_chunks, d_chunks, d_probes, S.key16_0.as<uint16_t>(), d_small);
        PCOB_CUDA_TRY(cudaMemcpyAsync(fl, d_small, 8, cudaMemcpyDeviceToHost, stream));
        PCOB_CUDA_TRY(cudaStreamSynchronize(stream));
        if (fl[1] == 0) {
          vrb[0] = fl[0];
          uint64_t stored_total = 0;
          for (uint64_t pn : sizes) stored_total += pn > e.order ? pn - e.order : 0;
          plan_solve_kernel<L><<<1, SOLVE_THREADS, 0, stream>>>(e, d_probes, d_chunks, d_plans, 0, uint32_t(stored_total));
          broadcast_plan_kernel<<<n_chunks, 128, 0, stream>>>(d_plans, n_chunks);
          return PCO_B200_OK;
        }
        fl[1] = 1;  // every page is narrow but their union is not: the two-kernel path below (it needs the 64-bit latents)
      }
      if (fl[1] == 0) {
        vrb[0] = fl[0];
        profiler().begin("plan_solve_kernel", stream);
        plan_solve_kernel<L><<<n_chunks, SOLVE_THREADS, 0, stream>>>(e, d_probes, d_chunks, d_plans, 0);
        profiler().end(stream);
        return PCO_B200_OK;
      }
      init_chunks_kernel<<<(n_chunks + 255) / 256, 256, 0, stream>>>(d_chunks, n_chunks);
    }
#endif
    profiler().begin("split_delta_kernel", stream);
    switch (e.mode) {
      case MODE_CLASSIC: split_delta_kernel<L, MODE_CLASSIC><<<n_chunks * tiles, SPLIT_THREADS, 0, stream>>>(e, tiles, d_lat[0], d_lat[1], d_chunks); break;
      case MODE_INT_MULT: split_delta_kernel<L, MODE_INT_MULT><<<n_chunks * tiles, SPLIT_THREADS, 0, stream>>>(e, tiles, d_lat[0], d_lat[1], d_chunks); break;
      case MODE_FLOAT_QUANT: split_delta_kernel<L, MODE_FLOAT_QUANT><<<n_chunks * tiles, SPLIT_THREADS, 0, stream>>>(e, tiles, d_lat[0], d_lat[1], d_chunks); break;
      default: split_delta_kernel<L, MODE_FLOAT_MULT><<<n_chunks * tiles, SPLIT_THREADS, 0, stream>>>(e, tiles, d_lat[0], d_lat[1], d_chunks); break;
    }
    profiler().end(stream);
    // ---- planner per var: range-reduced keys -> segmented radix sort over the significant bits -> plan
    if (shared_bins) {
      // pages of one wrapped chunk on the sort path: common minimum, the pages' keys as one gap-free segment, one sort,
      // the union's probes and plan, copied to every page's slot (classic mode: one latent var)
      PCOB_CUDA_TRY(cudaMemsetAsync(d_small, 0, 4, stream));
      union_range_kernel<<<1, 32, 0, stream>>>(e, n_chunks, d_chunks, d_small);
      uint32_t range_bits = 0;
      PCOB_CUDA_TRY(cudaMemcpyAsync(&range_bits, d_small, 4, cudaMemcpyDeviceToHost, stream));
      PCOB_CUDA_TRY(cudaStreamSynchronize(stream));
      vrb[0] = std::max<uint32_t>(range_bits, PLAN_MAX_COUNT_BITS + 1);  // binning and packing read the 64-bit latents on this path
      std::vector<uint64_t> prefix(sizes.size() + 1, 0);
      for (size_t i = 0; i < sizes.size(); i++) prefix[i + 1] = prefix[i] + (sizes[i] > e.order ? sizes[i] - e.order : 0);
      const uint64_t stored_total = prefix.back();
      PCOB_CUDA_TRY(S.keys_a.reserve(slots * sizeof(L) + 64));
      PCOB_CUDA_TRY(S.keys_b.reserve(slots * sizeof(L) + 64));
      PCOB_CUDA_TRY(S.seg.reserve((prefix.size() + 2) * 8 + 64));
      uint64_t* d_prefix = S.seg.as<uint64_t>();  // [0 .. n_pages]: compaction offsets; then {segment begin, segment end}
      PCOB_CUDA_TRY(cudaMemcpyAsync(d_prefix, prefix.data(), prefix.size() * 8, cudaMemcpyHostToDevice, stream));
      const uint64_t seg_host[2] = {0, stored_total};
      uint64_t* d_seg = d_prefix + prefix.size();
      PCOB_CUDA_TRY(cudaMemcpyAsync(d_seg, seg_host, 16, cudaMemcpyHostToDevice, stream));
      sort_keys_kernel<L><<<n_chunks * tiles, 256, 0, stream>>>(e, tiles, d_lat[0], S.keys_a.as<L>(), d_chunks, 0, d_prefix);
      const L* sorted = S.keys_a.as<L>();
      if (stored_total > 0) {
        cub::DoubleBuffer<L> db(S.keys_a.as<L>(), S.keys_b.as<L>());
        size_t tmp_bytes = 0;
        const int end_bit = int(std::max<uint32_t>(range_bits, 1));
        PCOB_CUDA_TRY(cub::DeviceSegmentedRadixSort::SortKeys(nullptr, tmp_bytes, db, int64_t(stored_total), int64_t(1), d_seg, d_seg + 1, 0, end_bit, stream));
        PCOB_CUDA_TRY(S.cub_tmp.reserve(tmp_bytes + 16));
        PCOB_CUDA_TRY(cub::DeviceSegmentedRadixSort::SortKeys(S.cub_tmp.p, tmp_bytes, db, int64_t(stored_total), int64_t(1), d_seg, d_seg + 1, 0, end_bit, stream));
        sorted = db.Current();
      }
      plan_probe_kernel<L, false><<<1, PLAN_THREADS, 16, stream>>>(e, sorted, d_chunks, d_probes, 0, range_bits, nullptr, uint32_t(stored_total));
      plan_solve_kernel<L><<<1, SOLVE_THREADS, 0, stream>>>(e, d_probes, d_chunks, d_plans, 0, uint32_t(stored_total));
      broadcast_plan_kernel<<<n_chunks, 128, 0, stream>>>(d_plans, n_chunks);
      return PCO_B200_OK;
    }
    for (uint32_t v = 0; v < e.n_vars; v++) {
      const uint32_t order_v = v == 0 ? e.order : 0;
      PCOB_CUDA_TRY(cudaMemsetAsync(d_small, 0, 4, stream));
      range_bits_kernel<<<(n_chunks + 255) / 256, 256, 0, stream>>>(d_chunks, n_chunks, int(v), d_small);
      uint32_t range_bits = 0;
      PCOB_CUDA_TRY(cudaMemcpyAsync(&range_bits, d_small, 4, cudaMemcpyDeviceToHost, stream));
      PCOB_CUDA_TRY(cudaStreamSynchronize(stream));
      vrb[v] = range_bits;
      if (range_bits <= PLAN_MAX_COUNT_BITS) {
        // small key range (the usual case once the chunk minimum is subtracted): counting histogram, no sort
        const size_t smem = ((size_t(1) << range_bits) + 1) * 4 + 16;
        profiler().begin("plan_probe_kernel_counting", stream);
        PCOB_CUDA_TRY((v == 0 ? S.key16_0 : S.key16_1).reserve(slots * 2 + 64));
      plan_probe_kernel<L, true><<<n_chunks, PLAN_THREADS, smem, stream>>>(e, d_lat[v], d_chunks, d_probes, int(v), range_bits,
                                                                            (v == 0 ? S.key16_0 : S.key16_1).as<uint16_t>());
        profiler().end(stream);
        profiler().begin("plan_solve_kernel", stream);
        plan_solve_kernel<L><<<n_chunks, SOLVE_THREADS, 0, stream>>>(e, d_probes, d_chunks, d_plans, int(v));
        profiler().end(stream);
        continue;
      }
      // wide key range: sort the range-reduced keys (the two key buffers exist only on this path)
      PCOB_CUDA_TRY(S.keys_a.reserve(slots * sizeof(L) + 64));
      PCOB_CUDA_TRY(S.keys_b.reserve(slots * sizeof(L) + 64));
      profiler().begin("sort_keys_kernel", stream);
      sort_keys_kernel<L><<<n_chunks * tiles, 256, 0, stream>>>(e, tiles, d_lat[v], S.keys_a.as<L>(), d_chunks, int(v));
      profiler().end(stream);
      uint64_t* seg_begin = S.seg.as<uint64_t>();
      uint64_t* seg_end = seg_begin + n_chunks;
      segment_offsets_kernel<<<(n_chunks + 255) / 256, 256, 0, stream>>>(e, order_v, seg_begin, seg_end);
      const L* sorted = S.keys_a.as<L>();
      {
        cub::DoubleBuffer<L> db(S.keys_a.as<L>(), S.keys_b.as<L>());
        size_t tmp_bytes = 0;
        PCOB_CUDA_TRY(cub::DeviceSegmentedRadixSort::SortKeys(nullptr, tmp_bytes, db, int64_t(slots), int64_t(n_chunks), seg_begin, seg_end, 0,
                                                              int(range_bits), stream));
        PCOB_CUDA_TRY(S.cub_tmp.reserve(tmp_bytes + 16));
        profiler().begin("cub_segmented_radix_sort", stream);
        PCOB_CUDA_TRY(cub::DeviceSegmentedRadixSort::SortKeys(S.cub_tmp.p, tmp_bytes, db, int64_t(slots), int64_t(n_chunks), seg_begin, seg_end, 0,
                                                              int(range_bits), stream));
        profiler().end(stream);
        sorted = db.Current();
      }
      profiler().begin("plan_probe_kernel_sorted", stream);
      plan_probe_kernel<L, false><<<n_chunks, PLAN_THREADS, 16, stream>>>(e, sorted, d_chunks, d_probes, int(v), range_bits, nullptr);
      profiler().end(stream);
      profiler().begin("plan_solve_kernel", stream);
      plan_solve_kernel<L><<<n_chunks, SOLVE_THREADS, 0, stream>>>(e, d_probes, d_chunks, d_plans, int(v));
      profiler().end(stream);
    }
    return PCO_B200_OK;
  };
  if (auto_delta) {
    // ---- DeltaSpec::Auto: sampled search over consecutive orders (see gather_sample_kernel)
    std::vector<uint64_t> s_starts(pages.size() + 1, 0), s_rows(pages.size() + 1, 0), s_sizes(pages.size(), 0);
    uint64_t max_ns = 0;
    for (size_t i = 0; i < pages.size(); i++) {
      const SampleGeom g = delta_sample_geom(pages[i]);
      const uint64_t ns = uint64_t(g.n_groups) * g.group_n;
      s_starts[i + 1] = s_starts[i] + ns;
      s_sizes[i] = ns;
      s_rows[i + 1] = s_rows[i] + ((ns + BATCH_N - 1) / BATCH_N) * BATCH_N;
      max_ns = std::max(max_ns, ns);
    }
    if (s_starts.back() > 0) {
      PCOB_CUDA_TRY(S.sample.reserve(s_starts.back() * sizeof(L) + 64));
      PCOB_CUDA_TRY(S.sample_starts.reserve(s_starts.size() * 16 + 16));
      uint64_t* d_ss = S.sample_starts.as<uint64_t>();
      PCOB_CUDA_TRY(cudaMemcpyAsync(d_ss, s_starts.data(), s_starts.size() * 8, cudaMemcpyHostToDevice, stream));
      PCOB_CUDA_TRY(cudaMemcpyAsync(d_ss + s_starts.size(), s_rows.data(), s_rows.size() * 8, cudaMemcpyHostToDevice, stream));
      gather_sample_kernel<L><<<n_chunks, 256, 0, stream>>>(static_cast<const L*>(d_nums), ep.chunk_starts, d_ss, S.sample.as<L>());
      EncParams es = ep;
      es.nums = S.sample.p;
      es.chunk_starts = d_ss;
      es.row_base = d_ss + s_starts.size();
      es.n_total = s_starts.back();
      es.max_chunk_n = uint32_t(max_ns);
      const uint32_t s_tiles = uint32_t((max_ns + SPLIT_TILE - 1) / SPLIT_TILE);
      unsigned long long* d_cost = reinterpret_cast<unsigned long long*>(d_small + 4);
      unsigned long long best_cost = 0;
      uint32_t best_order = 0;
      for (uint32_t k = 0; k <= MAX_ORDER; k++) {
        es.order = k;
        uint32_t vrb[MAX_VARS] = {64, 64};
        if (PcoB200Error e = front(es, s_tiles, size_t(s_rows.back()), vrb, s_sizes)) return e;
        PCOB_CUDA_TRY(cudaMemsetAsync(d_cost, 0, 8, stream));
        auto_cost_kernel<<<(n_chunks + 255) / 256, 256, 0, stream>>>(es, d_plans, d_cost);
        unsigned long long cost = 0;
        PCOB_CUDA_TRY(cudaMemcpyAsync(&cost, d_cost, 8, cudaMemcpyDeviceToHost, stream));
        PCOB_CUDA_TRY(cudaStreamSynchronize(stream));
        if (k == 0 || cost < best_cost) { best_cost = cost; best_order = k; }
        else break;  // "it's almost always convex" (chunk_compressor.rs:347-357)
      }
      ep.order = best_order;
    }
  }
  if (PcoB200Error e = front(ep, tiles_per_chunk, n_slots, var_range_bits, pages)) return e;
  fallback_kernel<<<(n_chunks + 255) / 256, 256, 0, stream>>>(ep, d_plans, d_chunks, shared_bins ? n_chunks : 0u);
  // ---- K3, K4
  const uint32_t groups_per_chunk = (bpc + 7) / 8;
  for (uint32_t v = 0; v < ep.n_vars; v++)
  {
    if (var_range_bits[v] <= PLAN_MAX_COUNT_BITS) {
      const uint32_t parts = (bpc + BINL_BATCHES - 1) / BINL_BATCHES;
      profiler().begin("bin_lut_kernel", stream);
      bin_lut_kernel<L><<<n_chunks * parts, BINL_THREADS, (size_t(1) << var_range_bits[v]) + 16, stream>>>(ep, bpc, parts, (v == 0 ? S.key16_0 : S.key16_1).as<uint16_t>(), d_plans, d_chunks, d_sym[v],
                                                                                                           S.ob_sum.as<uint32_t>(), int(v), var_range_bits[v]);
      profiler().end(stream);
      continue;
    }
    profiler().begin("bin_kernel", stream);
    bin_kernel<L><<<n_chunks * groups_per_chunk, BIN_THREADS, 0, stream>>>(ep, bpc, d_lat[v], d_plans, d_chunks, d_sym[v], S.ob_sum.as<uint32_t>(), int(v));
    profiler().end(stream);
  }
  // the ans kernel indexes (chunk, var) by blockIdx; both vars share the launch via separate symbol arrays
  profiler().begin("ans_encode_kernel", stream);
  ans_encode_kernel<<<n_chunks * MAX_VARS, ANS_THREADS, 0, stream>>>(ep, bpc, d_plans, d_chunks, d_sym[0], d_sym[1], d_ans[0], d_ans[1], S.ans_sum.as<uint32_t>(),
                                                                     S.entries.as<BatchEntry>());
  profiler().end(stream);
  // ---- layout, offsets, K5
  profiler().begin("layout_kernel", stream);
  layout_kernel<<<n_chunks, LAYOUT_THREADS, 0, stream>>>(ep, bpc, d_plans, d_chunks, S.ans_sum.as<uint32_t>(), S.ob_sum.as<uint32_t>(), S.entries.as<BatchEntry>());
  profiler().end(stream);
  chunk_offsets_kernel<<<1, 1024, 0, stream>>>(d_chunks, n_chunks, header.size(), chunks_only ? 0u : 1u, d_total);
  uint64_t total = 0;
  uint8_t* d_out = static_cast<uint8_t*>(dst);
  if (!dst_dev) {
    // the staging buffer is sized from the file size, so that is needed first
    PCOB_CUDA_TRY(cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, stream));
    PCOB_CUDA_TRY(cudaStreamSynchronize(stream));
    PCOB_CUDA_TRY(cudaGetLastError());
    if (total > dst_cap) return fail(PCO_B200_IO, "failed to write whole buffer (need " + std::to_string(total) + " bytes, dst_cap " + std::to_string(dst_cap) + ")");
    PCOB_CUDA_TRY(S.out.reserve(total + 64));
    d_out = S.out.as<uint8_t>();
  }
  // device destination: no host round trip here - pack_kernel leaves out any chunk that would not fit dst_cap, and the size
  // is checked when it is read back behind the kernel
  const uint64_t out_cap = dst_dev ? uint64_t(dst_cap) : total;
  profiler().begin("pack_kernel", stream);
  pack_kernel<L><<<n_chunks, PACK_THREADS, sizeof(PackSmem), stream>>>(ep, bpc, d_lat[0], d_lat[1], d_plans, d_chunks, d_sym[0], d_sym[1], d_ans[0], d_ans[1],
                                                                      S.entries.as<BatchEntry>(), d_out, out_cap,
                                                                      var_range_bits[0] <= PLAN_MAX_COUNT_BITS ? S.key16_0.as<uint16_t>() : nullptr,
                                                                      (ep.n_vars > 1 && var_range_bits[1] <= PLAN_MAX_COUNT_BITS) ? S.key16_1.as<uint16_t>() : nullptr);
  profiler().end(stream);
  if (!chunks_only) header_footer_kernel<<<1, 32, 0, stream>>>(d_out, out_cap, d_header, uint32_t(header.size()), d_total);
  PCOB_CUDA_TRY(cudaGetLastError());
  if (!dst_dev) PCOB_CUDA_TRY(cudaMemcpyAsync(dst, d_out, total, cudaMemcpyDeviceToHost, stream));
  // ---- optional side index
  if (index_dst != nullptr) {
    const uint64_t chunks_offset = sizeof(IndexHeader);
    std::vector<uint64_t> eoff(n_chunks);
    uint64_t off = (chunks_offset + uint64_t(n_chunks) * sizeof(IndexChunk) + 15) & ~uint64_t(15);
    for (uint32_t c = 0; c < n_chunks; c++) {
      eoff[c] = off;
      off += (uint64_t(ep.n_vars) * n_batches_of(uint32_t(pages[c])) * sizeof(BatchEntry) + 15) & ~uint64_t(15);
    }
    if (off > index_cap) return fail(PCO_B200_IO, "index buffer too small (need " + std::to_string(off) + " bytes)");
    // assemble on the device (the input staging buffer is free again), then one copy out
    DevBuf& idx = S.keys_b;  // sort buffers are idle now
    PCOB_CUDA_TRY(idx.reserve(off + 64));
    PCOB_CUDA_TRY(S.seg.reserve(size_t(n_chunks) * 16));
    PCOB_CUDA_TRY(cudaMemcpyAsync(S.seg.p, eoff.data(), size_t(n_chunks) * 8, cudaMemcpyHostToDevice, stream));
    IndexHeader ih;
    std::memset(&ih, 0, sizeof(ih));
    ih.magic = INDEX_MAGIC; ih.version = 1; ih.n_chunks = n_chunks; ih.n_total = n; ih.chunks_offset = chunks_offset;  // file_len, end_byte: emit_index_kernel
    PCOB_CUDA_TRY(cudaMemcpyAsync(idx.p, &ih, sizeof(ih), cudaMemcpyHostToDevice, stream));
    emit_index_kernel<<<n_chunks, 256, 0, stream>>>(ep, bpc, d_chunks, S.entries.as<BatchEntry>(), idx.as<uint8_t>(), chunks_offset, S.seg.as<uint64_t>(), d_total,
                                                    chunks_only ? 0u : 1u);
    PCOB_CUDA_TRY(cudaGetLastError());
    PCOB_CUDA_TRY(cudaMemcpyAsync(index_dst, idx.p, off, (flags & PCO_B200_INDEX_ON_DEVICE) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, stream));
    res->index_bytes = off;
  }
  if (dst_dev) PCOB_CUDA_TRY(cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, stream));  // read back behind the kernels
  PCOB_CUDA_TRY(cudaStreamSynchronize(stream));
  PCOB_CUDA_TRY(cudaGetLastError());
  if (total > dst_cap) return fail(PCO_B200_IO, "failed to write whole buffer (need " + std::to_string(total) + " bytes, dst_cap " + std::to_string(dst_cap) + ")");
  res->total_bytes = total;
  return PCO_B200_OK;
}

}  // namespace pcob200
