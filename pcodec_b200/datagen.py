"""Seeded synthetic inputs for the BASELINE configs, shaped after the reference's
pco_cli/generate_randoms.py distributions.  numpy versions feed tests; torch versions fill HBM for bench.py."""
import numpy as np

CHUNK_N = 1 << 18


def c1_u32_lomax(n=CHUNK_N, seed=0):
    """C1: u32, floor(lomax(a=0.5, median 1000)) clipped (generate_randoms.py:162-169)."""
    rng = np.random.default_rng(seed)
    a = 0.5
    scale = 1000.0 / (2 ** (1 / a) - 1)
    x = np.floor(rng.pareto(a, size=n) * scale)
    return np.clip(x, 0, 2**32 - 1).astype(np.uint32)


def c2_u64_cumsum_geometric(n=CHUNK_N, seed=0, p=0.001):
    """C2(i): u64 cumulative sum of geometric(p) increments (generate_randoms.py:157-159): first-order deltas form a
    smooth multi-bin distribution, so the tANS path is exercised."""
    rng = np.random.default_rng(seed)
    return np.cumsum(rng.geometric(p, size=n)).astype(np.uint64)


def c3_f64_decimal_sinusoid(n=CHUNK_N, seed=0):
    """C3: f64 round(1e5 cos(2 pi i / P)) * 0.01, P = n/103 (generate_randoms.py:234-239,283-285): FloatMult(0.01) primary is
    smooth under second-order deltas; the secondary is ULP noise."""
    i = np.arange(n, dtype=np.float64) + seed * 17
    return np.round(1e5 * np.cos(2 * np.pi * i / (n / 103.0))) * 0.01


def c5_sweep(dtype, n=CHUNK_N, seed=0):
    """C5: C2(i)-like walk cast to the dtype; normal walk for floats (generate_randoms.py:242-244)."""
    rng = np.random.default_rng(seed)
    dt = np.dtype(dtype)
    if dt.kind == "f":
        return np.cumsum(rng.normal(size=n)).astype(dt)
    steps = rng.geometric(0.05, size=n).astype(np.int64) - 10
    return np.cumsum(steps).astype(np.uint64).astype(dt.str.replace("i", "u")).view(dt)


def c2_u64_torch(n_chunks, chunk_n=CHUNK_N, seed=0, device="cuda", p=0.001):
    """C2(i) generated in HBM: one independent walk per chunk.  Returns an int64 tensor holding the u64 bit patterns."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty(n_chunks * chunk_n, dtype=torch.int64, device=device)
    rows = max(1, min(n_chunks, (1 << 25) // chunk_n))
    log1mp = float(np.log1p(-p))
    for s in range(0, n_chunks, rows):
        r = min(rows, n_chunks - s)
        u = torch.rand(r, chunk_n, dtype=torch.float64, device=device, generator=g).clamp_(min=1e-300)
        inc = torch.floor(torch.log(u) / log1mp).to(torch.int64) + 1
        out[s * chunk_n:(s + r) * chunk_n] = torch.cumsum(inc, dim=1).reshape(-1)
    return out
