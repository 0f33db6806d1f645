"""Deterministic synthetic inputs shaped like BASELINE.json's configs.

Host (numpy) generators serve the oracle-sized parity tests and the golden-vector script;
device (torch) generators build the full-size benchmark inputs directly in HBM.
Nothing here is on the product path.
"""
from __future__ import annotations

import numpy as np
from scipy.sparse import csr_matrix


def zipf_priors(m: int, avg_labels: float = 5.0) -> np.ndarray:
    """Label priors pi_j ~ 1/j scaled to `avg_labels` expected positives per row."""
    pri = 1.0 / np.arange(1, m + 1, dtype=np.float64)
    pri *= avg_labels / pri.sum()
    return np.clip(pri, 1e-5, 0.5)


def make_tie_free(eta: np.ndarray) -> np.ndarray:
    """Nudge duplicate values inside a row with nextafter until every row is tie-free."""
    eta = np.ascontiguousarray(eta)
    for i in range(eta.shape[0]):
        row = eta[i]
        for _ in range(64):
            order = np.argsort(row, kind="stable")
            srt = row[order]
            dup = np.nonzero(srt[1:] == srt[:-1])[0]
            if dup.size == 0:
                break
            # bump the later element of each equal pair one ulp up (repeat until clean)
            row[order[dup + 1]] = np.nextafter(srt[dup + 1], np.float32(2.0)).astype(row.dtype)
        else:  # pragma: no cover
            raise RuntimeError("could not make row tie-free")
    return eta


def dense_probs(n: int, m: int, seed: int, dtype=np.float32, tie_free: bool = True,
                spread: float = 2.0) -> np.ndarray:
    """eta_ij = sigmoid(logit(pi_j) + spread * N(0,1)), clipped to [1e-6, 1-1e-6]."""
    rng = np.random.default_rng(seed)
    pri = zipf_priors(m)
    logit = np.log(pri) - np.log1p(-pri)
    z = logit[None, :] + spread * rng.standard_normal((n, m))
    eta = 1.0 / (1.0 + np.exp(-z))
    eta = np.clip(eta, 1e-6, 1.0 - 1e-6).astype(dtype)
    if tie_free:
        eta = make_tie_free(eta)
    return eta


def csr_probs(n: int, m: int, nnz_per_row: int, seed: int, dtype=np.float32,
              ragged: bool = False) -> csr_matrix:
    """Rows of `nnz_per_row` distinct Zipf-weighted label ids (sorted, int32) with values
    0.001 + 0.98 * U^3 (strictly inside (0, 1), distinct within a row)."""
    rng = np.random.default_rng(seed)
    w = 1.0 / np.arange(1, m + 1, dtype=np.float64)
    w /= w.sum()
    cdf = np.cumsum(w)
    indptr = np.zeros(n + 1, dtype=np.int32)
    idx_rows, val_rows = [], []
    for i in range(n):
        nz = nnz_per_row
        if ragged:
            nz = int(rng.integers(max(1, nnz_per_row // 4), nnz_per_row + 1))
        nz = min(nz, m)
        got = np.empty(0, dtype=np.int64)
        while got.size < nz:
            cand = np.searchsorted(cdf, rng.random(2 * nz))
            got = np.unique(np.concatenate([got, np.minimum(cand, m - 1)]))
        if got.size > nz:
            got = np.sort(rng.choice(got, nz, replace=False))
        vals = (0.001 + 0.98 * rng.random(nz) ** 3).astype(dtype)
        # distinct values inside the row
        vals = make_tie_free(vals[None, :])[0]
        idx_rows.append(got.astype(np.int32))
        val_rows.append(vals)
        indptr[i + 1] = indptr[i] + nz
    return csr_matrix((np.concatenate(val_rows), np.concatenate(idx_rows), indptr), shape=(n, m))


# ------------------------------------------------------------------------------------------
# Device-side generators (bench.py).  Same distributions, torch RNG streams.
# ------------------------------------------------------------------------------------------

def dense_probs_device(n: int, m: int, seed: int, device, ld: int | None = None,
                       chunk_rows: int = 8192, spread: float = 2.0):
    """Row-major float32 [n, ld] (ld >= m) probability matrix generated in HBM."""
    import torch

    ld = m if ld is None else ld
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    pri = torch.from_numpy(zipf_priors(m)).to(device)
    logit = (torch.log(pri) - torch.log1p(-pri)).to(torch.float32)
    out = torch.zeros((n, ld), dtype=torch.float32, device=device)
    for s in range(0, n, chunk_rows):
        e = min(n, s + chunk_rows)
        z = torch.randn((e - s, m), generator=gen, device=device, dtype=torch.float32)
        z.mul_(spread).add_(logit[None, :])
        torch.sigmoid(z, out=z)
        z.clamp_(1e-6, 1.0 - 1e-6)
        out[s:e, :m] = z
    return out


def csr_probs_device(n: int, m: int, nnz_per_row: int, seed: int, device):
    """CSR (data f32, indices int32 sorted, indptr int64) with `nnz_per_row` distinct
    Zipf-ish labels per row, generated in HBM."""
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    # distinct ids: log-uniform start then strictly increasing by adding positive gaps
    u = torch.rand((n, nnz_per_row), generator=gen, device=device, dtype=torch.float64)
    gaps = torch.floor(torch.exp(u * np.log(2.0 * m / nnz_per_row))).to(torch.int64).clamp_(min=1)
    ids = torch.cumsum(gaps, dim=1) - 1
    over = ids[:, -1:] - (m - 1)
    ids = torch.where(over > 0, (ids.to(torch.float64) * ((m - 1) / ids[:, -1:].to(torch.float64))).to(torch.int64), ids)
    # enforce strict monotonicity after rescale
    ar = torch.arange(nnz_per_row, device=device)[None, :]
    ids = torch.maximum(ids, ar)
    ids = torch.cummax(ids - ar, dim=1).values + ar
    ids = ids.clamp_(max=m - 1 - (nnz_per_row - 1 - ar)).to(torch.int32)
    vals = (0.001 + 0.98 * torch.rand((n, nnz_per_row), generator=gen, device=device) ** 3).to(torch.float32)
    indptr = torch.arange(n + 1, device=device, dtype=torch.int64) * nnz_per_row
    return vals.reshape(-1).contiguous(), ids.reshape(-1).contiguous(), indptr
