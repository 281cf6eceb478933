"""Seeded synthetic mutation matrices of the benchmark shapes:
K_true Bernoulli(geno_p) genotypes, uniform cluster membership, false negatives / false
positives, then missing entries.  Returns (float64 [N,M] with NaN, z_true)."""
import numpy as np

# BASELINE.json configs: name -> generator and model settings
CONFIGS = {
    'C2': dict(cells=10_000, muts=500, k_true=20, fn=0.2, fp=0.01, miss=0.10, learning=True, pp=(0.25, 0.25)),
    'C3': dict(cells=100_000, muts=1_000, k_true=20, fn=0.2, fp=0.01, miss=0.10, learning=True, pp=(0.25, 0.25)),
    'C4': dict(cells=50_000, muts=5_000, k_true=20, fn=0.3, fp=1e-4, miss=0.10, learning=False, pp=(0.25, 0.25),
               FN=0.3, FP=1e-4, sm_prob=0.75),
    'C5': dict(cells=1_000_000, muts=50, k_true=10, fn=0.2, fp=0.01, miss=0.30, learning=True, pp=(1, 1)),
}


def make_matrix(cells, muts, k_true=20, fn=0.2, fp=0.01, miss=0.10, seed=0, geno_p=0.3, chunk=8192):
    rng = np.random.default_rng(seed)
    geno = rng.random((k_true, muts)) < geno_p
    z = rng.integers(0, k_true, size=cells)
    out = np.empty((cells, muts), dtype=np.float64)
    for r0 in range(0, cells, chunk):                  # bounded temporaries at 1M cells
        zz = z[r0:r0 + chunk]
        truth = geno[zz]
        flip = rng.random(truth.shape)
        blk = np.where(truth, flip >= fn, flip < fp).astype(np.float64)
        blk[rng.random(truth.shape) < miss] = np.nan
        out[r0:r0 + chunk] = blk
    return out, z
