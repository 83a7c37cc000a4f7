"""Seeded synthetic SIFT-128 descriptor collections (SURVEY.md §8d "Synthetic inputs").

Rows imitate real SIFT after the reference's VLSIFT scaling (feature_extractor_vl_sift.cpp:199-203:
512 x unit-norm): 128 gamma(0.6) draws, L2-normalised, clamped at 0.2, renormalised, then
q = min(255, floor(512 x)) so that ||q||^2 ~ 2.6e5.  To give non-trivial match lists, every image copies a
random 30-50 % of its rows from a shared "scene pool" and perturbs them with N(0, sigma=6) noise before
re-quantising.  The float variant (unit-norm rows, the CUDASIFT regime of
feature_extractor_cuda_sift.cpp:75-80) is the pre-quantisation vector.

Pure numpy; deterministic in (collection_seed, image_id).
"""
from __future__ import annotations

import numpy as np

DIM = 128
BASE_SEED = 0x5EED0000


def _sift_like_unit(rng: np.random.Generator, rows: int) -> np.ndarray:
    x = rng.standard_gamma(0.6, size=(rows, DIM)).astype(np.float32)
    x /= np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-12)
    np.minimum(x, 0.2, out=x)
    x /= np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-12)
    return x


def quantize_512(unit: np.ndarray) -> np.ndarray:
    """u8 rows from unit-norm floats: min(255, floor(512 x))."""
    return np.minimum(255.0, np.floor(unit * 512.0)).astype(np.uint8)


class Collection:
    """A synthetic image collection sharing one scene pool."""

    def __init__(self, rows_per_image: int, pool_factor: float = 2.0, seed: int = 0, noise_sigma: float = 6.0):
        self.rows = int(rows_per_image)
        self.seed = int(seed)
        self.sigma = float(noise_sigma)
        pool_rows = max(1, int(self.rows * pool_factor))
        prng = np.random.default_rng([BASE_SEED, self.seed, 0xFFFF])
        self.pool_unit = _sift_like_unit(prng, pool_rows)

    def image_unit(self, image_id: int, rows: int | None = None) -> np.ndarray:
        """Unit-norm float rows of one image (the float regime)."""
        rows = self.rows if rows is None else int(rows)
        rng = np.random.default_rng([BASE_SEED + int(image_id), self.seed])
        x = _sift_like_unit(rng, rows)
        frac = rng.uniform(0.3, 0.5)
        k = min(int(rows * frac), self.pool_unit.shape[0])
        if k > 0:
            dst = rng.choice(rows, size=k, replace=False)
            src = rng.choice(self.pool_unit.shape[0], size=k, replace=False)
            noisy = self.pool_unit[src] * 512.0 + rng.normal(0.0, self.sigma, size=(k, DIM)).astype(np.float32)
            np.maximum(noisy, 0.0, out=noisy)
            noisy /= np.maximum(np.linalg.norm(noisy, axis=1, keepdims=True), 1e-12)
            x[dst] = noisy
        return x

    def image_u8(self, image_id: int, rows: int | None = None) -> np.ndarray:
        """Integer regime: the rows the bit-exact contract is stated on."""
        return quantize_512(self.image_unit(image_id, rows))

    def image_f32_512(self, image_id: int, rows: int | None = None) -> np.ndarray:
        """Integer-valued float rows, i.e. what a cv::Mat CV_32FC1 holds after an explicit rounding step."""
        return self.image_u8(image_id, rows).astype(np.float32)


def exhaustive_pairs(n_images: int) -> np.ndarray:
    """All unordered pairs (i<j), grouped by i like the reference's outer idx1 loop (fine_matching_graph.cc:58-64)."""
    i, j = np.triu_indices(n_images, k=1)
    return np.stack([i, j], axis=1).astype(np.int32)


def gps_neighbour_pairs(n_images: int, k: int = 30, seed: int = 0) -> np.ndarray:
    """Guided pair list for an aerial block: each image -> its k nearest on a jittered flight grid (L1 distance,
    as initial_matching_graph.cc:142-161 uses), deduplicated to unordered pairs, grouped by the lower index."""
    rng = np.random.default_rng([BASE_SEED, seed, 0x6B5])
    side = int(np.ceil(np.sqrt(n_images)))
    gx, gy = np.meshgrid(np.arange(side), np.arange(side))
    xy = np.stack([gx.ravel(), gy.ravel()], axis=1)[:n_images].astype(np.float64)
    xy += rng.normal(0.0, 0.15, size=xy.shape)
    pairs = set()
    for i in range(n_images):
        d = np.abs(xy - xy[i]).sum(axis=1)
        d[i] = np.inf
        nn = np.argpartition(d, min(k, n_images - 1) - 1)[: min(k, n_images - 1)]
        for j in nn:
            a, b = (i, int(j)) if i < j else (int(j), i)
            pairs.add((a, b))
    out = np.array(sorted(pairs), dtype=np.int32).reshape(-1, 2)
    return out


def retrieval_pairs(n_images: int, partners: int = 40, seed: int = 0) -> np.ndarray:
    """Random retrieval-candidate pair list (web collection): `partners` random partners per image, deduplicated."""
    rng = np.random.default_rng([BASE_SEED, seed, 0x7E7])
    a = np.repeat(np.arange(n_images, dtype=np.int64), partners)
    b = rng.integers(0, n_images, size=a.shape[0])
    keep = a != b
    lo = np.minimum(a[keep], b[keep])
    hi = np.maximum(a[keep], b[keep])
    key = np.unique(lo * n_images + hi)
    return np.stack([key // n_images, key % n_images], axis=1).astype(np.int32)
