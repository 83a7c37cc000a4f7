"""GPU-side generator of the synthetic SIFT-128 collections (same recipe as synth.py, torch ops on the device), for the
BASELINE workloads whose tables are too large to draw with numpy in reasonable time (configs #3-#5: 2.6 - 42 GB).

Deterministic in (collection seed, image id) on a given GPU model and torch build: any rank can regenerate any image, which
is how bench.py checks rows that arrived over NCCL and re-creates, on the host, the exact bytes the GPU matched.
Bench plumbing only (synthetic inputs); nothing of the product path lives here.
"""
from __future__ import annotations

import torch

DIM = 128
BASE_SEED = 0x5EED0000


def _sift_like_unit(gen: torch.Generator, rows: int, device) -> torch.Tensor:
    x = torch._standard_gamma(torch.full((rows, DIM), 0.6, device=device, dtype=torch.float32), generator=gen)
    x /= x.norm(dim=1, keepdim=True).clamp_min(1e-12)
    x.clamp_(max=0.2)
    x /= x.norm(dim=1, keepdim=True).clamp_min(1e-12)
    return x


class GpuCollection:
    def __init__(self, rows_per_image: int, device, seed: int = 0, pool_factor: float = 2.0, noise_sigma: float = 6.0):
        self.rows, self.seed, self.sigma, self.device = int(rows_per_image), int(seed), float(noise_sigma), device
        g = torch.Generator(device=device)
        g.manual_seed(BASE_SEED * 7919 + self.seed * 104729 + 0xFFFF)
        self.pool_unit = _sift_like_unit(g, max(1, int(self.rows * pool_factor)), device)

    def image_u8(self, image_id: int, rows: int | None = None) -> torch.Tensor:
        rows = self.rows if rows is None else int(rows)
        g = torch.Generator(device=self.device)
        g.manual_seed((BASE_SEED + int(image_id)) * 7919 + self.seed * 104729)
        x = _sift_like_unit(g, rows, self.device)
        frac = 0.3 + 0.2 * torch.rand((), generator=g, device=self.device).item()
        k = min(int(rows * frac), self.pool_unit.shape[0])
        if k > 0:
            dst = torch.randperm(rows, generator=g, device=self.device)[:k]
            src = torch.randperm(self.pool_unit.shape[0], generator=g, device=self.device)[:k]
            noisy = self.pool_unit[src] * 512.0 + torch.randn((k, DIM), generator=g, device=self.device) * self.sigma
            noisy.clamp_(min=0.0)
            noisy /= noisy.norm(dim=1, keepdim=True).clamp_min(1e-12)
            x[dst] = noisy
        return torch.clamp(torch.floor(x * 512.0), max=255.0).to(torch.uint8)
