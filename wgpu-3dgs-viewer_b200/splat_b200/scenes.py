"""Synthetic scenes and cameras of SURVEY.md §8(d) — shared by tests/ and bench.py.

Deterministic (numpy default_rng, seed 0x3D65 + config index).  Positions ~ U([-10,10]^3),
log-scales ~ U(-5,-3), rotations = normalised N(0,1)^4, opacity logit ~ N(0,2),
DC colour u8 ~ U{0..255}^3, SH rest ~ N(0, 0.15^2).
"""
from __future__ import annotations

import numpy as np

from .api import GAUSSIAN_DTYPE

BASE_SEED = 0x3D65


def synthetic_gaussians(n: int, seed: int = BASE_SEED, extent: float = 10.0, log_scale=(-5.0, -3.0)) -> np.ndarray:
    rng = np.random.default_rng(seed)
    g = np.zeros(n, dtype=GAUSSIAN_DTYPE)
    g["pos"] = rng.uniform(-extent, extent, size=(n, 3)).astype(np.float32)
    g["scale"] = np.exp(rng.uniform(log_scale[0], log_scale[1], size=(n, 3))).astype(np.float32)
    q = rng.standard_normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    g["rot"] = q.astype(np.float32)
    logit = rng.normal(0.0, 2.0, size=n)
    alpha = np.clip(1.0 / (1.0 + np.exp(-logit)) * 255.0, 0, 255).astype(np.uint8)
    g["color"][:, :3] = rng.integers(0, 256, size=(n, 3), dtype=np.uint8)
    g["color"][:, 3] = alpha
    g["sh"] = rng.normal(0.0, 0.15, size=(n, 45)).astype(np.float32)
    return g


# camera presets: (pos, yaw, pitch)
CAMERA_OUTSIDE = ((0.0, 0.0, -30.0), 0.1, 0.1)
CAMERA_INSIDE = ((0.0, 0.0, 0.0), 0.1, 0.1)


def orbit_camera(k: int, count: int = 64, radius: float = 30.0):
    """Config 5a: camera k of `count` on a circle of `radius` looking at the scene centre."""
    a = 2.0 * np.pi * k / count
    pos = (float(-radius * np.sin(a)), 0.0, float(-radius * np.cos(a)))
    return pos, float(a + 0.02), 0.02


def single_red_gaussian() -> np.ndarray:
    """The reference's e2e fixture (tests/e2e/viewer.rs:42-48)."""
    g = np.zeros(1, dtype=GAUSSIAN_DTYPE)
    g["pos"][0] = (0, 0, 1)
    g["rot"][0] = (0, 0, 0, 1)
    g["scale"][0] = (1, 1, 1)
    g["color"][0] = (255, 0, 0, 255)
    return g
