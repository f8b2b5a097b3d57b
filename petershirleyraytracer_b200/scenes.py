"""Scene / camera definitions of the BASELINE.json configs.

The reference ships exactly one scene (programs/main.cc:62-63); the book-layout scenes below are synthetic
extensions built through the reference's public API surface (sphere = centre + radius, camera = four
public vec3 fields), generated from a fixed splitmix64 stream so every consumer (CUDA path, oracle,
reference build) sees identical doubles.
"""
from __future__ import annotations

import numpy as np

from . import Camera

_MASK = 0xFFFFFFFFFFFFFFFF


class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & _MASK

    def next_u64(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & _MASK
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK
        return z ^ (z >> 31)

    def u01(self) -> float:
        return (self.next_u64() >> 11) * (1.0 / 9007199254740992.0)


def default_scene():
    """programs/main.cc:62-63."""
    centres = np.array([[0.0, 0.0, -1.0], [0.0, -100.5, 0.0]])
    radii = np.array([0.5, 100.0])
    return centres, radii


def default_image_size(width: int = 400):
    """programs/main.cc:57-58: height = (int)(width / aspect_ratio)."""
    return width, int(width / (16.0 / 9.0))


def book_scene(grid_half: int = 11, seed: int = 42):
    """Book-layout random spheres: ground r=1000, a jittered grid of r=0.2 spheres,
    three r=1 spheres.  grid_half=11 -> ~485 spheres (configs 3/5); 158 -> ~99.9k (config 4)."""
    rng = SplitMix64(seed)
    c = [(0.0, -1000.0, 0.0)]
    r = [1000.0]
    for a in range(-grid_half, grid_half):
        for b in range(-grid_half, grid_half):
            rng.u01()  # mirrors the book's material draw
            x1, x2 = rng.u01(), rng.u01()
            cx, cy, cz = a + 0.9 * x1, 0.2, b + 0.9 * x2
            if ((cx - 4.0) ** 2 + (cy - 0.2) ** 2 + cz ** 2) ** 0.5 > 0.9:
                c.append((cx, cy, cz))
                r.append(0.2)
    for cx in (0.0, -4.0, 4.0):
        c.append((cx, 1.0, 0.0))
        r.append(1.0)
    return np.array(c, dtype=np.float64), np.array(r, dtype=np.float64)


def book_camera(width: int, height: int) -> Camera:
    """lookfrom (13,2,3) -> (0,0,0), vup (0,1,0), vfov 20 deg, no lens."""
    return Camera.look_at((13.0, 2.0, 3.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 20.0, width / height)


CONFIGS = {
    # name: (scene factory, camera factory(W,H), W, H, spp, max_depth)
    "c1_default": (default_scene, lambda w, h: Camera.default(), 400, 225, 100, 50),
    "c2_primary": (default_scene, lambda w, h: Camera.default(), 400, 225, 1, 1),
    "c3_book_1200x800": (lambda: book_scene(11), book_camera, 1200, 800, 500, 50),
    "c4_bvh_1920x1080": (lambda: book_scene(158), book_camera, 1920, 1080, 256, 50),
    "c5_book_4k": (lambda: book_scene(11), book_camera, 3840, 2160, 1024, 50),
}
