"""Synthetic particle clouds of BASELINE.json's configs: lattice spacing d = 2r, uniform jitter
in +-0.25 d per axis, float32 AoS (N, 3), numpy.random.default_rng(seed)."""
from __future__ import annotations

import numpy as np


def _lattice(nx: int, ny: int, nz: int, d: float, origin=(0.0, 0.0, 0.0)) -> np.ndarray:
    x = (np.arange(nx, dtype=np.float32) * np.float32(d) + np.float32(origin[0]))
    y = (np.arange(ny, dtype=np.float32) * np.float32(d) + np.float32(origin[1]))
    z = (np.arange(nz, dtype=np.float32) * np.float32(d) + np.float32(origin[2]))
    out = np.empty((nx, ny, nz, 3), dtype=np.float32)
    out[..., 0] = x[:, None, None]
    out[..., 1] = y[None, :, None]
    out[..., 2] = z[None, None, :]
    return out.reshape(-1, 3)


def _jitter(p: np.ndarray, d: float, rng) -> np.ndarray:
    # chunked to bound peak memory on 50 M-particle clouds
    step = 4_000_000
    for a in range(0, len(p), step):
        b = min(len(p), a + step)
        p[a:b] += rng.uniform(-0.25 * d, 0.25 * d, size=(b - a, 3)).astype(np.float32)
    return p


def jittered_cube(n_per_dim: int = 100, r: float = 0.025, seed: int = 1234) -> np.ndarray:
    """cfg-2: n^3 jittered lattice cube (1 M particles for n = 100)."""
    rng = np.random.default_rng(seed)
    d = 2.0 * r
    return _jitter(_lattice(n_per_dim, n_per_dim, n_per_dim, d), d, rng)


def dam_break(column=(200, 230, 200), sheet=(400, 10, 200), r: float = 0.01, seed: int = 2) -> np.ndarray:
    """cfg-3/4: fluid column at the origin plus a thin floor sheet adjacent in +x."""
    rng = np.random.default_rng(seed)
    d = 2.0 * r
    col = _lattice(*column, d)
    sh = _lattice(*sheet, d, origin=(column[0] * d, 0.0, 0.0))
    return _jitter(np.concatenate([col, sh], axis=0), d, rng)


def dam_break_10m() -> np.ndarray:
    return dam_break((200, 230, 200), (400, 10, 200), 0.01, 2)


def dam_break_50m() -> np.ndarray:
    return dam_break((340, 370, 340), (1063, 20, 340), 0.01, 3)


def dam_break_scaled(n_target: int, r: float = 0.01, seed: int = 3) -> np.ndarray:
    """A dam break with the proportions of cfg-4 scaled to roughly n_target particles."""
    f = (n_target / 50_000_400.0) ** (1.0 / 3.0)
    col = tuple(max(4, int(round(v * f))) for v in (340, 370, 340))
    sh = (max(4, int(round(1063 * f))), max(2, int(round(20 * f))), col[2])
    return dam_break(col, sh, r, seed)


def splash(n_body=(60, 65, 60), n_droplets: int = 12, r: float = 0.005, seed: int = 4) -> np.ndarray:
    """cfg-5 in miniature: a body of fluid plus lattice-ball droplets above it (sparse subdomains)."""
    rng = np.random.default_rng(seed)
    d = 2.0 * r
    parts = [_lattice(*n_body, d)]
    top = n_body[1] * d
    for _ in range(n_droplets):
        rad = rng.uniform(2.0, 6.0) * d
        ctr = np.array([rng.uniform(0, n_body[0] * d), top + rng.uniform(8, 30) * d, rng.uniform(0, n_body[2] * d)])
        m = int(np.ceil(rad / d))
        ball = _lattice(2 * m + 1, 2 * m + 1, 2 * m + 1, d, origin=tuple(ctr - m * d))
        ball = ball[np.linalg.norm(ball - ctr[None].astype(np.float32), axis=1) <= rad]
        parts.append(ball)
    return _jitter(np.concatenate(parts, axis=0), d, rng)


def splash_200m(scale: float = 1.0, seed: int = 4, overlap: bool = False) -> np.ndarray:
    """cfg-5: dam-break body 540x583x540 (r = 0.005) plus lattice-ball droplets of radius U[10, 40] d above it, added until
    N = 200 M (~340 droplets).  Droplet centres are rejection-sampled so that no two droplets touch (a gap of 2 d between their
    surfaces; the air space above the body is made as tall as that needs): particles of an incompressible fluid do not
    interpenetrate, and superimposed lattices would be a 2-3x rest-density cloud that no SPH frame contains.  `overlap=True`
    gives the round-2 first-run cloud (random centres in a shallow layer, heavily superimposed droplets) as a stress case for
    the dense-cluster escape routes.  `scale` < 1 shrinks every lattice count (tests / smaller boxes)."""
    r = 0.005
    d = 2.0 * r
    rng = np.random.default_rng(seed)
    nb = tuple(max(4, int(round(v * scale))) for v in (540, 583, 540))
    parts = [_lattice(*nb, d)]
    n = len(parts[0])
    target = int(200_000_000 * scale ** 3)
    top = nb[1] * d
    placed = []                                     # (centre, radius) of the droplets so far
    height = 500.0                                  # air space for droplet centres, in units of scale * d
    while n < target:
        rad = rng.uniform(10.0, 40.0) * scale * d
        for attempt in range(4000):
            ctr = np.array([rng.uniform(0, nb[0] * d), top + rad + rng.uniform(2.0, 60.0 if overlap else height) * scale * d, rng.uniform(0, nb[2] * d)])
            if overlap or all(np.linalg.norm(ctr - c0) >= rad + r0 + 2.0 * d for c0, r0 in placed):
                break
        else:
            height *= 1.5                           # crowded: raise the ceiling and draw again
            continue
        placed.append((ctr, rad))
        m = int(np.ceil(rad / d))
        ball = _lattice(2 * m + 1, 2 * m + 1, 2 * m + 1, d, origin=tuple(ctr - m * d))
        ball = ball[np.linalg.norm(ball - ctr[None].astype(np.float32), axis=1) <= rad]
        if n + len(ball) > target:
            ball = ball[: max(target - n, 0)]
        parts.append(ball)
        n += len(ball)
    return _jitter(np.concatenate(parts, axis=0), d, rng)


def load_vtk_points(path: str) -> np.ndarray:
    """Legacy-VTK BINARY (big-endian float) POINTS reader, enough for data/*_particles.vtk fixtures."""
    b = open(path, "rb").read()
    k = b.index(b"POINTS")
    e = b.index(b"\n", k)
    n = int(b[k:e].split()[1])
    return np.frombuffer(b[e + 1:e + 1 + 12 * n], dtype=">f4").reshape(n, 3).astype("<f4")
