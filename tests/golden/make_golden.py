"""Mints the known-answer vectors under tests/golden/ from the CPU oracle (SURVEY.md §8c:
the reference ships no tests or golden images, so the oracle's own KATs are frozen here to
catch regressions of the oracle itself).  Run from the repo root:

    python tests/golden/make_golden.py

The reference (Haskell + GLSL) cannot be executed in this image, so these vectors are
"parity unpinned": they pin the oracle to its committed state, not to a run of the reference.
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402

OUT = Path(__file__).resolve().parent


def synthetic_envs():
    """Five tiny deterministic cube maps (faces 6, 4, 4, 3, 3) so that the texture path is pinned
    without committing the 2 MB uffizi cubes."""
    rng = np.random.RandomState(1234)
    envs = []
    for w in (6, 4, 4, 3, 3):
        e = rng.uniform(0.0, 2.0, size=(6, w, w, 3)).astype(np.float32)
        envs.append(O.glsl("round_to_half", e).reshape(e.shape))
    return envs


def main():
    # (i) distance-estimator values on a lattice, per scene
    g = np.linspace(-1.1, 1.1, 5, dtype=np.float32)
    pts = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    de = {str(s): O.distance(s, 7.0, pts).view(np.uint32).tolist() for s in range(4)}
    (OUT / "de_lattice.json").write_text(json.dumps({"points": pts.view(np.uint32).tolist(), "time": 7.0, "de_bits": de}))
    # (iv) 64×64 renders per scene at in_time 0 and 7: white env and synthetic env
    envs = synthetic_envs()
    frames = {}
    for s in range(4):
        for t in (0.0, 7.0):
            frames["white_s%d_t%d" % (s, int(t))] = O.render(s, 64, 64, t)["fb"]
            frames["env_s%d_t%d" % (s, int(t))] = O.render(s, 64, 64, t, envs=envs)["fb"]
    np.savez_compressed(OUT / "frames_64.npz", **frames)
    # julia
    jl = {"j_t%d_s%d" % (int(t), sm): O.julia(96, 64, bool(sm), t) for t in (0.0, 7.0, 123.5) for sm in (0, 1)}
    np.savez_compressed(OUT / "julia_96x64.npz", **jl)
    # camera + cornell table
    cams = {"%d_%g" % (s, t): O.camera(s, t).view(np.uint32).tolist() for s in (0, 2) for t in (0.0, 7.0)}
    (OUT / "camera.json").write_text(json.dumps(cams))
    np.save(OUT / "cornell_vertices.npy", O.cornell_vertices())
    # (iii) convolution of the six-colour test env map, downsampled to 32×16 (fast), powers 1 and 8
    small = O.resize(O.test_latlong_envmap(), 32)
    np.savez_compressed(OUT / "test_envmap_conv.npz", small=small, p1=O.convolve(small, 1.0), p8=O.convolve(small, 8.0))
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
