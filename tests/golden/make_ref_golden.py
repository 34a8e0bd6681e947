"""Mints tests/golden/ref_shader.npz from oracle/_ref — the reference's OWN fragment shader
(/root/reference/fragment.shd) compiled by g++ (make -C oracle ref) — i.e. outputs of the
reference itself, run in the dev container.  Run from the repo root:

    python tests/golden/make_ref_golden.py

The GPU box has no /root/reference; these vectors are how the reference's results travel:
tests/test_oracle_cpu.py::test_ref_golden_* checks the C oracle against them and
tests/test_march_parity.py::test_cuda_equals_reference_shader_golden the CUDA kernels.
Platform conventions (built-in functions, GL sampler) are those of DESIGN.md §2 C0-C6.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle as O, oracle_ref as REF  # noqa: E402

OUT = Path(__file__).resolve().parent
f32 = np.float32


def synthetic_envs():
    rng = np.random.RandomState(1234)
    envs = []
    for w in (6, 4, 4, 3, 3):
        e = rng.uniform(0.0, 2.0, size=(6, w, w, 3)).astype(f32)
        envs.append(O.glsl("round_to_half", e).reshape(e.shape))
    return envs


def main():
    assert REF.available(), "oracle/_ref is not built (needs /root/reference)"
    envs = synthetic_envs()
    z = {}
    for s in range(4):
        for t in (0.0, 7.0):
            for name, e in (("white", None), ("env", envs)):
                r = REF.render(s, 64, 48, t, envs=e)
                k = "%s_s%d_t%d" % (name, s, int(t))
                z["fb_" + k] = r["fb"]
                z["lin_" + k] = r["linear"].view(np.uint32)
                z["hit_" + k] = np.packbits(r["hit"])
    # shader switches that are commented out in the reference
    for s, fl in ((0, O.FLAG_RAY_TRACING_TEST), (1, O.FLAG_CAM_ORTHO), (2, O.FLAG_CAM_ORTHO)):
        z["fb_flag%d_s%d" % (fl, s)] = REF.render(s, 64, 48, 1.0, envs=envs, flags=fl)["fb"]
    # crops of the BASELINE frames (configs 1, 3, 4) at their real sizes, white env
    z["fb_c1_crop"] = REF.render(2, 512, 512, 0.0, rect=(224, 240, 288, 272))["fb"]
    z["fb_c3_cornell_crop"] = REF.render(0, 3840, 2160, 0.0, rect=(1888, 1064, 1952, 1096))["fb"]
    z["fb_c3_csg_crop"] = REF.render(1, 3840, 2160, 0.0, rect=(1888, 1064, 1952, 1096))["fb"]
    z["fb_c4_crop"] = REF.render(2, 15360, 8640, 0.0, rect=(7648, 4304, 7712, 4336))["fb"]
    # distance estimator on a lattice, per variant
    g = np.linspace(-1.1, 1.1, 7, dtype=f32)
    pts = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3) + f32(0.013)
    z["de_points"] = pts.view(np.uint32)
    for s in range(4):
        z["de_s%d" % s] = REF.distance(s, 7.0, pts).view(np.uint32)
    np.savez_compressed(OUT / "ref_shader.npz", **z)
    print("wrote", OUT / "ref_shader.npz", (OUT / "ref_shader.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
