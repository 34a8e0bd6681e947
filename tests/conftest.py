import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

UFFIZI = ROOT / "data" / "uffizi_512.hdr"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure)."""
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def R():
    """The product package (ctypes over librmdf_b200.so)."""
    import rmdf_b200
    return rmdf_b200


@pytest.fixture(scope="session")
def uffizi(O):
    return O.hdr_decode(UFFIZI.read_bytes())


@pytest.fixture(scope="session")
def oracle_envs(O, uffizi):
    """The reference's five cube maps built entirely by the oracle:
    reflection (face 170) + cos^{1,8,64,512} of the 256×128 downsample through RGBE (face 85)."""
    cache = ROOT / "tests" / "golden" / "_oracle_envs.npz"
    if cache.exists():
        z = np.load(cache)
        return [z["e%d" % i] for i in range(5)]
    small = O.resize(uffizi, 256)
    envs = [O.latlong_to_cube(uffizi)]
    for p in (1.0, 8.0, 64.0, 512.0):
        envs.append(O.latlong_to_cube(O.rgbe_roundtrip(O.convolve(small, p))))
    try:
        np.savez_compressed(cache, **{"e%d" % i: e for i, e in enumerate(envs)})
    except OSError:
        pass
    return envs


@pytest.fixture(scope="session")
def sr(R):
    """One ShaderRenderer on cuda:0 for the whole session (GPU tests only)."""
    r = R.ShaderRenderer(0)
    yield r
    r.close()


def psnr(a, b, peak=1.0):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    mse = np.mean((a - b) ** 2)
    return float("inf") if mse == 0 else 10.0 * np.log10(peak * peak / mse)
