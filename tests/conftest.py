import glob
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(HERE, "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def load_construction(n, K, crc):
    z = np.load(os.path.join(GOLDEN, "construction_n%d_K%d_crc%d.npz" % (n, K, crc)))
    N = 1 << n
    frozen = np.unpackbits(z["frozen"])[:N]
    crcm = np.unpackbits(z["crc_matrix"], axis=-1)[:, :K] if crc else np.zeros((0, K), np.uint8)
    return dict(frozen=frozen, order=z["order"], crc_matrix=np.ascontiguousarray(crcm), bitrev=z["bitrev"])


def load_decode(path):
    z = np.load(path)
    K = int(z["K"])
    d = dict(n=int(z["n"]), K=K, crc=int(z["crc"]), L=int(z["L"]), llr=z["llr"],
             info=np.unpackbits(z["info"], axis=-1)[:, :K], decoded=np.unpackbits(z["decoded"], axis=-1)[:, :K])
    d["name"] = os.path.basename(path)[len("decode_"):-len(".npz")]
    return d


def decode_fixture_paths():
    return sorted(glob.glob(os.path.join(GOLDEN, "decode_*.npz")))


def load_edge(n, K, crc):
    z = np.load(os.path.join(GOLDEN, "edge_n%d_K%d_crc%d.npz" % (n, K, crc)))
    out = dict(n=n, K=K, crc=crc, llr=z["llr"])
    for L in (1, 2, 4, 32):
        out[L] = np.unpackbits(z["decoded_L%d" % L], axis=-1)[:, :K]
    return out


def p1_fixture_paths():
    return sorted(glob.glob(os.path.join(GOLDEN, "p1_*.npz")))


def load_p1(path):
    z = np.load(path)
    K = int(z["K"])
    return dict(n=int(z["n"]), K=K, crc=int(z["crc"]), L=int(z["L"]), n_edge=int(z["n_edge"]), p1=z["p1"], p0=z["p0"],
                decoded=np.unpackbits(z["decoded"], axis=-1)[:, :K])


CONSTRUCTIONS = [(9, 256, 0), (9, 256, 16), (11, 1024, 0), (11, 1024, 16), (5, 16, 4), (7, 64, 8), (10, 512, 8)]


@pytest.fixture(scope="session")
def native_libs():
    """Build (if stale) and return the in-tree native libraries; needed by GPU and ABI tests."""
    from polar_b200 import build
    return build.build()
