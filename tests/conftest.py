"""pytest configuration: markers, import paths, golden-fixture helpers."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "gf3-audio-modem_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, PKG_DIR):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def known_sequence():
    from oracle import gf3_oracle as orc
    return orc.load_known_sequence(os.path.join(PKG_DIR, "gf3b200", "known_sequence_4096.txt"))


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def oracle_params(cfg, known_sequence, encoding="XOR"):
    from oracle import gf3_oracle as orc
    N, cp, lo, hi, P, L = (int(x) for x in cfg[:6])
    return orc.Params(N=N, cp=cp, lo=lo, hi=hi, n_pilots=P, packet_len=L,
                      encoding=encoding, known_sequence=known_sequence)


STAGE_NAMES = ["w1024", "a2_4096", "b1_4096", "n2048"]
