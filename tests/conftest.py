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


# ----------------------------------------------------------------------------- decision-boundary policy
from oracle.parity import BOUNDARY_TOL, classify_bit_diffs  # noqa: E402,F401  (policy and wording: oracle/parity.py)


def assert_bits_match(got, ref_bits, ref_eq_data, what):
    """Bit-exact except decisions within BOUNDARY_TOL of a boundary; those are counted and printed."""
    c = classify_bit_diffs(got, ref_bits, ref_eq_data)
    print("%s: %d bits, %d differ (%d within 1e-5 of a boundary, %d within 1e-5*|point|, %d within the 1e-4*|point| constellation "
          "tolerance, %d beyond; worst margin %.2e); %d oracle points lie within 1e-5 of a boundary"
          % (what, c["n_bits"], c["n_diff"], c["near_1e5"], c["near_scaled"], c["within_eq_tol"], c["beyond"], c["worst_margin"],
             c["n_points_near_1e5"]))
    far = c["beyond"] + c["within_eq_tol"]
    assert far == 0, "%s: %d bit mismatches away from decision boundaries (worst margin %.3e)" % (what, far, c["worst_margin"])
    return c


def kat4_regenerate(g, known_sequence):
    """KAT-4 input (BASELINE.json configs[1]): the oracle's pinned transmit() of the golden payload with the
    golden seed, then the same channel as oracle/make_golden.py -> int16 recording.  The fixture's
    sha256 of the reference-made recording pins the regeneration."""
    import hashlib
    from oracle import gf3_oracle as orc
    from oracle.make_golden import kat4_signal
    p = orc.Params.from_mode("A2", known_sequence=known_sequence, encoding="XOR")
    bits_in = orc.load_file_bits("gr5ch2.wav", g["payload"])
    np.random.seed(int(g["seed"]))
    tx = orc.transmit(p, bits_in)
    r = kat4_signal(tx, g["gr5channel"], int(g["noise_seed"]))
    assert hashlib.sha256(r.tobytes()).hexdigest() == str(g["r_sha256"]), "KAT-4 recording differs from the reference-made one"
    return p, bits_in, r
