"""Decision-boundary bookkeeping shared by tests/ and bench.py's parity check.

TEST INFRASTRUCTURE ONLY (like the rest of oracle/): never imported by the product.

north star: demapped bits bit-exact; decisions whose float64 point lies within 1e-5 of a QPSK decision
boundary are listed separately.  The margin is ABSOLUTE on the oracle's float64 constellation point
(distance of the deciding component from 0: b0 <- imag, b1 <- real, OFDM.py:484-500).  For points of
magnitude above 1 a second class scales the band with the magnitude, because the fp32 constellation
tolerance (1e-4 relative) does; everything outside both classes is a parity failure.
"""
import numpy as np

BOUNDARY_TOL = 1e-5
EQ_RTOL = 1e-4


def classify_bit_diffs(got, ref_bits, ref_eq_data):
    """got / ref_bits: demapped bit vectors (same order); ref_eq_data: the oracle's float64 equalised
    points of the data carriers, one per bit pair, in bit order.  Returns a dict
      n_bits, n_diff,
      near_1e5     differing decisions within BOUNDARY_TOL (absolute) of a boundary,
      near_scaled  further differing decisions within BOUNDARY_TOL * |point| (|point| > 1),
      within_eq_tol  further differing decisions whose margin is below EQ_RTOL * |point|: the float32 constellation
                   tolerance the north star grants (1e-4 relative) is itself larger than the margin, which happens
                   for noise-amplified points in deep channel nulls (|point| in the tens); reported, and treated as
                   a failure by the tests,
      beyond       differing decisions outside all of these (parity failures),
      worst_margin largest margin of a differing decision,
      n_points_near_1e5  oracle points within BOUNDARY_TOL of a boundary, differing or not."""
    got = np.asarray(got).reshape(-1)
    ref_bits = np.asarray(ref_bits).reshape(-1)
    assert got.shape == ref_bits.shape, (got.shape, ref_bits.shape)
    pts = np.asarray(ref_eq_data).reshape(-1)
    assert 2 * len(pts) == len(got)
    allm = np.minimum(np.abs(pts.real), np.abs(pts.imag))
    bad = np.flatnonzero(got != ref_bits)
    out = dict(n_bits=int(len(got)), n_diff=int(len(bad)), near_1e5=0, near_scaled=0, within_eq_tol=0, beyond=0, worst_margin=0.0,
               n_points_near_1e5=int(np.sum(allm < BOUNDARY_TOL)))
    if len(bad):
        pt = pts[bad // 2]
        comp = np.where(bad % 2 == 0, np.abs(pt.imag), np.abs(pt.real))
        strict = comp < BOUNDARY_TOL
        scaled = ~strict & (comp < BOUNDARY_TOL * np.maximum(1.0, np.abs(pt)))
        eqtol = ~strict & ~scaled & (comp < EQ_RTOL * np.abs(pt))
        out.update(near_1e5=int(strict.sum()), near_scaled=int(scaled.sum()), within_eq_tol=int(eqtol.sum()),
                   beyond=int((~strict & ~scaled & ~eqtol).sum()), worst_margin=float(comp.max()))
    return out
