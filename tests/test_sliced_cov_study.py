"""CPU: the paper study of an integer-sliced covariance for M = 16 (tools/emulate_sliced_cov.py): the
fixed-point alignment error shrinks by 2^-7 per slice, and six 7-bit slices reach the rounding level of the fp64 path."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import emulate_sliced_cov as sc  # noqa: E402


def test_alignment_error_per_slice_count():
    rows, _ = sc.study(windows=3)
    by = {r[0]: r for r in rows}
    assert by[4][2] <= 1e-8 and by[6][2] <= 1e-12          # max |dR| / max |R|
    assert by[4][3] <= 1e-5 and by[6][3] <= 1e-10          # max relative error of P(theta): the 1e-5 gate needs >= 4 slices
    for s in (4, 5, 6, 7):
        assert by[s][2] < by[s - 1][2] / 20.0              # ~2^-7 per slice
