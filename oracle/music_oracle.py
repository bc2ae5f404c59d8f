"""CPU oracle (numpy, fp64) for gr-baz's MUSIC DOA block.  TEST INFRASTRUCTURE ONLY.

This file is a restatement of the reference algorithm; it is *not* product code.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and there only as the checker.  The product path
(``gr-baz_b200``) never routes through it.

PARITY UNPINNED BY THE REFERENCE.  The reference ships no tests, golden vectors or
fixtures for this block (``python/qa_baz.py:26-56`` and
``lib/qa_baz.cc:32-41`` are empty stubs), and the reference itself cannot be built or
imported here (GNU Radio, Boost, Armadillo, SWIG, Python 2 absent).  The arithmetic of
the path lives in Armadillo (un-vendored, no version pinned:
``CMakeLists-3.7.txt:195``, ``cmake/Modules/FindArmadillo.cmake:45-73``) on top of the
system LAPACK (``eig_sym`` -> zheev/zheevd).  This restatement therefore follows the
reference *source text* line by line and uses numpy's LAPACK ``zheevd`` binding
(``numpy.linalg.eigh``) - the same routine family Armadillo dispatches to - for the
eigendecomposition.  Our own golden vectors (``tests/golden``) are produced by this file.
What narrows the gap: ``oracle/_ref`` compiles the reference's own ``lib/baz_music_doa.cc``
unmodified against stand-in GNU Radio / Armadillo headers (``oracle/Makefile`` target
``ref``) and ``tests/test_ref_shim.py`` checks this file against it - the reference's control
flow is exercised as written, Armadillo's own arithmetic is not.

Reference lines restated (all under /root/reference):
  lib/baz_music_doa.cc:72-161   work()
  lib/baz_music_doa.cc:35-53    constructor checks
  python/music_doa_helper.py:29-46,55-56   steering table
  swig/baz_swig.i:564           complex128 -> complex64 rounding of the table
"""
from __future__ import annotations

import numpy as np

C_LIGHT = 299792458.0  # python/music_doa_helper.py:55


# --------------------------------------------------------------------------------------
# Steering table: python/music_doa_helper.py:29-46 (+ the SWIG c128->c64 rounding)
# --------------------------------------------------------------------------------------
def unit_vect(theta):
    """python/music_doa_helper.py:29-30"""
    return np.array([np.cos(theta), np.sin(theta)])


def calculate_antenna_array_response_literal(antenna_array, angular_resolution, l):
    """Literal restatement of python/music_doa_helper.py:32-46 (K*M Python loop, c128).

    Returns a nested list [K][M] of Python complex, exactly what the helper hands to SWIG.
    """
    response = []
    for step in range(0, angular_resolution):
        angle = (step * 360.0 / angular_resolution) * (np.pi / 180.0)  # :36
        response_step = []
        for antenna in antenna_array:
            phase_offset = np.inner(antenna, unit_vect(angle)) / l  # :40
            antenna_response = np.exp(-1j * 2.0 * np.pi * phase_offset)  # :41
            response_step += [antenna_response]
        response += [response_step]
    return response


def scaled_antenna_array(array_spacing, antenna_array):
    """python/music_doa_helper.py:56: positions = spacing * [x, y]."""
    return [[array_spacing * x, array_spacing * y] for [x, y] in antenna_array]


def steering_table_c64(antenna_array_scaled, angular_resolution, wavelength):
    """Table as the C++ block stores it: [K][M] complex64 (swig/baz_swig.i:564 marshals
    Python complex (c128) into std::vector<std::vector<gr_complex>>, i.e. rounds to c64)."""
    resp = calculate_antenna_array_response_literal(antenna_array_scaled, angular_resolution, wavelength)
    return np.asarray(resp, dtype=np.complex128).astype(np.complex64)


# --------------------------------------------------------------------------------------
# work(): lib/baz_music_doa.cc:72-161
# --------------------------------------------------------------------------------------
def check_params(m, n, nsamples, resolution, table=None):
    """Constructor asserts, lib/baz_music_doa.cc:45-50.  n must also be >= 1 and < m
    (grc/baz_music_doa.xml doc: "it is necessary that n<m"; m == n underflows
    ``cols(0, m-n-1)`` at :93)."""
    if not (m > 0 and 1 <= n < m):
        raise ValueError("need m > 0 and 1 <= n < m")
    if not (nsamples > 0 and nsamples % m == 0):
        raise ValueError("nsamples must be a positive multiple of m")
    if not resolution > 0:
        raise ValueError("resolution must be > 0")
    if table is not None:
        if table.shape != (resolution, m):
            raise ValueError("array_response must be [resolution][m]")


def covariance(in_c64, m):
    """:74-85.  in_c64: (nsamples,) complex64.  x(r, c) = in[c*m + r] (column-major reshape)."""
    data = in_c64.astype(np.complex128)  # :75-77 widen
    average_over = data.shape[0] // m  # :83
    x = data.reshape(average_over, m).T  # :82-84 (m rows, average_over cols)
    R = (x @ x.conj().T) / float(average_over)  # :85
    return R


def pick_top_n_literal(P, n, resolution):
    """Literal restatement of the insertion loop, :95,129-141.  P: fp64 strengths."""
    vDOAs = [(0.0, 0.0, -1)] * n  # (angle, strength, bin); bin is ours, for the index gate
    for step in range(resolution):
        strength = P[step]
        for i in range(n):
            if strength > vDOAs[i][1]:  # :132, strict, NaN never true
                angle = float(step) * 360.0 / float(resolution)  # :134
                vDOAs.insert(i, (angle, float(strength), step))  # :136
                vDOAs.pop()  # :137
                break
    return vDOAs


def pick_top_n(P, n, resolution):
    """Vectorised equivalent of pick_top_n_literal: n largest strictly-positive (non-NaN)
    values, ordered (value desc, bin asc); unfilled slots stay (0, 0, -1)."""
    P = np.asarray(P, dtype=np.float64)
    valid = np.nonzero(P > 0.0)[0]  # NaN > 0 is False
    order = valid[np.lexsort((valid, -P[valid]))][:n]
    out = [(float(k) * 360.0 / float(resolution), float(P[k]), int(k)) for k in order]
    out += [(0.0, 0.0, -1)] * (n - len(out))
    return out


def pick_local_maxima(P, n, resolution, exclusion=0):
    """NOT in the reference (opt-in): the reference's rule (:129-141) returns the n
    largest BINS, which for n >= 2 are usually neighbours on the flank of one peak.  This mode returns the n
    largest circular LOCAL MAXIMA instead.  Definition (ours - there is no reference behaviour to match):
      * bin k is a candidate iff P[k] > 0, P[k] > P[k-1] and P[k] >= P[k+1] (indices mod K; NaN compares false;
        on a plateau only the lowest bin qualifies);
      * repeat n times: take the candidate with the largest P (ties: lowest k) whose circular distance to every
        peak already taken is > exclusion bins;
      * output order = pick order (strength descending); unfilled slots stay (0, 0, -1) like the reference's
        initial pairs (:95)."""
    P = np.asarray(P, dtype=np.float64)
    K = int(resolution)
    with np.errstate(invalid="ignore"):
        cand = np.nonzero((P > 0.0) & (P > np.roll(P, 1)) & (P >= np.roll(P, -1)))[0]
    order = cand[np.lexsort((cand, -P[cand]))]
    taken = []
    for k in order:
        if len(taken) == n:
            break
        if all(min((k - t) % K, (t - k) % K) > exclusion for t in taken):
            taken.append(int(k))
    out = [(float(k) * 360.0 / float(K), float(P[k]), int(k)) for k in taken]
    out += [(0.0, 0.0, -1)] * (n - len(out))
    return out


def reduce_angles(angles, levels=None, weighted=False):
    """NOT in the reference: per angle slot, the circular mean over the windows
    (degrees in [0, 360)), the mean resultant length and the total weight.  A window counts iff levels is None or
    its level is > 0; weighted = weight by level.  Returns float32 arrays (mean_deg[n], resultant[n], weight[n])."""
    a = np.asarray(angles, dtype=np.float32).astype(np.float64)
    W, n = a.shape
    lv = np.ones((W, n)) if levels is None else np.asarray(levels, dtype=np.float32).astype(np.float64)
    with np.errstate(invalid="ignore"):
        valid = lv > 0.0
    wt = np.where(valid, lv if weighted else 1.0, 0.0)
    rad = a * (np.pi / 180.0)
    S = np.sum(wt * np.sin(rad), axis=0)
    C = np.sum(wt * np.cos(rad), axis=0)
    Wt = np.sum(wt, axis=0)
    deg = np.degrees(np.arctan2(S, C))
    deg = np.where(deg < 0.0, deg + 360.0, deg)
    with np.errstate(invalid="ignore", divide="ignore"):
        r = np.where(Wt > 0.0, np.sqrt(S * S + C * C) / Wt, 0.0)
    deg = np.where(Wt > 0.0, deg, 0.0)
    return deg.astype(np.float32), r.astype(np.float32), Wt.astype(np.float32)


def reduce_spectrum(spectrum):
    """NOT in the reference: mean over the windows of the float32 spectra, accumulated in fp64."""
    return np.mean(np.asarray(spectrum, dtype=np.float32).astype(np.float64), axis=0).astype(np.float32)


def work(in_c64, m, n, table_c64, want_spectrum=True, literal_pick=False, return_internals=False):
    """One window through lib/baz_music_doa.cc:72-161.

    in_c64   : (nsamples,) complex64, sample-interleaved antennas.
    table_c64: (K, m) complex64 array response.
    Returns dict with float32 'angles'[n], 'levels'[n], int32 'bins'[n] (ours: -1 = unfilled),
    float32 'spectrum'[K] (the optional port 2) and fp64 'P'[K] (the pre-cast strengths).
    """
    in_c64 = np.ascontiguousarray(in_c64, dtype=np.complex64)
    table_c64 = np.ascontiguousarray(table_c64, dtype=np.complex64)
    K = table_c64.shape[0]
    check_params(m, n, in_c64.shape[0], K, table_c64)

    R = covariance(in_c64, m)
    eigvals, eigvec = np.linalg.eigh(R)  # :88-90 ascending eigenvalues, eigenvectors in columns
    G = eigvec[:, 0 : m - n]  # :93

    a = table_c64.astype(np.complex128)  # :110-112 widen per step
    v = a @ G.conj()  # row k = (G^H a_k)^T, :116/118
    # arma::norm(v, 2) for complex: sqrt(sum_i |v_i|^2) with |.| = std::abs
    nrm = np.sqrt(np.sum(np.abs(v) ** 2, axis=1))
    with np.errstate(divide="ignore", invalid="ignore"):
        P = 1.0 / (nrm * nrm)  # 1.0 / pow(norm, 2)

    pick = pick_top_n_literal if literal_pick else pick_top_n
    doas = pick(P, n, K)
    res = {
        "angles": np.array([d[0] for d in doas], dtype=np.float32),  # :153 double -> float
        "levels": np.array([d[1] for d in doas], dtype=np.float32),  # :154
        "bins": np.array([d[2] for d in doas], dtype=np.int32),
        "P": P,
    }
    if want_spectrum:
        with np.errstate(over="ignore"):
            res["spectrum"] = P.astype(np.float32)  # :121
    if return_internals:
        res["R"] = R
        res["eigvals"] = eigvals
        res["noise_projector"] = G @ G.conj().T  # phase-free
    return res


def work_batch(in_c64, m, n, table_c64, want_spectrum=False):
    """W windows; in_c64: (W, nsamples).  Same per-window results as calling work() W times
    (the block is stateless across windows)."""
    in_c64 = np.asarray(in_c64)
    W = in_c64.shape[0]
    K = table_c64.shape[0]
    angles = np.zeros((W, n), np.float32)
    levels = np.zeros((W, n), np.float32)
    bins = np.zeros((W, n), np.int32)
    P = np.zeros((W, K), np.float64)
    spec = np.zeros((W, K), np.float32) if want_spectrum else None
    for w in range(W):
        r = work(in_c64[w], m, n, table_c64, want_spectrum=want_spectrum)
        angles[w], levels[w], bins[w], P[w] = r["angles"], r["levels"], r["bins"], r["P"]
        if want_spectrum:
            spec[w] = r["spectrum"]
    return {"angles": angles, "levels": levels, "bins": bins, "P": P, "spectrum": spec}
