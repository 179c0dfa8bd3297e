"""The module's scalar functions run on the HOST (no GPU, no launch): sigmoid_forward / sigmoid_backward / t_conorm_forward /
t_conorm_backward of the drop-in extension module are the host instantiation of the templates the kernels use, exactly as the
reference's are (K.cu:1230-1270, K.cpp:195-236; animations/distributions_to_csv.py:19 loops over thousands of points).
Checked against the C oracle and, where /root/reference is available, against the reference's own host functions."""
import math
import time

import pytest

import scenes
from gendr_b200.cuda import generalized_renderer as ext

XS = (0.0, 1e-3, 0.01, 0.05, 0.1, 0.3, 0.7, 1.0, 1.3, 2.0)


def _check(oracle):
    for did in range(18):
        shape = 2.0 if did in (14, 15) else 0.0
        for sign in (-1.0, 1.0):
            for x in XS:
                for shift in ((0.0,) if did < 12 else (0.0, 0.5)):
                    e, g = oracle.sigmoid_forward(did, sign, x * 0.1, 0.1, shape, shift), ext.sigmoid_forward(did, sign, x * 0.1, 0.1, shape, shift)
                    assert (math.isnan(e) and math.isnan(g)) or abs(g - e) <= 2e-7 + 2e-6 * abs(e), ('cdf', did, sign, x, shift, g, e)
                    e, g = oracle.sigmoid_backward(did, sign, x * 0.1, 0.1, shape, shift), ext.sigmoid_backward(did, sign, x * 0.1, 0.1, shape, shift)
                    assert (math.isnan(e) and math.isnan(g)) or abs(g - e) <= 1e-6 * max(1.0, abs(e)) + 2e-5 * abs(e), ('pdf', did, sign, x, shift, g, e)
    for tname, p in scenes.TCN_SWEEP[1:]:
        tid = scenes.TCN_SWEEP.index((tname, p))
        p = 0.0 if p is None else p
        for a in (0.0, 1e-4, 0.05, 0.3, 0.6, 0.95, 0.9999):
            for b in (1e-5, 0.01, 0.3, 0.6, 0.99):
                e, g = oracle.t_conorm_forward(tid, a, b, 0, p), ext.t_conorm_forward(tid, a, b, 0, p)
                assert abs(g - e) <= 2e-6, ('fold', tname, a, b, g, e)
                A = max(a, b)
                e, g = oracle.t_conorm_backward(tid, A, b, 0, p), ext.t_conorm_backward(tid, A, b, 0, p)
                assert abs(g - e) <= 1e-4 * max(1.0, abs(e)), ('dS', tname, A, b, g, e)


def test_host_scalars_vs_port_oracle(port_oracle):
    port_oracle.lib.gendr_oracle_set_mode(0)
    _check(port_oracle)


def test_host_scalars_vs_reference_host_functions(ref_oracle):
    """oracle/_ref exports the reference's own sigmoid_*_cuda / t_conorm_*_cuda host instantiations (unmodified source)."""
    _check(ref_oracle)


def test_survey_known_answers_on_host():
    for tid, p, want in ((2, 0., 0.72), (3, 0., 0.7627118), (4, 2., 0.7627119), (5, 2., 0.7375257), (6, 2., 0.6708204),
                         (7, 2., 0.6259115), (8, 2., 0.6093786), (9, -2., 0.6296504)):
        assert abs(ext.t_conorm_forward(tid, 0.3, 0.6, 0, p) - want) < 2e-6
    for did, lo, hi, pdf in ((4, 0.460172, 0.539828, 3.969525), (6, 0.475021, 0.524979, 2.49376), (8, 0.468274, 0.531726, 3.151583), (1, 0.45, 0.55, 5.0)):
        assert abs(ext.sigmoid_forward(did, -1., .01, .1, 1., 0.) - lo) < 2e-6
        assert abs(ext.sigmoid_forward(did, 1., .01, .1, 1., 0.) - hi) < 2e-6
        assert abs(ext.sigmoid_backward(did, 1., .01, .1, 1., 0.) - pdf) < 2e-5


def test_invalid_ids_and_speed():
    assert math.isnan(ext.sigmoid_forward(18, 1.0, 0.1, 0.1, 0.0, 0.0)) and math.isnan(ext.t_conorm_forward(0, 0.1, 0.2, 0, 0.0))
    assert math.isnan(ext.t_conorm_forward(10, 0.1, 0.2, 0, 0.0))
    t0 = time.perf_counter()
    for i in range(2000):                      # the CSV scripts of the reference evaluate thousands of points
        ext.sigmoid_forward(4, -1.0, i * 1e-4, 0.05, 0.0, 0.0)
    assert (time.perf_counter() - t0) / 2000 < 200e-6, 'scalar functions must not launch kernels'
