"""Helpers for the compact golden vectors at the BASELINE box sizes (tests/golden/make_golden_boxes.py)."""
import os

import numpy as np


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _box_case(golden_dir, N):
    """A compact golden of tests/golden/make_golden_boxes.py: the PD is regenerated from its seed (checksums stored)."""
    from manifoldem_python_b200 import synthetic
    g = _load(golden_dir, 'pd_box_N%d.npz' % N)
    pd = synthetic.make_pd(int(g['nS']), N, seed=int(g['seed']), snr=float(g['snr']))
    st = np.asarray(pd['stack'], dtype=np.float64)
    assert abs(st.sum() - float(g['stack_sum'])) <= 1e-9 * float(g['stack_abs_sum'])
    assert abs(np.abs(st).sum() - float(g['stack_abs_sum'])) <= 1e-9 * float(g['stack_abs_sum'])
    return g, pd


def check_box_outputs(out, g, d_rtol, img_tol):
    """D, imgAvg, the stored pixel subset of imgAll and the intensity against a compact golden (float32 storage: 6e-8)."""
    nS, stride = int(g['nS']), int(g['stride'])
    D, Dr = np.asarray(out['D'], dtype=np.float64), g['ref_D']
    off = ~np.eye(nS, dtype=bool)
    assert (np.abs(D - Dr)[off] / Dr[off]).max() <= d_rtol
    for k, got in (('imgAvg', np.asarray(out['imgAvg'], dtype=np.float64)),
                   ('imgAll_sub', np.asarray(out['imgAll'], dtype=np.float64).reshape(nS, -1)[:, ::stride]),
                   ('imgAllIntensity', np.asarray(out['imgAllIntensity'], dtype=np.float64))):
        ref = np.asarray(g['ref_' + k], dtype=np.float64).reshape(got.shape)
        assert np.abs(got - ref).max() <= img_tol * np.abs(ref).max(), (k, np.abs(got - ref).max() / np.abs(ref).max())
    assert np.allclose(np.asarray(out['Psis']), g['ref_Psis'], rtol=0, atol=1e-9)
