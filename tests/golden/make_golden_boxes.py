"""Compact golden vectors at the BASELINE box sizes (128, 256, 320) from the UNMODIFIED reference.

    python tests/golden/make_golden_boxes.py        (build container only: needs /root/reference)

The small-box vectors of make_golden.py pin the algorithm; these pin it where the library takes its size-specific kernels
(FFT-128 / 256 / 320, the rotation kernels for boxes that are a multiple of 32).  To stay small the files do not hold the
stack (synthetic.make_pd(nS, N, seed, snr) regenerates it; a float64 checksum is stored) and keep float32 copies of D's
inputs only where needed: ref_D (float64, nS x nS), ref_imgAvg (float32), every 13th pixel of ref_imgAll (float32) and
ref_imgAllIntensity.
Outputs: tests/golden/pd_box_N{128,256,320}.npz.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_golden as mg   # noqa: E402

CASES = [(128, 6, 11, 0.5), (256, 5, 12, 0.5), (320, 4, 13, 0.5)]      # N, nS, seed, snr
STRIDE = 13


def main():
    from manifoldem_python_b200 import synthetic
    ref = mg.load_reference()
    with tempfile.TemporaryDirectory() as tmp:
        for N, nS, seed, snr in CASES:
            pd = synthetic.make_pd(nS, N, seed=seed, snr=snr)
            res = mg.run_reference_pd(ref, pd, N, tmp=tmp)
            flat = lambda a: np.asarray(a, dtype=np.float64).reshape(nS, -1)     # noqa: E731
            np.savez_compressed(
                os.path.join(HERE, 'pd_box_N%d.npz' % N), N=N, nS=nS, seed=seed, snr=snr, stride=STRIDE,
                stack_sum=float(np.asarray(pd['stack'], dtype=np.float64).sum()),
                stack_abs_sum=float(np.abs(np.asarray(pd['stack'], dtype=np.float64)).sum()),
                ref_D=np.asarray(res['D'], dtype=np.float64),
                ref_imgAvg=np.asarray(res['imgAvg'], dtype=np.float32),
                ref_imgAll_sub=flat(res['imgAll'])[:, ::STRIDE].astype(np.float32),
                ref_imgAllIntensity=np.asarray(res['imgAllIntensity'], dtype=np.float32),
                ref_Psis=np.asarray(res['Psis'], dtype=np.float64), ref_PD=np.asarray(res['PD'], dtype=np.float64))
            print('N=%d nS=%d: D max %.4g' % (N, nS, np.asarray(res['D']).max()))


if __name__ == '__main__':
    main()
