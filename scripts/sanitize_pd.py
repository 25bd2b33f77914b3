"""One small PD at the bench box size through the device API, for compute-sanitizer:
    compute-sanitizer --tool memcheck  python scripts/sanitize_pd.py
    compute-sanitizer --tool racecheck python scripts/sanitize_pd.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifoldem_python_b200 import pd_stage, synthetic   # noqa: E402

nS, N = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (12, 256)
pd = synthetic.make_pd(nS, N, seed=5, snr=0.5)
em = pd['em']
yy, xx = np.mgrid[:N, :N]
msk2 = ((yy - N / 2) ** 2 / (0.42 * N) ** 2 + (xx - N / 2) ** 2 / (0.3 * N) ** 2) < 1
for m in (None, msk2):
    res = pd_stage.run_pd(pd['ind'], pd['q'], pd['df'], pd['stack'], pd['nStot'], N, em['pix_size'], em['Cs'], em['EkV'],
                          em['AmpContrast'], msk2=m)
    print('D', res['D'].shape, float(res['D'].max()), 'finite', bool(np.isfinite(res['D']).all()))
