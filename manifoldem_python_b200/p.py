"""Module-global configuration surface read by the hot path — same attribute names as the
reference's modules/p.py (:13-114 init, :116-225 create_dir), restricted to what the distance and
embedding stages touch (SURVEY.md §5 'Config / flag system').  When the reference's own `p`
module is importable first on sys.path, the drop-in modules use that one instead (see _cfg())."""
import os

import numpy as np


def init():
    g = globals()
    g.update(proj_name='', user_dir='', resProj=0, relion_data=False, ncpu=1, machinefile=False, eps=1e-10,
             avg_vol_file='', img_stack_file='', align_param_file='', mask_vol_file='', num_part=0,
             Cs=0.0, EkV=0.0, AmpContrast=0.0, gaussEnv=np.inf, nPix=0, pix_size=0.0,
             PDsizeThL=100, PDsizeThH=2000, numberofJobs=0,
             num_eigs=15, num_psiTrunc=8, num_psis=8, tune=3, rad=5, conOrderRange=50, nClass=50, trajName='1',
             psi2_dir='', psi2_prog='', psi2_file='', EL_dir='', EL_prog='', EL_file='',
             out_dir='', tess_file='', dist_dir='', dist_prog='', dist_file='', psi_dir='', psi_prog='', psi_file='')
    return None


def create_dir():
    """distances/ and diff_maps/ trees with their progress/ marker directories (p.py:125-133, :205-208)."""
    g = globals()
    out = os.path.join(g['user_dir'], 'outputs_{}'.format(g['proj_name']))
    g['out_dir'] = out
    g['dist_dir'] = os.path.join(out, 'distances/')
    g['dist_prog'] = os.path.join(g['dist_dir'], 'progress/')
    g['psi_dir'] = os.path.join(out, 'diff_maps/')
    g['psi_prog'] = os.path.join(g['psi_dir'], 'progress/')
    for d in (g['dist_prog'], g['psi_prog']):
        os.makedirs(d, exist_ok=True)
    g['tess_file'] = os.path.join(out, 'selecGCs')
    g['dist_file'] = '{}/IMGs_'.format(g['dist_dir'])
    g['psi_file'] = '{}/gC_trimmed_psi_'.format(g['psi_dir'])
    return None


init()
