import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, torch
import conftest
from nf_b200.PhaseSpace.flat_phase_space_generator import FlatInvertiblePhasespace
np.set_printoptions(precision=17, linewidth=220)
for case in ('edge0_m4','edge0_m5'):
    g = conftest.Golden('rambo_%s.npz'%case); m=g.meta
    ps = FlatInvertiblePhasespace(m["initial"], m["final"], pdf=None, pdf_active=False)
    mom, w, mask = ps.generateKinematics_batch(m["E_cm"], g.t("r").cuda(), return_cutmask=True, **m["cuts"])
    w=w.cpu().numpy(); ref=g['weight']; r=g['r']; n=len(m['final'])
    for i in range(len(w)):
        if ref[i]!=0 and abs(w[i]/ref[i]-1)>1e-9: print(case,i,r[i,:n-2],w[i],ref[i],w[i]/ref[i]-1)
