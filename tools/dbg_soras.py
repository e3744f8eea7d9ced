import sys; sys.path.insert(0, '.')
import numpy as np
import geneo4petsc_b200 as g
from oracle import geneo_oracle as go
mesh = go.gen_grid(3, 12, 1e-4, 2.0, "lin")
p = g.Problem().set_mesh(mesh.nb_node, mesh.elem_ptr, mesh.elem_idx, mesh.mat_val).decompose(4, True, 0)
pc = g.GeneoPC(["-geneo_lvl", "SORAS,2", "-geneo_tau", "0.3", "-geneo_optim", "0.5"]).setup(p)
rep = go.run_case(mesh, 4, go.GenEOOptions(lvl1="SORAS", lvl2="2", tau=0.3, optim=0.5), part=p.partition(), ksp="cg", rtol=1e-6)
for s in range(4):
    si = pc.sub_info(s)
    o = rep.pc.sub[s]
    ev = pc.sub_eigenvalues(s)
    print(s, si, 'oracle nev', o.z.shape[1], 'estim', o.estim, 'tauLoc', o.tau_loc, 'gammaLoc', o.gamma_loc)
    oe = np.array(o.eigvals)
    print('   mine  tau-part', np.sort(ev[ev < 1.0])[:6], ' n_gamma', int(np.sum(ev >= 1.0)), 'min gamma ev', ev[ev >= 1.0].min() if np.any(ev >= 1.0) else None)
    print('   oracle tau-part', np.sort(oe[oe < 1.0])[:6], ' n_gamma', int(np.sum(oe >= 1.0)), 'min gamma ev', oe[oe >= 1.0].min() if np.any(oe >= 1.0) else None)
