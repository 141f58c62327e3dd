import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
from bench import boss_blocks
from victor_b200 import CCFFit
model, data = boss_blocks()
fit = CCFFit(model, data, device=0)
prm = {"fsigma8": 0.47, "beta": 0.37, "sigma_v": 380.0, "epsilon": 1.0}
def med(fn, n=300):
    for _ in range(20): fn()
    t=[]
    for _ in range(n):
        t0=time.perf_counter(); fn(); t.append(time.perf_counter()-t0)
    return round(float(np.median(t))*1e6,1)
s = np.asarray(fit.s, float); mu = np.linspace(0,1,100)
print("theory_multipoles (dict)        us", med(lambda: fit.theory_multipoles(s, dict(prm), poles=[0,2])))
print("theory_multipole_vector         us", med(lambda: fit.theory_multipole_vector(s, dict(prm), [0,2])))
print("theory_xi 100x30                us", med(lambda: fit.theory_xi(s, mu, dict(prm))))
print("chi_squared                     us", med(lambda: fit.chi_squared(dict(prm))))
print("log_likelihood                  us", med(lambda: fit.log_likelihood(dict(prm))))
print("log_likelihood dispersion kw    us", med(lambda: fit.log_likelihood(dict(prm), rsd_model="dispersion")))
print("log_likelihood likelihood-interp us", med(lambda: fit.log_likelihood(dict(prm), beta_interpolation="likelihood")))
print("theory_xi_2D                    us", med(lambda: fit.theory_xi_2D(dict(prm)), n=50))
print("xi_2D_from_multipoles           us", med(lambda: fit.xi_2D_from_multipoles(dict(prm)), n=50))
fit.close()
