import numpy as np, sys
sys.path.insert(0, ".")
import vireo_b200 as vb
from oracle import vireo_oracle as O
AD, DP, _, _ = O.synth_counts(1500, 1200, 16, density=0.05, seed=2)
np.random.seed(1)
m = vb.Vireo(n_cell=1500, n_var=1200, n_donor=16)
m.fit(AD, DP, max_iter=2, min_iter=2, delay_fit_theta=1, verbose=False)
print("fit ok", m.ELBO_[-1])
