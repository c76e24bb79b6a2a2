import sys, warnings
sys.path.insert(0, "."); sys.path.insert(0, "tests")
warnings.filterwarnings("ignore")
import numpy as np
from test_gpu_estimators import _splitting_network
from dynetlsm_b200 import DynamicNetworkHDPLPCM
from dynetlsm_b200.diagnostics import ess
Y = _splitting_network(n=36, T=3, seed=7)
kw = dict(n_iter=700, tune=300, burn=300, n_features=2, n_components=6)
for ds, rs in ((5, (11, 12)), (6, (13, 14)), (7, (15, 16))):
    dev = DynamicNetworkHDPLPCM(random_state=ds, sampler="device", n_chains=6, **kw).fit(Y)
    reps = [DynamicNetworkHDPLPCM(random_state=s, sampler="replay", **kw).fit(Y) for s in rs]
    nb = 600
    a = dev.chains_["intercepts"][:, nb:, 0]; b = np.stack([r.intercepts_[nb:, 0] for r in reps])
    se = np.sqrt(a.var() / max(ess(a), 10) + b.var() / max(ess(b), 10))
    la, lb = dev.chains_["lambdas"][:, nb:], np.stack([r.lambdas_[nb:, 0] for r in reps])
    ka = dev.chains_["n_clusters"][:, nb:].mean(); kb = np.mean([[np.unique(zz).size for zz in r.zs_[nb:]] for r in reps])
    cooc = lambda zs: (zs[:, :, :, None] == zs[:, :, None, :]).mean(axis=0)
    ca = np.mean([cooc(dev.chains_["zs"][c, nb:]) for c in range(6)], axis=0); cb = np.mean([cooc(r.zs_[nb:]) for r in reps], axis=0)
    truth = np.random.RandomState(7).randint(0, 2, 36); same = truth[:, None] == truth[None, :]
    print("icpt diff %.4f (5se+.03=%.4f) lam %.4f k %.3f/%.3f cooc %.4f sep %.3f" % (abs(a.mean()-b.mean()), 5*se+0.03, abs(la.mean()-lb.mean()), ka, kb, np.abs(ca-cb).mean(), ca[0][same].mean()-ca[0][~same].mean()))
