"""Golden chains of the reference's finite-mixture estimator ``DynamicNetworkLPCM`` (lpcm.py:134-873).

TEST INFRASTRUCTURE ONLY.  Runs the UNMODIFIED reference (imported from /root/reference through
``ref_shims``, Cython kernels from oracle/_ref) in THIS container and writes
tests/golden/lpcm_*.npz.  Per sweep it records

  * what the reference hands to its conjugate block (lpcm.py:567-656): centred positions, labels,
    transition counts ``n``, occupancies ``nk`` and the numpy RandomState at that point, and
  * what the block leaves behind (initial / transition weights, means, variances, lambda, the two
    scale hyper-priors) plus the joint log-posterior of the stored sample,

so the product's host block can be pinned draw for draw on the CPU
(tests/test_lpcm_host.py) and the whole ``fit`` on the GPU (tests/test_gpu_lpcm.py).  The stored
traces of the model are taken BEFORE the post-hoc Procrustes rotation (lpcm.py:739-745).

  python oracle/make_golden_lpcm.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

import ref_shims  # noqa: E402


def _passthrough(it, *a, **k):
    return it


def run_lpcm(Y, keep, **kw):
    import dynetlsm.lsm as L
    import dynetlsm.lpcm as P
    L.tqdm = _passthrough
    P.tqdm = _passthrough
    sweeps, cur = [], {}
    orig_lab = P.sample_labels_block_lpcm
    orig_logp = P.DynamicNetworkLPCM.logp

    def labels(X, mu, sigma, lmbda, init_weights, trans_weights, random_state=None):
        z, n, nk, resp = orig_lab(X, mu, sigma, lmbda, init_weights, trans_weights,
                                  random_state=random_state)
        st = random_state.get_state()
        cur.clear()
        cur.update(X_centered=X.copy(), z_out=z.astype(np.int32), n_out=n.copy(),
                   nk_out=nk.astype(np.int32), mu_in=mu.copy(), sigma_in=sigma.copy(),
                   lmbda_in=np.array(lmbda, dtype=np.float64).copy(),
                   rng_keys=st[1].copy(), rng_pos=np.int64(st[2]),
                   rng_has_gauss=np.int64(st[3]), rng_gauss=np.float64(st[4]))
        return z, n, nk, resp

    def logp(self, X, intercept, mu, sigma, z, init_weights, trans_weights, lmbda, **k):
        v = orig_logp(self, X, intercept, mu, sigma, z, init_weights, trans_weights, lmbda, **k)
        if "z_out" in cur:
            rec = dict(cur)
            cur.clear()
            rec.update(logp=float(np.ravel(v)[0]), mu_next=mu.copy(), sigma_next=sigma.copy(),
                       lmbda_next=np.array(lmbda, dtype=np.float64).copy(),
                       init_next=init_weights.copy(), trans_next=trans_weights.copy(),
                       intercept_out=np.array(intercept, dtype=np.float64).copy(),
                       mean_variance_prior=np.float64(np.ravel(self.mean_variance_prior_)[0]),
                       b=np.float64(np.ravel(self.b_)[0]))
            if k.get("radii") is not None:
                rec["radii_out"] = k["radii"].copy()
            sweeps.append(rec)
        return v

    P.sample_labels_block_lpcm = labels
    P.DynamicNetworkLPCM.logp = logp
    try:
        model = P.DynamicNetworkLPCM(**kw).fit(Y)
    finally:
        P.sample_labels_block_lpcm = orig_lab
        P.DynamicNetworkLPCM.logp = orig_logp
    rec = {k: np.stack([np.asarray(s[k]) for s in sweeps[:keep]]) for k in sweeps[0]}
    rec["Y"] = model.Y_fit_.astype(np.int8)
    rec["logps"] = model.logps_[:keep + 1].copy()
    rec["selected_id"] = np.int64(model.selected_id_)
    rec["cooccurrence_probas"] = model.cooccurrence_probas_.copy()
    rec["z_hat"] = model.z_.astype(np.int32)
    rec["a"] = np.float64(model.a)
    rec["intercept_prior"] = np.atleast_1d(np.asarray(model.intercept_prior, dtype=np.float64)).copy()
    rec["intercept_variance_prior"] = np.float64(model.intercept_variance_prior)
    for nm in ("a0_", "b0_", "c0_", "d0_", "dirichlet_prior_"):
        rec[nm] = np.float64(getattr(model, nm))
    return rec, model


def main():
    os.makedirs(OUT, exist_ok=True)
    ref_shims.load_reference()
    from dynetlsm.datasets import load_monks, simple_splitting_dynamic_network

    # undirected, the community-splitting network of the HDP golden (make_golden.py)
    Ys, _ = simple_splitting_dynamic_network(n_nodes=36, n_time_steps=3, random_state=42)
    Ys = np.ascontiguousarray(Ys[:4])
    kw = dict(n_iter=25, tune=25, burn=10, tune_interval=8, n_components=4, random_state=3)
    rec, model = run_lpcm(Ys, keep=59, **kw)
    np.savez_compressed(os.path.join(OUT, "lpcm_undirected_split.npz"), **rec)

    # directed with radii, Sampson's monks, posterior-expected-VI point estimate
    Yd, _, _ = load_monks(is_directed=True)
    kw = dict(n_iter=15, tune=15, burn=10, tune_interval=5, n_components=3, is_directed=True,
              selection_type="vi", random_state=5)
    rec, model = run_lpcm(Yd, keep=39, **kw)
    np.savez_compressed(os.path.join(OUT, "lpcm_directed_monks.npz"), **rec)

    for f in sorted(os.listdir(OUT)):
        if f.startswith("lpcm_"):
            print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KB")


if __name__ == "__main__":
    main()
