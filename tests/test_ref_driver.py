"""oracle/ref_driver.py (the CPU cost model used by bench.py's reference arm) must reproduce the
live reference bit-for-bit.  Needs the reference tree -> only runs in the build container."""
import copy
import os
import sys

import numpy as np
import pytest

import ref_shims

pytestmark = pytest.mark.skipif(not ref_shims.reference_available(),
                                reason="reference sources not present (GPU box)")


def _mk_samplers(M, T, n, step, tune, interval):
    return [[M(step_size=step, tune=tune, tune_interval=interval) for _ in range(n)] for _ in range(T)]


@pytest.mark.parametrize("directed", [False, True])
def test_hot_path_sweep_matches_live_reference(directed):
    ref_shims.load_reference()
    import ref_driver as R
    from dynetlsm.metropolis import Metropolis
    from dynetlsm.sample_latent_positions import sample_latent_positions_mixture, sample_latent_positions
    from dynetlsm.sample_coefficients import sample_intercepts, sample_radii
    from dynetlsm.sample_labels import sample_labels_block
    from dynetlsm.latent_space import calculate_distances
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    name = "cfg1"
    w = bench.make_workload(name)
    T, n, d = w["X"].shape
    K = 4
    rng0 = np.random.RandomState(0)
    scale = 1.0 / n if directed else 1.0
    X0 = w["X"] * scale
    Y = w["Y"]
    if directed:
        Y = (rng0.rand(T, n, n) < 0.3).astype(np.float64)
        for t in range(T):
            np.fill_diagonal(Y[t], 0)
    radii = rng0.dirichlet(np.ones(n) * 5) if directed else None
    ic = np.array([0.3, 0.6]) if directed else np.array([0.8])
    mu, sigma = rng0.randn(K, d) * scale, rng0.gamma(2, 1, K) * scale ** 2
    z = rng0.randint(0, K, (T, n))
    wt = rng0.dirichlet(np.ones(K), size=(T, K))
    lm = np.array([0.8])
    step = 0.1 * scale

    # live reference
    rng = np.random.RandomState(5)
    Xr = X0.copy()
    icr = ic.copy()
    smp = _mk_samplers(Metropolis, T, n, step, 20, 3)
    ism = [Metropolis(step_size=0.1, tune=20) for _ in range(ic.size)]
    rsm = Metropolis(step_size=175000, tune=None, proposal_type="dirichlet")
    zr, rr = z.copy(), None if radii is None else radii.copy()
    for it in range(4):
        Xr = sample_latent_positions_mixture(Y, Xr, intercept=icr, mu=mu, sigma=sigma, lmbda=lm, z=zr,
                                             radii=rr, samplers=smp, is_directed=directed,
                                             squared=False, random_state=rng)
        Xr -= np.mean(Xr, axis=(0, 1))
        dist = calculate_distances(Xr)
        icr = sample_intercepts(Y, Xr, icr, intercept_prior=ic, intercept_variance_prior=2.0,
                                samplers=ism, radii=rr, dist=dist, is_directed=directed,
                                random_state=rng)
        if directed:
            rr = sample_radii(Y, Xr, intercepts=icr, radii=rr, sampler=rsm, dist=dist, random_state=rng)
        zr, ncr, nkr, _ = sample_labels_block(Xr, mu, sigma, lm, wt, random_state=rng)

    # restated driver over the same compiled kernels
    rng = np.random.RandomState(5)
    st = R.make_state(Y, X0, ic, is_directed=directed, radii=radii,
                      mixture=(mu, sigma, lm, z.copy(), wt), step_X=step, tune=20, tune_interval=3)
    for it in range(4):
        R.hot_path_sweep(st, rng)
    assert np.array_equal(st["X"], Xr)
    assert np.array_equal(st["intercept"], icr)
    assert np.array_equal(st["z"], zr)
    assert np.array_equal(st["ncount"], ncr)
    if directed:
        assert np.array_equal(st["radii"], rr)
    assert st["samplers"][1][3].step_size == smp[1][3].step_size

    # LSM prior variant
    rng = np.random.RandomState(6)
    Xr = X0.copy()
    smp = _mk_samplers(Metropolis, T, n, step, 20, 3)
    Xr = sample_latent_positions(Y, Xr, intercept=ic, tau_sq=2.0 * scale ** 2, sigma_sq=0.1 * scale ** 2,
                                 samplers=smp, radii=radii, is_directed=directed, random_state=rng)
    rng = np.random.RandomState(6)
    st = R.make_state(Y, X0, ic, is_directed=directed, radii=radii, step_X=step, tune=20,
                      tune_interval=3, tau_sq=2.0 * scale ** 2, sigma_sq=0.1 * scale ** 2)
    Xd = R.latent_sweep(Y, st["X"], st["intercept"], st["samplers"], rng, radii=radii,
                        is_directed=directed, tau_sq=st["tau_sq"], sigma_sq=st["sigma_sq"])
    assert np.array_equal(Xd, Xr)
