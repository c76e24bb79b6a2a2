"""The function-level seams keep the reference's signatures (sample_latent_positions.py:92/:149,
sample_coefficients.py:12/:91, sample_labels.py:134, the Cython kernels of network_likelihoods.py)
and, driven by the same RandomState, the reference's results: checked against the recorded
reference sweeps, with the RandomState rewound to the recorded draws."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


class RecordedRandomState(object):
    """Stands in for numpy's RandomState: hands back the reference's recorded raw draws."""

    def __init__(self, normals=None, uniforms=None, dirichlets=None):
        self.normals = list(np.asarray(normals).reshape(-1)) if normals is not None else []
        self.uniforms = list(np.asarray(uniforms).reshape(-1)) if uniforms is not None else []
        self.dirichlets = list(dirichlets) if dirichlets is not None else []

    def randn(self, d):
        out = np.array(self.normals[:d]); del self.normals[:d]
        return out

    def rand(self):
        return self.uniforms.pop(0)

    def dirichlet(self, alpha):
        return self.dirichlets.pop(0)

    def random_sample(self, shape):
        out = np.array(self.uniforms[:int(np.prod(shape))]).reshape(shape)
        del self.uniforms[:int(np.prod(shape))]
        return out


def _samplers(g, s, prefix, shape):
    from dynetlsm_b200 import Metropolis
    out = []
    step = np.reshape(g[prefix + "step"][s], shape)
    for idx in np.ndindex(*shape):
        m = Metropolis(step_size=float(step[idx]), tune=int(g["tune"]), tune_interval=int(g["tune_interval"]))
        m.n_accepted = int(np.reshape(g[prefix + "n_accepted"][s], shape)[idx])
        m.n_steps = int(np.reshape(g[prefix + "n_steps"][s], shape)[idx])
        m.steps_until_tune = int(np.reshape(g[prefix + "until"][s], shape)[idx])
        out.append(m)
    return np.array(out, dtype=object).reshape(shape)


def test_sample_latent_positions_seam_undirected(monkeypatch):
    import dynetlsm_b200.sample_latent_positions as SLP
    g = load_golden("lsm_undirected_monks.npz")
    Y = g["Y"].astype(np.float64)
    T, n, d = g["X_in"].shape[1:]
    monkeypatch.setattr(SLP, "check_random_state", lambda r: r)
    for s in (0, 7, 31, 60):
        grid = _samplers(g, s, "tuner_", (T, n))
        samplers = [list(row) for row in grid]
        X = g["X_in"][s].copy()
        rng = RecordedRandomState(g["eps"][s], np.exp(g["logu"][s]))
        # the recorded log-uniforms are exact; exp/log round-trips are not, so patch np.log's input:
        rng.uniforms = list(np.exp(g["logu"][s]).reshape(-1))
        out = SLP.sample_latent_positions(Y, X, g["intercept_in"][s], float(g["tau_sq"]),
                                          float(g["sigma_sq"]), samplers, random_state=rng)
        assert out is X
        agree = np.mean(np.all(X == g["X_out"][s], axis=-1))
        assert agree > 0.98          # log(exp(logu)) may differ from logu in the last ulp
        assert samplers[1][3].n_steps == int(g["tuner_n_steps"][s][1, 3]) + 1


def test_sample_labels_block_seam(monkeypatch):
    import dynetlsm_b200.sample_labels as SL
    g = load_golden("hdp_undirected_split.npz")
    monkeypatch.setattr(SL, "check_random_state", lambda r: r)
    for s in (0, 20, 58):
        rng = RecordedRandomState(uniforms=g["U"][s])
        z, cnt, nk, resp = SL.sample_labels_block(g["X_centered"][s], g["mu"][s], g["sigma"][s],
                                                  g["lmbda"][s], g["w"][s], random_state=rng)
        assert z.dtype == np.int64 and resp.shape == z.shape + (g["sigma"].shape[1],)
        assert np.array_equal(z, g["z_out"][s])
        assert np.array_equal(cnt, g["n_out"][s]) and np.array_equal(nk, g["nk_out"][s])
        assert np.array_equal(resp.sum(axis=2), np.ones_like(z))


def test_intercept_and_radii_seams(monkeypatch):
    import dynetlsm_b200.sample_coefficients as SC
    from dynetlsm_b200 import Metropolis
    g = load_golden("lsm_directed_monks.npz")
    Y = g["Y"].astype(np.float64)
    monkeypatch.setattr(SC, "check_random_state", lambda r: r)
    s = 12
    samplers = list(_samplers(g, s, "itun_", (2,)))
    rng = RecordedRandomState(g["i_eps"][s], np.exp(g["i_logu"][s]))
    ic = g["intercept_in"][s].copy()
    out = SC.sample_intercepts(Y, g["X_centered"][s], ic, intercept_prior=g["intercept_prior"],
                               intercept_variance_prior=float(g["intercept_variance_prior"]),
                               samplers=samplers, radii=g["radii_in"][s], is_directed=True,
                               random_state=rng)
    assert np.array_equal(out, g["intercept_out"][s])
    rs = Metropolis(step_size=float(g["rtun_step"][s][0]), tune=None, proposal_type="dirichlet")
    rng = RecordedRandomState(uniforms=[np.exp(g["r_logu"][s])], dirichlets=[g["r_proposal"][s].copy()])
    r = SC.sample_radii(Y, g["X_centered"][s], intercepts=out, radii=g["radii_in"][s].copy(),
                        sampler=rs, random_state=rng)
    assert np.array_equal(r, g["radii_out"][s])
    assert rs.n_steps == 1


def test_likelihood_seams(kernels_golden):
    from dynetlsm_b200.network_likelihoods import (partial_loglikelihood,
                                                   directed_partial_loglikelihood,
                                                   dynamic_network_loglikelihood_undirected,
                                                   dynamic_network_loglikelihood_directed)
    g = kernels_golden
    X, Yu, Yd = g["a_X"], g["a_Yu"].astype(np.float64), g["a_Yd"].astype(np.float64)
    n = X.shape[1]
    b, b_in, b_out = g["a_b"]
    v = partial_loglikelihood(Yu[1], X[1], b, 5)
    assert abs(v - g["a_k1"][1, 5]) <= 1e-10 * abs(v)
    v = directed_partial_loglikelihood(Yd[2], X[2] / n, g["a_radii"], b_in, b_out, 7)
    assert abs(v - g["a_k2"][2, 7]) <= 1e-10 * abs(v)
    v = dynamic_network_loglikelihood_undirected(Yu, X, b)
    assert abs(v - float(g["a_k5"])) <= 1e-10 * abs(v)
    v = dynamic_network_loglikelihood_directed(Yd, X / n, b_in, b_out, g["a_radii"])
    assert abs(v - float(g["a_k4"])) <= 1e-10 * abs(v)
