"""workloads.splitting_network restates the reference's simple_splitting_dynamic_network
(datasets/samples_generator.py:107-260); with the reference tree present (build container) the two
must agree bit for bit, so that bench.py and the parity tests run on SURVEY 8(d)'s inputs."""
import numpy as np
import pytest

import ref_shims
import workloads as W


@pytest.mark.skipif(not ref_shims.reference_available(), reason="reference sources not present (GPU box)")
@pytest.mark.parametrize("n,steps,directed,seed", [(120, 9, False, 42), (500, 9, False, 42), (60, 6, False, 7),
                                                   (300, 19, True, 42), (90, 7, True, 3)])
def test_splitting_network_equals_the_reference_generator(n, steps, directed, seed):
    ref_shims.load_reference()
    from dynetlsm.datasets import simple_splitting_dynamic_network
    Y, z = simple_splitting_dynamic_network(n_nodes=n, n_time_steps=steps, is_directed=directed,
                                            random_state=seed)
    g = W.splitting_network(n_nodes=n, n_time_steps=steps, is_directed=directed, random_state=seed)
    assert g["Y"].shape == Y.shape and np.array_equal(g["z"], z)
    assert np.array_equal(g["Y"], Y)


def test_configurations_have_the_survey_shapes():
    for name, dens in (("cfg2", (0.25, 0.4)), ("cfg4", (0.15, 0.45))):
        w = W.make_workload(name)
        assert w["Y"].shape == (w["T"], w["n"], w["n"]) and w["X"].shape == (w["T"], w["n"], 2)
        assert np.array_equal(w["Y"], w["Y"].transpose(0, 2, 1)) and set(np.unique(w["Y"])) <= {0.0, 1.0}
        assert dens[0] < w["density"] < dens[1]
        assert w["z"].max() < w["K"] and w["w"].shape == (w["T"], w["K"], w["K"])
        assert np.allclose(w["w"].sum(axis=2), 1.0)
    w = W.make_workload("cfg1")
    assert w["Y"].shape == (3, 18, 18)


def test_bytes_per_node_update_model_m1():
    assert W.bytes_per_node_update(dict(n=120, d=2, directed=False)) == 1935.0
    assert W.bytes_per_node_update(dict(n=500, d=2, directed=False)) == 8062.5
    assert W.bytes_per_node_update(dict(n=2000, d=2, directed=True)) == 48500.0
