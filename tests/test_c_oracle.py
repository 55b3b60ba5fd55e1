"""Pin the C restatement (oracle/oracle.c) to the NumPy oracle, which is itself pinned bit-for-bit to
the reference's golden vectors."""
import numpy as np
import pytest

from conftest import golden_names
from oracle import c_oracle as co, ref_numpy as orc


@pytest.mark.parametrize("name", golden_names("dist"))
def test_distances(golden, name):
    G = golden[f"{name}/G"]
    np.testing.assert_allclose(np.sqrt(co.pairwise_sqdist(G)), orc.pairwise_distances_f64(G), rtol=1e-12, atol=0)
    np.testing.assert_allclose(co.pairwise_dist_as_f32(G), golden[f"{name}/dist"], rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("name", golden_names("krum_idx"))
def test_krum(golden, name):
    G = golden[f"{name}/G"]; f = int(golden[f"{name}/f"])
    assert co.krum(G, len(G), f) == int(golden[f"{name}/krum_idx"])


@pytest.mark.parametrize("name", golden_names("bulyan"))
def test_bulyan(golden, name):
    G = golden[f"{name}/G"]; f = int(golden[f"{name}/f"])
    out, sel = co.bulyan(G, len(G), f)
    assert sel == list(golden[f"{name}/bulyan_sel"])
    np.testing.assert_allclose(out, golden[f"{name}/bulyan"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", golden_names("tm"))
def test_trimmed_mean(golden, name):
    G = golden[f"{name}/G"]; f = int(golden[f"{name}/f"])
    np.testing.assert_allclose(co.trimmed_mean(G, f), golden[f"{name}/tm"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(co.trimmed_mean(G, f), orc.trimmed_mean_f64(G, len(G), f), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("nm", ["a", "b", "c"])
def test_alie(golden, nm):
    rows = golden[f"alie_{nm}/rows"]; z = float(golden[f"alie_{nm}/z"])
    cr, mu, sd = co.alie(rows, z)
    np.testing.assert_allclose(cr, golden[f"alie_{nm}/grads0"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(sd, golden[f"alie_{nm}/stdev"], rtol=1e-5, atol=1e-7)


def test_mean(golden):
    G = golden["het_n10_d64_f2/G"]
    np.testing.assert_allclose(co.mean(G), golden["het_n10_d64_f2/mean"], rtol=1e-6)
    assert co.threads() >= 1
