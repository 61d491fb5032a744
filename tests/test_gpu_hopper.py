"""Parity of the hopper slip-risk block (values, Jacobian, Hessian) with the oracles."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _feats():
    from riskaversetrajopt_b200.hopper import hopper as hp
    st = np.random.get_state(); np.random.seed(1)
    f = hp.sample_friction_features(hp.M); np.random.set_state(st)
    return f


def _dense(shape, r, c, v):
    D = np.zeros(shape)
    np.add.at(D, (r, c), v)
    return D


@pytest.mark.parametrize("src", ["", "ref_"], ids=["oracle_a", "reference_exec"])
@pytest.mark.parametrize("method", ["saa", "baseline"])
def test_golden_values_jacobian_hessian(method, src):
    from riskaversetrajopt_b200.hopper import hopper as hp
    g = np.load(os.path.join(G, src + "hopper_M30.npz"))
    Z = g["Z"]
    m = hp.Model(hp.M, method, 0.2, _feats())
    nv = hp.num_vars(hp.M)
    gs = m.slip_risk_constraints(Z)
    assert gs.shape == g[method + "_g"].shape == (m.n_rows,)
    assert np.allclose(gs, g[method + "_g"], rtol=1e-9, atol=1e-12)
    r, c, v = m.slip_risk_jacobian(Z)
    J = _dense((m.n_rows, nv), r, c, v)
    Jg = _dense((m.n_rows, nv), g[method + "_jac_r"], g[method + "_jac_c"], g[method + "_jac_v"])
    assert np.allclose(J, Jg, rtol=1e-9, atol=1e-12)
    # the structural pattern covers every non-zero the reference's dense jacrev has
    assert set(zip(g[method + "_jac_r"], g[method + "_jac_c"])) <= set(zip(r, c))
    hr, hc, hv = m.slip_risk_hessian(Z, g[method + "_lam"])
    assert np.all(hr >= hc)
    H = _dense((nv, nv), hr, hc, hv)
    Hg = _dense((nv, nv), g[method + "_hess_r"], g[method + "_hess_c"], g[method + "_hess_v"])
    assert np.allclose(H, Hg, rtol=1e-9, atol=1e-11)


def test_callback_style_helpers_match_dense_oracle():
    from oracle.oracle_hopper import HopperOracleA
    from riskaversetrajopt_b200.hopper import hopper as hp
    M = 7
    rs = np.random.RandomState(3)
    f = tuple(x[:M] for x in _feats())
    nv = hp.num_vars(M)
    Z = rs.uniform(-1, 1, nv); Z[3::8][:31] = rs.uniform(0.5, 1.2, 31)
    m, a = hp.Model(M, 'saa', 0.3, f), HopperOracleA(M, 'saa', 0.3, *f)
    off = 5
    out = np.full(off + m.n_rows + 3, 7.0)
    m.eval_g_into(Z, out, off)
    assert np.allclose(out[off:off + m.n_rows], a.g(Z), rtol=1e-9, atol=1e-12) and out[0] == 7.0 and out[-1] == 7.0
    Jd = np.full((off + m.n_rows, nv), 1.0)
    m.eval_jac_g_into(Z, Jd, off)
    assert np.allclose(Jd[off:], a.jac(Z), rtol=1e-9, atol=1e-12) and np.all(Jd[:off] == 1.0)
    lam = rs.randn(m.n_rows)
    Hd = np.zeros((nv, nv))
    m.eval_h_add(Z, lam, Hd)
    assert np.allclose(Hd, a.hess(Z, lam), rtol=1e-9, atol=1e-11)


def test_large_M_and_fp32():
    from oracle.oracle_hopper import HopperOracleB
    from riskaversetrajopt_b200.hopper import hopper as hp
    M = 50_000
    rs = np.random.RandomState(5)
    f = (0.025 * np.sqrt(2 / 30) * rs.uniform(0, 1, (M, 30)), rs.uniform(0, np.pi, (M, 30)),
         rs.uniform(0, 2 * np.pi, (M, 30)))
    Z = rs.uniform(-1, 1, hp.num_vars(M))
    m = hp.Model(M, 'saa', 0.1, f)
    b = HopperOracleB(M, 'saa', 0.1, *f)
    assert np.allclose(m.slip_risk_constraints(Z), b.g(Z), rtol=1e-9, atol=1e-12)
    # Hessian sums: sum_i lam mu', sum_i lam mu'' against the closed form
    px, fx, fz, x2, x3 = m._contact_geometry(Z)
    lam = rs.randn(M, 20)
    _, dmu, hs = m._friction(px, lam.reshape(-1))
    mu_b, dmu_b, d2mu_b = b.friction(px)
    assert np.allclose(dmu, dmu_b, rtol=1e-9, atol=1e-13)
    assert np.allclose(hs[:, 0], (lam * dmu_b).sum(0), rtol=1e-9, atol=1e-10)
    assert np.allclose(hs[:, 1], (lam * d2mu_b).sum(0), rtol=1e-9, atol=1e-10)
    m32 = hp.Model(M, 'saa', 0.1, f, precision='fp32')
    g32 = m32.slip_risk_constraints(Z)
    assert np.allclose(g32, b.g(Z), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("pscale,atol", [(1.0, 1e-13), (2.0e4, 2e-10), (1.0e7, 2e-7)])
def test_friction_field_argument_ranges(pscale, atol):
    """The kernel's branch-free sincos covers |theta p + tau| < 1e5; a chunk of samples whose
    arguments may exceed that takes the library routine.  pscale = 2e4 puts the arguments on both
    sides of the switch (theta in [0, pi)), 1e7 far beyond it.  The tolerance follows the argument
    rounding: the oracle rounds theta*p and the sum separately, the kernel uses one FMA, so the
    arguments differ by up to ulp(|x|) and mu by up to sum_f I_f ulp(|x|)."""
    from oracle.oracle_hopper import HopperOracleB
    from riskaversetrajopt_b200.hopper import hopper as hp
    M = 4099                                    # ragged: not a multiple of the staged chunk
    rs = np.random.RandomState(11)
    f = (0.025 * np.sqrt(2 / 30) * rs.uniform(0, 1, (M, 30)), rs.uniform(0, np.pi, (M, 30)),
         rs.uniform(0, 2 * np.pi, (M, 30)))
    f[1][::7] *= 1e-3                           # some samples stay in the fast range at pscale = 2e4
    m = hp.Model(M, 'saa', 0.1, f)
    b = HopperOracleB(M, 'saa', 0.1, *f)
    px = pscale * rs.uniform(-1, 1, 20)
    lam = rs.randn(M, 20)
    mu, dmu, hs = m._friction(px, lam.reshape(-1))
    mu_b, dmu_b, d2mu_b = b.friction(px)
    assert np.allclose(mu, mu_b, rtol=1e-9, atol=atol)
    assert np.allclose(dmu, dmu_b, rtol=1e-9, atol=atol * np.pi)
    assert np.allclose(hs[:, 0], (lam * dmu_b).sum(0), rtol=1e-9, atol=atol * np.pi * M)
    assert np.allclose(hs[:, 1], (lam * d2mu_b).sum(0), rtol=1e-9, atol=atol * np.pi ** 2 * M)
    # determinism: the Hessian sums are reduced in a fixed order
    _, _, hs2 = m._friction(px, lam.reshape(-1))
    assert np.array_equal(hs, hs2)


def test_fp32_storage_mode_per_entry(built_lib=None):
    """precision='fp32' = FP32 STORAGE of g / the Jacobian entries, FP64 arithmetic: every entry is
    the FP64 entry rounded once, so the 1e-4 relative tolerance holds per entry even where mu' is
    small by cancellation between the 30 features."""
    from oracle.oracle_hopper import HopperOracleB
    from riskaversetrajopt_b200.hopper import hopper as hp
    M = 4099
    rs = np.random.RandomState(12)
    f = (0.025 * np.sqrt(2 / 30) * rs.uniform(0, 1, (M, 30)), rs.uniform(0, np.pi, (M, 30)),
         rs.uniform(0, 2 * np.pi, (M, 30)))
    Z = rs.uniform(-1, 1, hp.num_vars(M))
    m64, m32 = hp.Model(M, 'saa', 0.1, f), hp.Model(M, 'saa', 0.1, f, precision='fp32')
    g64, g32 = m64.slip_risk_constraints(Z), m32.slip_risk_constraints(Z)
    nz = g64 != 0
    assert np.max(np.abs(g32[nz] - g64[nz]) / np.abs(g64[nz])) < 1e-6
    (_, _, v64), (_, _, v32) = m64.slip_risk_jacobian(Z), m32.slip_risk_jacobian(Z)
    nz = v64 != 0
    assert np.max(np.abs(v32[nz] - v64[nz]) / np.abs(v64[nz])) < 1e-6
    b = HopperOracleB(M, 'saa', 0.1, *f)
    assert np.allclose(g64, b.g(Z), rtol=1e-9, atol=1e-12)


def test_monte_carlo_terms_vs_reference_execution():
    """no_slip_constraints_verification vmapped over the samples (hopper/hopper.py:910-925) and the
    AV@R closed form (:957): fixture = the reference's own functions executed (ref_hopper_M30.npz)."""
    from riskaversetrajopt_b200.hopper import hopper as hp
    g = np.load(os.path.join(G, "ref_hopper_M30.npz"))
    m = hp.Model(hp.M, 'saa', 0.2, _feats())
    sat, Zi, out3 = m.monte_carlo_constraints(g["Z"], t_risk=0.05)
    assert np.allclose(Zi, g["mc_Z"], rtol=1e-9, atol=1e-12)
    assert np.array_equal(sat, g["mc_Z"] <= 1e-6)
    assert np.isclose(out3[0], np.maximum(g["mc_Z"] - 0.05, 0).sum(), rtol=1e-10)
    assert out3[1] == np.count_nonzero(g["mc_Z"] <= 1e-6) and np.isclose(out3[2], g["mc_Z"].max(), rtol=1e-12)
    avar = m.monte_carlo_avar(g["Z"], 0.05, alpha=0.2)
    assert np.isclose(avar, 0.05 + np.mean(np.maximum(g["mc_Z"] - 0.05, 0)) / 0.2, rtol=1e-10)


@pytest.mark.parametrize("M", [1, 7, 33, 4099])
def test_device_assembly_ragged_sizes_and_baseline(M):
    """saa_hopper_g / _jac / _hess write the block straight into caller buffers: ragged M, both methods."""
    from oracle.oracle_hopper import HopperOracleB
    from riskaversetrajopt_b200.hopper import hopper as hp
    rs = np.random.RandomState(M)
    f = (0.025 * np.sqrt(2 / 30) * rs.uniform(0, 1, (M, 30)), rs.uniform(0, np.pi, (M, 30)),
         rs.uniform(0, 2 * np.pi, (M, 30)))
    Z = rs.uniform(-1, 1, hp.num_vars(M))
    for method in ('saa', 'baseline'):
        m, b = hp.Model(M, method, 0.3, f), HopperOracleB(M, method, 0.3, *f)
        gd = m.slip_risk_constraints_device(Z)
        assert gd.is_cuda and gd.shape == (m.n_rows,)
        assert np.allclose(gd.cpu().numpy(), b.g(Z), rtol=1e-9, atol=1e-12)
        if method == 'saa':
            assert gd[-1].item() == 0.0 and np.isclose(gd[0].item(), b.g(Z)[0], rtol=1e-12)
