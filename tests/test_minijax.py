"""The AD engine under the reference executor (oracle/refexec/minijax.py) against closed forms,
central finite differences and mpmath high-precision evaluation: it supplies the derivatives when
the reference's own source is executed, so it must be trustworthy on its own."""
import numpy as np
import pytest

from oracle.refexec import minijax as jnp


def _fd(f, x, h=1e-6):
    x = np.asarray(x, dtype=np.float64)
    f0 = np.asarray(f(x))
    J = np.zeros(f0.shape + x.shape)
    for idx in np.ndindex(*x.shape):
        e = np.zeros_like(x); e[idx] = h
        J[(Ellipsis,) + idx] = (np.asarray(f(x + e)) - np.asarray(f(x - e))) / (2 * h)
    return J


def test_elementwise_and_broadcast_rules():
    x = np.array([0.3, -1.2, 2.0])

    def f(v):
        return jnp.sin(v) * jnp.cos(2.0 * v) + jnp.sqrt(v * v + 1.0) / (1.0 + jnp.abs(v)) - v ** 3
    J = np.asarray(jnp.jacfwd(f)(x))
    c, s = np.cos, np.sin
    r = np.sqrt(x * x + 1)
    d = (c(x) * c(2 * x) - 2 * s(x) * s(2 * x) + (x / r) / (1 + np.abs(x))
         - r * np.sign(x) / (1 + np.abs(x)) ** 2 - 3 * x ** 2)
    assert np.allclose(J, np.diag(d), rtol=1e-14, atol=1e-15)


def test_abs_derivative_is_zero_at_zero_like_jax():
    J = np.asarray(jnp.jacfwd(lambda v: jnp.abs(v) * v)(np.array([0.0, 2.0, -3.0])))
    assert np.array_equal(np.diag(J), [0.0, 4.0, 6.0])


def test_matmul_norm_at_set_reshape():
    rs = np.random.RandomState(0)
    A, B = rs.randn(3, 4), rs.randn(4, 2)
    x = rs.randn(4)

    def f(v):
        z = jnp.zeros((2, 3))
        z = z.at[0, :].set(A @ v)
        z = z.at[1, :2].set(v @ B)
        z = z.at[1, 2].set(jnp.linalg.norm(v[1:3]))
        return jnp.reshape(z, (6,), 'F')
    J = np.asarray(jnp.jacfwd(f)(x))
    assert np.allclose(J, _fd(lambda v: np.asarray(f(v)), x), rtol=1e-7, atol=1e-8)
    n = np.linalg.norm(x[1:3])
    assert np.allclose(J[5], [0, x[1] / n, x[2] / n, 0], rtol=1e-14)          # 'F' order: (1,2) is last


def test_vmap_stacks_pytrees_and_nested_jacfwd_of_vmapped_function():
    rs = np.random.RandomState(1)
    W = rs.randn(5, 3)

    def per(w, v):
        return (jnp.sin(w @ v), jnp.array([w[0] * v[1], v[2] ** 2]))
    x = rs.randn(3)

    def f(v):
        a, b = jnp.vmap(per)(W, jnp.repeat(v[jnp.newaxis, :], 5, axis=0))
        return jnp.concatenate([a, jnp.mean(b, axis=0)])
    J = np.asarray(jnp.jacfwd(f)(x))
    assert J.shape == (7, 3)
    assert np.allclose(J, _fd(lambda v: np.asarray(f(v)), x), rtol=1e-7, atol=1e-8)


def test_hessian_by_complex_step_matches_closed_form():
    rs = np.random.RandomState(2)
    I, th, ta = rs.rand(6), rs.rand(6) * 3, rs.rand(6) * 6
    lam = rs.randn(2)

    def L(z):        # lam0 * (z1 - mu(z0) z2) + lam1 * z0 z2^2  with mu = sum I cos(th z0 + ta)
        mu = jnp.sum(I * jnp.cos(th * z[0] + ta))
        return lam[0] * (z[1] - mu * z[2]) + lam[1] * z[0] * z[2] * z[2]
    z = rs.randn(3)
    H = np.asarray(jnp.hessian(L)(z))
    mu1 = -np.sum(I * th * np.sin(th * z[0] + ta))
    mu2 = -np.sum(I * th * th * np.cos(th * z[0] + ta))
    Hc = np.zeros((3, 3))
    Hc[0, 0] = -lam[0] * mu2 * z[2]
    Hc[0, 2] = Hc[2, 0] = -lam[0] * mu1 + 2 * lam[1] * z[2]
    Hc[2, 2] = 2 * lam[1] * z[0]
    assert np.allclose(H, Hc, rtol=1e-13, atol=1e-14)


def test_dual_numbers_against_mpmath_50_digits():
    """One drone-like Euler step chain (|v| v drag, division by the mass) differentiated by the
    dual numbers vs mpmath's arbitrary-precision numerical derivative."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    dt, m, kp, kd, c = 2.5, 31.7, 0.05, 0.25, 0.2

    def chain(u, absf, lib):
        p, v = lib(-1.9), lib(0.3)
        for k in range(6):
            a = (u - kp * p - kd * v - c * absf(v) * v) / m
            p, v = p + dt * v, v + dt * a
        return p * p + v
    d_ad = float(np.asarray(jnp.jacfwd(lambda u: chain(u[0], jnp.abs, float))(np.array([0.7])))[0])
    d_mp = mp.diff(lambda u: chain(u, mp.fabs, mp.mpf), mp.mpf('0.7'))
    assert abs(d_ad - float(d_mp)) <= 1e-14 * abs(float(d_mp))
