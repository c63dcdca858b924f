"""Pin the CPU oracle against the reference's own behaviour (CPU only).

The trajectory fixtures hold what the UNMODIFIED reference did when fed the oracle's
counter-based draws (tests/golden/make_golden.py); the oracle must reproduce every integer
decision exactly and every float to 1e-10.
"""
import numpy as np
import pytest

from oracle import oracle as orc

from _helpers import load, oracle_from_fixture

TRAJ = ["traj_t1_d5", "traj_t4_groups_d6", "traj_t3_curved_ext_d4", "traj_t1_d20", "traj_t3_hot_tmax_d4",
        "traj_t2_prior_d4"]
FTOL = 1e-10


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10 (Salmon et al. SC'11)
    assert orc.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert orc.philox4x32_10([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert orc.philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_draw_primitives():
    words = [orc.draw_word(11, orc.PURPOSE_MH, 5, 3, 2, j) for j in range(4000)]
    assert len(set(words)) == len(words)
    u = np.array([orc.word_to_unit(w) for w in words])
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.02
    ints = np.array([orc.word_to_int(w, 7) for w in words])
    assert ints.min() == 0 and ints.max() == 6
    z = np.array([orc.word_to_normals(w) for w in words]).ravel()
    assert abs(z.mean()) < 0.05 and abs(z.var() - 1.0) < 0.06
    # extreme words stay finite; the pair is a rotation of one radius
    for w in (0, 2**64 - 1, 2**63, 2**32 - 1):
        a, b = orc.word_to_normals(w)
        assert np.isfinite(a) and np.isfinite(b)
    # half-turn symmetry of the angle: lo -> lo + 2^31 negates both normals
    a, b = orc.word_to_normals((123456 << 32) | 98765)
    c, d = orc.word_to_normals((123456 << 32) | (98765 + 2**31))
    assert a == -c and b == -d


def test_normals_distribution_ks_and_tails():
    """10^7 Box-Muller pairs of the engine's single-precision generator (word_to_normals, the same bits on the
    device: tests/test_gpu_parity.py::test_device_normals_equal_oracle_bits) against N(0, 1): Kolmogorov-Smirnov
    distance, tail masses out to 5 sigma within Poisson error, moments, independence of the pair."""
    from scipy import special

    n = 10_000_000
    words = np.random.default_rng(2024).integers(0, 2**64, n, dtype=np.uint64)
    z0, z1 = orc.word_to_normals_many(words)
    for z in (z0, z1):
        zs = np.sort(z)
        cdf = 0.5 * special.erfc(-zs / np.sqrt(2.0))
        i = np.arange(1, n + 1)
        ks = max(np.max(i / n - cdf), np.max(cdf - (i - 1) / n))
        assert ks < 1.63 / np.sqrt(n)  # 1 % critical value of the KS statistic
        assert abs(z.mean()) < 5 / np.sqrt(n) and abs(z.var() - 1.0) < 5 * np.sqrt(2.0 / n)
        assert abs(np.mean(z**4) - 3.0) < 5 * np.sqrt(96.0 / n)
        for k in (2.0, 3.0, 4.0, 5.0):
            expect = n * special.erfc(k / np.sqrt(2.0))  # two-sided tail mass
            got = np.count_nonzero(np.abs(z) > k)
            assert abs(got - expect) < 5 * np.sqrt(expect) + 1, (k, got, expect)
        assert np.abs(z).max() < 6.77  # the radius comes from 33 bits: |z| <= sqrt(66 ln 2)
    assert abs(np.mean(z0 * z1)) < 5 / np.sqrt(n)
    assert abs(np.mean(z0 * z0 * z1 * z1) - 1.0) < 5 * np.sqrt(8.0 / n)
    # the far tail keeps its resolution: radius from the smallest words
    tiny = (np.arange(1, 2000, dtype=np.uint64) << np.uint64(32))
    r = np.hypot(*orc.word_to_normals_many(tiny))
    expect = np.sqrt(-2.0 * np.log((2.0 * np.arange(1, 2000) + 1.0) * 2.0**-33))
    assert np.allclose(r, expect, rtol=3e-7)


def test_temperature_ladder_matches_reference_formula():
    # ref PTMCMCSampler.py:709-718
    lad = orc.temperature_ladder(20, 5)
    assert np.allclose(lad, (1 + np.sqrt(2 / 20.0)) ** np.arange(5), rtol=1e-14)
    lad = orc.temperature_ladder(20, 4, 1.0, 50.0)
    assert np.allclose(lad, np.exp(np.log(50.0) / 3) ** np.arange(4), rtol=1e-13)
    assert np.array_equal(orc.temperature_ladder(20, 1), [1.0])


def test_sym_factor_against_lapack():
    rng = np.random.default_rng(0)
    for n in (1, 2, 5, 20, 48):
        A = rng.standard_normal((n, n))
        cov = A @ A.T / n + 1e-3 * np.eye(n)
        U, S = orc.sym_factor(cov)
        Ur, Sr, _ = np.linalg.svd(cov)
        assert np.allclose(S, Sr, rtol=1e-11, atol=1e-14)
        assert np.allclose(U @ np.diag(S) @ U.T, cov, atol=1e-12)
        assert np.allclose(U.T @ U, np.eye(n), atol=1e-12)
        assert np.allclose(np.abs(U), np.abs(Ur), atol=1e-8)
        assert np.all(U[np.abs(U).argmax(axis=0), np.arange(n)] > 0)
    # diagonal input: exact, sorted, ties keep index order
    U, S = orc.sym_factor(np.diag([0.01, 0.04, 0.04, 0.02]))
    assert np.array_equal(S, [0.04, 0.04, 0.02, 0.01])
    assert np.array_equal(U.argmax(axis=0), [1, 2, 3, 0])


@pytest.mark.parametrize("name", TRAJ)
def test_trajectory_matches_reference(name):
    g = load(name)
    N, T = int(g["N"]), int(g["T"])
    o = oracle_from_fixture(g)
    tskip = int(g["kw_Tskip"])
    trace, swapmaps = o.set_trace(N, N // tskip)
    o.set_state(g["p0"][:, None, :])
    o.run(N)
    ch, lnl, lnp = o.chain()
    # integer bookkeeping: bit exact
    assert np.array_equal(trace[:, :, 0] & 0x7F, g["jump"])
    assert np.array_equal(trace[:, :, 0] >> 7, g["acc"])
    prop, acc, sw, nsw = o.counters()
    assert np.array_equal(prop[:, 0, :g["jump_prop"].shape[1]], g["jump_prop"])
    assert np.array_equal(acc[:, 0, :g["jump_acc"].shape[1]], g["jump_acc"])
    assert np.array_equal(acc[:, 0, :].sum(axis=1), g["naccepted"])
    assert np.array_equal(sw[:, 0], g["swap_acc"][-1])
    assert nsw == int(g["swap_proposed"])
    # floating point: 1e-10
    assert np.allclose(ch[1:, :, 0], g["x"], rtol=FTOL, atol=FTOL)
    assert np.allclose(lnl[1:, :, 0], g["lnl"], rtol=FTOL, atol=FTOL)
    assert np.allclose(lnp[1:, :, 0], g["lnp"], rtol=FTOL, atol=FTOL)
    cov, mu, m2, n = o.adapt()
    assert np.allclose(cov, g["cov"], rtol=1e-9, atol=1e-12)
    assert np.allclose(mu, g["mu"], rtol=1e-9, atol=1e-12)
    assert np.allclose(m2, g["m2"], rtol=1e-9, atol=1e-9)
    am, de = o.buffers()
    assert np.allclose(am[:, 0], g["am"], rtol=FTOL, atol=FTOL)
    assert np.allclose(de[:, 0], g["de"], rtol=FTOL, atol=FTOL)


@pytest.mark.parametrize("name", TRAJ)
def test_thinned_record_matches_reference_chain(name):
    """The reference's _chain/_lnlike/_lnprob rows (ref :331-335) for the T=1 chain."""
    g = load(name)
    N, thin = int(g["N"]), int(g["kw_thin"])
    o = oracle_from_fixture(g, thin=thin, record_hot=False)
    o.set_state(g["p0"][:, None, :])
    o.run(N)
    ch, lnl, lnp = o.chain()
    rows = N // thin + 1
    assert ch.shape[0] == rows
    assert np.allclose(ch[:, 0, 0], g["chain"][:rows], rtol=FTOL, atol=FTOL)
    assert np.allclose(lnl[:, 0, 0], g["chain_lnl"][:rows], rtol=FTOL, atol=FTOL)
    assert np.allclose(lnp[:, 0, 0], g["chain_lnp"][:rows], rtol=FTOL, atol=FTOL)


@pytest.mark.parametrize("name", ["traj_t1_d5", "traj_t1_d20"])
def test_own_factorisation_tracks_reference_covariance(name):
    """Without injected factors the trajectory forks at the first covariance update (LAPACK's
    eigenvector signs are arbitrary) but everything before it is identical, and the oracle's own
    factor of the reference covariance agrees with LAPACK up to those signs."""
    g = load(name)
    cu = int(g["kw_covUpdate"])
    o = oracle_from_fixture(g, inject=False)
    trace, _ = o.set_trace(cu)
    o.set_state(g["p0"][:, None, :])
    o.run(cu)
    assert np.array_equal(trace[:, :, 0] & 0x7F, g["jump"][:cu])
    assert np.array_equal(trace[:, :, 0] >> 7, g["acc"][:cu])
    o.run(1)  # triggers the first update
    cov, mu, m2, n = o.adapt()
    U, S = o.factor()
    d = int(g["d"])
    Ur, Sr, _ = np.linalg.svd(cov)
    assert np.allclose(S, Sr, rtol=1e-10)
    assert np.allclose(np.abs(U.reshape(d, d)), np.abs(Ur), atol=1e-7)
    assert np.allclose(S, g["S"][0], rtol=1e-9)


def test_walker_axis_is_w_independent_copies_before_adaptation():
    """W walkers = W runs of the single-walker algorithm with walker-keyed streams (until the
    pooled covariance first couples them)."""
    g = load("traj_t4_groups_d6")
    cu, T, d = int(g["kw_covUpdate"]), int(g["T"]), int(g["d"])
    W = 3
    rng = np.random.default_rng(0)
    x0 = rng.uniform(0, 10, (T, W, d))
    from _helpers import fixture_cycle, fixture_groups, fixture_target
    lk, lpar, pk, ppar = fixture_target(g)
    common = dict(seed=5, ladder=g["ladder"], groups=fixture_groups(g), cycle=fixture_cycle(g),
                  de_weight=50, cov_update=cu, burn=200, tskip=7, thin=1, logl_kind=lk, logl_params=lpar,
                  logp_kind=pk, logp_params=ppar, record_hot=True, max_rows=cu + 1)
    pooled = orc.Oracle(d, W, T, g["cov0"], **common)
    pooled.set_state(x0)
    pooled.run(cu)
    xs, lnls, _, _ = pooled.state()
    for w in range(W):
        single = orc.Oracle(d, 1, T, g["cov0"], walker_offset=w, **common)
        single.set_state(x0[:, w:w + 1])
        single.run(cu)
        x1, l1, _, _ = single.state()
        assert np.array_equal(x1[:, 0], xs[:, w])
        assert np.array_equal(l1[:, 0], lnls[:, w])


def test_threads_do_not_change_results():
    g = load("traj_t4_groups_d6")
    from _helpers import fixture_cycle, fixture_groups, fixture_target
    lk, lpar, pk, ppar = fixture_target(g)
    T, d, W = int(g["T"]), int(g["d"]), 16
    x0 = np.random.default_rng(1).uniform(0, 10, (T, W, d))
    res = []
    for nth in (1, 4):
        o = orc.Oracle(d, W, T, g["cov0"], seed=3, ladder=g["ladder"], groups=fixture_groups(g),
                       cycle=fixture_cycle(g), de_weight=50, cov_update=50, burn=100, tskip=7, thin=5,
                       logl_kind=lk, logl_params=lpar, logp_kind=pk, logp_params=ppar, max_rows=100,
                       nthreads=nth)
        o.set_state(x0)
        o.run(330)
        res.append((o.state(), o.adapt(), o.counters(), o.chain()))
    a, b = res
    for k in range(4):
        assert np.array_equal(a[0][k], b[0][k])
    assert np.array_equal(a[1][0], b[1][0]) and np.array_equal(a[1][2], b[1][2])
    assert all(np.array_equal(x, y) for x, y in zip(a[2][:3], b[2][:3]))
    assert np.array_equal(a[3][0], b[3][0])


def test_de_update_with_covupdate_larger_than_burn_is_an_error():
    # the reference raises a broadcast ValueError at ref :817 in this configuration
    o = orc.Oracle(3, 1, 1, np.eye(3), cov_update=50, burn=20, thin=1, max_rows=100,
                   logl_params=orc.gaussian_params(np.zeros(3), np.eye(3)),
                   logp_params=orc.uniform_params(-5 * np.ones(3), 5 * np.ones(3)))
    o.set_state(np.zeros((1, 1, 3)))
    with pytest.raises(ValueError):
        o.run(40)
