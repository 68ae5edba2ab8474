"""GPU parity tests, rows d1-d4: Vogiatzis filter, computeTau and the full seed update (epipolar matching included) against
the oracle. Tolerance: seed mean / variance within 1e-4 relative (north_star)."""
import numpy as np
import pytest

from svo_pro_universal_b200 import capi, synth

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4


def test_vogiatzis_update_batch(ctx, orc):
    rng = np.random.default_rng(3)
    n = 20000
    state = np.stack([rng.uniform(0.05, 1.0, n), rng.uniform(1e-5, 0.05, n), rng.uniform(5, 30, n), rng.uniform(5, 30, n)], 1)
    z = state[:, 0] + rng.normal(size=n) * np.sqrt(state[:, 1]) * rng.choice([0.5, 3.0, 30.0], n)
    tau2 = rng.uniform(1e-7, 1e-2, n)
    mu_range = rng.uniform(0.3, 2.0, n)
    z[:50] = -5.0; state[:50, 1] = 100.0            # negative-mean branch
    tau2[50:60] = np.nan                            # NaN norm_scale branch
    exp = state.copy()
    ok_exp = np.zeros(n, np.uint8)
    orc.lib().orc_update_filter_vogiatzis_batch(n, z.ctypes.data_as(orc.f64p), tau2.ctypes.data_as(orc.f64p), mu_range.ctypes.data_as(orc.f64p),
                                                exp.ctypes.data_as(orc.f64p), ok_exp.ctypes.data_as(orc.u8p), 4)
    got = state.copy()
    ok = capi.update_filter_vogiatzis(ctx, z, tau2, mu_range, got)
    assert np.array_equal(ok, ok_exp) and (ok_exp == 0).sum() >= 50
    fin = np.isfinite(exp).all(1)
    np.testing.assert_allclose(got[fin], exp[fin], rtol=1e-9)
    assert np.array_equal(np.isnan(got), np.isnan(exp))


def test_vogiatzis_sequence_of_64_observations(ctx, orc):
    """BASELINE config 4 arithmetic: 64 ordered updates per seed; parity must hold after the whole chain."""
    rng = np.random.default_rng(8)
    S, O = 5000, 64
    true_inv = rng.uniform(0.1, 0.6, S)
    state = np.tile(np.array([1 / 4.0, (1 / 1.5) ** 2 / 36.0, 10.0, 10.0]), (S, 1))
    exp = state.copy()
    got = state.copy()
    mu_range = np.full(S, 1 / 1.5)
    for o in range(O):
        outlier = rng.uniform(size=S) < 0.1
        z = np.where(outlier, rng.uniform(0.01, 0.66, S), true_inv + rng.normal(size=S) * 0.01)
        tau2 = np.full(S, 1e-4 / (o + 1))
        orc.lib().orc_update_filter_vogiatzis_batch(S, z.ctypes.data_as(orc.f64p), tau2.ctypes.data_as(orc.f64p), mu_range.ctypes.data_as(orc.f64p),
                                                    exp.ctypes.data_as(orc.f64p), None, 4)
        capi.update_filter_vogiatzis(ctx, z, tau2, mu_range, got)
    np.testing.assert_allclose(got[:, :2], exp[:, :2], rtol=REL_TOL)
    np.testing.assert_allclose(got, exp, rtol=1e-6)
    assert np.median(np.abs(got[:, 0] - true_inv)) < 5e-3


def test_compute_tau(ctx, orc):
    rng = np.random.default_rng(4)
    n = 1000
    T = np.zeros((n, 7)); T[:, 0] = 1.0
    T[:, 4:] = rng.normal(size=(n, 3)) * 0.1
    f = rng.normal(size=(n, 3)) * 0.2 + np.array([0, 0, 1.0]); f /= np.linalg.norm(f, axis=1)[:, None]
    z = rng.uniform(0.5, 10, n)
    ang = 0.00218
    got = capi.compute_tau(ctx, T, f, z, ang)
    exp = np.array([orc.lib().orc_compute_tau(T[i].ctypes.data_as(orc.f64p), np.ascontiguousarray(f[i]).ctypes.data_as(orc.f64p), float(z[i]), ang)
                    for i in range(n)])
    np.testing.assert_allclose(got, exp, rtol=1e-9)


@pytest.mark.parametrize("dkw,mkw", [(dict(), dict()), (dict(check_convergence=1, seed_convergence_sigma2_thresh=50.0), dict()),
                                      (dict(use_vogiatzis_update=0), dict(scan_on_unit_sphere=0)), (dict(check_visibility=0), dict())])
def test_update_seeds_full_chain(ctx, orc, dkw, mkw):
    """depth_filter_utils::updateSeed over ordered observations: matcher + tau + filter + convergence flags."""
    sq = synth.make_seed_sequence(17, n_seeds=600, n_obs=10)
    S, O = len(sq["px"]), len(sq["cur_imgs"])
    ref = capi.Pyramid(ctx, 1, 752, 480, 5)
    cur = capi.Pyramid(ctx, O, 752, 480, 5)
    ref.upload(sq["ref_img"]); cur.upload(np.stack(sq["cur_imgs"]))
    ref.build(); cur.build()
    cam = capi.Camera.from_dict(sq["cam"])
    ft = capi.make_features(sq["px"], sq["f"], sq["grad"], sq["type"].astype(np.int32), sq["level"])
    types = sq["type"].copy()
    types[:5] = synth.K_OUTLIER                         # already-diverged seeds are skipped
    types[5:10] = synth.K_CORNER_SEED_CONV              # converged seeds: skipped only with check_convergence
    state = sq["state"].copy()
    mu_range = np.full(S, sq["mu_range"])
    obs_frame = np.tile(np.arange(O, dtype=np.int32)[:, None], (1, S))
    obs_frame[3, ::7] = -1                              # e.g. cur frame == ref frame for some seeds: skipped
    obs_T = np.ascontiguousarray(obs_frame.clip(min=0))
    mopt = capi.matcher_options(**mkw)
    dopt = capi.depth_filter_options(**dkw)
    g_types, g_state = types.copy(), state.copy()
    n_succ, mr = capi.update_seeds(ctx, ref, cur, cam, cam, ft, g_types, g_state, mu_range, obs_frame, obs_T,
                                   np.ascontiguousarray(sq["T_cur_ref"]), mopt, dopt)
    # oracle: one observation at a time so skipped (seed, observation) pairs can be honoured
    keep = []
    rf = orc.make_frame(orc.create_img_pyramid(sq["ref_img"], 5), sq["cam"], keep=keep)
    cfs = [orc.make_frame(orc.create_img_pyramid(im, 5), sq["cam"], keep=keep) for im in sq["cur_imgs"]]
    oft = orc.make_features(sq["px"], sq["f"], sq["grad"], sq["type"].astype(np.int32), sq["level"])
    o_types, o_state = types.copy(), state.copy()
    oopt = orc.default_matcher_options(**mkw)
    n_exp = 0
    mr_exp = np.full((O, S), -1, np.int32)
    for o in range(O):
        act = np.flatnonzero(obs_frame[o] >= 0)
        sub_f = (orc.Feature * len(act))(*[oft[i] for i in act])
        t_sub, s_sub = np.ascontiguousarray(o_types[act]), np.ascontiguousarray(o_state[act])
        n, m, _ = orc.update_seeds(rf, [cfs[o]], sq["T_cur_ref"][o:o + 1], sub_f, t_sub, s_sub, sq["mu_range"], oopt,
                                   sigma2_thresh=dopt.seed_convergence_sigma2_thresh, check_visibility=dopt.check_visibility,
                                   check_convergence=dopt.check_convergence, use_vogiatzis=dopt.use_vogiatzis_update, n_threads=8)
        o_types[act], o_state[act] = t_sub, s_sub
        mr_exp[o, act] = m[0]
        n_exp += n
    assert np.array_equal(mr, mr_exp), np.argwhere(mr != mr_exp)[:10]
    assert np.array_equal(g_types, o_types)
    assert int(n_succ[0]) == n_exp and n_exp > 0.3 * S * O
    np.testing.assert_allclose(g_state[:, :2], o_state[:, :2], rtol=REL_TOL)
    np.testing.assert_allclose(g_state[:, 2:], o_state[:, 2:], rtol=1e-4)
    assert (g_state[:5] == state[:5]).all()
    good = (g_types == synth.K_CORNER_SEED_CONV) | (g_types == synth.K_EDGELET_SEED_CONV)
    if dopt.use_vogiatzis_update and good.sum() > 20:
        rel = np.abs(1.0 / g_state[good, 0] - sq["depth_true"][good]) / sq["depth_true"][good]
        assert np.median(rel) < 0.05  # converged seeds sit at the true depth
