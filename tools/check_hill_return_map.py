#!/usr/bin/env python
"""numpy check of the closed-form structure behind `hill_return_map` / `hill_return_map_plane_stress`
(calibr8_b200/csrc/models.cuh): with the associated Hill flow n = M s / hill(s), the flow rule and the yield
condition of the small_hill residual (src/small_hill.cpp:196-276, src/small_hill_plane_stress.cpp:194-276) reduce
to (I + c M) s = s_trial and ONE scalar equation in dgam.  For random trial states the scalar Newton below (the
same steps as the device code) must land on a state that satisfies the model's own residual equations to
rounding.  CPU only, no library needed:  python tools/check_hill_return_map.py"""
import numpy as np


def hill_params(R00, R11, R22, R01, R02, R12):
    i00, i11, i22 = 1 / R00 ** 2, 1 / R11 ** 2, 1 / R22 ** 2
    return dict(F=.5 * (i11 + i22 - i00), G=.5 * (i22 + i00 - i11), H=.5 * (i00 + i11 - i22),
                L=1.5 / R12 ** 2, M=1.5 / R02 ** 2, N=1.5 / R01 ** 2)


def hill_value(t, h):
    return np.sqrt(h['F'] * (t[1, 1] - t[2, 2]) ** 2 + h['G'] * (t[2, 2] - t[0, 0]) ** 2 + h['H'] * (t[0, 0] - t[1, 1]) ** 2
                   + 2 * (h['L'] * t[1, 2] ** 2 + h['M'] * t[0, 2] ** 2 + h['N'] * t[0, 1] ** 2))


def hill_normal(t, h, hv):
    n = np.zeros((3, 3))
    n[0, 0] = ((h['G'] + h['H']) * t[0, 0] - h['H'] * t[1, 1] - h['G'] * t[2, 2]) / hv
    n[1, 1] = ((h['F'] + h['H']) * t[1, 1] - h['H'] * t[0, 0] - h['F'] * t[2, 2]) / hv
    n[2, 2] = ((h['G'] + h['F']) * t[2, 2] - h['G'] * t[0, 0] - h['F'] * t[1, 1]) / hv
    n[0, 1] = n[1, 0] = h['N'] * t[0, 1] / hv
    n[0, 2] = n[2, 0] = h['M'] * t[0, 2] / hv
    n[1, 2] = n[2, 1] = h['L'] * t[1, 2] / hv
    return n


def sigma_y(a, Y, S, D):
    ex = np.exp(-D * a)
    return Y + S * (1 - ex), S * D * ex


def return_map_3d(s_tr, h, mu, Y, S, D, a0):
    Mn = np.array([[h['G'] + h['H'], -h['H'], -h['G']], [-h['H'], h['F'] + h['H'], -h['F']],
                   [-h['G'], -h['F'], h['G'] + h['F']]])
    sy0, dsy = sigma_y(a0, Y, S, D)
    dgam = max((hill_value(s_tr, h) - sy0) / (3 * mu + dsy), 0.)
    polish = False
    for it in range(40):
        sy, dsy = sigma_y(a0 + dgam, Y, S, D)
        c = 2 * mu * dgam / sy
        Ai = np.linalg.inv(np.eye(3) + c * Mn)
        i01, i02, i12 = 1 / (1 + c * h['N']), 1 / (1 + c * h['M']), 1 / (1 + c * h['L'])
        sn = Ai @ np.array([s_tr[0, 0], s_tr[1, 1], s_tr[2, 2]])
        s = np.zeros((3, 3)); s[0, 0], s[1, 1], s[2, 2] = sn
        s[0, 1] = s[1, 0] = s_tr[0, 1] * i01; s[0, 2] = s[2, 0] = s_tr[0, 2] * i02; s[1, 2] = s[2, 1] = s_tr[1, 2] * i12
        hv = hill_value(s, h); g = hv - sy
        if abs(g) < 1e-13 * mu:
            if polish:
                return s, dgam, it
            polish = True
        ms = Mn @ sn; dsn = -(Ai @ ms)
        dh = (ms @ dsn) / hv + 2 / hv * (h['N'] * s[0, 1] * (-h['N'] * s[0, 1] * i01) + h['M'] * s[0, 2] * (-h['M'] * s[0, 2] * i02)
                                         + h['L'] * s[1, 2] * (-h['L'] * s[1, 2] * i12))
        dg = dh * (2 * mu * (sy - dgam * dsy) / sy ** 2) - dsy
        nd = dgam - g / dg
        dgam = nd if nd > 0 else .5 * dgam
    raise RuntimeError("3-D return map did not converge")


def return_map_plane_stress(st, h, mu, lam, Y, S, D, a0):
    lp = 2 * mu * lam / (lam + 2 * mu)
    Mn = np.array([[h['G'] + h['H'], -h['H']], [-h['H'], h['F'] + h['H']]])
    CM = np.array([[2 * mu + lp, lp], [lp, 2 * mu + lp]]) @ Mn
    hv2 = lambda sn, t: np.sqrt(h['F'] * sn[1] ** 2 + h['G'] * sn[0] ** 2 + h['H'] * (sn[0] - sn[1]) ** 2 + 2 * h['N'] * t ** 2)
    sntr, ttr = np.array([st[0, 0], st[1, 1]]), st[0, 1]
    sy0, dsy = sigma_y(a0, Y, S, D)
    dgam = max((hv2(sntr, ttr) - sy0) / (3 * mu + dsy), 0.)
    polish = False
    for it in range(40):
        sy, dsy = sigma_y(a0 + dgam, Y, S, D)
        c = dgam / sy
        Ai = np.linalg.inv(np.eye(2) + c * CM)
        i01 = 1 / (1 + c * 2 * mu * h['N'])
        sn, t = Ai @ sntr, ttr * i01
        hv = hv2(sn, t); g = hv - sy
        if abs(g) < 1e-13 * mu:
            if polish:
                return sn, t, dgam, it
            polish = True
        dsn = -(Ai @ (CM @ sn))
        dh = ((Mn @ sn) @ dsn) / hv + 2 / hv * h['N'] * t * (-2 * mu * h['N'] * t * i01)
        dg = dh * ((sy - dgam * dsy) / sy ** 2) - dsy
        nd = dgam - g / dg
        dgam = nd if nd > 0 else .5 * dgam
    raise RuntimeError("plane-stress return map did not converge")


def main():
    rng = np.random.RandomState(0)
    E, nu = 1000., .25
    mu, lam = E / 2 / (1 + nu), E * nu / ((1 + nu) * (1 - 2 * nu))
    Y, S, D = 2., 10., 2.
    h = hill_params(1, .9, 1.1, 1, .95, 1.05)
    worst, its, n = 0., [], 0
    for _ in range(2000):
        e = rng.randn(3, 3) * rng.choice([1e-3, 5e-3, 3e-2]); e = .5 * (e + e.T)
        p0 = rng.randn(3, 3) * 1e-3; p0 = .5 * (p0 + p0.T); p0 -= np.trace(p0) / 3 * np.eye(3)
        a0 = abs(rng.randn()) * .01
        dev = e - np.trace(e) / 3 * np.eye(3)
        s_tr = 2 * mu * (dev - p0)
        if hill_value(s_tr, h) <= sigma_y(a0, Y, S, D)[0]:
            continue
        s, dgam, it = return_map_3d(s_tr, h, mu, Y, S, D, a0)
        hv = hill_value(s, h)
        p = p0 + dgam * hill_normal(s, h, hv)
        res = max(np.abs(2 * mu * (dev - p) - s).max() / mu, abs(hv - sigma_y(a0 + dgam, Y, S, D)[0]) / mu, abs(np.trace(p)))
        worst = max(worst, res); its.append(it); n += 1
    print(f"3-D: {n} yielding states, <= {max(its)} scalar steps (mean {np.mean(its):.1f}), worst residual {worst:.1e}")
    assert worst < 1e-14

    h2 = hill_params(1, .9, 1.1, .95, 1, 1)
    def cauchy(e, p):
        ezz = -(lam * np.trace(e) + 2 * mu * np.trace(p)) / (lam + 2 * mu)
        return 2 * mu * (e - p) + lam * (np.trace(e) + ezz) * np.eye(2)
    worst, its, n = 0., [], 0
    for _ in range(2000):
        e = rng.randn(2, 2) * rng.choice([1e-3, 5e-3, 3e-2]); e = .5 * (e + e.T)
        p0 = rng.randn(2, 2) * 1e-3; p0 = .5 * (p0 + p0.T); a0 = abs(rng.randn()) * .01
        st = cauchy(e, p0)
        s3 = np.zeros((3, 3)); s3[:2, :2] = st
        if hill_value(s3, h2) <= sigma_y(a0, Y, S, D)[0]:
            continue
        sn, t, dgam, it = return_map_plane_stress(st, h2, mu, lam, Y, S, D, a0)
        s3 = np.zeros((3, 3)); s3[0, 0], s3[1, 1] = sn; s3[0, 1] = s3[1, 0] = t
        hv = hill_value(s3, h2)
        p = p0 + dgam * hill_normal(s3, h2, hv)[:2, :2]
        res = max(np.abs(cauchy(e, p) - s3[:2, :2]).max() / mu, abs(hv - sigma_y(a0 + dgam, Y, S, D)[0]) / mu)
        worst = max(worst, res); its.append(it); n += 1
    print(f"plane stress: {n} yielding states, <= {max(its)} scalar steps (mean {np.mean(its):.1f}), worst residual {worst:.1e}")
    assert worst < 1e-14


if __name__ == "__main__":
    main()
