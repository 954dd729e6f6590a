"""The Horn solver of the product (rcvpose_b200/csrc/horn_core.h: largest root of the quartic by Newton + adjugate
eigenvector + inverse-iteration polish), compiled for the host, against the REAL reference's lmshorn outputs
(tests/golden/reference_golden.npz, util/horn.py:75-181) and against the oracle's restatement on random triples."""
import numpy as np
import pytest

from oracle import oracle
from tests import hostsim


@pytest.mark.parametrize("t", range(6))
def test_horn_core_matches_reference_golden(golden, t):
    P1, P2, want = golden["h%d_P1" % t], golden["h%d_P2" % t], golden["h%d_RT" % t]
    got = hostsim.horn(P1, P2)
    scale = max(1.0, float(np.abs(want[:3, 3]).max()))
    assert np.abs(got[:3, :3] - want[:3, :3]).max() < 1e-12
    assert np.abs(got[:3, 3] - want[:3, 3]).max() < 1e-12 * scale * 10
    assert np.array_equal(got[3], [0, 0, 0, 1])


def _rand_rot(rng):
    a = rng.normal(size=3); th = np.linalg.norm(a); k = a / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx


def test_horn_core_random_triples_vs_oracle_and_kabsch():
    rng = np.random.default_rng(3)
    worst = 0.0
    for _ in range(500):
        n = int(rng.choice([3, 3, 3, 4, 9]))
        P1 = rng.normal(size=(n, 3)) * rng.uniform(10, 200)
        R = _rand_rot(rng)
        P2 = P1 @ R.T + rng.normal(size=3) * 300 + rng.normal(size=(n, 3)) * rng.choice([0, 0.1, 2.5, 10])
        got = hostsim.horn(P1, P2)
        want = np.zeros((4, 4)); oracle.lmshorn(P1, P2, n, want)
        a, b = P1 - P1.mean(0), P2 - P2.mean(0)
        U, s, Vt = np.linalg.svd((a.T @ b).T)
        Rk = U @ np.diag([1, 1, np.sign(np.linalg.det(U @ Vt))]) @ Vt
        # ill-conditioned triples (nearly collinear): every solver agrees only to eps / gap; bound by Kabsch's own distance
        tol = max(1e-11, 20 * np.abs(Rk - want[:3, :3]).max())
        err = np.abs(got[:3, :3] - want[:3, :3]).max()
        assert err < tol, (err, tol)
        assert abs(np.linalg.det(got[:3, :3]) - 1) < 1e-12 and np.abs(got[:3, :3] @ got[:3, :3].T - np.eye(3)).max() < 1e-12
        worst = max(worst, err)
    assert worst < 1e-9


def test_horn_core_degenerate_inputs_give_a_rotation():
    rng = np.random.default_rng(4)
    P = rng.normal(size=(3, 3))
    cases = [(np.zeros((3, 3)), np.zeros((3, 3))),                                   # S = 0: identity
             (np.outer([0, 1, 2], [1.0, 2, 3]), np.outer([0, 1, 2], [3.0, -1, 2])),     # collinear both: rotation about the line is free
             (P, P),                                                                  # identity motion
             (P, -P)]                                                                 # reflection-like: best proper rotation
    for P1, P2 in cases:
        got = hostsim.horn(P1, P2)
        R = got[:3, :3]
        assert np.isfinite(got).all()
        assert np.abs(R @ R.T - np.eye(3)).max() < 1e-12 and abs(np.linalg.det(R) - 1) < 1e-12
    assert np.allclose(hostsim.horn(np.zeros((3, 3)), np.zeros((3, 3)))[:3, :3], np.eye(3))
    assert np.allclose(hostsim.horn(P, P)[:3, :3], np.eye(3), atol=1e-14)
    # collinear: the aligned line must map onto the target line
    P1, P2 = cases[1]
    got = hostsim.horn(P1, P2)
    a, b = P1 - P1.mean(0), P2 - P2.mean(0)
    assert np.allclose(a @ got[:3, :3].T / np.linalg.norm(a[2]), b / np.linalg.norm(b[2]), atol=1e-10)
