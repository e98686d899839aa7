"""The five BASELINE.json workloads: stateline config + built-in target + payload.

Each entry gives what both sides need: the engine's (config, plugin, payload) and the oracle's
(target id, target_params).  Pure host-side description, no compute."""
import numpy as np

# oracle target ids (oracle/sl_oracle.h)
T_DEMO_GAUSS, T_DOUBLE_WELL, T_GAUSS_FACT, T_GAUSS_CORR, T_ROSENBROCK = range(5)


def _householder_q(d, seed):
    """Orthogonal Q as a product of d seeded Householder reflectors (config #4)."""
    rng = np.random.default_rng(seed)
    Q = np.eye(d)
    for _ in range(d):
        v = rng.standard_normal(d)
        v /= np.linalg.norm(v)
        Q = Q - 2.0 * np.outer(Q @ v, v)
    return Q


def gauss_corr_U(d=100, seed=4):
    lam = np.logspace(0.0, 2.0, d)  # precision spectrum 1..100
    Q = _householder_q(d, seed)
    P = (Q * lam) @ Q.T
    P = 0.5 * (P + P.T)
    return np.ascontiguousarray(np.linalg.cholesky(P).T)  # upper: P = U^T U


def workload(name, S=None, T=None):
    """Returns dict(S, T, d, J, mn, mx, swapInterval, optAccept, optSwap, plugin, payload,
    oracle_target, oracle_params)."""
    if name in ("demo", "config1"):
        # src/bin/demo-config.json:1-14, likelihood src/bin/demo-worker.cpp:37-45
        w = dict(S=2, T=5, J=3, mn=[-10, 0, -10, -10], mx=[10, 10, 10, 2], plugin="demo_gauss", payload=None,
                 oracle_target=T_DEMO_GAUSS, oracle_params=None)
    elif name in ("double_well", "config2"):
        w = dict(S=64, T=8, J=2, mn=[-3, -3], mx=[3, 3], plugin="double_well", payload=None,
                 oracle_target=T_DOUBLE_WELL, oracle_params=None)
    elif name in ("gauss20", "config3"):
        d = 20
        mu = np.zeros(d)
        s = 2.0 ** ((np.arange(d) - 9.5) / 9.5)
        prm = np.concatenate([mu, 1.0 / s])
        w = dict(S=8192, T=8, J=10, mn=[-10.0] * d, mx=[10.0] * d, plugin="gauss_fact", payload=prm,
                 oracle_target=T_GAUSS_FACT, oracle_params=prm)
    elif name in ("gauss100", "config4"):
        d, J = 100, 4
        U = np.triu(gauss_corr_U(d, 4))
        w = dict(S=1024, T=8, J=J, mn=[-20.0] * d, mx=[20.0] * d, plugin="gauss_corr",
                 payload=np.concatenate([[float(J)], U.ravel(), U.T.ravel()]), oracle_target=T_GAUSS_CORR, oracle_params=U.ravel())
    elif name in ("rosenbrock50", "config5"):
        d, J = 50, 7
        w = dict(S=8192, T=16, J=J, mn=[-5.0] * d, mx=[5.0] * d, plugin="rosenbrock", payload=np.array([float(J)]),
                 oracle_target=T_ROSENBROCK, oracle_params=None)
    else:
        raise KeyError(name)
    w.update(swapInterval=10, optAccept=0.234, optSwap=0.3874, name=name)
    if S is not None:
        w["S"] = S
    if T is not None:
        w["T"] = T
    w["d"] = len(w["mn"])
    return w


def engine_config(w, **gpu):
    from .engine import make_config
    return make_config(w["S"], w["T"], w["mn"], w["mx"], w["J"], swapInterval=w["swapInterval"],
                       optimalAcceptRate=w["optAccept"], optimalSwapRate=w["optSwap"], **gpu)
