"""Design check for the int8 filter index worked out in DESIGN.md §7 (no kernel exists yet): quantise rows and
threshold-normalised queries, and verify on random and adversarial data that EVERY row whose exact fp32 score reaches
the threshold passes the integer test   acc >= 1/(t s_r) - 1/2 |x8_r|_1 - max_q(1/2 |q8_q|_1 + d/4),
and report how many extra rows pass compared with the bf16 filter's margin."""
import numpy as np


def run(n=400_000, d=64, nq=16, seed=0, scale_rows=False):
    rng = np.random.default_rng(seed)
    X = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    if scale_rows:
        X *= np.exp(rng.uniform(-3, 3, size=(n, 1))).astype(np.float32)   # row norms over 2.5 decades
    Q = (rng.standard_normal((nq, d)) / np.sqrt(d)).astype(np.float32)
    exact = X.astype(np.float64) @ Q.astype(np.float64).T                   # [n, nq]
    k = 400
    tau = np.sort(exact, axis=0)[-4 * k]                                   # ~4k rows reach it
    assert (tau > 0).all()
    Qn = Q.astype(np.float64) / tau[:, None]                               # exact' >= 1  <=>  exact >= tau
    s = np.abs(X).max(axis=1).astype(np.float64) / 127.0                   # per-row scale
    s[s == 0] = 1.0
    x8 = np.rint(X / s[:, None]).astype(np.int32)
    t = np.abs(Qn).max() / 127.0                                           # ONE scale for the pass
    q8 = np.rint(Qn / t).astype(np.int32)
    acc = x8 @ q8.T                                                        # exact in s32
    B = (0.5 * np.abs(q8).sum(axis=1) + 0.25 * d).max()
    T = 1.0 / (t * s) - 0.5 * np.abs(x8).sum(axis=1) - B                   # per-row scalar
    keep = acc >= T[:, None]
    must = exact >= tau[None, :]
    missed = int((must & ~keep).sum())
    # bf16 filter for comparison: approx >= tau - c |x||q|
    c = 1.05 / 256
    xb = (X.view(np.uint32) + 0x8000 & 0xFFFF0000).view(np.float32)        # round-to-nearest-ish bf16
    qb = (Q.view(np.uint32) + 0x8000 & 0xFFFF0000).view(np.float32)
    approx = xb.astype(np.float64) @ qb.astype(np.float64).T
    keep_b = approx >= tau[None, :] - c * np.linalg.norm(X, axis=1)[:, None] * np.linalg.norm(Q, axis=1)[None, :]
    return missed, keep.sum(axis=0).mean(), keep_b.sum(axis=0).mean(), must.sum(axis=0).mean()


if __name__ == "__main__":
    for kw in (dict(), dict(seed=1, scale_rows=True), dict(d=128, seed=2)):
        missed, k8, kb, m = run(**kw)
        print(kw, f"rows that must pass: {m:.0f}/query; pass int8: {k8:.0f} ({k8 / m:.2f}x); pass bf16: {kb:.0f} "
                  f"({kb / m:.2f}x); missed by int8: {missed}")
        assert missed == 0
