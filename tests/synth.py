"""Seeded synthetic tables shared by the tests (numpy PCG64; SURVEY §8d shapes scaled down)."""
import numpy as np


def rank_tables(n_items=5000, n_fields=32, fdim=16, seed=4, max_rows=3000):
    rng = np.random.default_rng(seed)
    rows_t = rng.integers(50, max_rows, size=n_fields)
    factors = [(rng.standard_normal((int(r), fdim)) * 0.05).astype(np.float32) for r in rows_t]
    linear = [(rng.standard_normal(int(r)) * 0.05).astype(np.float32) for r in rows_t]
    fields = np.stack([rng.integers(0, int(r), size=n_items) for r in rows_t], axis=1).astype(np.uint32)
    return fields, factors, linear


def mlp_weights(dims, seed=6):
    """Xavier-uniform bf16 weights [out][in] and small f32 biases."""
    import oracle
    rng = np.random.default_rng(seed)
    W, b = [], []
    for l in range(len(dims) - 1):
        fan_in, fan_out = dims[l], dims[l + 1]
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        w = rng.uniform(-lim, lim, size=(fan_out, fan_in)).astype(np.float32)
        W.append(oracle.f32_to_bf16(w))
        b.append((rng.standard_normal(fan_out) * 0.01).astype(np.float32))
    return W, b


def diversity(n_items=5000, dim=128, seed=7, dtype=np.float32):
    rng = np.random.default_rng(seed)
    D = rng.standard_normal((n_items, dim))
    D /= np.linalg.norm(D, axis=1, keepdims=True)
    return D.astype(dtype)
