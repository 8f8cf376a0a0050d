"""Golden vectors for the simplex-noise row, made by RUNNING THE REFERENCE's own code
(/root/reference/src/utils/simplex_noise.py) in this container under the real numba JIT (numba 0.65 is installed; the
reference decorates its kernels with @njit(cache=True) / @njit(parallel=True)). matplotlib is absent and only imported
for the reference's plotting helpers: it is replaced by an inert stub.

    python tests/golden/make_simplex_golden.py              # real numba; writes tests/golden/simplex_golden.npz
    python tests/golden/make_simplex_golden.py --no-numba   # njit -> identity, prange -> range (the same Python source,
                                                            # interpreted); --check compares with the committed file

Both modes produce bit-identical arrays (checked when this file was regenerated in round 2): the arithmetic is plain
IEEE fp64 without fastmath, so the JIT changes speed, not results.

The GPU box has no /root/reference; the tests there use the committed .npz."""
import sys
import types
from pathlib import Path

import numpy as np


def import_reference(stub_numba: bool):
    if stub_numba:
        nb = types.ModuleType("numba")
        nb.njit = lambda *a, **k: (a[0] if a and callable(a[0]) else (lambda f: f))
        nb.prange = range
        sys.modules["numba"] = nb
    else:
        import os

        os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")  # /root/reference is read-only (cache=True)
        import numba  # noqa: F401
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].animation = sys.modules["matplotlib.animation"]
    sys.path.insert(0, "/root/reference")
    from src.utils import simplex_noise as ref

    return ref


def main():
    import torch

    stub = "--no-numba" in sys.argv
    ref = import_reference(stub)
    out = {}
    # 1. raw lattice function on random points, cell-boundary points and ties
    perm, pgi = ref._init(424242)
    rng = np.random.default_rng(1)
    pts = [tuple(rng.uniform(-50, 50, 3)) for _ in range(4000)]
    pts += [(a / 4, b / 4, c / 4) for a in range(-4, 5) for b in range(-4, 5) for c in range(-4, 5)]
    out["noise3_seed"] = np.int64(424242)
    out["noise3_points"] = np.array(pts)
    out["noise3_values"] = np.array([ref._noise3(x, y, z, perm, pgi) for x, y, z in pts])
    # 2. permutation tables
    seeds = np.array([3, 424242, -9876543210, 9999999999, -1], dtype=np.int64)
    out["table_seeds"] = seeds
    out["table_perm"] = np.stack([ref._init(int(s))[0] for s in seeds])
    out["table_grad_index3"] = np.stack([ref._init(int(s))[1] for s in seeds])
    # 3. the call the trainer makes (src/trainers/reconstruct.py:133-139), seeds from numpy's global RNG
    for name, shape, ts in (("a", (2, 1, 16, 16), [10, 650]), ("b", (1, 3, 12, 20), [330])):
        np.random.seed(7)
        x = torch.zeros(shape)
        t = torch.tensor(ts).long()
        noise = ref.generate_simplex_noise(ref.Simplex_CLASS(), x, t, in_channels=shape[1])
        out[f"gen_{name}_shape"] = np.array(shape)
        out[f"gen_{name}_t"] = np.array(ts)
        out[f"gen_{name}_noise"] = noise.numpy()
    path = Path(__file__).parent / "simplex_golden.npz"
    if "--check" in sys.argv:
        old = np.load(path)
        same = all(np.array_equal(old[k], np.asarray(v)) for k, v in out.items() if k != "numba")
        print("bit-identical to the committed file:", same)
        sys.exit(0 if same else 1)
    import numba as _nb

    out["numba"] = np.array("stubbed" if stub else getattr(_nb, "__version__", "?"))
    np.savez_compressed(path, **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
