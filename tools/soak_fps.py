"""Soak of the scene-scale FPS routes: the bucketed form (alone and after a head of co-resident iterations) against the
co-resident kernel on random sizes / batch sizes / distributions, including degenerate geometry (planes, lines, duplicated
points, coordinates at 1e6 where fp32 is coarse, extents of 1e-21 where squares are denormal or underflow).  Every index must agree."""
import os, sys
import numpy as np, torch
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "point-cloud-compression_b200")]
from pcc_b200 import ops

KINDS = ["uniform", "plane", "line", "clusters", "dups", "offset", "tiny", "lattice", "aniso"]


def cloud(kind, n, rng):
    if kind == "uniform":
        p = rng.random((n, 3))
    elif kind == "plane":
        p = rng.random((n, 3)); p[:, 2] = 0.25
    elif kind == "line":
        t = rng.random(n); p = np.stack((t, 2 * t, -t), 1)
    elif kind == "clusters":
        c = rng.random((17, 3)) * 10; p = c[rng.integers(0, 17, n)] + rng.normal(0, 0.01, (n, 3))
    elif kind == "dups":
        base = rng.random((max(n // 7, 1), 3)); p = base[rng.integers(0, len(base), n)]
    elif kind == "offset":
        p = 1e6 + rng.random((n, 3)) * 3
    elif kind == "tiny":
        p = rng.random((n, 3)) * 1e-21
    elif kind == "lattice":
        p = (rng.integers(0, 40, (n, 3)) + 0.5) / 40
    else:
        p = rng.normal(0, 1, (n, 3)) * np.array([100.0, 1.0, 0.01])
    return p.astype(np.float32)


def main():
    rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
    bad = 0
    cases = [(1, 1_000_000, 3000, "clusters"), (1, 700_000, 2500, "plane"), (2, 300_000, 1300, "offset")]
    for _ in range(36):
        n = int(rng.choice([8193, 9000, 20_000, 65_536, 131_073, 250_000]))
        cases.append((int(rng.integers(1, 4)), n, int(rng.choice([1, 2, 130, 600, 1500])), str(rng.choice(KINDS))))
    for (B, N, S, kind) in cases:
        x = torch.from_numpy(np.stack([cloud(kind, N, rng) for _ in range(B)])).cuda()
        st = torch.from_numpy(rng.integers(0, N, B)).cuda()
        os.environ["PCC_FPS_PATH"] = "grid"; os.environ.pop("PCC_FPS_HEAD", None)
        ref = ops.fps(x, S, st, 1e10)
        ok = []
        for head in ("0", None, "3"):
            os.environ["PCC_FPS_PATH"] = "bucket"
            if head is None: os.environ.pop("PCC_FPS_HEAD", None)
            else: os.environ["PCC_FPS_HEAD"] = head
            ok.append(bool(torch.equal(ops.fps(x, S, st, 1e10), ref)))
        os.environ.pop("PCC_FPS_PATH", None); os.environ.pop("PCC_FPS_HEAD", None)
        ok.append(bool(torch.equal(ops.fps(x, S, st, 1e10), ref)))     # the default route
        bad += not all(ok)
        print(f"{kind:9s} B={B} N={N:8d} S={S:5d}: {'ok' if all(ok) else 'MISMATCH ' + str(ok)}", flush=True)
    print("soak:", "all identical" if bad == 0 else f"{bad} case(s) differ")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
